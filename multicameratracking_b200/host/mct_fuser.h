// mct_fuser.h -- host side of the geometric fuser: the reference's
// GeometryBasedInformationFuser in its only live mode, PRINCIPAL_AXIS_INTERSECTION (CameraNetwork.cpp:398).
//
// The reference builds this on OpenCV's C API (cvInv, cvMatMul, cvSVD, CvKalman); OpenCV is not a dependency here, so
// the few small-matrix operations are written out: 3x3 f64 inverse by cofactors, the null direction of the stacked
// ground-plane lines by a one-sided Jacobi SVD, and the 4-state / 2-measurement Kalman filter on f32 matrices with f64
// accumulation (the arithmetic of OpenCV's small GEMM).  The per-particle work (ground-plane pdf, rearrangement,
// re-scoring) runs on the device: mct_pf_ground_pdf, mct_pf_rearrange_ground, mct_track_score(noise = NULL).
#pragma once
#include <array>
#include <vector>

#include "mct_host.h"

namespace MultipleCameraTracking {
typedef std::array<double, 9> Homography;     // row-major 3x3, image plane -> ground plane (Camera.cpp:120-132)

class GeometryBasedInformationFuser {
public:
    // GeometryBasedInformationFuser.cpp:30-75: one fuser per object, the homographies of all cameras
    explicit GeometryBasedInformationFuser(const std::vector<Homography>& homographyMatrixList);
    // :316-413; particles = [numberOfCameras][2] foot points (image plane) of the highest-weight particle per camera
    void InitializeKalman(const float* particles);
    // :443-464 -> FuseInformationByUsingPrincipalAxisIntersection (:630-664): measurement = intersection of the
    // cameras' principal axes on the ground plane, measurement covariance diag(5, 5), Kalman predict + correct
    void FuseInformation(const float* particles);
    void GetGroundPlaneKalmanMeanMatrix(float mean[2]) const;              // :672-688
    void GetGroundPlaneKalmanCovarianceMatrix(float cov[4]) const;        // :696-712
    const float* StatePost() const { return m_statePost; }
    const float* ErrorCovPost() const { return m_errorCovPost; }
    bool IsInitialized() const { return m_initialized; }
    int NumberOfCameras() const { return (int)m_homographyMatrixList.size(); }

    // :220-284
    static std::array<double, 2> EstimateGroundPlanePrincipleAxisIntersection(const float* pixels, const std::vector<Homography>& homographyMatrixList);
    // :172-212 (f32 points in, f64 out)
    static void TransformWithHomography(const float* particles, int n, const Homography& H, double* out);
    static Homography Invert(const Homography& H);
private:
    void UpdateKalmanFilter(const float measurement[2], const float* measurementCov);   // :420-436
    std::vector<Homography> m_homographyMatrixList;
    // CvKalman(4, 2, 0), row-major f32
    float m_transition[16], m_measurementMatrix[8], m_processNoiseCov[16], m_measurementNoiseCov[4];
    float m_statePre[4], m_statePost[4], m_errorCovPre[16], m_errorCovPost[16];
    bool m_initialized;
};

// ParticleFilterTracker::UpdateParticlesWithGroundPDF (ParticleFilterTracker.cpp:1040-1125) for one object of one camera.
// Returns true when the particles were rearranged and re-scored.  The randgaus(0, 10) pairs are drawn from the tracker's
// MWC stream in the reference's order (Public.cpp:37-49).
bool UpdateParticlesWithGroundPDF(mct_ctx* ctx, ParticleFilterTracker* tracker, const float mean[2], const float cov[4],
                                  const Homography& homography, float* distanceOut = nullptr);
// The same for all objects of one camera (the object loop of FeedbackInformationFromGeometricFusion, CameraNetwork.cpp:105-121):
// ONE device pass + read-back for the ground pdfs of all objects, then rearrangement + re-score for the objects that need it.
// means [n][2], covs [n][4]; rearranged / distances (optional) [n].
void UpdateParticlesWithGroundPDF(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, const float* means, const float* covs,
                                  const Homography& homography, int* rearranged = nullptr, float* distances = nullptr);
}  // namespace MultipleCameraTracking

namespace Public {
float randgaus(const float mean, const float sigma);    // Public.cpp:37-49 (polar Box-Muller on randfloat())
}
