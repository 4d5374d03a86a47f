// mct_network.h -- the cross-camera schedule: CameraNetwork (CameraNetwork.{h,cpp}) with both fusers, for the B200 path.
//
//   reference (src/)                                              here
//   ------------------------------------------------------------  ----------------------------------------------------
//   CameraNetwork::TrackObjectsOnCurrentFrame (:537-613)          CameraNetwork::TrackObjectsOnCurrentFrame
//   Camera::TrackCameraFrame / Object::TrackObjectFrame           batched per camera: one score pass for all objects,
//     (Camera.cpp:334-397, Object.cpp:411-455)                    one re-weighting pass with the global classifiers
//   CameraNetwork::FuseGeometrically / Feedback... (:94-268)      FuseGeometrically: packed foot points -> exchange ->
//                                                                 per-object fuser on every process -> batched feedback
//   Camera::LearnGlobalAppearanceModel, AppearanceBasedInfor-     LearnGlobalAppearanceModel: per learner camera one
//     mationFuser::LearnGlobalAppearanceModel (:44-90),           round = training sets regenerated on every camera
//     CameraNetwork::GenerateTrainingSampleSetsFor...(:276-320)   (RNG order kept), FEATURE ROWS exchanged instead of
//                                                                 image pointers, drop rule + Update at the learner
//
// The reference iterates cameras inside one process and reads the other cameras through pointers.  Here the cameras of a
// network are spread over processes (one camera per process per GPU) or all live in this process (tests, single-GPU use);
// everything that crosses cameras goes through NetworkExchange.  Every (camera, object) pair owns its MWC stream
// (the reference has one global stream, which would serialise the cameras).
#pragma once
#include <array>
#include <memory>
#include <vector>

#include "mct_fuser.h"
#include "mct_host.h"

struct mcth_camera;

namespace MultipleCameraTracking {

// TrackerParameters.h:10-30 + config.cfg "Fusion setting"
struct NetworkParameters {
    int geometricFusionType = 0;             // Geometric_Fusion_Type: 0 none, 1 ground plane (principal-axis mode, CameraNetwork.cpp:398)
    int appearanceFusionType = 0;            // Appearance_Fusion_Type: 0 none, 1 culture colour, 2 multi-dimensional colour
    int appearanceFusionStrongClassifierType = 1;   // 1 MILBoost, 2 AdaBoost, 3 MILEnsemble (TrackerParameters.h:24-29)
    int appearanceFusionWeakClassifierType = 0;     // Classifier::WeakClassifierType
    int afPercentageOfWeakClassifiersSelected = 20;
    int afPercentageOfWeakClassifiersRetained = 0;
    int appearanceFusionRefreshRate = 1;
};

// What crosses processes.  Rank r holds camera r (one camera per process); a network whose cameras all live in this process
// needs no exchange object.
class NetworkExchange {
public:
    virtual ~NetworkExchange() {}
    virtual int Rank() const = 0;
    virtual int World() const = 0;
    // recv = [World][bytes] in rank order; host memory
    virtual void AllGatherHost(const void* send, void* recv, size_t bytes) = 0;
    // all-to-all of the device row buffers registered with CameraNetwork::SetRowBuffers, both [World][slotBytes]:
    // recv slot c = camera c's send slot Rank().  Must be ordered after the work already enqueued on `stream` and complete
    // (or stream-ordered) when it returns
    virtual void AllToAllRows(size_t slotBytes, void* stream) = 0;
};

class CameraNetwork {
public:
    // localCameras: the cameras of this process, trackers added (not yet initialised); cameraIds[i] = global index of
    // localCameras[i]; homographies: image -> ground plane, one per camera of the whole network
    CameraNetwork(const std::vector<mcth_camera*>& localCameras, const std::vector<int>& cameraIds, int numberOfCameras,
                  const std::vector<Homography>& homographies, const NetworkParameters& params, NetworkExchange* exchange = nullptr);
    ~CameraNetwork();
    // CameraNetwork ctor (:26-38): Camera::InitializeCameraTrackers for every camera (frames already set), the appearance
    // fusers' global classifiers (Object.cpp:203-216), InitializeGeometricFuser (:380-470)
    void InitializeCameraNetwork();
    // :537-613.  noise[i]: prediction noise of local camera i, host [n_obj][4][N].  The cameras' frames are already set.
    void TrackObjectsOnCurrentFrame(int frameIndex, const float* const* noise);
    // multi-process runs with a collective library: the embedding owns the exchange buffers (e.g. torch tensors for NCCL)
    size_t RowBufferBytes() const { return m_rowBytes * (size_t)m_numberOfCameras; }     // of EACH of the two buffers
    void SetRowBuffers(void* sendDev, void* recvDev);

    // introspection (parity tests)
    int NumberOfObjects() const { return m_numberOfObjects; }
    const GeometryBasedInformationFuser* GeometricFuser(int object) const { return m_geometricFusers.empty() ? nullptr : m_geometricFusers[object].get(); }
    Classifier::StrongClassifierBasePtr GlobalClassifier(int localCamera, int object) const { return m_globalClassifiers[localCamera][object]; }
    int LastRearranged(int localCamera, int object) const { return m_lastRearranged[localCamera][object]; }
    // host milliseconds spent in the stages of the last frame: score, reweight, geometric (exchange / fusers / feedback), update, learn, store
    const double* StageMilliseconds() const { return m_stageMs; }
private:
    void FuseGeometrically(int frameIndex, bool initialise);
    void UpdateParticleWeightsUsingMultiCameraAppearanceModel(int localCamera);
    void ForceParticleResampling(int localCamera);
    void LearnGlobalAppearanceModel();
    void GatherFootPoints(std::vector<float>& all);

    std::vector<mcth_camera*> m_cameras;
    std::vector<int> m_cameraIds;
    int m_numberOfCameras, m_numberOfObjects;
    std::vector<Homography> m_homographies;
    NetworkParameters m_params;
    NetworkExchange* m_exchange;
    std::vector<std::unique_ptr<GeometryBasedInformationFuser>> m_geometricFusers;           // per object, replicated on every process
    std::vector<std::vector<Classifier::StrongClassifierBasePtr>> m_globalClassifiers;        // [local camera][object]
    std::vector<std::vector<char>> m_globalTrained;
    std::vector<std::vector<int>> m_lastRearranged;
    // appearance-fuser exchange buffers, per local camera [numberOfCameras slots][object][F][ldRow] feature rows (positives from
    // column 0, negatives from column ldPos).  Send slot L = this camera's rows of learner round L; receive slot c = camera c's
    // rows of the round in which THIS camera learns.  m_rowBytes = bytes of one slot.
    int m_afFeatures, m_ldPos, m_ldRow;
    size_t m_rowBytes;
    std::vector<void*> m_sendRows, m_recvRows;             // per local camera (device)
    bool m_ownRowBuffers;
    double m_stageMs[8];
};
}  // namespace MultipleCameraTracking
