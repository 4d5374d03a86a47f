// mct_fuser.cpp -- see mct_fuser.h
#include "mct_fuser.h"

#include <algorithm>
#include <cmath>
#include <thread>
#include <cstring>

namespace Public {
float randgaus(const float mean, const float sigma)
{
    double x, y, r2;
    do {
        x = -1 + 2 * randfloat();
        y = -1 + 2 * randfloat();
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return (float)(sigma * y * std::sqrt(-2.0 * std::log(r2) / r2)) + mean;
}
}  // namespace Public

namespace MultipleCameraTracking {
namespace {
// C[m x n] = A[m x k] * B[k x n] (+ D), f32 storage, f64 accumulation
void gemm32(const float* A, const float* B, const float* D, float* C, int m, int k, int n, bool transB = false)
{
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int q = 0; q < k; q++) s += (double)A[i * k + q] * (double)(transB ? B[j * k + q] : B[q * n + j]);
            if (D) s += (double)D[i * n + j];
            C[i * n + j] = (float)s;
        }
}
}  // namespace

Homography GeometryBasedInformationFuser::Invert(const Homography& H)
{
    Homography t{};
    const double d = H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) + H[2] * (H[3] * H[7] - H[4] * H[6]);
    if (d == 0.0) return t;
    const double r = 1.0 / d;
    t[0] = (H[4] * H[8] - H[5] * H[7]) * r; t[1] = (H[2] * H[7] - H[1] * H[8]) * r; t[2] = (H[1] * H[5] - H[2] * H[4]) * r;
    t[3] = (H[5] * H[6] - H[3] * H[8]) * r; t[4] = (H[0] * H[8] - H[2] * H[6]) * r; t[5] = (H[2] * H[3] - H[0] * H[5]) * r;
    t[6] = (H[3] * H[7] - H[4] * H[6]) * r; t[7] = (H[1] * H[6] - H[0] * H[7]) * r; t[8] = (H[0] * H[4] - H[1] * H[3]) * r;
    return t;
}

void GeometryBasedInformationFuser::TransformWithHomography(const float* particles, int n, const Homography& H, double* out)
{
    for (int i = 0; i < n; i++) {
        const double x = particles[2 * i], y = particles[2 * i + 1];
        const double t0 = H[0] * x + H[1] * y + H[2], t1 = H[3] * x + H[4] * y + H[5], t2 = H[6] * x + H[7] * y + H[8];
        out[2 * i] = t0 / t2; out[2 * i + 1] = t1 / t2;
    }
}

std::array<double, 2> GeometryBasedInformationFuser::EstimateGroundPlanePrincipleAxisIntersection(const float* pixels,
                                                                                                 const std::vector<Homography>& list)
{
    const int C = (int)list.size();
    if (C < 2) throw MctHostError(MCT_E_ARG, "principal-axis intersection needs at least two cameras");
    // the vertical image line x = foot_x of every camera, mapped to the ground plane: [1, 0, -x] * inv(H_c)
    const int R = C < 3 ? 3 : C;
    std::vector<double> A((size_t)R * 3, 0.0);
    for (int c = 0; c < C; c++) {
        const Homography Hi = Invert(list[c]);
        const double l2 = -1.0 * (double)pixels[2 * c];
        for (int j = 0; j < 3; j++) A[(size_t)c * 3 + j] = 1.0 * Hi[j] + 0.0 * Hi[3 + j] + l2 * Hi[6 + j];
    }
    // one-sided Jacobi SVD: rotate column pairs of A until they are orthogonal, accumulating V; the right singular vector
    // of the smallest singular value (cvSVD's V column 2) is the column that ends with the smallest norm
    double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double a = 0, b = 0, g = 0;
                for (int r = 0; r < R; r++) { a += A[r * 3 + p] * A[r * 3 + p]; b += A[r * 3 + q] * A[r * 3 + q]; g += A[r * 3 + p] * A[r * 3 + q]; }
                if (g == 0.0 || std::fabs(g) <= 1e-300) continue;
                off = std::fmax(off, std::fabs(g) / std::sqrt(a * b + 1e-300));
                const double zeta = (b - a) / (2.0 * g);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int r = 0; r < R; r++) {
                    const double x = A[r * 3 + p], y = A[r * 3 + q];
                    A[r * 3 + p] = cs * x - sn * y; A[r * 3 + q] = sn * x + cs * y;
                }
                for (int r = 0; r < 3; r++) {
                    const double x = V[r * 3 + p], y = V[r * 3 + q];
                    V[r * 3 + p] = cs * x - sn * y; V[r * 3 + q] = sn * x + cs * y;
                }
            }
        if (off < 1e-15) break;
    }
    int best = 0; double bn = -1;
    for (int j = 0; j < 3; j++) {
        double n2 = 0; for (int r = 0; r < R; r++) n2 += A[r * 3 + j] * A[r * 3 + j];
        if (bn < 0 || n2 < bn) { bn = n2; best = j; }
    }
    return {V[0 * 3 + best] / V[2 * 3 + best], V[1 * 3 + best] / V[2 * 3 + best]};
}

GeometryBasedInformationFuser::GeometryBasedInformationFuser(const std::vector<Homography>& list)
    : m_homographyMatrixList(list), m_initialized(false)
{
    if (list.size() < 2) throw MctHostError(MCT_E_ARG, "geometric fusion needs at least two cameras");
    std::memset(m_statePre, 0, sizeof(m_statePre)); std::memset(m_statePost, 0, sizeof(m_statePost));
    std::memset(m_errorCovPre, 0, sizeof(m_errorCovPre)); std::memset(m_errorCovPost, 0, sizeof(m_errorCovPost));
}

void GeometryBasedInformationFuser::InitializeKalman(const float* particles)
{
    static const float A[16] = {1, 0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1};      // linear motion
    std::memcpy(m_transition, A, sizeof(A));
    std::memset(m_measurementMatrix, 0, sizeof(m_measurementMatrix));
    m_measurementMatrix[0] = 1; m_measurementMatrix[5] = 1;                               // cvSetIdentity on the 2x4 matrix
    for (int i = 0; i < 16; i++) { m_processNoiseCov[i] = (i % 5 == 0) ? 1.0f : 0.0f; m_errorCovPost[i] = (i % 5 == 0) ? 100.0f : 0.0f; }
    m_measurementNoiseCov[0] = 10; m_measurementNoiseCov[1] = 0; m_measurementNoiseCov[2] = 0; m_measurementNoiseCov[3] = 10;
    const std::array<double, 2> m = EstimateGroundPlanePrincipleAxisIntersection(particles, m_homographyMatrixList);
    m_statePost[0] = (float)m[0]; m_statePost[1] = (float)m[1]; m_statePost[2] = 0.0f; m_statePost[3] = 0.0f;
    m_initialized = true;
}

void GeometryBasedInformationFuser::UpdateKalmanFilter(const float z[2], const float* measurementCov)
{
    if (measurementCov) std::memcpy(m_measurementNoiseCov, measurementCov, sizeof(m_measurementNoiseCov));
    // cvKalmanPredict: x' = A x, P' = A P A^T + Q (and both copied to the posteriors)
    float t1[16];
    gemm32(m_transition, m_statePost, nullptr, m_statePre, 4, 4, 1);
    gemm32(m_transition, m_errorCovPost, nullptr, t1, 4, 4, 4);
    gemm32(t1, m_transition, m_processNoiseCov, m_errorCovPre, 4, 4, 4, true);
    std::memcpy(m_statePost, m_statePre, sizeof(m_statePre));
    std::memcpy(m_errorCovPost, m_errorCovPre, sizeof(m_errorCovPre));
    // cvKalmanCorrect: t2 = H P', t3 = t2 H^T + R, K^T = t3^-1 t2, x = x' + K (z - H x'), P = P' - K t2
    float t2[8], t3[4], t4[8], K[8], t5[2];
    gemm32(m_measurementMatrix, m_errorCovPre, nullptr, t2, 2, 4, 4);
    gemm32(t2, m_measurementMatrix, m_measurementNoiseCov, t3, 2, 4, 2, true);
    const double det = (double)t3[0] * t3[3] - (double)t3[1] * t3[2];
    const double i00 = t3[3] / det, i01 = -t3[1] / det, i10 = -t3[2] / det, i11 = t3[0] / det;
    for (int j = 0; j < 4; j++) {
        t4[j] = (float)(i00 * t2[j] + i01 * t2[4 + j]);
        t4[4 + j] = (float)(i10 * t2[j] + i11 * t2[4 + j]);
    }
    for (int i = 0; i < 4; i++) { K[i * 2] = t4[i]; K[i * 2 + 1] = t4[4 + i]; }
    for (int i = 0; i < 2; i++) {
        double s = 0; for (int q = 0; q < 4; q++) s += (double)m_measurementMatrix[i * 4 + q] * (double)m_statePre[q];
        t5[i] = (float)((double)z[i] - s);
    }
    gemm32(K, t5, m_statePre, m_statePost, 4, 2, 1);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0; for (int q = 0; q < 2; q++) s += (double)K[i * 2 + q] * (double)t2[q * 4 + j];
            m_errorCovPost[i * 4 + j] = (float)((double)m_errorCovPre[i * 4 + j] - s);
        }
}

void GeometryBasedInformationFuser::FuseInformation(const float* particles)
{
    if (!m_initialized) throw MctHostError(MCT_E_STATE, "GeometryBasedInformationFuser: InitializeKalman has not run");
    const std::array<double, 2> g = EstimateGroundPlanePrincipleAxisIntersection(particles, m_homographyMatrixList);
    const float position[2] = {(float)g[0], (float)g[1]};
    const float covArray[4] = {5, 0, 0, 5};                  // constant measurement covariance (:652-653)
    UpdateKalmanFilter(position, covArray);
}

void GeometryBasedInformationFuser::GetGroundPlaneKalmanMeanMatrix(float mean[2]) const { mean[0] = m_statePost[0]; mean[1] = m_statePost[1]; }
void GeometryBasedInformationFuser::GetGroundPlaneKalmanCovarianceMatrix(float cov[4]) const
{
    cov[0] = m_errorCovPost[0]; cov[1] = m_errorCovPost[1]; cov[2] = m_errorCovPost[4]; cov[3] = m_errorCovPost[5];
}

// the part of UpdateParticlesWithGroundPDF behind the pdf evaluation (:1088-1123)
static bool feedback_after_pdf(mct_ctx* ctx, ParticleFilterTracker* t, const mct_ground_pdf_result& res, const float mean[2],
                               const Homography& H, float* distanceOut)
{
    const vectorf& cur = t->GetCurrentTrackerState();
    mct_pf* pf = t->GetParticleFilter()->Handle();
    if (distanceOut) *distanceOut = -1.0f;
    if (!(res.total_weight > 0.0)) return false;                                   // :1088
    // FindMeanFootPositionOnImagePlane (:1168-1194): the Kalman mean back through the inverse homography
    double img[2];
    GeometryBasedInformationFuser::TransformWithHomography(mean, 1, GeometryBasedInformationFuser::Invert(H), img);
    const float meanFoot[2] = {(float)img[0], (float)img[1]};
    // :1093-1099 (the average particle's foot uses cy + h0 * sy, without the 1/2 the other foot points carry)
    const float footX = res.average[0], footY = res.average[1] + cur[3] * res.average[3];
    const double distance = std::sqrt(std::pow((double)(meanFoot[0] - footX), 2) + std::pow((double)(footY - meanFoot[1]), 2));
    if (distanceOut) *distanceOut = (float)distance;
    if (distance < 20) return false;                                               // :1100-1105
    // RearrangeParticlesBasedOnGroundLocation: (xRand, yRand) = randgaus(0, 10) per particle, in particle order
    const int N = t->GetParticleFilter()->GetNumberOfParticles();
    std::vector<float> nz((size_t)2 * N);
    Public::SetCurrentRng(&t->Rng());
    for (int r = 0; r < N; r++) { nz[r] = Public::randgaus(0, 10.0f); nz[(size_t)N + r] = Public::randgaus(0, 10.0f); }
    int rc = mct_pf_rearrange_ground(pf, meanFoot[0], meanFoot[1], nz.data(), cur[2], cur[3]);
    if (rc) throw MctHostError(rc, mct_last_error(ctx));
    // UpdateParticleWeightsForRearrangedParticles (:1334-1482): score in place, exponent 1, weights reset, resample
    mct_clf* clf = t->GetClassifier()->Handle();
    mct_track_params tp; tp.w0 = cur[2]; tp.h0 = cur[3]; tp.exponent = 1; tp.force_reset = 1; tp.resample = 1;
    const float u01 = (float)(t->Rng().Peek() * 2.3283064365386962890625e-10);
    mct_pf_summary summ;
    rc = mct_track_score(ctx, 1, &pf, &clf, &tp, nullptr, 0, &u01, &summ);
    if (rc) throw MctHostError(rc, mct_last_error(ctx));
    if (summ.resampled) (void)t->Rng().Next();
    t->GetParticleFilter()->AdoptSummary(summ);
    return true;
}

bool UpdateParticlesWithGroundPDF(mct_ctx* ctx, ParticleFilterTracker* t, const float mean[2], const float cov[4],
                                  const Homography& H, float* distanceOut)
{
    const vectorf& cur = t->GetCurrentTrackerState();
    mct_ground_pdf_result res;
    int rc = mct_pf_ground_pdf(t->GetParticleFilter()->Handle(), H.data(), mean, cov, cur[3], &res, nullptr);
    if (rc) throw MctHostError(rc, mct_last_error(ctx));
    return feedback_after_pdf(ctx, t, res, mean, H, distanceOut);
}

void UpdateParticlesWithGroundPDF(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, const float* means, const float* covs,
                                  const Homography& H, int* rearranged, float* distances)
{
    const int n = (int)trackers.size();
    if (n == 0) return;
    std::vector<mct_pf*> pfs(n); std::vector<float> h0(n);
    std::vector<mct_ground_pdf_result> res(n);
    for (int o = 0; o < n; o++) { pfs[o] = trackers[o]->GetParticleFilter()->Handle(); h0[o] = trackers[o]->GetCurrentTrackerState()[3]; }
    int rc = mct_fuser_ground_pdf(ctx, n, pfs.data(), H.data(), means, covs, h0.data(), res.data());
    if (rc) throw MctHostError(rc, mct_last_error(ctx));
    // the host part of :1088-1105 per object; the objects that rearrange then go through ONE batched rearrangement pass and ONE
    // batched re-scoring call (every object draws from its own MWC stream, so the order across objects is free)
    std::vector<int> R;
    std::vector<float> meanFoot, nz, w0v, h0v;
    std::vector<size_t> nzAt;
    const Homography Hinv = GeometryBasedInformationFuser::Invert(H);
    for (int o = 0; o < n; o++) {
        ParticleFilterTracker* t = trackers[o];
        const vectorf& cur = t->GetCurrentTrackerState();
        if (rearranged) rearranged[o] = 0;
        if (distances) distances[o] = -1.0f;
        if (!(res[o].total_weight > 0.0)) continue;
        double img[2];
        GeometryBasedInformationFuser::TransformWithHomography(means + 2 * o, 1, Hinv, img);
        const float mf[2] = {(float)img[0], (float)img[1]};
        const float footX = res[o].average[0], footY = res[o].average[1] + cur[3] * res[o].average[3];
        const double distance = std::sqrt(std::pow((double)(mf[0] - footX), 2) + std::pow((double)(footY - mf[1]), 2));
        if (distances) distances[o] = (float)distance;
        if (distance < 20) continue;
        const int N = t->GetParticleFilter()->GetNumberOfParticles();
        nzAt.push_back(nz.size());
        nz.resize(nz.size() + (size_t)2 * N);
        R.push_back(o); meanFoot.push_back(mf[0]); meanFoot.push_back(mf[1]); w0v.push_back(cur[2]); h0v.push_back(cur[3]);
        if (rearranged) rearranged[o] = 1;
    }
    if (R.empty()) return;
    const int m = (int)R.size();
    // RearrangeParticlesBasedOnGroundLocation: (xRand, yRand) = randgaus(0, 10) per particle, in particle order, from the object's
    // own MWC stream.  The polar method rejects a data-dependent number of draws, so a stream is sequential (~30 ns per value);
    // the objects' streams are independent and are drawn side by side on host threads.
    auto draw = [&](int k) {
        ParticleFilterTracker* t = trackers[R[k]];
        const int N = t->GetParticleFilter()->GetNumberOfParticles();
        float* z = nz.data() + nzAt[k];
        Public::SetCurrentRng(&t->Rng());
        for (int r = 0; r < N; r++) { z[r] = Public::randgaus(0, 10.0f); z[(size_t)N + r] = Public::randgaus(0, 10.0f); }
    };
    {
        const int nth = std::min(m, std::max(1, (int)std::thread::hardware_concurrency() / 8));
        if (nth <= 1) for (int k = 0; k < m; k++) draw(k);
        else {
            std::vector<std::thread> th;
            for (int q = 1; q < nth; q++) th.emplace_back([&, q] { for (int k = q; k < m; k += nth) draw(k); });
            for (int k = 0; k < m; k += nth) draw(k);
            for (std::thread& x : th) x.join();
        }
    }
    std::vector<mct_pf*> rp(m); std::vector<mct_clf*> rc2(m); std::vector<mct_track_params> tp(m); std::vector<float> u01(m);
    std::vector<mct_pf_summary> summ(m);
    for (int k = 0; k < m; k++) {
        ParticleFilterTracker* t = trackers[R[k]];
        rp[k] = t->GetParticleFilter()->Handle(); rc2[k] = t->GetClassifier()->Handle();
        tp[k].w0 = w0v[k]; tp[k].h0 = h0v[k]; tp[k].exponent = 1; tp[k].force_reset = 1; tp[k].resample = 1;
        u01[k] = (float)(t->Rng().Peek() * 2.3283064365386962890625e-10);
    }
    rc = mct_pf_rearrange_ground_many(ctx, m, rp.data(), meanFoot.data(), nz.data(), w0v.data(), h0v.data());
    if (rc) throw MctHostError(rc, mct_last_error(ctx));
    // UpdateParticleWeightsForRearrangedParticles (:1334-1482): score in place, exponent 1, weights reset, resample
    rc = mct_track_score(ctx, m, rp.data(), rc2.data(), tp.data(), nullptr, 0, u01.data(), summ.data());
    if (rc) throw MctHostError(rc, mct_last_error(ctx));
    for (int k = 0; k < m; k++) {
        ParticleFilterTracker* t = trackers[R[k]];
        if (summ[k].resampled) (void)t->Rng().Next();
        t->GetParticleFilter()->AdoptSummary(summ[k]);
    }
}
}  // namespace MultipleCameraTracking
