// mct_network.cpp -- see mct_network.h.  All reference citations are /root/reference/src/.
#include "mct_network.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mct_host_c.h"

namespace MultipleCameraTracking {
namespace {
inline void chk(mct_ctx* ctx, int rc) { if (rc) throw MctHostError(rc, mct_last_error(ctx)); }
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline int round_up32(int v) { return (v + 31) / 32 * 32; }
}  // namespace

CameraNetwork::CameraNetwork(const std::vector<mcth_camera*>& localCameras, const std::vector<int>& cameraIds, int numberOfCameras,
                             const std::vector<Homography>& homographies, const NetworkParameters& params, NetworkExchange* exchange)
    : m_cameras(localCameras), m_cameraIds(cameraIds), m_numberOfCameras(numberOfCameras), m_numberOfObjects(0), m_homographies(homographies),
      m_params(params), m_exchange(exchange), m_afFeatures(0), m_ldPos(0), m_ldRow(0), m_rowBytes(0), m_ownRowBuffers(false)
{
    if (m_cameras.empty() || m_cameras.size() != m_cameraIds.size()) throw MctHostError(MCT_E_ARG, "CameraNetwork: no local cameras");
    if (m_exchange && m_cameras.size() != 1) throw MctHostError(MCT_E_ARG, "CameraNetwork: one camera per process when the network spans processes");
    if (!m_exchange && (int)m_cameras.size() != m_numberOfCameras) throw MctHostError(MCT_E_ARG, "CameraNetwork: cameras missing and no exchange given");
    if (m_exchange && (m_exchange->World() != m_numberOfCameras || m_cameraIds[0] != m_exchange->Rank()))
        throw MctHostError(MCT_E_ARG, "CameraNetwork: rank r must hold camera r");
    m_numberOfObjects = (int)m_cameras[0]->trackers.size();
    for (mcth_camera* c : m_cameras)
        if ((int)c->trackers.size() != m_numberOfObjects) throw MctHostError(MCT_E_ARG, "CameraNetwork: every camera tracks the same objects");
    if (m_params.geometricFusionType && (int)m_homographies.size() != m_numberOfCameras)
        throw MctHostError(MCT_E_ARG, "CameraNetwork: one homography per camera is needed");
    if (m_params.appearanceFusionRefreshRate < 1) m_params.appearanceFusionRefreshRate = 1;
    std::memset(m_stageMs, 0, sizeof(m_stageMs));
    m_lastRearranged.assign(m_cameras.size(), std::vector<int>(m_numberOfObjects, 0));
}

CameraNetwork::~CameraNetwork()
{
    if (m_ownRowBuffers)
        for (size_t i = 0; i < m_cameras.size(); i++) {
            if (i < m_sendRows.size()) mct_dev_free(m_cameras[i]->ctx, m_sendRows[i]);
            if (i < m_recvRows.size()) mct_dev_free(m_cameras[i]->ctx, m_recvRows[i]);
        }
}

void CameraNetwork::SetRowBuffers(void* sendDev, void* recvDev)
{
    if (m_cameras.size() != 1) throw MctHostError(MCT_E_ARG, "SetRowBuffers: one camera per process");
    m_sendRows.assign(1, sendDev); m_recvRows.assign(1, recvDev); m_ownRowBuffers = false;
}

// ------------------------------------------------------------------------------------------------ initialisation
extern "C" int mcth_camera_init_trackers(mcth_camera* c);

void CameraNetwork::InitializeCameraNetwork()
{
    // Object::InitializeObjectParameters (Object.cpp:203-216): the appearance fuser (global classifier, learning rate 0,
    // AppearanceBasedInformationFuser.cpp:28) exists before the tracker; colour features draw nothing from the RNG
    if (m_params.appearanceFusionType) {
        m_globalClassifiers.resize(m_cameras.size()); m_globalTrained.resize(m_cameras.size());
        for (size_t i = 0; i < m_cameras.size(); i++) {
            mcth_camera* cam = m_cameras[i];
            for (int o = 0; o < m_numberOfObjects; o++) {
                const mcth_object_params& p = cam->params[o];
                Features::FeatureParametersPtr fp(new Features::FeatureParameters());
                // Object::GenerateDefaultAppearanceFusionFeatureParameters (Object.cpp:75-104): default colour parameters (RGB, 8 bins)
                fp->type = m_params.appearanceFusionType == 1 ? Features::CULTURE_COLOR_HISTOGRAM : Features::MULTI_DIMENSIONAL_COLOR_HISTOGRAM;
                fp->m_numberOfBins = 8; fp->m_useHSVColorSpace = false;
                fp->m_width = (uint)p.init_xywh[2]; fp->m_height = (uint)p.init_xywh[3];
                const int F = (int)fp->GetFeatureDimension();
                const int K = int(F * m_params.afPercentageOfWeakClassifiersSelected / 100);      // Object.cpp:122-123
                Classifier::StrongClassifierParametersBasePtr cp;
                if (m_params.appearanceFusionStrongClassifierType == 2) cp.reset(new Classifier::AdaBoostClassifierParameters(K, F));
                else if (m_params.appearanceFusionStrongClassifierType == 3) {
                    // Object.cpp:149-157: the RETAINED count is passed where the constructor expects a percentage
                    const int retained = int(F * m_params.afPercentageOfWeakClassifiersRetained / 100);
                    cp.reset(new Classifier::MILEnsembleClassifierParameters(K, F, (float)retained));
                } else cp.reset(new Classifier::MILBoostClassifierParameters(K, F));
                cp->m_weakClassifierType = m_params.appearanceFusionStrongClassifierType == 3 ? Classifier::PERCEPTRON
                                                                                              : (Classifier::WeakClassifierType)m_params.appearanceFusionWeakClassifierType;
                cp->m_learningRate = 0;
                cp->m_featureParametersPtr = fp;
                m_globalClassifiers[i].push_back(Classifier::StrongClassifierFactory::CreateAndInitializeClassifier(cp, cam->ctx));
                m_afFeatures = F;
            }
            m_globalTrained[i].assign(m_numberOfObjects, 0);
        }
    }
    // Camera::InitializeCameraTrackers (Camera.cpp:278-321)
    for (mcth_camera* cam : m_cameras) {
        const int rc = mcth_camera_init_trackers(cam);
        if (rc) throw MctHostError(rc, cam->err);
    }
    if (m_params.appearanceFusionType) {
        // exchange buffers sized for the largest training sets the trackers can produce: positives = the disc of radius
        // m_posRadiusTrain (or m_maxNumPositiveExamples unique particles), negatives <= m_numberOfNegativeTrainingSamples
        int maxPos = 1, maxNeg = 1;
        for (mcth_camera* cam : m_cameras)
            for (const mcth_object_params& p : cam->params) {
                const int r = (int)std::ceil(p.pos_radius_train);
                int pos = (2 * r + 1) * (2 * r + 1);
                if (p.pos_strategy != 0) pos = std::max(pos, p.max_num_pos_examples);
                maxPos = std::max(maxPos, pos); maxNeg = std::max(maxNeg, p.n_neg_train);
            }
        if (m_exchange) {
            int mine[2] = {maxPos, maxNeg};
            std::vector<int> all((size_t)2 * m_numberOfCameras);
            m_exchange->AllGatherHost(mine, all.data(), sizeof(mine));
            for (int c = 0; c < m_numberOfCameras; c++) { maxPos = std::max(maxPos, all[2 * c]); maxNeg = std::max(maxNeg, all[2 * c + 1]); }
        }
        m_ldPos = round_up32(maxPos);
        m_ldRow = m_ldPos + round_up32(maxNeg);
        m_rowBytes = (size_t)m_numberOfObjects * m_afFeatures * m_ldRow * sizeof(float);
        if (!m_exchange) {
            m_sendRows.assign(m_cameras.size(), nullptr); m_recvRows.assign(m_cameras.size(), nullptr);
            for (size_t i = 0; i < m_cameras.size(); i++) {
                chk(m_cameras[i]->ctx, mct_dev_alloc(m_cameras[i]->ctx, m_rowBytes * m_numberOfCameras, &m_sendRows[i]));
                chk(m_cameras[i]->ctx, mct_dev_alloc(m_cameras[i]->ctx, m_rowBytes * m_numberOfCameras, &m_recvRows[i]));
            }
            m_ownRowBuffers = true;
        }
    }
    // CameraNetwork::InitializeGeometricFuser (:380-470): one fuser per object, Kalman state from the cameras' foot points
    if (m_params.geometricFusionType) {
        for (int o = 0; o < m_numberOfObjects; o++) m_geometricFusers.emplace_back(new GeometryBasedInformationFuser(m_homographies));
        FuseGeometrically(0, true);
    }
}

// ------------------------------------------------------------------------------------------------ geometric fuser
// the payload of the geometric fuser: foot point of the highest-weight particle of every object of every camera
// (ParticleFilterTracker.cpp:1223-1238), [camera][object][2]
void CameraNetwork::GatherFootPoints(std::vector<float>& all)
{
    const int n = m_numberOfObjects;
    all.assign((size_t)m_numberOfCameras * n * 2, 0.0f);
    std::vector<float> mine((size_t)m_cameras.size() * n * 2);
    for (size_t i = 0; i < m_cameras.size(); i++) {
        mcth_camera* cam = m_cameras[i];
        std::vector<mct_pf*> pfs(n); std::vector<float> h0(n);
        for (int o = 0; o < n; o++) { pfs[o] = cam->trackers[o]->GetParticleFilter()->Handle(); h0[o] = cam->trackers[o]->GetCurrentTrackerState()[3]; }
        chk(cam->ctx, mct_fuser_foot_points(cam->ctx, n, pfs.data(), h0.data(), mine.data() + i * n * 2));
    }
    if (m_exchange) m_exchange->AllGatherHost(mine.data(), all.data(), (size_t)n * 2 * sizeof(float));
    else for (size_t i = 0; i < m_cameras.size(); i++) std::memcpy(all.data() + (size_t)m_cameraIds[i] * n * 2, mine.data() + i * n * 2, (size_t)n * 2 * sizeof(float));
}

// FuseGeometrically (:175-268) + FeedbackInformationFromGeometricFusion (:94-128)
void CameraNetwork::FuseGeometrically(int frameIndex, bool initialise)
{
    const int n = m_numberOfObjects;
    double t0 = now_ms();
    std::vector<float> all;
    GatherFootPoints(all);
    double t1 = now_ms();
    std::vector<float> means((size_t)n * 2), covs((size_t)n * 4), foot((size_t)m_numberOfCameras * 2);
    for (int o = 0; o < n; o++) {
        for (int c = 0; c < m_numberOfCameras; c++) { foot[2 * c] = all[((size_t)c * n + o) * 2]; foot[2 * c + 1] = all[((size_t)c * n + o) * 2 + 1]; }
        if (initialise) m_geometricFusers[o]->InitializeKalman(foot.data());
        else m_geometricFusers[o]->FuseInformation(foot.data());
        m_geometricFusers[o]->GetGroundPlaneKalmanMeanMatrix(&means[2 * o]);
        m_geometricFusers[o]->GetGroundPlaneKalmanCovarianceMatrix(&covs[4 * o]);
    }
    double t2 = now_ms();
    if (!initialise && frameIndex > 0)
        for (size_t i = 0; i < m_cameras.size(); i++)
            UpdateParticlesWithGroundPDF(m_cameras[i]->ctx, m_cameras[i]->trackers, means.data(), covs.data(), m_homographies[m_cameraIds[i]],
                                         m_lastRearranged[i].data(), nullptr);
    double t3 = now_ms();
    m_stageMs[2] = t1 - t0; m_stageMs[3] = t2 - t1; m_stageMs[4] = t3 - t2;
}

// ------------------------------------------------------------------------------------------------ appearance fuser
// Object::UpdateParticleWeightUsingMultiCameraAppearanceModel (Object.cpp:496-528) for all objects of a camera in one batched
// pass: test set in place, global Classify, sigmoid (AppearanceBasedInformationFuser.cpp:98-121), UpdateParticleWeights(true, false)
void CameraNetwork::UpdateParticleWeightsUsingMultiCameraAppearanceModel(int i)
{
    mcth_camera* cam = m_cameras[i];
    const int n = m_numberOfObjects;
    std::vector<mct_pf*> pfs(n); std::vector<mct_clf*> clfs(n); std::vector<mct_track_params> tp(n); std::vector<float> u01(n);
    std::vector<mct_pf_summary> summ(n);
    for (int o = 0; o < n; o++) {
        ParticleFilterTracker* t = cam->trackers[o];
        pfs[o] = t->GetParticleFilter()->Handle(); clfs[o] = m_globalClassifiers[i][o]->Handle();
        tp[o].w0 = t->GetCurrentTrackerState()[2]; tp[o].h0 = t->GetCurrentTrackerState()[3];
        tp[o].exponent = 0; tp[o].force_reset = 0; tp[o].resample = 1;
        u01[o] = (float)(t->Rng().Peek() * 2.3283064365386962890625e-10);
    }
    chk(cam->ctx, mct_track_score(cam->ctx, n, pfs.data(), clfs.data(), tp.data(), nullptr, 0, u01.data(), summ.data()));
    for (int o = 0; o < n; o++) {
        if (summ[o].resampled) (void)cam->trackers[o]->Rng().Next();
        cam->trackers[o]->GetParticleFilter()->AdoptSummary(summ[o]);
    }
}

// Object.cpp:449 -> ParticleFilterTracker::ForceParticleResampling -> ResampleParticles(false): one draw per object
void CameraNetwork::ForceParticleResampling(int i)
{
    mcth_camera* cam = m_cameras[i];
    const int n = m_numberOfObjects;
    std::vector<mct_pf*> pfs(n); std::vector<float> u01(n);
    for (int o = 0; o < n; o++) {
        pfs[o] = cam->trackers[o]->GetParticleFilter()->Handle();
        u01[o] = (float)(cam->trackers[o]->Rng().Next() * 2.3283064365386962890625e-10);
    }
    chk(cam->ctx, mct_pf_resample_many(cam->ctx, n, pfs.data(), 0, u01.data()));
    for (int o = 0; o < n; o++) cam->trackers[o]->GetParticleFilter()->Invalidate();
}

// CameraNetwork.cpp:589-600 -> Camera::LearnGlobalAppearanceModel (Camera.cpp:722-740) -> AppearanceBasedInformationFuser::
// LearnGlobalAppearanceModel (:44-90).  The reference runs, for every learner camera c and object o, GenerateTrainingSampleSets-
// ForAppearanceFusion over ALL cameras (CameraNetwork.cpp:276-320): every tracker regenerates its training sets once per learner
// (the negative ring is thinned with fresh randfloat() draws each time), and the learner drops pooled samples with its own draws
// (:66-82).  What orders the rounds is therefore the HOST side only (each tracker's MWC stream: regenerate, regenerate, ...,
// with the drop draws of the round in which its camera learns in between); the device work of a round depends on nothing from
// the other rounds.  So:
//   round L = 0 .. C-1 : every camera regenerates its sets (host RNG) and launches the feature rows of all its objects into send
//                        slot L (asynchronous); the (P, Nn) counts of the round go round (small host exchange); the learner
//                        camera L draws its drop decisions (host RNG) and keeps the column offsets;
//   then               : ONE all-to-all of the row buffers (camera c's slot L goes to camera L) and ONE batched Update of the
//                        global classifiers per camera -- all cameras learn at the same time instead of one after the other.
void CameraNetwork::LearnGlobalAppearanceModel()
{
    const int n = m_numberOfObjects, F = m_afFeatures, C = m_numberOfCameras;
    const size_t nl = m_cameras.size();
    const size_t slotFloats = (size_t)n * F * m_ldRow;                         // one camera's rows of one round
    std::vector<int> counts((size_t)C * n * 2), mine(nl * n * 2);
    struct LearnerPlan { std::vector<int> off, Pk, Nk; };
    std::vector<LearnerPlan> plan(nl);
    const bool batched = m_params.appearanceFusionType == 2;
    // MCT_HOST_TIMING=1: where the host time of this stage goes (stderr, every 16 calls)
    static const bool timing = getenv("MCT_HOST_TIMING") != nullptr;
    static double tacc[6] = {0, 0, 0, 0, 0, 0}; static int tcalls = 0;
    double tm = now_ms();
    auto lap = [&](int k) { if (timing) { const double t = now_ms(); tacc[k] += t - tm; tm = t; } };
    // positives of the first round per local camera: later rounds regenerate the same boxes (the tracker state does not change
    // between the rounds and the positive strategy draws nothing unless it thins), which is checked, not assumed
    std::vector<std::vector<vectori>> pc0(nl), pr0(nl);
    std::vector<std::vector<vectorf>> psx0(nl), psy0(nl);
    for (int learner = 0; learner < C; learner++) {
        // 1. every camera: GenerateTrainingSampleSet for every object (host: RNG order), feature rows of all objects into send
        //    slot `learner` in one batched device pass
        for (size_t i = 0; i < nl; i++) {
            mcth_camera* cam = m_cameras[i];
            std::vector<vectori> pc(n), pr(n), nc(n), nr(n);
            std::vector<vectorf> psx(n), psy(n), nsx(n), nsy(n);
            std::vector<mct_boxes> pb(n), nb(n);
            std::vector<mct_clf*> gl(n);
            ParticleFilterTracker::EnsureSummaries(cam->ctx, cam->trackers);
            for (int o = 0; o < n; o++) {
                ParticleFilterTracker* t = cam->trackers[o];
                t->GenerateTrainingSampleSet(&cam->frame, &cam->frame, &cam->frame);
                t->PositiveSet().ToBoxes(pc[o], pr[o], psx[o], psy[o]); t->NegativeSet().ToBoxes(nc[o], nr[o], nsx[o], nsy[o]);
                const int P = (int)pc[o].size(), Nn = (int)nc[o].size();
                if (P > m_ldPos || Nn > m_ldRow - m_ldPos) throw MctHostError(MCT_E_ARG, "appearance fusion: training set larger than the exchange buffer");
                pb[o] = {P, pc[o].data(), pr[o].data(), psx[o].data(), psy[o].data()};
                nb[o] = {Nn, nc[o].data(), nr[o].data(), nsx[o].data(), nsy[o].data()};
                gl[o] = m_globalClassifiers[i][o]->Handle();
                mine[(i * n + o) * 2] = P; mine[(i * n + o) * 2 + 1] = Nn;
            }
            lap(0);
            float* rows = (float*)m_sendRows[i] + (size_t)learner * slotFloats;
            if (batched) {
                bool same = learner > 0;
                if (learner == 0) { pc0[i] = pc; pr0[i] = pr; psx0[i] = psx; psy0[i] = psy; }
                else for (int o = 0; o < n && same; o++) same = pc[o] == pc0[i][o] && pr[o] == pr0[i][o] && psx[o] == psx0[i][o] && psy[o] == psy0[i][o];
                chk(cam->ctx, mct_train_features_many_dev(cam->ctx, n, gl.data(), pb.data(), nb.data(), rows, (size_t)F * m_ldRow, m_ldRow, m_ldPos,
                                                          same ? (const float*)m_sendRows[i] : nullptr));
            } else {
                for (int o = 0; o < n; o++) {
                    float* base = rows + (size_t)o * F * m_ldRow;
                    chk(cam->ctx, mct_features_compute_dev(gl[o], &pb[o], base, m_ldRow));
                    chk(cam->ctx, mct_features_compute_dev(gl[o], &nb[o], base + m_ldPos, m_ldRow));
                }
            }
        }
        lap(1);
        // 2. the counts of the round to everybody (the learner needs them before its next regeneration)
        if (m_exchange) m_exchange->AllGatherHost(mine.data(), counts.data(), (size_t)n * 2 * sizeof(int));
        else for (size_t i = 0; i < nl; i++) std::memcpy(counts.data() + (size_t)m_cameraIds[i] * n * 2, mine.data() + i * n * 2, (size_t)n * 2 * sizeof(int));
        lap(2);
        int li = -1;                                   // local index of the learner camera, if it lives here
        for (size_t i = 0; i < nl; i++) if (m_cameraIds[i] == learner) li = (int)i;
        if (li < 0) continue;
        // 3. the learner's drop rule: pooled order = positives of camera 0..C-1, then negatives of camera 0..C-1; offsets address
        //    its receive buffer [camera][object][F][ldRow]
        mcth_camera* cam = m_cameras[li];
        LearnerPlan& lp = plan[li];
        lp.Pk.assign(n, 0); lp.Nk.assign(n, 0);
        for (int o = 0; o < n; o++) {
            Public::SetCurrentRng(&cam->trackers[o]->Rng());
            int P = 0, Nn = 0;
            for (int c = 0; c < C; c++)
                for (int s = 0; s < counts[((size_t)c * n + o) * 2]; s++)
                    if (Public::randfloat() > 0.5f) { lp.off.push_back((int)(c * slotFloats + (size_t)o * F * m_ldRow + s)); P++; }
            for (int c = 0; c < C; c++)
                for (int s = 0; s < counts[((size_t)c * n + o) * 2 + 1]; s++)
                    if (Public::randfloat() > 0.5f) { lp.off.push_back((int)(c * slotFloats + (size_t)o * F * m_ldRow + m_ldPos + s)); Nn++; }
            if (P < 1 || Nn < 1) throw MctHostError(MCT_E_EMPTY, "appearance fusion: every pooled sample of a class was dropped");
            lp.Pk[o] = P; lp.Nk[o] = Nn;
            m_globalTrained[li][o] = 1;
        }
    }
    lap(3);
    // 4. rows to their learners
    if (m_exchange) {
        m_exchange->AllToAllRows(m_rowBytes, mct_ctx_stream(m_cameras[0]->ctx));
    } else {
        for (size_t i = 0; i < nl; i++) chk(m_cameras[i]->ctx, mct_sync(m_cameras[i]->ctx));
        for (size_t li = 0; li < nl; li++)
            for (size_t i = 0; i < nl; i++)
                chk(m_cameras[li]->ctx, mct_dev_copy(m_cameras[li]->ctx, (char*)m_recvRows[li] + (size_t)m_cameraIds[i] * m_rowBytes,
                                                     (char*)m_sendRows[i] + (size_t)m_cameraIds[li] * m_rowBytes, m_rowBytes));
    }
    lap(4);
    // 5. every local camera updates its global classifiers (one batched pass per camera, the cameras side by side)
    for (size_t li = 0; li < nl; li++) {
        mcth_camera* cam = m_cameras[li];
        LearnerPlan& lp = plan[li];
        std::vector<mct_clf*> gl(n);
        for (int o = 0; o < n; o++) gl[o] = m_globalClassifiers[li][o]->Handle();
        if (batched) {
            chk(cam->ctx, mct_track_update_from_gathered(cam->ctx, n, gl.data(), (const float*)m_recvRows[li], lp.off.data(), m_ldRow, lp.Pk.data(), lp.Nk.data()));
        } else {
            size_t at = 0;
            for (int o = 0; o < n; o++) {
                chk(cam->ctx, mct_classifier_update_from_gathered(gl[o], (const float*)m_recvRows[li], lp.off.data() + at, m_ldRow, lp.Pk[o], lp.Nk[o]));
                at += (size_t)lp.Pk[o] + lp.Nk[o];
            }
        }
    }
    for (size_t li = 0; li < nl; li++) chk(m_cameras[li]->ctx, mct_sync(m_cameras[li]->ctx));      // the offsets above are host vectors
    lap(5);
    if (timing && ++tcalls == 16) {
        fprintf(stderr, "LearnGlobalAppearanceModel host ms per frame: training sets %.3f, feature-row launches %.3f, counts exchange %.3f, drop rule %.3f, "
                        "row exchange %.3f, update + drain %.3f\n", tacc[0] / 16, tacc[1] / 16, tacc[2] / 16, tacc[3] / 16, tacc[4] / 16, tacc[5] / 16);
        tcalls = 0; for (double& x : tacc) x = 0;
    }
}

// ------------------------------------------------------------------------------------------------ the frame
void CameraNetwork::TrackObjectsOnCurrentFrame(int frameIndex, const float* const* noise)
{
    const bool fusion = m_params.geometricFusionType || m_params.appearanceFusionType;
    double t0 = now_ms();
    std::memset(m_stageMs, 0, sizeof(m_stageMs));
    // :549-554 -> Camera::TrackCameraFrame -> Object::TrackObjectFrame (Object.cpp:411-455)
    for (size_t i = 0; i < m_cameras.size(); i++) {
        mcth_camera* cam = m_cameras[i];
        // no fusion: TrackObjectAndSaveState (score, store, update); fusion: TrackObjectWithoutSaveState
        ParticleFilterTracker::TrackCameraObjects(cam->ctx, cam->trackers, frameIndex, &cam->frame, &cam->frame, &cam->frame, noise[i], fusion ? (1 | 4) : 3);
    }
    double t1 = now_ms();
    m_stageMs[0] = t1 - t0;
    if (!fusion) return;                                                                                       // :557-561
    for (size_t i = 0; i < m_cameras.size(); i++) {
        if (m_params.appearanceFusionType && frameIndex > 2 && frameIndex % m_params.appearanceFusionRefreshRate == 0)      // Object.cpp:441-446
            UpdateParticleWeightsUsingMultiCameraAppearanceModel((int)i);
        ForceParticleResampling((int)i);                                                                       // Object.cpp:449
    }
    double t2 = now_ms();
    m_stageMs[1] = t2 - t1;
    if (m_params.geometricFusionType) FuseGeometrically(frameIndex, false);                                    // :564-581
    double t3 = now_ms();
    for (size_t i = 0; i < m_cameras.size(); i++) {                                                            // :583-587
        mcth_camera* cam = m_cameras[i];
        ParticleFilterTracker::TrackCameraObjects(cam->ctx, cam->trackers, frameIndex, &cam->frame, &cam->frame, &cam->frame, nullptr, 2);
    }
    double t4 = now_ms();
    m_stageMs[5] = t4 - t3;
    if (m_params.appearanceFusionType && (frameIndex + 1) % m_params.appearanceFusionRefreshRate == 0) LearnGlobalAppearanceModel();   // :589-600
    double t5 = now_ms();
    m_stageMs[6] = t5 - t4;
    for (size_t i = 0; i < m_cameras.size(); i++) {                                                            // :602-608
        mcth_camera* cam = m_cameras[i];
        ParticleFilterTracker::TrackCameraObjects(cam->ctx, cam->trackers, frameIndex, &cam->frame, &cam->frame, &cam->frame, nullptr, 8);
    }
    m_stageMs[7] = now_ms() - t5;
}
}  // namespace MultipleCameraTracking
