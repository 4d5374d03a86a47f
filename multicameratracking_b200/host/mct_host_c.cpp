// mct_host_c.cpp -- flat C driver over the C++ host classes, for ctypes (tests, bench.py).
// One "camera" = one mct_ctx + its Matrixu frame + the ParticleFilterTrackers of its objects, i.e. the slice of
// Camera/Object (Camera.cpp:334-397, Object.cpp:198-358) that feeds the hot path.
#include <cstring>
#include <string>
#include <vector>

#include "mct_host.h"
#include "mct_fuser.h"

using namespace MultipleCameraTracking;

#include "mct_host_c.h"

extern "C" {
const char* mcth_last_error(mcth_camera* c) { return c ? c->err.c_str() : "null camera"; }

int mcth_camera_create(int device, void* stream, int n_frames, mcth_camera** out)
{
    *out = nullptr;
    mct_ctx* ctx = nullptr;
    int rc = mct_ctx_create(device, stream, &ctx);
    if (rc) return rc;
    mcth_camera* c = new mcth_camera{ctx, Matrixu(ctx), {}, {}, {}, {}, n_frames, ""};
    *out = c;
    return 0;
}
void mcth_camera_destroy(mcth_camera* c)
{
    if (!c) return;
    for (auto* t : c->simple) delete t;
    for (auto* t : c->trackers) delete t;
    mct_ctx_destroy(c->ctx);
    delete c;
}
mct_ctx* mcth_camera_ctx(mcth_camera* c) { return c->ctx; }

#define GUARD(body)                                                        \
    try { body; return 0; }                                                \
    catch (const MctHostError& e) { c->err = e.what(); return e.code ? e.code : -1; } \
    catch (const std::exception& e) { c->err = e.what(); return -100; }

int mcth_camera_set_frame(mcth_camera* c, const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2,
                          int W, int H, int pitch, int is_device)
{
    GUARD(c->frame.SetFrame(gray, c0, c1, c2, W, H, pitch, is_device));
}

int mcth_camera_set_frame_bgr(mcth_camera* c, const uint8_t* bgr, int W, int H, int pitch_bytes, int use_hsv, int is_device)
{
    GUARD(c->frame.SetFrameBGR(bgr, W, H, pitch_bytes, use_hsv != 0, is_device));
}
int mcth_camera_prefetch_frame_bgr(mcth_camera* c, const uint8_t* bgr, int W, int H, int pitch_bytes, int use_hsv)
{
    GUARD(c->frame.PrefetchFrameBGR(bgr, W, H, pitch_bytes, use_hsv != 0));
}
int mcth_camera_prefetch_frame(mcth_camera* c, const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2,
                               int W, int H, int pitch)
{
    GUARD(c->frame.PrefetchFrame(gray, c0, c1, c2, W, H, pitch));
}

// announce the Gaussian rows of the NEXT mcth_camera_track call so that they are uploaded during the current frame
int mcth_camera_prefetch_noise(mcth_camera* c, const float* noise, size_t n_floats)
{
    int rc = mct_noise_prefetch(c->ctx, noise, n_floats);
    if (rc) c->err = mct_last_error(c->ctx);
    return rc;
}

int mcth_camera_add_object(mcth_camera* c, const mcth_object_params* p)
{
    GUARD({
        c->params.push_back(*p);
        c->trackers.push_back(new ParticleFilterTracker(c->ctx, p->rng_seed));
    });
}

// ---- SimpleTracker objects (the reference's default local tracker) ----
static void make_params(const mcth_object_params& p, Classifier::StrongClassifierParametersBasePtr& cp, SimpleTrackerParametersPtr& tp)
{
    Features::FeatureParametersPtr fp(new Features::FeatureParameters());
    fp->type = (Features::FeatureType)p.feature_type;
    fp->m_width = (uint)p.init_xywh[2]; fp->m_height = (uint)p.init_xywh[3];
    fp->m_featureDimensionHaar = (uint)p.n_haar; fp->m_numberOfBins = (uint)p.n_bins;
    const int F = (int)fp->GetFeatureDimension();
    const int K = int(F * p.percent_selected / 100);
    if (p.strong_type == MCT_STRONG_MILANYBOOST) cp.reset(new Classifier::MILAnyBoostClassifierParameters(K, F));
    else if (p.strong_type == MCT_STRONG_MILENSEMBLE) cp.reset(new Classifier::MILEnsembleClassifierParameters(K, F, p.percent_retained));
    else if (p.strong_type == MCT_STRONG_ADABOOST) cp.reset(new Classifier::AdaBoostClassifierParameters(K, F));
    else cp.reset(new Classifier::MILBoostClassifierParameters(K, F));
    cp->m_weakClassifierType = (Classifier::WeakClassifierType)p.weak_type;
    cp->m_featureParametersPtr = fp;
    tp.reset(new SimpleTrackerParameters());
    tp->m_posRadiusTrain = p.pos_radius_train; tp->m_init_posTrainRadius = p.init_pos_radius_train;
    tp->m_numberOfNegativeTrainingSamples = (uint)p.n_neg_train; tp->m_init_negNumTrain = (uint)p.init_n_neg_train;
    tp->m_searchWindSize = (uint)p.search_win;
    tp->m_initState.assign(p.init_xywh, p.init_xywh + 4);
}
int mcth_camera_add_simple_object(mcth_camera* c, const mcth_object_params* p)
{
    GUARD({
        c->simple_params.push_back(*p);
        c->simple.push_back(new SimpleTracker(c->ctx, p->rng_seed));
    });
}
int mcth_camera_init_simple(mcth_camera* c)
{
    GUARD({
        for (size_t o = 0; o < c->simple.size(); o++) {
            Classifier::StrongClassifierParametersBasePtr cp; SimpleTrackerParametersPtr tp;
            make_params(c->simple_params[o], cp, tp);
            c->simple[o]->InitializeTrackerWithParameters(&c->frame, &c->frame, 0, (uint)c->n_frames, tp, cp, &c->frame);
        }
    });
}
// one frame for all SimpleTracker objects: out_state [n][4] = m_states row (x, y, w, h), out_resp [n] = best response
int mcth_camera_track_simple(mcth_camera* c, int frameind, float* out_state, double* out_resp)
{
    GUARD({
        SimpleTracker::TrackCameraObjects(c->ctx, c->simple, frameind, &c->frame, &c->frame, &c->frame, out_resp);
        if (out_state)
            for (size_t o = 0; o < c->simple.size(); o++)
                for (int q = 0; q < 4; q++) out_state[o * 4 + q] = c->simple[o]->GetCurrentTrackerState()[q];
    });
}
int mcth_simple_selectors(mcth_camera* c, int o, int* idx)
{
    try { vectori s = c->simple[o]->GetClassifier()->GetSelectorList(); for (size_t i = 0; i < s.size(); i++) idx[i] = s[i]; return (int)s.size(); }
    catch (const std::exception& e) { c->err = e.what(); return -1; }
}
mct_clf* mcth_simple_classifier(mcth_camera* c, int o) { return c->simple[o]->GetClassifier()->Handle(); }
uint64_t mcth_simple_rng_state(mcth_camera* c, int o) { return c->simple[o]->Rng().state; }
int mcth_simple_last_test_size(mcth_camera* c, int o) { return c->simple[o]->LastTestSetSize(); }

// frame 0: InitializeTracker for every object (Camera::InitializeCameraTrackers, Camera.cpp:278-321)
int mcth_camera_init_trackers(mcth_camera* c)
{
    GUARD({
        for (size_t o = 0; o < c->trackers.size(); o++) {
            const mcth_object_params& p = c->params[o];
            // Object::InitializeObjectParameters (Object.cpp:198-358)
            Features::FeatureParametersPtr fp(new Features::FeatureParameters());
            fp->type = (Features::FeatureType)p.feature_type;
            fp->m_width = (uint)p.init_xywh[2]; fp->m_height = (uint)p.init_xywh[3];
            fp->m_featureDimensionHaar = (uint)p.n_haar; fp->m_numberOfBins = (uint)p.n_bins;
            const int F = (int)fp->GetFeatureDimension();
            const int K = int(F * p.percent_selected / 100);
            Classifier::StrongClassifierParametersBasePtr cp;
            if (p.strong_type == MCT_STRONG_MILANYBOOST) cp.reset(new Classifier::MILAnyBoostClassifierParameters(K, F));
            else if (p.strong_type == MCT_STRONG_MILENSEMBLE) cp.reset(new Classifier::MILEnsembleClassifierParameters(K, F, p.percent_retained));
            else if (p.strong_type == MCT_STRONG_ADABOOST) cp.reset(new Classifier::AdaBoostClassifierParameters(K, F));
            else cp.reset(new Classifier::MILBoostClassifierParameters(K, F));
            cp->m_weakClassifierType = (Classifier::WeakClassifierType)p.weak_type;
            cp->m_featureParametersPtr = fp;
            TrackerParametersPtr tp(new ParticleFilterTrackerParameters());
            tp->m_posRadiusTrain = p.pos_radius_train; tp->m_init_posTrainRadius = p.init_pos_radius_train;
            tp->m_numberOfNegativeTrainingSamples = (uint)p.n_neg_train; tp->m_init_negNumTrain = (uint)p.init_n_neg_train;
            tp->m_searchWindSize = (uint)p.search_win;
            tp->m_outputTrajectoryOption = (PFTracker_TrajectoryOption)p.trajectory_option;
            tp->m_numberOfParticles = p.n_particles;
            tp->m_positiveSampleStrategy = p.pos_strategy; tp->m_negativeSampleStrategy = p.neg_strategy;
            if (p.max_num_pos_examples > 0) tp->m_maxNumPositiveExamples = p.max_num_pos_examples;
            tp->m_standardDeviationX = p.sd_x; tp->m_standardDeviationY = p.sd_y;
            tp->m_standardDeviationScaleX = p.sd_sx; tp->m_standardDeviationScaleY = p.sd_sy;
            tp->m_initState.assign(p.init_xywh, p.init_xywh + 4);
            c->trackers[o]->InitializeTrackerWithParameters(&c->frame, &c->frame, 0, (uint)c->n_frames, tp, cp, &c->frame);
        }
    });
}

// one frame for all objects: phases bit0 = score path (TrackObjectOnTheGivenFrame + StoreObjectState),
// bit1 = UpdateClassifier.  noise [n_obj][4][N]; out_boxes [n_obj][4] (m_states rows).
int mcth_camera_track(mcth_camera* c, int frameind, const float* noise, int phases, int* out_boxes)
{
    GUARD({
        ParticleFilterTracker::TrackCameraObjects(c->ctx, c->trackers, frameind, &c->frame, &c->frame, &c->frame, noise, phases);
        if (out_boxes && (((phases & 1) && !(phases & 4)) || (phases & 8)))
            for (size_t o = 0; o < c->trackers.size(); o++)
                for (int q = 0; q < 4; q++) out_boxes[o * 4 + q] = c->trackers[o]->States()[(size_t)frameind * 4 + q];
    });
}

// One video frame in one call (the body of the reference's per-frame loop, CameraNetwork.cpp:549-554 / Camera.cpp:449-494):
// make frame t current, announce frame t+1 and its Gaussian rows (uploaded while frame t is tracked; NULL = nothing to
// announce), track.  planes / next_planes: gray, c0, c1, c2 host pointers.
int mcth_camera_step(mcth_camera* c, int frameind, const uint8_t* const* planes, const uint8_t* const* next_planes, int W, int H, int pitch,
                     const float* noise, const float* next_noise, size_t n_noise_floats, int phases, int* out_boxes)
{
    if (phases & 16) {
        // STREAMING score path (phases = 1 | 16): the call returns frame t's boxes after it has already queued frame t + 1 (the
        // announced one), so the device never waits for the host between two frames.  Every call still uploads one frame + its
        // noise from host memory and reads one frame's results back.  A run ends with a call that announces nothing.
        GUARD({
            if ((phases & ~16) != 1) throw MctHostError(MCT_E_ARG, "streaming step: score-only frames (phases = 1 | 16)");
            if (!c->stream_pending || c->stream_gray != (const void*)planes[0] || c->stream_noise != (const void*)noise) {
                if (c->stream_pending) throw MctHostError(MCT_E_STATE, "streaming step: the frame announced by the previous call was queued and must be tracked next");
                c->frame.SetFrame(planes[0], planes[1], planes[2], planes[3], W, H, pitch, 0);
                ParticleFilterTracker::IssueScore(c->ctx, c->trackers, noise);
            }
            c->stream_pending = 0;
            if (next_planes && next_noise) {
                c->frame.PrefetchFrame(next_planes[0], next_planes[1], next_planes[2], next_planes[3], W, H, pitch);
                c->frame.FlushPrefetch();                    // upload + preparation beside the frame in flight
                const int rc = mct_noise_prefetch(c->ctx, next_noise, n_noise_floats);
                if (rc) throw MctHostError(rc, mct_last_error(c->ctx));
                c->frame.SetFrame(next_planes[0], next_planes[1], next_planes[2], next_planes[3], W, H, pitch, 0);
                ParticleFilterTracker::IssueScore(c->ctx, c->trackers, next_noise);
                c->stream_pending = 1; c->stream_gray = next_planes[0]; c->stream_noise = next_noise;
            }
            ParticleFilterTracker::FinishScore(c->ctx, c->trackers, frameind);
            if (out_boxes)
                for (size_t o = 0; o < c->trackers.size(); o++)
                    for (int q = 0; q < 4; q++) out_boxes[o * 4 + q] = c->trackers[o]->States()[(size_t)frameind * 4 + q];
        });
    }
    GUARD({
        c->frame.SetFrame(planes[0], planes[1], planes[2], planes[3], W, H, pitch, 0);
        if (next_planes) c->frame.PrefetchFrame(next_planes[0], next_planes[1], next_planes[2], next_planes[3], W, H, pitch);
        if (next_noise) {
            const int rc = mct_noise_prefetch(c->ctx, next_noise, n_noise_floats);
            if (rc) throw MctHostError(rc, mct_last_error(c->ctx));
        }
        ParticleFilterTracker::TrackCameraObjects(c->ctx, c->trackers, frameind, &c->frame, &c->frame, &c->frame, noise, phases);
        if (out_boxes && (((phases & 1) && !(phases & 4)) || (phases & 8)))
            for (size_t o = 0; o < c->trackers.size(); o++)
                for (int q = 0; q < 4; q++) out_boxes[o * 4 + q] = c->trackers[o]->States()[(size_t)frameind * 4 + q];
    });
}

// score path with device-resident noise, asynchronous (bench.py "value" leg)
int mcth_camera_track_resident(mcth_camera* c, int frameind, const float* noise_dev)
{
    (void)frameind;
    GUARD(ParticleFilterTracker::ScoreCameraObjectsAsync(c->ctx, c->trackers, noise_dev));
}

// geometric-fuser payload of this camera: foot point of the highest-weight particle of every object, packed on
// the device at dev_out [n_obj][2] (ready for an allgather on the same stream)
int mcth_camera_pack_foot_points(mcth_camera* c, float* dev_out)
{
    const int n = (int)c->trackers.size();
    std::vector<mct_pf*> pfs(n); std::vector<float> h0(n);
    for (int o = 0; o < n; o++) { pfs[o] = c->trackers[o]->GetParticleFilter()->Handle(); h0[o] = c->trackers[o]->GetCurrentTrackerState()[3]; }
    int rc = mct_fuser_pack_foot_points(c->ctx, n, pfs.data(), h0.data(), dev_out);
    if (rc) c->err = mct_last_error(c->ctx);
    return rc;
}

// C-ABI handles of object o (particle filter, tracker classifier): for callers that drive extra device work on the camera's
// own context, e.g. the appearance fuser's global classifier (network.AppearanceFusion)
int mcth_object_handles(mcth_camera* c, int o, mct_pf** pf, mct_clf** clf)
{
    if (!c || o < 0 || o >= (int)c->trackers.size()) return MCT_E_ARG;
    if (pf) *pf = c->trackers[o]->GetParticleFilter()->Handle();
    if (clf) *clf = c->trackers[o]->GetClassifier()->Handle();
    return 0;
}
// the tracker's MWC stream: next randfloat() (consumed) -- the appearance fuser's sample dropping draws from it
float mcth_object_randfloat(mcth_camera* c, int o)
{
    Public::SetCurrentRng(&c->trackers[o]->Rng());
    return Public::randfloat();
}

// the same payload exchanged with all cameras over NVLink peer memory (mct_p2p_* set up on c->ctx beforehand):
// host_out [world][n_obj][2]
int mcth_camera_exchange_foot_points(mcth_camera* c, float* host_out, int timeout_ms)
{
    const int n = (int)c->trackers.size();
    std::vector<mct_pf*> pfs(n); std::vector<float> h0(n);
    for (int o = 0; o < n; o++) { pfs[o] = c->trackers[o]->GetParticleFilter()->Handle(); h0[o] = c->trackers[o]->GetCurrentTrackerState()[3]; }
    int rc = mct_fuser_exchange_foot_points(c->ctx, n, pfs.data(), h0.data(), host_out, timeout_ms);
    if (rc) c->err = mct_last_error(c->ctx);
    return rc;
}

// ---- geometric fuser ----
// feedback of the fused ground-plane estimate into object o of this camera (UpdateParticlesWithGroundPDF)
int mcth_camera_ground_feedback(mcth_camera* c, int o, const double* H9, const float* mean2, const float* cov4, int* rearranged,
                                float* distance)
{
    GUARD({
        if (o < 0 || o >= (int)c->trackers.size()) throw MctHostError(MCT_E_ARG, "object index out of range");
        Homography H; for (int i = 0; i < 9; i++) H[i] = H9[i];
        const bool r = UpdateParticlesWithGroundPDF(c->ctx, c->trackers[o], mean2, cov4, H, distance);
        if (rearranged) *rearranged = r ? 1 : 0;
    });
}

// all objects of the camera: one device pass for the pdfs (means [n][2], covs [n][4]; rearranged / distance [n])
int mcth_camera_ground_feedback_all(mcth_camera* c, const double* H9, const float* means, const float* covs, int* rearranged, float* distance)
{
    GUARD({
        Homography H; for (int i = 0; i < 9; i++) H[i] = H9[i];
        UpdateParticlesWithGroundPDF(c->ctx, c->trackers, means, covs, H, rearranged, distance);
    });
}

struct mcth_fuser { GeometryBasedInformationFuser* f; std::string err; };
int mcth_fuser_create(int n_cameras, const double* H /* [n_cameras][9] */, mcth_fuser** out)
{
    *out = nullptr;
    try {
        std::vector<Homography> list(n_cameras > 0 ? n_cameras : 0);
        for (int c = 0; c < n_cameras; c++) for (int i = 0; i < 9; i++) list[c][i] = H[(size_t)c * 9 + i];
        mcth_fuser* f = new mcth_fuser();
        f->f = new GeometryBasedInformationFuser(list);
        *out = f;
        return 0;
    } catch (const MctHostError& e) { return e.code ? e.code : -1; }
    catch (const std::exception&) { return -100; }
}
void mcth_fuser_destroy(mcth_fuser* f) { if (f) { delete f->f; delete f; } }
const char* mcth_fuser_last_error(mcth_fuser* f) { return f ? f->err.c_str() : "null fuser"; }
#define FGUARD(body)                                                       \
    try { body; return 0; }                                                \
    catch (const MctHostError& e) { f->err = e.what(); return e.code ? e.code : -1; } \
    catch (const std::exception& e) { f->err = e.what(); return -100; }
int mcth_fuser_initialize(mcth_fuser* f, const float* foot /* [n_cameras][2] */) { FGUARD(f->f->InitializeKalman(foot)); }
int mcth_fuser_fuse(mcth_fuser* f, const float* foot) { FGUARD(f->f->FuseInformation(foot)); }
int mcth_fuser_get(mcth_fuser* f, float* mean2, float* cov4, float* state4, float* P16)
{
    FGUARD({
        if (mean2) f->f->GetGroundPlaneKalmanMeanMatrix(mean2);
        if (cov4) f->f->GetGroundPlaneKalmanCovarianceMatrix(cov4);
        if (state4) memcpy(state4, f->f->StatePost(), 16);
        if (P16) memcpy(P16, f->f->ErrorCovPost(), 64);
    });
}
// one frame for the fusers of all objects: allp [n_cameras][n_obj][2] (the all-gathered foot points); a fuser that is not
// initialised yet is initialised from this frame (fresh[o] = 1, no feedback for it); means [n_obj][2], covs [n_obj][4]
int mcth_fuser_fuse_many(mcth_fuser** fs, int n_obj, const float* allp, int n_cameras, float* means, float* covs, int* fresh)
{
    for (int o = 0; o < n_obj; o++) {
        mcth_fuser* f = fs[o];
        try {
            std::vector<float> foot((size_t)n_cameras * 2);
            for (int c = 0; c < n_cameras; c++) { foot[2 * c] = allp[((size_t)c * n_obj + o) * 2]; foot[2 * c + 1] = allp[((size_t)c * n_obj + o) * 2 + 1]; }
            const bool first = !f->f->IsInitialized();
            if (first) f->f->InitializeKalman(foot.data()); else f->f->FuseInformation(foot.data());
            if (fresh) fresh[o] = first ? 1 : 0;
            f->f->GetGroundPlaneKalmanMeanMatrix(means + 2 * o);
            f->f->GetGroundPlaneKalmanCovarianceMatrix(covs + 4 * o);
        } catch (const MctHostError& e) { f->err = e.what(); return e.code ? e.code : -1; }
        catch (const std::exception& e) { f->err = e.what(); return -100; }
    }
    return 0;
}
int mcth_fuser_intersection(int n_cameras, const double* H, const float* foot, double* out2)
{
    try {
        std::vector<Homography> list(n_cameras > 0 ? n_cameras : 0);
        for (int c = 0; c < n_cameras; c++) for (int i = 0; i < 9; i++) list[c][i] = H[(size_t)c * 9 + i];
        const std::array<double, 2> g = GeometryBasedInformationFuser::EstimateGroundPlanePrincipleAxisIntersection(foot, list);
        out2[0] = g[0]; out2[1] = g[1];
        return 0;
    } catch (...) { return -1; }
}
// randgaus(0, sigma) stream of a fresh MWC state (Public.cpp:37-49), for the parity tests
int mcth_randgaus(int64_t seed, float sigma, int n, float* out)
{
    Public::RngStream s(seed);
    Public::SetCurrentRng(&s);
    for (int i = 0; i < n; i++) out[i] = Public::randgaus(0, sigma);
    Public::SetCurrentRng(nullptr);
    return 0;
}

int mcth_camera_sync(mcth_camera* c) { return mct_sync(c->ctx); }
long long mcth_camera_launches(mcth_camera* c) { return mct_launch_count(c->ctx); }

// introspection for parity tests
int mcth_object_selectors(mcth_camera* c, int o, int* idx)
{
    try { vectori s = c->trackers[o]->GetClassifier()->GetSelectorList(); for (size_t i = 0; i < s.size(); i++) idx[i] = s[i]; return (int)s.size(); }
    catch (const std::exception& e) { c->err = e.what(); return -1; }
}
int mcth_object_state(mcth_camera* c, int o, float* cur6)
{
    const vectorf& s = c->trackers[o]->GetCurrentTrackerState();
    for (int i = 0; i < 6; i++) cur6[i] = s[i];
    return 0;
}
int mcth_object_particles(mcth_camera* c, int o, float* Nx5)
{
    return mct_pf_get_state(c->trackers[o]->GetParticleFilter()->Handle(), Nx5);
}
int mcth_object_last_train(mcth_camera* c, int o, int which, int* col, int* row, int cap)
{
    const Classifier::SampleSet& s = which == 0 ? c->trackers[o]->PositiveSet() : c->trackers[o]->NegativeSet();
    int n = (int)s.Size() < cap ? (int)s.Size() : cap;
    for (int i = 0; i < n; i++) { col[i] = s[i].m_col; row[i] = s[i].m_row; }
    return n;
}
uint64_t mcth_object_rng_state(mcth_camera* c, int o) { return c->trackers[o]->Rng().state; }
// whole frame on the device (default-strategy training sets from k_train_gen_default) on / off for this process
// layout check for bindings (tests/test_abi.py: the ctypes mirror of mcth_object_params)
int mcth_abi_object_params_size(void) { return (int)sizeof(mcth_object_params); }
void mcth_set_device_training_sets(int on) { ParticleFilterTracker::SetDeviceTrainingSets(on != 0); }
int mcth_get_device_training_sets(void) { return ParticleFilterTracker::DeviceTrainingSets() ? 1 : 0; }
// m_states row of a frame (x, y, w, h as stored by StoreObjectState)
int mcth_object_box(mcth_camera* c, int o, int frameind, int* out4)
{
    if (!c || o < 0 || o >= (int)c->trackers.size()) return MCT_E_ARG;
    const std::vector<int>& st = c->trackers[o]->States();
    if (frameind < 0 || (size_t)frameind * 4 + 3 >= st.size()) return MCT_E_ARG;
    for (int q = 0; q < 4; q++) out4[q] = st[(size_t)frameind * 4 + q];
    return 0;
}
}

// ------------------------------------------------------------------------------------------------------------------
// CameraNetwork (mct_network.h) behind a flat C interface.  exchange == NULL: all cameras of the network live in this process.
#include "mct_network.h"

struct mcth_network_params {
    int geometric_fusion, appearance_fusion, af_strong, af_weak, af_pct_selected, af_pct_retained, af_refresh;
};
struct mcth_exchange {
    int rank, world;
    void* user;
    void (*allgather_host)(void* user, const void* send, void* recv, size_t bytes);
    void (*alltoall_rows)(void* user, size_t slot_bytes, void* stream);
};
namespace {
struct CallbackExchange : public NetworkExchange {
    mcth_exchange e;
    explicit CallbackExchange(const mcth_exchange& x) : e(x) {}
    int Rank() const override { return e.rank; }
    int World() const override { return e.world; }
    void AllGatherHost(const void* send, void* recv, size_t bytes) override { e.allgather_host(e.user, send, recv, bytes); }
    void AllToAllRows(size_t slotBytes, void* stream) override { e.alltoall_rows(e.user, slotBytes, stream); }
};
}  // namespace
struct mcth_network { CameraNetwork* net; CallbackExchange* ex; std::string err; };

extern "C" {
const char* mcth_network_last_error(mcth_network* n) { return n ? n->err.c_str() : "null network"; }
int mcth_network_create(mcth_camera* const* cams, const int* cam_ids, int n_local, int n_cameras, const double* H,
                        const mcth_network_params* p, const mcth_exchange* exchange, mcth_network** out)
{
    *out = nullptr;
    mcth_network* n = new mcth_network{nullptr, nullptr, ""};
    try {
        std::vector<mcth_camera*> cl(cams, cams + n_local);
        std::vector<int> ids(cam_ids, cam_ids + n_local);
        std::vector<Homography> hs;
        if (H) { hs.resize(n_cameras); for (int c = 0; c < n_cameras; c++) for (int i = 0; i < 9; i++) hs[c][i] = H[(size_t)c * 9 + i]; }
        NetworkParameters np;
        np.geometricFusionType = p->geometric_fusion; np.appearanceFusionType = p->appearance_fusion;
        np.appearanceFusionStrongClassifierType = p->af_strong; np.appearanceFusionWeakClassifierType = p->af_weak;
        np.afPercentageOfWeakClassifiersSelected = p->af_pct_selected; np.afPercentageOfWeakClassifiersRetained = p->af_pct_retained;
        np.appearanceFusionRefreshRate = p->af_refresh;
        if (exchange) n->ex = new CallbackExchange(*exchange);
        n->net = new CameraNetwork(cl, ids, n_cameras, hs, np, n->ex);
        *out = n;
        return 0;
    } catch (const MctHostError& e) { n->err = e.what(); *out = n; return e.code ? e.code : -1; }
    catch (const std::exception& e) { n->err = e.what(); *out = n; return -100; }
}
void mcth_network_destroy(mcth_network* n) { if (n) { delete n->net; delete n->ex; delete n; } }
#define NGUARD(body)                                                       \
    try { body; return 0; }                                                \
    catch (const MctHostError& e) { n->err = e.what(); return e.code ? e.code : -1; } \
    catch (const std::exception& e) { n->err = e.what(); return -100; }
int mcth_network_init(mcth_network* n) { NGUARD(n->net->InitializeCameraNetwork()); }
int mcth_network_track(mcth_network* n, int frame, const float* const* noise) { NGUARD(n->net->TrackObjectsOnCurrentFrame(frame, noise)); }
size_t mcth_network_row_bytes(mcth_network* n) { return n->net->RowBufferBytes(); }
int mcth_network_set_row_buffers(mcth_network* n, void* send_dev, void* recv_dev) { NGUARD(n->net->SetRowBuffers(send_dev, recv_dev)); }
int mcth_network_kalman(mcth_network* n, int o, float* state4, float* P16)
{
    const GeometryBasedInformationFuser* f = n->net->GeometricFuser(o);
    if (!f) return -1;
    memcpy(state4, f->StatePost(), 16); memcpy(P16, f->ErrorCovPost(), 64);
    return 0;
}
int mcth_network_global_selectors(mcth_network* n, int local_cam, int o, int* idx)
{
    try { vectori s = n->net->GlobalClassifier(local_cam, o)->GetSelectorList(); for (size_t i = 0; i < s.size(); i++) idx[i] = s[i]; return (int)s.size(); }
    catch (const std::exception& e) { n->err = e.what(); return -1; }
}
mct_clf* mcth_network_global_classifier(mcth_network* n, int local_cam, int o) { return n->net->GlobalClassifier(local_cam, o)->Handle(); }
int mcth_network_last_rearranged(mcth_network* n, int local_cam, int o) { return n->net->LastRearranged(local_cam, o); }
int mcth_network_stage_ms(mcth_network* n, double* out8) { memcpy(out8, n->net->StageMilliseconds(), 8 * sizeof(double)); return 0; }
}
