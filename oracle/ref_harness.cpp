/*
 * ref_harness.cpp -- C entry points that DRIVE THE UNMODIFIED REFERENCE SOURCES (oracle/_ref/libmct_ref.so).
 *
 * TEST INFRASTRUCTURE ONLY.  This file contains no restatement of any algorithm: every function below builds the
 * reference's own objects (Matrixu, SampleSet, StrongClassifierFactory, ParticleFilter, ParticleFilterTracker, ...)
 * from plain arrays, calls the reference's own methods, and copies the results back out.  The reference .cpp files
 * are compiled where they lie under /root/reference/src against the header shims in oracle/ref_shim/ (Intel IPP,
 * OpenCV 2.3.1 C API, boost::shared_ptr -- none of which is in this image).  Build recipe: oracle/build_ref.py.
 *
 * Private members are read through "#define private public" in THIS translation unit only (object layout does not
 * depend on access specifiers); the reference translation units are compiled untouched.
 *
 * Used by tests/test_ref_pins_oracle.py to pin oracle/mct_oracle.c row by row, and by
 * tests/golden/make_ref_golden.py to produce reference-generated fixtures.
 */
#include <iostream>
#include <fstream>
#include <sstream>
#include <iomanip>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <cassert>
#include <algorithm>
#include <numeric>
#include <vector>
#include <string>
#include <list>
#include <memory>
#include <typeinfo>
#include <new>
#include <exception>
#include <stdint.h>

#define private public
#define protected public
#include "Public.h"
#include "Matrix.h"
#include "SampleSet.h"
#include "StrongClassifierBase.h"
#include "StrongClassifierFactory.h"
#include "HaarFeature.h"
#include "HaarFeatureVector.h"
#include "MILBoostClassifier.h"
#include "MILAnyBoostClassifier.h"
#include "MILEnsembleClassifier.h"
#include "AdaBoostClassifier.h"
#include "ParticleFilter.h"
#include "ParticleFilterTracker.h"
#include "SimpleTracker.h"
#include "GeometryBasedInformationFuser.h"
#include "AppearanceBasedInformationFuser.h"
#include "CameraNetwork.h"
#undef private
#undef protected

using namespace MultipleCameraTracking;

extern "C" void mct_shim_set_noise(const float* rows, size_t n);
extern "C" size_t mct_shim_noise_consumed();
extern "C" { extern uint64_t* mct_shim_rng_override; }

/* abortError -> exit(0) inside the reference throws MctRefAbort (ref_shim/boost/shared_ptr.hpp); entry points that can reach
 * an assertion of the reference return -1 / NULL instead of ending the process */
static int g_aborts = 0;
#define REF_TRY(body, fail)                                                                         \
    try { body; }                                                                                   \
    catch (const MctRefAbort&) { g_aborts++; mct_shim_rng_override = 0; mct_shim_set_noise(NULL, 0); return fail; }
extern "C" int ref_abort_count(void) { return g_aborts; }

namespace {
struct Frame {
    Matrixu gray, color, hsv;
    int W, H;
};
struct Clf {
    Classifier::StrongClassifierParametersBasePtr params;
    Classifier::StrongClassifierBasePtr clf;
    int w0, h0;
};
void fill_set(Classifier::SampleSet& set, Frame* f, int n, const int* col, const int* row, const float* sx, const float* sy, int w0, int h0)
{
    for (int i = 0; i < n; i++)
        set.PushBackSample(&f->gray, col[i], row[i], w0, h0, 1.0f, &f->color, &f->hsv, sx ? sx[i] : 1.0f, sy ? sy[i] : 1.0f);
}
Classifier::StrongClassifierParametersBasePtr make_clf_params(int feature_type, int n_haar, int nb, int weak_type, int strong_type, int K,
                                                               float lr, int use_hsv, float pct_retained)
{
    Features::FeatureParametersPtr fp;
    switch (feature_type) {
    case Features::HAAR_LIKE: fp = Features::FeatureParametersPtr(new Features::HaarFeatureParameters(n_haar)); break;
    case Features::CULTURE_COLOR_HISTOGRAM: fp = Features::FeatureParametersPtr(new Features::CultureColorHistogramParameters()); break;
    case Features::MULTI_DIMENSIONAL_COLOR_HISTOGRAM:
        fp = Features::FeatureParametersPtr(new Features::MultiDimensionalColorHistogramParameters(use_hsv != 0, nb)); break;
    default: fp = Features::FeatureParametersPtr(new Features::HaarAndColorHistogramFeatureParameters(n_haar, use_hsv != 0, nb)); break;
    }
    int F = fp->GetFeatureDimension();
    Classifier::StrongClassifierParametersBasePtr p;
    switch (strong_type) {
    case Classifier::ONLINE_ADABOOST: p.reset(new Classifier::AdaBoostClassifierParameters(K, F)); break;
    case Classifier::ONLINE_ENSEMBLE_BOOST_MIL: p.reset(new Classifier::MILEnsembleClassifierParameters(K, F, pct_retained)); break;
    case Classifier::ONLINE_ANY_BOOST_MIL: p.reset(new Classifier::MILAnyBoostClassifierParameters(K, F)); break;
    default: p.reset(new Classifier::MILBoostClassifierParameters(K, F)); break;
    }
    p->m_weakClassifierType = (Classifier::WeakClassifierType)weak_type;
    p->m_learningRate = lr;
    p->m_storeFeatureHistory = false;      /* no "Haar Features" PNG dump (StrongClassifierBase.cpp:36-39, dropped: SURVEY a14) */
    p->m_featureParametersPtr = fp;
    return p;
}
}  // namespace

extern "C" {

/* ---- RNG: Public.cpp:10-23 ---- */
void ref_seed(int seed) { randinitalize(seed); }
float ref_randfloat(void) { return randfloat(); }
int ref_randint(int a, int b) { return randint(a, b); }
int ref_cvround(double v) { return cvRound(v); }
void ref_set_log(int verbose)
{
    g_verboseMode = verbose != 0;
    g_detailedLog = false;
    if (!g_logFile.is_open()) g_logFile.open("/dev/null");
    /* the reference also prints progress straight to std::cout (selector lists, foot points, ...) */
    if (verbose) std::cout.clear(); else std::cout.setstate(std::ios_base::failbit);
}

/* ---- frames: Matrixu planes + initII (Matrix.cpp:33-47, Camera.cpp:449-494) ---- */
void* ref_frame_create(const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, int W, int H, int pitch)
{
    Frame* f = new Frame();
    f->W = W; f->H = H;
    f->gray.Resize(H, W, 1);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) f->gray(y, x) = gray[(size_t)y * pitch + x];
    f->gray.initII();
    if (c0) {
        f->color.Resize(H, W, 3);
        f->hsv.Resize(H, W, 3);
        const uint8_t* pl[3] = {c0, c1, c2};
        for (int k = 0; k < 3; k++)
            for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
                f->color(y, x, k) = pl[k][(size_t)y * pitch + x];
                f->hsv(y, x, k) = pl[k][(size_t)y * pitch + x];
            }
    }
    return f;
}
void ref_frame_free(void* fr) { delete (Frame*)fr; }
void ref_integral(void* fr, float* out)
{
    Frame* f = (Frame*)fr;
    for (int y = 0; y <= f->H; y++) for (int x = 0; x <= f->W; x++) out[(size_t)y * (f->W + 1) + x] = f->gray.ii(y, x, 0);
}
float ref_sum_rect(void* fr, int x, int y, int w, int h)
{
    IppiRect r; r.x = x; r.y = y; r.width = w; r.height = h;
    return ((Frame*)fr)->gray.sumRect(r, 0);
}
/* Matrix.cpp:70-100, 518-580: packed BGR IplImage -> planes, conv2BW, conv2HSV (bug included: writes into *this) */
void ref_bgr_to_planes(const uint8_t* bgr, int W, int H, int pitch_bytes, uint8_t* gray, uint8_t* rgb3 /* [3][H][W] */)
{
    IplImage hdr;
    cvInitImageHeader(&hdr, cvSize(W, H), IPL_DEPTH_8U, 3);
    hdr.imageData = (char*)bgr; hdr.widthStep = pitch_bytes;
    Matrixu color(H, W, 3), bw;
    color.IplImage2Matrix(&hdr);
    color.conv2BW(bw);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        gray[(size_t)y * W + x] = bw(y, x);
        for (int k = 0; k < 3; k++) rgb3[((size_t)k * H + y) * W + x] = color(y, x, k);
    }
}

/* ---- strong classifier through the reference's factory (StrongClassifierFactory.cpp:14-43) ---- */
void* ref_clf_create(int feature_type, int n_haar, int nb, int w0, int h0, int weak_type, int strong_type, int K, float lr,
                     int use_hsv, float pct_retained)
{
    Clf* c = new Clf();
    c->params = make_clf_params(feature_type, n_haar, nb, weak_type, strong_type, K, lr, use_hsv, pct_retained);
    c->params->m_featureParametersPtr->m_width = w0;
    c->params->m_featureParametersPtr->m_height = h0;
    c->w0 = w0; c->h0 = h0;
    c->clf = Classifier::StrongClassifierFactory::CreateAndInitializeClassifier(c->params);
    return c;
}
void ref_clf_free(void* c) { delete (Clf*)c; }
int ref_clf_num_features(void* c) { return ((Clf*)c)->clf->GetNumberOfFeatures(); }
int ref_clf_num_selected(void* c) { return ((Clf*)c)->params->m_numberOfSelectedWeakClassifiers; }
/* the Haar table HaarFeature::Generate drew (HaarFeature.cpp:25-69) */
int ref_clf_haar_table(void* cp, int* nrect, int* rect /* [F][6][4] */, float* weight /* [F][6] */)
{
    Clf* c = (Clf*)cp;
    Features::HaarFeatureVector* hv = dynamic_cast<Features::HaarFeatureVector*>(c->clf->m_featureVectorPtr.get());
    if (!hv) return 0;
    int F = (int)hv->m_featureList.size();
    for (int f = 0; f < F; f++) {
        Features::HaarFeature* h = dynamic_cast<Features::HaarFeature*>(hv->m_featureList[f].get());
        nrect[f] = (int)h->m_rects.size();
        for (int k = 0; k < nrect[f]; k++) {
            rect[(f * 6 + k) * 4 + 0] = h->m_rects[k].x; rect[(f * 6 + k) * 4 + 1] = h->m_rects[k].y;
            rect[(f * 6 + k) * 4 + 2] = h->m_rects[k].width; rect[(f * 6 + k) * 4 + 3] = h->m_rects[k].height;
            weight[f * 6 + k] = h->m_weights[k];
        }
    }
    return F;
}
/* FeatureVector::Compute -> [F][n] */
void ref_clf_features(void* cp, void* fr, int n, const int* col, const int* row, const float* sx, const float* sy, float* out, int ld)
{
    Clf* c = (Clf*)cp;
    Classifier::SampleSet set;
    fill_set(set, (Frame*)fr, n, col, row, sx, sy, c->w0, c->h0);
    c->clf->m_featureVectorPtr->Compute(set);
    int F = c->clf->GetNumberOfFeatures();
    for (int f = 0; f < F; f++) for (int i = 0; i < n; i++) out[(size_t)f * ld + i] = set.GetFeatureValue(i, f);
}
int ref_clf_update(void* cp, void* fr, int np, const int* pcol, const int* prow, const float* psx, const float* psy, int nn,
                    const int* ncol, const int* nrow, const float* nsx, const float* nsy)
{
    Clf* c = (Clf*)cp;
    Classifier::SampleSet pos, neg;
    fill_set(pos, (Frame*)fr, np, pcol, prow, psx, psy, c->w0, c->h0);
    fill_set(neg, (Frame*)fr, nn, ncol, nrow, nsx, nsy, c->w0, c->h0);
    REF_TRY(c->clf->Update(pos, neg), -1);
    return 0;
}
void ref_clf_classify(void* cp, void* fr, int n, const int* col, const int* row, const float* sx, const float* sy, int log_ratio, float* out)
{
    Clf* c = (Clf*)cp;
    Classifier::SampleSet set;
    fill_set(set, (Frame*)fr, n, col, row, sx, sy, c->w0, c->h0);
    vectorf r = c->clf->Classify(set, log_ratio != 0);
    for (int i = 0; i < n; i++) out[i] = r[i];
}
int ref_clf_get_selectors(void* cp, int* idx)
{
    Clf* c = (Clf*)cp;
    int n = (int)c->clf->m_selectorList.size();
    for (int i = 0; i < n; i++) idx[i] = c->clf->m_selectorList[i];
    return n;
}
int ref_clf_get_alphas(void* cp, float* a)
{
    Classifier::AdaBoostClassifier* ab = dynamic_cast<Classifier::AdaBoostClassifier*>(((Clf*)cp)->clf.get());
    if (!ab) return 0;
    for (size_t i = 0; i < ab->m_alphaList.size(); i++) a[i] = ab->m_alphaList[i];
    return (int)ab->m_alphaList.size();
}
static void weak_state(Classifier::WeakClassifierBase* w, float* out, int F, int f)
{
    if (Classifier::OnlineStumpsWeakClassifier* s = dynamic_cast<Classifier::OnlineStumpsWeakClassifier*>(w)) {
        float v[8] = {s->m_mu0, s->m_mu1, s->m_sig0, s->m_sig1, s->m_n0, s->m_n1, s->m_e0, s->m_e1};
        for (int k = 0; k < 8; k++) out[(size_t)k * F + f] = v[k];
    } else if (Classifier::WeightedStumpsWeakClassifier* s = dynamic_cast<Classifier::WeightedStumpsWeakClassifier*>(w)) {
        float v[8] = {s->m_mu0, s->m_mu1, s->m_sig0, s->m_sig1, s->m_n0, s->m_n1, s->m_e0, s->m_e1};
        for (int k = 0; k < 8; k++) out[(size_t)k * F + f] = v[k];
    } else if (Classifier::PerceptronWeakClassifier* p = dynamic_cast<Classifier::PerceptronWeakClassifier*>(w)) {
        for (int k = 0; k < 8; k++) out[(size_t)k * F + f] = 0;
        for (size_t k = 0; k < p->m_weightList.size() && k < 8; k++) out[(size_t)k * F + f] = p->m_weightList[k];
    }
}
void ref_clf_get_weak_state(void* cp, float* out8xF)
{
    Clf* c = (Clf*)cp;
    int F = (int)c->clf->m_weakClassifierPtrList.size();
    for (int f = 0; f < F; f++) weak_state(c->clf->m_weakClassifierPtrList[f], out8xF, F, f);
}

/* ---- one weak classifier on given feature values, explicit weights allowed (a11-a13) ---- */
/* n_updates successive Update calls on the same rows; then ClassifyF on xtest.  wpos/wneg may be NULL for STUMP / PERCEPTRON. */
void ref_weak_run(int weak_type, float lr, int n_updates, const float* xpos, int P, const float* xneg, int Nn, const float* wpos,
                  const float* wneg, float* state8, const float* xtest, int nt, float* out, int* is_valid)
{
    Classifier::SampleSet pos, neg, tst;
    pos.Resize(P); neg.Resize(Nn); tst.Resize(nt);
    pos.ResizeFeatures(1); neg.ResizeFeatures(1); tst.ResizeFeatures(1);
    for (int i = 0; i < P; i++) pos.GetFeatureValue(i, 0) = xpos[i];
    for (int i = 0; i < Nn; i++) neg.GetFeatureValue(i, 0) = xneg[i];
    for (int i = 0; i < nt; i++) tst.GetFeatureValue(i, 0) = xtest[i];
    Classifier::WeakClassifierBase* w;
    if (weak_type == Classifier::STUMP) w = new Classifier::OnlineStumpsWeakClassifier(0);
    else if (weak_type == Classifier::WEIGHTED_STUMP) w = new Classifier::WeightedStumpsWeakClassifier(0);
    else w = new Classifier::PerceptronWeakClassifier(0);
    w->SetLearningRate(lr);
    vectorf vp, vn;
    if (wpos) vp.assign(wpos, wpos + P);
    if (wneg) vn.assign(wneg, wneg + Nn);
    for (int u = 0; u < n_updates; u++) w->Update(pos, neg, wpos ? &vp : NULL, wneg ? &vn : NULL);
    weak_state(w, state8, 1, 0);
    for (int i = 0; i < nt; i++) out[i] = w->ClassifyF(tst, i);
    if (is_valid) *is_valid = w->IsValidWeakClassifier() ? 1 : 0;
    delete w;
}

/* ---- samplers: SampleSet.cpp:82-361 ---- */
int ref_sample_ring(void* fr, int x, int y, int w, int h, float outer_r, float inner_r, int max_n, float sx, float sy, int* col, int* row, int cap)
{
    Frame* f = (Frame*)fr;
    Classifier::SampleSet s;
    s.SampleImage(&f->gray, x, y, w, h, outer_r, inner_r, max_n, &f->color, &f->hsv, sx, sy);
    int n = (int)s.Size();
    for (int i = 0; i < n && i < cap; i++) { col[i] = s[i].m_col; row[i] = s[i].m_row; }
    return n;
}
int ref_sample_rect(void* fr, int x, int y, int w, int h, float maxdx, float maxdy, float mindx, float mindy, int max_n, float sx,
                    float sy, int* col, int* row, int cap)
{
    Frame* f = (Frame*)fr;
    Classifier::SampleSet s;
    s.SampleImage(&f->gray, x, y, w, h, maxdx, maxdy, mindx, mindy, max_n, &f->color, &f->hsv, sx, sy);
    int n = (int)s.Size();
    for (int i = 0; i < n && i < cap; i++) { col[i] = s[i].m_col; row[i] = s[i].m_row; }
    return n;
}

/* ---- particle filter: ParticleFilter.cpp ---- */
void* ref_pf_create(int N, float img_w, float img_h, float cx, float cy, float sx, float sy, float max_scale_change, float max_scale, float min_scale)
{
    ParticleFilter* p = new ParticleFilter(img_w, img_h);
    p->Initialize(N, cx, cy, sx, sy, max_scale_change, max_scale, min_scale);
    return p;
}
void ref_pf_free(void* p) { delete (ParticleFilter*)p; }
static void copy_mat(const Matrixf& m, int N, float* out) { for (int r = 0; r < N; r++) for (int c = 0; c < 5; c++) out[r * 5 + c] = m(r, c); }
void ref_pf_get(void* pp, float* state, float* resampled, int* idx, int* filter_state)
{
    ParticleFilter* p = (ParticleFilter*)pp;
    int N = p->m_numberOfParticles;
    if (state) copy_mat(p->m_particlesStateMatrix, N, state);
    if (resampled) copy_mat(p->m_particlesResampledMatrix, N, resampled);
    if (idx) for (int i = 0; i < N && i < (int)p->m_resampleIndexList.size(); i++) idx[i] = p->m_resampleIndexList[i];
    if (filter_state) *filter_state = (int)p->m_filterState;
}
void ref_pf_set(void* pp, const float* state, const float* resampled, int filter_state)
{
    ParticleFilter* p = (ParticleFilter*)pp;
    int N = p->m_numberOfParticles;
    for (int r = 0; r < N; r++) for (int c = 0; c < 5; c++) {
        if (state) p->m_particlesStateMatrix(r, c) = state[r * 5 + c];
        if (resampled) p->m_particlesResampledMatrix(r, c) = resampled[r * 5 + c];
    }
    if (filter_state >= 0) p->m_filterState = (ParticleFilter::State)filter_state;
    p->m_isOrderedUniqueParticlesFound = false;
}
/* noise rows [4][N] are consumed by the four cv::randn calls of PredictWithBrownianMotion (:248-301) */
void ref_pf_predict(void* pp, const float* noise4xN, float sdx, float sdy, float sdsx, float sdsy, float w0, float h0)
{
    ParticleFilter* p = (ParticleFilter*)pp;
    mct_shim_set_noise(noise4xN, (size_t)4 * p->m_numberOfParticles);
    p->PredictWithBrownianMotion(sdx, sdy, sdsx, sdsy, w0, h0);
    mct_shim_set_noise(NULL, 0);
}
void ref_pf_update_all_weights(void* pp, const float* prob, int n, int only_nonzero, int force_reset, int resample)
{
    vectorf v(prob, prob + n);
    ((ParticleFilter*)pp)->UpdateAllParticlesWeight(v, only_nonzero != 0, force_reset != 0, resample != 0);
}
void ref_pf_update_particle_weight(void* pp, int i, float prob, int force_reset) { ((ParticleFilter*)pp)->UpdateParticleWeight(i, prob, force_reset != 0); }
void ref_pf_normalize(void* pp) { ((ParticleFilter*)pp)->NormalizeParticleWeights(); }
void ref_pf_resample(void* pp, int force) { ((ParticleFilter*)pp)->ResampleParticles(force != 0); }
void ref_pf_resample_if_required(void* pp) { ((ParticleFilter*)pp)->ResampleIfRequired(); }
void ref_pf_get_average(void* pp, float* out4) { vectorf v; ((ParticleFilter*)pp)->GetAverageofAllParticles(v); for (int i = 0; i < 4; i++) out4[i] = v[i]; }
void ref_pf_get_highest_weight(void* pp, float* out5) { vectorf v; ((ParticleFilter*)pp)->GetHighestWeightParticle(v); for (int i = 0; i < 5; i++) out5[i] = v[i]; }
int ref_pf_ordered_unique(void* pp, float* out, int cap)
{
    ParticleFilter* p = (ParticleFilter*)pp;
    int n = p->GetNumberOfOrderedUniqueParticles();
    for (int i = 0; i < n && i < cap; i++) { vectorf v; p->GetOrderedUniqueParticles(i, v); for (int k = 0; k < 5; k++) out[i * 5 + k] = v[k]; }
    return n;
}
void ref_pf_get_highest_unique_close(void* pp, const float* closest2, float* out5)
{
    vectorf v, c(closest2, closest2 + 2);
    ((ParticleFilter*)pp)->GetHighestOrderedUniqueParticleCloseToTheGivenState(v, c);
    for (int i = 0; i < 5; i++) out5[i] = v[i];
}
void ref_pf_rearrange(void* pp, const float* mean_foot2, int change_scale, float w0, float h0)
{
    vectorf m(mean_foot2, mean_foot2 + 2), var(2, 0.0f);
    ((ParticleFilter*)pp)->RearrangeParticlesBasedOnGroundLocation(m, var, change_scale != 0, w0, h0);
}

/* ---- ParticleFilterTracker (ParticleFilterTracker.cpp) ---- */
typedef struct {
    int n_particles;
    float sd_x, sd_y, sd_sx, sd_sy;
    float pos_radius_train, init_pos_radius_train;
    int n_neg_train, init_n_neg_train;
    int max_pos_train;
    int search_win;
    int trajectory_option;
    int pos_strategy, neg_strategy;
    int max_num_pos_examples;
} ref_tracker_params;           /* same layout as orc_tracker_params */

struct Trk {
    ParticleFilterTrackerPtr pft;
    SimpleTrackerPtr simple;
    Tracker* base;
    int n_frames, frame;
};
static void fill_tp(ParticleFilterTrackerParameters* p, const ref_tracker_params* tp, const float* init_xywh)
{
    p->m_posRadiusTrain = tp->pos_radius_train;
    p->m_numberOfNegativeTrainingSamples = tp->n_neg_train;
    p->m_init_negNumTrain = tp->init_n_neg_train;
    p->m_init_posTrainRadius = tp->init_pos_radius_train;
    p->m_maximumNumberOfPositiveTrainingSamples = tp->max_pos_train;
    p->m_initState.resize(5);
    for (int i = 0; i < 4; i++) p->m_initState[i] = init_xywh[i];
    p->m_initState[4] = 0;
    p->m_initializeWithFaceDetection = false;
    p->m_debugv = false;
    p->m_isColor = true;
    p->m_displayTrainingSampleCenterOnly = false;
    p->m_searchWindSize = tp->search_win;
    p->m_negSampleStrategy = 1;
    p->m_shouldNotUseSigmoid = DEFAULT_PFTRACKER_NOT_USE_SIGMOIDAL;
    p->m_numberOfParticles = tp->n_particles;
    p->m_standardDeviationX = tp->sd_x; p->m_standardDeviationY = tp->sd_y;
    p->m_standardDeviationScaleX = tp->sd_sx; p->m_standardDeviationScaleY = tp->sd_sy;
    p->m_maxNumPositiveExamples = tp->max_num_pos_examples;
    p->m_numOfDisplayedParticles = 0;
    p->m_positiveSampleStrategy = (PFTracker_Positive_Sample_Strategy)tp->pos_strategy;
    p->m_negativeSampleStrategy = (PFTracker_Negative_Sample_Strategy)tp->neg_strategy;
    p->m_outputTrajectoryOption = (PFTracker_TrajectoryOption)tp->trajectory_option;
}
/* particle_filter = 1: ParticleFilterTracker, 0: SimpleTracker.  The strong classifier is created INSIDE InitializeTracker
 * (ParticleFilterTracker.cpp:128), i.e. the Haar table is drawn from the global stream right before the first samplers run. */
void* ref_tracker_create(int particle_filter, const ref_tracker_params* tp, int feature_type, int n_haar, int nb, int weak_type,
                         int strong_type, int K, float lr, int use_hsv, float pct_retained, void* frame0, const float* init_xywh,
                         int n_frames)
{
    ref_set_log(0);
    Frame* f = (Frame*)frame0;
    Trk* t = new Trk();
    t->n_frames = n_frames; t->frame = 0;
    Classifier::StrongClassifierParametersBasePtr cp = make_clf_params(feature_type, n_haar, nb, weak_type, strong_type, K, lr, use_hsv, pct_retained);
    ParticleFilterTrackerParametersPtr p(new ParticleFilterTrackerParameters());
    fill_tp(p.get(), tp, init_xywh);
    Matrixu* color = f->color.rows() ? &f->color : NULL;
    Matrixu* hsv = f->hsv.rows() ? &f->hsv : NULL;
    if (particle_filter) {
        t->pft.reset(new ParticleFilterTracker(NULL, AppearanceBasedInformationFuserPtr(), false));
        t->base = t->pft.get();
    } else {
        t->simple.reset(new SimpleTracker(NULL));
        t->base = t->simple.get();
    }
    REF_TRY(t->base->InitializeTrackerWithParameters(color, &f->gray, 0, n_frames, p, cp, NULL, NULL, hsv, NULL), (delete t, (void*)NULL));
    return t;
}
void ref_tracker_free(void* t) { delete (Trk*)t; }
/* phases bit0: TrackObjectOnTheGivenFrame + StoreObjectState; bit1: UpdateClassifier (TrackObjectAndSaveState :239-258) */
static void tracker_track_body(void* tp, void* fr, const float* noise4xN, int out_box[4], int phases);
int ref_tracker_track(void* tp, void* fr, const float* noise4xN, int out_box[4], int phases)
{
    REF_TRY(tracker_track_body(tp, fr, noise4xN, out_box, phases), -1);
    return 0;
}
static void tracker_track_body(void* tp, void* fr, const float* noise4xN, int out_box[4], int phases)
{
    Trk* t = (Trk*)tp; Frame* f = (Frame*)fr;
    Matrixu* color = f->color.rows() ? &f->color : NULL;
    Matrixu* hsv = f->hsv.rows() ? &f->hsv : NULL;
    t->frame++;
    if (t->pft) {
        int N = t->pft->m_particleFilterTrackerParamsPtr->m_numberOfParticles;
        mct_shim_set_noise(noise4xN, (size_t)4 * N);
        if (phases & 1) {
            t->pft->TrackObjectWithoutSaveState(color, &f->gray, hsv);
            t->pft->StoreObjectState(t->frame, NULL);
        }
        if (phases & 2) t->pft->UpdateClassifier(color, &f->gray, NULL, hsv);
        mct_shim_set_noise(NULL, 0);
        for (int k = 0; k < 4; k++) out_box[k] = (int)t->pft->m_states(t->frame, k);
    } else {
        t->simple->TrackObjectAndSaveState(t->frame, color, &f->gray, NULL, NULL, hsv);
        for (int k = 0; k < 4; k++) out_box[k] = (int)t->simple->m_states(t->frame, k);
    }
}
void ref_tracker_state(void* tp, float* out6)
{
    Trk* t = (Trk*)tp;
    const vectorf& s = t->base->GetCurrentTrackerState();
    for (int i = 0; i < 6; i++) out6[i] = s[i];
}
void ref_tracker_states_row(void* tp, int frame, float* out4)
{
    Trk* t = (Trk*)tp;
    SimpleTracker* s = t->pft ? (SimpleTracker*)t->pft.get() : t->simple.get();
    for (int k = 0; k < 4; k++) out4[k] = s->m_states(frame, k);
}
void* ref_tracker_pf(void* tp) { return ((Trk*)tp)->pft ? ((Trk*)tp)->pft->m_particleFilterPtr.get() : NULL; }
int ref_tracker_selectors(void* tp, int* idx)
{
    Trk* t = (Trk*)tp;
    SimpleTracker* s = t->pft ? (SimpleTracker*)t->pft.get() : t->simple.get();
    int n = (int)s->m_strongClassifierBasePtr->m_selectorList.size();
    for (int i = 0; i < n; i++) idx[i] = s->m_strongClassifierBasePtr->m_selectorList[i];
    return n;
}
void ref_tracker_weak_state(void* tp, float* out8xF)
{
    Trk* t = (Trk*)tp;
    SimpleTracker* s = t->pft ? (SimpleTracker*)t->pft.get() : t->simple.get();
    int F = (int)s->m_strongClassifierBasePtr->m_weakClassifierPtrList.size();
    for (int f = 0; f < F; f++) weak_state(s->m_strongClassifierBasePtr->m_weakClassifierPtrList[f], out8xF, F, f);
}
int ref_tracker_test_set_size(void* tp) { Trk* t = (Trk*)tp; return t->pft ? (int)t->pft->m_testSampleSet.Size() : (int)t->simple->m_testSampleSet.Size(); }

}  // extern "C"

/* ============================================================================================================
 * Multi-camera schedule: CameraNetwork::TrackObjectsOnCurrentFrame (CameraNetwork.cpp:537-613) over the reference's own
 * Object / ParticleFilterTracker / GeometryBasedInformationFuser / AppearanceBasedInformationFuser classes.
 *
 * Why not the reference's CameraNetwork / Camera objects themselves: Camera reads videos, homographies and ground-truth
 * files from disk inside its constructor (Camera.cpp:90-190), and the per-pair streams need a hook between objects.  The loops below are those of CameraNetwork.cpp / Camera.cpp line by line (cited), everything inside
 * them is the reference's code.  Each (camera, object) pair draws from its own MWC stream (mct_shim_rng_override), the
 * convention of the sharded B200 path; with every pair seeded to share ONE stream object the draws are the reference's
 * global sequence.
 * ============================================================================================================ */
typedef struct {
    int n_cameras, n_objects, n_frames;
    int feature_type, n_haar, n_bins, use_hsv, strong_type, weak_type, pct_selected, pct_retained;
    int n_particles;
    float sd_x, sd_y, sd_sx, sd_sy;
    int pos_radius_train, init_pos_radius_train, n_neg_train, init_n_neg_train, search_win, neg_sample_strategy;
    int trajectory_option, pos_strategy, neg_strategy, max_num_pos_examples;
    int geometric_fusion;                 /* 0 / 1 */
    int appearance_fusion;                /* 0 none, 1 culture colour, 2 MDCH */
    int af_strong, af_weak, af_pct_selected, af_pct_retained, af_refresh;
    int shared_stream;                    /* 1: all pairs draw from stream [0][0] (the reference's single global stream) */
} ref_net_cfg;

namespace {
struct Net;
struct NetBase : public CameraNetworkBase {
    Net* net;
    virtual void GenerateTrainingSampleSetsForAppearanceFusion(const int objectId, Classifier::SampleSet& pos, Classifier::SampleSet& neg);
    virtual void TrackObjectsOnCurrentFrame(const int) {}
    virtual void SaveCameraNetworkState() {}
};
struct Net {
    ref_net_cfg cfg;
    CameraTrackingParametersPtr ctp;
    std::vector<CvMat*> H;
    std::vector<std::vector<ObjectPtr> > obj;          /* [camera][object] */
    std::vector<Frame*> frames;                        /* current frame of every camera */
    std::vector<GeometryBasedInformationFuserPtr> gf;  /* per object */
    std::vector<uint64_t> streams;                     /* [camera][object] */
    std::vector<CvMat*> ground;                        /* m_groundPlaneParticlesPtrList */
    NetBase* base;
    CameraNetworkBasePtr basePtr;
    uint64_t* stream(int c, int o) { return cfg.shared_stream ? &streams[0] : &streams[(size_t)c * cfg.n_objects + o]; }
    void select(int c, int o) { mct_shim_rng_override = stream(c, o); }
    Matrixu* color(int c) { return frames[c]->color.rows() ? &frames[c]->color : NULL; }
    Matrixu* gray(int c) {
        /* Camera::PrepareCurrentFrameForTracking (Camera.cpp:449-494): a gray image exists only for the Haar feature types */
        return (cfg.feature_type == Features::HAAR_LIKE || cfg.feature_type == Features::HAAR_COLOR_HISTOGRAM) ? &frames[c]->gray : NULL;
    }
    Matrixu* hsv(int c) { return ctp->m_HSVRequired ? &frames[c]->hsv : NULL; }
};
/* CameraNetwork.cpp:276-320 */
void NetBase::GenerateTrainingSampleSetsForAppearanceFusion(const int objectId, Classifier::SampleSet& positiveSampleSet, Classifier::SampleSet& negativeSampleSet)
{
    uint64_t* saved = mct_shim_rng_override;
    for (int cameraIndex = 0; cameraIndex < net->cfg.n_cameras; cameraIndex++) {
        Classifier::SampleSet objectPositiveSampleSet, objectNegativeSampleSet;
        net->select(cameraIndex, objectId);
        net->obj[cameraIndex][objectId]->GenerateTrainingSampleSetsForAppearanceFusion(objectPositiveSampleSet, objectNegativeSampleSet,
                                                                                     net->color(cameraIndex), net->gray(cameraIndex), net->hsv(cameraIndex));
        for (int i = 0; i < (int)objectPositiveSampleSet.Size(); i++) { objectPositiveSampleSet[i].m_cameraID = cameraIndex; positiveSampleSet.PushBackSample(objectPositiveSampleSet[i]); }
        for (int i = 0; i < (int)objectNegativeSampleSet.Size(); i++) negativeSampleSet.PushBackSample(objectNegativeSampleSet[i]);
    }
    mct_shim_rng_override = saved;
}
}  // namespace

extern "C" {
static void* net_create_body(const ref_net_cfg* cfg, const double* homographies, const float* init_xywh, const int64_t* seeds);
void* ref_net_create(const ref_net_cfg* cfg, const double* homographies /* [C][9] */, const float* init_xywh /* [C][O][4] */,
                     const int64_t* seeds /* [C][O] */)
{
    REF_TRY(return net_create_body(cfg, homographies, init_xywh, seeds), (void*)NULL);
}
static void* net_create_body(const ref_net_cfg* cfg, const double* homographies, const float* init_xywh, const int64_t* seeds)
{
    ref_set_log(0);
    Net* n = new Net();
    n->cfg = *cfg;
    InputParameters& g = g_configInput;
    memset(&g, 0, sizeof(g));
    g.m_numOfFrames = cfg->n_frames; g.m_startFrameIndex = 0; g.m_trialNumber = 1;
    g.m_loadVideoWithColor = 1; g.m_loadVideoFromImgs = 1;
    strcpy(g.m_outputDirectoryNameCstr, "/tmp"); strcpy(g.m_inputDirectoryNameCstr, "/tmp");
    strcpy(g.m_dataFilesNameCstr, "mct"); strcpy(g.m_intializationDirectoryCstr, "init");
    g.m_trackerFeatureType = cfg->feature_type + 1; g.m_useHSVColor = cfg->use_hsv; g.m_numofBinsColor = cfg->n_bins;
    g.m_trackerFeatureParameter = cfg->n_haar;
    g.m_percentageOfWeakClassifiersSelected = cfg->pct_selected; g.m_percentageOfWeakClassifiersRetained = cfg->pct_retained;
    g.m_localTrackerType = 1;
    g.m_posRadiusTrain = cfg->pos_radius_train; g.m_initPosRadiusTrain = cfg->init_pos_radius_train;
    g.m_numNegExamples = cfg->n_neg_train; g.m_initNumNegExampes = cfg->init_n_neg_train;
    g.m_searchWindowSize = cfg->search_win; g.m_negSampleStrategy = cfg->neg_sample_strategy;
    g.m_numOfParticles = cfg->n_particles;
    g.m_PFTrackerStdDevX = cfg->sd_x; g.m_PFTrackerStdDevY = cfg->sd_y; g.m_PFTrackerStdDevScaleX = cfg->sd_sx; g.m_PFTrackerStdDevScaleY = cfg->sd_sy;
    g.m_PFTrackerNumDispParticles = 0; g.m_PfTrackerMaxNumPositiveExamples = cfg->max_num_pos_examples;
    g.m_PFOutputTrajectoryOption = cfg->trajectory_option;
    g.m_PfTrackerPositiveExampleStrategy = cfg->pos_strategy; g.m_PfTrackerNegativeExampleStrategy = cfg->neg_strategy;
    g.m_geometricFusionType = cfg->geometric_fusion; g.m_appearanceFusionType = cfg->appearance_fusion;
    g.m_appearanceFusionStrongClassifierType = cfg->af_strong; g.m_appearanceFusionWeakClassifierType = cfg->af_weak + 1;
    g.m_AFRefreshRate = cfg->af_refresh > 0 ? cfg->af_refresh : 1;
    g.m_AFpercentageOfWeakClassifiersSelected = cfg->af_pct_selected; g.m_AFpercentageOfWeakClassifiersRetained = cfg->af_pct_retained;
    n->ctp.reset(new CameraTrackingParameters(&g, PARTICLE_FILTER_TRACKER, (Classifier::StrongClassifierType)cfg->strong_type,
                                              (Classifier::WeakClassifierType)cfg->weak_type, (Features::FeatureType)cfg->feature_type,
                                              (GeometricFusionType)cfg->geometric_fusion, cfg->n_particles,
                                              (AppearanceFusionType)cfg->appearance_fusion,
                                              (AppearanceFusionStrongClassifierType)cfg->af_strong,
                                              (Classifier::WeakClassifierType)cfg->af_weak));
    n->streams.resize((size_t)cfg->n_cameras * cfg->n_objects);
    n->obj.resize(cfg->n_cameras);
    n->frames.assign(cfg->n_cameras, (Frame*)NULL);
    for (int c = 0; c < cfg->n_cameras; c++) {
        CvMat* H = cvCreateMat(3, 3, CV_64FC1);
        for (int i = 0; i < 9; i++) H->data.db[i] = homographies ? homographies[c * 9 + i] : (i % 4 == 0 ? 1.0 : 0.0);
        n->H.push_back(H);
        for (int o = 0; o < cfg->n_objects; o++) {
            int64_t seed = seeds[(size_t)c * cfg->n_objects + o];
            n->streams[(size_t)c * cfg->n_objects + o] = seed ? (uint64_t)seed : (uint64_t)(int64_t)-1;
            /* Camera::InitializeCameraParameters (Camera.cpp:139-150, 251-270): objects with their initial state */
            ObjectPtr op(new Object(o, c, true, n->ctp, H));
            vectorf st(5);
            for (int k = 0; k < 4; k++) st[k] = init_xywh[((size_t)c * cfg->n_objects + o) * 4 + k];
            st[4] = 0;
            n->select(c, o);
            op->InitializeObjectParameters(st, "/tmp/mct_ref");      /* creates the appearance fuser's classifier too (Object.cpp:203-216) */
            n->obj[c].push_back(op);
        }
    }
    n->base = new NetBase(); n->base->net = n;
    n->basePtr = n->base;                                  /* CameraNetworkBasePtr is a raw pointer (CameraNetworkBase.h) */
    mct_shim_rng_override = 0;
    return n;
}
void ref_net_free(void* np)
{
    Net* n = (Net*)np;
    if (!n) return;
    mct_shim_rng_override = 0;
    delete n->base;
    delete n;
}
/* frame 0: Camera::InitializeCameraTrackers (Camera.cpp:278-321) for every camera in order, then CameraNetwork::InitializeGeometricFuser (:380-470) */
static void net_init_body(void* np, void* const* frames);
int ref_net_init(void* np, void* const* frames)
{
    REF_TRY(net_init_body(np, frames), -1);
    return 0;
}
static void net_init_body(void* np, void* const* frames)
{
    Net* n = (Net*)np;
    const ref_net_cfg& cfg = n->cfg;
    for (int c = 0; c < cfg.n_cameras; c++) {
        n->frames[c] = (Frame*)frames[c];
        for (int o = 0; o < cfg.n_objects; o++) {
            n->select(c, o);
            n->obj[c][o]->InitializeObjectTracker(n->color(c), n->gray(c), 0, cfg.n_frames, NULL, NULL, n->hsv(c), NULL);
        }
    }
    if (cfg.geometric_fusion) {
        n->gf.resize(cfg.n_objects); n->ground.assign(cfg.n_objects, (CvMat*)NULL);
        for (int o = 0; o < cfg.n_objects; o++) {
            n->ground[o] = cvCreateMat(cfg.n_cameras, 2, CV_32FC1);
            for (int c = 0; c < cfg.n_cameras; c++) {
                CvMat* p = n->obj[c][o]->GetAverageParticleFootPositionOnImagePlaneForGeometricFusion();
                GeometryBasedInformationFuser::CopyMatrix(p, n->ground[o], 0, c, 2);
            }
            n->gf[o].reset(new GeometryBasedInformationFuser(cfg.n_cameras, n->ground[o], GeometryBasedInformationFuser::PRINCIPAL_AXIS_INTERSECTION, n->H));
        }
    }
    mct_shim_rng_override = 0;
}
/* one frame.  noise [C][O][4][N]: the cv::randn rows of each tracker's PredictWithBrownianMotion */
static void net_track_body(void* np, int frameIndex, void* const* frames, const float* noise);
int ref_net_track(void* np, int frameIndex, void* const* frames, const float* noise)
{
    REF_TRY(net_track_body(np, frameIndex, frames, noise), -1);
    return 0;
}
static void net_track_body(void* np, int frameIndex, void* const* frames, const float* noise)
{
    Net* n = (Net*)np;
    const ref_net_cfg& cfg = n->cfg;
    const size_t per = (size_t)4 * cfg.n_particles;
    /* CameraNetwork.cpp:549-554 -> Camera::TrackCameraFrame (Camera.cpp:334-397) -> Object::TrackObjectFrame (Object.cpp:411-455) */
    for (int c = 0; c < cfg.n_cameras; c++) {
        n->frames[c] = (Frame*)frames[c];
        for (int o = 0; o < cfg.n_objects; o++) {
            n->select(c, o);
            mct_shim_set_noise(noise + ((size_t)c * cfg.n_objects + o) * per, per);
            n->obj[c][o]->TrackObjectFrame(frameIndex, n->color(c), n->gray(c), NULL, NULL, n->hsv(c));
            mct_shim_set_noise(NULL, 0);
        }
    }
    if (!cfg.geometric_fusion && !cfg.appearance_fusion) { mct_shim_rng_override = 0; return; }                 /* :557-561 */
    if (cfg.geometric_fusion) {
        /* FuseGeometrically (:175-268), principal-axis branch */
        for (int o = 0; o < cfg.n_objects; o++) {
            n->ground[o] = cvCreateMat(cfg.n_cameras, 2, CV_32FC1);                                               /* (:191, leaked there too) */
            for (int c = 0; c < cfg.n_cameras; c++) {
                CvMat* p = n->obj[c][o]->GetAverageParticleFootPositionOnImagePlaneForGeometricFusion();
                GeometryBasedInformationFuser::CopyMatrix(p, n->ground[o], 0, c, 2);
            }
        }
        for (int o = 0; o < cfg.n_objects; o++) n->gf[o]->FuseInformation(n->ground[o]);
        if (frameIndex > 0) {
            /* FeedbackInformationFromGeometricFusion (:94-128) */
            for (int c = 0; c < cfg.n_cameras; c++)
                for (int o = 0; o < cfg.n_objects; o++) {
                    CvMat* mean = n->gf[o]->GetGroundPlaneKalmanMeanMatrix(frameIndex);
                    CvMat* cov = n->gf[o]->GetGroundPlaneKalmanCovarianceMatrix();
                    n->select(c, o);
                    n->obj[c][o]->UpdateParticleWeightsWithGroundPDF(mean, cov, n->color(c), n->gray(c), n->hsv(c));
                    cvReleaseMat(&mean); cvReleaseMat(&cov);
                }
        }
    }
    /* :583-587 -> Camera::LearnLocalAppearanceModel (Camera.cpp:610-640) */
    for (int c = 0; c < cfg.n_cameras; c++)
        for (int o = 0; o < cfg.n_objects; o++) {
            n->select(c, o);
            n->obj[c][o]->UpdateParticleFilterTrackerAppearanceModel(n->color(c), n->gray(c), NULL, n->hsv(c));
        }
    /* :589-600 -> Camera::LearnGlobalAppearanceModel (Camera.cpp:722-740) */
    if (cfg.appearance_fusion && (frameIndex + 1) % n->ctp->m_appearanceFusionRefreshRate == 0)
        for (int c = 0; c < cfg.n_cameras; c++)
            for (int o = 0; o < cfg.n_objects; o++) {
                n->select(c, o);
                n->obj[c][o]->LearnGlobalAppearanceModel(n->basePtr);
            }
    /* :602-608 -> Camera::StoreAllParticleFilterTrackerState (Camera.cpp:553-575) */
    for (int c = 0; c < cfg.n_cameras; c++)
        for (int o = 0; o < cfg.n_objects; o++) n->obj[c][o]->StoreParticleFilterTrackerState(frameIndex);
    mct_shim_rng_override = 0;
}
static ParticleFilterTracker* net_pft(Net* n, int c, int o) { return static_cast<ParticleFilterTracker*>(n->obj[c][o]->m_trackerPtr.get()); }
void ref_net_box(void* np, int c, int o, int frame, float* out4)
{
    ParticleFilterTracker* t = net_pft((Net*)np, c, o);
    for (int k = 0; k < 4; k++) out4[k] = t->m_states(frame, k);
}
void* ref_net_pf(void* np, int c, int o) { return net_pft((Net*)np, c, o)->m_particleFilterPtr.get(); }
void ref_net_state(void* np, int c, int o, float* out6) { const vectorf& s = net_pft((Net*)np, c, o)->GetCurrentTrackerState(); for (int i = 0; i < 6; i++) out6[i] = s[i]; }
int ref_net_selectors(void* np, int c, int o, int global, int* idx)
{
    Net* n = (Net*)np;
    Classifier::StrongClassifierBase* clf = global ? n->obj[c][o]->m_appearanceFuserPtr->m_strongClassifierBasePtr.get()
                                                   : net_pft(n, c, o)->m_strongClassifierBasePtr.get();
    for (size_t i = 0; i < clf->m_selectorList.size(); i++) idx[i] = clf->m_selectorList[i];
    return (int)clf->m_selectorList.size();
}
void ref_net_weak_state(void* np, int c, int o, int global, float* out8xF)
{
    Net* n = (Net*)np;
    Classifier::StrongClassifierBase* clf = global ? n->obj[c][o]->m_appearanceFuserPtr->m_strongClassifierBasePtr.get()
                                                   : net_pft(n, c, o)->m_strongClassifierBasePtr.get();
    int F = (int)clf->m_weakClassifierPtrList.size();
    for (int f = 0; f < F; f++) weak_state(clf->m_weakClassifierPtrList[f], out8xF, F, f);
}
int ref_net_kalman(void* np, int o, float* state4, float* P16)
{
    Net* n = (Net*)np;
    if (n->gf.empty()) return -1;
    CvKalman* k = n->gf[o]->m_pKalman;
    for (int i = 0; i < 4; i++) state4[i] = k->state_post->data.fl[i];
    for (int i = 0; i < 16; i++) P16[i] = k->error_cov_post->data.fl[i];
    return 0;
}
uint64_t ref_net_rng_state(void* np, int c, int o) { return *((Net*)np)->stream(c, o); }
int ref_net_last_train(void* np, int c, int o, int which, int* col, int* row, int cap)
{
    ParticleFilterTracker* t = net_pft((Net*)np, c, o);
    Classifier::SampleSet& s = which == 0 ? t->m_positiveSampleSet : t->m_negativeSampleSet;
    int m = (int)s.Size();
    for (int i = 0; i < m && i < cap; i++) { col[i] = s[i].m_col; row[i] = s[i].m_row; }
    return m;
}
}  // extern "C"
