/*
 * mct_oracle.h -- CPU ORACLE for the MultiCameraTracking appearance-scoring hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (multicameratracking_b200/) never links, imports or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's algorithm, function by function; every
 * function cites the reference file:line it follows (paths relative to /root/reference/src).
 *
 * PARITY PINNING STATUS (see DESIGN.md "Oracle"):
 *   - PINNED TO THE REFERENCE'S OWN CODE: oracle/_ref/libmct_ref.so is the unmodified reference (the .cpp files under /root/reference/src)
 *               compiled where it lies against header shims for IPP / OpenCV / Boost (oracle/build_ref.py, oracle/ref_shim/)
 *               and driven through oracle/ref_harness.cpp; tests/test_ref_pins_oracle.py runs it and this file on the same
 *               seeded inputs for every row of SURVEY 8a (bit-exact for integer / index / ordered-f32 work, 1e-5 elsewhere),
 *               the tracker loop and the multi-camera schedule included.
 *   - also pinned: MWC RNG stream, BGR->gray / HSV, cvRound, cvInvert / SVD / Kalman against cv2 4.13 (tests/golden/cv2_*.json),
 *               hand-computed integral / sumRect / resampler vectors (tests/golden/hand.json).
 *   - STATED ASSUMPTIONS THAT REMAIN (oracle/ref_shim/ipp.h): closed-source IPP arithmetic (ippiIntegral accumulation, ippiMean /
 *               StdDev summation) is the shim's restatement -- exact integer / f64 accumulation, rounded once -- and std::sort's
 *               order of exactly tied scores is unspecified by the language (the _ref build keeps ties in index order; the
 *               native-libstdc++ build is tested to differ in nothing else).
 */
#ifndef MCT_ORACLE_H
#define MCT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG: Public.h:156, Public.cpp:10-23 (OpenCV cvRNG multiply-with-carry) ---- */
typedef struct { uint64_t state; } orc_rng;
void     orc_rng_seed(orc_rng* r, int64_t seed);
uint32_t orc_rng_next(orc_rng* r);
int      orc_randint(orc_rng* r, int mn, int mx);
float    orc_randfloat(orc_rng* r);
int      orc_cvround(double v);

/* ---- integral image + rect sum: Matrix.cpp:33-67 ---- */
void  orc_integral(const uint8_t* img, int W, int H, int pitch, float* ii /* (H+1)*(W+1) dense */);
float orc_sum_rect(const float* ii, int W, int x, int y, int w, int h);

/* ---- frame pre-processing upstream of the path: Camera.cpp:449-494, Matrix.cpp:70-100 (IplImage BGR -> planes, channel 0 = R,
 * 1 = G, 2 = B), Matrix.cpp:518-533 (conv2BW = cvCvtColor BGR2GRAY), :535-580 (conv2HSV = cvCvtColor BGR2HSV; the reference writes
 * the result into the wrong matrix -- SURVEY 8f row 1 -- here the INTENDED conversion).  OpenCV 8-bit formulas: gray =
 * (R*4899 + G*9617 + B*1868 + 8192) >> 14; HSV = the fixed-point table algorithm (hsv_shift 12, H range 180).  Pinned against
 * cv2 4.13 by tests/golden/cv2_gray.json and cv2_hsv.json.  use_hsv: planes c0,c1,c2 = H,S,V instead of R,G,B. */
void orc_bgr_to_planes(const uint8_t* bgr, int W, int H, int pitch_bytes, int use_hsv,
                       uint8_t* gray, uint8_t* c0, uint8_t* c1, uint8_t* c2 /* dense [H][W] */);

/* ---- boxes = Sample (Sample.h:41-51), SoA ---- */
typedef struct {
    int n;
    const int* col; const int* row;        /* m_col, m_row (top-left) */
    const float* sx; const float* sy;      /* m_scaleX, m_scaleY */
} orc_boxes;

/* ---- Haar: HaarFeature.cpp:25-69 (Generate), :111-141 (Compute) ---- */
#define ORC_MAX_RECTS 6
typedef struct {
    int F;
    int*   nrect;    /* [F] */
    int*   rect;     /* [F][6][4] x,y,w,h in the w0 x h0 blob frame */
    float* weight;   /* [F][6] */
    int*   channel;  /* [F] (always 0) */
} orc_haar_table;
orc_haar_table* orc_haar_generate(orc_rng* r, int F, int w0, int h0);
void orc_haar_free(orc_haar_table* t);
/* out is feature-major [F][n] with leading dimension ld (SampleSet.h:41) */
void orc_haar_compute(const orc_haar_table* t, const float* ii, int W, const orc_boxes* b, float* out, int ld);

/* ---- colour histograms ---- */
/* MultiDimensionalColorHistogram.cpp:32-127; planes c0,c1,c2 = R,G,B (or H,S,V); out rows [0,nb^3) */
void orc_mdch_compute(const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, int pitch,
                      int nb, int w0, int h0, const orc_boxes* b, float* out, int ld);
/* CultureColorHistogram.cpp:25-111; mode 0 = reference_exact (all mass in bin 0), 1 = intended */
void orc_cch_compute(const uint8_t* hue, int pitch, int w0, int h0, const orc_boxes* b,
                     int mode, float* out, int ld);

/* ---- strong classifier (StrongClassifierBase + MILBoost / MILAnyBoost + weak clfs) ---- */
enum { ORC_FEAT_HAAR = 0, ORC_FEAT_CCH = 1, ORC_FEAT_MDCH = 2, ORC_FEAT_HAAR_MDCH = 3 };   /* FeatureParameters.h:12-16 */
enum { ORC_WEAK_STUMP = 0, ORC_WEAK_WSTUMP = 1, ORC_WEAK_PERCEPTRON = 2 };                 /* WeakClassifierBase.h:13-16 */
enum { ORC_STRONG_ADABOOST = 0, ORC_STRONG_MILBOOST = 2, ORC_STRONG_MILENSEMBLE = 3, ORC_STRONG_MILANYBOOST = 4 };                              /* ClassifierParameters.h:11-18 */

typedef struct {
    const uint8_t* gray; const uint8_t* c0; const uint8_t* c1; const uint8_t* c2;
    int W, H, pitch;
    const float* ii;    /* integral of gray, dense (H+1)*(W+1) */
} orc_frame;

typedef struct orc_clf orc_clf;
orc_clf* orc_clf_create(int feature_type, int n_haar, int nb, int w0, int h0, int weak_type,
                        int strong_type, int K, float lr, const orc_haar_table* haar /* copied; may be NULL */);
void orc_clf_free(orc_clf* c);
int  orc_clf_num_features(const orc_clf* c);
/* MILEnsemble only: Percentage_Of_Weak_Classifiers_Retained (MILEnsembleClassifierParameters, ClassifierParameters.h:105-117) */
void orc_clf_set_retained(orc_clf* c, float percentage);
/* FeatureVector::Compute: out [F][n], ld */
void orc_clf_features(const orc_clf* c, const orc_frame* f, const orc_boxes* b, float* out, int ld);
/* StrongClassifierBase::Update(pos,neg); wpos/wneg NULL => uniform 1.0 (SURVEY a12 decision) */
void orc_clf_update(orc_clf* c, const orc_frame* f, const orc_boxes* pos, const orc_boxes* neg,
                    const float* wpos, const float* wneg);
/* same, from precomputed feature matrices [F][P], [F][Nn] (used by the appearance fuser row) */
void orc_clf_update_from_features(orc_clf* c, const float* fpos, int P, const float* fneg, int Nn,
                                  const float* wpos, const float* wneg);
void orc_clf_classify(const orc_clf* c, const orc_frame* f, const orc_boxes* b, int log_ratio, float* out);
void orc_clf_classify_from_features(const orc_clf* c, const float* feat, int n, int ld, int log_ratio, float* out);
int  orc_clf_get_selectors(const orc_clf* c, int* idx);
/* weak state rows: mu0,mu1,sig0,sig1,n0,n1,e0,e1 (stumps) or w0,thr (perceptron, rows 0,1) */
void orc_clf_get_weak_state(const orc_clf* c, float* out8xF);
int  orc_clf_get_alphas(const orc_clf* c, float* alphas);   /* AdaBoost voting weights of the selected weak classifiers */
void orc_clf_set_threads(int n);   /* honours the reference's (compiled-out) omp pragmas when built with -fopenmp */

/* ---- likelihood map: ParticleFilterTracker.cpp:541-566 ---- */
void orc_likelihood_map(float* h, int n, int exponent);

/* ---- particle filter: ParticleFilter.{h,cpp} ---- */
enum { ORC_PF_UNINIT = 0, ORC_PF_INIT = 1, ORC_PF_PREDICTED = 2, ORC_PF_RESAMPLED = 3 };
typedef struct {
    int N;
    float img_w, img_h, max_scale_change, max_scale, min_scale;
    float* state;      /* [N][5] cx,cy,sx,sy,w  (m_particlesStateMatrix) */
    float* resampled;  /* [N][5]                (m_particlesResampledMatrix) */
    int*   resample_idx;
    int    filter_state, time_index;
    /* ordered-unique cache (FindOrderedUniqueParticles) */
    float* uniq; int n_uniq; int uniq_valid;
    int    last_resampled;   /* set by resample_if_required */
    float  last_neff_inv;
} orc_pf;
orc_pf* orc_pf_create(int N, float img_w, float img_h);
void orc_pf_free(orc_pf* p);
void orc_pf_initialize(orc_pf* p, float cx, float cy, float sx, float sy, float max_scale_change, float max_scale, float min_scale);
void orc_pf_predict(orc_pf* p, const float* noise4xN, float w0, float h0);
void orc_pf_update_particle_weight(orc_pf* p, int i, float prob, int force_reset);
/* returns 1 if a resample happened; *consumed_rng = number of RNG draws taken from r */
int  orc_pf_update_all_weights(orc_pf* p, const float* prob, int only_nonzero, int force_reset, int resample, orc_rng* r);
void orc_pf_normalize(orc_pf* p);
int  orc_pf_resample_if_required(orc_pf* p, orc_rng* r);
void orc_pf_resample(orc_pf* p, int force_state_change, orc_rng* r);
void orc_pf_resample_u01(orc_pf* p, int force_state_change, float u01);
void orc_pf_find_ordered_unique(orc_pf* p);
void orc_pf_get_average(const orc_pf* p, float* out4);
void orc_pf_get_highest_weight(const orc_pf* p, float* out5);
void orc_pf_get_highest_unique_close(orc_pf* p, const float* closest2, float* out5);
/* GenerateTestSampleSet (ParticleFilterTracker.cpp:1247-1305): fills boxes of the valid particles,
   zeroes weights of out-of-frame ones; returns count */
int  orc_pf_generate_test_set(orc_pf* p, float w0, float h0, int frame_w, int frame_h,
                              int* col, int* row, float* sx, float* sy);

/* ---- SampleSet samplers: SampleSet.cpp:82-361 ---- */
/* ring sampler; returns count written (<= cap) */
int orc_sample_ring(orc_rng* r, int img_w, int img_h, int x, int y, int w, int h,
                    float outer_r, float inner_r, int max_n, float sx, float sy, int* col, int* row, int cap);
int orc_sample_rect(orc_rng* r, int img_w, int img_h, int x, int y, int w, int h,
                    float maxdx, float maxdy, float mindx, float mindy, int max_n, float sx, float sy,
                    int* col, int* row, int cap);

/* ---- ParticleFilterTracker frame loop (InitializeTracker, TrackObjectAndSaveState) ---- */
typedef struct {
    int n_particles;
    float sd_x, sd_y, sd_sx, sd_sy;
    float pos_radius_train, init_pos_radius_train;
    int   n_neg_train, init_n_neg_train;
    int   max_pos_train;          /* m_maximumNumberOfPositiveTrainingSamples */
    int   search_win;
    int   trajectory_option;      /* 0 highest weight, 1 average */
    int   pos_strategy, neg_strategy;   /* only 0/0 restated (config.cfg defaults) */
    int   max_num_pos_examples;
} orc_tracker_params;
typedef struct orc_tracker orc_tracker;
orc_tracker* orc_tracker_create(const orc_tracker_params* tp, orc_clf* clf /* owned by caller */,
                                const orc_frame* f0, const float init_xywh[4], orc_rng* r /* shared stream */);
void orc_tracker_free(orc_tracker* t);
/* one frame: predict (noise given), score, weight update/resample, store state, update classifier.
   out_box = m_states row (x,y,w,h rounded).  phases: bit0 = score, bit1 = update classifier */
void orc_tracker_track(orc_tracker* t, const orc_frame* f, const float* noise4xN, int out_box[4], int phases);
int  orc_tracker_last_valid(const orc_tracker* t);
orc_pf* orc_tracker_pf(orc_tracker* t);
const float* orc_tracker_state(const orc_tracker* t);     /* m_currentStateList[6] */
/* last training sets (for parity tests of the update path) */
int  orc_tracker_last_train(const orc_tracker* t, int which /*0 pos,1 neg*/, int* col, int* row, float* sx, float* sy, int cap);

/* SimpleTracker (SimpleTracker.cpp), default strategies; shares orc_tracker_params (n_particles etc. unused) */
typedef struct orc_simple orc_simple;
orc_simple* orc_simple_create(const orc_tracker_params* tp, orc_clf* clf /* owned by caller */, const orc_frame* f0,
                              const float init_xywh[4], orc_rng* r /* shared stream */);
void orc_simple_free(orc_simple* t);
double orc_simple_track(orc_simple* t, const orc_frame* f, float out_state[4]);
int orc_simple_last_n(const orc_simple* t);
int orc_simple_last_best(const orc_simple* t);

#ifdef __cplusplus
}
#endif
#endif
