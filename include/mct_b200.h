/*
 * mct_b200.h -- C-ABI of the B200-native appearance-scoring path (libmct_b200.so).
 *
 * The reference (santhosh-kumar/MultiCameraTracking) has no plugin/FFI ABI: its seam is a set of
 * C++ virtual interfaces.  Each entry point below names the reference interface it
 * replaces (file:line under /root/reference/src).  The C++ host mirror in
 * multicameratracking_b200/host/ keeps the reference's class names and forwards to these calls.
 *
 * Conventions
 *   - every call returns 0 (MCT_OK) or a negative MCT_E_* code; it never exits or throws
 *     (the reference's abortError() calls exit(0), Public.h:288-308).  mct_last_error() gives text.
 *   - one mct_ctx per camera == per GPU, bound to one CUDA stream; calls on a ctx are stream-ordered
 *     and not thread-safe; distinct ctxs are independent.
 *   - plain pointers + sizes only.  "host" pointers are ordinary (ideally pinned) host memory;
 *     pointers documented "dev" are CUDA device pointers on the ctx's device.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns MCT_E_CUDA.
 *   - feature matrices are feature-major [F][ld] f32 exactly like SampleSet::m_featureMatrix
 *     (SampleSet.h:41-43, 92-93).
 */
#ifndef MCT_B200_H
#define MCT_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MCT_OK            0
#define MCT_E_ARG        -1   /* invalid argument */
#define MCT_E_CUDA       -2   /* CUDA runtime error / no device */
#define MCT_E_STATE      -3   /* call out of order (e.g. no frame set) */
#define MCT_E_EMPTY      -4   /* no valid test sample (reference: ASSERT abort, ParticleFilterTracker.cpp:513) */
#define MCT_E_NOMEM      -5
#define MCT_E_UNSUPPORTED -6

#define MCT_MAX_RECTS 6       /* DefaultParameters.h:14-15 */

/* Features::FeatureType (FeatureParameters.h:12-16) */
enum { MCT_FEAT_HAAR = 0, MCT_FEAT_CCH = 1, MCT_FEAT_MDCH = 2, MCT_FEAT_HAAR_MDCH = 3 };
/* Classifier::WeakClassifierType (WeakClassifierBase.h:13-16) */
enum { MCT_WEAK_STUMP = 0, MCT_WEAK_WSTUMP = 1, MCT_WEAK_PERCEPTRON = 2 };
/* Classifier::StrongClassifierType (ClassifierParameters.h:11-18) */
enum { MCT_STRONG_ADABOOST = 0, MCT_STRONG_MILBOOST = 2, MCT_STRONG_MILENSEMBLE = 3, MCT_STRONG_MILANYBOOST = 4 };   /* ClassifierParameters.h:11-18 */
/* ParticleFilter::State (ParticleFilter.h:33-37) */
enum { MCT_PF_UNINIT = 0, MCT_PF_INIT = 1, MCT_PF_PREDICTED = 2, MCT_PF_RESAMPLED = 3 };

typedef struct mct_ctx mct_ctx;
typedef struct mct_clf mct_clf;
typedef struct mct_pf  mct_pf;

/* ---------------------------------------------------------------- context ---------------- */
/* cuda_stream: a cudaStream_t to run on (e.g. torch's current stream) or NULL to create one. */
int  mct_ctx_create(int device, void* cuda_stream, mct_ctx** out);
int  mct_ctx_destroy(mct_ctx* ctx);
int  mct_sync(mct_ctx* ctx);
const char* mct_last_error(mct_ctx* ctx);           /* ctx may be NULL: last global error */
const char* mct_version(void);
/* number of kernels this library has launched on the ctx since creation (bench: gpu_launches) */
long long mct_launch_count(mct_ctx* ctx);
void* mct_ctx_stream(mct_ctx* ctx);
/* measurement hooks (bench.py): per-kernel CUDA-event timing on the ctx stream, and the number of test
 * samples scored since the last reset (sum of n_valid over mct_track_score calls, counted on the device).
 * names: [max_kernels][48] chars. */
int mct_profile_enable(mct_ctx* ctx, int on);
int mct_profile_read(mct_ctx* ctx, int max_kernels, char* names, float* total_ms, int* launches, int* n_out);
int mct_scored_particles(mct_ctx* ctx, unsigned long long* out, int reset);

/* ---------------------------------------------------------------- frame ------------------ */
/* Replaces Matrixu planes + Matrixu::initII()/FreeII() per frame (Matrix.cpp:33-47, Camera.cpp:491,389).
 * gray is required; c0,c1,c2 are the colour planes the MDCH/CCH features read (R,G,B in that order,
 * Matrix.cpp:92-94, or H,S,V) and may be NULL when only Haar features are used.
 * is_device = 0: host pointers, copied H2D asynchronously on the ctx stream;
 * is_device = 1: device pointers adopted without a copy (must stay valid until the next frame_set); the planes are
 *                taken to be produced by work already enqueued on the ctx stream (ordered after it);
 * is_device = 2: same, but the planes are already complete, so the frame preparation may start at once and overlap
 *                the previous frame's kernels (it runs on an internal stream into a second buffer set).
 * Computes the (H+1)x(W+1) f32 integral image of gray: exact u32 prefix sums rounded once to f32. */
int mct_frame_set(mct_ctx* ctx, const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2,
                  int W, int H, int pitch, int is_device);
/* The same from the PACKED BGR frame the reference's Camera reads (IplImage; Camera.cpp:449-494): the split into planar
 * R,G,B (Matrix.cpp:70-100), the gray conversion (conv2BW, Matrix.cpp:518-533: OpenCV's fixed-point BGR2GRAY) and, with
 * use_hsv, the 8-bit BGR2HSV conversion (conv2HSV, Matrix.cpp:535-580 -- the intended result; the reference stores it in
 * the wrong matrix) run on the device, so only 3 bytes per pixel cross PCIe.  pitch_bytes >= 3*W.  is_device as above. */
int mct_frame_set_bgr(mct_ctx* ctx, const uint8_t* bgr, int W, int H, int pitch_bytes, int use_hsv, int is_device);
int mct_frame_prefetch_bgr(mct_ctx* ctx, const uint8_t* bgr_host, int W, int H, int pitch_bytes, int use_hsv);
/* Issues the announced frame's upload and preparation now instead of at the end of the next mct_track_score* call: for a caller
 * that queues frame t + 1 (mct_frame_set + mct_track_score_async) while frame t is still being tracked. */
int mct_frame_prefetch_flush(mct_ctx* ctx);
/* Optional double buffering of the upload (Camera.cpp:405-495 reads frame t+1 from the video while frame t is
 * tracked): copies the NEXT frame's host planes (pinned memory for a truly asynchronous copy) into the ctx's back
 * buffers on a separate copy stream, overlapping the kernels of the current frame.  A following mct_frame_set with
 * the same pointers and geometry adopts the prefetched buffers instead of copying again.  The copy and the frame
 * preparation are issued lazily (behind the current frame's noise copy and beside its particle-filter update in
 * mct_track_score*, at the latest in the adopting mct_frame_set), so the host planes must stay valid and unchanged until
 * the first mct_sync / mct_track_collect after the adopting mct_frame_set. */
int mct_frame_prefetch(mct_ctx* ctx, const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2,
                       int W, int H, int pitch);
/* dense H*W host copies of the current frame's device planes gray, c0, c1, c2 (NULL pointers are skipped; tests) */
int mct_frame_get_planes(mct_ctx* ctx, uint8_t* gray, uint8_t* c0, uint8_t* c1, uint8_t* c2);
/* dense (H+1)*(W+1) host copy of the integral image (tests) */
int mct_frame_get_integral(mct_ctx* ctx, float* out_host);
/* Matrixu::sumRect (Matrix.cpp:49-67): n rects {x,y,w,h} -> ((br+tl)-tr)-bl in f32.  Host in/out. */
int mct_sum_rects(mct_ctx* ctx, const int* rects_xywh, int n, float* out_host);

/* ---------------------------------------------------------------- boxes == SampleSet ------ */
/* Sample.h:41-51 as SoA; all boxes refer to the ctx's current frame (the reference stores raw
 * Matrixu* in every Sample).  host arrays. */
typedef struct {
    int n;
    const int*   col;   /* m_col  (left) */
    const int*   row;   /* m_row  (top)  */
    const float* sx;    /* m_scaleX */
    const float* sy;    /* m_scaleY */
} mct_boxes;

/* ---------------------------------------------------------------- strong classifier ------- */
/* ClassifierParameters.h:38-65 + FeatureParameters.h */
typedef struct {
    int   feature_type;      /* MCT_FEAT_* */
    int   n_haar;            /* Tracker_Feature_Parameter (ignored for colour-only types) */
    int   n_bins;            /* Color_Number_Of_Bins (8) */
    int   w0, h0;            /* m_width, m_height of the object blob at scale 1 */
    int   weak_type;         /* MCT_WEAK_* */
    int   strong_type;       /* MCT_STRONG_* */
    int   n_selected;        /* K; clamped to F/2 when out of range (StrongClassifierBase.cpp:20-23) */
    float learning_rate;     /* 0.85 (DefaultParameters.h:18); 0 for the appearance fuser */
    int   cch_mode;          /* 0 = reference_exact, 1 = intended */
    int   max_train;         /* capacity for P + Nn training boxes (0 -> 4096) */
    int   max_test;          /* capacity for test boxes per classify call (0 -> 65536) */
    float pct_retained;      /* MILEnsemble only: Percentage_Of_Weak_Classifiers_Retained, in percent
                              * (MILEnsembleClassifierParameters, ClassifierParameters.h:105-117; stored / 100.0f, :52) */
} mct_clf_params;

/* Haar table generated on the host with the reference RNG order (HaarFeature.cpp:25-69). */
typedef struct {
    int n_features;
    const int*   nrect;      /* [F] 2..6 */
    const int*   rects;      /* [F][6][4] x,y,w,h in the w0 x h0 frame */
    const float* weights;    /* [F][6] */
} mct_haar_table;

/* StrongClassifierFactory::CreateAndInitializeClassifier (StrongClassifierFactory.cpp:14-43) +
 * StrongClassifierBase ctor (StrongClassifierBase.cpp:8-129). */
int mct_classifier_create(mct_ctx* ctx, const mct_clf_params* p, const mct_haar_table* haar, mct_clf** out);
int mct_classifier_destroy(mct_clf* clf);
/* StrongClassifierBase::GetNumberOfFeatures (StrongClassifierBase.h:47-51) */
int mct_classifier_num_features(mct_clf* clf);
/* FeatureVector::Compute (FeatureVector.h:40; HaarFeatureVector.cpp:48-79,
 * MultiDimensionalColorHistogramFeatureVector.cpp:52-90, CultureColorHistogramFeatureVector.cpp:48-81,
 * HaarAndColorHistogramFeatureVector.cpp:33-44): out_host is [F][boxes->n]. */
int mct_features_compute(mct_clf* clf, const mct_boxes* boxes, float* out_host);
/* StrongClassifierBase::Update(pos, neg) (MILBoostClassifier.cpp:15-157, MILAnyBoostClassifier.cpp:14-254).
 * wpos/wneg: WeightedStumps sample weights, NULL => uniform 1.0. */
int mct_classifier_update(mct_clf* clf, const mct_boxes* pos, const mct_boxes* neg,
                          const float* wpos, const float* wneg);
/* Same from already-computed feature rows [F][P], [F][Nn] (host) -- the appearance fuser's pooled
 * sample sets arrive as feature rows after the cross-camera allgather. */
/* device-resident forms for the appearance fuser (AppearanceBasedInformationFuser.cpp:44-90 with the pooled SampleSets replaced by
 * feature rows): the rows [F][n] are written to / read from device memory with the given leading dimensions, so that they
 * can go through an NCCL all-gather without touching the host.  Both calls have consumed / filled the caller's device buffers
 * when they return (the staging copies are waited for), so a stream-ordered allocator may recycle them at once. */
int mct_features_compute_dev(mct_clf* clf, const mct_boxes* boxes, float* dev_out, int ld_out);
int mct_classifier_update_from_features_dev(mct_clf* clf, const float* dev_fpos, int ld_pos, int P, const float* dev_fneg,
                                            int ld_neg, int Nn);
int mct_classifier_update_from_features(mct_clf* clf, const float* fpos, int P, const float* fneg, int Nn,
                                        const float* wpos, const float* wneg);
/* Learner side of AppearanceBasedInformationFuser::LearnGlobalAppearanceModel (AppearanceBasedInformationFuser.cpp:44-90):
 * the sample sets CameraNetwork::GenerateTrainingSampleSetsForAppearanceFusion (CameraNetwork.cpp:276-320) pools arrive as
 * feature rows in device exchange buffers.  Training column j (positives first) = the feature column starting at element
 * col_offset_host[j] of dev_rows, feature rows row_stride elements apart; the caller lists only the samples that survived
 * the randfloat() > 0.5 drop rule (:66-82), in pooled order. */
int mct_classifier_update_from_gathered(mct_clf* clf, const float* dev_rows, const int* col_offset_host, long long row_stride,
                                        int P, int Nn);
/* The two halves of one learner round of the appearance fuser for ALL objects of a camera (batched forms of
 * mct_features_compute_dev / mct_classifier_update_from_gathered):
 *  - every camera: feature rows of the freshly generated training sets of all its objects under the objects' global
 *    classifiers, written into the exchange buffer (object o at dev_rows + o * obj_stride floats, feature rows ld floats apart,
 *    positives from column 0, negatives from column ld_pos); stream-ordered, the box arrays may be reused at once.
 *    reuse_pos_rows (or NULL): the rows an earlier call wrote for IDENTICAL positive boxes (same layout) -- the learner rounds
 *    of one frame regenerate the same positives and only re-thin the negative ring, so the positive columns are copied from
 *    there and only the negative boxes are evaluated;
 *  - the learner: per object the kept columns gathered into the global classifier's training matrix and ONE batched Update.
 *    col_offsets_host: the objects' offset lists back to back (P[0] + Nn[0] entries, then P[1] + Nn[1], ...). */
int mct_train_features_many_dev(mct_ctx* ctx, int n_obj, mct_clf* const* clfs, const mct_boxes* pos, const mct_boxes* neg,
                                float* dev_rows, size_t obj_stride, int ld, int ld_pos, const float* reuse_pos_rows);
int mct_track_update_from_gathered(mct_ctx* ctx, int n_obj, mct_clf* const* clfs, const float* dev_rows, const int* col_offsets_host,
                                   long long row_stride, const int* P, const int* Nn);
/* device buffers on the ctx's device (exchange buffers) and a copy ordered on the ctx's stream (any direction, peers included) */
int mct_dev_alloc(mct_ctx* ctx, size_t bytes, void** out);
int mct_dev_free(mct_ctx* ctx, void* p);
int mct_dev_copy(mct_ctx* ctx, void* dst, const void* src, size_t bytes);
/* StrongClassifierBase::Classify(set, isLogRatioEnabled) (MILBoostClassifier.cpp:339-374). */
int mct_classifier_classify(mct_clf* clf, const mct_boxes* boxes, int log_ratio, float* out_host);
/* SimpleTracker (the reference's default local tracker, SimpleTracker.cpp:280-342): the search windows of all objects of a
 * camera are classified in one batched pass and only the first-maximum index and its response come back per object.
 * boxes: [n_obj] test sample sets (host); best_idx / best_resp: [n_obj] host. */
int mct_classify_argmax(mct_ctx* ctx, int n_obj, mct_clf* const* clfs, const mct_boxes* boxes, int log_ratio,
                        int* best_idx, float* best_resp);
int mct_classifier_classify_from_features(mct_clf* clf, const float* feat, int n, int log_ratio, float* out_host);
/* m_selectorList (StrongClassifierBase.h:63) */
int mct_classifier_get_selectors(mct_clf* clf, int* idx_host, int* n);
/* weak-classifier state, 8 rows x F: mu0,mu1,sig0,sig1,n0,n1,e0,e1 (OnlineStumpsWeakClassifier.h:33-40);
 * perceptron: rows 0,1 = m_weightList[0], [1] */
/* AdaBoost (AdaBoostClassifier.cpp): voting weights m_alphaList of the selected weak classifiers, in selector order */
int mct_classifier_get_alphas(mct_clf* clf, float* alphas_host, int* n);
int mct_classifier_get_weak_state(mct_clf* clf, float* out8xF_host);
/* last Update's weak predictions [F][P+Nn] (positives first) -- parity tests */
int mct_classifier_get_predictions(mct_clf* clf, float* out_host, int* P, int* Nn);
/* cluster width used by the MIL selection kernel (0 = auto) */
int mct_classifier_set_cluster(mct_clf* clf, int cluster_size);

/* ---------------------------------------------------------------- particle filter --------- */
/* ParticleFilter (ParticleFilter.h:28-156).  Device state is SoA cx,cy,sx,sy,w; host views are the
 * reference's row-major [N][5] (ParticleFilter.h:139). */
int mct_pf_create(mct_ctx* ctx, int n_particles, float image_w, float image_h, mct_pf** out);
int mct_pf_destroy(mct_pf* pf);
/* ParticleFilter::Initialize (ParticleFilter.cpp:15-75) */
int mct_pf_initialize(mct_pf* pf, float cx, float cy, float sx, float sy,
                      float max_scale_change, float max_scale, float min_scale);
/* ParticleFilter::PredictWithBrownianMotion (ParticleFilter.cpp:236-341).  noise4xN (host, or dev when
 * noise_is_device) holds the four cv::randn(1xN,0,sigma) rows (x, y, scaleX, scaleY): the host owns the
 * Gaussian stream.  w0/h0 feed CheckForParticleRefinement (:328). */
int mct_pf_predict(mct_pf* pf, const float* noise4xN, int noise_is_device, float w0, float h0);
/* ParticleFilter::UpdateAllParticlesWeight (ParticleFilter.cpp:680-717): prob is the compacted list
 * for the particles with non-zero weight.  u01 = the randfloat() the reference would draw inside
 * ResampleParticles (:936); *resampled tells the host whether that draw was consumed. */
int mct_pf_update_all_weights(mct_pf* pf, const float* prob_host, int n_prob, int only_nonzero,
                              int force_reset, int resample, float u01, int* resampled);
/* ParticleFilter::ResampleParticles(force) (ParticleFilter.cpp:920-978) */
int mct_pf_resample(mct_pf* pf, int force_state_change, float u01);
/* the same for all objects of a camera in one launch, stream-ordered (no synchronisation): Object.cpp:449 forces a
 * resampling of every object every frame when a fuser is on.  u01 [n_obj] host. */
int mct_pf_resample_many(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, int force_state_change, const float* u01);
/* m_particlesStateMatrix / m_particlesResampledMatrix / m_resampleIndexList */
int mct_pf_get_state(mct_pf* pf, float* out_Nx5_host);
int mct_pf_set_state(mct_pf* pf, const float* in_Nx5_host, int filter_state);
int mct_pf_get_resampled(mct_pf* pf, float* out_Nx5_host);
int mct_pf_get_resample_indices(mct_pf* pf, int* out_N_host);

/* Read-out used by StoreObjectState / GenerateTrainingSampleSet (ParticleFilterTracker.cpp:286-393,
 * 647-688), computed on the device so only this struct crosses PCIe per object per frame. */
typedef struct {
    int   n_valid;          /* test samples scored this frame */
    int   resampled;        /* 1 if ResampleIfRequired fired (the u01 draw was consumed) */
    int   filter_state;     /* MCT_PF_* after the call */
    int   n_unique;         /* rows of m_particlesOrderedUniqueMatrix */
    float neff_inv;         /* sum w^2 after normalisation (ParticleFilter.cpp:860-875) */
    float max_response;     /* max H over the test set (return value of TrackObjectOnTheGivenFrame) */
    float average[4];       /* GetAverageofAllParticles (ParticleFilter.cpp:601-633) */
    float ordered_first[5]; /* GetOrderedUniqueParticles(0) (ParticleFilter.cpp:507-522, 728-815) */
    float highest_close[5]; /* GetHighestOrderedUniqueParticleCloseToTheGivenState (:557-593) */
} mct_pf_summary;
int mct_pf_get_summary(mct_pf* pf, mct_pf_summary* out_host);
/* the read-outs of several objects of one camera in one launch and one synchronisation
 * (Camera::StoreAllParticleFilterTrackerState, Camera.cpp:556-574) */
int mct_pf_get_summary_many(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, mct_pf_summary* out_host /* [n_obj] */);

/* ---------------------------------------------------------------- fused per-frame path ---- */
/* One call = TrackObjectOnTheGivenFrame (ParticleFilterTracker.cpp:468-639) for n_obj objects of this
 * camera in a handful of batched launches: predict -> GenerateTestSampleSet (:1247-1305) -> Classify
 * (selected features only; identical result to the reference, which computes all F and reads K) ->
 * likelihood map ((H-min)/range)^exponent (:541-566) -> UpdateAllParticlesWeight -> Normalize ->
 * ResampleIfRequired -> read-out.  noise: [n_obj][4][N] (host or dev); u01: [n_obj] host.  noise == NULL selects the
 * re-score variant (ParticleFilterTracker.cpp:1334-1482): the particles are scored where they are, without prediction.
 * summaries: [n_obj] host, valid after the call returns (the call waits for the frame's read-outs).
 * u01 == NULL: the draw of ResampleParticles comes from every object's device-resident stream (mct_pf_rng_set). */
typedef struct {
    float w0, h0;           /* m_currentStateList[2], [3] */
    int   exponent;         /* 20 (:556); 1 for the re-score variant (:1407); 0 = the appearance fuser's map
                             * sigmoid(H) instead of the min-max map (AppearanceBasedInformationFuser.cpp:98-121) */
    int   force_reset;      /* shouldForceReset */
    int   resample;         /* shouldResample */
} mct_track_params;
int mct_track_score(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, mct_clf* const* clfs,
                    const mct_track_params* params /* [n_obj] */,
                    const float* noise, int noise_is_device, const float* u01,
                    mct_pf_summary* summaries);
/* Optional: announce the host noise block ([n_obj][4][N], n_floats in total) of the NEXT mct_track_score* call.  It is
 * uploaded on the internal copy stream while the current frame is being tracked; the next mct_track_score* called with
 * the same host pointer then finds it on the device instead of copying it in front of its first kernel.  The host block
 * must stay valid and unchanged until that call has been collected. */
int mct_noise_prefetch(mct_ctx* ctx, const float* noise_host, size_t n_floats);
/* asynchronous variant: no stream sync, summaries land in the ctx's pinned buffer; collect with mct_track_collect, which waits for
 * the read-outs of the OLDEST call not collected yet (an event behind that call's last kernel; work queued later -- the frame's
 * UpdateClassifier, the next frame -- keeps running).  Two calls may be in flight: the read-outs alternate between two slots;
 * a third uncollected call retires the oldest.  With no call pending, mct_track_collect synchronises the stream and returns the
 * latest read-outs again.  An empty training set met by the previous frame's mct_track_update_default is reported here. */
int mct_track_score_async(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, mct_clf* const* clfs,
                          const mct_track_params* params, const float* noise, int noise_is_device,
                          const float* u01);
int mct_track_collect(mct_ctx* ctx, int n_obj, mct_pf_summary* summaries);
/* UpdateClassifier (ParticleFilterTracker.cpp:1013-1032) for n_obj objects, batched, asynchronous. */
int mct_track_update(mct_ctx* ctx, int n_obj, mct_clf* const* clfs,
                     const mct_boxes* pos /* [n_obj] */, const mct_boxes* neg /* [n_obj] */);
/* per-object H (log-ratio response) of the last mct_track_score, dense over particles (tests) */
int mct_pf_get_last_response(mct_pf* pf, float* out_N_host, int* valid_N_host);

/* ---------------------------------------------------------------- cross-camera exchange --- */
/* Packs the per-object payload the fusers exchange: foot point of the highest-weight
 * particle (cx, cy + sy*h0/2) (ParticleFilterTracker.cpp:1223-1238) into dev_out[n_obj][2] (dev),
 * ready for an NCCL allgather on the same stream. */
/* ---- geometric-fuser feedback into one camera's particles (ParticleFilterTracker.cpp:1040-1125) ----
 * mct_pf_ground_pdf: EstimateGroundPoint(false) (:1134-1160) + TransformWithHomography (GeometryBasedInformationFuser.cpp:
 * 172-212) + MultiVariateNormalPdf (:111-163, exponent without the 1/2 as in the reference) for every particle, their total,
 * and GetAverageofAllParticles (ParticleFilter.cpp:601-633) for the 20-pixel rearrangement test (:1090-1105) the caller
 * makes.  H9: image -> ground homography (row-major f64); mean2 / cov4: the Kalman posterior (f32).  weights_host: optional
 * [N] read-back of the per-particle pdf (all 1 when the covariance is unusable, :1063-1077). */
typedef struct {
    float total_weight;
    float average[4];
    int   usable_covariance;
} mct_ground_pdf_result;
int mct_pf_ground_pdf(mct_pf* pf, const double* H9, const float* mean2, const float* cov4, float h0,
                      mct_ground_pdf_result* out, float* weights_host);
/* The same for all objects of a camera in one launch and one read-back (FeedbackInformationFromGeometricFusion's loop,
 * CameraNetwork.cpp:105-121): H9 is the camera's homography, mean2 [n_obj][2], cov4 [n_obj][4], h0 [n_obj], out [n_obj]. */
int mct_fuser_ground_pdf(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const double* H9, const float* mean2,
                         const float* cov4, const float* h0, mct_ground_pdf_result* out);
/* RearrangeParticlesBasedOnGroundLocation (ParticleFilter.cpp:120-179, scale-changing form) followed by
 * CheckForParticleRefinement (:173).  noise2xN: host [2][N], the (xRand, yRand) = randgaus(0, 10) pairs in particle order
 * (drawn by the caller from the reference's MWC stream, Public.cpp:37-49).  The re-scoring that follows in the reference
 * (UpdateParticleWeightsForRearrangedParticles, ParticleFilterTracker.cpp:1334-1482) is mct_track_score with noise == NULL
 * (no prediction), exponent 1, force_reset 1, resample 1. */
int mct_pf_rearrange_ground(mct_pf* pf, float mean_foot_x, float mean_foot_y, const float* noise2xN, float w0, float h0);
/* the same for several objects of a camera in one pass: mean_foot_xy [n_obj][2]; noise host, object o's [2][N] block behind
 * the blocks of the objects in front of it; w0 / h0 [n_obj] */
int mct_pf_rearrange_ground_many(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const float* mean_foot_xy, const float* noise,
                                 const float* w0, const float* h0);

/* The same exchange WITHOUT a collective library: every camera's receive buffer is exported with CUDA IPC and the kernel that
 * computes the foot points stores them straight into all cameras' buffers over NVLink, then a bounded wait gathers
 * everybody's frame (replaces CameraNetwork.cpp:182-216 + the all-gather).  Set-up: mct_p2p_create on every rank, exchange the
 * 64-byte handles by any means (e.g. torch.distributed.all_gather_object), mct_p2p_open with all of them, barrier.
 * mct_fuser_exchange_foot_points: host_out [world][n_obj][2]; every camera calls it once per frame; timeout_ms <= 0 -> 2000. */
int mct_p2p_create(mct_ctx* ctx, int world, int rank, unsigned char* ipc_handle_out64);
int mct_p2p_open(mct_ctx* ctx, const unsigned char* all_handles);
int mct_p2p_destroy(mct_ctx* ctx);
int mct_fuser_exchange_foot_points(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const float* h0, float* host_out, int timeout_ms);

int mct_fuser_pack_foot_points(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const float* h0, float* dev_out);
/* the same payload read back to the host, host_out [n_obj][2] (what CameraNetwork::FuseGeometrically, CameraNetwork.cpp:182-216,
 * collects from every camera) */
int mct_fuser_foot_points(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const float* h0, float* host_out);

/* ---------------------------------------------------------------- training samples from the particles -------
 * ParticleFilterTracker::GeneratePositiveTrainingSampleSet, the strategies that read the particle filter
 * (ParticleFilterTracker.cpp:743-892; PFTracker_Positive_Sample_Strategy, TrackerParameters.h:42-51), for all objects
 * of a camera in one device pass over the device-resident particles:
 *   1 SAMPLE_POS_GREEDY           the first min(#ordered unique, max_pos) ordered unique particles with non-zero weight
 *                                 whose box lies inside the frame (FindOrderedUniqueParticles, ParticleFilter.cpp:728-815);
 *   2 SAMPLE_POS_SEMI_GREEDY      the same, additionally closer than 8 px to the current state's centre;
 *   3 SAMPLE_POS_PARTICLE_RANDOM  every in-frame resampled particle kept iff randfloat() < max_pos / N: draw k of the
 *                                 tracker's multiply-with-carry stream belongs to the k-th in-frame particle, reproduced
 *                                 in parallel by jumping the generator ahead (state_n = A^n state_0 mod (A 2^32 - 1)).
 * The read-out also carries what SAMPLE_NEG_PARTICLE_RANDOM (:939-995) needs from the particles: the largest
 * |x - centre|, |y - centre| over the ordered unique particles.  Boxes come back as [n_obj][cap] host arrays. */
typedef struct {
    int   strategy;                  /* 1, 2 or 3 (0 = only the negative-strategy read-out) */
    int   max_pos;                   /* m_maxNumPositiveExamples */
    float cur[6];                    /* m_currentStateList: leftX, topY, width, height, scaleX, scaleY */
    int   frame_w, frame_h;
    unsigned long long rng_state;    /* the tracker's cvRNG state (strategy 3) */
} mct_train_gen;
typedef struct {
    int   n_pos;                     /* boxes written for this object (<= cap) */
    int   n_unique;                  /* GetNumberOfOrderedUniqueParticles */
    float max_dx, max_dy;            /* over the ordered unique particles */
    unsigned long long rng_state;    /* state after the draws of strategy 3 (unchanged otherwise) */
    int   n_draws, pad;
} mct_train_gen_out;
int mct_pf_training_boxes(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const mct_train_gen* gen, int cap,
                          int* col, int* row, float* sx, float* sy, mct_train_gen_out* out);

/* SampleSet::SampleImage, ring form (SampleSet.cpp:82-168) and rectangle form (:180-262), followed by
 * SelectSamplesUniformlyFromLargerSet (:331-361), for a batch of regions in one device pass: candidates are enumerated in the
 * reference's row-major order, candidate k owns draw k of the tracker's RNG stream (jump-ahead, see mct_pf_training_boxes) and
 * survives iff !(randfloat() > (max_n + 1) / n_candidates); the survivors keep their order and are cut at max_n.  No draws
 * happen when the probability is >= 1.  col / row: [n_regions][cap] host arrays. */
typedef struct {
    int   kind;                      /* 0 ring: p = {outerCircleRadius, innerCircleRadius}; 1 rectangle: p = {maxDX, maxDY, minDX, minDY} */
    int   x, y, width, height;       /* SampleImage arguments */
    float p[4];
    int   max_n;                     /* maximumNumberOfSamples */
    float scale_x, scale_y;
    int   frame_w, frame_h;          /* image cols / rows */
    unsigned long long rng_state;
} mct_sample_region;
typedef struct {
    int n;                           /* samples kept (<= cap) */
    int n_candidates;                /* size of the larger set */
    int n_draws, pad;                /* randfloat() calls consumed */
    unsigned long long rng_state;    /* state after them */
} mct_sample_region_out;
int mct_sample_regions(mct_ctx* ctx, int n_regions, const mct_sample_region* regions, int cap, int* col, int* row,
                       mct_sample_region_out* out);

/* ---------------------------------------------------------------- the whole frame on the device ---- */
/* The tracker's multiply-with-carry stream (Public.cpp:15-23, `rng_state`) as DEVICE-resident state of the particle filter:
 * after mct_pf_rng_set, mct_track_score* called with u01 == NULL takes the randfloat() of ResampleParticles
 * (ParticleFilter.cpp:940) from that stream -- consumed iff the filter resamples, as in the reference -- and
 * mct_track_update_default draws the thinning decisions of SelectSamplesUniformlyFromLargerSet from it.  mct_pf_rng_get
 * synchronises the stream and returns the current state (the host mirror adopts it when host-side code needs the stream). */
int mct_pf_rng_set(mct_pf* pf, unsigned long long state);
int mct_pf_rng_get(mct_pf* pf, unsigned long long* state);

/* GenerateTrainingSampleSet with the default strategies (SAMPLE_POS_SIMPLETRACKER / SAMPLE_NEG_SIMPLETRACKER,
 * ParticleFilterTracker.cpp:647-741, 913-937) + UpdateClassifier (:1013-1032) for n_obj objects, queued behind the frame's
 * mct_track_score* with NO host round trip: one kernel reads the particle read-out where k_pf_update left it, derives
 * m_currentStateList (:655-672), enumerates the positive disc and the negative ring (SampleSet.cpp:82-168), thins them with
 * the device-resident stream (:331-361; candidate k takes draw k by jump-ahead), writes the boxes, the set sizes and the plan
 * of the colour-histogram kernels straight into the device-side descriptors; the update kernels are launched from upper
 * bounds.  Same boxes, same order, same draws as the host enumeration (mct_track_update).  Asynchronous. */
typedef struct {
    float w0, h0;                 /* m_currentStateList[2], [3] */
    int   trajectory_option;      /* m_outputTrajectoryOption: 0 = highest-weight particle, 1 = average of all particles */
    float pos_radius;             /* m_posRadiusTrain */
    int   max_pos;                /* m_maximumNumberOfPositiveTrainingSamples */
    float neg_outer, neg_inner;   /* 1.5 * m_searchWindSize, m_posRadiusTrain + 15 (:926-930) */
    int   max_neg;                /* m_numberOfNegativeTrainingSamples */
    float scale_hint_x, scale_hint_y;   /* last known state scale; sizes the launches only: a wrong hint costs time, never results */
} mct_train_default;
typedef struct {
    int P, Nn;                    /* sizes of the positive / negative set */
    int status;                   /* MCT_OK; MCT_E_EMPTY: no particle in the frame or an empty set -- the classifier was left untouched */
    int mdch_fast;                /* the colour rows took the pixel-list kernels (diagnostic) */
    unsigned long long rng_state; /* the stream after this frame's draws */
    float cur[6];                 /* m_currentStateList after GenerateTrainingSampleSet */
    int n_draws, pad;             /* randfloat() calls of the two thinning passes */
} mct_train_default_out;
int mct_track_update_default(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, mct_clf* const* clfs, const mct_train_default* gen /* [n_obj] */);
/* synchronises the stream; out [n_obj] (or NULL: only the status is checked); returns the first object's error, if any */
int mct_track_update_default_collect(mct_ctx* ctx, int n_obj, mct_train_default_out* out);
/* the boxes generated for object `obj` by the last mct_track_update_default (positives first; tests / diagnostics; synchronises) */
int mct_track_update_default_boxes(mct_ctx* ctx, int obj, int cap, int* col, int* row, float* sx, float* sy, int* P, int* Nn);

/* sizeof of the public structs in declaration order (mct_boxes, mct_clf_params, mct_haar_table, mct_pf_summary, mct_track_params,
 * mct_ground_pdf_result, mct_train_gen, mct_train_gen_out, mct_sample_region, mct_sample_region_out, mct_train_default,
 * mct_train_default_out): a binding in another language checks its own layouts against it.  Returns the number of entries. */
int mct_abi_struct_sizes(int* out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* MCT_B200_H */
