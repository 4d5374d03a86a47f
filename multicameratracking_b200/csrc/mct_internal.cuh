// mct_internal.cuh -- shared host structs, launch helpers and device arithmetic helpers.
// Compiled for sm_100a only, with -fmad=false so every f32 result follows the reference's
// operation order (no FMA contraction); see DESIGN.md "Arithmetic contract".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/mct_b200.h"

#define MCT_II_LEAD 3          // integral rows start 3 floats into a 16B-aligned row so that column 1 is 16B aligned
#define MCT_FIX_SHIFT_MAX 20   // MDCH fixed-point histogram scale 2^20
#define MCT_MDCH_SEG_PIX 4096  // pixels per fixed-point segment (u32 cannot overflow: 4096 * 2^20 = 2^32 excluded by weight <= 1 and shift rule)
#define MCT_MDCH_TRAIN_PIX 8191 // largest box of the pixel-list training kernels: one fixed-point segment at 2^clz(pixels) >= 2^19 (8191 * 2^19 < 2^32)

// ------------------------------------------------------------------ per-object device status
struct PfDev {
    int   filter_state;
    int   n_valid;
    int   resampled;
    int   n_unique;
    float neff_inv;
    float max_response;
    int   time_index;
    int   refine;        // CheckForParticleRefinement hand-over between the predict kernel and k_pf_update (k_pf.cu)
};
// Device-resident tracker state of one object (mct_pf::d_trk): what GenerateTrainingSampleSet reads from the particle filter
// (ParticleFilterTracker.cpp:655-672), kept next to the mapped host read-out so that the training-sample kernel of the same
// frame needs no host round trip, and the tracker's multiply-with-carry stream when it lives on the device (mct_pf_rng_set).
struct PfTrk {
    unsigned long long rng;      // Public::rng_state of this tracker (Public.cpp:15-23)
    float avg[4];                // GetAverageofAllParticles
    float first[4];              // GetOrderedUniqueParticles(0)
    int   n_valid, pad;
};

// One selected weak classifier, ready for Classify: Haar rect table (or the featN row of a colour feature) + the
// 8-float weak state.  The table of a classifier is rebuilt on the device after every Update (k_build_colormap)
// so that the per-frame kernels fetch it with one bulk copy instead of chasing selector -> feature -> state.
struct SelEntry {
    int4  r[MCT_MAX_RECTS];
    float w[MCT_MAX_RECTS];
    float st[8];
    int   nrect;      // >0: Haar feature; 0: colour feature, `row` indexes featN
    int   row;
};

// ------------------------------------------------------------------ host-side objects
struct TrainDesc;
struct FrameSet {              // one of the two sets of per-frame device buffers
    uint8_t* planes[4]; size_t plane_cap;
    float* ii_base; size_t ii_cap;
    uint16_t* binimg; size_t binimg_cap;
};
struct PrepTarget {            // what one frame-prep launch reads and writes
    const uint8_t *gray, *c0, *c1, *c2; int W, H, pitch;
    float* ii_base; int ii_pitch; uint16_t* binimg;
};
struct GroundOut;
struct GroundArgs;
// peer-memory exchange of the geometric fuser's payload (k_pf.cu: k_foot_points_push / _wait)
#define MCT_P2P_MAXW 16
struct P2PBuf {
    unsigned flags[2][MCT_P2P_MAXW];                       // [parity][camera] sequence number of the last frame published
    float data[2][MCT_P2P_MAXW][64][2];                    // [parity][camera][object] foot point
};
struct P2PPeers { P2PBuf* p[MCT_P2P_MAXW]; };
struct mct_ctx {
    int device;
    int sms;                               // multiprocessors of this ctx's device
    cudaStream_t stream;
    bool own_stream;
    int W, H, pitch;                       // current frame geometry (u8 planes)
    const uint8_t* gray; const uint8_t* c0; const uint8_t* c1; const uint8_t* c2;   // device planes in use
    uint8_t* own_planes[4]; size_t own_plane_cap;
    // Double buffering of the frame: the planes / integral image / bin image of frame t+1 are produced on
    // prep_stream into the BACK set while the compute stream still works on frame t (mct_api.cu, "frame").
    FrameSet back;
    cudaStream_t prep_stream; cudaEvent_t ev_ready, ev_mark[2], ev_gather; int mark_idx, has_gather;
    unsigned prep_epoch;
    const uint8_t* pf_src[4]; int pf_W, pf_H, pf_pitch, pf_pending, pf_with_bins, pf_prepped, pf_copied, pf_kind, pf_hsv;
    uint8_t* d_bgr_stage; size_t bgr_stage_cap;         // packed BGR upload staging (consumed by k_bgr_unpack on the prep stream)
    float* ii_base; int ii_pitch; size_t ii_cap;       // integral image storage; element (y,x) at ii_base[MCT_II_LEAD + y*ii_pitch + x]
    uint32_t* bandsum; size_t bandsum_cap;
    unsigned* prep_flags;                              // [256] band-ready flags of k_frame_prep (value = frame id)
    uint16_t* binimg; size_t binimg_cap; int binimg_nb; long long binimg_frame;   // MDCH joint-bin image of the colour planes
    long long frame_id;
    void* d_desc; void* h_desc; size_t desc_cap;       // batched object descriptors (device + pinned host)
    void* desc_last_h; const void* desc_last_d; int desc_last_n;   // last published descriptor set (skip identical uploads)
    mct_pf_summary* h_summ; mct_pf_summary* h_summ_dev; int summ_cap;   // pinned, device-mapped read-out buffer (kernels write it directly)
    float* d_noise_stage; size_t noise_stage_cap;
    float* d_noise_pf[2]; size_t noise_pf_cap[2]; int noise_pf_cur;         // prefetched noise blocks (mct_noise_prefetch)
    const float* npf_ready_src; size_t npf_ready_n; int npf_ready;          // block already on the device, in d_noise_pf[noise_pf_cur]
    const float* npf_req_src; size_t npf_req_n; int npf_req;               // block announced for the next call
    cudaEvent_t ev_noise;
    TrainDesc* h_tdesc; TrainDesc* d_tdesc; cudaEvent_t ev_train;   // batched UpdateClassifier descriptors (pinned + device)
    cudaStream_t train_stream; cudaEvent_t ev_tfork, ev_tjoin;      // the training Haar rows run beside the MDCH rows
    int* h_train_stage; int* d_train_stage; size_t train_stage_cap;   // packed training boxes of all objects (4 x 4-byte arrays)      // staging for host-provided prediction noise (one H2D per call)
    float* d_tmp; size_t tmp_cap;                      // scratch (sum_rects etc.)
    float* h_pin; size_t pin_cap;                      // pinned staging for small host arrays
    long long launches;
    unsigned long long* d_scored;                      // device counter: test samples scored (sum of n_valid)
    void* prof;                                        // per-kernel event timing state (NULL = off)
    char err[512];
    void* d_argmax; int argmax_cap;   // read-out of k_response_argmax [n_obj] {idx, val}
    uint32_t* d_mlist; size_t mlist_cap; uint32_t* d_macc; size_t macc_cap; uint32_t* d_mtot;   // k_mdch_train_* work buffers
    P2PPeers p2p; int p2p_world, p2p_rank; unsigned p2p_seq; unsigned* d_p2p_ticket; float* h_p2p_out; float* h_p2p_out_dev; int* h_p2p_status; int* h_p2p_status_dev;
    // whole frame on the device (mct_track_update_default): per-object generator arguments (uploaded when they change), the mapped
    // read-out of k_train_gen_default, the descriptor set last uploaded for it, the event behind the frame's k_pf_update
    struct TrainGenDev* d_gen; void* gen_shadow; int gen_shadow_n;
    mct_train_default_out* h_gen_out; mct_train_default_out* h_gen_out_dev;
    void* tdesc_shadow; int tdesc_shadow_n; int tdesc_gen_valid;
    mct_clf* gen_clfs[64]; int gen_n; int gen_pending;
    long long gen_seq, score_gen_seq, gen_checked_seq;   // read-outs alternate between two buffers (gen_seq & 1); see mct_track_collect
    int weak_split_done;         // launch_train_features queued the weak-classifier update itself (one part per feature stream)
    cudaEvent_t ev_score[2]; int has_ev_score;      // behind the k_pf_update of the score call that wrote read-out slot 0 / 1
    long long score_issued, score_collected;        // fused score calls queued / collected: at most two in flight (slot = count & 1)
    long long score_gen_seq_slot[2];
    GroundOut* d_ground_out;     // read-out of k_ground_pdf [DESC_MAX_OBJ]
    GroundArgs* d_ground_args;   // its per-object arguments [DESC_MAX_OBJ]
};

struct mct_clf {
    mct_ctx* ctx;
    mct_clf_params p;
    int F, K, n_haar, n_color;
    // Haar table
    int4* d_rects; float* d_wts; int* d_nrect;
    int haar_mxw, haar_mxh;                             // max over the table of rect.x + rect.w / rect.y + rect.h (blob frame)
    // weak state [8][F], trained flags
    float* d_weak; int* d_trained;
    // selection
    int* d_sel; int* d_nsel;
    SelEntry* d_seltab; int* d_selhdr;                  // [K] prepared selected-classifier table; hdr = {nsel, ext_w, ext_h, 0}
    int cluster;                                       // requested cluster width (0 auto)
    // training buffers
    int max_train, ldT;
    int* d_tcol; int* d_trow; float* d_tsx; float* d_tsy;   // P then Nn boxes
    float* d_tw;                                        // sample weights (P then Nn)
    float* d_featT;                                     // [F][ldT]
    float* d_pred;                                      // [F][ldT]
    int lastP, lastNn;
    // test buffers (grow-only)
    int test_cap, ldN;
    int* d_col; int* d_row; float* d_sx; float* d_sy;
    float* d_featN; size_t featN_cap;                   // [rows][ldN]
    float* d_H;
    // selected colour rows map for classify (built on the device from the selector list after every Update)
    int* d_outbins; int* d_colorrow;                    // [n_color]: bins to write / colour feature -> row in featN
    uint16_t* d_slotmap;                                // [n_color]: joint bin -> private-histogram slot (0 = not selected)
    int* d_nout;                                        // device: number of selected colour bins
    float* d_rowst;                                     // [n_color][8]: weak state of the selector that owns colour row r (fused responses)
    int* d_tsort; size_t tsort_cap;                     // tiled Haar matrix: tile starts + box order (k_tile_sort)
    int* d_big; int big_cap;                            // per test box: 1 = left to the generic (large / odd box) kernel
    int* d_nbig;                                        // device counter of such boxes
    float* d_ada_cnt; float* d_alpha;                   // AdaBoost state ([4][K][F], initial value 1; [K])
    int* d_keep; float* d_ensH;                         // MILEnsemble: [F] retained flags, [ldT] hypothesis left by the retain step
};

// geometric-fuser feedback: per-call arguments and read-out of k_ground_pdf
struct GroundArgs {
    double H[9];         // image -> ground homography
    double a;            // 1 / pow(sqrt(2 pi), 2)
    float mean[2];       // Kalman posterior mean
    float ci[4];         // inverse covariance (cvInvert, f32)
    float b;             // 1 / sqrt(det)
    float h0;
    int usable;          // det >= 0 && cov[0] > 0 (else every weight is 1)
};
struct GroundOut { float total_weight; float average[4]; };

struct mct_pf {
    mct_ctx* ctx;
    int N, ld;
    float img_w, img_h, msc, maxs, mins;
    float* d_state;      // 5*ld : cx,cy,sx,sy,w
    float* d_res;        // 5*ld
    int*   d_residx;
    int*   d_col; int* d_row; int* d_valid;
    float* d_H;
    float* d_noise;      // 4*ld
    float* d_prob;       // ld
    float* d_scratch;    // 2*ld (cdf / weights when N too large for smem)
    PfDev* d_st;
    PfTrk* d_trk;
    mct_pf_summary* d_summ;
    int initialized;
};

// ------------------------------------------------------------------ batched object descriptor
struct ObjDesc {
    // particle filter
    float* cx; float* cy; float* sx; float* sy; float* w;
    float* rcx; float* rcy; float* rsx; float* rsy; float* rw;
    int* residx; int* col; int* row; int* valid;
    float* H; long long noise_off; float* scratch;      // noise_off: float offset of this object's 4 x N noise block in StepArgs::noise_base
    PfDev* st; mct_pf_summary* summ; unsigned long long* scored;
    PfTrk* trk;
    int N, ld;
    float img_w, img_h, msc, maxs, mins;
    float w0, h0;
    int exponent, force_reset, resample, pad0;
    // classifier
    const int4* rects; const float* wts; const int* nrect;
    const float* weak; const int* sel; const int* nsel;
    const SelEntry* seltab; const int* selhdr;
    const float* alpha;                         // AdaBoost voting weights per selector (NULL for the MIL classifiers)
    float* featN; int ldN; int n_color_rows;
    const int* colorrow; const int* outbins;
    const uint16_t* slotmap; const int* nout; int* big; int* nbig;
    int F, n_haar, nb, weak_type, feature_type, cch_mode, K, cw0, ch0;
    int color_resp;      // 1: the per-frame colour kernel leaves ClassifyF(x) of the owning selector in featN instead of the bin value x
    const float* rowst;  // [n_color_rows][8] weak state per colour row (valid when color_resp)
};

// ------------------------------------------------------------------ batched training descriptor (UpdateClassifier)
// One entry per object of a camera: all objects' classifiers are updated by the same handful of launches
// (blockIdx.y / .z = object) instead of five launches + eight copies per object.
// One group of training boxes of one object (the positives, or the negatives): same box size, corners inside one bounding
// region.  Filled on the host (mct_track_update); n == 0 or ok == 0 sends the object to the generic kernel.
#define MTR_EXT_WORDS 12288    // largest extended weight table of k_mdch_train_main (words of shared memory)
struct MdchGroup {
    int first, n;              // box index range [first, first + n) in the object's training set
    int rows, cols;            // common box size in pixels
    int x0, y0, RW, RH;        // bounding region of the boxes (image coordinates of its corner, extent)
    int EW, EH;                // EW != 0: the walk uses the extended weight table E[(dy + RH - rows) * (2 RW - cols) + dx + RW - cols], zero outside the box
    int list_off, npx;         // this group's slice of the sorted pixel list
    int nchunk, chunk_px;      // list chunks (one CTA each)
    int shift, pad;
};
struct TrainDesc {
    const int* col; const int* row; const float* sx; const float* sy; const float* tw;   // P positives then Nn negatives
    int P, Nn, ldT, pad0;
    float* featT; float* pred;
    const int4* rects; const float* wts; const int* nrect;
    float* weak; int* trained; int* sel; int* nsel;
    uint16_t* slotmap; int* colorrow; int* outbins; int* nout; SelEntry* seltab; int* selhdr; float* rowst;
    int F, K, n_haar, n_color, nb, w0, h0, weak_type, strong_type, feature_type, cch_mode, pad1;
    float lr; int mdch_fast;               // mdch_fast: both groups eligible for the pixel-list kernels
    MdchGroup grp[2];
    uint32_t* mlist;                       // sorted pixel lists of both groups ((bin << 22) | offset in the extended table)
    uint32_t* macc;                        // [n_color][ldT] u32 accumulators (zero between launches)
    uint32_t* mtot;                        // [2] fixed-point total of each group's weight table
    float* ada_cnt; float* alpha;          // AdaBoost: [4][K][F] running TP / FP / TN / FN weights, [K] voting weights
    int* keep; float* ensH;                // MILEnsemble: [F] retained flags (NULL for the other types), [ldT] hypothesis
    float pct_retained; int pad2;          // MILEnsemble: m_percentageOfRetainedWeakClassifiers (already / 100.0f)
};

// Per-call values that change every frame travel as kernel parameters, so that the ObjDesc array itself is
// unchanged from frame to frame and its upload can be skipped (mct_api.cu: desc_publish).
#define DESC_MAX_OBJ 64
struct StepArgs {
    const float* noise_base;       // [sum over objects of 4 * N] Gaussian rows (device)
    float u01[DESC_MAX_OBJ];       // the randfloat() of ResampleParticles per object
    int summ_off = 0;              // the read-outs of this call go to ObjDesc::summ + summ_off (the fused score path alternates between two slots)
    int dev_rng = 0;               // 1: the draw comes from the object's device-resident stream (PfTrk::rng), which advances iff it resamples
};

// per-object arguments of the batched RearrangeParticlesBasedOnGroundLocation launch
struct RearrangeArgs { float mfx[DESC_MAX_OBJ], mfy[DESC_MAX_OBJ]; long long noise_off[DESC_MAX_OBJ]; };

// ------------------------------------------------------------------ error helpers
extern char g_mct_err[512];
static inline int mct_fail(mct_ctx* ctx, int code, const char* fmt, const char* a = "", int b = 0)
{
    char* dst = ctx ? ctx->err : g_mct_err;
    snprintf(dst, 512, fmt, a, b);
    if (ctx) snprintf(g_mct_err, 512, "%s", dst);
    return code;
}
#define MCT_CUDA(ctx, call)                                                                 \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            (void)cudaGetLastError();      /* do not let it resurface at the next launch check */ \
            return mct_fail((ctx), MCT_E_CUDA, "CUDA error: %s (line %d)", cudaGetErrorString(e__), __LINE__); \
        }                                                                                    \
    } while (0)
#define MCT_KERNEL_CHECK(ctx)                                                                \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return mct_fail((ctx), MCT_E_CUDA, "kernel launch failed: %s (line %d)", cudaGetErrorString(e__), __LINE__); \
    } while (0)

// optional per-kernel CUDA-event timing on the launching stream (bench.py roofline); see mct_api.cu
void mct_prof_begin(mct_ctx* ctx, const char* name);
void mct_prof_end(mct_ctx* ctx);
void mct_prof_begin_on(mct_ctx* ctx, const char* name, cudaStream_t st);
void mct_prof_end_on(mct_ctx* ctx, cudaStream_t st);
#define MCT_LAUNCH(ctx, name, ...)                                                           \
    do {                                                                                     \
        mct_prof_begin((ctx), (name));                                                       \
        __VA_ARGS__;                                                                         \
        mct_prof_end((ctx));                                                                 \
        MCT_KERNEL_CHECK(ctx);                                                               \
    } while (0)

#define MCT_LAUNCH_ON(ctx, st, name, ...)                                                    \
    do {                                                                                     \
        mct_prof_begin_on((ctx), (name), (st));                                              \
        __VA_ARGS__;                                                                         \
        mct_prof_end_on((ctx), (st));                                                        \
        MCT_KERNEL_CHECK(ctx);                                                               \
    } while (0)

// Programmatic dependent launch (sm_90+): a consumer launched with this attribute may become resident while its producer (the
// previous kernel of the stream) is still running; it executes up to pdl_wait() -- which returns once the producer has completed
// and its writes are visible -- and the producer allows it with pdl_trigger().  Takes ~1.7 us off a launch boundary
// (tools/pdl_lab.cu).  Only used where the consumer cannot pile up in front of the producer (one CTA per SM, or a handful of CTAs);
// a kernel that calls pdl_wait() without the attribute runs as before.  MCT_NO_PDL=1 launches everything the plain way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static inline bool mct_pdl_enabled() { static int v = -1; if (v < 0) v = getenv("MCT_NO_PDL") ? 0 : 1; return v != 0; }
#define MCT_LAUNCH_PDL(ctx, name, kern, grid, block, smem, ...)                               \
    do {                                                                                     \
        cudaLaunchConfig_t cfg__ = {};                                                       \
        cfg__.gridDim = (grid); cfg__.blockDim = (block); cfg__.dynamicSmemBytes = (smem); cfg__.stream = (ctx)->stream; \
        cudaLaunchAttribute at__[1];                                                         \
        at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                     \
        at__[0].val.programmaticStreamSerializationAllowed = 1;                              \
        cfg__.attrs = at__; cfg__.numAttrs = mct_pdl_enabled() ? 1 : 0;                      \
        mct_prof_begin((ctx), (name));                                                       \
        cudaError_t le__ = cudaLaunchKernelEx(&cfg__, kern, __VA_ARGS__);                    \
        mct_prof_end(ctx);                                                                   \
        if (le__ != cudaSuccess) { (void)cudaGetLastError(); return mct_fail((ctx), MCT_E_CUDA, "kernel launch failed: %s (line %d)", cudaGetErrorString(le__), __LINE__); } \
        MCT_KERNEL_CHECK(ctx);                                                               \
    } while (0)

// Per-device launch state.  Function attributes (opt-in dynamic shared memory, non-portable cluster sizes) and __constant__
// uploads belong to ONE device: every launcher keeps its high-water mark in a table indexed by ctx->device, so ctxs on several
// GPUs of one process are independent (INTEGRATION.md "threads and devices").  The mutex makes check + set atomic for host
// threads that drive different ctxs.
#define MCT_MAX_DEV 64
#include <mutex>
extern std::mutex g_mct_attr_mu;
template <typename KernT>
static inline int mct_smem_attr(mct_ctx* ctx, KernT kern, size_t* high_water /* [MCT_MAX_DEV] */, size_t need, size_t floor_bytes)
{
    std::lock_guard<std::mutex> lock(g_mct_attr_mu);
    size_t& cur = high_water[ctx->device];
    if (cur < floor_bytes) cur = floor_bytes;
    if (need > cur) {
        MCT_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        cur = need;
    }
    return MCT_OK;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ------------------------------------------------------------------ kernel launchers (defined per .cu)
// integral image (+ bin image when an MDCH classifier uses this ctx) of `t` on stream `st`; *with_bins tells whether
// the bin image was produced
int launch_frame_prep(mct_ctx* ctx, const PrepTarget& t, cudaStream_t st, int* with_bins);
// packed BGR -> gray + planar R,G,B (or H,S,V) planes of pitch dpitch
int launch_bgr_unpack(mct_ctx* ctx, const uint8_t* d_bgr, int spitch, int W, int H, int use_hsv, uint8_t* const planes[4], int dpitch,
                      cudaStream_t st);
int launch_bin_image(mct_ctx* ctx, int nb);
int launch_sum_rects(mct_ctx* ctx, const int* d_rects, int n, float* d_out);
// full [F][ld] feature matrix of one box list (FeatureVector::Compute)
int mct_feature_matrix_full(mct_clf* clf, const int* d_col, const int* d_row, const float* d_sx, const float* d_sy,
                            int n, float* d_out, int ld);
// colour rows needed by Classify: the selected bins only (MDCH) / all 11 (CCH)
int mct_feature_color_rows(mct_clf* clf, const int* d_col, const int* d_row, const float* d_sx, const float* d_sy,
                           const int* d_valid, int n, const int* d_outbins, int n_out, float* d_out, int ld);
// MDCH rows of the selected bins for the test boxes of n_obj objects in one batched pass (persistent
// private-histogram kernel + generic fallback for non-conforming boxes); slotmap == NULL marks a non-MDCH object
int launch_build_colormap(mct_clf* clf);
int launch_mdch_selected(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int max_slots, const StepArgs& sa, int pre);
int launch_classify(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int K_max, int from_matrix, int log_ratio);
int launch_classify_staged(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int log_ratio);
int launch_pf_predict(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, const StepArgs& sa);
int launch_pf_testset(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max);
int launch_pf_refine(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj);
int launch_foot_points_exchange(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, const float* d_h0, const P2PPeers& peers, int world, int rank,
                                unsigned seq, unsigned* d_ticket, float* d_out, unsigned long long timeout_ns, int* d_status);
int launch_response_argmax(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, void* d_out);
int launch_pf_rearrange(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, const float* d_noise2, const RearrangeArgs& ra);
int launch_ground_pdf(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, const GroundArgs* d_ga, float* d_wout, GroundOut* d_out);
// fused predict + test set (+ per-frame reset of the MDCH fallback counters) for the batched per-frame path
int launch_pf_predict_testset(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, const StepArgs& sa);
// with_summary: run the read-out (k_pf_summary body) at the end of the same kernel
int launch_pf_update(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int use_H, const StepArgs& sa, int with_summary);
int launch_pf_resample_only(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int force, const StepArgs& sa);
int launch_pf_summary(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max);
int launch_pf_expand_prob(mct_ctx* ctx, const ObjDesc* d_desc, const float* d_prob_compact, int n_prob, int only_nonzero);
int launch_weak_update(mct_clf* clf, int P, int Nn, int has_weights);
int launch_mil_select(mct_clf* clf, int P, int Nn);
// MILEnsemble: RetainBestPerformingWeakClassifiers, in front of the weak update (no-op for the other strong types)
int launch_ensemble_retain(mct_clf* clf, int P, int Nn);
int launch_ensemble_retain_batch(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj);
// batched UpdateClassifier of n_obj classifiers (device descriptor array d, host copy h)
// launch bounds of the training-feature kernels when the set sizes and the colour-histogram plan are decided on the device
// (mct_track_update_default): grids and shared memory from upper bounds, CTAs outside the actual plan leave at once
struct TrainLaunchBounds { int n_max, nch, tab_words, chunk_words, npx_max, n_train_max; };
// weak_too: the caller runs the weak-classifier update next; with both feature families on their own streams it is then queued here,
// split by family (ctx->weak_split_done tells launch_weak_update_batch)
int launch_train_features(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj, const TrainLaunchBounds* lb = nullptr, bool weak_too = false);
// per-object arguments of k_train_gen_default (device array, re-uploaded only when they change)
struct TrainGenDev {
    PfTrk* trk; const PfDev* st;
    int* stage; int mcap;                 // [4][mcap] col | row | sx | sy of the object's training boxes (positives first)
    int cap_pos, cap_neg;
    float w0, h0; int opt;                // m_currentStateList[2], [3]; 0 = highest-weight particle, 1 = average
    float pos_radius; int max_pos;
    float neg_outer, neg_inner; int max_neg;
    float maxs;                           // ParticleFilter::GetMaxScale()
    int frame_w, frame_h;
    int mdch, cw0, ch0, dim;              // colour-histogram rows wanted; classifier blob size; nb^3
    int npx_b[2], tab_b, nch_b;           // what the launch was sized for (a plan beyond it takes the generic kernel)
};
int launch_train_gen_default(mct_ctx* ctx, int n_obj, const TrainGenDev* d_gen, TrainDesc* d_tdesc, mct_train_default_out* d_out, bool fresh_args);
int launch_weak_update_batch(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj);
int launch_weak_update_part(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj, int part, cudaStream_t st);
int launch_mil_select_batch(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj, const int* cluster_req);
int launch_build_colormap_batch(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj);
int launch_foot_points(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, const float* d_h0, float* d_out);
int launch_sample_regions(mct_ctx* ctx, int n, const mct_sample_region* d_reg, int cap, int* d_col, int* d_row, mct_sample_region_out* d_out);
int launch_train_from_particles(mct_ctx* ctx, const ObjDesc* d_desc, const ObjDesc* h_desc, int n_obj, const mct_train_gen* d_gen, int max_pos_max,
                                int cap, int* d_col, int* d_row, float* d_sx, float* d_sy, mct_train_gen_out* d_out);

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
// OnlineStumps/WeightedStumps::ClassifyF (OnlineStumpsWeakClassifier.cpp:83-90): p_c in f32 via expf,
// then f64 logs, result rounded to f32.  Perceptron::ClassifyF (PerceptronWeakClassifier.cpp:78-89).
__device__ __forceinline__ float weak_classify_dev(int weak_type, float x, float mu0, float mu1,
                                                   float n0, float n1, float e0, float e1)
{
    if (weak_type == MCT_WEAK_PERCEPTRON) {
        // rows 0,1 of the state hold m_weightList[0], [1]
        double fv = (double)x;
        return (fv * (double)mu0 > (double)mu1) ? 1.0f : -1.0f;
    }
    float d0 = __fsub_rn(x, mu0), d1 = __fsub_rn(x, mu1);
    float a0 = __fmul_rn(__fmul_rn(d0, d0), e0);
    float a1 = __fmul_rn(__fmul_rn(d1, d1), e1);
    double p0 = (double)__fmul_rn(expf(a0), n0);
    double p1 = (double)__fmul_rn(expf(a1), n1);
    return (float)(log(1e-5 + p1) - log(1e-5 + p0));
}
// the same value with ONE f64 log (of the f64 quotient): identical after rounding to f32 except for sub-ulp ties; used where
// the f64 logs dominate (test-set classification, training predictions)
__device__ __forceinline__ float weak_classify_q_dev(int weak_type, float x, float mu0, float mu1,
                                                     float n0, float n1, float e0, float e1)
{
    if (weak_type == MCT_WEAK_PERCEPTRON) return ((double)x * (double)mu0 > (double)mu1) ? 1.0f : -1.0f;
    const float d0 = __fsub_rn(x, mu0), d1 = __fsub_rn(x, mu1);
    const float a0 = __fmul_rn(__fmul_rn(d0, d0), e0);
    const float a1 = __fmul_rn(__fmul_rn(d1, d1), e1);
    const double p0 = (double)__fmul_rn(expf(a0), n0);
    const double p1 = (double)__fmul_rn(expf(a1), n1);
    return (float)log((1e-5 + p1) / (1e-5 + p0));
}
// AdaBoostClassifier::Classify vote of one selected weak classifier (AdaBoostClassifier.cpp:196-203): +alpha if the weak
// classifier says "object" (OnlineStumps::Classify compares the two class likelihoods, .cpp:68-75; the other weak types take
// the sign of ClassifyF), -alpha otherwise
__device__ __forceinline__ bool weak_classify_bool_dev(int weak_type, float x, float mu0, float mu1, float n0, float n1, float e0, float e1)
{
    if (weak_type == MCT_WEAK_STUMP) {
        const float d0 = __fsub_rn(x, mu0), d1 = __fsub_rn(x, mu1);
        const float p0 = __fmul_rn(expf(__fmul_rn(__fmul_rn(d0, d0), e0)), n0);
        const float p1 = __fmul_rn(expf(__fmul_rn(__fmul_rn(d1, d1), e1)), n1);
        return p1 > p0;
    }
    return weak_classify_dev(weak_type, x, mu0, mu1, n0, n1, e0, e1) > 0.0f;
}
// Public.h:169-172
// (1.0f / y correctly rounded == __frcp_rn(y): same bits as the division, without its generic slow-path check)
__device__ __forceinline__ float sigmoid_dev(float x) { return __frcp_rn(__fadd_rn(1.0f, expf(-x))); }

// ---- per-particle steps shared by the particle-filter kernels and the fused prologue of k_mdch_sel
// ParticleFilter::UpdateParticleWeight scale-ratio clamp (ParticleFilter.cpp:658-665): 1.2 and 0.9 are
// double literals, so the comparison and the product are evaluated in double.
__device__ __forceinline__ void scale_ratio_clamp(float& sx, float& sy)
{
    if ((double)__fdiv_rn(sx, sy) > 1.2) sy = (float)(0.9 * (double)sx);
    else if ((double)__fdiv_rn(sy, sx) > 1.2) sx = (float)(0.9 * (double)sy);
}

// ------------------------------------------------------------------------------------------
// PredictWithBrownianMotion (ParticleFilter.cpp:236-341).  x and y use the scale *before* its update.
__device__ __forceinline__ float scale_step_dev(float s, float n, float msc, float maxs, float mins)
{
    if (msc > fabsf(n)) s = __fadd_rn(s, n);
    else if (n < 0) s = __fsub_rn(s, msc);
    else s = __fadd_rn(s, msc);
    if (s > maxs) s = maxs;
    if (s < mins) s = mins;
    return s;
}

// CheckForParticleRefinement (ParticleFilter.cpp:83-112) looks at the RESAMPLED matrix: a particle counts when its box's top-left
// corner lies inside the image.  When no particle counts, ForceParticleFilterRefinement (:187-226) pulls the resampled centres
// back towards the frame (quirk kept: the vertical correction uses scaleX, RS(row, 2)).
__device__ __forceinline__ bool refine_valid_one(const ObjDesc& d, int i)
{
    const float left = __fsub_rn(d.rcx[i], __fdiv_rn(__fmul_rn(d.rsx[i], d.w0), 2.0f));
    const float top = __fsub_rn(d.rcy[i], __fdiv_rn(__fmul_rn(d.rsy[i], d.h0), 2.0f));
    return !(left < 0 || left >= d.img_w || top < 0 || top >= d.img_h);
}
__device__ __forceinline__ void force_refine_one(const ObjDesc& d, int i)
{
    const float cx = d.rcx[i], cy = d.rcy[i], sx = d.rsx[i], sy = d.rsy[i];
    const float left = __fsub_rn(cx, __fdiv_rn(__fmul_rn(sx, d.w0), 2.0f));
    const float top = __fsub_rn(cy, __fdiv_rn(__fmul_rn(sy, d.h0), 2.0f));
    const float fx = cx, fy = __fadd_rn(cy, __fdiv_rn(__fmul_rn(sy, d.h0), 2.0f));
    const float up = __fsub_rn(fy, __fdiv_rn(__fmul_rn(sx, d.h0), 2.0f));
    if (left < 0) d.rcx[i] = fmaxf(fx, 0.0f);
    if (left >= d.img_w) d.rcx[i] = fminf(fx, __fsub_rn(d.img_w, 1.0f));
    if (top < 0) d.rcy[i] = fmaxf(up, 0.0f);
    if (top >= d.img_h) d.rcy[i] = fminf(up, __fsub_rn(d.img_h, 1.0f));
    d.rw[i] = 1.0f;
}
// The fused per-frame path spreads the check over two kernels: the predict kernel ORs {2 = check pending, 1 = some
// particle counts} into PfDev::refine, k_pf_update (one CTA per object, the next kernel that touches the resampled
// matrix) applies the refinement when the check is pending and nothing counted, and clears the flag.
#define MCT_REFINE_PENDING 2
#define MCT_REFINE_VALID 1
__device__ __forceinline__ void predict_one(const ObjDesc& d, int i, const float* __restrict__ nz)
{
    const int N = d.N;
    float sx = d.sx[i], sy = d.sy[i];
    d.cx[i] = (float)__float2int_rn(__fadd_rn(d.cx[i], __fmul_rn(sx, nz[i])));
    d.cy[i] = (float)__float2int_rn(__fadd_rn(d.cy[i], __fmul_rn(sy, nz[(size_t)N + i])));
    d.sx[i] = scale_step_dev(sx, nz[(size_t)2 * N + i], d.msc, d.maxs, d.mins);
    d.sy[i] = scale_step_dev(sy, nz[(size_t)3 * N + i], d.msc, d.maxs, d.mins);
}
// ------------------------------------------------------------------------------------------
// GenerateTestSampleSet (ParticleFilterTracker.cpp:1247-1305): box per particle; out-of-frame particles
// get UpdateParticleWeight(p, 0) (weight *= 0 and the scale-ratio clamp); in-frame particles with a
// non-zero weight form the test set (dense mask instead of the reference's compaction).
__device__ __forceinline__ void testset_one(const ObjDesc& d, int i, int fw, int fh)
{
    float sx = d.sx[i], sy = d.sy[i];
    const float wf = __fmul_rn(sx, d.w0), hf = __fmul_rn(sy, d.h0);
    const float left = (float)__float2int_rn(__fsub_rn(d.cx[i], __fdiv_rn(wf, 2.0f)));
    const float top = (float)__float2int_rn(__fsub_rn(d.cy[i], __fdiv_rn(hf, 2.0f)));
    const float width = (float)__float2int_rn(wf), height = (float)__float2int_rn(hf);
    int v = 0;
    if (left <= 0 || __fadd_rn(left, width) >= (float)(fw - 1) || top <= 0 || __fadd_rn(top, height) >= (float)(fh - 1)) {
        d.w[i] = __fmul_rn(0.0f, d.w[i]);
        scale_ratio_clamp(sx, sy);
        d.sx[i] = sx; d.sy[i] = sy;
    } else if (d.w[i] != 0) {
        v = 1;
    }
    d.col[i] = (int)left; d.row[i] = (int)top;
    d.valid[i] = v;
}

// The descriptor of this CTA's object, copied once into shared memory: every `d.field` through the global array is an
// L2 round trip in front of the access it feeds, which is what bounds the small latency-bound kernels.
__device__ __forceinline__ const ObjDesc& load_desc(const ObjDesc* __restrict__ g, ObjDesc* s)
{
    const int n = (int)(sizeof(ObjDesc) / sizeof(int));
    for (int i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<int*>(s)[i] = __ldg(reinterpret_cast<const int*>(g) + i);
    __syncthreads();
    return *s;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Matrixu::sumRect (Matrix.cpp:49-67) on the padded integral image; coordinates are clamped to the
// table so that an out-of-frame rect (reference: debug assert / UB) cannot fault.
__device__ __forceinline__ float sum_rect_dev(const float* __restrict__ ii, int iip, int W, int H, int x, int y, int w, int h)
{
    int x0 = clampi(x, 0, W), x1 = clampi(x + w, 0, W);
    int y0 = clampi(y, 0, H), y1 = clampi(y + h, 0, H);
    const float* r0 = ii + (size_t)y0 * iip;
    const float* r1 = ii + (size_t)y1 * iip;
    float tl = __ldg(r0 + x0), tr = __ldg(r0 + x1), br = __ldg(r1 + x1), bl = __ldg(r1 + x0);
    return __fsub_rn(__fsub_rn(__fadd_rn(br, tl), tr), bl);
}
// HaarFeature::Compute (HaarFeature.cpp:111-141)
__device__ __forceinline__ float haar_eval_dev(const float* __restrict__ ii, int iip, int W, int H,
                                               const int4* rects, const float* wts, int nrect,
                                               int col, int row, float sx, float sy)
{
    float sum = 0.0f;
    for (int k = 0; k < nrect; k++) {
        int4 r = rects[k];
        int x = __float2int_rn(__fmul_rn((float)r.x, sx)) + col;
        int y = __float2int_rn(__fmul_rn((float)r.y, sy)) + row;
        int h = __float2int_rn(__fmul_rn((float)r.w, sy));     // .w = height
        int w = __float2int_rn(__fmul_rn((float)r.z, sx));     // .z = width
        sum = __fadd_rn(sum, __fmul_rn(wts[k], sum_rect_dev(ii, iip, W, H, x, y, w, h)));
    }
    return __fdiv_rn(sum, __fmul_rn(sx, sy));
}

// ---- the same on a region of the integral image staged in shared memory (origin ox, oy; pitch rp)
__device__ __forceinline__ float sum_rect_region(const float* __restrict__ reg, int rp, int ox, int oy, int W, int H,
                                                 int x, int y, int w, int h)
{
    const int x0 = clampi(x, 0, W) - ox, x1 = clampi(x + w, 0, W) - ox;
    const int y0 = clampi(y, 0, H) - oy, y1 = clampi(y + h, 0, H) - oy;
    const float* r0 = reg + y0 * rp;
    const float* r1 = reg + y1 * rp;
    const float tl = r0[x0], tr = r0[x1], br = r1[x1], bl = r1[x0];
    return __fsub_rn(__fsub_rn(__fadd_rn(br, tl), tr), bl);
}
__device__ __forceinline__ float haar_eval_region(const float* __restrict__ reg, int rp, int ox, int oy, int W, int H,
                                                  const int4* rects, const float* wts, int nrect,
                                                  int col, int row, float sx, float sy)
{
    float sum = 0.0f;
    for (int k = 0; k < nrect; k++) {
        const int4 r = rects[k];
        const int x = __float2int_rn(__fmul_rn((float)r.x, sx)) + col;
        const int y = __float2int_rn(__fmul_rn((float)r.y, sy)) + row;
        const int h = __float2int_rn(__fmul_rn((float)r.w, sy));
        const int w = __float2int_rn(__fmul_rn((float)r.z, sx));
        sum = __fadd_rn(sum, __fmul_rn(wts[k], sum_rect_region(reg, rp, ox, oy, W, H, x, y, w, h)));
    }
    return __fdiv_rn(sum, __fmul_rn(sx, sy));
}

// ---- TMA 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}


__device__ __forceinline__ float warp_sum_f(float v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif
