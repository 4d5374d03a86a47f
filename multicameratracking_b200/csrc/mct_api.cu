// mct_api.cu -- C-ABI entry points (include/mct_b200.h): context, frame, classifier, particle filter and
// the fused per-frame path.  Host logic only; every compute step is a CUDA kernel in k_*.cu.
// There is no CPU fallback anywhere in this library.
#include "mct_internal.cuh"
#include <new>
#include <vector>

char g_mct_err[512] = "";
std::mutex g_mct_attr_mu;

#define DESC_RING 32

struct DescRing {
    ObjDesc* h;            // pinned [DESC_RING][DESC_MAX_OBJ]
    ObjDesc* d;            // device [DESC_RING][DESC_MAX_OBJ]
    cudaEvent_t ev[DESC_RING];
    int next;
};

// ---- optional per-kernel timing: CUDA events recorded on the launching stream around every launch ----
struct ProfEntry { const char* name; cudaEvent_t a, b; };
struct ProfState {
    bool on;
    std::vector<ProfEntry> entries;      // recorded launches since the last read
    std::vector<cudaEvent_t> pool;       // reusable events
    const char* cur;
    cudaEvent_t cur_a;
};
static cudaEvent_t prof_event(ProfState* ps)
{
    if (!ps->pool.empty()) { cudaEvent_t e = ps->pool.back(); ps->pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void mct_prof_begin_on(mct_ctx* ctx, const char* name, cudaStream_t st)
{
    ProfState* ps = (ProfState*)ctx->prof;
    if (!ps || !ps->on) return;
    ps->cur = name; ps->cur_a = prof_event(ps);
    cudaEventRecord(ps->cur_a, st);
}
void mct_prof_end_on(mct_ctx* ctx, cudaStream_t st)
{
    ProfState* ps = (ProfState*)ctx->prof;
    if (!ps || !ps->on || !ps->cur) return;
    cudaEvent_t b = prof_event(ps);
    cudaEventRecord(b, st);
    ps->entries.push_back({ps->cur, ps->cur_a, b});
    ps->cur = nullptr;
}
void mct_prof_begin(mct_ctx* ctx, const char* name) { mct_prof_begin_on(ctx, name, ctx->stream); }
void mct_prof_end(mct_ctx* ctx) { mct_prof_end_on(ctx, ctx->stream); }

static DescRing* ring_of(mct_ctx* ctx) { return (DescRing*)ctx->h_desc; }

// Reserve a descriptor slot: returns host pointer to fill; call desc_commit afterwards.
static int desc_acquire(mct_ctx* ctx, ObjDesc** h, ObjDesc** d, int* slot)
{
    DescRing* r = ring_of(ctx);
    int s = r->next;
    r->next = (r->next + 1) % DESC_RING;
    MCT_CUDA(ctx, cudaEventSynchronize(r->ev[s]));
    *h = r->h + (size_t)s * DESC_MAX_OBJ;
    *d = r->d + (size_t)s * DESC_MAX_OBJ;
    *slot = s;
    return MCT_OK;
}
static int desc_commit(mct_ctx* ctx, int slot, int n)
{
    DescRing* r = ring_of(ctx);
    MCT_CUDA(ctx, cudaMemcpyAsync(r->d + (size_t)slot * DESC_MAX_OBJ, r->h + (size_t)slot * DESC_MAX_OBJ,
                                  (size_t)n * sizeof(ObjDesc), cudaMemcpyHostToDevice, ctx->stream));
    MCT_CUDA(ctx, cudaEventRecord(r->ev[slot], ctx->stream));
    return MCT_OK;
}

// Publish n descriptors (built by the caller in `src`) and return their device address.  The array is uploaded
// only when it differs from the last published one: per-frame values (noise, u01) travel as kernel parameters
// (StepArgs), so in steady state no descriptor copy is issued at all.  A published set lives in a ring slot that
// is only reused 32 publishes later, after its copy event has completed.
static int desc_publish(mct_ctx* ctx, const ObjDesc* src, int n, const ObjDesc** d_out)
{
    if (ctx->desc_last_d && ctx->desc_last_n == n && !memcmp(ctx->desc_last_h, src, (size_t)n * sizeof(ObjDesc))) {
        *d_out = (const ObjDesc*)ctx->desc_last_d;
        return MCT_OK;
    }
    ObjDesc *h, *d; int slot, rc;
    if ((rc = desc_acquire(ctx, &h, &d, &slot))) return rc;
    memcpy(h, src, (size_t)n * sizeof(ObjDesc));
    if ((rc = desc_commit(ctx, slot, n))) return rc;
    memcpy(ctx->desc_last_h, src, (size_t)n * sizeof(ObjDesc));
    ctx->desc_last_d = d; ctx->desc_last_n = n;
    *d_out = d;
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------ context
// Staging buffers of UpdateClassifier parked between contexts: a harness that re-creates its cameras (bench.py's network legs build
// fresh trackers every few frames) otherwise pays the pinned / device allocations of the first batched Update inside a frame
// -- about a millisecond per context with peer mappings in place (one camera per process).  One parked set per device.
#include <mutex>
struct StageCache { int* h_stage; int* d_stage; size_t cap; TrainDesc* h_tdesc; TrainDesc* d_tdesc; };
static StageCache g_stage_cache[MCT_MAX_DEV];
static std::mutex g_stage_mu;

extern "C" const char* mct_version(void) { return "mct-b200 0.1 (sm_100a)"; }
// sizeof of every public struct, in the order of their declaration in mct_b200.h: lets a binding check its own layout
// (tests/test_abi.py compares the ctypes mirrors in capi.py) without touching a device
extern "C" int mct_abi_struct_sizes(int* out, int cap)
{
    const int v[] = {(int)sizeof(mct_boxes), (int)sizeof(mct_clf_params), (int)sizeof(mct_haar_table), (int)sizeof(mct_pf_summary),
                     (int)sizeof(mct_track_params), (int)sizeof(mct_ground_pdf_result), (int)sizeof(mct_train_gen), (int)sizeof(mct_train_gen_out),
                     (int)sizeof(mct_sample_region), (int)sizeof(mct_sample_region_out), (int)sizeof(mct_train_default),
                     (int)sizeof(mct_train_default_out)};
    const int n = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < n && i < cap && out; i++) out[i] = v[i];
    return n;
}
extern "C" const char* mct_last_error(mct_ctx* ctx) { return ctx ? ctx->err : g_mct_err; }
extern "C" long long mct_launch_count(mct_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* mct_ctx_stream(mct_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int mct_ctx_create(int device, void* cuda_stream, mct_ctx** out)
{
    if (!out) return mct_fail(nullptr, MCT_E_ARG, "mct_ctx_create: out is NULL%s");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return mct_fail(nullptr, MCT_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                        cudaGetErrorString(e));
    if (device < 0 || device >= ndev || device >= MCT_MAX_DEV) return mct_fail(nullptr, MCT_E_ARG, "device index out of range%s (%d)", "", device);
    MCT_CUDA(nullptr, cudaSetDevice(device));
    mct_ctx* ctx = new (std::nothrow) mct_ctx();
    if (!ctx) return mct_fail(nullptr, MCT_E_NOMEM, "out of host memory%s");
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    MCT_CUDA(nullptr, cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device));
    ctx->binimg_frame = -1;
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else { MCT_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    DescRing* r = new (std::nothrow) DescRing();
    if (!r) return mct_fail(nullptr, MCT_E_NOMEM, "out of host memory%s");
    MCT_CUDA(nullptr, cudaMallocHost(&r->h, sizeof(ObjDesc) * DESC_RING * DESC_MAX_OBJ));
    MCT_CUDA(nullptr, cudaMalloc(&r->d, sizeof(ObjDesc) * DESC_RING * DESC_MAX_OBJ));
    for (int i = 0; i < DESC_RING; i++) MCT_CUDA(nullptr, cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming));
    r->next = 0;
    ctx->h_desc = r;
    ctx->desc_last_h = malloc(sizeof(ObjDesc) * DESC_MAX_OBJ);
    if (!ctx->desc_last_h) return mct_fail(nullptr, MCT_E_NOMEM, "out of host memory%s");
    ctx->summ_cap = DESC_MAX_OBJ;
    // [0, 64) and [64, 128): the two alternating read-out slots of the fused score path (two calls may be in flight);
    // [128, 192): the stand-alone read-out calls (mct_pf_get_summary_many)
    MCT_CUDA(nullptr, cudaHostAlloc(&ctx->h_summ, sizeof(mct_pf_summary) * DESC_MAX_OBJ * 3, cudaHostAllocMapped));
    MCT_CUDA(nullptr, cudaHostGetDevicePointer((void**)&ctx->h_summ_dev, ctx->h_summ, 0));
    memset(ctx->h_summ, 0, sizeof(mct_pf_summary) * DESC_MAX_OBJ * 3);
    MCT_CUDA(nullptr, cudaMalloc(&ctx->d_scored, sizeof(unsigned long long)));
    MCT_CUDA(nullptr, cudaMemset(ctx->d_scored, 0, sizeof(unsigned long long)));
    *out = ctx;
    return MCT_OK;
}

extern "C" int mct_profile_enable(mct_ctx* ctx, int on)
{
    if (!ctx) return MCT_E_ARG;
    if (!ctx->prof) { ProfState* ps = new (std::nothrow) ProfState(); if (!ps) return MCT_E_NOMEM; ps->on = false; ps->cur = nullptr; ctx->prof = ps; }
    ((ProfState*)ctx->prof)->on = on != 0;
    return MCT_OK;
}
extern "C" int mct_profile_read(mct_ctx* ctx, int max_kernels, char* names, float* total_ms, int* launches, int* n_out)
{
    if (!ctx || !names || !total_ms || !launches || !n_out) return MCT_E_ARG;
    *n_out = 0;
    ProfState* ps = (ProfState*)ctx->prof;
    if (!ps) return MCT_OK;
    MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->prep_stream) MCT_CUDA(ctx, cudaStreamSynchronize(ctx->prep_stream));
    int n = 0;
    for (ProfEntry& e : ps->entries) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        int k = -1;
        for (int i = 0; i < n; i++) if (!strncmp(names + (size_t)i * 48, e.name, 47)) { k = i; break; }
        if (k < 0 && n < max_kernels) { k = n++; snprintf(names + (size_t)k * 48, 48, "%s", e.name); total_ms[k] = 0; launches[k] = 0; }
        if (k >= 0) { total_ms[k] += ms; launches[k] += 1; }
        ps->pool.push_back(e.a); ps->pool.push_back(e.b);
    }
    ps->entries.clear();
    *n_out = n;
    return MCT_OK;
}
extern "C" int mct_scored_particles(mct_ctx* ctx, unsigned long long* out, int reset)
{
    if (!ctx || !out) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_scored, sizeof(*out), cudaMemcpyDeviceToHost, ctx->stream));
    MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (reset) MCT_CUDA(ctx, cudaMemsetAsync(ctx->d_scored, 0, sizeof(*out), ctx->stream));
    return MCT_OK;
}

extern "C" int mct_ctx_destroy(mct_ctx* ctx)
{
    if (!ctx) return MCT_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->prep_stream) cudaStreamSynchronize(ctx->prep_stream);
    if (ctx->d_scored) cudaFree(ctx->d_scored);
    mct_p2p_destroy(ctx);
    if (ctx->d_ground_out) cudaFree(ctx->d_ground_out);
    if (ctx->d_ground_args) cudaFree(ctx->d_ground_args);
    if (ctx->d_argmax) cudaFree(ctx->d_argmax);
    if (ctx->d_mlist) cudaFree(ctx->d_mlist);
    if (ctx->d_macc) cudaFree(ctx->d_macc);
    if (ctx->d_mtot) cudaFree(ctx->d_mtot);
    if (ctx->prof) {
        ProfState* ps = (ProfState*)ctx->prof;
        for (ProfEntry& e : ps->entries) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        for (cudaEvent_t e : ps->pool) cudaEventDestroy(e);
        delete ps;
    }
    for (int i = 0; i < 4; i++) if (ctx->own_planes[i]) cudaFree(ctx->own_planes[i]);
    if (ctx->prep_stream) {
        cudaStreamSynchronize(ctx->prep_stream); cudaStreamDestroy(ctx->prep_stream);
        cudaEventDestroy(ctx->ev_ready); cudaEventDestroy(ctx->ev_mark[0]); cudaEventDestroy(ctx->ev_mark[1]); cudaEventDestroy(ctx->ev_gather); cudaEventDestroy(ctx->ev_noise);
    }
    for (int i = 0; i < 4; i++) if (ctx->back.planes[i]) cudaFree(ctx->back.planes[i]);
    if (ctx->back.ii_base) cudaFree(ctx->back.ii_base);
    if (ctx->back.binimg) cudaFree(ctx->back.binimg);
    if (ctx->ii_base) cudaFree(ctx->ii_base);
    if (ctx->bandsum) cudaFree(ctx->bandsum);
    if (ctx->prep_flags) cudaFree(ctx->prep_flags);
    if (ctx->binimg) cudaFree(ctx->binimg);
    if (ctx->d_tmp) cudaFree(ctx->d_tmp);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->h_summ) cudaFreeHost(ctx->h_summ);
    if (ctx->d_noise_stage) cudaFree(ctx->d_noise_stage);
    for (int i = 0; i < 2; i++) if (ctx->d_noise_pf[i]) cudaFree(ctx->d_noise_pf[i]);
    if (ctx->d_bgr_stage) cudaFree(ctx->d_bgr_stage);
    {
        std::lock_guard<std::mutex> lk(g_stage_mu);
        StageCache& sc = g_stage_cache[ctx->device < MCT_MAX_DEV ? ctx->device : 0];
        if (ctx->h_tdesc) {
            if (!sc.h_tdesc && ctx->device < MCT_MAX_DEV) { sc.h_tdesc = ctx->h_tdesc; sc.d_tdesc = ctx->d_tdesc; }
            else { cudaFreeHost(ctx->h_tdesc); cudaFree(ctx->d_tdesc); }
            cudaEventDestroy(ctx->ev_train);
        }
        if (ctx->h_train_stage && ctx->d_train_stage && !sc.h_stage && ctx->device < MCT_MAX_DEV) {
            sc.h_stage = ctx->h_train_stage; sc.d_stage = ctx->d_train_stage; sc.cap = ctx->train_stage_cap;
            ctx->h_train_stage = nullptr; ctx->d_train_stage = nullptr;
        }
    }
    if (ctx->train_stream) {
        cudaStreamSynchronize(ctx->train_stream); cudaStreamDestroy(ctx->train_stream);
        cudaEventDestroy(ctx->ev_tfork); cudaEventDestroy(ctx->ev_tjoin);
    }
    if (ctx->h_train_stage) cudaFreeHost(ctx->h_train_stage);
    if (ctx->d_train_stage) cudaFree(ctx->d_train_stage);
    if (ctx->d_gen) cudaFree(ctx->d_gen);
    if (ctx->h_gen_out) cudaFreeHost(ctx->h_gen_out);
    free(ctx->gen_shadow); free(ctx->tdesc_shadow);
    if (ctx->has_ev_score) { cudaEventDestroy(ctx->ev_score[0]); cudaEventDestroy(ctx->ev_score[1]); }
    free(ctx->desc_last_h);
    DescRing* r = ring_of(ctx);
    if (r) {
        for (int i = 0; i < DESC_RING; i++) cudaEventDestroy(r->ev[i]);
        cudaFreeHost(r->h); cudaFree(r->d);
        delete r;
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MCT_OK;
}

extern "C" int mct_sync(mct_ctx* ctx)
{
    if (!ctx) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCT_OK;
}

static int ensure_tmp(mct_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->tmp_cap) return MCT_OK;
    if (ctx->d_tmp) { MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_tmp); }
    MCT_CUDA(ctx, cudaMalloc(&ctx->d_tmp, bytes));
    ctx->tmp_cap = bytes;
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------ frame
// Two sets of per-frame buffers (planes, integral image, bin image).  Frame t+1 is uploaded and prepared on
// prep_stream into the BACK set while the compute stream is still busy with frame t (its particle-filter update
// occupies only n_obj SMs), then the sets are swapped and the compute stream waits for ev_ready.  A set is only
// overwritten after the compute-stream marker recorded at the frame_set call that retired it.
static int frame_streams(mct_ctx* ctx)
{
    if (ctx->prep_stream) return MCT_OK;
    MCT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->prep_stream, cudaStreamNonBlocking));
    MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
    MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_mark[0], cudaEventDisableTiming));
    MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_mark[1], cudaEventDisableTiming));
    MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_gather, cudaEventDisableTiming));
    MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_noise, cudaEventDisableTiming));
    return MCT_OK;
}
static void frame_swap(mct_ctx* ctx)
{
    FrameSet cur;
    for (int i = 0; i < 4; i++) cur.planes[i] = ctx->own_planes[i];
    cur.plane_cap = ctx->own_plane_cap; cur.ii_base = ctx->ii_base; cur.ii_cap = ctx->ii_cap;
    cur.binimg = ctx->binimg; cur.binimg_cap = ctx->binimg_cap;
    for (int i = 0; i < 4; i++) ctx->own_planes[i] = ctx->back.planes[i];
    ctx->own_plane_cap = ctx->back.plane_cap; ctx->ii_base = ctx->back.ii_base; ctx->ii_cap = ctx->back.ii_cap;
    ctx->binimg = ctx->back.binimg; ctx->binimg_cap = ctx->back.binimg_cap;
    ctx->back = cur;
}
// grow the buffers of one set (rare: first frames / geometry change); growth waits for the whole device
static int frame_ensure(mct_ctx* ctx, uint8_t** planes, size_t* plane_cap, float** ii_base, size_t* ii_cap, uint16_t** binimg,
                        size_t* binimg_cap, int W, int H, bool host_planes)
{
    const size_t pneed = (size_t)round_up(W, 16) * H;
    const size_t ineed = (size_t)MCT_II_LEAD + (size_t)(H + 1) * round_up(W + 1, 4) + 4;
    const size_t bneed = (size_t)W * H;
    const bool grow = (host_planes && pneed > *plane_cap) || ineed > *ii_cap || (ctx->binimg_nb > 0 && bneed > *binimg_cap);
    if (!grow) return MCT_OK;
    MCT_CUDA(ctx, cudaDeviceSynchronize());
    if (host_planes && pneed > *plane_cap) {
        for (int i = 0; i < 4; i++) {
            if (planes[i]) cudaFree(planes[i]);
            MCT_CUDA(ctx, cudaMalloc(&planes[i], pneed));
        }
        *plane_cap = pneed;
    }
    if (ineed > *ii_cap) {
        if (*ii_base) cudaFree(*ii_base);
        MCT_CUDA(ctx, cudaMalloc(ii_base, ineed * sizeof(float)));
        *ii_cap = ineed;
    }
    if (ctx->binimg_nb > 0 && bneed > *binimg_cap) {
        if (*binimg) cudaFree(*binimg);
        MCT_CUDA(ctx, cudaMalloc(binimg, bneed * sizeof(uint16_t)));
        *binimg_cap = bneed;
    }
    return MCT_OK;
}

// Frame prep of the prefetched (back) set.  It is deferred until the gather kernels of the current frame have been
// launched (mct_track_score_async calls this right after the classify kernel), so that it runs beside the
// particle-filter update, which occupies only n_obj SMs, instead of competing with the one-CTA-per-SM kernels.
// Host -> device upload of one frame into `planes` on stream st: four planes, or the packed BGR image followed by the
// unpack kernel (kind 1).  For device-resident BGR the unpack reads the caller's buffer directly.
static int frame_upload(mct_ctx* ctx, int kind, const uint8_t* const src[4], int W, int H, int pitch, int is_device, int use_hsv,
                        uint8_t* const planes[4], cudaStream_t st)
{
    const int dpp = round_up(W, 16);
    if (kind == 0) {
        if (is_device) return MCT_OK;
        for (int i = 0; i < 4; i++)
            if (src[i]) MCT_CUDA(ctx, cudaMemcpy2DAsync(planes[i], dpp, src[i], pitch, W, H, cudaMemcpyHostToDevice, st));
        return MCT_OK;
    }
    const uint8_t* d_bgr = src[0];
    int spitch = pitch;
    if (!is_device) {
        spitch = round_up(3 * W, 16);
        const size_t need = (size_t)spitch * H;
        if (need > ctx->bgr_stage_cap) {
            MCT_CUDA(ctx, cudaDeviceSynchronize());
            if (ctx->d_bgr_stage) cudaFree(ctx->d_bgr_stage);
            MCT_CUDA(ctx, cudaMalloc(&ctx->d_bgr_stage, need));
            ctx->bgr_stage_cap = need;
        }
        if (pitch == spitch) MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_bgr_stage, src[0], (size_t)spitch * H, cudaMemcpyHostToDevice, st));
        else MCT_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_bgr_stage, spitch, src[0], pitch, (size_t)3 * W, H, cudaMemcpyHostToDevice, st));
        d_bgr = ctx->d_bgr_stage;
    }
    return launch_bgr_unpack(ctx, d_bgr, spitch, W, H, use_hsv, planes, dpp, st);
}

// Upload of the prefetched planes into the back set.  Also deferred: mct_track_score_async issues it right AFTER the
// (small, latency-critical) noise copy of the current frame so that the 1.2 MB of planes never sit in front of it.
static int prefetch_copy(mct_ctx* ctx)
{
    if (!ctx->pf_pending || ctx->pf_copied) return MCT_OK;
    const int W = ctx->pf_W, H = ctx->pf_H, dpp = round_up(W, 16);
    // the back set was the current set until the latest mct_frame_set: its readers are covered by the latest marker
    MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->prep_stream, ctx->ev_mark[ctx->mark_idx], 0));
    (void)dpp;
    int rc = frame_upload(ctx, ctx->pf_kind, ctx->pf_src, W, H, ctx->pf_pitch, 0, ctx->pf_hsv, ctx->back.planes, ctx->prep_stream);
    if (rc) return rc;
    ctx->pf_copied = 1;
    return MCT_OK;
}
static int prefetch_prep(mct_ctx* ctx)
{
    if (!ctx->pf_pending || ctx->pf_prepped) return MCT_OK;
    int rc0 = prefetch_copy(ctx);
    if (rc0) return rc0;
    const int W = ctx->pf_W, H = ctx->pf_H, dpp = round_up(W, 16), iip = round_up(W + 1, 4);
    const bool bgr = ctx->pf_kind == 1;
    PrepTarget t = {ctx->back.planes[0], (bgr || ctx->pf_src[1]) ? ctx->back.planes[1] : nullptr, (bgr || ctx->pf_src[2]) ? ctx->back.planes[2] : nullptr,
                    (bgr || ctx->pf_src[3]) ? ctx->back.planes[3] : nullptr, W, H, dpp, ctx->back.ii_base, iip, ctx->back.binimg};
    int with_bins = 0, rc;
    if (ctx->has_gather) MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->prep_stream, ctx->ev_gather, 0));
    if ((rc = launch_frame_prep(ctx, t, ctx->prep_stream, &with_bins))) return rc;
    MCT_CUDA(ctx, cudaEventRecord(ctx->ev_ready, ctx->prep_stream));
    ctx->pf_with_bins = with_bins; ctx->pf_prepped = 1;
    return MCT_OK;
}

static int frame_set_impl(mct_ctx* ctx, int kind, const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2,
                          int W, int H, int pitch, int is_device, int use_hsv)
{
    if (!ctx) return MCT_E_ARG;
    if (!gray || W < 4 || H < 4 || pitch < (kind ? 3 * W : W)) return mct_fail(ctx, MCT_E_ARG, "mct_frame_set: bad frame geometry%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = frame_streams(ctx))) return rc;
    const uint8_t* src[4] = {gray, c0, c1, c2};
    const int dpp = round_up(W, 16), iip = round_up(W + 1, 4);
    // marker: everything enqueued on the compute stream so far (all readers of the current set) ...
    cudaEvent_t old_mark = ctx->ev_mark[ctx->mark_idx];
    ctx->mark_idx ^= 1;
    MCT_CUDA(ctx, cudaEventRecord(ctx->ev_mark[ctx->mark_idx], ctx->stream));
    const bool adopt = !is_device && ctx->pf_pending && ctx->pf_kind == kind && ctx->pf_hsv == use_hsv && ctx->pf_W == W && ctx->pf_H == H &&
                       ctx->pf_pitch == pitch && ctx->pf_src[0] == gray && ctx->pf_src[1] == c0 && ctx->pf_src[2] == c1 && ctx->pf_src[3] == c2;
    const bool own = kind == 1 || !is_device;               // planes live in the ctx's own buffers
    int with_bins = 0;
    if (adopt) {
        // the back set already holds this frame, uploaded by mct_frame_prefetch (and normally prepared beside the
        // previous frame's particle-filter update)
        if ((rc = prefetch_prep(ctx))) return rc;
        frame_swap(ctx);
        with_bins = ctx->pf_with_bins;
        ctx->pf_pending = 0;
    } else {
        ctx->pf_pending = 0;
        // ... the back set was retired one frame_set call ago: its readers are covered by old_mark
        if ((rc = frame_ensure(ctx, ctx->back.planes, &ctx->back.plane_cap, &ctx->back.ii_base, &ctx->back.ii_cap, &ctx->back.binimg,
                               &ctx->back.binimg_cap, W, H, own)))
            return rc;
        frame_swap(ctx);
        MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->prep_stream, old_mark, 0));
        // do not compete with the gather kernels of the frame in flight: start beside its particle-filter update
        if (ctx->has_gather) MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->prep_stream, ctx->ev_gather, 0));
        // is_device == 1: the caller's device planes are ordered after the work already enqueued on the compute stream
        if (is_device == 1) MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->prep_stream, ctx->ev_mark[ctx->mark_idx], 0));
        if ((rc = frame_upload(ctx, kind, src, W, H, pitch, is_device, use_hsv, ctx->own_planes, ctx->prep_stream))) return rc;
    }
    if (!own) {
        ctx->gray = gray; ctx->c0 = c0; ctx->c1 = c1; ctx->c2 = c2;
        ctx->pitch = pitch;
    } else {
        ctx->gray = ctx->own_planes[0];
        ctx->c0 = (kind == 1 || c0) ? ctx->own_planes[1] : nullptr;
        ctx->c1 = (kind == 1 || c1) ? ctx->own_planes[2] : nullptr;
        ctx->c2 = (kind == 1 || c2) ? ctx->own_planes[3] : nullptr;
        ctx->pitch = dpp;
    }
    ctx->W = W; ctx->H = H; ctx->ii_pitch = iip;
    ctx->frame_id++;
    if (!adopt) {
        PrepTarget t = {ctx->gray, ctx->c0, ctx->c1, ctx->c2, W, H, ctx->pitch, ctx->ii_base, iip, ctx->binimg};
        if ((rc = launch_frame_prep(ctx, t, ctx->prep_stream, &with_bins))) return rc;
        MCT_CUDA(ctx, cudaEventRecord(ctx->ev_ready, ctx->prep_stream));
    }
    ctx->binimg_frame = with_bins ? ctx->frame_id : -1;
    MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_ready, 0));
    return MCT_OK;
}

extern "C" int mct_frame_set(mct_ctx* ctx, const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2,
                             int W, int H, int pitch, int is_device)
{
    return frame_set_impl(ctx, 0, gray, c0, c1, c2, W, H, pitch, is_device, 0);
}
extern "C" int mct_frame_set_bgr(mct_ctx* ctx, const uint8_t* bgr, int W, int H, int pitch_bytes, int use_hsv, int is_device)
{
    return frame_set_impl(ctx, 1, bgr, nullptr, nullptr, nullptr, W, H, pitch_bytes, is_device, use_hsv ? 1 : 0);
}

static int frame_prefetch_impl(mct_ctx* ctx, int kind, const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2,
                               int W, int H, int pitch, int use_hsv)
{
    if (!ctx) return MCT_E_ARG;
    if (!gray || W < 4 || H < 4 || pitch < (kind ? 3 * W : W)) return mct_fail(ctx, MCT_E_ARG, "mct_frame_prefetch: bad frame geometry%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = frame_streams(ctx))) return rc;
    if ((rc = frame_ensure(ctx, ctx->back.planes, &ctx->back.plane_cap, &ctx->back.ii_base, &ctx->back.ii_cap, &ctx->back.binimg,
                           &ctx->back.binimg_cap, W, H, true)))
        return rc;
    const uint8_t* src[4] = {gray, c0, c1, c2};
    for (int i = 0; i < 4; i++) ctx->pf_src[i] = src[i];
    ctx->pf_W = W; ctx->pf_H = H; ctx->pf_pitch = pitch; ctx->pf_pending = 1; ctx->pf_with_bins = 0; ctx->pf_prepped = 0; ctx->pf_copied = 0;
    ctx->pf_kind = kind; ctx->pf_hsv = use_hsv;
    return MCT_OK;      // copy and frame prep follow in mct_track_score_async (or at the adopting mct_frame_set)
}
extern "C" int mct_frame_prefetch(mct_ctx* ctx, const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2,
                                  int W, int H, int pitch)
{
    return frame_prefetch_impl(ctx, 0, gray, c0, c1, c2, W, H, pitch, 0);
}
extern "C" int mct_frame_prefetch_bgr(mct_ctx* ctx, const uint8_t* bgr_host, int W, int H, int pitch_bytes, int use_hsv)
{
    return frame_prefetch_impl(ctx, 1, bgr_host, nullptr, nullptr, nullptr, W, H, pitch_bytes, use_hsv ? 1 : 0);
}

// issue the announced frame's upload and preparation NOW (instead of at the end of the next mct_track_score*): for callers that
// queue frame t + 1 while frame t is still being tracked
extern "C" int mct_frame_prefetch_flush(mct_ctx* ctx)
{
    if (!ctx) return MCT_E_ARG;
    if (!ctx->pf_pending) return MCT_OK;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = prefetch_copy(ctx))) return rc;
    return prefetch_prep(ctx);
}
extern "C" int mct_frame_get_planes(mct_ctx* ctx, uint8_t* gray, uint8_t* c0, uint8_t* c1, uint8_t* c2)
{
    if (!ctx) return MCT_E_ARG;
    if (ctx->frame_id == 0) return mct_fail(ctx, MCT_E_STATE, "no frame set%s");
    uint8_t* dst[4] = {gray, c0, c1, c2};
    const uint8_t* src[4] = {ctx->gray, ctx->c0, ctx->c1, ctx->c2};
    for (int i = 0; i < 4; i++)
        if (dst[i] && src[i])
            MCT_CUDA(ctx, cudaMemcpy2DAsync(dst[i], ctx->W, src[i], ctx->pitch, ctx->W, ctx->H, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

extern "C" int mct_frame_get_integral(mct_ctx* ctx, float* out_host)
{
    if (!ctx || !out_host) return MCT_E_ARG;
    if (!ctx->ii_base || ctx->frame_id == 0) return mct_fail(ctx, MCT_E_STATE, "no frame set%s");
    MCT_CUDA(ctx, cudaMemcpy2DAsync(out_host, (size_t)(ctx->W + 1) * sizeof(float), ctx->ii_base + MCT_II_LEAD,
                                    (size_t)ctx->ii_pitch * sizeof(float), (size_t)(ctx->W + 1) * sizeof(float),
                                    ctx->H + 1, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

extern "C" int mct_sum_rects(mct_ctx* ctx, const int* rects, int n, float* out_host)
{
    if (!ctx || (n > 0 && (!rects || !out_host))) return MCT_E_ARG;
    if (!ctx->ii_base || ctx->frame_id == 0) return mct_fail(ctx, MCT_E_STATE, "no frame set%s");
    if (n <= 0) return MCT_OK;
    int rc = ensure_tmp(ctx, (size_t)n * 5 * sizeof(float));
    if (rc) return rc;
    int* d_r = (int*)ctx->d_tmp;
    float* d_o = ctx->d_tmp + (size_t)4 * n;
    MCT_CUDA(ctx, cudaMemcpyAsync(d_r, rects, (size_t)n * 4 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_sum_rects(ctx, d_r, n, d_o))) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(out_host, d_o, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

// ------------------------------------------------------------------------------------------ classifier
static int clf_ensure_test(mct_clf* clf, int n, int rows)
{
    mct_ctx* ctx = clf->ctx;
    if (n > clf->test_cap) {
        MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int cap = round_up(n, 256);
        if (clf->d_col) { cudaFree(clf->d_col); cudaFree(clf->d_row); cudaFree(clf->d_sx); cudaFree(clf->d_sy); cudaFree(clf->d_H); }
        MCT_CUDA(ctx, cudaMalloc(&clf->d_col, (size_t)cap * 4)); MCT_CUDA(ctx, cudaMalloc(&clf->d_row, (size_t)cap * 4));
        MCT_CUDA(ctx, cudaMalloc(&clf->d_sx, (size_t)cap * 4)); MCT_CUDA(ctx, cudaMalloc(&clf->d_sy, (size_t)cap * 4));
        MCT_CUDA(ctx, cudaMalloc(&clf->d_H, (size_t)cap * 4));
        clf->test_cap = cap;
    }
    int ldN = round_up(n > 0 ? n : 1, 32);
    size_t need = (size_t)(rows > 0 ? rows : 1) * ldN;
    if (need > clf->featN_cap) {
        MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (clf->d_featN) cudaFree(clf->d_featN);
        MCT_CUDA(ctx, cudaMalloc(&clf->d_featN, need * sizeof(float)));
        clf->featN_cap = need;
    }
    clf->ldN = ldN;
    if (clf->p.feature_type == MCT_FEAT_MDCH || clf->p.feature_type == MCT_FEAT_HAAR_MDCH) {
        if (n > clf->big_cap) {
            MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (clf->d_big) cudaFree(clf->d_big);
            int cap = round_up(n, 256);
            MCT_CUDA(ctx, cudaMalloc(&clf->d_big, (size_t)cap * 4));
            clf->big_cap = cap;
        }
    }
    return MCT_OK;
}

static int upload_boxes(mct_ctx* ctx, const mct_boxes* b, int* d_col, int* d_row, float* d_sx, float* d_sy, int off)
{
    if (b->n <= 0) return MCT_OK;
    if (!b->col || !b->row || !b->sx || !b->sy) return mct_fail(ctx, MCT_E_ARG, "box arrays are NULL%s");
    size_t bytes = (size_t)b->n * 4;
    MCT_CUDA(ctx, cudaMemcpyAsync(d_col + off, b->col, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(d_row + off, b->row, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(d_sx + off, b->sx, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(d_sy + off, b->sy, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return MCT_OK;
}

extern "C" int mct_classifier_create(mct_ctx* ctx, const mct_clf_params* p, const mct_haar_table* haar, mct_clf** out)
{
    if (!ctx || !p || !out) return MCT_E_ARG;
    *out = nullptr;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int ncol = p->n_bins * p->n_bins * p->n_bins;
    int n_haar = 0, n_color = 0;
    switch (p->feature_type) {
        case MCT_FEAT_HAAR: n_haar = p->n_haar; break;
        case MCT_FEAT_CCH: n_color = 11; break;
        case MCT_FEAT_MDCH: n_color = ncol; break;
        case MCT_FEAT_HAAR_MDCH: n_haar = p->n_haar; n_color = ncol; break;
        default: return mct_fail(ctx, MCT_E_ARG, "unknown feature type%s (%d)", "", p->feature_type);
    }
    if ((p->feature_type == MCT_FEAT_MDCH || p->feature_type == MCT_FEAT_HAAR_MDCH) &&
        (p->n_bins < 1 || p->n_bins > 8 || 256 % p->n_bins != 0))
        return mct_fail(ctx, MCT_E_UNSUPPORTED, "Color_Number_Of_Bins must divide 256 and be <= 8%s (%d)", "", p->n_bins);
    if (p->weak_type < 0 || p->weak_type > 2) return mct_fail(ctx, MCT_E_ARG, "unknown weak classifier type%s (%d)", "", p->weak_type);
    if (p->strong_type != MCT_STRONG_MILBOOST && p->strong_type != MCT_STRONG_MILANYBOOST && p->strong_type != MCT_STRONG_ADABOOST &&
        p->strong_type != MCT_STRONG_MILENSEMBLE)
        return mct_fail(ctx, MCT_E_UNSUPPORTED, "unknown strong classifier type%s (%d)", "", p->strong_type);
    if (p->strong_type == MCT_STRONG_MILENSEMBLE && !(p->pct_retained >= 0.0f && p->pct_retained <= 100.0f))
        return mct_fail(ctx, MCT_E_ARG, "Percentage_Of_Weak_Classifiers_Retained outside [0, 100]%s");
    if (n_haar > 0 && (!haar || haar->n_features != n_haar || !haar->nrect || !haar->rects || !haar->weights))
        return mct_fail(ctx, MCT_E_ARG, "Haar table missing or of the wrong size%s");
    if (p->w0 < 4 || p->h0 < 4) return mct_fail(ctx, MCT_E_ARG, "object blob too small%s");
    int F = n_haar + n_color;
    if (F < 1) return mct_fail(ctx, MCT_E_ARG, "classifier has no features%s");
    mct_clf* c = new (std::nothrow) mct_clf();
    if (!c) return mct_fail(ctx, MCT_E_NOMEM, "out of host memory%s");
    memset(c, 0, sizeof(*c));
    c->ctx = ctx; c->p = *p; c->F = F; c->n_haar = n_haar; c->n_color = n_color;
    int K = p->n_selected;
    if (K > F || K < 1) K = F / 2;                       // StrongClassifierBase.cpp:20-23
    if (K < 1) K = 1;
    c->K = K;
    c->max_train = p->max_train > 0 ? p->max_train : 4096;
    c->ldT = round_up(c->max_train, 32);
    if (n_haar > 0) {
        std::vector<int4> r((size_t)n_haar * MCT_MAX_RECTS);
        for (int f = 0; f < n_haar; f++) {
            if (haar->nrect[f] < 1 || haar->nrect[f] > MCT_MAX_RECTS) { delete c; return mct_fail(ctx, MCT_E_ARG, "Haar feature with a bad rect count%s (%d)", "", f); }
            for (int k = 0; k < MCT_MAX_RECTS; k++) {
                const int* s = haar->rects + ((size_t)f * MCT_MAX_RECTS + k) * 4;
                r[(size_t)f * MCT_MAX_RECTS + k] = make_int4(s[0], s[1], s[2], s[3]);
                if (k < haar->nrect[f]) {
                    if (s[0] + s[2] > c->haar_mxw) c->haar_mxw = s[0] + s[2];
                    if (s[1] + s[3] > c->haar_mxh) c->haar_mxh = s[1] + s[3];
                }
            }
        }
        MCT_CUDA(ctx, cudaMalloc(&c->d_rects, r.size() * sizeof(int4)));
        MCT_CUDA(ctx, cudaMalloc(&c->d_wts, (size_t)n_haar * MCT_MAX_RECTS * 4));
        MCT_CUDA(ctx, cudaMalloc(&c->d_nrect, (size_t)n_haar * 4));
        MCT_CUDA(ctx, cudaMemcpy(c->d_rects, r.data(), r.size() * sizeof(int4), cudaMemcpyHostToDevice));
        MCT_CUDA(ctx, cudaMemcpy(c->d_wts, haar->weights, (size_t)n_haar * MCT_MAX_RECTS * 4, cudaMemcpyHostToDevice));
        MCT_CUDA(ctx, cudaMemcpy(c->d_nrect, haar->nrect, (size_t)n_haar * 4, cudaMemcpyHostToDevice));
    }
    // weak state: mu = 0, sig = 1 (OnlineStumpsWeakClassifier.cpp:53-60); derived rows filled by the first Update
    std::vector<float> w((size_t)8 * F, 0.0f);
    for (int f = 0; f < F; f++) { w[(size_t)2 * F + f] = 1.0f; w[(size_t)3 * F + f] = 1.0f; }
    MCT_CUDA(ctx, cudaMalloc(&c->d_weak, w.size() * 4));
    MCT_CUDA(ctx, cudaMemcpy(c->d_weak, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    MCT_CUDA(ctx, cudaMalloc(&c->d_trained, (size_t)F * 4));
    MCT_CUDA(ctx, cudaMemset(c->d_trained, 0, (size_t)F * 4));
    MCT_CUDA(ctx, cudaMalloc(&c->d_sel, (size_t)F * 4));
    MCT_CUDA(ctx, cudaMemset(c->d_sel, 0, (size_t)F * 4));
    MCT_CUDA(ctx, cudaMalloc(&c->d_nsel, 4));
    MCT_CUDA(ctx, cudaMemset(c->d_nsel, 0, 4));
    if (p->strong_type == MCT_STRONG_MILENSEMBLE) {
        // m_selectorList.resize(K, 0) (StrongClassifierBase.cpp:91): K copies of feature 0.  MILEnsemble alone does not clear the list
        // before its first retain step, so untrained weak classifier 0 is retained by the first Update and never trained while it stays
        // (pinned against the reference build: tests/test_ref_pins_oracle.py::test_strong_classifier_update_and_classify[*ensemble])
        MCT_CUDA(ctx, cudaMemcpy(c->d_nsel, &K, 4, cudaMemcpyHostToDevice));
    }
    MCT_CUDA(ctx, cudaMalloc(&c->d_seltab, (size_t)K * sizeof(SelEntry)));
    MCT_CUDA(ctx, cudaMemset(c->d_seltab, 0, (size_t)K * sizeof(SelEntry)));
    MCT_CUDA(ctx, cudaMalloc(&c->d_selhdr, 16));
    MCT_CUDA(ctx, cudaMemset(c->d_selhdr, 0, 16));
    if (p->strong_type == MCT_STRONG_ADABOOST) {
        // running TP / FP / TN / FN weights, initial value 1 (AdaBoostClassifier.cpp:27-36), and the voting weights
        const size_t nc = (size_t)4 * K * c->F;
        std::vector<float> ones(nc, 1.0f);
        MCT_CUDA(ctx, cudaMalloc(&c->d_ada_cnt, nc * 4));
        MCT_CUDA(ctx, cudaMemcpy(c->d_ada_cnt, ones.data(), nc * 4, cudaMemcpyHostToDevice));
        MCT_CUDA(ctx, cudaMalloc(&c->d_alpha, (size_t)K * 4));
        MCT_CUDA(ctx, cudaMemset(c->d_alpha, 0, (size_t)K * 4));
    }
    if (p->strong_type == MCT_STRONG_MILENSEMBLE) {
        // retained flags; hypothesis left by the retain step followed by the round's nll / sorted rows
        MCT_CUDA(ctx, cudaMalloc(&c->d_keep, (size_t)F * 4));
        MCT_CUDA(ctx, cudaMemset(c->d_keep, 0, (size_t)F * 4));
        MCT_CUDA(ctx, cudaMalloc(&c->d_ensH, ((size_t)c->ldT + 2 * (size_t)F) * 4));
        MCT_CUDA(ctx, cudaMemset(c->d_ensH, 0, ((size_t)c->ldT + 2 * (size_t)F) * 4));
    }
    MCT_CUDA(ctx, cudaMalloc(&c->d_tcol, (size_t)c->ldT * 4)); MCT_CUDA(ctx, cudaMalloc(&c->d_trow, (size_t)c->ldT * 4));
    MCT_CUDA(ctx, cudaMalloc(&c->d_tsx, (size_t)c->ldT * 4)); MCT_CUDA(ctx, cudaMalloc(&c->d_tsy, (size_t)c->ldT * 4));
    MCT_CUDA(ctx, cudaMalloc(&c->d_tw, (size_t)c->ldT * 4));
    MCT_CUDA(ctx, cudaMalloc(&c->d_featT, (size_t)F * c->ldT * 4));
    MCT_CUDA(ctx, cudaMalloc(&c->d_pred, (size_t)F * c->ldT * 4));
    if (n_color > 0) {
        std::vector<int> id(n_color);
        for (int i = 0; i < n_color; i++) id[i] = i;
        MCT_CUDA(ctx, cudaMalloc(&c->d_colorrow, (size_t)n_color * 4));
        MCT_CUDA(ctx, cudaMemcpy(c->d_colorrow, id.data(), (size_t)n_color * 4, cudaMemcpyHostToDevice));
        MCT_CUDA(ctx, cudaMalloc(&c->d_outbins, (size_t)n_color * 4));
        MCT_CUDA(ctx, cudaMemcpy(c->d_outbins, id.data(), (size_t)n_color * 4, cudaMemcpyHostToDevice));
        MCT_CUDA(ctx, cudaMalloc(&c->d_slotmap, (size_t)n_color * 2));
        MCT_CUDA(ctx, cudaMemset(c->d_slotmap, 0, (size_t)n_color * 2));
        MCT_CUDA(ctx, cudaMalloc(&c->d_nout, 4)); MCT_CUDA(ctx, cudaMemset(c->d_nout, 0, 4));
        MCT_CUDA(ctx, cudaMalloc(&c->d_rowst, (size_t)n_color * 8 * 4)); MCT_CUDA(ctx, cudaMemset(c->d_rowst, 0, (size_t)n_color * 8 * 4));
        MCT_CUDA(ctx, cudaMalloc(&c->d_nbig, 8)); MCT_CUDA(ctx, cudaMemset(c->d_nbig, 0, 8));     // [0] flagged boxes, [1] k_mdch_sel arrival counter
    }
    *out = c;
    return MCT_OK;
}

extern "C" int mct_classifier_destroy(mct_clf* c)
{
    if (!c) return MCT_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    void* ptrs[] = {c->d_rects, c->d_wts, c->d_nrect, c->d_weak, c->d_trained, c->d_sel, c->d_nsel, c->d_seltab, c->d_selhdr, c->d_tcol, c->d_trow,
                    c->d_tsx, c->d_tsy, c->d_tw, c->d_featT, c->d_pred, c->d_col, c->d_row, c->d_sx, c->d_sy, c->d_featN,
                    c->d_H, c->d_outbins, c->d_colorrow, c->d_slotmap, c->d_nout, c->d_big, c->d_nbig, c->d_tsort, c->d_ada_cnt, c->d_alpha,
                    c->d_keep, c->d_ensH, c->d_rowst};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete c;
    return MCT_OK;
}

extern "C" int mct_classifier_num_features(mct_clf* c) { return c ? c->F : MCT_E_ARG; }
extern "C" int mct_classifier_set_cluster(mct_clf* c, int cluster)
{
    if (!c || cluster < 0 || cluster > 16 || (cluster & (cluster - 1))) return MCT_E_ARG;
    c->cluster = cluster;
    return MCT_OK;
}

static int need_frame(mct_clf* c)
{
    mct_ctx* ctx = c->ctx;
    if (ctx->frame_id == 0) return mct_fail(ctx, MCT_E_STATE, "no frame set%s");
    if (c->n_color > 0 && c->p.feature_type != MCT_FEAT_CCH && (!ctx->c0 || !ctx->c1 || !ctx->c2))
        return mct_fail(ctx, MCT_E_STATE, "colour planes not set for a colour-histogram classifier%s");
    if (c->p.feature_type == MCT_FEAT_CCH && !ctx->c0) return mct_fail(ctx, MCT_E_STATE, "hue plane not set%s");
    return MCT_OK;
}

extern "C" int mct_features_compute(mct_clf* c, const mct_boxes* boxes, float* out_host)
{
    if (!c || !boxes || (boxes->n > 0 && !out_host)) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    int rc = need_frame(c);
    if (rc) return rc;
    const int n = boxes->n;
    if (n <= 0) return MCT_OK;
    if ((rc = clf_ensure_test(c, n, c->F))) return rc;
    if ((rc = upload_boxes(ctx, boxes, c->d_col, c->d_row, c->d_sx, c->d_sy, 0))) return rc;
    if ((rc = mct_feature_matrix_full(c, c->d_col, c->d_row, c->d_sx, c->d_sy, n, c->d_featN, c->ldN))) return rc;
    MCT_CUDA(ctx, cudaMemcpy2DAsync(out_host, (size_t)n * 4, c->d_featN, (size_t)c->ldN * 4, (size_t)n * 4, c->F,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

static int clf_train_common(mct_clf* c, int P, int Nn, const float* wpos, const float* wneg)
{
    mct_ctx* ctx = c->ctx;
    int has_w = (wpos && wneg) ? 1 : 0;
    if (has_w) {
        MCT_CUDA(ctx, cudaMemcpyAsync(c->d_tw, wpos, (size_t)P * 4, cudaMemcpyHostToDevice, ctx->stream));
        MCT_CUDA(ctx, cudaMemcpyAsync(c->d_tw + P, wneg, (size_t)Nn * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    int rc;
    if ((rc = launch_ensemble_retain(c, P, Nn))) return rc;
    if ((rc = launch_weak_update(c, P, Nn, has_w))) return rc;
    if ((rc = launch_mil_select(c, P, Nn))) return rc;
    if ((rc = launch_build_colormap(c))) return rc;
    c->lastP = P; c->lastNn = Nn;
    return MCT_OK;
}

extern "C" int mct_classifier_update(mct_clf* c, const mct_boxes* pos, const mct_boxes* neg, const float* wpos, const float* wneg)
{
    if (!c || !pos || !neg) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    int rc = need_frame(c);
    if (rc) return rc;
    const int P = pos->n, Nn = neg->n;
    if (P < 1 || Nn < 1) return mct_fail(ctx, MCT_E_ARG, "Update needs at least one positive and one negative sample%s");
    if (P + Nn > c->max_train) return mct_fail(ctx, MCT_E_ARG, "training set exceeds max_train%s (%d)", "", P + Nn);
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((rc = upload_boxes(ctx, pos, c->d_tcol, c->d_trow, c->d_tsx, c->d_tsy, 0))) return rc;
    if ((rc = upload_boxes(ctx, neg, c->d_tcol, c->d_trow, c->d_tsx, c->d_tsy, P))) return rc;
    if ((rc = mct_feature_matrix_full(c, c->d_tcol, c->d_trow, c->d_tsx, c->d_tsy, P + Nn, c->d_featT, c->ldT))) return rc;
    return clf_train_common(c, P, Nn, wpos, wneg);
}

// device-resident variants for the appearance fuser's exchange (rows stay in HBM between the feature kernels, the NCCL
// all-gather and the update): dev_out / dev_fpos / dev_fneg are device pointers with the given leading dimensions
extern "C" int mct_features_compute_dev(mct_clf* c, const mct_boxes* boxes, float* dev_out, int ld_out)
{
    if (!c || !boxes || (boxes->n > 0 && (!dev_out || ld_out < boxes->n))) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    int rc = need_frame(c);
    if (rc) return rc;
    const int n = boxes->n;
    if (n <= 0) return MCT_OK;
    if ((rc = clf_ensure_test(c, n, c->F))) return rc;
    if ((rc = upload_boxes(ctx, boxes, c->d_col, c->d_row, c->d_sx, c->d_sy, 0))) return rc;
    if ((rc = mct_feature_matrix_full(c, c->d_col, c->d_row, c->d_sx, c->d_sy, n, c->d_featN, c->ldN))) return rc;
    MCT_CUDA(ctx, cudaMemcpy2DAsync(dev_out, (size_t)ld_out * 4, c->d_featN, (size_t)c->ldN * 4, (size_t)n * 4, c->F,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    return mct_sync(ctx);          // the host box arrays may be reused by the caller
}
extern "C" int mct_classifier_update_from_features_dev(mct_clf* c, const float* dev_fpos, int ld_pos, int P, const float* dev_fneg,
                                                       int ld_neg, int Nn)
{
    if (!c || !dev_fpos || !dev_fneg || ld_pos < P || ld_neg < Nn) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    if (P < 1 || Nn < 1) return mct_fail(ctx, MCT_E_ARG, "Update needs at least one positive and one negative sample%s");
    if (P + Nn > c->max_train) return mct_fail(ctx, MCT_E_ARG, "training set exceeds max_train%s (%d)", "", P + Nn);
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaMemcpy2DAsync(c->d_featT, (size_t)c->ldT * 4, dev_fpos, (size_t)ld_pos * 4, (size_t)P * 4, c->F,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpy2DAsync(c->d_featT + P, (size_t)c->ldT * 4, dev_fneg, (size_t)ld_neg * 4, (size_t)Nn * 4, c->F,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    // the caller's buffers have been consumed when this call returns (a caching allocator may hand them out again at once)
    MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return clf_train_common(c, P, Nn, nullptr, nullptr);
}
// Appearance fuser, learner side (AppearanceBasedInformationFuser.cpp:44-90): the pooled sample sets of all cameras arrive as
// feature rows in exchange buffers; column j of this Update's training matrix is the feature column that starts at element
// col_offset[j] of dev_rows (rows row_stride elements apart).  The kept samples (randfloat() > 0.5, :66-82) are gathered
// straight into the classifier's training matrix: no intermediate pooled copy.
__global__ void __launch_bounds__(256) k_gather_columns(const float* __restrict__ src, const int* __restrict__ off, long long row_stride,
                                                        int n, int F, float* __restrict__ dst, int ld)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float* s = src + off[j];
    for (int f = blockIdx.y; f < F; f += gridDim.y) dst[(size_t)f * ld + j] = s[(size_t)f * row_stride];
}
// all classifiers of a camera in one launch (blockIdx.z = object; the object's column offsets hang off TrainDesc::col)
__global__ void __launch_bounds__(256) k_gather_columns_batch(const TrainDesc* __restrict__ descs, const float* __restrict__ src, long long row_stride)
{
    const TrainDesc& d = descs[blockIdx.z];
    const int n = d.P + d.Nn;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float* s = src + d.col[j];
    for (int f = blockIdx.y; f < d.F; f += gridDim.y) d.featT[(size_t)f * d.ldT + j] = s[(size_t)f * row_stride];
}
extern "C" int mct_classifier_update_from_gathered(mct_clf* c, const float* dev_rows, const int* col_offset_host, long long row_stride, int P, int Nn)
{
    if (!c || !dev_rows || !col_offset_host || row_stride < 1) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    if (P < 1 || Nn < 1) return mct_fail(ctx, MCT_E_ARG, "Update needs at least one positive and one negative sample%s");
    if (P + Nn > c->max_train) return mct_fail(ctx, MCT_E_ARG, "training set exceeds max_train%s (%d)", "", P + Nn);
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaMemcpyAsync(c->d_tcol, col_offset_host, (size_t)(P + Nn) * 4, cudaMemcpyHostToDevice, ctx->stream));
    const int n = P + Nn;
    MCT_LAUNCH(ctx, "k_gather_columns", k_gather_columns<<<dim3(ceil_div(n, 256), c->F < 64 ? c->F : 64), 256, 0, ctx->stream>>>(
                   dev_rows, c->d_tcol, row_stride, n, c->F, c->d_featT, c->ldT));
    return clf_train_common(c, P, Nn, nullptr, nullptr);
}
// plain device buffers on the ctx's device (exchange buffers of the appearance fuser) and stream-ordered copies between them
extern "C" int mct_dev_alloc(mct_ctx* ctx, size_t bytes, void** out)
{
    if (!ctx || !out) return MCT_E_ARG;
    *out = nullptr;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return MCT_OK;
}
extern "C" int mct_dev_free(mct_ctx* ctx, void* p)
{
    if (!ctx) return MCT_E_ARG;
    if (!p) return MCT_OK;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MCT_CUDA(ctx, cudaFree(p));
    return MCT_OK;
}
extern "C" int mct_dev_copy(mct_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx || (bytes && (!dst || !src))) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    return MCT_OK;
}
extern "C" int mct_classifier_update_from_features(mct_clf* c, const float* fpos, int P, const float* fneg, int Nn,
                                                   const float* wpos, const float* wneg)
{
    if (!c || !fpos || !fneg) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    if (P < 1 || Nn < 1) return mct_fail(ctx, MCT_E_ARG, "Update needs at least one positive and one negative sample%s");
    if (P + Nn > c->max_train) return mct_fail(ctx, MCT_E_ARG, "training set exceeds max_train%s (%d)", "", P + Nn);
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaMemcpy2DAsync(c->d_featT, (size_t)c->ldT * 4, fpos, (size_t)P * 4, (size_t)P * 4, c->F,
                                    cudaMemcpyHostToDevice, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpy2DAsync(c->d_featT + P, (size_t)c->ldT * 4, fneg, (size_t)Nn * 4, (size_t)Nn * 4, c->F,
                                    cudaMemcpyHostToDevice, ctx->stream));
    return clf_train_common(c, P, Nn, wpos, wneg);
}

static void fill_clf_desc(ObjDesc* d, mct_clf* c)
{
    d->rects = c->d_rects; d->wts = c->d_wts; d->nrect = c->d_nrect;
    d->weak = c->d_weak; d->sel = c->d_sel; d->nsel = c->d_nsel; d->seltab = c->d_seltab; d->selhdr = c->d_selhdr;
    d->alpha = c->p.strong_type == MCT_STRONG_ADABOOST ? c->d_alpha : nullptr;
    d->featN = c->d_featN; d->ldN = c->ldN; d->n_color_rows = c->n_color;
    d->colorrow = c->d_colorrow; d->outbins = c->d_outbins;
    const bool mdch = c->p.feature_type == MCT_FEAT_MDCH || c->p.feature_type == MCT_FEAT_HAAR_MDCH;
    d->slotmap = mdch ? c->d_slotmap : nullptr; d->nout = c->d_nout; d->big = c->d_big; d->nbig = c->d_nbig;
    d->rowst = c->d_rowst; d->color_resp = 0;
    d->cw0 = c->p.w0; d->ch0 = c->p.h0;
    d->F = c->F; d->n_haar = c->n_haar; d->nb = c->p.n_bins; d->weak_type = c->p.weak_type;
    d->feature_type = c->p.feature_type; d->cch_mode = c->p.cch_mode; d->K = c->K;
}

static int classify_boxes_on_device(mct_clf* c, int n, int from_matrix, int log_ratio)
{
    mct_ctx* ctx = c->ctx;
    ObjDesc hd; ObjDesc* h = &hd; const ObjDesc* d; int rc;
    memset(h, 0, sizeof(ObjDesc));
    fill_clf_desc(h, c);
    h->N = n; h->ld = c->ldN;
    h->col = c->d_col; h->row = c->d_row; h->sx = c->d_sx; h->sy = c->d_sy; h->valid = nullptr; h->H = c->d_H;
    if ((rc = desc_publish(ctx, h, 1, &d))) return rc;
    if (!from_matrix && c->n_color > 0) {
        if (h->slotmap) {
            if ((rc = launch_bin_image(ctx, c->p.n_bins))) return rc;
            MCT_CUDA(ctx, cudaMemsetAsync(c->d_nbig, 0, 4, ctx->stream));
            StepArgs sa0; memset(&sa0, 0, sizeof(sa0));
            if ((rc = launch_mdch_selected(ctx, d, 1, n, (c->K < c->n_color ? c->K : c->n_color) + 1, sa0, 0))) return rc;
        } else if ((rc = mct_feature_color_rows(c, c->d_col, c->d_row, c->d_sx, c->d_sy, nullptr, n, nullptr, c->n_color,
                                                c->d_featN, c->ldN)))
            return rc;
    }
    return launch_classify(ctx, d, 1, n, c->K, from_matrix, log_ratio);
}

extern "C" int mct_classifier_classify(mct_clf* c, const mct_boxes* boxes, int log_ratio, float* out_host)
{
    if (!c || !boxes || (boxes->n > 0 && !out_host)) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    int rc = need_frame(c);
    if (rc) return rc;
    const int n = boxes->n;
    if (n <= 0) return MCT_OK;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((rc = clf_ensure_test(c, n, c->n_color))) return rc;
    if ((rc = upload_boxes(ctx, boxes, c->d_col, c->d_row, c->d_sx, c->d_sy, 0))) return rc;
    if ((rc = classify_boxes_on_device(c, n, 0, log_ratio))) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(out_host, c->d_H, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

// SimpleTracker::TrackObjectOnTheGivenFrame's device part for all objects of a camera (SimpleTracker.cpp:280-342): the
// strong classifiers score their search windows in one batched pass and only {best index, best response} per object comes
// back.  boxes[o]: the test sample set of object o (host).  NaN responses are never picked unless all are NaN.
extern "C" int mct_classify_argmax(mct_ctx* ctx, int n_obj, mct_clf* const* clfs, const mct_boxes* boxes, int log_ratio,
                                   int* best_idx, float* best_resp)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !clfs || !boxes || !best_idx || !best_resp) return MCT_E_ARG;
    if (ctx->frame_id == 0) return mct_fail(ctx, MCT_E_STATE, "no frame set%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    ObjDesc h[DESC_MAX_OBJ];
    int rc, n_max = 0, K_max = 1, any_mdch = 0, max_slots = 1, nb = 0;
    for (int o = 0; o < n_obj; o++) {
        mct_clf* c = clfs[o];
        if (!c || c->ctx != ctx) return mct_fail(ctx, MCT_E_ARG, "classifier handle belongs to another ctx%s");
        if (boxes[o].n < 1) return mct_fail(ctx, MCT_E_EMPTY, "object %s%d has an empty search window", "", o);
        if ((rc = need_frame(c))) return rc;
        if ((rc = clf_ensure_test(c, boxes[o].n, c->n_color))) return rc;
        if ((rc = upload_boxes(ctx, &boxes[o], c->d_col, c->d_row, c->d_sx, c->d_sy, 0))) return rc;
        memset(&h[o], 0, sizeof(ObjDesc));
        fill_clf_desc(&h[o], c);
        h[o].N = boxes[o].n; h[o].ld = c->ldN;
        h[o].col = c->d_col; h[o].row = c->d_row; h[o].sx = c->d_sx; h[o].sy = c->d_sy; h[o].valid = nullptr; h[o].H = c->d_H;
        if (boxes[o].n > n_max) n_max = boxes[o].n;
        if (c->K > K_max) K_max = c->K;
        if (c->n_color > 0 && h[o].slotmap) {
            any_mdch = 1; nb = c->p.n_bins;
            int ms = (c->K < c->n_color ? c->K : c->n_color) + 1;
            if (ms > max_slots) max_slots = ms;
        }
    }
    const ObjDesc* d;
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    for (int o = 0; o < n_obj; o++) {          // CCH rows per object (the reference's degenerate histogram; rarely configured)
        mct_clf* c = clfs[o];
        if (c->n_color > 0 && !h[o].slotmap &&
            (rc = mct_feature_color_rows(c, c->d_col, c->d_row, c->d_sx, c->d_sy, nullptr, boxes[o].n, nullptr, c->n_color, c->d_featN, c->ldN)))
            return rc;
    }
    if (any_mdch) {
        StepArgs sa0; memset(&sa0, 0, sizeof(sa0));
        if ((rc = launch_bin_image(ctx, nb))) return rc;
        if ((rc = launch_mdch_selected(ctx, d, n_obj, n_max, max_slots, sa0, 0))) return rc;
    }
    if ((rc = launch_classify(ctx, d, n_obj, n_max, K_max, 0, log_ratio))) return rc;
    if (ctx->argmax_cap < n_obj) {
        if (ctx->d_argmax) cudaFree(ctx->d_argmax);
    if (ctx->d_mlist) cudaFree(ctx->d_mlist);
    if (ctx->d_macc) cudaFree(ctx->d_macc);
    if (ctx->d_mtot) cudaFree(ctx->d_mtot);
        MCT_CUDA(ctx, cudaMalloc(&ctx->d_argmax, (size_t)DESC_MAX_OBJ * 8));
        ctx->argmax_cap = DESC_MAX_OBJ;
    }
    if ((rc = launch_response_argmax(ctx, d, n_obj, ctx->d_argmax))) return rc;
    struct { int idx; float val; } out[DESC_MAX_OBJ];
    MCT_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_argmax, (size_t)n_obj * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = mct_sync(ctx))) return rc;
    for (int o = 0; o < n_obj; o++) { best_idx[o] = out[o].idx; best_resp[o] = out[o].val; }
    return MCT_OK;
}

extern "C" int mct_classifier_classify_from_features(mct_clf* c, const float* feat, int n, int log_ratio, float* out_host)
{
    if (!c || !feat || n < 0 || (n > 0 && !out_host)) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    if (n == 0) return MCT_OK;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = clf_ensure_test(c, n, c->F))) return rc;
    MCT_CUDA(ctx, cudaMemcpy2DAsync(c->d_featN, (size_t)c->ldN * 4, feat, (size_t)n * 4, (size_t)n * 4, c->F,
                                    cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = classify_boxes_on_device(c, n, 1, log_ratio))) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(out_host, c->d_H, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

extern "C" int mct_classifier_get_selectors(mct_clf* c, int* idx_host, int* n)
{
    if (!c || !idx_host || !n) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    MCT_CUDA(ctx, cudaMemcpyAsync(n, c->d_nsel, 4, cudaMemcpyDeviceToHost, ctx->stream));
    int rc = mct_sync(ctx);
    if (rc) return rc;
    if (*n > 0) {
        MCT_CUDA(ctx, cudaMemcpyAsync(idx_host, c->d_sel, (size_t)(*n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        return mct_sync(ctx);
    }
    return MCT_OK;
}
extern "C" int mct_classifier_get_alphas(mct_clf* c, float* alphas_host, int* n)
{
    if (!c || !alphas_host || !n) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    *n = 0;
    if (c->p.strong_type != MCT_STRONG_ADABOOST) return MCT_OK;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int ns = 0;
    MCT_CUDA(ctx, cudaMemcpyAsync(&ns, c->d_nsel, 4, cudaMemcpyDeviceToHost, ctx->stream));
    int rc = mct_sync(ctx);
    if (rc) return rc;
    if (ns > c->K) ns = c->K;
    MCT_CUDA(ctx, cudaMemcpyAsync(alphas_host, c->d_alpha, (size_t)ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
    *n = ns;
    return mct_sync(ctx);
}
extern "C" int mct_classifier_get_weak_state(mct_clf* c, float* out)
{
    if (!c || !out) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    MCT_CUDA(ctx, cudaMemcpyAsync(out, c->d_weak, (size_t)8 * c->F * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}
extern "C" int mct_classifier_get_predictions(mct_clf* c, float* out, int* P, int* Nn)
{
    if (!c || !out || !P || !Nn) return MCT_E_ARG;
    mct_ctx* ctx = c->ctx;
    if (c->lastP < 0 && ctx->gen_pending) {              // the last Update took its set sizes on the device: fetch them
        int rc = mct_track_update_default_collect(ctx, ctx->gen_n, nullptr);
        if (rc && rc != MCT_E_EMPTY) return rc;
    }
    *P = c->lastP; *Nn = c->lastNn;
    int M = c->lastP + c->lastNn;
    if (M <= 0) return MCT_OK;
    MCT_CUDA(ctx, cudaMemcpy2DAsync(out, (size_t)M * 4, c->d_pred, (size_t)c->ldT * 4, (size_t)M * 4, c->F,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

// ------------------------------------------------------------------------------------------ particle filter
extern "C" int mct_pf_create(mct_ctx* ctx, int N, float image_w, float image_h, mct_pf** out)
{
    if (!ctx || !out || N < 1) return MCT_E_ARG;
    *out = nullptr;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    mct_pf* p = new (std::nothrow) mct_pf();
    if (!p) return mct_fail(ctx, MCT_E_NOMEM, "out of host memory%s");
    memset(p, 0, sizeof(*p));
    p->ctx = ctx; p->N = N; p->ld = round_up(N, 32); p->img_w = image_w; p->img_h = image_h;
    size_t ld = p->ld;
    MCT_CUDA(ctx, cudaMalloc(&p->d_state, 5 * ld * 4)); MCT_CUDA(ctx, cudaMalloc(&p->d_res, 5 * ld * 4));
    MCT_CUDA(ctx, cudaMalloc(&p->d_residx, ld * 4));
    MCT_CUDA(ctx, cudaMalloc(&p->d_col, ld * 4)); MCT_CUDA(ctx, cudaMalloc(&p->d_row, ld * 4));
    MCT_CUDA(ctx, cudaMalloc(&p->d_valid, ld * 4)); MCT_CUDA(ctx, cudaMalloc(&p->d_H, ld * 4));
    MCT_CUDA(ctx, cudaMalloc(&p->d_noise, 4 * ld * 4)); MCT_CUDA(ctx, cudaMalloc(&p->d_prob, ld * 4));
    MCT_CUDA(ctx, cudaMalloc(&p->d_scratch, 2 * (ld + ld / 64 + 64) * 4));   // sw | cdf in the padded chain layout (k_pf.cu PLEN)
    MCT_CUDA(ctx, cudaMalloc(&p->d_st, sizeof(PfDev))); MCT_CUDA(ctx, cudaMemset(p->d_st, 0, sizeof(PfDev)));
    MCT_CUDA(ctx, cudaMalloc(&p->d_trk, sizeof(PfTrk))); MCT_CUDA(ctx, cudaMemset(p->d_trk, 0, sizeof(PfTrk)));
    MCT_CUDA(ctx, cudaMalloc(&p->d_summ, sizeof(mct_pf_summary)));
    MCT_CUDA(ctx, cudaMemset(p->d_summ, 0, sizeof(mct_pf_summary)));
    MCT_CUDA(ctx, cudaMemset(p->d_state, 0, 5 * ld * 4)); MCT_CUDA(ctx, cudaMemset(p->d_res, 0, 5 * ld * 4));
    MCT_CUDA(ctx, cudaMemset(p->d_residx, 0, ld * 4)); MCT_CUDA(ctx, cudaMemset(p->d_valid, 0, ld * 4));
    MCT_CUDA(ctx, cudaMemset(p->d_H, 0, ld * 4));
    *out = p;
    return MCT_OK;
}
extern "C" int mct_pf_destroy(mct_pf* p)
{
    if (!p) return MCT_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    void* ptrs[] = {p->d_state, p->d_res, p->d_residx, p->d_col, p->d_row, p->d_valid, p->d_H, p->d_noise, p->d_prob,
                    p->d_scratch, p->d_st, p->d_summ, p->d_trk};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete p;
    return MCT_OK;
}

static void fill_pf_desc(ObjDesc* d, mct_pf* p)
{
    size_t ld = p->ld;
    d->cx = p->d_state; d->cy = p->d_state + ld; d->sx = p->d_state + 2 * ld; d->sy = p->d_state + 3 * ld; d->w = p->d_state + 4 * ld;
    d->rcx = p->d_res; d->rcy = p->d_res + ld; d->rsx = p->d_res + 2 * ld; d->rsy = p->d_res + 3 * ld; d->rw = p->d_res + 4 * ld;
    d->residx = p->d_residx; d->col = p->d_col; d->row = p->d_row; d->valid = p->d_valid;
    d->H = p->d_H; d->noise_off = 0; d->scratch = p->d_scratch; d->st = p->d_st; d->summ = p->d_summ;
    d->trk = p->d_trk;
    d->scored = p->ctx->d_scored;
    d->N = p->N; d->ld = p->ld;
    d->img_w = p->img_w; d->img_h = p->img_h; d->msc = p->msc; d->maxs = p->maxs; d->mins = p->mins;
    d->exponent = 20; d->force_reset = 0; d->resample = 1;
}

// host staging of an [N][5] row-major matrix <-> device SoA
static int pf_upload_matrix(mct_pf* p, float* d_dst, const float* src_Nx5)
{
    mct_ctx* ctx = p->ctx;
    std::vector<float> soa((size_t)5 * p->ld, 0.0f);
    for (int i = 0; i < p->N; i++)
        for (int j = 0; j < 5; j++) soa[(size_t)j * p->ld + i] = src_Nx5[(size_t)i * 5 + j];
    MCT_CUDA(ctx, cudaMemcpyAsync(d_dst, soa.data(), soa.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    return mct_sync(ctx);
}
static int pf_download_matrix(mct_pf* p, const float* d_src, float* dst_Nx5)
{
    mct_ctx* ctx = p->ctx;
    std::vector<float> soa((size_t)5 * p->ld);
    MCT_CUDA(ctx, cudaMemcpyAsync(soa.data(), d_src, soa.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    int rc = mct_sync(ctx);
    if (rc) return rc;
    for (int i = 0; i < p->N; i++)
        for (int j = 0; j < 5; j++) dst_Nx5[(size_t)i * 5 + j] = soa[(size_t)j * p->ld + i];
    return MCT_OK;
}

extern "C" int mct_pf_initialize(mct_pf* p, float cx, float cy, float sx, float sy, float msc, float maxs, float mins)
{
    if (!p) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    if (msc < 0 || !(maxs > 0 && maxs > sx && maxs > sy) || !(mins > 0 && mins < sx && mins < sy))
        return mct_fail(ctx, MCT_E_ARG, "ParticleFilter::Initialize: scale limits violated%s");
    p->msc = msc; p->maxs = maxs; p->mins = mins;
    std::vector<float> m((size_t)p->N * 5);
    for (int i = 0; i < p->N; i++) { m[i * 5 + 0] = cx; m[i * 5 + 1] = cy; m[i * 5 + 2] = sx; m[i * 5 + 3] = sy; m[i * 5 + 4] = 1.0f; }
    int rc;
    if ((rc = pf_upload_matrix(p, p->d_res, m.data()))) return rc;
    if ((rc = pf_upload_matrix(p, p->d_state, m.data()))) return rc;
    PfDev st; memset(&st, 0, sizeof(st)); st.filter_state = MCT_PF_INIT;
    MCT_CUDA(ctx, cudaMemcpyAsync(p->d_st, &st, sizeof(st), cudaMemcpyHostToDevice, ctx->stream));
    p->initialized = 1;
    return mct_sync(ctx);
}

extern "C" int mct_pf_get_state(mct_pf* p, float* out) { return (!p || !out) ? MCT_E_ARG : pf_download_matrix(p, p->d_state, out); }
extern "C" int mct_pf_get_resampled(mct_pf* p, float* out) { return (!p || !out) ? MCT_E_ARG : pf_download_matrix(p, p->d_res, out); }
extern "C" int mct_pf_set_state(mct_pf* p, const float* in, int filter_state)
{
    if (!p || !in) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    int rc = pf_upload_matrix(p, p->d_state, in);
    if (rc) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(&p->d_st->filter_state, &filter_state, 4, cudaMemcpyHostToDevice, ctx->stream));
    return mct_sync(ctx);
}
extern "C" int mct_pf_get_resample_indices(mct_pf* p, int* out)
{
    if (!p || !out) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    MCT_CUDA(ctx, cudaMemcpyAsync(out, p->d_residx, (size_t)p->N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

static int pf_single_desc(mct_pf* p, const ObjDesc** d_out, StepArgs* sa, float w0, float h0, int force_reset, int resample, float u01)
{
    mct_ctx* ctx = p->ctx;
    ObjDesc hd;
    memset(&hd, 0, sizeof(ObjDesc));
    fill_pf_desc(&hd, p);
    hd.w0 = w0; hd.h0 = h0; hd.force_reset = force_reset; hd.resample = resample;
    memset(sa, 0, sizeof(*sa));
    sa->noise_base = p->d_noise; sa->u01[0] = u01;
    return desc_publish(ctx, &hd, 1, d_out);
}

extern "C" int mct_pf_predict(mct_pf* p, const float* noise, int noise_is_device, float w0, float h0)
{
    if (!p || !noise) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    if (!p->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaMemcpyAsync(p->d_noise, noise, (size_t)4 * p->N * 4,
                                  noise_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    const ObjDesc* d; StepArgs sa; int rc;
    if ((rc = pf_single_desc(p, &d, &sa, w0, h0, 0, 1, 0.0f))) return rc;
    if ((rc = launch_pf_predict(ctx, d, 1, p->N, sa))) return rc;
    if (w0 > 0 && h0 > 0 && (rc = launch_pf_refine(ctx, d, 1))) return rc;      // CheckForParticleRefinement (:328)
    return mct_sync(ctx);     // the host noise buffer may be reused by the caller
}

// ---------------------------------------------------------------------------------------------------------------------
// Geometric-fuser feedback
static void ground_args_fill(GroundArgs* ga, const double* H9, const float* mean2, const float* cov4, float h0)
{
    memset(ga, 0, sizeof(*ga));
    for (int i = 0; i < 9; i++) ga->H[i] = H9[i];
    ga->mean[0] = mean2[0]; ga->mean[1] = mean2[1]; ga->h0 = h0;
    // cvDet / cvInvert of the 2x2 CV_32F covariance: f64 products, f32 store (GeometryBasedInformationFuser.cpp:119-133)
    const double det = (double)cov4[0] * (double)cov4[3] - (double)cov4[1] * (double)cov4[2];
    ga->usable = (det >= 0.0 && cov4[0] > 0) ? 1 : 0;       // ParticleFilterTracker.cpp:1063
    if (ga->usable) {
        if (det != 0.0) {
            const double r = 1.0 / det;
            ga->ci[0] = (float)((double)cov4[3] * r); ga->ci[1] = (float)(-(double)cov4[1] * r);
            ga->ci[2] = (float)(-(double)cov4[2] * r); ga->ci[3] = (float)((double)cov4[0] * r);
        }
        const float detf = (float)det;
        ga->b = 1.0f / sqrtf(detf);
        const double s2pi = sqrt(6.283185307179586);
        ga->a = 1.0 / (s2pi * s2pi);
    }
}
static int ground_buffers(mct_ctx* ctx)
{
    if (!ctx->d_ground_out) MCT_CUDA(ctx, cudaMalloc(&ctx->d_ground_out, sizeof(GroundOut) * DESC_MAX_OBJ));
    if (!ctx->d_ground_args) MCT_CUDA(ctx, cudaMalloc(&ctx->d_ground_args, sizeof(GroundArgs) * DESC_MAX_OBJ));
    return MCT_OK;
}
// all objects of a camera in one launch and one read-back (H9: the camera's homography; mean2 / cov4 / h0: per object)
extern "C" int mct_fuser_ground_pdf(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const double* H9, const float* mean2,
                                    const float* cov4, const float* h0, mct_ground_pdf_result* out)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !H9 || !mean2 || !cov4 || !h0 || !out) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    ObjDesc h[DESC_MAX_OBJ];
    GroundArgs ga[DESC_MAX_OBJ];
    for (int o = 0; o < n_obj; o++) {
        mct_pf* p = pfs[o];
        if (!p || p->ctx != ctx) return mct_fail(ctx, MCT_E_ARG, "particle filter belongs to another ctx%s");
        if (!p->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
        memset(&h[o], 0, sizeof(ObjDesc));
        fill_pf_desc(&h[o], p);
        h[o].h0 = h0[o];
        ground_args_fill(&ga[o], H9, mean2 + 2 * o, cov4 + 4 * o, h0[o]);
    }
    int rc;
    if ((rc = ground_buffers(ctx))) return rc;
    const ObjDesc* d;
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_ground_args, ga, sizeof(GroundArgs) * n_obj, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_ground_pdf(ctx, d, n_obj, ctx->d_ground_args, nullptr, ctx->d_ground_out))) return rc;
    GroundOut go[DESC_MAX_OBJ];
    MCT_CUDA(ctx, cudaMemcpyAsync(go, ctx->d_ground_out, sizeof(GroundOut) * n_obj, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = mct_sync(ctx))) return rc;
    for (int o = 0; o < n_obj; o++) {
        out[o].total_weight = go[o].total_weight;
        for (int q = 0; q < 4; q++) out[o].average[q] = go[o].average[q];
        out[o].usable_covariance = ga[o].usable;
    }
    return MCT_OK;
}

extern "C" int mct_pf_ground_pdf(mct_pf* p, const double* H9, const float* mean2, const float* cov4, float h0,
                                 mct_ground_pdf_result* out, float* weights_host)
{
    if (!p || !H9 || !mean2 || !cov4 || !out) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    if (!p->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    GroundArgs ga;
    ground_args_fill(&ga, H9, mean2, cov4, h0);
    const ObjDesc* d; StepArgs sa; int rc;
    if ((rc = pf_single_desc(p, &d, &sa, 0, h0, 0, 1, 0.0f))) return rc;
    if ((rc = ground_buffers(ctx))) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_ground_args, &ga, sizeof(GroundArgs), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_ground_pdf(ctx, d, 1, ctx->d_ground_args, weights_host ? p->d_prob : nullptr, ctx->d_ground_out))) return rc;
    GroundOut go;
    MCT_CUDA(ctx, cudaMemcpyAsync(&go, ctx->d_ground_out, sizeof(go), cudaMemcpyDeviceToHost, ctx->stream));
    if (weights_host) MCT_CUDA(ctx, cudaMemcpyAsync(weights_host, p->d_prob, (size_t)p->N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = mct_sync(ctx))) return rc;
    out->total_weight = go.total_weight;
    for (int q = 0; q < 4; q++) out->average[q] = go.average[q];
    out->usable_covariance = ga.usable;
    return MCT_OK;
}

extern "C" int mct_pf_rearrange_ground(mct_pf* p, float mean_foot_x, float mean_foot_y, const float* noise2xN, float w0, float h0)
{
    if (!p || !noise2xN || !(w0 > 0) || !(h0 > 0)) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    if (!p->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaMemcpyAsync(p->d_noise, noise2xN, (size_t)2 * p->N * 4, cudaMemcpyHostToDevice, ctx->stream));
    const ObjDesc* d; StepArgs sa; int rc;
    if ((rc = pf_single_desc(p, &d, &sa, w0, h0, 0, 1, 0.0f))) return rc;
    RearrangeArgs ra; ra.mfx[0] = mean_foot_x; ra.mfy[0] = mean_foot_y; ra.noise_off[0] = 0;
    if ((rc = launch_pf_rearrange(ctx, d, 1, p->N, p->d_noise, ra))) return rc;
    if ((rc = launch_pf_refine(ctx, d, 1))) return rc;                            // :173 CheckForParticleRefinement
    return mct_sync(ctx);     // the host noise buffer may be reused by the caller
}

// the same for several objects of a camera in one pass (the geometric fuser's feedback rearranges every object whose average
// particle lies 20 px from the fused estimate): noise host [sum over objects of 2 * N], object o's block after the blocks of
// the objects in front of it; mean_foot_xy [n_obj][2]
extern "C" int mct_pf_rearrange_ground_many(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const float* mean_foot_xy, const float* noise,
                                            const float* w0, const float* h0)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !mean_foot_xy || !noise || !w0 || !h0) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    ObjDesc h[DESC_MAX_OBJ]; const ObjDesc* d; RearrangeArgs ra;
    size_t ntot = 0; int n_max = 0;
    for (int o = 0; o < n_obj; o++) {
        mct_pf* p = pfs[o];
        if (!p || p->ctx != ctx || !p->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
        if (!(w0[o] > 0) || !(h0[o] > 0)) return MCT_E_ARG;
        memset(&h[o], 0, sizeof(ObjDesc)); fill_pf_desc(&h[o], p);
        h[o].w0 = w0[o]; h[o].h0 = h0[o]; h[o].resample = 1;
        ra.mfx[o] = mean_foot_xy[2 * o]; ra.mfy[o] = mean_foot_xy[2 * o + 1]; ra.noise_off[o] = (long long)ntot;
        ntot += (size_t)2 * p->N;
        if (p->N > n_max) n_max = p->N;
    }
    if (ntot > ctx->noise_stage_cap) {
        MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->d_noise_stage) cudaFree(ctx->d_noise_stage);
        MCT_CUDA(ctx, cudaMalloc(&ctx->d_noise_stage, ntot * sizeof(float)));
        ctx->noise_stage_cap = ntot;
    }
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_noise_stage, noise, ntot * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    int rc;
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    if ((rc = launch_pf_rearrange(ctx, d, n_obj, n_max, ctx->d_noise_stage, ra))) return rc;
    if ((rc = launch_pf_refine(ctx, d, n_obj))) return rc;
    return mct_sync(ctx);     // the host noise buffer may be reused by the caller
}

extern "C" int mct_pf_update_all_weights(mct_pf* p, const float* prob, int n_prob, int only_nonzero, int force_reset,
                                         int resample, float u01, int* resampled)
{
    if (!p || (n_prob > 0 && !prob)) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    if (!p->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
    if (n_prob > p->N) return mct_fail(ctx, MCT_E_ARG, "more probabilities than particles%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_prob > 0) MCT_CUDA(ctx, cudaMemcpyAsync(p->d_prob, prob, (size_t)n_prob * 4, cudaMemcpyHostToDevice, ctx->stream));
    const ObjDesc* d; StepArgs sa; int rc;
    if ((rc = pf_single_desc(p, &d, &sa, 0, 0, force_reset, resample, u01))) return rc;
    if ((rc = launch_pf_expand_prob(ctx, d, p->d_prob, n_prob, only_nonzero))) return rc;
    if ((rc = launch_pf_update(ctx, d, 1, p->N, 0, sa, 0))) return rc;
    PfDev st;
    MCT_CUDA(ctx, cudaMemcpyAsync(&st, p->d_st, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = mct_sync(ctx))) return rc;
    if (resampled) *resampled = st.resampled;
    return MCT_OK;
}

extern "C" int mct_pf_resample(mct_pf* p, int force, float u01)
{
    if (!p) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    if (!p->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    const ObjDesc* d; StepArgs sa; int rc;
    if ((rc = pf_single_desc(p, &d, &sa, 0, 0, 0, 1, u01))) return rc;
    if ((rc = launch_pf_resample_only(ctx, d, 1, p->N, force, sa))) return rc;
    return mct_sync(ctx);
}

// ResampleParticles(force) for all objects of a camera in one launch (Object.cpp:449: the fusion path forces a resampling
// of every object every frame); u01 [n_obj] host
extern "C" int mct_pf_resample_many(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, int force, const float* u01)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !u01) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    ObjDesc h[DESC_MAX_OBJ]; const ObjDesc* d; StepArgs sa;
    memset(&sa, 0, sizeof(sa));
    int n_max = 0;
    for (int o = 0; o < n_obj; o++) {
        if (!pfs[o] || !pfs[o]->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
        memset(&h[o], 0, sizeof(ObjDesc)); fill_pf_desc(&h[o], pfs[o]);
        h[o].resample = 1;
        sa.u01[o] = u01[o];
        if (pfs[o]->N > n_max) n_max = pfs[o]->N;
    }
    int rc;
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    return launch_pf_resample_only(ctx, d, n_obj, n_max, force, sa);
}

extern "C" int mct_pf_get_summary(mct_pf* p, mct_pf_summary* out)
{
    if (!p || !out) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    const ObjDesc* d; StepArgs sa; int rc;
    if ((rc = pf_single_desc(p, &d, &sa, 0, 0, 0, 1, 0.0f))) return rc;
    if ((rc = launch_pf_summary(ctx, d, 1, p->N))) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(out, p->d_summ, sizeof(*out), cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

// read-outs of several objects of a camera: one launch, one synchronisation (Camera::StoreAllParticleFilterTrackerState)
extern "C" int mct_pf_get_summary_many(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, mct_pf_summary* out)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !out) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    ObjDesc h[DESC_MAX_OBJ]; const ObjDesc* d;
    int n_max = 0;
    for (int o = 0; o < n_obj; o++) {
        if (!pfs[o] || pfs[o]->ctx != ctx || !pfs[o]->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
        memset(&h[o], 0, sizeof(ObjDesc)); fill_pf_desc(&h[o], pfs[o]);
        h[o].summ = ctx->h_summ_dev + 2 * DESC_MAX_OBJ + o;    // straight into mapped pinned memory (the stand-alone region)
        h[o].resample = 1;
        if (pfs[o]->N > n_max) n_max = pfs[o]->N;
    }
    int rc;
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    if ((rc = launch_pf_summary(ctx, d, n_obj, n_max))) return rc;
    if ((rc = mct_sync(ctx))) return rc;
    memcpy(out, ctx->h_summ + 2 * DESC_MAX_OBJ, sizeof(mct_pf_summary) * n_obj);
    return MCT_OK;
}

extern "C" int mct_pf_get_last_response(mct_pf* p, float* out, int* valid)
{
    if (!p || !out) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    MCT_CUDA(ctx, cudaMemcpyAsync(out, p->d_H, (size_t)p->N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (valid) MCT_CUDA(ctx, cudaMemcpyAsync(valid, p->d_valid, (size_t)p->N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

// ------------------------------------------------------------------------------------------ fused path
extern "C" int mct_track_score_async(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, mct_clf* const* clfs,
                                     const mct_track_params* params, const float* noise, int noise_is_device,
                                     const float* u01)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !clfs || !params) return MCT_E_ARG;
    if (ctx->frame_id == 0) return mct_fail(ctx, MCT_E_STATE, "no frame set%s");
    const bool rescore = noise == nullptr;      // UpdateParticleWeightsForRearrangedParticles: no prediction step
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc, n_max = 0, K_max = 1;
    size_t ntot = 0;
    ObjDesc h[DESC_MAX_OBJ];
    StepArgs sa;
    for (int o = 0; o < n_obj; o++) {
        mct_pf* p = pfs[o]; mct_clf* c = clfs[o];
        if (!p || !c || p->ctx != ctx || c->ctx != ctx) return mct_fail(ctx, MCT_E_ARG, "object handle belongs to another ctx%s");
        if (!p->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
        if ((rc = need_frame(c))) return rc;
        if ((rc = clf_ensure_test(c, p->N, c->n_color))) return rc;
        memset(&h[o], 0, sizeof(ObjDesc));
        fill_pf_desc(&h[o], p);
        fill_clf_desc(&h[o], c);
        h[o].noise_off = (long long)ntot;
        // MDCH objects of the MIL classifiers: the colour kernel evaluates the selected colour weak classifiers where the bin
        // values are (shared memory) and leaves their responses in featN; the classify kernel only adds them in selector order
        h[o].color_resp = (h[o].slotmap != nullptr && h[o].alpha == nullptr && !getenv("MCT_NO_COLOR_RESP")) ? 1 : 0;
        h[o].summ = ctx->h_summ_dev + o;               // the read-out is written straight into mapped pinned memory (+ StepArgs::summ_off)
        h[o].w0 = params[o].w0; h[o].h0 = params[o].h0; h[o].exponent = params[o].exponent;
        h[o].force_reset = params[o].force_reset; h[o].resample = params[o].resample;
        sa.u01[o] = u01 ? u01[o] : 0.0f;
        ntot += (size_t)4 * p->N;
        if (p->N > n_max) n_max = p->N;
        if (c->K > K_max) K_max = c->K;
    }
    sa.dev_rng = u01 ? 0 : 1;                          // u01 == NULL: every object draws from its device-resident stream (mct_pf_rng_set)
    // read-out slot of this call: two fused score calls may be in flight (the caller queues frame t + 1 before it collects frame t);
    // a third one retires the oldest, whose read-outs are then lost (the free-running throughput path never collects)
    if (ctx->score_issued - ctx->score_collected >= 2) ctx->score_collected = ctx->score_issued - 1;
    const int slot = (int)(ctx->score_issued & 1);
    sa.summ_off = slot * DESC_MAX_OBJ;
    // prediction noise: device pointers are used in place; a block announced with mct_noise_prefetch is already on the
    // device; any other host noise is staged with ONE copy
    if (noise_is_device || rescore) sa.noise_base = noise;
    else if (ctx->npf_req && ctx->prep_stream && ctx->npf_req_src == noise && ctx->npf_req_n == ntot) {
        // announced for THIS call and not uploaded yet (the caller queues the next frame while the current one is still being
        // tracked): upload it now on the copy stream, beside the kernels in flight, instead of in front of this call's first kernel
        const int nb2 = ctx->noise_pf_cur ^ 1;
        if (ctx->npf_req_n > ctx->noise_pf_cap[nb2]) {
            MCT_CUDA(ctx, cudaDeviceSynchronize());
            if (ctx->d_noise_pf[nb2]) cudaFree(ctx->d_noise_pf[nb2]);
            MCT_CUDA(ctx, cudaMalloc(&ctx->d_noise_pf[nb2], ctx->npf_req_n * sizeof(float)));
            ctx->noise_pf_cap[nb2] = ctx->npf_req_n;
        }
        // (the other buffer was last read by the score call two calls ago: the event of this call's slot still marks its end)
        if (ctx->has_ev_score && ctx->score_issued >= 2) MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->prep_stream, ctx->ev_score[slot], 0));
        MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_noise_pf[nb2], noise, ntot * sizeof(float), cudaMemcpyHostToDevice, ctx->prep_stream));
        MCT_CUDA(ctx, cudaEventRecord(ctx->ev_noise, ctx->prep_stream));
        MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_noise, 0));
        ctx->noise_pf_cur = nb2; ctx->npf_req = 0; ctx->npf_ready = 0;
        sa.noise_base = ctx->d_noise_pf[nb2];
    }
    else if (ctx->npf_ready && ctx->npf_ready_src == noise && ctx->npf_ready_n == ntot) {
        MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_noise, 0));
        sa.noise_base = ctx->d_noise_pf[ctx->noise_pf_cur];
        ctx->npf_ready = 0;
    } else {
        if (ntot > ctx->noise_stage_cap) {
            MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->d_noise_stage) cudaFree(ctx->d_noise_stage);
            MCT_CUDA(ctx, cudaMalloc(&ctx->d_noise_stage, ntot * sizeof(float)));
            ctx->noise_stage_cap = ntot;
        }
        MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_noise_stage, noise, ntot * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        sa.noise_base = ctx->d_noise_stage;
    }
    const ObjDesc* d;
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    // Objects with MDCH colour features: the predict / test-set step of ALL objects rides in the prologue of k_mdch_sel
    // (one launch less on the frame's critical path).  CCH objects compute their rows per object and need the boxes first.
    int any_mdch = 0, any_cch = 0, max_slots = 1, nb = 0;
    for (int o = 0; o < n_obj; o++) {
        mct_clf* c = clfs[o];
        if (c->n_color <= 0) continue;
        if (h[o].slotmap) {
            any_mdch = 1; nb = c->p.n_bins;
            int ms = (c->K < c->n_color ? c->K : c->n_color) + 1;
            if (ms > max_slots) max_slots = ms;
        } else any_cch = 1;
    }
    const bool fused_pre = any_mdch && !any_cch;
    if (!fused_pre) {
        if (rescore) { if ((rc = launch_pf_testset(ctx, d, n_obj, n_max))) return rc; }
        else if ((rc = launch_pf_predict_testset(ctx, d, n_obj, n_max, sa))) return rc;
    }
    for (int o = 0; o < n_obj && any_cch; o++) {
        mct_clf* c = clfs[o]; mct_pf* p = pfs[o];
        if (c->n_color <= 0 || h[o].slotmap) continue;
        float* sx = p->d_state + 2 * (size_t)p->ld; float* sy = p->d_state + 3 * (size_t)p->ld;
        if ((rc = mct_feature_color_rows(c, p->d_col, p->d_row, sx, sy, p->d_valid, p->N, nullptr, c->n_color,
                                         c->d_featN, c->ldN)))
            return rc;
    }
    if (any_mdch) {
        if ((rc = launch_bin_image(ctx, nb))) return rc;
        if ((rc = launch_mdch_selected(ctx, d, n_obj, n_max, max_slots, sa, fused_pre ? (rescore ? 2 : 1) : 0))) return rc;
    }
    (void)K_max;
    if ((rc = launch_classify_staged(ctx, d, n_obj, n_max, 1))) return rc;
    // from here on the GPU is nearly idle (the update runs on n_obj SMs): next frame's preparation may start
    if (ctx->prep_stream) {
        MCT_CUDA(ctx, cudaEventRecord(ctx->ev_gather, ctx->stream));
        ctx->has_gather = 1;
    }
    if ((rc = launch_pf_update(ctx, d, n_obj, n_max, 1, sa, 1))) return rc;
    // the read-outs are complete behind this point: mct_track_collect waits for it, not for whatever the caller queues next
    // (the frame's UpdateClassifier when it runs from the device, mct_track_update_default)
    if (!ctx->has_ev_score) {
        MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_score[0], cudaEventDisableTiming));
        MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_score[1], cudaEventDisableTiming));
        ctx->has_ev_score = 1;
    }
    MCT_CUDA(ctx, cudaEventRecord(ctx->ev_score[slot], ctx->stream));
    ctx->score_gen_seq_slot[slot] = ctx->gen_seq;      // every device-side UpdateClassifier queued so far completes in front of that event
    ctx->score_issued++;
    // Only now, with this frame's whole chain queued, the housekeeping for the NEXT frame: every API call in front of the
    // first launch is host time the GPU sits idle for (the caller waits for this frame's summaries before the next call)
    if (ctx->prep_stream && (rc = prefetch_copy(ctx))) return rc;      // next frame's planes (behind this frame's noise copy)
    if (ctx->npf_req && ctx->prep_stream) {
        // next call's noise: into the other prefetch buffer (last read two frames ago; those kernels are covered by the
        // marker of the latest mct_frame_set, which the copy stream waits for)
        const int nb = ctx->noise_pf_cur ^ 1;
        if (ctx->npf_req_n > ctx->noise_pf_cap[nb]) {
            MCT_CUDA(ctx, cudaDeviceSynchronize());
            if (ctx->d_noise_pf[nb]) cudaFree(ctx->d_noise_pf[nb]);
            MCT_CUDA(ctx, cudaMalloc(&ctx->d_noise_pf[nb], ctx->npf_req_n * sizeof(float)));
            ctx->noise_pf_cap[nb] = ctx->npf_req_n;
        }
        MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->prep_stream, ctx->ev_mark[ctx->mark_idx], 0));
        MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_noise_pf[nb], ctx->npf_req_src, ctx->npf_req_n * sizeof(float), cudaMemcpyHostToDevice, ctx->prep_stream));
        MCT_CUDA(ctx, cudaEventRecord(ctx->ev_noise, ctx->prep_stream));
        ctx->noise_pf_cur = nb;
        ctx->npf_ready_src = ctx->npf_req_src; ctx->npf_ready_n = ctx->npf_req_n; ctx->npf_ready = 1; ctx->npf_req = 0;
    }
    if (ctx->prep_stream && ctx->has_gather && (rc = prefetch_prep(ctx))) return rc;
    return MCT_OK;
}

extern "C" int mct_noise_prefetch(mct_ctx* ctx, const float* noise_host, size_t n_floats)
{
    if (!ctx || !noise_host || n_floats == 0) return MCT_E_ARG;
    int rc = frame_streams(ctx);
    if (rc) return rc;
    ctx->npf_req_src = noise_host; ctx->npf_req_n = n_floats; ctx->npf_req = 1;
    return MCT_OK;
}

extern "C" int mct_track_collect(mct_ctx* ctx, int n_obj, mct_pf_summary* summaries)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !summaries) return MCT_E_ARG;
    // the oldest fused score call not collected yet (or, with none pending, the latest one again)
    int slot;
    if (ctx->score_collected < ctx->score_issued) {
        slot = (int)(ctx->score_collected & 1);
        MCT_CUDA(ctx, cudaEventSynchronize(ctx->ev_score[slot]));
        ctx->score_collected++;
    } else {
        slot = (int)((ctx->score_issued - 1) & 1);
        if (ctx->score_issued < 1) slot = 0;
        int rc = mct_sync(ctx); if (rc) return rc;
    }
    memcpy(summaries, ctx->h_summ + slot * DESC_MAX_OBJ, sizeof(mct_pf_summary) * n_obj);
    ctx->score_gen_seq = ctx->score_gen_seq_slot[slot];
    if (ctx->h_gen_out && ctx->score_gen_seq > ctx->gen_checked_seq) {
        // the device-side UpdateClassifier of the PREVIOUS frame has completed (it was queued in front of this frame's kernels):
        // an empty training set it met is reported here, one frame late
        const mct_train_default_out* res = ctx->h_gen_out + (ctx->score_gen_seq & 1) * DESC_MAX_OBJ;
        ctx->gen_checked_seq = ctx->score_gen_seq;
        if (ctx->gen_checked_seq == ctx->gen_seq) ctx->gen_pending = 0;
        for (int o = 0; o < ctx->gen_n; o++) {
            if (ctx->gen_clfs[o] && ctx->gen_checked_seq == ctx->gen_seq) { ctx->gen_clfs[o]->lastP = res[o].P; ctx->gen_clfs[o]->lastNn = res[o].Nn; }
            if (res[o].status != MCT_OK)
                return mct_fail(ctx, res[o].status, "object %s%d: the previous frame's training set was empty (classifier left as it was)", "", o);
        }
    }
    for (int o = 0; o < n_obj; o++)
        if (summaries[o].n_valid == 0)
            return mct_fail(ctx, MCT_E_EMPTY, "object %s%d has no valid test sample (all particles out of frame)", "", o);
    return MCT_OK;
}

extern "C" int mct_track_score(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, mct_clf* const* clfs,
                               const mct_track_params* params, const float* noise, int noise_is_device, const float* u01,
                               mct_pf_summary* summaries)
{
    // MCT_HOST_TIMING=1: host time spent issuing the frame's work vs waiting for its summaries (stderr, every 64 calls)
    static int timing = -1;
    if (timing < 0) { const char* e = getenv("MCT_HOST_TIMING"); timing = e && e[0] == '1'; }
    if (!timing) {
        int rc = mct_track_score_async(ctx, n_obj, pfs, clfs, params, noise, noise_is_device, u01);
        if (rc) return rc;
        return mct_track_collect(ctx, n_obj, summaries);
    }
    static double t_issue = 0, t_wait = 0; static int calls = 0;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    const double a = now();
    int rc = mct_track_score_async(ctx, n_obj, pfs, clfs, params, noise, noise_is_device, u01);
    const double b = now();
    if (rc) return rc;
    rc = mct_track_collect(ctx, n_obj, summaries);
    t_issue += b - a; t_wait += now() - b;
    if (++calls == 64) { fprintf(stderr, "mct_track_score: issue %.1f us, wait %.1f us per call\n", t_issue / calls, t_wait / calls); calls = 0; t_issue = t_wait = 0; }
    return rc;
}

static void fill_train_desc(TrainDesc* t, mct_clf* c)
{
    memset(t, 0, sizeof(*t));
    t->ldT = c->ldT; t->featT = c->d_featT; t->pred = c->d_pred;
    t->rects = c->d_rects; t->wts = c->d_wts; t->nrect = c->d_nrect;
    t->weak = c->d_weak; t->trained = c->d_trained; t->sel = c->d_sel; t->nsel = c->d_nsel;
    t->slotmap = c->d_slotmap; t->colorrow = c->d_colorrow; t->outbins = c->d_outbins; t->nout = c->d_nout; t->rowst = c->d_rowst;
    t->seltab = c->d_seltab; t->selhdr = c->d_selhdr;
    t->ada_cnt = c->d_ada_cnt; t->alpha = c->d_alpha;
    const bool ens = c->p.strong_type == MCT_STRONG_MILENSEMBLE;
    t->keep = ens ? c->d_keep : nullptr; t->ensH = c->d_ensH; t->pct_retained = ens ? c->p.pct_retained / 100.0f : 0.0f;
    t->F = c->F; t->K = c->K; t->n_haar = c->n_haar; t->n_color = c->n_color; t->nb = c->p.n_bins; t->w0 = c->p.w0; t->h0 = c->p.h0;
    t->weak_type = c->p.weak_type; t->strong_type = c->p.strong_type; t->feature_type = c->p.feature_type; t->cch_mode = c->p.cch_mode;
    t->lr = c->p.learning_rate;
}

// Host side of the pixel-list MDCH kernels (k_features.cu, k_mdch_train_*): one group = boxes [first, first + n) of the same
// size; returns false when the group has to take the generic per-box kernel.
static bool mdch_group_plan(MdchGroup* g, const mct_boxes* b, int first, int w0, int h0, int dim)
{
    memset(g, 0, sizeof(*g));
    g->first = first; g->n = b->n;
    if (b->n < 1 || dim > 1024) return false;
    const float sx = b->sx[0], sy = b->sy[0];
    int x0 = b->col[0], x1 = b->col[0], y0 = b->row[0], y1 = b->row[0];
    for (int i = 0; i < b->n; i++) {
        if (b->sx[i] != sx || b->sy[i] != sy) return false;
        x0 = std::min(x0, b->col[i]); x1 = std::max(x1, b->col[i]);
        y0 = std::min(y0, b->row[i]); y1 = std::max(y1, b->row[i]);
    }
    g->rows = (int)lrintf((float)h0 * sy); g->cols = (int)lrintf((float)w0 * sx);      // __float2int_rn(__fmul_rn(.)), as the kernels
    if (g->rows < 1 || g->cols < 1 || g->rows * g->cols > MCT_MDCH_TRAIN_PIX) return false;
    g->x0 = x0; g->y0 = y0; g->RW = x1 - x0 + g->cols; g->RH = y1 - y0 + g->rows;
    g->EW = ((2 * g->RW - g->cols) * (2 * g->RH - g->rows) <= MTR_EXT_WORDS) ? 1 : 0; g->EH = 0;
    g->npx = g->RW * g->RH;
    if (g->npx > 65536 || g->RW > 2047 || g->RH > 2047) return false;
    // one CTA walks chunk_px list entries: S slices side by side (S = threads / boxes per pass), ~160 entries per thread
    const int T = 384;
    const int B = std::min((g->n + 31) & ~31, T), S = T / B;
    g->chunk_px = S * 160;
    g->nchunk = (g->npx + g->chunk_px - 1) / g->chunk_px;
    int shift = __builtin_clz((unsigned)(g->rows * g->cols));
    g->shift = shift > MCT_FIX_SHIFT_MAX ? MCT_FIX_SHIFT_MAX : shift;
    return true;
}

// UpdateClassifier (ParticleFilterTracker.cpp:1013-1032) for all objects of a camera: ONE copy of the packed training
// boxes, ONE descriptor copy and five batched launches (Haar rows, MDCH rows, weak update + predictions, greedy
// selection with one cluster per object, selected-classifier tables) instead of 8 copies + 5 launches per object.
// stages the training boxes of all objects, fills the TrainDescs and launches the feature kernels: afterwards every classifier's
// training matrix d_featT holds its rows [F][P + Nn] (stream-ordered, no synchronisation)
static int train_stage_and_features(mct_ctx* ctx, int n_obj, mct_clf* const* clfs, const mct_boxes* pos, const mct_boxes* neg, std::vector<int>& creq, bool weak_too = false);
static int train_descs_ready(mct_ctx* ctx);

extern "C" int mct_track_update(mct_ctx* ctx, int n_obj, mct_clf* const* clfs, const mct_boxes* pos, const mct_boxes* neg)
{
    if (!ctx || n_obj < 1 || !clfs || !pos || !neg) return MCT_E_ARG;
    bool batchable = n_obj <= DESC_MAX_OBJ;
    size_t total = 0;
    for (int o = 0; o < n_obj; o++) {
        if (!clfs[o] || clfs[o]->ctx != ctx) return mct_fail(ctx, MCT_E_ARG, "classifier handle belongs to another ctx%s");
        if (clfs[o]->p.feature_type == MCT_FEAT_CCH) batchable = false;      // CCH classifiers take the per-object path
        total += (size_t)pos[o].n + (size_t)neg[o].n;
    }
    if (!batchable) {
        for (int o = 0; o < n_obj; o++) {
            int rc = mct_classifier_update(clfs[o], &pos[o], &neg[o], nullptr, nullptr);
            if (rc) return rc;
        }
        return MCT_OK;
    }
    std::vector<int> creq;
    int rc;
    if ((rc = train_stage_and_features(ctx, n_obj, clfs, pos, neg, creq, true))) return rc;
    if ((rc = launch_ensemble_retain_batch(ctx, ctx->d_tdesc, ctx->h_tdesc, n_obj))) return rc;
    if ((rc = launch_weak_update_batch(ctx, ctx->d_tdesc, ctx->h_tdesc, n_obj))) return rc;
    if ((rc = launch_mil_select_batch(ctx, ctx->d_tdesc, ctx->h_tdesc, n_obj, creq.data()))) return rc;
    if ((rc = launch_build_colormap_batch(ctx, ctx->d_tdesc, ctx->h_tdesc, n_obj))) return rc;
    return MCT_OK;
}

static void stage_take_parked(mct_ctx* ctx, size_t need_boxes)
{
    std::lock_guard<std::mutex> lk(g_stage_mu);
    if (ctx->device >= MCT_MAX_DEV) return;
    StageCache& sc = g_stage_cache[ctx->device];
    if (!sc.h_stage || sc.cap < need_boxes) return;
    ctx->h_train_stage = sc.h_stage; ctx->d_train_stage = sc.d_stage; ctx->train_stage_cap = sc.cap;
    sc.h_stage = nullptr; sc.d_stage = nullptr; sc.cap = 0;
}
static int train_descs_ready(mct_ctx* ctx)
{
    if (!ctx->h_tdesc) {
        {
            std::lock_guard<std::mutex> lk(g_stage_mu);
            StageCache& sc = g_stage_cache[ctx->device < MCT_MAX_DEV ? ctx->device : 0];
            if (sc.h_tdesc && ctx->device < MCT_MAX_DEV) { ctx->h_tdesc = sc.h_tdesc; ctx->d_tdesc = sc.d_tdesc; sc.h_tdesc = nullptr; sc.d_tdesc = nullptr; }
        }
        if (!ctx->h_tdesc) {
            MCT_CUDA(ctx, cudaMallocHost(&ctx->h_tdesc, sizeof(TrainDesc) * DESC_MAX_OBJ));
            MCT_CUDA(ctx, cudaMalloc(&ctx->d_tdesc, sizeof(TrainDesc) * DESC_MAX_OBJ));
        }
        MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_train, cudaEventDisableTiming));
        MCT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->train_stream, cudaStreamNonBlocking));
        MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_tfork, cudaEventDisableTiming));
        MCT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_tjoin, cudaEventDisableTiming));
    }
    MCT_CUDA(ctx, cudaEventSynchronize(ctx->ev_train));                // the previous call's staging copies have executed
    return MCT_OK;
}

static int train_stage_and_features(mct_ctx* ctx, int n_obj, mct_clf* const* clfs, const mct_boxes* pos, const mct_boxes* neg, std::vector<int>& creq_out, bool weak_too)
{
    size_t total = 0;
    for (int o = 0; o < n_obj; o++) total += (size_t)pos[o].n + (size_t)neg[o].n;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = train_descs_ready(ctx))) return rc;
    if (total > ctx->train_stage_cap && !ctx->h_train_stage) stage_take_parked(ctx, total);
    if (total > ctx->train_stage_cap) {
        MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->h_train_stage) cudaFreeHost(ctx->h_train_stage);
        if (ctx->d_train_stage) cudaFree(ctx->d_train_stage);
        size_t cap = total + 1024;
        MCT_CUDA(ctx, cudaMallocHost(&ctx->h_train_stage, cap * 4 * sizeof(int)));
        MCT_CUDA(ctx, cudaMalloc(&ctx->d_train_stage, cap * 4 * sizeof(int)));
        ctx->train_stage_cap = cap;
    }
    // pack: per object [col | row | sx | sy] x (P + Nn)
    size_t off = 0, mlist_words = 0, macc_words = 0;
    std::vector<int> creq(n_obj, 0);
    for (int o = 0; o < n_obj; o++) {
        mct_clf* c = clfs[o];
        if ((rc = need_frame(c))) return rc;
        const int P = pos[o].n, Nn = neg[o].n, M = P + Nn;
        if (P < 1 || Nn < 1) return mct_fail(ctx, MCT_E_ARG, "Update needs at least one positive and one negative sample%s");
        if (M > c->max_train) return mct_fail(ctx, MCT_E_ARG, "training set exceeds max_train%s (%d)", "", M);
        if (!pos[o].col || !pos[o].row || !pos[o].sx || !pos[o].sy || !neg[o].col || !neg[o].row || !neg[o].sx || !neg[o].sy)
            return mct_fail(ctx, MCT_E_ARG, "box arrays are NULL%s");
        int* hs = ctx->h_train_stage + off * 4;
        memcpy(hs, pos[o].col, (size_t)P * 4);             memcpy(hs + P, neg[o].col, (size_t)Nn * 4);
        memcpy(hs + M, pos[o].row, (size_t)P * 4);         memcpy(hs + M + P, neg[o].row, (size_t)Nn * 4);
        memcpy(hs + 2 * M, pos[o].sx, (size_t)P * 4);      memcpy(hs + 2 * M + P, neg[o].sx, (size_t)Nn * 4);
        memcpy(hs + 3 * M, pos[o].sy, (size_t)P * 4);      memcpy(hs + 3 * M + P, neg[o].sy, (size_t)Nn * 4);
        TrainDesc* t = &ctx->h_tdesc[o];
        fill_train_desc(t, c);
        int* ds = ctx->d_train_stage + off * 4;
        t->col = ds; t->row = ds + M; t->sx = (const float*)(ds + 2 * M); t->sy = (const float*)(ds + 3 * M); t->tw = nullptr;
        t->P = P; t->Nn = Nn;
        c->lastP = P; c->lastNn = Nn;
        creq[o] = c->cluster;
        off += (size_t)M;
        // MDCH training rows: both groups of equal-size boxes inside a small region -> pixel-list kernels
        t->mdch_fast = 0;
        if ((c->p.feature_type == MCT_FEAT_MDCH || c->p.feature_type == MCT_FEAT_HAAR_MDCH) && !getenv("MCT_MDCH_TRAIN_GENERIC")) {
            const bool ok0 = mdch_group_plan(&t->grp[0], &pos[o], 0, c->p.w0, c->p.h0, c->n_color);
            const bool ok1 = mdch_group_plan(&t->grp[1], &neg[o], P, c->p.w0, c->p.h0, c->n_color);
            if (ok0 && ok1) {
                t->mdch_fast = 1;
                t->grp[0].list_off = 0; t->grp[1].list_off = t->grp[0].npx;
                mlist_words += (size_t)t->grp[0].npx + t->grp[1].npx;
                macc_words += (size_t)c->n_color * c->ldT;
            }
        }
    }
    // work buffers of the pixel-list kernels (accumulators are zero between launches: k_mdch_train_finish clears what it reads)
    if (mlist_words > ctx->mlist_cap || macc_words > ctx->macc_cap || (mlist_words > 0 && !ctx->d_mtot)) {
        MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (mlist_words > ctx->mlist_cap) {
            if (ctx->d_mlist) cudaFree(ctx->d_mlist);
            MCT_CUDA(ctx, cudaMalloc(&ctx->d_mlist, (mlist_words + 4096) * 4)); ctx->mlist_cap = mlist_words + 4096;
        }
        if (macc_words > ctx->macc_cap) {
            if (ctx->d_macc) cudaFree(ctx->d_macc);
            MCT_CUDA(ctx, cudaMalloc(&ctx->d_macc, (macc_words + 4096) * 4)); ctx->macc_cap = macc_words + 4096;
            MCT_CUDA(ctx, cudaMemset(ctx->d_macc, 0, ctx->macc_cap * 4));
        }
        if (!ctx->d_mtot) MCT_CUDA(ctx, cudaMalloc(&ctx->d_mtot, (size_t)DESC_MAX_OBJ * 2 * 4));
    }
    {
        size_t lo = 0, ao = 0;
        for (int o = 0; o < n_obj; o++) {
            TrainDesc* t = &ctx->h_tdesc[o];
            if (!t->mdch_fast) continue;
            t->mlist = ctx->d_mlist + lo; t->macc = ctx->d_macc + ao; t->mtot = ctx->d_mtot + 2 * o;
            lo += (size_t)t->grp[0].npx + t->grp[1].npx; ao += (size_t)clfs[o]->n_color * clfs[o]->ldT;
        }
    }
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_train_stage, ctx->h_train_stage, total * 4 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_tdesc, ctx->h_tdesc, sizeof(TrainDesc) * n_obj, cudaMemcpyHostToDevice, ctx->stream));
    ctx->tdesc_gen_valid = 0;
    MCT_CUDA(ctx, cudaEventRecord(ctx->ev_train, ctx->stream));
    if ((rc = launch_train_features(ctx, ctx->d_tdesc, ctx->h_tdesc, n_obj, nullptr, weak_too))) return rc;
    creq_out = creq;
    return MCT_OK;
}

// Appearance fuser, every camera's side of one learner round (CameraNetwork.cpp:276-320 with SampleSets replaced by feature rows):
// the feature rows of all objects' freshly generated training sets, computed in ONE batched pass under each object's global
// classifier and written into the exchange buffer: object o at dev_rows + o * obj_stride floats, rows `ld` floats apart,
// positives from column 0, negatives from column ld_pos.  Stream-ordered, no synchronisation; the box arrays may be reused.
// reuse_pos_rows != NULL: rows of an earlier call (same layout) whose POSITIVE boxes were identical -- the learner rounds of one
// frame regenerate the same positives and only re-thin the negative ring; the positive columns are then copied from there and
// only the negative boxes go through the feature kernels.
extern "C" int mct_train_features_many_dev(mct_ctx* ctx, int n_obj, mct_clf* const* clfs, const mct_boxes* pos, const mct_boxes* neg,
                                           float* dev_rows, size_t obj_stride, int ld, int ld_pos, const float* reuse_pos_rows)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !clfs || !pos || !neg || !dev_rows || ld_pos < 1 || ld <= ld_pos) return MCT_E_ARG;
    for (int o = 0; o < n_obj; o++) {
        if (!clfs[o] || clfs[o]->ctx != ctx) return mct_fail(ctx, MCT_E_ARG, "classifier handle belongs to another ctx%s");
        if (clfs[o]->p.feature_type == MCT_FEAT_CCH) return mct_fail(ctx, MCT_E_UNSUPPORTED, "batched training rows: culture colour classifiers take mct_features_compute_dev%s");
        if (pos[o].n > ld_pos || neg[o].n > ld - ld_pos) return mct_fail(ctx, MCT_E_ARG, "training set larger than the exchange buffer%s");
    }
    std::vector<int> creq;
    int rc;
    std::vector<mct_boxes> pos1;
    if (reuse_pos_rows) {                       // one positive box keeps the staging / planning code on its usual path
        pos1.assign(pos, pos + n_obj);
        for (int o = 0; o < n_obj; o++) if (pos1[o].n > 1) pos1[o].n = 1;
    }
    const mct_boxes* pp = reuse_pos_rows ? pos1.data() : pos;
    if ((rc = train_stage_and_features(ctx, n_obj, clfs, pp, neg, creq))) return rc;
    for (int o = 0; o < n_obj; o++) {
        mct_clf* c = clfs[o];
        float* base = dev_rows + (size_t)o * obj_stride;
        if (reuse_pos_rows)
            MCT_CUDA(ctx, cudaMemcpy2DAsync(base, (size_t)ld * 4, reuse_pos_rows + (size_t)o * obj_stride, (size_t)ld * 4, (size_t)pos[o].n * 4, c->F,
                                            cudaMemcpyDeviceToDevice, ctx->stream));
        else
            MCT_CUDA(ctx, cudaMemcpy2DAsync(base, (size_t)ld * 4, c->d_featT, (size_t)c->ldT * 4, (size_t)pos[o].n * 4, c->F, cudaMemcpyDeviceToDevice, ctx->stream));
        MCT_CUDA(ctx, cudaMemcpy2DAsync(base + ld_pos, (size_t)ld * 4, c->d_featT + pp[o].n, (size_t)c->ldT * 4, (size_t)neg[o].n * 4, c->F,
                                        cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return MCT_OK;
}

// The learner's side for all its objects at once: per object the kept columns are gathered into the classifier's training matrix
// (see mct_classifier_update_from_gathered), then ONE batched Update (weak classifiers, selection, colour map) for all of them.
// col_offsets_host: the objects' offset lists back to back (P[0] + Nn[0] entries, then P[1] + Nn[1], ...).
extern "C" int mct_track_update_from_gathered(mct_ctx* ctx, int n_obj, mct_clf* const* clfs, const float* dev_rows, const int* col_offsets_host,
                                              long long row_stride, const int* P, const int* Nn)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !clfs || !dev_rows || !col_offsets_host || !P || !Nn || row_stride < 1) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = train_descs_ready(ctx))) return rc;
    std::vector<int> creq(n_obj, 0);
    size_t total = 0;
    int M_max = 0, F_max = 0;
    for (int o = 0; o < n_obj; o++) {
        mct_clf* c = clfs[o];
        if (!c || c->ctx != ctx) return mct_fail(ctx, MCT_E_ARG, "classifier handle belongs to another ctx%s");
        if (P[o] < 1 || Nn[o] < 1) return mct_fail(ctx, MCT_E_ARG, "Update needs at least one positive and one negative sample%s");
        const int M = P[o] + Nn[o];
        if (M > c->max_train) return mct_fail(ctx, MCT_E_ARG, "training set exceeds max_train%s (%d)", "", M);
        total += (size_t)M; M_max = std::max(M_max, M); F_max = std::max(F_max, c->F);
    }
    // the column offsets of all objects travel with ONE copy through the pinned box staging buffer (one copy + one gather launch
    // per object before: 16 small copies from pageable memory and 16 launches per camera-frame at configs[3])
    const size_t need_boxes = (total + 3) / 4 + 1;
    if (need_boxes > ctx->train_stage_cap && !ctx->h_train_stage) stage_take_parked(ctx, need_boxes);
    if (need_boxes > ctx->train_stage_cap) {
        MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->h_train_stage) cudaFreeHost(ctx->h_train_stage);
        if (ctx->d_train_stage) cudaFree(ctx->d_train_stage);
        const size_t cap = need_boxes + 1024;
        MCT_CUDA(ctx, cudaMallocHost(&ctx->h_train_stage, cap * 4 * sizeof(int)));
        MCT_CUDA(ctx, cudaMalloc(&ctx->d_train_stage, cap * 4 * sizeof(int)));
        ctx->train_stage_cap = cap;
    }
    memcpy(ctx->h_train_stage, col_offsets_host, total * sizeof(int));
    size_t off = 0;
    for (int o = 0; o < n_obj; o++) {
        mct_clf* c = clfs[o];
        const int M = P[o] + Nn[o];
        TrainDesc* t = &ctx->h_tdesc[o];
        fill_train_desc(t, c);
        t->col = ctx->d_train_stage + off; t->row = nullptr; t->sx = nullptr; t->sy = nullptr; t->tw = nullptr;
        t->P = P[o]; t->Nn = Nn[o]; t->mdch_fast = 0;
        c->lastP = P[o]; c->lastNn = Nn[o];
        creq[o] = c->cluster;
        off += (size_t)M;
    }
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_train_stage, ctx->h_train_stage, total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_tdesc, ctx->h_tdesc, sizeof(TrainDesc) * n_obj, cudaMemcpyHostToDevice, ctx->stream));
    ctx->tdesc_gen_valid = 0;
    MCT_CUDA(ctx, cudaEventRecord(ctx->ev_train, ctx->stream));
    MCT_LAUNCH(ctx, "k_gather_columns", k_gather_columns_batch<<<dim3(ceil_div(M_max, 256), F_max < 64 ? F_max : 64, n_obj), 256, 0, ctx->stream>>>(
                   ctx->d_tdesc, dev_rows, row_stride));
    ctx->weak_split_done = 0;                          // (no feature streams here: the weak update below is the whole one)
    if ((rc = launch_ensemble_retain_batch(ctx, ctx->d_tdesc, ctx->h_tdesc, n_obj))) return rc;
    if ((rc = launch_weak_update_batch(ctx, ctx->d_tdesc, ctx->h_tdesc, n_obj))) return rc;
    if ((rc = launch_mil_select_batch(ctx, ctx->d_tdesc, ctx->h_tdesc, n_obj, creq.data()))) return rc;
    if ((rc = launch_build_colormap_batch(ctx, ctx->d_tdesc, ctx->h_tdesc, n_obj))) return rc;
    return MCT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The whole frame on the device: the tracker's stream as device-resident state, and UpdateClassifier with the
// default-strategy training sets generated by k_train_gen_default (k_pf.cu) behind the frame's k_pf_update.
extern "C" int mct_pf_rng_set(mct_pf* p, unsigned long long state)
{
    if (!p) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaMemcpyAsync(&p->d_trk->rng, &state, sizeof(state), cudaMemcpyHostToDevice, ctx->stream));
    return mct_sync(ctx);                       // `state` is a stack value of the caller's frame
}
extern "C" int mct_pf_rng_get(mct_pf* p, unsigned long long* state)
{
    if (!p || !state) return MCT_E_ARG;
    mct_ctx* ctx = p->ctx;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    MCT_CUDA(ctx, cudaMemcpyAsync(state, &p->d_trk->rng, sizeof(*state), cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

// candidates of SampleSet::SampleImage's ring form when the frame does not clip it (SampleSet.cpp:106-140)
static int ring_count_unclipped(float outer, float inner)
{
    static float s_o[8], s_i[8]; static int s_n[8], s_used = 0;      // a handful of (outer, inner) pairs per process
    for (int k = 0; k < s_used; k++) if (s_o[k] == outer && s_i[k] == inner) return s_n[k];
    const int R = (int)outer;
    const float o2 = outer * outer, i2 = inner * inner;
    int n = 0;
    for (int dy = -R; dy <= R; dy++)
        for (int dx = -R; dx <= R; dx++) { const int d = dx * dx + dy * dy; n += ((float)d < o2 && (float)d >= i2) ? 1 : 0; }
    if (s_used < 8) { s_o[s_used] = outer; s_i[s_used] = inner; s_n[s_used] = n; s_used++; }
    return n;
}

extern "C" int mct_track_update_default(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, mct_clf* const* clfs, const mct_train_default* gen)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !clfs || !gen) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if (!ctx->d_gen) {
        MCT_CUDA(ctx, cudaMalloc(&ctx->d_gen, sizeof(TrainGenDev) * DESC_MAX_OBJ));
        MCT_CUDA(ctx, cudaHostAlloc(&ctx->h_gen_out, sizeof(mct_train_default_out) * DESC_MAX_OBJ * 2, cudaHostAllocMapped));
        MCT_CUDA(ctx, cudaHostGetDevicePointer(&ctx->h_gen_out_dev, ctx->h_gen_out, 0));
        memset(ctx->h_gen_out, 0, sizeof(mct_train_default_out) * DESC_MAX_OBJ * 2);
        ctx->gen_shadow = calloc(DESC_MAX_OBJ, sizeof(TrainGenDev));
        ctx->tdesc_shadow = calloc(DESC_MAX_OBJ, sizeof(TrainDesc));
        if (!ctx->gen_shadow || !ctx->tdesc_shadow) return mct_fail(ctx, MCT_E_NOMEM, "out of host memory%s");
    }
    if (!ctx->h_tdesc && (rc = train_descs_ready(ctx))) return rc;
    std::vector<TrainGenDev> hg(n_obj); std::vector<TrainDesc> ht(n_obj);
    TrainGenDev* g = hg.data(); TrainDesc* tds = ht.data();
    std::vector<int> creq(n_obj, 0);
    TrainLaunchBounds lb = {0, 0, 0, 0, 0, 0};
    size_t stage_words = 0, mlist_words = 0, macc_words = 0;
    for (int o = 0; o < n_obj; o++) {
        mct_pf* p = pfs[o]; mct_clf* c = clfs[o];
        if (!p || !c || p->ctx != ctx || c->ctx != ctx) return mct_fail(ctx, MCT_E_ARG, "object handle belongs to another ctx%s");
        if (!p->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
        if (c->p.feature_type == MCT_FEAT_CCH) return mct_fail(ctx, MCT_E_UNSUPPORTED, "device-side training sets: culture colour classifiers take mct_track_update%s");
        if ((rc = need_frame(c))) return rc;
        const mct_train_default& q = gen[o];
        if (!(q.pos_radius > 0.0f) || !(q.neg_outer > q.neg_inner) || q.max_pos < 0 || q.max_neg < 0) return mct_fail(ctx, MCT_E_ARG, "training-sample radii / counts out of range%s");
        const int cnt_p = ring_count_unclipped(q.pos_radius, 0.0f), cnt_n = ring_count_unclipped(q.neg_outer, q.neg_inner);
        const int Pb = std::max(1, (int)std::min<long long>(cnt_p, (long long)q.max_pos + 1)), Nb = std::max(1, (int)std::min<long long>(cnt_n, (long long)q.max_neg + 1));
        const int Mb = Pb + Nb;
        if (Mb > c->max_train) return mct_fail(ctx, MCT_E_ARG, "training set exceeds max_train%s (%d)", "", Mb);
        const int mcap = round_up(Mb, 4);
        TrainGenDev& d = g[o];
        memset(&d, 0, sizeof(d));
        d.trk = p->d_trk; d.st = p->d_st; d.mcap = mcap; d.cap_pos = Pb; d.cap_neg = Nb;
        d.w0 = q.w0; d.h0 = q.h0; d.opt = q.trajectory_option;
        d.pos_radius = q.pos_radius; d.max_pos = q.max_pos; d.neg_outer = q.neg_outer; d.neg_inner = q.neg_inner; d.max_neg = q.max_neg;
        d.maxs = p->maxs; d.frame_w = ctx->W; d.frame_h = ctx->H;
        d.cw0 = c->p.w0; d.ch0 = c->p.h0; d.dim = c->n_color;
        d.mdch = (c->p.feature_type == MCT_FEAT_MDCH || c->p.feature_type == MCT_FEAT_HAAR_MDCH) && !getenv("MCT_MDCH_TRAIN_GENERIC");
        TrainDesc* t = &tds[o];
        fill_train_desc(t, c);
        t->P = Pb; t->Nn = Nb;                       // upper bounds: the launches are sized from them, the kernel writes the real sizes
        t->mdch_fast = d.mdch;
        creq[o] = c->cluster;
        lb.n_max = std::max(lb.n_max, Mb);
        if (d.mdch) {
            // bounds of the pixel-list plan from the hinted scale (+ 2 pixels of slack on the box size)
            float shx = q.scale_hint_x > 0 ? q.scale_hint_x : 1.0f, shy = q.scale_hint_y > 0 ? q.scale_hint_y : 1.0f;
            // (test hook: a deliberately wrong hint, so that the plans leave the launch bounds and the objects take the generic kernel)
            if (const char* e = getenv("MCT_TEST_HINT_SCALE")) { const float k = (float)atof(e); if (k > 0) { shx *= k; shy *= k; } }
            const float hsx[2] = {shx, std::min(shx, p->maxs)}, hsy[2] = {shy, std::min(shy, p->maxs)};
            const int R[2] = {(int)q.pos_radius, (int)q.neg_outer}, nb_[2] = {Pb, Nb};
            for (int gi = 0; gi < 2; gi++) {
                const int cols = (int)lrintf((float)c->p.w0 * hsx[gi]) + 2, rows = (int)lrintf((float)c->p.h0 * hsy[gi]) + 2;
                const int RW = 2 * R[gi] + cols, RH = 2 * R[gi] + rows;
                const int npx = std::min(RW * RH, 65536);
                d.npx_b[gi] = npx;
                const int ext = (2 * RW - cols) * (2 * RH - rows);
                d.tab_b = std::max(d.tab_b, ext <= MTR_EXT_WORDS ? ext : std::min(rows * cols, MCT_MDCH_TRAIN_PIX));
                const int B = std::min((nb_[gi] + 31) & ~31, 384), S = 384 / B;
                d.nch_b = std::max(d.nch_b, (npx + S * 160 - 1) / (S * 160));
                lb.npx_max = std::max(lb.npx_max, npx);
            }
            lb.nch = std::max(lb.nch, d.nch_b); lb.tab_words = std::max(lb.tab_words, d.tab_b);
            lb.chunk_words = 12 * 160;               // the longest chunk a group can ask for (32 boxes or fewer: 12 slices of 160)
            lb.n_train_max = std::max(lb.n_train_max, Mb);
            mlist_words += (size_t)d.npx_b[0] + d.npx_b[1];
            macc_words += (size_t)c->n_color * c->ldT;
        }
        stage_words += (size_t)4 * mcap;
    }
    // buffers (grow-only; growing synchronises)
    if (stage_words > ctx->train_stage_cap * 4 && !ctx->h_train_stage) stage_take_parked(ctx, stage_words / 4 + 1);
    if (stage_words > ctx->train_stage_cap * 4) {
        MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->h_train_stage) cudaFreeHost(ctx->h_train_stage);
        if (ctx->d_train_stage) cudaFree(ctx->d_train_stage);
        const size_t cap = stage_words / 4 + 1024;
        MCT_CUDA(ctx, cudaMallocHost(&ctx->h_train_stage, cap * 4 * sizeof(int)));
        MCT_CUDA(ctx, cudaMalloc(&ctx->d_train_stage, cap * 4 * sizeof(int)));
        ctx->train_stage_cap = cap;
        ctx->tdesc_gen_valid = 0;
    }
    if (mlist_words > ctx->mlist_cap || macc_words > ctx->macc_cap || !ctx->d_mtot) {
        MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (mlist_words > ctx->mlist_cap) {
            if (ctx->d_mlist) cudaFree(ctx->d_mlist);
            MCT_CUDA(ctx, cudaMalloc(&ctx->d_mlist, (mlist_words + 4096) * 4)); ctx->mlist_cap = mlist_words + 4096;
        }
        if (macc_words > ctx->macc_cap) {
            if (ctx->d_macc) cudaFree(ctx->d_macc);
            MCT_CUDA(ctx, cudaMalloc(&ctx->d_macc, (macc_words + 4096) * 4)); ctx->macc_cap = macc_words + 4096;
            MCT_CUDA(ctx, cudaMemset(ctx->d_macc, 0, ctx->macc_cap * 4));
        }
        if (!ctx->d_mtot) MCT_CUDA(ctx, cudaMalloc(&ctx->d_mtot, (size_t)DESC_MAX_OBJ * 2 * 4));
        ctx->tdesc_gen_valid = 0;
    }
    {
        size_t so = 0, lo = 0, ao = 0;
        for (int o = 0; o < n_obj; o++) {
            TrainGenDev& d = g[o]; TrainDesc* t = &tds[o];
            d.stage = ctx->d_train_stage + so;
            t->col = d.stage; t->row = d.stage + d.mcap; t->sx = (const float*)(d.stage + 2 * d.mcap); t->sy = (const float*)(d.stage + 3 * d.mcap); t->tw = nullptr;
            so += (size_t)4 * d.mcap;
            if (d.mdch) {
                t->mlist = ctx->d_mlist + lo; t->macc = ctx->d_macc + ao; t->mtot = ctx->d_mtot + 2 * o;
                lo += (size_t)d.npx_b[0] + d.npx_b[1]; ao += (size_t)clfs[o]->n_color * clfs[o]->ldT;
            }
        }
    }
    // arguments and descriptors travel only when they differ from what the device already holds (the kernel rewrites every
    // per-frame field of the descriptors itself)
    bool fresh_args = false;
    if (ctx->gen_shadow_n != n_obj || memcmp(ctx->gen_shadow, g, sizeof(TrainGenDev) * n_obj) != 0) {
        fresh_args = true;
        memcpy(ctx->gen_shadow, g, sizeof(TrainGenDev) * n_obj); ctx->gen_shadow_n = n_obj;
        MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_gen, ctx->gen_shadow, sizeof(TrainGenDev) * n_obj, cudaMemcpyHostToDevice, ctx->stream));
        MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // pageable source: rare (a changed hint or object set)
    }
    if (!ctx->tdesc_gen_valid || ctx->tdesc_shadow_n != n_obj || memcmp(ctx->tdesc_shadow, tds, sizeof(TrainDesc) * n_obj) != 0) {
        fresh_args = true;
        if ((rc = train_descs_ready(ctx))) return rc;
        memcpy(ctx->tdesc_shadow, tds, sizeof(TrainDesc) * n_obj); ctx->tdesc_shadow_n = n_obj;
        memcpy(ctx->h_tdesc, tds, sizeof(TrainDesc) * n_obj);
        MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_tdesc, ctx->h_tdesc, sizeof(TrainDesc) * n_obj, cudaMemcpyHostToDevice, ctx->stream));
        MCT_CUDA(ctx, cudaEventRecord(ctx->ev_train, ctx->stream));
        ctx->tdesc_gen_valid = 1;
    }
    for (int o = 0; o < n_obj; o++) { ctx->gen_clfs[o] = clfs[o]; clfs[o]->lastP = -1; clfs[o]->lastNn = -1; }
    ctx->gen_n = n_obj; ctx->gen_pending = 1; ctx->gen_seq++;
    if ((rc = launch_train_gen_default(ctx, n_obj, ctx->d_gen, ctx->d_tdesc, ctx->h_gen_out_dev + (ctx->gen_seq & 1) * DESC_MAX_OBJ, fresh_args))) return rc;
    if ((rc = launch_train_features(ctx, ctx->d_tdesc, tds, n_obj, &lb, true))) return rc;
    if ((rc = launch_ensemble_retain_batch(ctx, ctx->d_tdesc, tds, n_obj))) return rc;
    if ((rc = launch_weak_update_batch(ctx, ctx->d_tdesc, tds, n_obj))) return rc;
    if ((rc = launch_mil_select_batch(ctx, ctx->d_tdesc, tds, n_obj, creq.data()))) return rc;
    if ((rc = launch_build_colormap_batch(ctx, ctx->d_tdesc, tds, n_obj))) return rc;
    return MCT_OK;
}

extern "C" int mct_track_update_default_collect(mct_ctx* ctx, int n_obj, mct_train_default_out* out)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ) return MCT_E_ARG;
    if (!ctx->h_gen_out || ctx->gen_n < n_obj) return mct_fail(ctx, MCT_E_STATE, "no device-side UpdateClassifier to collect%s");
    int rc = mct_sync(ctx);
    if (rc) return rc;
    ctx->gen_pending = 0; ctx->gen_checked_seq = ctx->gen_seq;
    const mct_train_default_out* res = ctx->h_gen_out + (ctx->gen_seq & 1) * DESC_MAX_OBJ;
    if (out) memcpy(out, res, sizeof(mct_train_default_out) * n_obj);
    for (int o = 0; o < n_obj; o++) {
        const mct_train_default_out& r = res[o];
        if (ctx->gen_clfs[o]) { ctx->gen_clfs[o]->lastP = r.status == MCT_OK ? r.P : 0; ctx->gen_clfs[o]->lastNn = r.status == MCT_OK ? r.Nn : 0; }
    }
    for (int o = 0; o < n_obj; o++)
        if (res[o].status != MCT_OK)
            return mct_fail(ctx, res[o].status, "object %s%d: empty training set (classifier left as it was)", "", o);
    return MCT_OK;
}

// the boxes k_train_gen_default produced for object `obj` of the last mct_track_update_default (tests / diagnostics): positives
// first, then negatives; synchronises
extern "C" int mct_track_update_default_boxes(mct_ctx* ctx, int obj, int cap, int* col, int* row, float* sx, float* sy, int* P, int* Nn)
{
    if (!ctx || obj < 0 || cap < 0 || !P || !Nn) return MCT_E_ARG;
    if (!ctx->h_gen_out || obj >= ctx->gen_n || obj >= ctx->gen_shadow_n) return mct_fail(ctx, MCT_E_STATE, "no device-side UpdateClassifier to read%s");
    int rc = mct_sync(ctx);
    if (rc) return rc;
    const mct_train_default_out& r = ctx->h_gen_out[(ctx->gen_seq & 1) * DESC_MAX_OBJ + obj];
    *P = r.P; *Nn = r.Nn;
    const TrainGenDev& g = ((const TrainGenDev*)ctx->gen_shadow)[obj];
    const int n = std::min(r.P + r.Nn, cap);
    if (n <= 0) return MCT_OK;
    if (col) MCT_CUDA(ctx, cudaMemcpy(col, g.stage, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (row) MCT_CUDA(ctx, cudaMemcpy(row, g.stage + g.mcap, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (sx) MCT_CUDA(ctx, cudaMemcpy(sx, g.stage + 2 * g.mcap, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (sy) MCT_CUDA(ctx, cudaMemcpy(sy, g.stage + 3 * g.mcap, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return MCT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Peer-memory exchange of the geometric fuser's payload (no NCCL): every camera's receive buffer is exported with CUDA IPC,
// the cameras open each other's buffers once, and per frame ONE kernel computes the foot points and stores them into all
// buffers over NVLink (k_pf.cu).
extern "C" int mct_p2p_create(mct_ctx* ctx, int world, int rank, unsigned char* ipc_handle_out64)
{
    if (!ctx || world < 1 || world > MCT_P2P_MAXW || rank < 0 || rank >= world || !ipc_handle_out64) return MCT_E_ARG;
    if (ctx->p2p_world) return mct_fail(ctx, MCT_E_STATE, "peer exchange already created%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    P2PBuf* mine = nullptr;
    MCT_CUDA(ctx, cudaMalloc(&mine, sizeof(P2PBuf)));
    MCT_CUDA(ctx, cudaMemset(mine, 0, sizeof(P2PBuf)));
    MCT_CUDA(ctx, cudaMalloc(&ctx->d_p2p_ticket, 4)); MCT_CUDA(ctx, cudaMemset(ctx->d_p2p_ticket, 0, 4));
    MCT_CUDA(ctx, cudaHostAlloc(&ctx->h_p2p_out, sizeof(float) * MCT_P2P_MAXW * 64 * 2, cudaHostAllocMapped));
    MCT_CUDA(ctx, cudaHostGetDevicePointer(&ctx->h_p2p_out_dev, ctx->h_p2p_out, 0));
    MCT_CUDA(ctx, cudaHostAlloc(&ctx->h_p2p_status, sizeof(int), cudaHostAllocMapped));
    MCT_CUDA(ctx, cudaHostGetDevicePointer(&ctx->h_p2p_status_dev, ctx->h_p2p_status, 0));
    MCT_CUDA(ctx, cudaDeviceSynchronize());
    cudaIpcMemHandle_t hnd;
    MCT_CUDA(ctx, cudaIpcGetMemHandle(&hnd, mine));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_out64, &hnd, 64);
    memset(&ctx->p2p, 0, sizeof(ctx->p2p));
    ctx->p2p.p[rank] = mine; ctx->p2p_world = world; ctx->p2p_rank = rank; ctx->p2p_seq = 0;
    return MCT_OK;
}
extern "C" int mct_p2p_open(mct_ctx* ctx, const unsigned char* all_handles /* [world][64] */)
{
    if (!ctx || !all_handles || !ctx->p2p_world) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->p2p_world; r++) {
        if (r == ctx->p2p_rank) continue;
        cudaIpcMemHandle_t hnd; memcpy(&hnd, all_handles + (size_t)r * 64, 64);
        void* p = nullptr;
        MCT_CUDA(ctx, cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
        ctx->p2p.p[r] = (P2PBuf*)p;
    }
    return MCT_OK;
}
extern "C" int mct_p2p_destroy(mct_ctx* ctx)
{
    if (!ctx || !ctx->p2p_world) return MCT_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < ctx->p2p_world; r++) {
        if (!ctx->p2p.p[r]) continue;
        if (r == ctx->p2p_rank) cudaFree(ctx->p2p.p[r]); else cudaIpcCloseMemHandle(ctx->p2p.p[r]);
    }
    if (ctx->d_p2p_ticket) cudaFree(ctx->d_p2p_ticket);
    if (ctx->h_p2p_out) cudaFreeHost(ctx->h_p2p_out);
    if (ctx->h_p2p_status) cudaFreeHost(ctx->h_p2p_status);
    memset(&ctx->p2p, 0, sizeof(ctx->p2p));
    ctx->p2p_world = 0; ctx->d_p2p_ticket = nullptr; ctx->h_p2p_out = nullptr; ctx->h_p2p_status = nullptr;
    return MCT_OK;
}
// one frame: host_out [world][n_obj][2] = the foot points of every camera (index = rank); all cameras must call it once per frame
extern "C" int mct_fuser_exchange_foot_points(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const float* h0, float* host_out, int timeout_ms)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !h0 || !host_out) return MCT_E_ARG;
    if (!ctx->p2p_world) return mct_fail(ctx, MCT_E_STATE, "peer exchange not created%s");
    for (int r = 0; r < ctx->p2p_world; r++) if (!ctx->p2p.p[r]) return mct_fail(ctx, MCT_E_STATE, "peer buffers not opened%s");
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_tmp(ctx, (size_t)n_obj * 4);
    if (rc) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp, h0, (size_t)n_obj * 4, cudaMemcpyHostToDevice, ctx->stream));
    ObjDesc h[DESC_MAX_OBJ]; const ObjDesc* d;
    for (int o = 0; o < n_obj; o++) { memset(&h[o], 0, sizeof(ObjDesc)); fill_pf_desc(&h[o], pfs[o]); }
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    const unsigned seq = ++ctx->p2p_seq;
    *ctx->h_p2p_status = 0;
    const unsigned long long tns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ull;
    if ((rc = launch_foot_points_exchange(ctx, d, n_obj, (const float*)ctx->d_tmp, ctx->p2p, ctx->p2p_world, ctx->p2p_rank, seq, ctx->d_p2p_ticket,
                                          ctx->h_p2p_out_dev, tns, ctx->h_p2p_status_dev)))
        return rc;
    if ((rc = mct_sync(ctx))) return rc;
    if (*ctx->h_p2p_status) return mct_fail(ctx, MCT_E_STATE, "peer exchange timed out: a camera did not publish its frame%s");
    memcpy(host_out, ctx->h_p2p_out, sizeof(float) * (size_t)ctx->p2p_world * n_obj * 2);
    return MCT_OK;
}

extern "C" int mct_fuser_pack_foot_points(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const float* h0, float* dev_out)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !h0 || !dev_out) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_tmp(ctx, (size_t)n_obj * 4);
    if (rc) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp, h0, (size_t)n_obj * 4, cudaMemcpyHostToDevice, ctx->stream));
    ObjDesc h[DESC_MAX_OBJ]; const ObjDesc* d;
    for (int o = 0; o < n_obj; o++) { memset(&h[o], 0, sizeof(ObjDesc)); fill_pf_desc(&h[o], pfs[o]); }
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    if ((rc = launch_foot_points(ctx, d, n_obj, ctx->d_tmp, dev_out))) return rc;
    return mct_sync(ctx);
}

// the same payload read back to the host: host_out [n_obj][2] (CameraNetwork::FuseGeometrically gathers it from every camera)
extern "C" int mct_fuser_foot_points(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const float* h0, float* host_out)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !h0 || !host_out) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_tmp(ctx, (size_t)n_obj * 4 + (size_t)n_obj * 8);
    if (rc) return rc;
    float* d_out = (float*)ctx->d_tmp + n_obj;
    MCT_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmp, h0, (size_t)n_obj * 4, cudaMemcpyHostToDevice, ctx->stream));
    ObjDesc h[DESC_MAX_OBJ]; const ObjDesc* d;
    for (int o = 0; o < n_obj; o++) { memset(&h[o], 0, sizeof(ObjDesc)); fill_pf_desc(&h[o], pfs[o]); }
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    if ((rc = launch_foot_points(ctx, d, n_obj, (const float*)ctx->d_tmp, d_out))) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(host_out, d_out, (size_t)n_obj * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

extern "C" int mct_pf_training_boxes(mct_ctx* ctx, int n_obj, mct_pf* const* pfs, const mct_train_gen* gen, int cap,
                                     int* col, int* row, float* sx, float* sy, mct_train_gen_out* out)
{
    if (!ctx || n_obj < 1 || n_obj > DESC_MAX_OBJ || !pfs || !gen || cap < 1 || !col || !row || !sx || !sy || !out) return MCT_E_ARG;
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    int max_pos = 1;
    for (int o = 0; o < n_obj; o++) {
        if (!pfs[o] || !pfs[o]->initialized) return mct_fail(ctx, MCT_E_STATE, "particle filter not initialised%s");
        if (gen[o].strategy < 0 || gen[o].strategy > 3 || gen[o].max_pos < 0) return mct_fail(ctx, MCT_E_ARG, "bad training-sample strategy%s");
        if (gen[o].max_pos > max_pos) max_pos = gen[o].max_pos;
    }
    // scratch: gen[n] | out[n] | col,row,sx,sy [n][cap]
    const size_t off_out = (size_t)n_obj * sizeof(mct_train_gen), off_box = off_out + (size_t)n_obj * sizeof(mct_train_gen_out);
    const size_t box_bytes = (size_t)n_obj * cap * 4;
    int rc = ensure_tmp(ctx, off_box + 4 * box_bytes);
    if (rc) return rc;
    unsigned char* base = (unsigned char*)ctx->d_tmp;
    mct_train_gen* d_gen = (mct_train_gen*)base;
    mct_train_gen_out* d_out = (mct_train_gen_out*)(base + off_out);
    int* d_col = (int*)(base + off_box); int* d_row = (int*)(base + off_box + box_bytes);
    float* d_sx = (float*)(base + off_box + 2 * box_bytes); float* d_sy = (float*)(base + off_box + 3 * box_bytes);
    MCT_CUDA(ctx, cudaMemcpyAsync(d_gen, gen, (size_t)n_obj * sizeof(mct_train_gen), cudaMemcpyHostToDevice, ctx->stream));
    ObjDesc h[DESC_MAX_OBJ]; const ObjDesc* d;
    for (int o = 0; o < n_obj; o++) { memset(&h[o], 0, sizeof(ObjDesc)); fill_pf_desc(&h[o], pfs[o]); }
    if ((rc = desc_publish(ctx, h, n_obj, &d))) return rc;
    if ((rc = launch_train_from_particles(ctx, d, h, n_obj, d_gen, max_pos, cap, d_col, d_row, d_sx, d_sy, d_out))) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)n_obj * sizeof(mct_train_gen_out), cudaMemcpyDeviceToHost, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(col, d_col, box_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(row, d_row, box_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(sx, d_sx, box_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(sy, d_sy, box_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}

extern "C" int mct_sample_regions(mct_ctx* ctx, int n_regions, const mct_sample_region* regions, int cap, int* col, int* row,
                                  mct_sample_region_out* out)
{
    if (!ctx || n_regions < 1 || !regions || cap < 1 || !col || !row || !out) return MCT_E_ARG;
    for (int i = 0; i < n_regions; i++) {
        const mct_sample_region& g = regions[i];
        if (g.kind < 0 || g.kind > 1 || g.max_n < 0) return mct_fail(ctx, MCT_E_ARG, "bad sample region%s (%d)", "", i);
        // the reference asserts these (SampleSet.cpp:97, :203-204)
        if (g.kind == 0 ? !(g.p[0] > g.p[1]) : !(g.p[0] > g.p[2] && g.p[1] > g.p[3]))
            return mct_fail(ctx, MCT_E_ARG, "sample region: outer extent must exceed the inner one%s (%d)", "", i);
    }
    MCT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t off_out = (size_t)n_regions * sizeof(mct_sample_region), off_box = off_out + (size_t)n_regions * sizeof(mct_sample_region_out);
    const size_t box_bytes = (size_t)n_regions * cap * 4;
    int rc = ensure_tmp(ctx, off_box + 2 * box_bytes);
    if (rc) return rc;
    unsigned char* base = (unsigned char*)ctx->d_tmp;
    mct_sample_region* d_reg = (mct_sample_region*)base;
    mct_sample_region_out* d_out = (mct_sample_region_out*)(base + off_out);
    int* d_col = (int*)(base + off_box); int* d_row = (int*)(base + off_box + box_bytes);
    MCT_CUDA(ctx, cudaMemcpyAsync(d_reg, regions, off_out, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_sample_regions(ctx, n_regions, d_reg, cap, d_col, d_row, d_out))) return rc;
    MCT_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)n_regions * sizeof(mct_sample_region_out), cudaMemcpyDeviceToHost, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(col, d_col, box_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MCT_CUDA(ctx, cudaMemcpyAsync(row, d_row, box_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return mct_sync(ctx);
}
