// k_classify.cu -- strong-classifier response H(x) = sum_{k in selected} ClassifyF_k(f_k(x))
// (MILBoostClassifier::Classify, MILBoostClassifier.cpp:339-374; MILAnyBoost identical, :262-296).
//
// One thread per box, blockIdx.y = object.  The K selected weak classifiers (Haar rect tables + the
// 8-float Gaussian state) are staged once per CTA in shared memory and read as warp-uniform
// broadcasts.  Haar features are evaluated on the fly from the integral image (the reference
// materialises all F and reads K; results are identical), colour features are read from the
// [rows][ldN] matrix produced by k_mdch / k_cch.  H accumulates in selector order in f32, exactly like
// the reference's responseList[i] += weak[k] loop.
#include "mct_internal.cuh"
#include <math_constants.h>

__global__ void __launch_bounds__(256)
k_classify(const ObjDesc* __restrict__ descs, const float* __restrict__ ii, int iip, int W, int H,
           int from_matrix, int log_ratio)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    SelEntry* se = (SelEntry*)smraw;
    const ObjDesc& d = descs[blockIdx.y];
    const int n = d.N;
    if (blockIdx.x * blockDim.x >= n) return;
    const int nsel = *d.nsel;
    const int F = d.F;
    for (int s = threadIdx.x; s < nsel; s += blockDim.x) {
        const int f = d.sel[s];
        SelEntry& e = se[s];
#pragma unroll
        for (int q = 0; q < 8; q++) e.st[q] = d.weak[(size_t)q * F + f];
        if (d.alpha) e.st[2] = d.alpha[s];                     // AdaBoost: the vote weight rides in the (unused) sigma0 slot
        if (f < d.n_haar && !from_matrix) {
            e.nrect = d.nrect[f];
            for (int k = 0; k < MCT_MAX_RECTS; k++) {
                e.r[k] = d.rects[(size_t)f * MCT_MAX_RECTS + k];
                e.w[k] = d.wts[(size_t)f * MCT_MAX_RECTS + k];
            }
            e.row = 0;
        } else {
            e.nrect = 0;
            e.row = from_matrix ? f : d.colorrow[f - d.n_haar];
        }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (d.valid && !d.valid[i]) { d.H[i] = 0.0f; return; }
    const int c = d.col[i], r = d.row[i];
    const float fx = d.sx[i], fy = d.sy[i];
    float Hs = 0.0f;
    for (int s = 0; s < nsel; s++) {
        const SelEntry& e = se[s];
        float x;
        if (e.nrect > 0) x = haar_eval_dev(ii, iip, W, H, e.r, e.w, e.nrect, c, r, fx, fy);
        else x = d.featN[(size_t)e.row * d.ldN + i];
        if (d.alpha) Hs = __fadd_rn(Hs, weak_classify_bool_dev(d.weak_type, x, e.st[0], e.st[1], e.st[4], e.st[5], e.st[6], e.st[7]) ? e.st[2] : -e.st[2]);
        else if (d.color_resp && e.nrect == 0) Hs = __fadd_rn(Hs, x);          // featN already holds the response (k_mdch_sel)
        else Hs = __fadd_rn(Hs, weak_classify_dev(d.weak_type, x, e.st[0], e.st[1], e.st[4], e.st[5], e.st[6], e.st[7]));
    }
    if (!log_ratio) Hs = sigmoid_dev(d.alpha ? __fmul_rn(2.0f, Hs) : Hs);     // AdaBoost: sigmoid(2 H) (:208)
    d.H[i] = Hs;
}

// ------------------------------------------------------------------------------------------
// k_classify_staged: the per-frame (test-set) form.  Persistent CTAs, each owning a contiguous chunk of ONE
// object's boxes (the same chunking as k_mdch_sel):
//   1. the K selected weak classifiers are staged in shared memory (as above);
//   2. the bounding region of the chunk's boxes in the INTEGRAL IMAGE is staged once into shared memory
//      (particles of an object cluster within a few sigma, so the region is ~150 KB at VGA); the 16 corner
//      gathers per Haar feature then hit shared memory (random bank conflicts, ~3 wavefronts) instead of L1 with
//      32 distinct lines per warp instruction -- the L1 wavefront rate was what bounded k_classify;
//   3. four lanes share a box: lane q evaluates selectors q, q+4, ...; the responses are then added IN SELECTOR
//      ORDER through shuffles, so H equals the reference's sequential f32 accumulation bit for bit;
//   4. log(1e-5+p1) - log(1e-5+p0) is evaluated as one f64 log of the f64 quotient (half the f64 log count;
//      identical after rounding to f32 except for sub-ulp ties).
// Falls back to global gathers (same arithmetic) when the region does not fit.
// ClassifyF with one f64 log: (float)log((1e-5 + p1) / (1e-5 + p0))
__device__ __forceinline__ float weak_classify_ratio_dev(int weak_type, float x, const float* st)
{
    if (weak_type == MCT_WEAK_PERCEPTRON) return ((double)x * (double)st[0] > (double)st[1]) ? 1.0f : -1.0f;
    const float d0 = __fsub_rn(x, st[0]), d1 = __fsub_rn(x, st[1]);
    const float a0 = __fmul_rn(__fmul_rn(d0, d0), st[6]);
    const float a1 = __fmul_rn(__fmul_rn(d1, d1), st[7]);
    const double p0 = (double)__fmul_rn(expf(a0), st[4]);
    const double p1 = (double)__fmul_rn(expf(a1), st[5]);
    return (float)log((1e-5 + p1) / (1e-5 + p0));
}

#ifdef MCT_CLS_DEBUG
// per-CTA phase stamps of the last k_classify_staged launch (tools/cls_stamps.py): [cta][8] globaltimer values
__device__ unsigned long long g_cls_dbg[256 * 8];
__device__ __forceinline__ unsigned long long cls_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define CLS_STAMP(k) do { if (threadIdx.x == 0) g_cls_dbg[((blockIdx.y * gridDim.x + blockIdx.x) & 255) * 8 + (k)] = cls_gtime(); } while (0)
extern "C" int mct_debug_cls_stamps(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_cls_dbg, sizeof(unsigned long long) * 256 * 8); }
#else
#define CLS_STAMP(k)
#endif
#define CLS_THREADS 512
#ifndef CLS_UF
#define CLS_UF 4              // selectors in flight per lane (8 measured the same)
#endif
#define CLS_SMEM (224 * 1024)
__global__ void __launch_bounds__(CLS_THREADS, 1)
k_classify_staged(const ObjDesc* __restrict__ descs, const float* __restrict__ ii, int iip, int W, int H, int log_ratio,
                  int smem_bytes)
{
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ int s_bb[4], s_ctr[3];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ ObjDesc s_desc;
    CLS_STAMP(0);
    const ObjDesc& d = load_desc(descs + blockIdx.y, &s_desc);
    const int tid = threadIdx.x, lane = tid & 31;
    const int N = d.N;
    CLS_STAMP(1);
    const int chunk = (N + gridDim.x - 1) / gridDim.x;
    const int i_begin = blockIdx.x * chunk, i_end = min(N, i_begin + chunk);
    if (i_begin >= i_end) return;
    SelEntry* se = (SelEntry*)smraw;
    const int Kcap = d.K;
    const size_t se_b = (size_t)max(Kcap, 1) * sizeof(SelEntry);               // 160 B entries: a multiple of 16
    if (tid == 0) {
        s_bb[0] = 1 << 30; s_bb[1] = 1 << 30; s_bb[2] = -(1 << 30); s_bb[3] = -(1 << 30);
        s_ctr[0] = 0; s_ctr[1] = 0; s_ctr[2] = 0;
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ---- 1. the prepared selected-classifier table arrives with ONE bulk copy while the boxes are scanned
    if (tid == 0) {
        mbar_expect_tx(&s_bar[0], (unsigned)se_b);
        bulk_g2s(se, d.seltab, (unsigned)se_b, &s_bar[0]);
    }
    const int nsel = __ldg(d.selhdr), ext_w = __ldg(d.selhdr + 1), ext_h = __ldg(d.selhdr + 2);
    // everything above reads what earlier frames' kernels left (descriptor, selector table); the boxes, the valid flags and the colour
    // responses come from the producer in front of this launch
    pdl_wait();
    pdl_trigger();
    // ---- 2. integral-image region of the chunk's valid boxes (Haar selectors only)
    if (ext_w > 0) {
        int x0 = 1 << 30, y0 = 1 << 30, x1 = -(1 << 30), y1 = -(1 << 30);
        int sxc = 0, syc = 0, nb = 0;                                          // sum of the box centres (window placement below)
        for (int i = i_begin + tid; i < i_end; i += blockDim.x) {
            // (all five loads are issued before the validity test: one round trip to L2 instead of two)
            const int vld = d.valid != nullptr ? d.valid[i] : 1;
            const int c = d.col[i], r = d.row[i];
            const float bsx = d.sx[i], bsy = d.sy[i];
            if (vld == 0) continue;
            // round(a*s) + round(b*s) <= (a+b)*s + 1
            const int ex = (int)ceilf((float)ext_w * bsx) + 2, ey = (int)ceilf((float)ext_h * bsy) + 2;
            x0 = min(x0, c); y0 = min(y0, r); x1 = max(x1, c + ex); y1 = max(y1, r + ey);
            sxc += c + (ex >> 1); syc += r + (ey >> 1); nb++;
        }
        if (nb) { atomicAdd(&s_ctr[0], sxc); atomicAdd(&s_ctr[1], syc); atomicAdd(&s_ctr[2], nb); }
        for (int o = 16; o > 0; o >>= 1) {
            x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
            x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
        if (lane == 0 && x1 >= x0) { atomicMin(&s_bb[0], x0); atomicMin(&s_bb[1], y0); atomicMax(&s_bb[2], x1); atomicMax(&s_bb[3], y1); }
    }
    __syncthreads();
    CLS_STAMP(2);
    float* region = (float*)(smraw + se_b);
    int bx0 = clampi(s_bb[0], 0, W), by0 = clampi(s_bb[1], 0, H);
    int bx1 = clampi(s_bb[2], 0, W), by1 = clampi(s_bb[3], 0, H);
    // When the bounding region of ALL boxes does not fit, a window of the same aspect around the mean box centre is staged
    // instead (the bounding box is set by a few outliers of the particle cloud); boxes inside the window evaluate from shared
    // memory, only the outliers gather from global memory.  (The launch lasts as long as its slowest CTA.)
    bool partial = false;
    {
        const size_t budget = ((size_t)smem_bytes - se_b) / 4;                // floats
        const size_t full = (size_t)(bx1 - bx0 + 12) * (size_t)(by1 - by0 + 1);
        if (ext_w > 0 && s_bb[2] >= s_bb[0] && s_ctr[2] > 0 && full > budget) {
            const float f = sqrtf((float)budget / (float)full) * 0.97f;
            const int ww = (int)((float)(bx1 - bx0 + 1) * f), wh = (int)((float)(by1 - by0 + 1) * f);
            const int mx = s_ctr[0] / s_ctr[2], my = s_ctr[1] / s_ctr[2];
            if (ww > 2 * ext_w + 8 && wh > 2 * ext_h + 8) {
                const int nx0 = clampi(mx - ww / 2, bx0, bx1 - ww + 1), ny0 = clampi(my - wh / 2, by0, by1 - wh + 1);
                bx0 = nx0; by0 = ny0; bx1 = nx0 + ww - 1; by1 = ny0 + wh - 1;
                partial = true;
            }
        }
    }
    // x origin such that the source rows are 16-byte aligned (the table has a 3-float lead, MCT_II_LEAD)
    const int ox = ((bx0 + 3) & ~3) - 3;
    const int RH = by1 - by0 + 1;
    const int RWc = (bx1 - ox + 1 + 3) & ~3;                                   // floats copied per row (multiple of 4)
    const int RWp = (RWc & 4) ? RWc : RWc + 4;                                 // pitch/4 odd: rows spread over the banks
    const bool have = ext_w > 0 && s_bb[2] >= s_bb[0];
    const bool staged = have && se_b + (size_t)RWp * RH * 4 <= (size_t)smem_bytes && (size_t)RWc * RH * 4 < (1u << 20);
    if (staged) {
        // The region arrives through 16-byte loads of ALL threads, six in flight per thread (two or three round trips to L2 for
        // ~150 KB).  One bulk copy per row (~200 UBLKCP of ~700 B per CTA) was measured at 11.7 us of mbarrier wait per launch
        // (profiles/r02a: 31 % of the kernel's stall samples): the copy engine serialises small copies.
        const int q4 = RWc >> 2, n4 = q4 * RH;
        const float* src0 = ii + (size_t)by0 * iip + ox;
        for (int g0 = tid; g0 < n4; g0 += CLS_THREADS * 6) {
            float4 v[6]; int off[6];
#pragma unroll
            for (int u = 0; u < 6; u++) {
                const int g = g0 + u * CLS_THREADS;
                off[u] = -1;
                if (g < n4) {
                    const int y = g / q4, x4 = g - y * q4;
                    v[u] = __ldg(reinterpret_cast<const float4*>(src0 + (size_t)y * iip) + x4);
                    off[u] = y * RWp + 4 * x4;
                }
            }
#pragma unroll
            for (int u = 0; u < 6; u++) if (off[u] >= 0) *reinterpret_cast<float4*>(region + off[u]) = v[u];
        }
    }
    CLS_STAMP(3);
    mbar_wait(&s_bar[0], 0);
    __syncthreads();
    CLS_STAMP(4);
    // ---- 3. four lanes per box, two selectors in flight per lane
    const int q = tid & 3;
    const unsigned gmask = 0xfu << (lane & ~3);
    for (int base = i_begin; base < i_end; base += CLS_THREADS / 4) {
        const int i = base + (tid >> 2);
        const bool live = i < i_end;
        const bool ok = live && (d.valid == nullptr || d.valid[i] != 0);
        int c = 0, r = 0; float fx = 1.0f, fy = 1.0f;
        if (ok) { c = d.col[i]; r = d.row[i]; fx = d.sx[i]; fy = d.sy[i]; }
        bool in_win = staged;
        if (partial && ok) {
            const int ex = (int)ceilf((float)ext_w * fx) + 2, ey = (int)ceilf((float)ext_h * fy) + 2;
            in_win = staged && c >= bx0 && r >= by0 && c + ex <= bx1 && r + ey <= by1;
        }
        float Hs = 0.0f;
        for (int s0 = 0; s0 < nsel; s0 += 4 * CLS_UF) {
            float v[CLS_UF];
            // the inputs of CLS_UF selectors per lane are fetched before any of them is evaluated: the colour rows come from L2
            // (one round trip per selector), so the trips of a step overlap instead of queueing behind each other's f64 log
            float x[CLS_UF]; bool on[CLS_UF];
#pragma unroll
            for (int u = 0; u < CLS_UF; u++) {
                const int s = s0 + 4 * u + q;
                on[u] = ok && s < nsel;
                x[u] = 0.0f; v[u] = 0.0f;
                if (on[u]) {
                    const SelEntry& e = se[s];
                    if (e.nrect > 0) {
                        if (in_win) x[u] = haar_eval_region(region, RWp, ox, by0, W, H, e.r, e.w, e.nrect, c, r, fx, fy);
                        else x[u] = haar_eval_dev(ii, iip, W, H, e.r, e.w, e.nrect, c, r, fx, fy);
                    } else x[u] = d.featN[(size_t)e.row * d.ldN + i];
                }
            }
#pragma unroll
            for (int u = 0; u < CLS_UF; u++) {
                if (on[u]) {
                    const SelEntry& e = se[s0 + 4 * u + q];
                    if (d.color_resp && e.nrect == 0) v[u] = x[u];             // the colour kernel left the response itself in featN
                    else v[u] = d.alpha ? (weak_classify_bool_dev(d.weak_type, x[u], e.st[0], e.st[1], e.st[4], e.st[5], e.st[6], e.st[7]) ? e.st[2] : -e.st[2])
                                   : weak_classify_ratio_dev(d.weak_type, x[u], e.st);
                }
            }
            // ordered accumulation: responseList[i] += weak[k] in selector order (MILBoostClassifier.cpp:351-361)
#pragma unroll
            for (int u = 0; u < CLS_UF; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float vj = __shfl_sync(gmask, v[u], (lane & ~3) + j);
                    if (s0 + 4 * u + j < nsel) Hs = __fadd_rn(Hs, vj);
                }
        }
        if (live && q == 0) {
            if (ok && !log_ratio) Hs = sigmoid_dev(d.alpha ? __fmul_rn(2.0f, Hs) : Hs);
            d.H[i] = ok ? Hs : 0.0f;
        }
    }
    CLS_STAMP(5);
#ifdef MCT_CLS_DEBUG
    __syncthreads();
    CLS_STAMP(6);
    if (threadIdx.x == 0) g_cls_dbg[((blockIdx.y * gridDim.x + blockIdx.x) & 255) * 8 + 7] = (unsigned long long)(staged ? 1 : 0) | ((unsigned long long)(partial ? 1 : 0) << 1) | ((unsigned long long)nsel << 8) | ((unsigned long long)(RWp * RH) << 24);
#endif
}

int launch_classify_staged(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int log_ratio)
{
    if (n_obj <= 0 || n_max <= 0) return MCT_OK;
    static size_t s_attr[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_classify_staged, s_attr, CLS_SMEM, 0)) return rc_;
    const int s_sms = ctx->sms;
    int per_obj = s_sms / n_obj;
    if (per_obj < 1) per_obj = 1;
    if (per_obj > ceil_div(n_max, 32)) per_obj = ceil_div(n_max, 32);
    dim3 g(per_obj, n_obj);
    MCT_LAUNCH_PDL(ctx, "k_classify_staged", k_classify_staged, g, dim3(CLS_THREADS), (size_t)CLS_SMEM,
                   d_desc, (const float*)(ctx->ii_base + MCT_II_LEAD), ctx->ii_pitch, ctx->W, ctx->H, log_ratio, (int)CLS_SMEM);
    return MCT_OK;
}

int launch_classify(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int K_max, int from_matrix, int log_ratio)
{
    if (n_obj <= 0 || n_max <= 0) return MCT_OK;
    size_t smem = (size_t)(K_max > 0 ? K_max : 1) * sizeof(SelEntry);
    static size_t s_attr[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_classify, s_attr, smem, 48 * 1024)) return rc_;
    dim3 g(ceil_div(n_max, 256), n_obj);
    MCT_LAUNCH(ctx, "k_classify", k_classify<<<g, 256, smem, ctx->stream>>>(d_desc, ctx->ii_base ? ctx->ii_base + MCT_II_LEAD : nullptr, ctx->ii_pitch,
                                              ctx->W, ctx->H, from_matrix, log_ratio));
    return MCT_OK;
}


// ------------------------------------------------------------------------------------------
// SimpleTracker's decision (SimpleTracker.cpp:316-323): index and value of the largest response over the search window,
// first occurrence on ties (max_idx, Public.h:256-268).  One CTA per object over d.H[0 .. N).
struct ArgmaxOut { int idx; float val; };
__global__ void __launch_bounds__(256) k_response_argmax(const ObjDesc* __restrict__ descs, ArgmaxOut* __restrict__ out)
{
    __shared__ float s_v[8];
    __shared__ int s_i[8];
    const ObjDesc& d = descs[blockIdx.x];
    float bv = -CUDART_INF_F; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < d.N; i += blockDim.x) {
        const float v = d.H[i];
        if (v > bv || (v == bv && i < bi) || bi == 0x7fffffff) { bv = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++)
            if (s_i[w] != 0x7fffffff && (bi == 0x7fffffff || s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi))) { bv = s_v[w]; bi = s_i[w]; }
        out[blockIdx.x].idx = bi == 0x7fffffff ? -1 : bi;
        out[blockIdx.x].val = bv;
    }
}
int launch_response_argmax(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, void* d_out)
{
    MCT_LAUNCH(ctx, "k_response_argmax", k_response_argmax<<<n_obj, 256, 0, ctx->stream>>>(d_desc, (ArgmaxOut*)d_out));
    return MCT_OK;
}
