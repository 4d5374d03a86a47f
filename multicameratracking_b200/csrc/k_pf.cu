// k_pf.cu -- ParticleFilter device path (ParticleFilter.cpp) + GenerateTestSampleSet
// (ParticleFilterTracker.cpp:1247-1305) + likelihood map (:541-566).
//
// Bit-exactness contract (north_star: resampling indices bit-exact given identical weights):
//   NormalizeParticleWeights (:987-1011) and the CDF (:927-933) are *sequential f32 add chains* in the
//   reference; a parallel scan reassociates and can flip an index by one ulp.  k_pf_update therefore
//   stages the weights in shared memory and lets one lane run the N-long chains in reference order
//   (4 cycles per dependent FADD, ~4 us for N = 2000), while everything elementwise (likelihood map,
//   weight products, divisions, the threshold search) runs on the whole CTA.  The N_eff test (:860-875)
//   is decided from an f64 block reduction; the serial f32/f64 chain is only run when the reduction
//   lands within the chain's rounding band of the threshold, so the decision always equals the chain's.
#include "mct_internal.cuh"
#include <math_constants.h>

#define PF_THREADS 512

#ifdef MCT_PF_DEBUG
__device__ unsigned long long g_pf_dbg[64];      // [0..31] last launch without resampling, [32..63] last launch that resampled
__shared__ unsigned long long s_pf_dbg[32];
__device__ __forceinline__ unsigned long long pf_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PF_STAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) s_pf_dbg[k] = pf_gtime(); } while (0)
#define PF_COMMIT(res) do { if (blockIdx.x == 0 && threadIdx.x == 0) for (int q_ = 0; q_ < 32; q_++) g_pf_dbg[((res) ? 32 : 0) + q_] = s_pf_dbg[q_]; } while (0)
extern "C" int mct_debug_pf_stamps(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_pf_dbg, sizeof(unsigned long long) * 64); }
#else
#define PF_STAMP(k)
#define PF_COMMIT(res)
#endif

__device__ __forceinline__ float block_reduce_minmax(float v, bool is_max, float* sh)
{
    for (int o = 16; o > 0; o >>= 1) {
        float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : fminf(v, t);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < (blockDim.x >> 5)) ? sh[lane] : (is_max ? -CUDART_INF_F : CUDART_INF_F);
        for (int o = 16; o > 0; o >>= 1) {
            float t = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_max ? fmaxf(v, t) : fminf(v, t);
        }
        if (lane == 0) sh[0] = v;
    }
    __syncthreads();
    return sh[0];
}
__device__ __forceinline__ double block_reduce_sum_d(double v, double* sh)
{
    v = warp_sum_d(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
        v = warp_sum_d(v);
        if (lane == 0) sh[0] = v;
    }
    __syncthreads();
    return sh[0];
}
// count + min + max in one pass: two barriers instead of nine
__device__ __forceinline__ void block_reduce_cnt_minmax(int& cnt, float& mn, float& mx, float* shf, int* shi)
{
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) { shi[warp] = cnt; shf[warp] = mn; shf[32 + warp] = mx; }
    __syncthreads();
    cnt = 0; mn = CUDART_INF_F; mx = -CUDART_INF_F;
    for (int q = 0; q < nw; q++) { cnt += shi[q]; mn = fminf(mn, shf[q]); mx = fmaxf(mx, shf[32 + q]); }
}
// five f64 sums in one pass
__device__ __forceinline__ void block_reduce_sum_d5(double* acc, double* sh /* [5][32] */)
{
    for (int q = 0; q < 5; q++) acc[q] = warp_sum_d(acc[q]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) for (int q = 0; q < 5; q++) sh[q * 32 + warp] = acc[q];
    __syncthreads();
    for (int q = 0; q < 5; q++) { double t = 0.0; for (int k = 0; k < nw; k++) t += sh[q * 32 + k]; acc[q] = t; }
}
__device__ __forceinline__ int block_reduce_sum_i(int v, int* sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < (blockDim.x >> 5)) ? sh[lane] : 0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[0] = v;
    }
    __syncthreads();
    return sh[0];
}

// pow(base, n) for the integer exponents of the likelihood map (20, or 1 for the re-score variant) by binary
// exponentiation in f64: <= 6 roundings of 1.1e-16 each, so after the cast to f32 it equals the correctly rounded
// pow() the reference calls, without CUDA's several-hundred-instruction generic pow.
__device__ __forceinline__ double ipow_d(double x, int n)
{
    if (n < 0) return pow(x, (double)n);
    double r = 1.0, b = x;
    while (n) { if (n & 1) r *= b; b *= b; n >>= 1; }
    return r;
}

// Correctly rounded f32 quotient through one f64 division: 53 >= 2*24 + 2 bits, so the double rounding is innocuous
// (Figueroa) and the result equals __fdiv_rn bit for bit -- but f32 denormals (the ^20 likelihood map produces many)
// are normal numbers in f64, so the division never takes the slow special-case path.
__device__ __forceinline__ float fdiv_exact(float a, float b) { return (float)((double)a / (double)b); }

// The randfloat() of ResampleParticles (ParticleFilter.cpp:940): from the call's arguments, or peeked from the object's
// device-resident multiply-with-carry stream (Public.cpp:15-23); resample_consume advances that stream (ONE thread, after a barrier
// behind every thread's peek) when the filter did resample -- the reference draws only inside ResampleParticles.
__device__ __forceinline__ float resample_draw(const ObjDesc& d, const StepArgs& sa, int obj)
{
    if (!sa.dev_rng) return sa.u01[obj];
    const unsigned long long s = d.trk->rng;
    const unsigned long long n = (s & 0xffffffffull) * 4164903690ull + (s >> 32);
    return (float)((double)(unsigned)(n & 0xffffffffull) * 2.3283064365386962890625e-10);
}
__device__ __forceinline__ void resample_consume(const ObjDesc& d, const StepArgs& sa)
{
    if (!sa.dev_rng) return;
    const unsigned long long s = d.trk->rng;
    d.trk->rng = (s & 0xffffffffull) * 4164903690ull + (s >> 32);
}
// the part of the read-out that GenerateTrainingSampleSet consumes, mirrored into device memory (PfTrk)
__device__ __forceinline__ void trk_set_avg(const ObjDesc& d, int q, float v) { if (d.trk) d.trk->avg[q] = v; }
__device__ __forceinline__ void trk_set_first(const ObjDesc& d, float x, float y, float sx, float sy)
{
    if (d.trk) { d.trk->first[0] = x; d.trk->first[1] = y; d.trk->first[2] = sx; d.trk->first[3] = sy; }
}

__global__ void __launch_bounds__(256) k_pf_predict(const ObjDesc* __restrict__ descs, const StepArgs sa)
{
    const ObjDesc& d = descs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { d.st->filter_state = MCT_PF_PREDICTED; d.st->time_index += 1; }
    if (i >= d.N) return;
    predict_one(d, i, sa.noise_base + d.noise_off);
}
int launch_pf_predict(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, const StepArgs& sa)
{
    dim3 g(ceil_div(n_max, 256), n_obj);
    MCT_LAUNCH(ctx, "k_pf_predict", k_pf_predict<<<g, 256, 0, ctx->stream>>>(d_desc, sa));
    return MCT_OK;
}

__global__ void __launch_bounds__(256) k_pf_testset(const ObjDesc* __restrict__ descs, int fw, int fh)
{
    pdl_trigger();                             // the classify kernel behind this one may become resident (it waits for these boxes)
    const ObjDesc& d = descs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && d.nbig) *d.nbig = 0;                 // (the re-score variant starts here: no predict kernel in front)
    if (i >= d.N) return;
    testset_one(d, i, fw, fh);
}
// batched per-frame form: PredictWithBrownianMotion + GenerateTestSampleSet in one pass over the particles
__global__ void __launch_bounds__(256) k_pf_predict_testset(const ObjDesc* __restrict__ descs, const StepArgs sa, int fw, int fh)
{
    __shared__ ObjDesc s_desc;
    pdl_trigger();                             // the classify kernel behind this one may become resident (it waits for these boxes)
    const ObjDesc& d = load_desc(descs + blockIdx.y, &s_desc);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        d.st->filter_state = MCT_PF_PREDICTED; d.st->time_index += 1;
        if (d.nbig) *d.nbig = 0;                       // per-frame reset of the MDCH generic-path counter
    }
    const bool in = i < d.N;
    const int any = __syncthreads_or(in && refine_valid_one(d, i));
    if (threadIdx.x == 0) atomicOr(&d.st->refine, MCT_REFINE_PENDING | (any ? MCT_REFINE_VALID : 0));
    if (!in) return;
    predict_one(d, i, sa.noise_base + d.noise_off);
    testset_one(d, i, fw, fh);
}
int launch_pf_predict_testset(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, const StepArgs& sa)
{
    dim3 g(ceil_div(n_max, 256), n_obj);
    MCT_LAUNCH(ctx, "k_pf_predict_testset", k_pf_predict_testset<<<g, 256, 0, ctx->stream>>>(d_desc, sa, ctx->W, ctx->H));
    return MCT_OK;
}
int launch_pf_testset(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max)
{
    dim3 g(ceil_div(n_max, 256), n_obj);
    MCT_LAUNCH(ctx, "k_pf_testset", k_pf_testset<<<g, 256, 0, ctx->stream>>>(d_desc, ctx->W, ctx->H));
    return MCT_OK;
}

// Sequential f32 chains in reference order, run by ONE WARP out of registers.
//   NormalizeParticleWeights (:987-1011) and the CDF (:927-933) are dependent FADD chains (4 cycles per element); any
//   load, store or register move inside the chain costs extra (measured: 6.4 cycles per element with shared-memory
//   operands, 12+ with the CDF stores).  Here lane L of warp 0 holds elements [64 L, 64 L + 64) of a 2048-element tile
//   in registers; the lanes run their 64 adds one after the other (the running value hops from lane to lane with a
//   shuffle), so the chain is pure FADD, and the CDF values stay in registers until the whole tile is done and are
//   then written out by all lanes at once.
//   Layout: element e of sw / cdf lives at index PIDX(e) = e + (e >> 6): one pad word per 64 elements makes the
//   lane-blocked register loads and stores bank-conflict free.
#define MCT_CH_PER 64
#define PIDX(e) ((e) + ((e) >> 6))
#define PLEN(n) ((n) + ((n) >> 6) + 4)
// returns the sequential sum of w[0..n); all 32 lanes of the calling warp must call it
__device__ __forceinline__ float warp_chain_sum(const float* __restrict__ w, int n, int lane)
{
    float carry = 0.0f;
    for (int t0 = 0; t0 < n; t0 += 32 * MCT_CH_PER) {
        float r[MCT_CH_PER];
        const int e0 = t0 + lane * MCT_CH_PER;
#pragma unroll
        for (int k = 0; k < MCT_CH_PER; k++) r[k] = (e0 + k < n) ? w[PIDX(e0) + k] : 0.0f;     // x + 0.0f == x: padding is inert
        for (int L = 0; L < 32; L++) {
            // branch-free: lanes other than L add +0.0f (exact identity for the non-negative running value), so the
            // warp never diverges and the selects sit in the shadow of the 4-cycle dependent adds
            const bool mine = lane == L;
            float c = carry;
#pragma unroll
            for (int k = 0; k < MCT_CH_PER; k++) c = __fadd_rn(c, mine ? r[k] : 0.0f);
            carry = __shfl_sync(0xffffffffu, c, L);
        }
    }
    return carry;
}
// cdf[0] = 0, cdf[i] = cdf[i-1] + w[i]  (ParticleFilter.cpp:927-933; w[0] excluded -- reference quirk)
__device__ __forceinline__ void warp_chain_cdf(const float* __restrict__ w, float* __restrict__ cdf, int n, int lane)
{
    float carry = 0.0f;
    for (int t0 = 0; t0 < n; t0 += 32 * MCT_CH_PER) {
        float r[MCT_CH_PER];
        const int e0 = t0 + lane * MCT_CH_PER;
#pragma unroll
        for (int k = 0; k < MCT_CH_PER; k++) r[k] = (e0 + k < n) ? w[PIDX(e0) + k] : 0.0f;
        if (e0 == 0) r[0] = 0.0f;                         // w[0] does not enter the CDF
        for (int L = 0; L < 32; L++) {
            const bool mine = lane == L;
            float c = carry;
#pragma unroll
            for (int k = 0; k < MCT_CH_PER; k++) { c = __fadd_rn(c, mine ? r[k] : 0.0f); r[k] = mine ? c : r[k]; }
            carry = __shfl_sync(0xffffffffu, c, L);
        }
#pragma unroll
        for (int k = 0; k < MCT_CH_PER; k++) if (e0 + k < n) cdf[PIDX(e0) + k] = r[k];
    }
}

// ------------------------------------------------------------------------------------------
// ResampleParticles body (ParticleFilter.cpp:920-978).  sw = weights after the first normalisation
// (shared or global), cdf = scratch of the same length.
__device__ void resample_body(const ObjDesc& d, float* sw, float* cdf, int force, float* sh_f, float u01)
{
    const int N = d.N, tid = threadIdx.x, nt = blockDim.x;
    PF_STAMP(6);
    // NormalizeParticleWeights again (:924): chain on warp 0, the divisions on all threads (tiny weights take the
    // slow denormal path of the IEEE division: keep them off the serial warp)
    if (tid < 32) { const float t2 = warp_chain_sum(sw, N, tid); if (tid == 0) sh_f[0] = t2; }
    __syncthreads();
    const float total = sh_f[0];
#pragma unroll 4
    for (int i = tid; i < N; i += nt) sw[PIDX(i)] = (total != 0) ? fdiv_exact(sw[PIDX(i)], total) : 1e-2f;
    __syncthreads();
    // CDF (:927-933): warp 0, out of registers
    if (tid < 32) warp_chain_cdf(sw, cdf, N, tid);
    __syncthreads();
    PF_STAMP(7);
    const float invN = __fdiv_rn(1.0f, (float)N);
    const float seed = __fmul_rn(invN, u01);
    int top = 1;
    while ((top << 1) <= N) top <<= 1;
    for (int i0 = tid; i0 < N; i0 += 4 * nt) {
        // idx = min(N-1, #{j : cdf[j] < thr})  == the reference's monotone walk; four branch-free
        // lower-bound searches run in lock-step so that their shared-memory latencies overlap
        float thr[4]; int lo[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { thr[k] = __fadd_rn(seed, __fmul_rn(invN, (float)(i0 + k * nt))); lo[k] = 0; }
        for (int step = top; step > 0; step >>= 1) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int m = lo[k] + step;
                if (m <= N && cdf[PIDX(m - 1)] < thr[k]) lo[k] = m;
            }
        }
        int idx[4]; float gx[4], gy[4], gsx[4], gsy[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            idx[k] = lo[k] < N - 1 ? lo[k] : N - 1;
            gx[k] = d.cx[idx[k]]; gy[k] = d.cy[idx[k]]; gsx[k] = d.sx[idx[k]]; gsy[k] = d.sy[idx[k]];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k * nt;
            if (i < N) {
                d.residx[i] = idx[k];
                d.rcx[i] = gx[k]; d.rcy[i] = gy[k]; d.rsx[i] = gsx[k]; d.rsy[i] = gsy[k];
                d.rw[i] = 1.0f;
            }
        }
    }
    __syncthreads();
    PF_STAMP(8);
    if (force) {
        for (int i0 = tid; i0 < N; i0 += 4 * nt) {
            float gx[4], gy[4], gsx[4], gsy[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = min(i0 + k * nt, N - 1);
                gx[k] = d.rcx[i]; gy[k] = d.rcy[i]; gsx[k] = d.rsx[i]; gsy[k] = d.rsy[i];
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = i0 + k * nt;
                if (i < N) { d.cx[i] = gx[k]; d.cy[i] = gy[k]; d.sx[i] = gsx[k]; d.sy[i] = gsy[k]; d.w[i] = 1.0f; }
            }
        }
        if (tid == 0) d.st->filter_state = MCT_PF_RESAMPLED;
    } else {
        for (int i = tid; i < N; i += nt) d.w[i] = sw[PIDX(i)];
    }
}

__device__ void summary_body(const ObjDesc& d, float* scratch, int summ_off = 0);

// mode 0: d.H holds the classifier response; build the likelihood map ((H-min)/range)^exponent over the
//         valid particles (ParticleFilterTracker.cpp:541-566) and use it as prob.
// mode 1: d.H already holds prob, dense by particle (mask d.valid).
// Then UpdateAllParticlesWeight (:680-717): w <- p*w (or p), scale clamp, Normalize, ResampleIfRequired.
__global__ void __launch_bounds__(PF_THREADS) k_pf_update(const ObjDesc* __restrict__ descs, const StepArgs sa, int mode, int smem_floats,
                                                          int with_summary)
{
    extern __shared__ __align__(16) float smf[];
    __shared__ float sh_f[64];
    __shared__ double sh_d[32];
    __shared__ int sh_i[32];
    __shared__ int s_decide;
    __shared__ ObjDesc s_desc;
    const ObjDesc& d = load_desc(descs + blockIdx.x, &s_desc);
    const int N = d.N, tid = threadIdx.x, nt = blockDim.x;
    const int N4 = (PLEN(N) + 3) & ~3;
    float* sw = (2 * N4 <= smem_floats) ? smf : d.scratch;
    float* cdf = (2 * N4 <= smem_floats) ? smf + N4 : d.scratch + PLEN(d.ld);

    PF_STAMP(0);
    const int refine = d.st->refine;          // consumed after the first reduction: the load overlaps it
    // 0. valid count, min / max of H over the test set
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    int cnt = 0;
    // (all per-particle loops below fetch four particles per thread before using any: the loads are L2 round trips)
    for (int i0 = tid; i0 < N; i0 += 4 * nt) {
        int vv[4]; float hh[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const int i = i0 + k * nt; vv[k] = (i < N) ? d.valid[i] : 0; hh[k] = (i < N) ? d.H[i] : 0.0f; }
#pragma unroll
        for (int k = 0; k < 4; k++) if (vv[k]) { mn = fminf(mn, hh[k]); mx = fmaxf(mx, hh[k]); cnt++; }
    }
    block_reduce_cnt_minmax(cnt, mn, mx, sh_f, sh_i);
    if (refine == MCT_REFINE_PENDING)         // CheckForParticleRefinement of this frame's predict found no particle in frame
        for (int i = tid; i < N; i += nt) force_refine_one(d, i);
    __syncthreads();
    if (tid == 0) {
        d.st->refine = 0;
        d.st->n_valid = cnt; d.st->resampled = 0;
        if (mode == 0) {
            d.st->max_response = cnt > 0 ? mx : 0.0f;
            if (d.scored) atomicAdd(d.scored, (unsigned long long)cnt);
        }
    }
    if (mode == 0 && cnt == 0) {              // reference: ASSERT abort (ParticleFilterTracker.cpp:513)
        __syncthreads();
        if (with_summary) summary_body(d, nullptr, sa.summ_off);    // the host sees n_valid == 0 -> MCT_E_EMPTY
        return;
    }
    const double dmn = (double)mn, range = (double)mx - (double)mn;
    PF_STAMP(1);
    // 1. weight update
    for (int i0 = tid; i0 < N; i0 += 4 * nt) {
        int vv[4]; float hh[4], ww[4], xs[4], ys[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k * nt;
            const bool in = i < N;
            vv[k] = in ? d.valid[i] : 0; hh[k] = in ? d.H[i] : 0.0f; ww[k] = in ? d.w[i] : 0.0f;
            xs[k] = in ? d.sx[i] : 1.0f; ys[k] = in ? d.sy[i] : 1.0f;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k * nt;
            if (i >= N) continue;
            float w = ww[k];
            if (vv[k]) {
                float p;
                if (mode == 0) {
                    // exponent 0 selects the appearance fuser's map pow(sigmoid(H), 1) (AppearanceBasedInformationFuser.cpp:108-112)
                    if (d.exponent == 0) p = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-hh[k])));
                    else p = (range != 0) ? (float)ipow_d(((double)hh[k] - dmn) / range, d.exponent) : 1.0f;
                }
                else p = hh[k];
                w = d.force_reset ? p : __fmul_rn(p, w);
                float sx = xs[k], sy = ys[k];
                scale_ratio_clamp(sx, sy);
                d.sx[i] = sx; d.sy[i] = sy;
            }
            sw[PIDX(i)] = w;
        }
    }
    __syncthreads();
    PF_STAMP(2);
    // 2. NormalizeParticleWeights: sequential f32 total (:991-995), warp 0 out of registers
    if (tid < 32) { const float t1 = warp_chain_sum(sw, N, tid); if (tid == 0) sh_f[0] = t1; }
    __syncthreads();
    PF_STAMP(3);
    const float total = sh_f[0];
    __syncthreads();
    double s2 = 0.0;
#pragma unroll 4
    for (int i = tid; i < N; i += nt) {
        float w = (total != 0) ? fdiv_exact(sw[PIDX(i)], total) : 1e-2f;
        sw[PIDX(i)] = w;
        s2 += (double)w * (double)w;
    }
    s2 = block_reduce_sum_d(s2, sh_d);
    PF_STAMP(4);
    // 3. ResampleIfRequired, WEIGHT_BASED (:845-910)
    if (tid == 0) {
        int decide = 0;
        float neff_inv = (float)s2;
        if (d.resample && s2 != 0.0) {
            const double thr = (double)(float)((double)N * 0.01);
            const double m = (double)N * 1.2e-7 + 1e-6;           // rounding band of the sequential chain
            const double hi = 1.0 / (s2 * (1.0 - m)), lo = 1.0 / (s2 * (1.0 + m));
            if (hi < thr) decide = 1;
            else if (lo >= thr) decide = 0;
            else {
                float e = 0.0f;                                    // exact reference chain
                for (int i = 0; i < N; i++) e = (float)((double)e + (double)sw[PIDX(i)] * (double)sw[PIDX(i)]);
                neff_inv = e;
                decide = (e != 0) && (__fdiv_rn(1.0f, e) < (float)thr);
            }
        }
        d.st->neff_inv = neff_inv;
        s_decide = decide;
    }
    __syncthreads();
    PF_STAMP(5);
    if (s_decide) {
        resample_body(d, sw, cdf, 1, sh_f, resample_draw(d, sa, blockIdx.x));
        if (tid == 0) { d.st->resampled = 1; resample_consume(d, sa); }
    } else {
        for (int i = tid; i < N; i += nt) d.w[i] = sw[PIDX(i)];
    }
    PF_STAMP(10);
    if (with_summary) {
        __syncthreads();                       // the read-out reads state / weights written by other threads above
        summary_body(d, (2 * N4 <= smem_floats) ? smf : nullptr, sa.summ_off);   // sw / cdf are dead: reuse them as scratch (2 * N4 floats)
    }
    PF_STAMP(15);
    PF_COMMIT(s_decide);
}

// ------------------------------------------------------------------------------------------
// k_pf_update_fast: the same step for N <= 4 * PF_THREADS particles with the particle matrix staged ONCE.
//   * every thread owns particles tid + k * PF_THREADS (k < 4); the five state columns are read once (one round trip to
//     L2) into shared memory, every later pass (weight update, both normalisations, the resampler's gather, the copy back
//     into the state, the read-out) works out of shared memory and registers;
//   * the three ordered f32 chains run on warp 0 with CH elements per lane, CH chosen so that 32 * CH just covers N
//     (a chain over N = 1000 no longer walks a 2048-element tile);
//   * the resampled rows are written to the resampled matrix AND back into the state from the registers that did the gather;
//   * the read-out is one pass + ONE block reduction (average sums, run statistics / arg-max together).
// Results are bit-identical to k_pf_update (same expressions, same chain order); k_pf_update stays for larger N.
struct PfRed { double acc[5]; float mx; int a, b; };      // a / b: ties, runs (RESAMPLED) or first / last arg-max (PREDICTED)

template <int CH>
__device__ __forceinline__ float warp_chain_sum_t(const float* __restrict__ w, int n, int lane)
{
    float r[CH];
    const int e0 = lane * CH;
#pragma unroll
    for (int k = 0; k < CH; k++) r[k] = (e0 + k < n) ? w[PIDX(e0) + k] : 0.0f;
    float carry = 0.0f;
    const int nl = (n + CH - 1) / CH;                         // lanes that hold elements
    for (int L = 0; L < nl; L++) {
        const bool mine = lane == L;
        float c = carry;
#pragma unroll
        for (int k = 0; k < CH; k++) c = __fadd_rn(c, mine ? r[k] : 0.0f);
        carry = __shfl_sync(0xffffffffu, c, L);
    }
    return carry;
}
template <int CH>
__device__ __forceinline__ void warp_chain_cdf_t(const float* __restrict__ w, float* __restrict__ cdf, int n, int lane)
{
    float r[CH];
    const int e0 = lane * CH;
#pragma unroll
    for (int k = 0; k < CH; k++) r[k] = (e0 + k < n) ? w[PIDX(e0) + k] : 0.0f;
    if (e0 == 0) r[0] = 0.0f;                                 // w[0] does not enter the CDF (reference quirk, :927-933)
    float carry = 0.0f;
    const int nl = (n + CH - 1) / CH;
    for (int L = 0; L < nl; L++) {
        // the owning lane turns its block into prefix sums in place (r[k] += r[k-1]: one dependent FADD per element, nothing
        // else in the loop body; tools/chain_lab.cu: 6.3 cycles per element against 7.9 for the select form)
        float c = carry;
        if (lane == L) {
            r[0] = __fadd_rn(c, r[0]);
#pragma unroll
            for (int k = 1; k < CH; k++) r[k] = __fadd_rn(r[k - 1], r[k]);
            c = r[CH - 1];
        }
        carry = __shfl_sync(0xffffffffu, c, L);
    }
#pragma unroll
    for (int k = 0; k < CH; k++) if (e0 + k < n) cdf[PIDX(e0) + k] = r[k];
}

template <int CH>
__global__ void __launch_bounds__(PF_THREADS, 1) k_pf_update_fast(const ObjDesc* __restrict__ descs, const StepArgs sa, int mode, int with_summary)
{
    extern __shared__ __align__(16) float smf[];
    __shared__ float sh_f[64];
    __shared__ double sh_d[32];
    __shared__ int sh_i[32];
    __shared__ int s_decide;
    __shared__ PfRed s_red[PF_THREADS / 32];
    __shared__ ObjDesc s_desc;
    const ObjDesc& d = load_desc(descs + blockIdx.x, &s_desc);
    const int N = d.N, tid = threadIdx.x;
    constexpr int nt = PF_THREADS;
    const int NP = (PLEN(N) + 3) & ~3, N4 = (N + 3) & ~3;
    float* sw = smf; float* cdf = smf + NP;
    float* ccx = cdf + NP; float* ccy = ccx + N4; float* csx = ccy + N4; float* csy = csx + N4;
    pdl_wait();                                // (the descriptor is older than the frame; everything below reads the producer's output)
    pdl_trigger();

    PF_STAMP(0);
    const int refine = d.st->refine;
    // 0. the particle matrix, once
    int vv[4]; float hh[4], ww[4], xs[4], ys[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = tid + k * nt;
        const bool in = i < N;
        vv[k] = in ? d.valid[i] : 0; hh[k] = in ? d.H[i] : 0.0f; ww[k] = in ? d.w[i] : 0.0f;
        xs[k] = in ? d.sx[i] : 1.0f; ys[k] = in ? d.sy[i] : 1.0f;
        const float x = in ? d.cx[i] : 0.0f, y = in ? d.cy[i] : 0.0f;
        if (in) { ccx[i] = x; ccy[i] = y; }
    }
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) if (vv[k]) { mn = fminf(mn, hh[k]); mx = fmaxf(mx, hh[k]); cnt++; }
    block_reduce_cnt_minmax(cnt, mn, mx, sh_f, sh_i);
    if (refine == MCT_REFINE_PENDING)
        for (int i = tid; i < N; i += nt) force_refine_one(d, i);
    const float max_response = cnt > 0 ? mx : 0.0f;
    if (tid == 0) {
        d.st->refine = 0;
        d.st->n_valid = cnt; d.st->resampled = 0;
        if (mode == 0) {
            d.st->max_response = max_response;
            if (d.scored) atomicAdd(d.scored, (unsigned long long)cnt);
        }
    }
    if (mode == 0 && cnt == 0) {
        __syncthreads();
        if (with_summary) summary_body(d, nullptr, sa.summ_off);
        return;
    }
    const double dmn = (double)mn, range = (double)mx - (double)mn;
    PF_STAMP(1);
    // 1. weight update
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = tid + k * nt;
        if (i >= N) continue;
        float w = ww[k];
        float sx = xs[k], sy = ys[k];
        if (vv[k]) {
            float p;
            if (mode == 0) {
                if (d.exponent == 0) p = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-hh[k])));
                else p = (range != 0) ? (float)ipow_d(((double)hh[k] - dmn) / range, d.exponent) : 1.0f;
            }
            else p = hh[k];
            w = d.force_reset ? p : __fmul_rn(p, w);
            scale_ratio_clamp(sx, sy);
            d.sx[i] = sx; d.sy[i] = sy;
        }
        csx[i] = sx; csy[i] = sy;
        sw[PIDX(i)] = w;
        ww[k] = w;
    }
    __syncthreads();
    PF_STAMP(2);
    // 2. NormalizeParticleWeights (:987-1011)
    if (tid < 32) { const float t1 = warp_chain_sum_t<CH>(sw, N, tid); if (tid == 0) sh_f[0] = t1; }
    __syncthreads();
    PF_STAMP(3);
    const float total = sh_f[0];
    double s2 = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = tid + k * nt;
        if (i >= N) continue;
        const float w = (total != 0) ? fdiv_exact(ww[k], total) : 1e-2f;
        ww[k] = w;
        sw[PIDX(i)] = w;
        s2 += (double)w * (double)w;
    }
    s2 = block_reduce_sum_d(s2, sh_d);
    PF_STAMP(4);
    // 3. ResampleIfRequired, WEIGHT_BASED (:845-910)
    float neff_inv = (float)s2;
    if (tid == 0) {
        int decide = 0;
        if (d.resample && s2 != 0.0) {
            const double thr = (double)(float)((double)N * 0.01);
            const double m = (double)N * 1.2e-7 + 1e-6;
            const double hi = 1.0 / (s2 * (1.0 - m)), lo = 1.0 / (s2 * (1.0 + m));
            if (hi < thr) decide = 1;
            else if (lo >= thr) decide = 0;
            else {
                float e = 0.0f;
                for (int i = 0; i < N; i++) e = (float)((double)e + (double)sw[PIDX(i)] * (double)sw[PIDX(i)]);
                neff_inv = e;
                decide = (e != 0) && (__fdiv_rn(1.0f, e) < (float)thr);
            }
        }
        d.st->neff_inv = neff_inv;
        s_decide = decide;
    }
    __syncthreads();
    PF_STAMP(5);
    const int decide = s_decide;
    const int fs_in = d.st->filter_state;
    float gx[4], gy[4], gsx[4], gsy[4];
    int* sidx = reinterpret_cast<int*>(sw);              // resample index list (sw is dead once the CDF exists)
    if (decide) {
        // ResampleParticles(true) (:920-978)
        PF_STAMP(6);
        if (tid < 32) { const float t2 = warp_chain_sum_t<CH>(sw, N, tid); if (tid == 0) sh_f[1] = t2; }
        __syncthreads();
        PF_STAMP(16);
        const float total2 = sh_f[1];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = tid + k * nt;
            if (i < N) sw[PIDX(i)] = (total2 != 0) ? fdiv_exact(ww[k], total2) : 1e-2f;
        }
        __syncthreads();
        PF_STAMP(17);
        if (tid < 32) warp_chain_cdf_t<CH>(sw, cdf, N, tid);
        __syncthreads();
        PF_STAMP(7);
        const float invN = __fdiv_rn(1.0f, (float)N);
        const float seed = __fmul_rn(invN, resample_draw(d, sa, blockIdx.x));
        int top = 1;
        while ((top << 1) <= N) top <<= 1;
        float thr[4]; int lo[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { thr[k] = __fadd_rn(seed, __fmul_rn(invN, (float)(tid + k * nt))); lo[k] = 0; }
        for (int step = top; step > 0; step >>= 1) {
            // branch-free: the four loads of a step are independent and in flight together
            int m[4]; float v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { m[k] = lo[k] + step; v[k] = cdf[PIDX(min(m[k], N) - 1)]; }
#pragma unroll
            for (int k = 0; k < 4; k++) lo[k] = (m[k] <= N && v[k] < thr[k]) ? m[k] : lo[k];
        }
        PF_STAMP(18);
        int idx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            idx[k] = lo[k] < N - 1 ? lo[k] : N - 1;
            gx[k] = ccx[idx[k]]; gy[k] = ccy[idx[k]]; gsx[k] = csx[idx[k]]; gsy[k] = csy[idx[k]];
        }
        __syncthreads();                                  // every search has read the CDF / sw region: sidx may overwrite it
        PF_STAMP(19);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = tid + k * nt;
            if (i < N) {
                sidx[i] = idx[k];
                d.residx[i] = idx[k];
                d.rcx[i] = gx[k]; d.rcy[i] = gy[k]; d.rsx[i] = gsx[k]; d.rsy[i] = gsy[k]; d.rw[i] = 1.0f;
                d.cx[i] = gx[k]; d.cy[i] = gy[k]; d.sx[i] = gsx[k]; d.sy[i] = gsy[k]; d.w[i] = 1.0f;
            }
        }
        if (tid == 0) { d.st->filter_state = MCT_PF_RESAMPLED; d.st->resampled = 1; resample_consume(d, sa); }
        PF_STAMP(8);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) { const int i = tid + k * nt; if (i < N) d.w[i] = ww[k]; }
    }
    PF_STAMP(10);
    if (!with_summary) { PF_COMMIT(decide); return; }
    const int fs = decide ? MCT_PF_RESAMPLED : fs_in;
    if (!decide && fs != MCT_PF_PREDICTED) {
        // (re-score of an already resampled filter, or a filter that was never predicted: the generic read-out)
        __syncthreads();
        summary_body(d, smf, sa.summ_off);
        PF_COMMIT(decide);
        return;
    }
    // ---- read-out: one pass, one reduction
    PfRed me;
#pragma unroll
    for (int q = 0; q < 5; q++) me.acc[q] = 0.0;
    me.mx = -CUDART_INF_F; me.a = decide ? 0 : N; me.b = decide ? 0 : -1;
    float* rwf = cdf;                                     // run weight per run start, -1 elsewhere (RESAMPLED)
    if (decide) {
        __syncthreads();                                  // sidx complete
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = tid + k * nt;
            if (i >= N) continue;
            // every weight is exactly 1.0f: x * 1.0f == x and an in-order f32 sum of L ones is exactly (float)L
            me.acc[0] += (double)gx[k]; me.acc[1] += (double)gy[k]; me.acc[2] += (double)gsx[k]; me.acc[3] += (double)gsy[k]; me.acc[4] += 1.0;
            const int v = sidx[i];
            const bool start = (i == 0) || (v != sidx[i - 1]);
            float r = -1.0f;
            if (start) {
                int lo2 = i + 1, hi2 = N;                 // first j > i with sidx[j] != v
                while (lo2 < hi2) { const int mid = (lo2 + hi2) >> 1; if (sidx[mid] == v) lo2 = mid + 1; else hi2 = mid; }
                r = (float)(lo2 - i);
                me.b++;
                if (r > me.mx) { me.mx = r; me.a = 1; } else if (r == me.mx) me.a++;
            }
            rwf[i] = r;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = tid + k * nt;
            if (i >= N) continue;
            const float w = ww[k];
            me.acc[0] += (double)__fmul_rn(ccx[i], w); me.acc[1] += (double)__fmul_rn(ccy[i], w);
            me.acc[2] += (double)__fmul_rn(csx[i], w); me.acc[3] += (double)__fmul_rn(csy[i], w);
            me.acc[4] += (double)w;
            if (w > me.mx) { me.mx = w; me.a = i; me.b = i; } else if (w == me.mx) { me.a = min(me.a, i); me.b = max(me.b, i); }
        }
    }
    PF_STAMP(11);
    auto merge = [&](PfRed& x, const PfRed& y) {
#pragma unroll
        for (int q = 0; q < 5; q++) x.acc[q] += y.acc[q];
        if (decide) {
            x.b += y.b;
            if (y.mx > x.mx) { x.mx = y.mx; x.a = y.a; } else if (y.mx == x.mx) x.a += y.a;
        } else {
            if (y.mx > x.mx) { x.mx = y.mx; x.a = y.a; x.b = y.b; }
            else if (y.mx == x.mx) { x.a = min(x.a, y.a); x.b = max(x.b, y.b); }
        }
    };
    auto shfl = [&](const PfRed& x, int o) {
        PfRed y;
#pragma unroll
        for (int q = 0; q < 5; q++) y.acc[q] = __shfl_xor_sync(0xffffffffu, x.acc[q], o);
        y.mx = __shfl_xor_sync(0xffffffffu, x.mx, o); y.a = __shfl_xor_sync(0xffffffffu, x.a, o); y.b = __shfl_xor_sync(0xffffffffu, x.b, o);
        return y;
    };
    // (the f64 sums are combined in a fixed tree: lanes by xor-shuffle, then the 16 warp partials in warp order)
    for (int o = 16; o > 0; o >>= 1) { const PfRed y = shfl(me, o); merge(me, y); }
    PF_STAMP(12);
    if ((tid & 31) == 0) s_red[tid >> 5] = me;
    __syncthreads();
    PF_STAMP(13);
    if (tid == 0) {
        PfRed t = s_red[0];
        for (int q = 1; q < nt / 32; q++) merge(t, s_red[q]);
        // (the record is assembled in registers and leaves as five 16-byte stores: the destination is mapped host memory, every
        //  store a posted write over PCIe)
        __align__(16) mct_pf_summary rec;
        mct_pf_summary* out = &rec;
        for (int q = 0; q < 4; q++) { const float v = (float)(t.acc[q] / t.acc[4]); out->average[q] = v; trk_set_avg(d, q, v); }
        out->n_valid = cnt; out->resampled = decide; out->filter_state = fs;
        out->neff_inv = neff_inv; out->max_response = (mode == 0) ? max_response : d.st->max_response;
        if (decide) {
            // row (ties-1) of the ordered-unique matrix = state of the (ties-1)-th run in index order (:799-806 pairing quirk)
            const int want = t.a - 1;
            int ord = -1, b = 0;
            if (want > 0) for (int i = 0; i < N; i++) if (rwf[i] >= 0.0f) { ord++; if (ord == want) { b = i; break; } }
            const int s0 = sidx[0], sb = sidx[b];
            out->ordered_first[0] = ccx[s0]; out->ordered_first[1] = ccy[s0]; out->ordered_first[2] = csx[s0];
            out->ordered_first[3] = csy[s0]; out->ordered_first[4] = t.mx;
            trk_set_first(d, ccx[s0], ccy[s0], csx[s0], csy[s0]);
            out->highest_close[0] = ccx[sb]; out->highest_close[1] = ccy[sb]; out->highest_close[2] = csx[sb];
            out->highest_close[3] = csy[sb]; out->highest_close[4] = t.mx;
            out->n_unique = t.b;
        } else {
            const int a = t.a < N ? t.a : 0, b = t.b >= 0 ? t.b : 0;
            out->ordered_first[0] = ccx[a]; out->ordered_first[1] = ccy[a]; out->ordered_first[2] = csx[a];
            out->ordered_first[3] = csy[a]; out->ordered_first[4] = sw[PIDX(a)];
            trk_set_first(d, ccx[a], ccy[a], csx[a], csy[a]);
            out->highest_close[0] = ccx[b]; out->highest_close[1] = ccy[b]; out->highest_close[2] = csx[b];
            out->highest_close[3] = csy[b]; out->highest_close[4] = sw[PIDX(b)];
            out->n_unique = N;
        }
        static_assert(sizeof(mct_pf_summary) == 80, "read-out record: five 16-byte words");
        uint4* dst = reinterpret_cast<uint4*>(d.summ + sa.summ_off);
        const uint4* src = reinterpret_cast<const uint4*>(&rec);
#pragma unroll
        for (int q = 0; q < 5; q++) dst[q] = src[q];
    }
    PF_STAMP(15);
    PF_COMMIT(decide);
}

static int pf_update_smem(mct_ctx* ctx, int n_max, int* smem_floats, size_t* smem_bytes)
{
    // both chains' operands in shared memory when 2*N floats fit in 200 KB, else global scratch
    size_t want = (size_t)2 * round_up(PLEN(n_max), 4) * sizeof(float);
    if (want > 200 * 1024) want = 0;
    *smem_bytes = want;
    *smem_floats = (int)(want / sizeof(float));
    (void)ctx;
    return MCT_OK;
}

int launch_pf_update(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int use_H, const StepArgs& sa, int with_summary)
{
    if (n_max <= 4 * PF_THREADS && !getenv("MCT_PF_GENERIC")) {
        const size_t fb = ((size_t)2 * round_up(PLEN(n_max), 4) + (size_t)4 * round_up(n_max, 4)) * sizeof(float);
        const int mode = use_H ? 0 : 1;
        if (n_max <= 32 * 16) {
            static size_t s_a16[MCT_MAX_DEV];
            if (int rc_ = mct_smem_attr(ctx, k_pf_update_fast<16>, s_a16, fb, 32 * 1024)) return rc_;
            MCT_LAUNCH_PDL(ctx, "k_pf_update", k_pf_update_fast<16>, dim3(n_obj), dim3(PF_THREADS), fb, d_desc, sa, mode, with_summary);
        } else if (n_max <= 32 * 32) {
            static size_t s_a32[MCT_MAX_DEV];
            if (int rc_ = mct_smem_attr(ctx, k_pf_update_fast<32>, s_a32, fb, 32 * 1024)) return rc_;
            MCT_LAUNCH_PDL(ctx, "k_pf_update", k_pf_update_fast<32>, dim3(n_obj), dim3(PF_THREADS), fb, d_desc, sa, mode, with_summary);
        } else {
            static size_t s_a64[MCT_MAX_DEV];
            if (int rc_ = mct_smem_attr(ctx, k_pf_update_fast<64>, s_a64, fb, 32 * 1024)) return rc_;
            MCT_LAUNCH_PDL(ctx, "k_pf_update", k_pf_update_fast<64>, dim3(n_obj), dim3(PF_THREADS), fb, d_desc, sa, mode, with_summary);
        }
        return MCT_OK;
    }
    int sf; size_t sb;
    pf_update_smem(ctx, n_max, &sf, &sb);
    static size_t s_attr[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_pf_update, s_attr, sb, 32 * 1024)) return rc_;
    const int nt = PF_THREADS;
    MCT_LAUNCH(ctx, "k_pf_update", k_pf_update<<<n_obj, nt, sb, ctx->stream>>>(d_desc, sa, use_H ? 0 : 1, sf, with_summary));
    return MCT_OK;
}

// ResampleParticles(force) on its own (ParticleFilter.cpp:920-978)
__global__ void __launch_bounds__(PF_THREADS) k_pf_resample(const ObjDesc* __restrict__ descs, const StepArgs sa, int force, int smem_floats)
{
    extern __shared__ __align__(16) float smf[];
    __shared__ float sh_f[32];
    __shared__ ObjDesc s_desc;
    const ObjDesc& d = load_desc(descs + blockIdx.x, &s_desc);
    const int N = d.N, N4 = (PLEN(N) + 3) & ~3;
    float* sw = (2 * N4 <= smem_floats) ? smf : d.scratch;
    float* cdf = (2 * N4 <= smem_floats) ? smf + N4 : d.scratch + PLEN(d.ld);
    for (int i = threadIdx.x; i < N; i += blockDim.x) sw[PIDX(i)] = d.w[i];
    __syncthreads();
    resample_body(d, sw, cdf, force, sh_f, sa.u01[blockIdx.x]);
}
int launch_pf_resample_only(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int force, const StepArgs& sa)
{
    int sf; size_t sb;
    pf_update_smem(ctx, n_max, &sf, &sb);
    static size_t s_attr[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_pf_resample, s_attr, sb, 32 * 1024)) return rc_;
    MCT_LAUNCH(ctx, "k_pf_resample", k_pf_resample<<<n_obj, PF_THREADS, sb, ctx->stream>>>(d_desc, sa, force, sf));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// UpdateAllParticlesWeight takes a *compacted* prob list (ParticleFilter.cpp:696-704): the p2-th entry
// belongs to the p2-th particle whose weight is non-zero.  Expand it to a dense array + mask.
__global__ void __launch_bounds__(PF_THREADS) k_pf_expand_prob(const ObjDesc* __restrict__ descs,
                                                               const float* __restrict__ prob, int n_prob, int only_nonzero)
{
    __shared__ int sh_i[32];
    __shared__ int s_base;
    const ObjDesc& d = descs[0];
    const int N = d.N, tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < N; i0 += nt) {
        const int i = i0 + tid;
        int nz = 0;
        if (i < N) nz = only_nonzero ? (d.w[i] != 0) : 1;
        // block exclusive scan of nz
        int v = nz;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) sh_i[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int t = sh_i[lane];
            int s = t;
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += u; }
            sh_i[lane] = s - t;        // exclusive warp offsets
        }
        __syncthreads();
        const int excl = s_base + sh_i[warp] + v - nz;
        if (i < N) {
            const int ok = nz && excl < n_prob;
            d.valid[i] = ok;
            d.H[i] = ok ? prob[excl] : 0.0f;
        }
        __syncthreads();
        if (tid == nt - 1) s_base = excl + nz;
        __syncthreads();
    }
}
int launch_pf_expand_prob(mct_ctx* ctx, const ObjDesc* d_desc, const float* d_prob_compact, int n_prob, int only_nonzero)
{
    MCT_LAUNCH(ctx, "k_pf_expand_prob", k_pf_expand_prob<<<1, PF_THREADS, 0, ctx->stream>>>(d_desc, d_prob_compact, n_prob, only_nonzero));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// Read-out (StoreObjectState / GenerateTrainingSampleSet inputs):
//   average        GetAverageofAllParticles (:601-633)  (f64 tree instead of the f32 chain; 1e-6 rel)
//   ordered_first  GetOrderedUniqueParticles(0): FindOrderedUniqueParticles (:728-815)
//   highest_close  GetHighestOrderedUniqueParticleCloseToTheGivenState (:557-593): closestDistance is
//                  never updated, so the *last* row tied with the top weight wins.
// PREDICTED: rows = particles sorted by weight (descending, ties by index).
// RESAMPLED: the resample index list is non-decreasing, so sorting it is the identity; unique runs are
//            contiguous; run weights are summed in order; the reference then pairs the states in *index*
//            order with the weights in *sorted* order (:799-806) -- kept.
// scratch: 2 * round_up(N, 4) floats of shared memory (or NULL): the resample index list and the run weights are
// staged there so that the per-run binary searches and tie scans do not chase L2 latency.
__device__ void summary_body(const ObjDesc& d, float* scratch, int summ_off)
{
    __shared__ float sh_f[32];
    __shared__ double sh_d5[5 * 32];
    __shared__ int sh_i[32];
    __shared__ int s_first, s_last;
    const int N = d.N, tid = threadIdx.x, nt = blockDim.x;
    mct_pf_summary* out = d.summ + summ_off;
    const int fs = d.st->filter_state;

    double acc[5] = {0, 0, 0, 0, 0};
    for (int i0 = tid; i0 < N; i0 += 4 * nt) {
        float w[4], x[4], y[4], sx[4], sy[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k * nt;
            const bool in = i < N;
            w[k] = in ? d.w[i] : 0.0f; x[k] = in ? d.cx[i] : 0.0f; y[k] = in ? d.cy[i] : 0.0f;
            sx[k] = in ? d.sx[i] : 0.0f; sy[k] = in ? d.sy[i] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k * nt >= N) continue;
            acc[0] += (double)__fmul_rn(x[k], w[k]); acc[1] += (double)__fmul_rn(y[k], w[k]);
            acc[2] += (double)__fmul_rn(sx[k], w[k]); acc[3] += (double)__fmul_rn(sy[k], w[k]);
            acc[4] += (double)w[k];
        }
    }
    PF_STAMP(11);
    block_reduce_sum_d5(acc, sh_d5);
    PF_STAMP(12);
    if (tid == 0) {
        for (int q = 0; q < 4; q++) { const float v = (float)(acc[q] / acc[4]); out->average[q] = v; trk_set_avg(d, q, v); }
        out->n_valid = d.st->n_valid; out->resampled = d.st->resampled; out->filter_state = fs;
        out->neff_inv = d.st->neff_inv; out->max_response = d.st->max_response;
    }
    if (fs == MCT_PF_PREDICTED) {
        float mx = -CUDART_INF_F;
        for (int i = tid; i < N; i += nt) mx = fmaxf(mx, d.w[i]);
        mx = block_reduce_minmax(mx, true, sh_f);
        if (tid == 0) { s_first = N; s_last = -1; }
        __syncthreads();
        int lf = N, ll = -1;
        for (int i = tid; i < N; i += nt) if (d.w[i] == mx) { lf = min(lf, i); ll = max(ll, i); }
        if (lf < N) atomicMin(&s_first, lf);
        if (ll >= 0) atomicMax(&s_last, ll);
        __syncthreads();
        if (tid == 0) {
            int a = s_first < N ? s_first : 0, b = s_last >= 0 ? s_last : 0;
            out->ordered_first[0] = d.cx[a]; out->ordered_first[1] = d.cy[a]; out->ordered_first[2] = d.sx[a];
            out->ordered_first[3] = d.sy[a]; out->ordered_first[4] = d.w[a];
            trk_set_first(d, d.cx[a], d.cy[a], d.sx[a], d.sy[a]);
            out->highest_close[0] = d.cx[b]; out->highest_close[1] = d.cy[b]; out->highest_close[2] = d.sx[b];
            out->highest_close[3] = d.sy[b]; out->highest_close[4] = d.w[b];
            out->n_unique = N;
        }
    } else if (fs == MCT_PF_RESAMPLED) {
        // run weights: the run-start thread sums its run in order.  Right after a forced resample every
        // weight is exactly 1.0f, and an in-order f32 sum of L ones is exactly (float)L (L < 2^24): the run end
        // is then found by binary search in the non-decreasing index list instead of walking the run.
        const int* ridx = d.residx;
        float* rw = d.rw;                          // without scratch the resampled-weight column doubles as run-weight scratch
        if (scratch) {
            int* si = reinterpret_cast<int*>(scratch);
            for (int i = tid; i < N; i += nt) si[i] = d.residx[i];
            ridx = si; rw = scratch + ((N + 3) & ~3);
        }
        int not_one = 0;
        for (int i = tid; i < N; i += nt) not_one += (d.w[i] != 1.0f);
        not_one = block_reduce_sum_i(not_one, sh_i);          // (also orders the staging stores above)
        float mx = -CUDART_INF_F;
        int nruns = 0;
        for (int i = tid; i < N; i += nt) {
            const int v = ridx[i];
            const bool start = (i == 0) || (v != ridx[i - 1]);
            float r = 0.0f;
            if (start) {
                nruns++;
                if (!not_one) {
                    int lo = i + 1, hi = N;                      // first j > i with residx[j] != v
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (ridx[mid] == v) lo = mid + 1; else hi = mid; }
                    r = (float)(lo - i);
                } else {
                    r = d.w[i];
                    for (int j = i + 1; j < N && ridx[j] == v; j++) r = __fadd_rn(r, d.w[j]);
                }
                mx = fmaxf(mx, r);
            }
            rw[i] = start ? r : -1.0f;
        }
        PF_STAMP(13);
        mx = block_reduce_minmax(mx, true, sh_f);
        nruns = block_reduce_sum_i(nruns, sh_i);
        int ties = 0;
        for (int i = tid; i < N; i += nt) if (rw[i] == mx) ties++;
        ties = block_reduce_sum_i(ties, sh_i);
        if (tid == 0) {
            // row (ties-1) of the ordered-unique matrix = state of the (ties-1)-th run in index order
            int want = ties - 1, ord = -1, b = 0;
            if (want > 0) {
                for (int i = 0; i < N; i++) if (rw[i] >= 0.0f) { ord++; if (ord == want) { b = i; break; } }
            }
            out->ordered_first[0] = d.cx[0]; out->ordered_first[1] = d.cy[0]; out->ordered_first[2] = d.sx[0];
            out->ordered_first[3] = d.sy[0]; out->ordered_first[4] = mx;
            trk_set_first(d, d.cx[0], d.cy[0], d.sx[0], d.sy[0]);
            out->highest_close[0] = d.cx[b]; out->highest_close[1] = d.cy[b]; out->highest_close[2] = d.sx[b];
            out->highest_close[3] = d.sy[b]; out->highest_close[4] = mx;
            out->n_unique = nruns;
        }
        if (!scratch) {
            __syncthreads();
            for (int i = tid; i < N; i += nt) d.rw[i] = 1.0f;   // restore m_particlesResampledMatrix weights
        }
    } else if (tid == 0) {
        for (int q = 0; q < 5; q++) { out->ordered_first[q] = 0; out->highest_close[q] = 0; }
        trk_set_first(d, 0.0f, 0.0f, 0.0f, 0.0f);
        out->n_unique = 0;
    }
}
__global__ void __launch_bounds__(PF_THREADS) k_pf_summary(const ObjDesc* __restrict__ descs)
{
    __shared__ ObjDesc s_desc;
    summary_body(load_desc(descs + blockIdx.x, &s_desc), nullptr);
}
int launch_pf_summary(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max)
{
    (void)n_max;
    MCT_LAUNCH(ctx, "k_pf_summary", k_pf_summary<<<n_obj, PF_THREADS, 0, ctx->stream>>>(d_desc));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// Foot point of the highest-weight particle (GetAverageParticleOnGroundMatrix,
// ParticleFilterTracker.cpp:1223-1238; GetHighestWeightParticle :529-549 = first maximum).
__global__ void __launch_bounds__(256) k_foot_points(const ObjDesc* __restrict__ descs, const float* __restrict__ h0,
                                                     float* __restrict__ out)
{
    __shared__ float sh_f[32];
    __shared__ int s_first;
    const ObjDesc& d = descs[blockIdx.x];
    const int N = d.N, tid = threadIdx.x, nt = blockDim.x;
    float mx = -CUDART_INF_F;
    for (int i = tid; i < N; i += nt) mx = fmaxf(mx, d.w[i]);
    mx = block_reduce_minmax(mx, true, sh_f);
    if (tid == 0) s_first = N;
    __syncthreads();
    int lf = N;
    for (int i = tid; i < N; i += nt) if (d.w[i] == mx) lf = min(lf, i);
    if (lf < N) atomicMin(&s_first, lf);
    __syncthreads();
    if (tid == 0) {
        int a = s_first < N ? s_first : 0;
        out[2 * blockIdx.x + 0] = d.cx[a];
        out[2 * blockIdx.x + 1] = __fadd_rn(d.cy[a], __fdiv_rn(__fmul_rn(d.sy[a], h0[blockIdx.x]), 2.0f));
    }
}
int launch_foot_points(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, const float* d_h0, float* d_out)
{
    MCT_LAUNCH(ctx, "k_foot_points", k_foot_points<<<n_obj, 256, 0, ctx->stream>>>(d_desc, d_h0, d_out));
    return MCT_OK;
}


// ------------------------------------------------------------------------------------------
// The geometric fuser's exchange over NVLink peer memory, fused with the kernel that produces the payload:
// k_foot_points_push computes this camera's foot points and stores them straight into slot [rank] of EVERY camera's
// receive buffer (peer pointers opened through CUDA IPC), then publishes the frame's sequence number in every buffer's
// flag word; k_foot_points_wait (same stream, next launch) waits until all cameras' flags carry the sequence number and
// hands the gathered points to the host through mapped pinned memory.  Two buffers by sequence parity: a camera can run at
// most one frame ahead of the slowest one (it needs everybody's flag of frame t before it finishes frame t), so the frame it
// may already be writing never collides with the one still being read.  The wait is bounded (timeout -> error flag).
__global__ void __launch_bounds__(256) k_foot_points_push(const ObjDesc* __restrict__ descs, const float* __restrict__ h0, P2PPeers peers,
                                                          int world, int rank, unsigned seq, unsigned* __restrict__ ticket)
{
    __shared__ float sh_f[32];
    __shared__ int s_first;
    const ObjDesc& d = descs[blockIdx.x];
    const int N = d.N, tid = threadIdx.x, nt = blockDim.x, par = seq & 1u;
    float mx = -CUDART_INF_F;
    for (int i = tid; i < N; i += nt) mx = fmaxf(mx, d.w[i]);
    mx = block_reduce_minmax(mx, true, sh_f);
    if (tid == 0) s_first = N;
    __syncthreads();
    int lf = N;
    for (int i = tid; i < N; i += nt) if (d.w[i] == mx) lf = min(lf, i);
    if (lf < N) atomicMin(&s_first, lf);
    __syncthreads();
    if (tid != 0) return;
    const int a = s_first < N ? s_first : 0;
    const float fx = d.cx[a], fy = __fadd_rn(d.cy[a], __fdiv_rn(__fmul_rn(d.sy[a], h0[blockIdx.x]), 2.0f));
    for (int r = 0; r < world; r++) {
        float* slot = peers.p[r]->data[par][rank][blockIdx.x];
        slot[0] = fx; slot[1] = fy;
    }
    __threadfence_system();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {                 // the last object of this camera publishes the frame
        *ticket = 0u;
        __threadfence_system();
        for (int r = 0; r < world; r++) {
            unsigned* f = &peers.p[r]->flags[par][rank];
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
        }
    }
}
__global__ void __launch_bounds__(64) k_foot_points_wait(P2PBuf* __restrict__ mine, int world, int n_obj, unsigned seq, float* __restrict__ out,
                                                         unsigned long long timeout_ns, int* __restrict__ status)
{
    const int par = seq & 1u, tid = threadIdx.x;
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    if (tid < world) {
        unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const unsigned* f = &mine->flags[par][tid];
        for (;;) {
            unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v == seq) break;
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) { s_bad = 1; break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (tid == 0) *status = s_bad;
    for (int q = tid; q < world * n_obj * 2; q += blockDim.x) {
        const int r = q / (n_obj * 2), rem = q - r * n_obj * 2;
        out[q] = mine->data[par][r][rem >> 1][rem & 1];
    }
}
int launch_foot_points_exchange(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, const float* d_h0, const P2PPeers& peers, int world, int rank,
                                unsigned seq, unsigned* d_ticket, float* d_out, unsigned long long timeout_ns, int* d_status)
{
    MCT_LAUNCH(ctx, "k_foot_points_push", k_foot_points_push<<<n_obj, 256, 0, ctx->stream>>>(d_desc, d_h0, peers, world, rank, seq, d_ticket));
    MCT_LAUNCH(ctx, "k_foot_points_wait", k_foot_points_wait<<<1, 64, 0, ctx->stream>>>(peers.p[rank], world, n_obj, seq, d_out, timeout_ns, d_status));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// Geometric-fuser feedback
//
// k_pf_refine: CheckForParticleRefinement + ForceParticleFilterRefinement in one CTA per object, for the call sites
// outside the fused per-frame path (mct_pf_predict, mct_pf_rearrange_ground).
__global__ void __launch_bounds__(256) k_pf_refine(const ObjDesc* __restrict__ descs)
{
    const ObjDesc& d = descs[blockIdx.x];
    int ok = 0;
    for (int i = threadIdx.x; i < d.N; i += blockDim.x) ok |= refine_valid_one(d, i) ? 1 : 0;
    const int any = __syncthreads_or(ok);
    if (!any)
        for (int i = threadIdx.x; i < d.N; i += blockDim.x) force_refine_one(d, i);
    if (threadIdx.x == 0) d.st->refine = 0;
}
int launch_pf_refine(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj)
{
    MCT_LAUNCH(ctx, "k_pf_refine", k_pf_refine<<<n_obj, 256, 0, ctx->stream>>>(d_desc));
    return MCT_OK;
}

// RearrangeParticlesBasedOnGroundLocation (ParticleFilter.cpp:120-179, shouldChangeTheScaleAccordingToOffset): every
// particle's foot point is pulled onto the fused ground location (seen in this camera's image) plus noise by changing its
// scale while its box's top-left corner stays where it was.  nz: [2][N] = the (xRand, yRand) pairs in particle order.
// (batched form: object blockIdx.y takes its mean foot point from ra and its noise block at nz + ra.noise_off[object])
__global__ void __launch_bounds__(256) k_pf_rearrange(const ObjDesc* __restrict__ descs, const float* __restrict__ nz_base, const RearrangeArgs ra)
{
    const ObjDesc& d = descs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.N) return;
    const float* nz = nz_base + ra.noise_off[blockIdx.y];
    const float mfx = ra.mfx[blockIdx.y], mfy = ra.mfy[blockIdx.y];
    const float cx = d.cx[i], cy = d.cy[i], sx = d.sx[i], sy = d.sy[i];
    const float hw = __fdiv_rn(d.w0, 2.0f), hh = __fdiv_rn(d.h0, 2.0f);
    const float kx = (float)(2.0 / (double)d.w0), ky = (float)(1.0 / (double)d.h0);
    const float xpos = fmaxf(0.0f, __fsub_rn(cx, __fmul_rn(hw, sx)));
    const float ypos = fmaxf(0.0f, __fsub_rn(cy, __fmul_rn(hh, sy)));
    const float fy = __fadd_rn(cy, __fmul_rn(hh, sy));
    const float offx = __fadd_rn(__fsub_rn(mfx, cx), nz[i]);
    const float offy = __fadd_rn(__fsub_rn(mfy, fy), nz[(size_t)d.N + i]);
    const float nsx = fminf(fmaxf(__fadd_rn(__fmul_rn(kx, offx), sx), d.mins), d.maxs);
    const float nsy = fminf(fmaxf(__fadd_rn(__fmul_rn(ky, offy), sy), d.mins), d.maxs);
    d.sx[i] = nsx; d.sy[i] = nsy;
    d.cx[i] = __fadd_rn(xpos, __fmul_rn(hw, nsx));
    d.cy[i] = __fadd_rn(ypos, __fmul_rn(hh, nsy));
}
int launch_pf_rearrange(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, const float* d_noise2, const RearrangeArgs& ra)
{
    dim3 g(ceil_div(n_max, 256), n_obj);
    MCT_LAUNCH(ctx, "k_pf_rearrange", k_pf_rearrange<<<g, 256, 0, ctx->stream>>>(d_desc, d_noise2, ra));
    return MCT_OK;
}

// UpdateParticlesWithGroundPDF, the per-particle part (ParticleFilterTracker.cpp:1054-1086): foot point
// (cx, cy + sy*h0/2) -> ground plane through the camera's homography (f64, GeometryBasedInformationFuser.cpp:172-212)
// -> multivariate normal pdf under the Kalman posterior (:111-163; the exponent has no 1/2, as in the reference), the total
// weight, and GetAverageofAllParticles for the rearrangement test that follows on the host.
__global__ void __launch_bounds__(256) k_ground_pdf(const ObjDesc* __restrict__ descs, const GroundArgs* __restrict__ gas, float* __restrict__ wout,
                                                    GroundOut* __restrict__ outs)
{
    const GroundArgs ga = gas[blockIdx.x];
    GroundOut* out = outs + blockIdx.x;
    __shared__ double sh_d5[5 * 32];
    __shared__ double sh_d[32];
    const ObjDesc& d = descs[blockIdx.x];
    const int N = d.N, tid = threadIdx.x, nt = blockDim.x;
    double tot = 0.0, acc[5] = {0, 0, 0, 0, 0};
    for (int i = tid; i < N; i += nt) {
        const float cx = d.cx[i], cy = d.cy[i], sx = d.sx[i], sy = d.sy[i], w = d.w[i];
        float pdf = 1.0f;
        if (ga.usable) {
            const double x = (double)cx, y = (double)__fadd_rn(cy, __fdiv_rn(__fmul_rn(sy, ga.h0), 2.0f));
            const double t0 = __dadd_rn(__dadd_rn(__dmul_rn(ga.H[0], x), __dmul_rn(ga.H[1], y)), ga.H[2]);
            const double t1 = __dadd_rn(__dadd_rn(__dmul_rn(ga.H[3], x), __dmul_rn(ga.H[4], y)), ga.H[5]);
            const double t2 = __dadd_rn(__dadd_rn(__dmul_rn(ga.H[6], x), __dmul_rn(ga.H[7], y)), ga.H[8]);
            const float gx = (float)(t0 / t2), gy = (float)(t1 / t2);
            const float d0 = __fsub_rn(gx, ga.mean[0]), d1 = __fsub_rn(gy, ga.mean[1]);
            const float q0 = (float)__dadd_rn(__dmul_rn((double)d0, (double)ga.ci[0]), __dmul_rn((double)d1, (double)ga.ci[2]));
            const float q1 = (float)__dadd_rn(__dmul_rn((double)d0, (double)ga.ci[1]), __dmul_rn((double)d1, (double)ga.ci[3]));
            const float prod = (float)__dadd_rn(__dmul_rn((double)q0, (double)d0), __dmul_rn((double)q1, (double)d1));
            const float c = expf(-prod);
            pdf = (float)(ga.a * (double)ga.b * (double)c);
            if (pdf < 0) pdf = 0.0f;
        }
        if (wout) wout[i] = pdf;
        tot += (double)pdf;
        acc[0] += (double)__fmul_rn(cx, w); acc[1] += (double)__fmul_rn(cy, w);
        acc[2] += (double)__fmul_rn(sx, w); acc[3] += (double)__fmul_rn(sy, w); acc[4] += (double)w;
    }
    tot = block_reduce_sum_d(tot, sh_d);
    block_reduce_sum_d5(acc, sh_d5);
    if (tid == 0) {
        out->total_weight = (float)tot;
        for (int q = 0; q < 4; q++) out->average[q] = (float)(acc[q] / acc[4]);
    }
}
// one CTA per object; d_wout (optional per-particle read-back) is only meaningful for n_obj == 1
int launch_ground_pdf(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, const GroundArgs* d_ga, float* d_wout, GroundOut* d_out)
{
    MCT_LAUNCH(ctx, "k_ground_pdf", k_ground_pdf<<<n_obj, 256, 0, ctx->stream>>>(d_desc, d_ga, d_wout, d_out));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// k_train_from_particles: GeneratePositiveTrainingSampleSet strategies GREEDY / SEMI_GREEDY / PARTICLE_RANDOM
// (ParticleFilterTracker.cpp:743-892) and the particle read-out of SAMPLE_NEG_PARTICLE_RANDOM (:939-964).
// One CTA per object.  Ordered unique particles (ParticleFilter.cpp:728-815):
//   RESAMPLED  rows = run starts of the (non-decreasing) resample index list in index order; the weights column holds
//              the run weights sorted descending (:799-806), so "weight != 0" holds for the first nz rows, nz = runs with
//              a non-zero member (weights are non-negative);
//   PREDICTED  rows = particles by weight descending, ties by index: exact rank by counting.
// PARTICLE_RANDOM consumes one randfloat() per in-frame resampled particle, in particle order: the k-th in-frame
// particle gets draw k, computed by jumping the multiply-with-carry generator (Public.cpp:15-23: state = lo * A + hi)
// ahead: state_n = A^n * state_0 mod (A * 2^32 - 1) once the state is below the modulus (two plain steps).
#define MWC_A 4164903690ull
#define MWC_M ((MWC_A << 32) - 1ull)
__device__ __forceinline__ unsigned long long mwc_step(unsigned long long s) { return (s & 0xffffffffull) * MWC_A + (s >> 32); }
// a * b mod M for a, b < M by Barrett reduction: q = floor(hi * mu / 2^64) with mu = floor(2^128 / M) = 2^64 + MWC_MU_LO
// underestimates the quotient by at most 2 (checked on 2e5 random pairs and by the parity tests against the sequential stream);
// a generic 128-bit modulo costs ~1000 instructions per call and made the jump-ahead the slowest part of the kernels
#define MWC_MU_LO 0x07fe96ee8b5642a8ull
__device__ __forceinline__ unsigned long long mwc_mulmod(unsigned long long a, unsigned long long b)
{
    const unsigned long long lo = a * b, hi = __umul64hi(a, b);
    const unsigned long long q = hi + __umul64hi(hi, MWC_MU_LO);
    const unsigned long long qlo = q * MWC_M, qhi = __umul64hi(q, MWC_M);
    unsigned long long rlo = lo - qlo;
    unsigned long long rhi = hi - qhi - (lo < qlo ? 1ull : 0ull);
    while (rhi != 0ull || rlo >= MWC_M) { rhi -= (rlo < MWC_M ? 1ull : 0ull); rlo -= MWC_M; }
    return rlo;
}
// A^(2^k) mod M, k = 0..31
__constant__ unsigned long long c_mwc_pw[32] = {
    0x00000000f83f630aull, 0xf0badf95453cbc64ull, 0x5bb3ca800e155d1full, 0x17d406e4d579f267ull, 0xb4e603a4edcddfc5ull, 0x6cbe3044aad17ec3ull,
    0xdd6cfbc2b97fc301ull, 0x4ddc089a9cb56ba1ull, 0x233e7431a129a78dull, 0xdc6d3202a2123589ull, 0x0e4f486bf3391191ull, 0x3899ee7ead25bea8ull,
    0x7d789128b6481463ull, 0x02f797587e026b06ull, 0x71505662b47474f9ull, 0x7ffea99b8bf62055ull, 0x4c8dab023bc54a8eull, 0xcdee9b4a3bd4bf9aull,
    0x783227907a7eae0eull, 0x86f2f38cf3cd17b9ull, 0xc73dfa3fdfcc6159ull, 0xf6e194fd40a1c691ull, 0xba2151430bc11a33ull, 0x80c6cf827191eec5ull,
    0xdf99a36ffd0b83a5ull, 0x73a1b24b33fe2045ull, 0x01fb902f184234bdull, 0xacfff6883014bc66ull, 0xb4a1db6974c2abf6ull, 0xa7367cf9ce7befc8ull,
    0x107eabd8a843b13bull, 0xb0d2605fe7d077ccull};
// pw[k] = A^(2^k) mod M; s0 must already be <= M (two plain steps taken)
__device__ __forceinline__ unsigned long long mwc_jump(unsigned long long s, unsigned n, const unsigned long long* pw)
{
    for (int k = 0; n; k++, n >>= 1) if (n & 1u) s = mwc_mulmod(s, pw[k]);
    return s;
}
// state after n draws from s0
__device__ __forceinline__ unsigned long long mwc_after(unsigned long long s0, unsigned long long s2, unsigned n, const unsigned long long* pw)
{
    if (n == 0) return s0;
    if (n == 1) return mwc_step(s0);
    return mwc_jump(s2, n - 2, pw);
}
// in-order position of the flagged threads of this tile (exclusive) and the tile total
__device__ __forceinline__ int block_compact(bool flag, int* sh /* [33] */, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int inwarp = __popc(bal & ((1u << lane) - 1u));
    __syncthreads();
    if (lane == 0) sh[warp] = __popc(bal);
    __syncthreads();
    int before = 0, tot = 0;
    for (int w = 0; w < nw; w++) { const int c = sh[w]; if (w < warp) before += c; tot += c; }
    *total = tot;
    return before + inwarp;
}
__device__ __forceinline__ bool particle_box_in_frame(float a0, float a1, float a2, float a3, float w0, float h0, int fw, int fh, int* lx, int* ty)
{
    const float wf = __fmul_rn(a2, w0), hf = __fmul_rn(a3, h0);
    const int leftX = __float2int_rn(__fsub_rn(a0, __fdiv_rn(wf, 2.0f))), topY = __float2int_rn(__fsub_rn(a1, __fdiv_rn(hf, 2.0f)));
    const int width = __float2int_rn(wf), height = __float2int_rn(hf);
    *lx = leftX; *ty = topY;
    return leftX > 0 && topY > 0 && leftX + width < fw - 1 && topY + height < fh - 1;
}
__global__ void __launch_bounds__(PF_THREADS)
k_train_from_particles(const ObjDesc* __restrict__ descs, const mct_train_gen* __restrict__ gens, int cap, int* __restrict__ col, int* __restrict__ row,
                       float* __restrict__ sx, float* __restrict__ sy, mct_train_gen_out* __restrict__ outs)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ ObjDesc s_desc;
    __shared__ int sh_c[33];
    __shared__ float sh_f[32];
    __shared__ int sh_i[32];
    __shared__ unsigned long long s_pw[32];
    const ObjDesc& d = load_desc(descs + blockIdx.x, &s_desc);
    const mct_train_gen g = gens[blockIdx.x];
    const int N = d.N, tid = threadIdx.x, nt = blockDim.x;
    const int fs = d.st->filter_state;
    int* o_col = col + (size_t)blockIdx.x * cap; int* o_row = row + (size_t)blockIdx.x * cap;
    float* o_sx = sx + (size_t)blockIdx.x * cap; float* o_sy = sy + (size_t)blockIdx.x * cap;
    float* sw = (float*)smraw;                         // [N] weights (PREDICTED rank pass)
    int* rowidx = (int*)(sw + ((N + 3) & ~3));         // [max_pos] particle index of the first ordered unique rows
    const int max_pos = g.max_pos > 0 ? g.max_pos : 0;
    const float w0 = g.cur[2], h0 = g.cur[3];
    const float ccx = __fadd_rn(g.cur[0], __fdiv_rn(__fmul_rn(g.cur[2], g.cur[4]), 2.0f));
    const float ccy = __fadd_rn(g.cur[1], __fdiv_rn(__fmul_rn(g.cur[3], g.cur[5]), 2.0f));

    // ---- ordered unique particles: first max_pos rows, count, non-zero rows, largest centre distances
    int n_uniq = 0, nz = 0x7fffffff;
    float mdx = 0.0f, mdy = 0.0f;
    if (fs == MCT_PF_RESAMPLED) {
        int base = 0, nzc = 0;
        for (int i0 = 0; i0 < N; i0 += nt) {
            const int i = i0 + tid;
            const bool start = i < N && (i == 0 || d.residx[i] != d.residx[i - 1]);
            int tot;
            const int q = base + block_compact(start, sh_c, &tot);
            if (start) {
                if (q < max_pos) rowidx[q] = i;
                const int v = d.residx[i];
                bool any = false;
                for (int j = i; j < N && d.residx[j] == v; j++) if (d.w[j] != 0.0f) { any = true; break; }
                nzc += any ? 1 : 0;
                mdx = fmaxf(mdx, fabsf(__fsub_rn(d.cx[i], ccx))); mdy = fmaxf(mdy, fabsf(__fsub_rn(d.cy[i], ccy)));
            }
            base += tot;
        }
        n_uniq = base;
        nz = block_reduce_sum_i(nzc, sh_i);
    } else if (fs == MCT_PF_PREDICTED) {
        // Only the first max_pos rows are needed.  Weights are non-negative, so their bit patterns order like the values: a radix
        // select (4 passes of 8 bits over a shared histogram) finds the max_pos-th largest pattern, and exact ranks (ties by
        // index) are then counted only for the particles at or above it, one warp per candidate.
        __shared__ unsigned s_hist[256];
        __shared__ unsigned s_prefix, s_want;
        for (int i = tid; i < N; i += nt) {
            sw[i] = d.w[i];
            mdx = fmaxf(mdx, fabsf(__fsub_rn(d.cx[i], ccx))); mdy = fmaxf(mdy, fabsf(__fsub_rn(d.cy[i], ccy)));
        }
        if (tid == 0) { s_prefix = 0u; s_want = (unsigned)min(max(max_pos, 1), N); }
        __syncthreads();
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int b = tid; b < 256; b += nt) s_hist[b] = 0u;
            __syncthreads();
            const unsigned prefix = s_prefix, himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int i = tid; i < N; i += nt) {
                const unsigned u = __float_as_uint(sw[i]);
                if ((u & himask) == prefix) atomicAdd(&s_hist[(u >> shift) & 0xffu], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned want = s_want, b = 255;
                for (;; b--) { const unsigned c = s_hist[b]; if (c >= want || b == 0) break; want -= c; }
                s_want = want; s_prefix = prefix | (b << shift);
            }
            __syncthreads();
        }
        const float thr = __uint_as_float(s_prefix);                     // the min(max_pos, N)-th largest weight
        {
            const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
            for (int i0 = warp * 32; i0 < N; i0 += nw * 32) {
                const int i = i0 + lane;
                unsigned cand = __ballot_sync(0xffffffffu, i < N && sw[i] >= thr);
                while (cand) {
                    const int src = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int ci = i0 + src;
                    const float v = sw[ci];
                    int cnt = 0;
                    for (int j = lane; j < N; j += 32) { const float o = sw[j]; cnt += (o > v || (o == v && j < ci)) ? 1 : 0; }
                    cnt = __reduce_add_sync(0xffffffffu, cnt);
                    if (lane == 0 && cnt < max_pos) rowidx[cnt] = ci;
                }
            }
        }
        n_uniq = N;
    }
    mdx = block_reduce_minmax(mdx, true, sh_f);
    mdy = block_reduce_minmax(mdy, true, sh_f);
    __syncthreads();

    int n_pos = 0, n_draws = 0;
    unsigned long long rng_after = g.rng_state;
    if (g.strategy == 1 || g.strategy == 2) {
        const int n = min(n_uniq, max_pos);
        int base = 0;
        for (int q0 = 0; q0 < n; q0 += nt) {
            const int q = q0 + tid;
            bool ok = false; int lx = 0, ty = 0; float a2 = 0, a3 = 0;
            if (q < n) {
                const int i = rowidx[q];
                const float a0 = d.cx[i], a1 = d.cy[i]; a2 = d.sx[i]; a3 = d.sy[i];
                const bool wnz = fs == MCT_PF_RESAMPLED ? q < nz : d.w[i] != 0.0f;
                ok = wnz && particle_box_in_frame(a0, a1, a2, a3, w0, h0, g.frame_w, g.frame_h, &lx, &ty);
                if (g.strategy == 2) {
                    const double ddx = (double)__fsub_rn(ccx, a0), ddy = (double)__fsub_rn(ccy, a1);
                    const float dist = (float)sqrt(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));
                    ok = ok && dist < 8.0f;
                }
            }
            int tot;
            const int pos = base + block_compact(ok, sh_c, &tot);
            if (ok && pos < cap) { o_col[pos] = lx; o_row[pos] = ty; o_sx[pos] = a2; o_sy[pos] = a3; }
            base += tot;
        }
        n_pos = min(base, cap);
    } else if (g.strategy == 3) {
        const unsigned long long s0 = g.rng_state, s2 = mwc_step(mwc_step(s0));
        if (tid < 32) s_pw[tid] = c_mwc_pw[tid];
        __syncthreads();
        const float prob = __fdiv_rn((float)g.max_pos, (float)N);
        int base_in = 0, base_keep = 0;
        for (int i0 = 0; i0 < N; i0 += nt) {
            const int i = i0 + tid;
            int lx = 0, ty = 0; float a2 = 0, a3 = 0;
            bool in = false;
            if (i < N) { a2 = d.rsx[i]; a3 = d.rsy[i]; in = particle_box_in_frame(d.rcx[i], d.rcy[i], a2, a3, w0, h0, g.frame_w, g.frame_h, &lx, &ty); }
            int tot_in;
            const int k = base_in + block_compact(in, sh_c, &tot_in);
            bool keep = false;
            if (in) {
                const unsigned long long s = mwc_after(s0, s2, (unsigned)k + 1u, s_pw);
                const float rf = (float)((double)(unsigned)(s & 0xffffffffull) * 2.3283064365386962890625e-10);
                keep = rf < prob;
            }
            int tot_keep;
            const int pos = base_keep + block_compact(keep, sh_c, &tot_keep);
            if (keep && pos < cap && pos < max_pos) { o_col[pos] = lx; o_row[pos] = ty; o_sx[pos] = a2; o_sy[pos] = a3; }
            base_in += tot_in; base_keep += tot_keep;
        }
        n_draws = base_in;
        n_pos = min(min(base_keep, max_pos), cap);
        rng_after = mwc_after(s0, s2, (unsigned)base_in, s_pw);
    }
    if (tid == 0) {
        mct_train_gen_out o;
        o.n_pos = n_pos; o.n_unique = n_uniq; o.max_dx = mdx; o.max_dy = mdy; o.rng_state = rng_after; o.n_draws = n_draws; o.pad = 0;
        outs[blockIdx.x] = o;
    }
}
int launch_train_from_particles(mct_ctx* ctx, const ObjDesc* d_desc, const ObjDesc* h_desc, int n_obj, const mct_train_gen* d_gen, int max_pos_max,
                                int cap, int* d_col, int* d_row, float* d_sx, float* d_sy, mct_train_gen_out* d_out)
{
    int n_max = 0;
    for (int o = 0; o < n_obj; o++) n_max = max(n_max, h_desc[o].N);
    const size_t smem = ((size_t)((n_max + 3) & ~3) + (size_t)max(max_pos_max, 1)) * 4;
    if (smem > 200 * 1024) return mct_fail(ctx, MCT_E_UNSUPPORTED, "too many particles for the training-sample kernel%s (%d)", "", n_max);
    static size_t s_attr[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_train_from_particles, s_attr, smem, 48 * 1024)) return rc_;
    MCT_LAUNCH(ctx, "k_train_from_particles", k_train_from_particles<<<n_obj, PF_THREADS, smem, ctx->stream>>>(d_desc, d_gen, cap, d_col, d_row, d_sx, d_sy, d_out));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// k_sample_regions: SampleSet::SampleImage (ring :82-168, rectangle :180-262) + SelectSamplesUniformlyFromLargerSet (:331-361).
// One CTA per region; pass 1 counts the candidates (the keep probability needs the count), pass 2 gives candidate k draw k.
__device__ __forceinline__ bool region_candidate(const mct_sample_region& g, int r, int c, float o2, float i2)
{
    if (g.kind == 0) {
        const int dist = (g.y - r) * (g.y - r) + (g.x - c) * (g.x - c);
        return (float)dist < o2 && (float)dist >= i2;
    }
    const int dx = abs(g.x - c), dy = abs(g.y - r);
    return (float)dx >= g.p[2] || (float)dy >= g.p[3];
}
// body shared by k_sample_regions and k_train_gen_default: all threads of the CTA call it; the result lands in *res (shared memory)
// and is visible to every thread on return.  bbox (shared int[4], or NULL): min col, max col, min row, max row of the KEPT samples.
// The candidates (a sparse ring inside its bounding square) are first compacted, in row-major order, into a shared list; the
// thinning then runs densely over that list: a thread owns a block of consecutive candidates, jumps the stream ONCE to its block's
// first draw and steps it from there (a jump is ~10 modular multiplications of ~80 instructions, a step is three) -- with one jump
// per candidate inside the square's row-major walk nearly every warp paid the full jump for a few live lanes (39 us per frame).
// Regions whose candidates do not fit the list keep the walk with one jump per candidate.
__device__ __forceinline__ int block_excl_scan_i(int v, int* sh /* [33] */, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    int before = 0, tot = 0;
    for (int w = 0; w < nw; w++) { const int c = sh[w]; if (w < warp) before += c; tot += c; }
    *total = tot;
    return before + inc - v;
}
#define SR_LIST_CAP 6144
// Ring / disc form by rows: the candidates of a row are at most two runs of consecutive columns, known in closed form
// ((float)dist < o2 && (float)dist >= i2 for an integer dist  <=>  ceil(i2) <= dist <= ceil(o2) - 1; the same bounds as the host
// sampler, host/mct_host.cpp SampleImage).  One thread per row describes its runs, one block scan numbers the candidates, and a
// thread then owns a block of consecutive candidates: no per-cell work at all.  Returns false (nothing written) when the region
// does not fit this plan; the caller then takes the cell walk.
#define SR_ROWS_CAP 512
__device__ __forceinline__ int isqrt_floor_dev(int v)
{
    int r = (int)sqrtf((float)v);
    while (r * r > v) r--;
    while ((r + 1) * (r + 1) <= v) r++;
    return r;
}
__device__ bool sample_ring_rows(const mct_sample_region& g, int cap, int* __restrict__ o_col, int* __restrict__ o_row,
                                 mct_sample_region_out* res, int* bbox, int* sh_c, const unsigned long long* s_pw, int* s_buf,
                                 int minrow, int mincol, int maxcol, int RH, float o2, float i2)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    int* a0 = s_buf; int* an = s_buf + SR_ROWS_CAP; int* b0 = s_buf + 2 * SR_ROWS_CAP; int* bn = s_buf + 3 * SR_ROWS_CAP;
    int* pre = s_buf + 4 * SR_ROWS_CAP;                                   // [RH + 1] candidates in front of row r
    const int dHi = (int)ceilf(o2) - 1, dLo = (int)ceilf(i2);
    int cnt = 0;
    if (tid < RH) {
        const int r = minrow + tid;
        const int dy2 = (g.y - r) * (g.y - r);
        int sa = 0, na = 0, sb = 0, nb = 0;
        if (dHi - dy2 >= 0 && mincol <= maxcol) {
            const int b = isqrt_floor_dev(dHi - dy2);                     // |x - c| <= b
            const int a = (dLo - dy2 > 0) ? isqrt_floor_dev(dLo - dy2 - 1) + 1 : 0;   // |x - c| >= a
            if (a <= b) {
                if (a == 0) {
                    const int c0 = max(mincol, g.x - b), c1 = min(maxcol, g.x + b);
                    if (c0 <= c1) { sa = c0; na = c1 - c0 + 1; }
                } else {
                    int c0 = max(mincol, g.x - b), c1 = min(maxcol, g.x - a);
                    if (c0 <= c1) { sa = c0; na = c1 - c0 + 1; }
                    c0 = max(mincol, g.x + a); c1 = min(maxcol, g.x + b);
                    if (c0 <= c1) { sb = c0; nb = c1 - c0 + 1; }
                }
            }
        }
        a0[tid] = sa; an[tid] = na; b0[tid] = sb; bn[tid] = nb;
        cnt = na + nb;
    }
    PF_STAMP(20);
    int n_cand;
    const int before = block_excl_scan_i(cnt, sh_c, &n_cand);
    PF_STAMP(21);
    if (n_cand > 32 * nt) { __syncthreads(); return false; }               // (uniform: every thread sees the same total)
    if (tid < RH) pre[tid] = before;
    if (tid == 0) pre[RH] = n_cand;
    __syncthreads();
    const int n_list = n_cand;
    const float prob = __fdiv_rn((float)(g.max_n + 1), (float)(unsigned)n_list);
    const bool thin = prob < 1.0f;
    const unsigned long long s0 = g.rng_state, s2 = mwc_step(mwc_step(s0));
    const int B = max(thin ? 8 : 1, (n_cand + nt - 1) / nt);               // (a jump per thread when thinning: longer blocks)
    const int k0 = tid * B, k1 = min(n_cand, k0 + B);
    // the stream after all draws: by a thread that owns no candidates when there is one (beside the others' jumps)
    unsigned long long rng_after = s0;
    const int fin = (n_cand + B - 1) / B < nt ? nt - 1 : 0;
    if (thin && tid == fin) rng_after = mwc_after(s0, s2, (unsigned)n_list, s_pw);
    unsigned keepmask = 0u;
    int r0 = 0;
    if (k0 < k1) {
        int lo = 0, hi = RH - 1;                                          // last row with pre[row] <= k0
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pre[mid] <= k0) lo = mid; else hi = mid - 1; }
        r0 = lo;
        if (thin) {
            unsigned long long st = mwc_after(s0, s2, (unsigned)k0, s_pw);
            for (int k = k0; k < k1; k++) {
                st = mwc_step(st);
                const float rf = (float)((double)(unsigned)(st & 0xffffffffull) * 2.3283064365386962890625e-10);
                if (!(rf > prob)) keepmask |= 1u << (k - k0);
            }
        } else keepmask = (k1 - k0 >= 32) ? 0xffffffffu : ((1u << (k1 - k0)) - 1u);
    }
    PF_STAMP(22);
    int tot_k;
    int pos = block_excl_scan_i(__popc(keepmask), sh_c, &tot_k);
    PF_STAMP(23);
    int bx0 = 0x7fffffff, bx1 = -0x7fffffff, by0 = 0x7fffffff, by1 = -0x7fffffff;
    if (k0 < k1) {
        int r = r0, off = k0 - pre[r0];
        for (int k = k0; k < k1; k++) {
            while (off >= an[r] + bn[r]) { off -= an[r] + bn[r]; r++; }   // (rows without candidates are skipped)
            if ((keepmask >> (k - k0)) & 1u) {
                if (pos < cap && (!thin || pos < g.max_n)) {
                    const int c = off < an[r] ? a0[r] + off : b0[r] + (off - an[r]);
                    const int rr = minrow + r;
                    o_col[pos] = c; o_row[pos] = rr;
                    bx0 = min(bx0, c); bx1 = max(bx1, c); by0 = min(by0, rr); by1 = max(by1, rr);
                }
                pos++;
            }
            off++;
        }
    }
    if (bbox && bx0 <= bx1) { atomicMin(&bbox[0], bx0); atomicMax(&bbox[1], bx1); atomicMin(&bbox[2], by0); atomicMax(&bbox[3], by1); }
    if (tid == fin) {
        mct_sample_region_out o;
        o.n = min(thin ? min(tot_k, g.max_n) : n_list, cap);
        o.n_candidates = n_cand; o.n_draws = thin ? n_list : 0; o.pad = 0;
        o.rng_state = rng_after;
        *res = o;
    }
    __syncthreads();
    return true;
}
__device__ void sample_region_body(const mct_sample_region& g, int cap, int* __restrict__ o_col, int* __restrict__ o_row,
                                   mct_sample_region_out* res, int* bbox, int* sh_c /* [33] */, int* sh_i /* [32] */,
                                   const unsigned long long* s_pw /* [32] shared copy of c_mwc_pw */, int* s_list /* [SR_LIST_CAP] shared */)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int scaledW = __float2int_rn(__fmul_rn((float)g.width, g.scale_x)), scaledH = __float2int_rn(__fmul_rn((float)g.height, g.scale_y));
    const int nrows = g.frame_h - scaledH - 1, ncols = g.frame_w - scaledW - 1;
    const int ry = (int)(g.kind == 0 ? g.p[0] : g.p[1]), rx = (int)g.p[0];
    const int minrow = max(0, g.y - ry), maxrow = min(nrows - 1, g.y + ry);
    const int mincol = max(0, g.x - rx), maxcol = min(ncols - 1, g.x + rx);
    const int RW = maxcol - mincol + 1, RH = maxrow - minrow + 1;
    const int cells = (RW > 0 && RH > 0) ? RW * RH : 0;
    const float o2 = __fmul_rn(g.p[0], g.p[0]), i2 = __fmul_rn(g.p[1], g.p[1]);
    if (bbox && tid == 0) { bbox[0] = 0x7fffffff; bbox[1] = -0x7fffffff; bbox[2] = 0x7fffffff; bbox[3] = -0x7fffffff; }
    if (g.kind == 0 && RH <= SR_ROWS_CAP && RH <= nt && sample_ring_rows(g, cap, o_col, o_row, res, bbox, sh_c, s_pw, s_list,
                                                                          minrow, mincol, maxcol, RH, o2, i2)) return;
    // the candidates, counted and (as far as the list reaches) compacted in row-major order
    int n_cand = 0;
    for (int e0 = 0; e0 < cells; e0 += nt) {
        const int e = e0 + tid;
        const bool cand = e < cells && region_candidate(g, minrow + e / max(RW, 1), mincol + e % max(RW, 1), o2, i2);
        int tot_c;
        const int k = n_cand + block_compact(cand, sh_c, &tot_c);
        if (cand && k < SR_LIST_CAP) s_list[k] = e;
        n_cand += tot_c;
    }
    __syncthreads();
    (void)sh_i;
    const int n_list = n_cand;
    const float prob = __fdiv_rn((float)(g.max_n + 1), (float)(unsigned)n_list);
    const bool thin = prob < 1.0f;
    const unsigned long long s0 = g.rng_state, s2 = mwc_step(mwc_step(s0));
    int bx0 = 0x7fffffff, bx1 = -0x7fffffff, by0 = 0x7fffffff, by1 = -0x7fffffff;
    int n_kept;
    if (n_cand <= SR_LIST_CAP && (n_cand + nt - 1) / nt <= 32) {
        // candidate k takes draw k: blocks of B consecutive candidates per thread
        const int B = max(8, (n_cand + nt - 1) / nt);                    // <= 12 for a full list
        const int k0 = tid * B, k1 = min(n_cand, k0 + B);
        unsigned keepmask = 0u;
        if (k0 < k1) {
            if (thin) {
                unsigned long long st = mwc_after(s0, s2, (unsigned)k0, s_pw);
                for (int k = k0; k < k1; k++) {
                    st = mwc_step(st);
                    const float rf = (float)((double)(unsigned)(st & 0xffffffffull) * 2.3283064365386962890625e-10);
                    if (!(rf > prob)) keepmask |= 1u << (k - k0);
                }
            } else keepmask = (k1 - k0 >= 32) ? 0xffffffffu : ((1u << (k1 - k0)) - 1u);
        }
        int tot_k;
        int pos = block_excl_scan_i(__popc(keepmask), sh_c, &tot_k);
        for (int k = k0; k < k1; k++) {
            if (!((keepmask >> (k - k0)) & 1u)) continue;
            if (pos < cap && (!thin || pos < g.max_n)) {
                const int e = s_list[k];
                const int r = minrow + e / RW, c = mincol + e % RW;
                o_col[pos] = c; o_row[pos] = r;
                bx0 = min(bx0, c); bx1 = max(bx1, c); by0 = min(by0, r); by1 = max(by1, r);
            }
            pos++;
        }
        n_kept = tot_k;
    } else {
        int base_c = 0, base_k = 0;
        for (int e0 = 0; e0 < cells; e0 += nt) {
            const int e = e0 + tid;
            const int r = minrow + e / max(RW, 1), c = mincol + e % max(RW, 1);
            const bool cand = e < cells && region_candidate(g, r, c, o2, i2);
            int tot_c;
            const int k = base_c + block_compact(cand, sh_c, &tot_c);
            bool keep = cand && k < n_list;
            if (keep && thin) {
                const unsigned long long s = mwc_after(s0, s2, (unsigned)k + 1u, s_pw);
                const float rf = (float)((double)(unsigned)(s & 0xffffffffull) * 2.3283064365386962890625e-10);
                keep = !(rf > prob);
            }
            int tot_k;
            const int pos = base_k + block_compact(keep, sh_c, &tot_k);
            if (keep && pos < cap && (!thin || pos < g.max_n)) {
                o_col[pos] = c; o_row[pos] = r;
                bx0 = min(bx0, c); bx1 = max(bx1, c); by0 = min(by0, r); by1 = max(by1, r);
            }
            base_c += tot_c; base_k += tot_k;
        }
        n_kept = base_k;
    }
    if (bbox && bx0 <= bx1) { atomicMin(&bbox[0], bx0); atomicMax(&bbox[1], bx1); atomicMin(&bbox[2], by0); atomicMax(&bbox[3], by1); }
    if (tid == 0) {
        mct_sample_region_out o;
        o.n = min(thin ? min(n_kept, g.max_n) : n_list, cap);
        o.n_candidates = n_cand; o.n_draws = thin ? n_list : 0; o.pad = 0;
        o.rng_state = thin ? mwc_after(s0, s2, (unsigned)n_list, s_pw) : s0;
        *res = o;
    }
    __syncthreads();
}
__global__ void __launch_bounds__(PF_THREADS)
k_sample_regions(const mct_sample_region* __restrict__ regs, int cap, int* __restrict__ col, int* __restrict__ row, mct_sample_region_out* __restrict__ outs)
{
    __shared__ int sh_c[33];
    __shared__ int sh_i[32];
    __shared__ unsigned long long s_pw[32];
    __shared__ mct_sample_region_out s_res;
    __shared__ int s_list[SR_LIST_CAP];
    const mct_sample_region g = regs[blockIdx.x];
    if (threadIdx.x < 32) s_pw[threadIdx.x] = c_mwc_pw[threadIdx.x];
    __syncthreads();
    sample_region_body(g, cap, col + (size_t)blockIdx.x * cap, row + (size_t)blockIdx.x * cap, &s_res, nullptr, sh_c, sh_i, s_pw, s_list);
    if (threadIdx.x == 0) outs[blockIdx.x] = s_res;
}
int launch_sample_regions(mct_ctx* ctx, int n, const mct_sample_region* d_reg, int cap, int* d_col, int* d_row, mct_sample_region_out* d_out)
{
    MCT_LAUNCH(ctx, "k_sample_regions", k_sample_regions<<<n, PF_THREADS, 0, ctx->stream>>>(d_reg, cap, d_col, d_row, d_out));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// k_train_gen_default: GenerateTrainingSampleSet with the default strategies for all objects of a camera, on the device, queued
// behind the frame's k_pf_update (one CTA per object):
//   m_currentStateList from the particle read-out (ParticleFilterTracker.cpp:655-672; PfTrk holds what k_pf_update left),
//   positives = disc of m_posRadiusTrain around the state's corner (:727-741), negatives = ring between m_posRadiusTrain + 15 and
//   1.5 m_searchWindSize (:926-937), both thinned like SelectSamplesUniformlyFromLargerSet with the tracker's own stream (positives
//   first), boxes written behind each other into the object's slot of the training stage; then the set sizes and the plan of the
//   pixel-list colour-histogram kernels (the device form of mdch_group_plan, mct_api.cu) go straight into the object's TrainDesc.
// A plan outside what the host sized the launches for (TrainGenDev::*_b) sends the object to the generic per-box kernel.
__device__ bool gen_group_plan(MdchGroup* g, int first, int n, float sx, float sy, const int* bbox, const TrainGenDev& q, int gi)
{
    MdchGroup z = {};
    z.first = first; z.n = n;
    *g = z;
    if (n < 1 || q.dim > 1024) return false;
    z.rows = __float2int_rn(__fmul_rn((float)q.ch0, sy)); z.cols = __float2int_rn(__fmul_rn((float)q.cw0, sx));
    if (z.rows < 1 || z.cols < 1 || z.rows * z.cols > MCT_MDCH_TRAIN_PIX) return false;
    z.x0 = bbox[0]; z.y0 = bbox[2]; z.RW = bbox[1] - bbox[0] + z.cols; z.RH = bbox[3] - bbox[2] + z.rows;
    z.npx = z.RW * z.RH;
    if (z.npx > 65536 || z.RW > 2047 || z.RH > 2047 || z.npx > q.npx_b[gi]) return false;
    const int ext = (2 * z.RW - z.cols) * (2 * z.RH - z.rows);
    z.EW = (ext <= MTR_EXT_WORDS && ext <= q.tab_b) ? 1 : 0; z.EH = 0;
    if (!z.EW && z.rows * z.cols > q.tab_b) return false;
    const int T = 384;
    const int B = min((n + 31) & ~31, T), S = T / B;
    z.chunk_px = S * 160;
    z.nchunk = (z.npx + z.chunk_px - 1) / z.chunk_px;
    if (z.nchunk > q.nch_b) return false;
    const int shift = __clz(z.rows * z.cols);
    z.shift = shift > MCT_FIX_SHIFT_MAX ? MCT_FIX_SHIFT_MAX : shift;
    *g = z;
    return true;
}
__global__ void __launch_bounds__(PF_THREADS)
k_train_gen_default(const TrainGenDev* __restrict__ gens, TrainDesc* __restrict__ descs, mct_train_default_out* __restrict__ outs)
{
    __shared__ int sh_c[33];
    __shared__ int sh_i[32];
    __shared__ unsigned long long s_pw[32];
    __shared__ mct_sample_region_out s_res[2];
    __shared__ int s_bbox[2][4];
    __shared__ int s_list[SR_LIST_CAP];
    PF_STAMP(0);
    const TrainGenDev q = gens[blockIdx.x];
    TrainDesc& t = descs[blockIdx.x];
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid < 32) s_pw[tid] = c_mwc_pw[tid];
    __syncthreads();
    pdl_wait();                                // the read-out mirror and the stream state come from the frame's k_pf_update
    PF_STAMP(1);
    const float* a = q.opt == 0 ? q.trk->first : q.trk->avg;
    const float a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3];
    const int n_valid = q.st->n_valid;
    const unsigned long long rng0 = q.trk->rng;
    float cur[6];
    {
        const float wf = __fmul_rn(a2, q.w0), hf = __fmul_rn(a3, q.h0);
        const float x = __fsub_rn(a0, __fdiv_rn(wf, 2.0f)), y = __fsub_rn(a1, __fdiv_rn(hf, 2.0f));
        cur[0] = x < 0.0f ? 0.0f : x; cur[1] = y < 0.0f ? 0.0f : y; cur[2] = q.w0; cur[3] = q.h0; cur[4] = a2; cur[5] = a3;
    }
    int* col = q.stage; int* row = q.stage + q.mcap;
    float* sx = (float*)(q.stage + 2 * q.mcap); float* sy = (float*)(q.stage + 3 * q.mcap);
    int P = 0, Nn = 0;
    unsigned long long rng = rng0;
    int n_draws = 0;
    float nsx = 1.0f, nsy = 1.0f;
    const bool usable = n_valid > 0 && a0 == a0 && a1 == a1 && a2 == a2 && a3 == a3;
    PF_STAMP(2);
    if (usable) {
        mct_sample_region g;
        g.kind = 0; g.x = (int)cur[0]; g.y = (int)cur[1]; g.width = (int)cur[2]; g.height = (int)cur[3];
        g.p[0] = q.pos_radius; g.p[1] = 0.0f; g.p[2] = 0.0f; g.p[3] = 0.0f;
        g.max_n = q.max_pos; g.scale_x = cur[4]; g.scale_y = cur[5]; g.frame_w = q.frame_w; g.frame_h = q.frame_h; g.rng_state = rng;
        sample_region_body(g, q.cap_pos, col, row, &s_res[0], s_bbox[0], sh_c, sh_i, s_pw, s_list);
        P = s_res[0].n; rng = s_res[0].rng_state; n_draws = s_res[0].n_draws;
        PF_STAMP(3);
        nsx = fminf(cur[4], q.maxs); nsy = fminf(cur[5], q.maxs);
        g.x = __float2int_rn(cur[0]); g.y = __float2int_rn(cur[1]); g.width = __float2int_rn(cur[2]); g.height = __float2int_rn(cur[3]);
        g.p[0] = q.neg_outer; g.p[1] = q.neg_inner;
        g.max_n = q.max_neg; g.scale_x = nsx; g.scale_y = nsy; g.rng_state = rng;
        sample_region_body(g, q.cap_neg, col + P, row + P, &s_res[1], s_bbox[1], sh_c, sh_i, s_pw, s_list);
        Nn = s_res[1].n; rng = s_res[1].rng_state; n_draws += s_res[1].n_draws;
        PF_STAMP(4);
        for (int i = tid; i < P + Nn; i += nt) { sx[i] = i < P ? cur[4] : nsx; sy[i] = i < P ? cur[5] : nsy; }
    }
    if (tid == 0) {
        const bool ok = usable && P >= 1 && Nn >= 1;
        int fast = 0;
        if (ok && q.mdch) {
            const bool ok0 = gen_group_plan(&t.grp[0], 0, P, cur[4], cur[5], s_bbox[0], q, 0);
            const bool ok1 = gen_group_plan(&t.grp[1], P, Nn, nsx, nsy, s_bbox[1], q, 1);
            if (ok0 && ok1) { fast = 1; t.grp[0].list_off = 0; t.grp[1].list_off = t.grp[0].npx; }
        }
        t.P = ok ? P : 0; t.Nn = ok ? Nn : 0; t.mdch_fast = fast;
        q.trk->rng = rng;
        mct_train_default_out o;
        o.P = P; o.Nn = Nn; o.status = ok ? MCT_OK : MCT_E_EMPTY; o.mdch_fast = fast; o.rng_state = rng;
        for (int k = 0; k < 6; k++) o.cur[k] = cur[k];
        o.n_draws = n_draws; o.pad = 0;
        outs[blockIdx.x] = o;
    }
    PF_STAMP(5);
    PF_COMMIT(0);
}
int launch_train_gen_default(mct_ctx* ctx, int n_obj, const TrainGenDev* d_gen, TrainDesc* d_tdesc, mct_train_default_out* d_out, bool fresh_args)
{
    // (arguments or descriptors uploaded in front of this launch: the plain launch keeps the kernel behind the copies)
    if (fresh_args) MCT_LAUNCH(ctx, "k_train_gen_default", k_train_gen_default<<<n_obj, PF_THREADS, 0, ctx->stream>>>(d_gen, d_tdesc, d_out));
    else MCT_LAUNCH_PDL(ctx, "k_train_gen_default", k_train_gen_default, dim3(n_obj), dim3(PF_THREADS), (size_t)0, d_gen, d_tdesc, d_out);
    return MCT_OK;
}
