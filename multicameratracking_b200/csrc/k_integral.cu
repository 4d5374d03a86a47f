// k_integral.cu -- integral image (Matrixu::initII, Matrix.cpp:33-47), rect-sum probe
// (Matrixu::sumRect, Matrix.cpp:49-67) and the MDCH joint-bin image.
//
// Integral: exact u32 prefix sums rounded once to f32 (order independent, so bit-exact against the
// oracle by construction).  Two launches:
//   k_band_colsum   : per band of BH rows, per column, the u32 column sum            (reads the frame once)
//   k_integral_band : one CTA per band: (a) column offsets from the bands above + scan over x,
//                     (b) warp-shuffle row prefix of every band row into shared memory,
//                     (c) thread-per-column walk down the band, coalesced f32 stores.
// HBM-bound in principle (W*H u8 in, (W+1)*(H+1) f32 out); at 640x480 it is launch-latency bound.
#include "mct_internal.cuh"
#include <math.h>

__global__ void k_band_colsum(const uint8_t* __restrict__ img, int W, int H, int pitch, int BH,
                              uint32_t* __restrict__ bandsum)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= W) return;
    int b = blockIdx.y;
    int r0 = b * BH, r1 = min(H, r0 + BH);
    uint32_t s = 0;
    for (int r = r0; r < r1; r++) s += img[(size_t)r * pitch + c];
    bandsum[(size_t)b * W + c] = s;
}

// inclusive warp scan
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(1024)
k_integral_band(const uint8_t* __restrict__ img, int W, int H, int pitch, int BH,
                const uint32_t* __restrict__ bandsum, float* __restrict__ ii, int iip)
{
    extern __shared__ uint32_t sm[];
    uint32_t* base = sm;                 // [W]   inclusive-over-x sum of everything above the band
    uint32_t* rp = sm + W;               // [BH][W] row prefix sums of the band rows
    const int b = blockIdx.x;
    const int r0 = b * BH;
    const int nr = min(BH, H - r0);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

    // (a1) column sums of all bands above
    for (int c = tid; c < W; c += nt) {
        uint32_t s = 0;
        for (int bb = 0; bb < b; bb++) s += bandsum[(size_t)bb * W + c];
        base[c] = s;
    }
    __syncthreads();
    // (a2) warp 0 scans base[] over x in chunks of 128 columns (4 per lane) with a running carry;
    // (b)  the other warps (and warp 0 afterwards) build the row prefixes.
    if (warp == 0) {
        uint32_t carry = 0;
        for (int c0 = 0; c0 < W; c0 += 128) {
            int c = c0 + lane * 4;
            uint32_t v0 = c + 0 < W ? base[c + 0] : 0, v1 = c + 1 < W ? base[c + 1] : 0;
            uint32_t v2 = c + 2 < W ? base[c + 2] : 0, v3 = c + 3 < W ? base[c + 3] : 0;
            v1 += v0; v2 += v1; v3 += v2;
            uint32_t incl = warp_incl_scan(v3, lane);
            uint32_t off = carry + incl - v3;
            if (c + 0 < W) base[c + 0] = v0 + off;
            if (c + 1 < W) base[c + 1] = v1 + off;
            if (c + 2 < W) base[c + 2] = v2 + off;
            if (c + 3 < W) base[c + 3] = v3 + off;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    for (int r = warp; r < nr; r += nwarps) {
        const uint8_t* src = img + (size_t)(r0 + r) * pitch;
        uint32_t* dst = rp + (size_t)r * W;
        uint32_t carry = 0;
        for (int c0 = 0; c0 < W; c0 += 128) {
            int c = c0 + lane * 4;
            uint32_t v0 = c + 0 < W ? src[c + 0] : 0, v1 = c + 1 < W ? src[c + 1] : 0;
            uint32_t v2 = c + 2 < W ? src[c + 2] : 0, v3 = c + 3 < W ? src[c + 3] : 0;
            v1 += v0; v2 += v1; v3 += v2;
            uint32_t incl = warp_incl_scan(v3, lane);
            uint32_t off = carry + incl - v3;
            if (c + 0 < W) dst[c + 0] = v0 + off;
            if (c + 1 < W) dst[c + 1] = v1 + off;
            if (c + 2 < W) dst[c + 2] = v2 + off;
            if (c + 3 < W) dst[c + 3] = v3 + off;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    // (c) walk down the band: ii[y+1][x+1] = base[x] + sum_{rows <= y in band} rp[row][x]
    for (int c = tid; c < W; c += nt) {
        uint32_t acc = base[c];
        for (int r = 0; r < nr; r++) {
            acc += rp[(size_t)r * W + c];
            ii[(size_t)(r0 + r + 1) * iip + c + 1] = __uint2float_rn(acc);
        }
    }
    // zero column 0 of the band rows; band 0 also zeroes row 0
    for (int r = tid; r < nr; r += nt) ii[(size_t)(r0 + r + 1) * iip] = 0.0f;
    if (b == 0)
        for (int c = tid; c <= W; c += nt) ii[c] = 0.0f;
}

// ------------------------------------------------------------------------------------------
// Single-launch frame preparation (default path): integral image of all bands in ONE kernel, plus (optionally)
// the MDCH joint-bin image on extra CTAs of the same launch.
//   CTA b < n_bands : rows [b*BH, b*BH+BH): warp-per-row shuffle prefix -> shared memory; thread-per-column walk
//                     turns it into the band-local integral; the band's last row (= band column totals, already
//                     prefixed over x) is published to bandtot[b][*] with a release flag (flag value = frame id,
//                     so the flags never need a reset); the CTA then acquires the flags of all bands above,
//                     sums their totals per column (exact u32) and stores f32(base + local) with coalesced rows.
//                     The frame is read from HBM exactly once and the table written once.
//   CTA b >= n_bands: bin image, 4 pixels per thread (uchar4 loads, one 8-byte store).
// Blocks wait on one another; see the ticket scheme at the top of the kernel for why this cannot deadlock.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#ifdef MCT_PREP_DEBUG
__device__ unsigned long long g_prep_dbg[256 * 8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PREP_STAMP(k) do { if (threadIdx.x == 0) g_prep_dbg[s_ticket * 8 + (k)] = gtime(); } while (0)
#else
#define PREP_STAMP(k)
#endif

// flag words of k_frame_prep, one per 128-byte line: [256] band flags, [32] group flags, the ticket and the done counter
#define PREP_GROUP 8
#define PREP_CPT 4                // columns per thread whose loads are in flight together
#define PREP_BAND_FLAG(b) ((b) * 32)
#define PREP_GROUP_FLAG(g) ((256 + (g)) * 32)
#define PREP_TICKET (288 * 32)
#define PREP_DONE (289 * 32)
#define PREP_FLAG_WORDS (290 * 32)
struct PrepArgs {
    const uint8_t* gray; int W, H, pitch, BH, n_bands;
    uint32_t* bandtot; uint32_t* grptot; unsigned* flags; unsigned epoch; float* ii; int iip;
    const uint8_t* c0; const uint8_t* c1; const uint8_t* c2; int nb; int bw_shift; uint16_t* binimg; int n_binblocks;
};

__global__ void __launch_bounds__(512) k_frame_prep(const PrepArgs a)
{
    extern __shared__ uint32_t sm[];
    __shared__ int s_ticket;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int W = a.W, H = a.H;
    // Work is handed out in ARRIVAL order (ticket counter), so the band that gets ticket b knows that every band
    // < b is held by a CTA that is already running: the waits below cannot deadlock whatever the dispatch order or
    // residency, and no cooperative launch is needed.  The last CTA to finish resets the two counters.
    if (tid == 0) s_ticket = (int)atomicAdd(a.flags + PREP_TICKET, 1u);
    __syncthreads();
    const int ticket = s_ticket;
    PREP_STAMP(0);
    if (ticket >= a.n_bands) {
        // ---- joint-bin image (MultiDimensionalColorHistogram.cpp:96-104); bw = 256/nb is a power of two here
        const int q = ticket - a.n_bands;
        const int groups = (W >> 2) * H;                                  // launch_integral guarantees W % 4 == 0, pitch % 4 == 0
        const int wq = W >> 2, nb = a.nb, sh = a.bw_shift;
#pragma unroll 4
        for (int g = q * nt + tid; g < groups; g += a.n_binblocks * nt) {
            const int y = g / wq, x = (g - y * wq) << 2;
            const size_t o = (size_t)y * a.pitch + x;
            const uchar4 r = *reinterpret_cast<const uchar4*>(a.c0 + o);
            const uchar4 gg = *reinterpret_cast<const uchar4*>(a.c1 + o);
            const uchar4 bb = *reinterpret_cast<const uchar4*>(a.c2 + o);
            const unsigned b0 = (r.x >> sh) + nb * (gg.x >> sh) + nb * nb * (bb.x >> sh);
            const unsigned b1 = (r.y >> sh) + nb * (gg.y >> sh) + nb * nb * (bb.y >> sh);
            const unsigned b2 = (r.z >> sh) + nb * (gg.z >> sh) + nb * nb * (bb.z >> sh);
            const unsigned b3 = (r.w >> sh) + nb * (gg.w >> sh) + nb * nb * (bb.w >> sh);
            *reinterpret_cast<uint2*>(a.binimg + (size_t)y * W + x) = make_uint2(b0 | (b1 << 16), b2 | (b3 << 16));
        }
        __syncthreads();
        PREP_STAMP(4);
        if (tid == 0 && atomicAdd(a.flags + PREP_DONE, 1u) == gridDim.x - 1) { a.flags[PREP_TICKET] = 0; a.flags[PREP_DONE] = 0; }
        return;
    }
    const int b = ticket, BH = a.BH;
    const int r0 = b * BH, nr = min(BH, H - r0);
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    uint32_t* rp = sm;                                                    // [BH][W]
    const bool vec = ((a.pitch | (int)(size_t)a.gray) & 3) == 0;
    // (1) row prefixes: ALL chunk loads of a row are issued before the first scan (the frame comes from HBM: one
    //     latency per row instead of one per 128-pixel chunk)
    for (int r = warp; r < nr; r += nwarps) {
        const uint8_t* src = a.gray + (size_t)(r0 + r) * a.pitch;
        uint32_t* dst = rp + (size_t)r * W;
        uchar4 px[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int c = k * 128 + lane * 4;
            px[k] = make_uchar4(0, 0, 0, 0);
            if (c < W) {
                if (c + 3 < W && vec) px[k] = *reinterpret_cast<const uchar4*>(src + c);
                else {
                    px[k].x = src[c];
                    if (c + 1 < W) px[k].y = src[c + 1];
                    if (c + 2 < W) px[k].z = src[c + 2];
                    if (c + 3 < W) px[k].w = src[c + 3];
                }
            }
        }
        uint32_t carry = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int c = k * 128 + lane * 4;
            if (k * 128 < W) {
                uint32_t v0 = px[k].x, v1 = px[k].y, v2 = px[k].z, v3 = px[k].w;
                v1 += v0; v2 += v1; v3 += v2;
                const uint32_t incl = warp_incl_scan(v3, lane);
                const uint32_t off = carry + incl - v3;
                if (c + 3 < W && (W & 3) == 0) *reinterpret_cast<uint4*>(dst + c) = make_uint4(v0 + off, v1 + off, v2 + off, v3 + off);
                else {
                    if (c + 0 < W) dst[c + 0] = v0 + off;
                    if (c + 1 < W) dst[c + 1] = v1 + off;
                    if (c + 2 < W) dst[c + 2] = v2 + off;
                    if (c + 3 < W) dst[c + 3] = v3 + off;
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        // frames wider than 2048 pixels: remaining chunks the simple way
        for (int c0 = 2048; c0 < W; c0 += 128) {
            const int c = c0 + lane * 4;
            uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;
            if (c + 0 < W) v0 = src[c + 0];
            if (c + 1 < W) v1 = src[c + 1];
            if (c + 2 < W) v2 = src[c + 2];
            if (c + 3 < W) v3 = src[c + 3];
            v1 += v0; v2 += v1; v3 += v2;
            const uint32_t incl = warp_incl_scan(v3, lane);
            const uint32_t off = carry + incl - v3;
            if (c + 0 < W) dst[c + 0] = v0 + off;
            if (c + 1 < W) dst[c + 1] = v1 + off;
            if (c + 2 < W) dst[c + 2] = v2 + off;
            if (c + 3 < W) dst[c + 3] = v3 + off;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    PREP_STAMP(1);
    // (2) band-local integral in place + publish the band totals
    uint32_t* tot = a.bandtot + (size_t)b * W;
    for (int c = tid; c < W; c += nt) {
        uint32_t acc = 0;
        for (int r = 0; r < nr; r++) { acc += rp[(size_t)r * W + c]; rp[(size_t)r * W + c] = acc; }
        tot[c] = acc;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release_u32(a.flags + PREP_BAND_FLAG(b), a.epoch);
    PREP_STAMP(2);
    // (3) Two levels instead of "every band sums every band above it" (B^2 / 2 rows of totals read from L2, 59 dependent-ish loads
    //     for the last band at VGA, 67 at 1080p): bands form groups of PREP_GROUP; the LAST band of a group adds up the group's
    //     totals and publishes them; a band's base = the totals of the groups in front of its own + the totals of the bands in
    //     front of it inside its group: at most (PREP_GROUP - 1) + n_groups rows.  Every wait is for a lower ticket (bands in
    //     front, leaders of groups in front), so the arrival-order argument above still holds.  Flags live 128 bytes apart and
    //     are polled by one lane each with a back-off: the former all-threads spin on two cache lines delayed the very stores it
    //     was waiting for (5 us between the last publication and the first waiter getting through at VGA).
    const int g = b / PREP_GROUP, first = g * PREP_GROUP, last = min(first + PREP_GROUP, a.n_bands) - 1;
    if (tid < b - first) { while (ld_acquire_u32(a.flags + PREP_BAND_FLAG(first + tid)) != a.epoch) __nanosleep(40); }
    __syncthreads();
    if (b == last) {
        uint32_t* gt = a.grptot + (size_t)g * W;
        // (all loads of a thread's columns are issued before any is used: one round trip to L2 per wait level, not one per column)
        for (int c0 = tid; c0 < W; c0 += PREP_CPT * nt) {
            uint32_t v[PREP_CPT][PREP_GROUP - 1];
#pragma unroll
            for (int k = 0; k < PREP_CPT; k++)
#pragma unroll
                for (int bb = 0; bb < PREP_GROUP - 1; bb++) {
                    const int c = c0 + k * nt;
                    v[k][bb] = (c < W && first + bb < b) ? __ldcg(a.bandtot + (size_t)(first + bb) * W + c) : 0u;
                }
#pragma unroll
            for (int k = 0; k < PREP_CPT; k++) {
                const int c = c0 + k * nt;
                if (c >= W) continue;
                uint32_t acc = rp[(size_t)(nr - 1) * W + c];              // own total
#pragma unroll
                for (int bb = 0; bb < PREP_GROUP - 1; bb++) acc += v[k][bb];
                gt[c] = acc;
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_u32(a.flags + PREP_GROUP_FLAG(g), a.epoch);
    }
    for (int t = tid; t < g; t += nt) { while (ld_acquire_u32(a.flags + PREP_GROUP_FLAG(t)) != a.epoch) __nanosleep(40); }
    __syncthreads();
    PREP_STAMP(3);
    // (4) store f32(base + local); a warp writes 32 consecutive columns of a row
    float* ii = a.ii;
    const int iip = a.iip;
    for (int c0 = tid; c0 < W; c0 += PREP_CPT * nt) {
        uint32_t base[PREP_CPT];
#pragma unroll
        for (int k = 0; k < PREP_CPT; k++) {
            const int c = c0 + k * nt;
            uint32_t acc = 0;
            if (c < W) {
#pragma unroll 8
                for (int gg = 0; gg < g; gg++) acc += __ldcg(a.grptot + (size_t)gg * W + c);
#pragma unroll
                for (int bb = 0; bb < PREP_GROUP - 1; bb++) if (first + bb < b) acc += __ldcg(a.bandtot + (size_t)(first + bb) * W + c);
            }
            base[k] = acc;
        }
#pragma unroll
        for (int k = 0; k < PREP_CPT; k++) {
            const int c = c0 + k * nt;
            if (c >= W) continue;
            for (int r = 0; r < nr; r++)
                ii[(size_t)(r0 + r + 1) * iip + c + 1] = __uint2float_rn(base[k] + rp[(size_t)r * W + c]);
        }
    }
    for (int r = tid; r < nr; r += nt) ii[(size_t)(r0 + r + 1) * iip] = 0.0f;
    if (b == 0)
        for (int c = tid; c <= W; c += nt) ii[c] = 0.0f;
    __syncthreads();
    PREP_STAMP(4);
    if (tid == 0 && atomicAdd(a.flags + PREP_DONE, 1u) == gridDim.x - 1) { a.flags[PREP_TICKET] = 0; a.flags[PREP_DONE] = 0; }
}

static int launch_integral_two_pass(mct_ctx* ctx, const PrepTarget& t, cudaStream_t st);

int launch_frame_prep(mct_ctx* ctx, const PrepTarget& t, cudaStream_t st, int* with_bins_out)
{
    const int W = t.W, H = t.H;
    *with_bins_out = 0;
    const int s_sms = ctx->sms;
    // band height: enough bands to spread over the SMs; BH*W u32 must fit in shared memory
    int BH = 8;
    while (ceil_div(H, BH) > s_sms / 2 && BH < 64) BH <<= 1;
    const size_t smem = (size_t)BH * W * sizeof(uint32_t);
    const int n_bands = ceil_div(H, BH);
    if (smem > 200 * 1024 || n_bands > 256) return launch_integral_two_pass(ctx, t, st);
    // the bin image rides along when an MDCH classifier has used this ctx before and the planes allow 4-pixel vectors
    const int cnb = ctx->binimg_nb;
    int with_bins = cnb > 0 && t.binimg && t.c0 && t.c1 && t.c2 && (W & 3) == 0 && (t.pitch & 3) == 0 &&
                    ((((size_t)t.c0) | ((size_t)t.c1) | ((size_t)t.c2)) & 3) == 0 && (256 % cnb) == 0 &&
                    ((256 / cnb) & (256 / cnb - 1)) == 0;
    if (!ctx->prep_flags) {
        MCT_CUDA(ctx, cudaMalloc(&ctx->prep_flags, PREP_FLAG_WORDS * sizeof(unsigned)));
        MCT_CUDA(ctx, cudaMemsetAsync(ctx->prep_flags, 0, PREP_FLAG_WORDS * sizeof(unsigned), st));
    }
    const int n_groups = ceil_div(n_bands, PREP_GROUP);
    size_t need = (size_t)(n_bands + n_groups) * W;                               // band totals, then group totals
    if (need > ctx->bandsum_cap) {
        if (ctx->bandsum) { MCT_CUDA(ctx, cudaStreamSynchronize(st)); cudaFree(ctx->bandsum); }
        MCT_CUDA(ctx, cudaMalloc(&ctx->bandsum, need * sizeof(uint32_t)));
        ctx->bandsum_cap = need;
    }
    PrepArgs a;
    a.gray = t.gray; a.W = W; a.H = H; a.pitch = t.pitch; a.BH = BH; a.n_bands = n_bands;
    a.bandtot = ctx->bandsum; a.grptot = ctx->bandsum + (size_t)n_bands * W; a.flags = ctx->prep_flags;
    a.epoch = ++ctx->prep_epoch; a.ii = t.ii_base + MCT_II_LEAD; a.iip = t.ii_pitch;
    a.c0 = t.c0; a.c1 = t.c1; a.c2 = t.c2; a.nb = cnb; a.bw_shift = 0; a.binimg = t.binimg;
    a.n_binblocks = 0;
    if (with_bins) {
        const int bw = 256 / cnb;
        while ((1 << a.bw_shift) < bw) a.bw_shift++;
        a.n_binblocks = s_sms - n_bands;
        if (a.n_binblocks > 64) a.n_binblocks = 64;
        if (a.n_binblocks < 1) { with_bins = 0; a.n_binblocks = 0; }
    }
    static size_t s_attr[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_frame_prep, s_attr, smem, 0)) return rc_;
    const int nt = 32 * (BH < 4 ? 4 : (BH > 16 ? 16 : BH));
    MCT_LAUNCH_ON(ctx, st, "k_frame_prep", k_frame_prep<<<n_bands + a.n_binblocks, nt, smem, st>>>(a));
    *with_bins_out = with_bins;
    return MCT_OK;
}

static int launch_integral_two_pass(mct_ctx* ctx, const PrepTarget& t, cudaStream_t st)
{
    const int W = t.W, H = t.H;
    // band height: (BH + 1) * W u32 must fit in ~160 KB of shared memory
    int BH = (160 * 1024 / 4 - W) / W;
    if (BH > 32) BH = 32;
    if (BH < 1) return mct_fail(ctx, MCT_E_UNSUPPORTED, "frame too wide for the integral kernel%s (W=%d)", "", W);
    int nb = ceil_div(H, BH);
    size_t need = (size_t)nb * W;
    if (need > ctx->bandsum_cap) {
        if (ctx->bandsum) { MCT_CUDA(ctx, cudaStreamSynchronize(st)); cudaFree(ctx->bandsum); }
        MCT_CUDA(ctx, cudaMalloc(&ctx->bandsum, need * sizeof(uint32_t)));
        ctx->bandsum_cap = need;
    }
    dim3 g1(ceil_div(W, 128), nb);
    MCT_LAUNCH_ON(ctx, st, "k_band_colsum", k_band_colsum<<<g1, 128, 0, st>>>(t.gray, W, H, t.pitch, BH, ctx->bandsum));
    size_t smem = (size_t)(BH + 1) * W * sizeof(uint32_t);
    static size_t s_attr[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_integral_band, s_attr, smem, 0)) return rc_;
    int nt = W >= 1024 ? 1024 : round_up(W, 32);
    if (nt < 128) nt = 128;
    MCT_LAUNCH_ON(ctx, st, "k_integral_band", k_integral_band<<<nb, nt, smem, st>>>(t.gray, W, H, t.pitch, BH, ctx->bandsum,
                                                                                t.ii_base + MCT_II_LEAD, t.ii_pitch));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// Frame pre-processing from the packed BGR image (Camera.cpp:449-494):
//   planes   Matrix.cpp:70-100     channel 0 = R, 1 = G, 2 = B
//   gray     Matrix.cpp:518-533    cvCvtColor BGR2GRAY = (R*4899 + G*9617 + B*1868 + 8192) >> 14
//   HSV      Matrix.cpp:535-580    cvCvtColor BGR2HSV, 8-bit: fixed-point tables (hsv_shift 12, H in [0,180))
// Integer arithmetic only: bit-exact against cv2 (tests/golden/cv2_gray.json, cv2_hsv.json via the oracle).
__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];
__global__ void __launch_bounds__(256)
k_bgr_unpack(const uint8_t* __restrict__ bgr, int spitch, int W, int H, int use_hsv, uint8_t* __restrict__ gray,
             uint8_t* __restrict__ c0, uint8_t* __restrict__ c1, uint8_t* __restrict__ c2, int dpitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const uint8_t* p = bgr + (size_t)y * spitch + 3 * x;
    const int b = p[0], g = p[1], r = p[2];
    const size_t o = (size_t)y * dpitch + x;
    gray[o] = (uint8_t)((r * 4899 + g * 9617 + b * 1868 + 8192) >> 14);
    if (!use_hsv) { c0[o] = (uint8_t)r; c1[o] = (uint8_t)g; c2[o] = (uint8_t)b; return; }
    const int v = max(max(b, g), r), vmin = min(min(b, g), r), diff = v - vmin;
    const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
    const int s = (diff * c_sdiv[v] + (1 << 11)) >> 12;
    int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    h = (h * c_hdiv[diff] + (1 << 11)) >> 12;
    h += h < 0 ? 180 : 0;
    c0[o] = (uint8_t)h; c1[o] = (uint8_t)s; c2[o] = (uint8_t)v;
}
int launch_bgr_unpack(mct_ctx* ctx, const uint8_t* d_bgr, int spitch, int W, int H, int use_hsv, uint8_t* const planes[4], int dpitch,
                      cudaStream_t st)
{
    static int s_tables_dev[MCT_MAX_DEV];             // the __constant__ tables are per device
    int& s_tables = s_tables_dev[ctx->device];
    if (!s_tables) {
        int sdiv[256], hdiv[256];
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; i++) {       // saturate_cast<int>(double) == cvRound (half to even)
            sdiv[i] = (int)lrint((double)(255 << 12) / (1.0 * i));
            hdiv[i] = (int)lrint((double)(180 << 12) / (6.0 * i));
        }
        MCT_CUDA(ctx, cudaMemcpyToSymbolAsync(c_sdiv, sdiv, sizeof(sdiv), 0, cudaMemcpyHostToDevice, st));
        MCT_CUDA(ctx, cudaMemcpyToSymbolAsync(c_hdiv, hdiv, sizeof(hdiv), 0, cudaMemcpyHostToDevice, st));
        MCT_CUDA(ctx, cudaStreamSynchronize(st));          // the host tables live on this stack frame
        s_tables = 1;
    }
    dim3 g(ceil_div(W, 256), H);
    MCT_LAUNCH_ON(ctx, st, "k_bgr_unpack", k_bgr_unpack<<<g, 256, 0, st>>>(d_bgr, spitch, W, H, use_hsv, planes[0], planes[1], planes[2], planes[3], dpitch));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void k_sum_rects(const float* __restrict__ ii, int iip, int W, int H, const int* __restrict__ rects,
                            int n, float* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = sum_rect_dev(ii, iip, W, H, rects[4 * i + 0], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]);
}
int launch_sum_rects(mct_ctx* ctx, const int* d_rects, int n, float* d_out)
{
    if (n <= 0) return MCT_OK;
    MCT_LAUNCH(ctx, "k_sum_rects", k_sum_rects<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(ctx->ii_base + MCT_II_LEAD, ctx->ii_pitch, ctx->W, ctx->H,
                                                          d_rects, n, d_out));
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// Joint colour bin per pixel (MultiDimensionalColorHistogram.cpp:96-104):
//   bin = floor(c0/bw) + nb*floor(c1/bw) + nb^2*floor(c2/bw), bw = 256/nb (integer division).
// It depends on the frame only, so it is computed once per frame instead of once per particle pixel.
__global__ void k_bin_image(const uint8_t* __restrict__ c0, const uint8_t* __restrict__ c1, const uint8_t* __restrict__ c2,
                            int W, int H, int pitch, int nb, float bw, uint16_t* __restrict__ out)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W || y >= H) return;
    size_t o = (size_t)y * pitch + x;
    unsigned rb = (unsigned)floorf(__fdiv_rn((float)c0[o], bw));
    unsigned gb = (unsigned)floorf(__fdiv_rn((float)c1[o], bw));
    unsigned bb = (unsigned)floorf(__fdiv_rn((float)c2[o], bw));
    out[(size_t)y * W + x] = (uint16_t)(rb + nb * gb + nb * nb * bb);
}
int launch_bin_image(mct_ctx* ctx, int nb)
{
    if (ctx->binimg_frame == ctx->frame_id && ctx->binimg_nb == nb) return MCT_OK;
    if (!ctx->c0 || !ctx->c1 || !ctx->c2) return mct_fail(ctx, MCT_E_STATE, "colour planes not set%s", "");
    size_t need = (size_t)ctx->W * ctx->H;
    if (need > ctx->binimg_cap) {
        if (ctx->binimg) { MCT_CUDA(ctx, cudaDeviceSynchronize()); cudaFree(ctx->binimg); }
        MCT_CUDA(ctx, cudaMalloc(&ctx->binimg, need * sizeof(uint16_t)));
        ctx->binimg_cap = need;
    }
    dim3 g(ceil_div(ctx->W, 256), ctx->H);
    MCT_LAUNCH(ctx, "k_bin_image", k_bin_image<<<g, 256, 0, ctx->stream>>>(ctx->c0, ctx->c1, ctx->c2, ctx->W, ctx->H, ctx->pitch, nb,
                                            (float)(256 / nb), ctx->binimg));
    ctx->binimg_nb = nb;
    ctx->binimg_frame = ctx->frame_id;
    return MCT_OK;
}
