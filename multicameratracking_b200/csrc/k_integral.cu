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

int launch_integral(mct_ctx* ctx)
{
    const int W = ctx->W, H = ctx->H;
    // band height: (BH + 1) * W u32 must fit in ~160 KB of shared memory
    int BH = (160 * 1024 / 4 - W) / W;
    if (BH > 32) BH = 32;
    if (BH < 1) return mct_fail(ctx, MCT_E_UNSUPPORTED, "frame too wide for the integral kernel%s (W=%d)", "", W);
    int nb = ceil_div(H, BH);
    size_t need = (size_t)nb * W;
    if (need > ctx->bandsum_cap) {
        if (ctx->bandsum) cudaFree(ctx->bandsum);
        MCT_CUDA(ctx, cudaMalloc(&ctx->bandsum, need * sizeof(uint32_t)));
        ctx->bandsum_cap = need;
    }
    dim3 g1(ceil_div(W, 128), nb);
    MCT_LAUNCH(ctx, "k_band_colsum", k_band_colsum<<<g1, 128, 0, ctx->stream>>>(ctx->gray, W, H, ctx->pitch, BH, ctx->bandsum));
    size_t smem = (size_t)(BH + 1) * W * sizeof(uint32_t);
    static size_t s_attr = 0;
    if (smem > s_attr) {
        MCT_CUDA(ctx, cudaFuncSetAttribute(k_integral_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s_attr = smem;
    }
    int nt = W >= 1024 ? 1024 : round_up(W, 32);
    if (nt < 128) nt = 128;
    MCT_LAUNCH(ctx, "k_integral_band", k_integral_band<<<nb, nt, smem, ctx->stream>>>(ctx->gray, W, H, ctx->pitch, BH, ctx->bandsum,
                                                  ctx->ii_base + MCT_II_LEAD, ctx->ii_pitch));
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
        if (ctx->binimg) cudaFree(ctx->binimg);
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
