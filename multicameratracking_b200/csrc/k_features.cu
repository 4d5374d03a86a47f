// k_features.cu -- feature extraction over box lists (FeatureVector::Compute family).
//   k_haar_matrix : HaarFeatureVector::Compute (HaarFeatureVector.cpp:48-79) -> rows [0, n_haar)
//   k_mdch        : MultiDimensionalColorHistogram::Compute (MultiDimensionalColorHistogram.cpp:32-127)
//   k_cch         : CultureColorHistogram::Compute (CultureColorHistogram.cpp:25-111)
// Output is feature-major [F][ld] f32 (SampleSet.h:41-43), coalesced over the box index.
#include "mct_internal.cuh"

// ------------------------------------------------------------------------------------------
// Haar: thread per box, blockIdx.y tiles the features (table reads are warp-uniform broadcasts).
#define HAAR_FT 16
__global__ void __launch_bounds__(256)
k_haar_matrix(const float* __restrict__ ii, int iip, int W, int H,
              const int4* __restrict__ rects, const float* __restrict__ wts, const int* __restrict__ nrect, int n_haar,
              const int* __restrict__ col, const int* __restrict__ row, const float* __restrict__ sx,
              const float* __restrict__ sy, int n, float* __restrict__ out, int ld)
{
    __shared__ int4 s_r[HAAR_FT * MCT_MAX_RECTS];
    __shared__ float s_w[HAAR_FT * MCT_MAX_RECTS];
    __shared__ int s_n[HAAR_FT];
    const int f0 = blockIdx.y * HAAR_FT;
    const int nf = min(HAAR_FT, n_haar - f0);
    for (int t = threadIdx.x; t < nf * MCT_MAX_RECTS; t += blockDim.x) {
        s_r[t] = rects[(size_t)f0 * MCT_MAX_RECTS + t];
        s_w[t] = wts[(size_t)f0 * MCT_MAX_RECTS + t];
    }
    for (int t = threadIdx.x; t < nf; t += blockDim.x) s_n[t] = nrect[f0 + t];
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = col[i], r = row[i];
    const float fx = sx[i], fy = sy[i];
    for (int f = 0; f < nf; f++)
        out[(size_t)(f0 + f) * ld + i] =
            haar_eval_dev(ii, iip, W, H, s_r + f * MCT_MAX_RECTS, s_w + f * MCT_MAX_RECTS, s_n[f], c, r, fx, fy);
}

// ------------------------------------------------------------------------------------------
// MDCH.  One warp per box, a CTA handles a tile of MDCH_TILE boxes so that the [bin][box] output can be
// written with coalesced rows.  Per box:
//   - the per-pixel weight exp(-(dx^2/(2 varX) + dy^2/(2 varY))) follows the reference's promotion
//     (double quotient sum rounded to f32, then expf);
//   - weights are accumulated in shared memory as u32 fixed point (scale 2^shift) with native
//     ATOMS.ADD -- float shared atomics are CAS loops on sm_100a and order-nondeterministic; fixed
//     point is exact-order-independent, so the histogram is run-to-run bit-stable.  Weights lie in
//     [e^-1, 1], so the relative quantisation error per pixel is <= 2^-20/0.37 = 2.6e-6;
//   - boxes with more than MCT_MDCH_SEG_PIX pixels are processed in row segments, each flushed to an
//     f32 accumulator, so the u32 bins cannot overflow at any scale;
//   - normalizeVec (Public.h:270-275): bins / sum(bins).
#define MDCH_TILE 32
#define MDCH_WARPS 8
__global__ void __launch_bounds__(MDCH_WARPS * 32)
k_mdch(const uint16_t* __restrict__ binimg, int W, int H, int nb, int w0, int h0,
       const int* __restrict__ col, const int* __restrict__ row, const float* __restrict__ sx,
       const float* __restrict__ sy, const int* __restrict__ valid, int n,
       const int* __restrict__ outbins, int n_out, float* __restrict__ out, int ld)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const int dim = nb * nb * nb;
    uint32_t* hist_all = (uint32_t*)smraw;                         // [MDCH_WARPS][dim]
    float* acc_all = (float*)(hist_all + MDCH_WARPS * dim);        // [MDCH_WARPS][dim] f32 accumulators
    float* tile = acc_all + MDCH_WARPS * dim;                      // [n_out][MDCH_TILE+1]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* hist = hist_all + warp * dim;
    float* acc = acc_all + warp * dim;
    const int i0 = blockIdx.x * MDCH_TILE;

    for (int bi = warp; bi < MDCH_TILE; bi += MDCH_WARPS) {
        const int i = i0 + bi;
        if (i >= n) break;                                          // warp-uniform
        if (valid && !valid[i]) {
            for (int o = lane; o < n_out; o += 32) tile[o * (MDCH_TILE + 1) + bi] = 0.0f;
            continue;
        }
        const int c0 = col[i], r0 = row[i];
        const float fx = sx[i], fy = sy[i];
        const int rows = __float2int_rn(__fmul_rn((float)h0, fy));
        const int cols = __float2int_rn(__fmul_rn((float)w0, fx));
        const float ncols = (float)cols, nrows = (float)rows;
        const float hx = __fdiv_rn(ncols, 2.0f), hy = __fdiv_rn(nrows, 2.0f);
        const float varX = __fmul_rn(hx, hx), varY = __fmul_rn(hy, hy);      // exact: half-integers squared
        const float cX = __fadd_rn((float)c0, hx), cY = __fadd_rn((float)r0, hy);
        const double ivx = 1.0 / (double)__fmul_rn(2.0f, varX);
        const double ivy = 1.0 / (double)__fmul_rn(2.0f, varY);
        for (int b = lane; b < dim; b += 32) { hist[b] = 0u; acc[b] = 0.0f; }
        __syncwarp();
        const int rows_per_seg = max(1, MCT_MDCH_SEG_PIX / max(cols, 1));
        for (int rs = 0; rs < rows; rs += rows_per_seg) {
            const int re = min(rows, rs + rows_per_seg);
            const int npix = (re - rs) * cols;
            int shift = __clz(max(npix, 1));
            if (shift > MCT_FIX_SHIFT_MAX) shift = MCT_FIX_SHIFT_MAX;
            const float scale = (float)(1u << shift);
            // lanes stride over the flattened segment; (r, c) tracked incrementally
            int c = lane, r = rs;
            while (c >= cols && r < re) { c -= cols; r++; }
            while (r < re) {
                const int yy = clampi(r0 + r, 0, H - 1), xx = clampi(c0 + c, 0, W - 1);
                const unsigned bin = binimg[(size_t)yy * W + xx];
                const double dx = (double)__fsub_rn(cX, (float)(c0 + c));
                const double dy = (double)__fsub_rn(cY, (float)(r0 + r));
                const float wd = (float)(dx * dx * ivx + dy * dy * ivy);
                const float wgt = expf(-wd);
                atomicAdd(&hist[bin], __float2uint_rn(__fmul_rn(wgt, scale)));
                c += 32;
                while (c >= cols && r < re) { c -= cols; r++; }
            }
            __syncwarp();
            const float inv = __fdiv_rn(1.0f, scale);
            for (int b = lane; b < dim; b += 32) {
                acc[b] = __fadd_rn(acc[b], __fmul_rn((float)hist[b], inv));
                hist[b] = 0u;
            }
            __syncwarp();
        }
        float s = 0.0f;
        for (int b = lane; b < dim; b += 32) s += acc[b];
        s = warp_sum_f(s);
        for (int o = lane; o < n_out; o += 32) {
            const int b = outbins ? outbins[o] : o;
            tile[o * (MDCH_TILE + 1) + bi] = __fdiv_rn(acc[b], s);
        }
        __syncwarp();
    }
    __syncthreads();
    // coalesced write: each output row gets MDCH_TILE consecutive boxes
    const int nbx = min(MDCH_TILE, n - i0);
    for (int t = threadIdx.x; t < n_out * MDCH_TILE; t += blockDim.x) {
        const int o = t / MDCH_TILE, bi = t % MDCH_TILE;
        if (bi < nbx) out[(size_t)o * ld + i0 + bi] = tile[o * (MDCH_TILE + 1) + bi];
    }
}

// ------------------------------------------------------------------------------------------
// CCH: 11 unnormalised counts per box.  mode 0 = reference_exact (every pixel -> bin 0 because the
// distance is always taken to HC[0], CultureColorHistogram.cpp:88-90), mode 1 = intended nearest hue
// centroid (first minimum).  Integer counts -> exact.
__constant__ int c_HC[11] = {5, 206, 181, 160, 81, 46, 28, 20, 110, 121, 133};
__global__ void __launch_bounds__(256)
k_cch(const uint8_t* __restrict__ hue, int pitch, int W, int H, int w0, int h0,
      const int* __restrict__ col, const int* __restrict__ row, const float* __restrict__ sx,
      const float* __restrict__ sy, const int* __restrict__ valid, int n, int mode, float* __restrict__ out, int ld)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    int cnt[11];
#pragma unroll
    for (int k = 0; k < 11; k++) cnt[k] = 0;
    if (!valid || valid[i]) {
        const int rows = __float2int_rn(__fmul_rn((float)h0, sy[i]));
        const int cols = __float2int_rn(__fmul_rn((float)w0, sx[i]));
        if (mode == 0) {
            if (lane == 0) cnt[0] = rows * cols;
        } else {
            const int c0 = col[i], r0 = row[i];
            for (int p = lane; p < rows * cols; p += 32) {
                const int r = p / cols, c = p - r * cols;
                const int yy = clampi(r0 + r, 0, H - 1), xx = clampi(c0 + c, 0, W - 1);
                const int Hh = hue[(size_t)yy * pitch + xx];
                int best = 999999, ind = 0;
#pragma unroll
                for (int k = 0; k < 11; k++) {
                    int d = abs(Hh - c_HC[k]);
                    if (best > d) { best = d; ind = k; }
                }
#pragma unroll
                for (int k = 0; k < 11; k++) cnt[k] += (ind == k);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 11; k++) {
        int v = cnt[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) out[(size_t)k * ld + i] = (float)v;
    }
}

// ------------------------------------------------------------------------------------------
static int launch_haar_rows(mct_clf* clf, const int* d_col, const int* d_row, const float* d_sx, const float* d_sy,
                            int n, float* d_out, int ld)
{
    mct_ctx* ctx = clf->ctx;
    dim3 g(ceil_div(n, 256), ceil_div(clf->n_haar, HAAR_FT));
    MCT_LAUNCH(ctx, "k_haar_matrix", k_haar_matrix<<<g, 256, 0, ctx->stream>>>(ctx->ii_base + MCT_II_LEAD, ctx->ii_pitch, ctx->W, ctx->H, clf->d_rects,
                                              clf->d_wts, clf->d_nrect, clf->n_haar, d_col, d_row, d_sx, d_sy, n, d_out, ld));
    return MCT_OK;
}

static size_t mdch_smem(int dim, int n_out)
{
    return (size_t)MDCH_WARPS * dim * (sizeof(uint32_t) + sizeof(float)) + (size_t)n_out * (MDCH_TILE + 1) * sizeof(float);
}

static int launch_mdch_rows(mct_clf* clf, const int* d_col, const int* d_row, const float* d_sx, const float* d_sy,
                            const int* d_valid, int n, const int* d_outbins, int n_out, float* d_out, int ld)
{
    mct_ctx* ctx = clf->ctx;
    if (n_out <= 0) return MCT_OK;
    int rc = launch_bin_image(ctx, clf->p.n_bins);
    if (rc) return rc;
    const int dim = clf->n_color;
    size_t smem = mdch_smem(dim, n_out);
    static size_t s_attr = 0;
    if (smem > s_attr) {
        MCT_CUDA(ctx, cudaFuncSetAttribute(k_mdch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s_attr = smem;
    }
    MCT_LAUNCH(ctx, "k_mdch", k_mdch<<<ceil_div(n, MDCH_TILE), MDCH_WARPS * 32, smem, ctx->stream>>>(
        ctx->binimg, ctx->W, ctx->H, clf->p.n_bins, clf->p.w0, clf->p.h0, d_col, d_row, d_sx, d_sy, d_valid, n,
        d_outbins, n_out, d_out, ld));
    return MCT_OK;
}

// full [F][ld] matrix (FeatureVector::Compute)
int mct_feature_matrix_full(mct_clf* clf, const int* d_col, const int* d_row, const float* d_sx, const float* d_sy,
                            int n, float* d_out, int ld)
{
    mct_ctx* ctx = clf->ctx;
    if (n <= 0) return MCT_OK;
    int rc;
    if (clf->n_haar > 0 && (rc = launch_haar_rows(clf, d_col, d_row, d_sx, d_sy, n, d_out, ld))) return rc;
    if (clf->p.feature_type == MCT_FEAT_MDCH || clf->p.feature_type == MCT_FEAT_HAAR_MDCH) {
        rc = launch_mdch_rows(clf, d_col, d_row, d_sx, d_sy, nullptr, n, nullptr, clf->n_color,
                              d_out + (size_t)clf->n_haar * ld, ld);
        if (rc) return rc;
    } else if (clf->p.feature_type == MCT_FEAT_CCH) {
        if (!ctx->c0) return mct_fail(ctx, MCT_E_STATE, "hue plane not set%s", "");
        MCT_LAUNCH(ctx, "k_cch", k_cch<<<ceil_div(n, 8), 256, 0, ctx->stream>>>(ctx->c0, ctx->pitch, ctx->W, ctx->H, clf->p.w0, clf->p.h0, d_col,
                                                       d_row, d_sx, d_sy, nullptr, n, clf->p.cch_mode, d_out, ld));
    }
    return MCT_OK;
}

// colour rows needed by classify: only the selected bins (rows 0..n_out) for MDCH; all 11 for CCH
int mct_feature_color_rows(mct_clf* clf, const int* d_col, const int* d_row, const float* d_sx, const float* d_sy,
                           const int* d_valid, int n, const int* d_outbins, int n_out, float* d_out, int ld)
{
    mct_ctx* ctx = clf->ctx;
    if (n <= 0) return MCT_OK;
    if (clf->p.feature_type == MCT_FEAT_MDCH || clf->p.feature_type == MCT_FEAT_HAAR_MDCH)
        return launch_mdch_rows(clf, d_col, d_row, d_sx, d_sy, d_valid, n, d_outbins, n_out, d_out, ld);
    if (clf->p.feature_type == MCT_FEAT_CCH) {
        if (!ctx->c0) return mct_fail(ctx, MCT_E_STATE, "hue plane not set%s", "");
        MCT_LAUNCH(ctx, "k_cch", k_cch<<<ceil_div(n, 8), 256, 0, ctx->stream>>>(ctx->c0, ctx->pitch, ctx->W, ctx->H, clf->p.w0, clf->p.h0, d_col,
                                                       d_row, d_sx, d_sy, d_valid, n, clf->p.cch_mode, d_out, ld));
    }
    return MCT_OK;
}
