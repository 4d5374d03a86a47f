// k_features.cu -- feature extraction over box lists (FeatureVector::Compute family).
//   k_haar_matrix : HaarFeatureVector::Compute (HaarFeatureVector.cpp:48-79) -> rows [0, n_haar)
//   k_mdch        : MultiDimensionalColorHistogram::Compute (MultiDimensionalColorHistogram.cpp:32-127)
//   k_cch         : CultureColorHistogram::Compute (CultureColorHistogram.cpp:25-111)
// Output is feature-major [F][ld] f32 (SampleSet.h:41-43), coalesced over the box index.
#include "mct_internal.cuh"
#include <stdlib.h>

// ------------------------------------------------------------------------------------------
// Haar: thread per box, blockIdx.y tiles the features (table reads are warp-uniform broadcasts).
#define HAAR_FT 16
__device__ __forceinline__ void
haar_matrix_body(const float* __restrict__ ii, int iip, int W, int H,
              const int4* __restrict__ rects, const float* __restrict__ wts, const int* __restrict__ nrect, int n_haar,
              const int* __restrict__ col, const int* __restrict__ row, const float* __restrict__ sx,
              const float* __restrict__ sy, int n, float* __restrict__ out, int ld)
{
    __shared__ int4 s_r[HAAR_FT * MCT_MAX_RECTS];
    __shared__ float s_w[HAAR_FT * MCT_MAX_RECTS];
    __shared__ int s_n[HAAR_FT];
    const int f0 = blockIdx.y * HAAR_FT;
    const int nf = min(HAAR_FT, n_haar - f0);
    if (nf <= 0 || (int)(blockIdx.x * blockDim.x) >= n) return;       // block-uniform (batched launches use the largest grid)
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

__global__ void __launch_bounds__(256)
k_haar_matrix(const float* __restrict__ ii, int iip, int W, int H,
              const int4* __restrict__ rects, const float* __restrict__ wts, const int* __restrict__ nrect, int n_haar,
              const int* __restrict__ col, const int* __restrict__ row, const float* __restrict__ sx,
              const float* __restrict__ sy, int n, float* __restrict__ out, int ld)
{
    haar_matrix_body(ii, iip, W, H, rects, wts, nrect, n_haar, col, row, sx, sy, n, out, ld);
}
// batched form: blockIdx.z = object (training sets of all objects of a camera)
__global__ void __launch_bounds__(256)
k_haar_matrix_batch(const TrainDesc* __restrict__ descs, const float* __restrict__ ii, int iip, int W, int H)
{
    const TrainDesc& d = descs[blockIdx.z];
    haar_matrix_body(ii, iip, W, H, d.rects, d.wts, d.nrect, d.n_haar, d.col, d.row, d.sx, d.sy, d.P + d.Nn, d.featT, d.ldT);
}

// ------------------------------------------------------------------------------------------
// Tiled Haar matrix (large box lists anywhere in the frame, e.g. the N x F sweep of BASELINE configs[4]).
// k_haar_matrix is bound by the L1 wavefront rate: a warp's 32 corner loads touch up to 32 different lines.  Here the
// boxes are first binned by position into TS x TS tiles (k_tile_sort: counting sort in one CTA); one CTA per
// (tile, feature chunk) then stages the integral-image region of its tile in shared memory with TMA bulk row copies and
// gathers from there (bank conflicts instead of line misses).  Boxes whose extent exceeds the staged region (large
// scales) take the global path inside the same kernel.  Arithmetic is identical to k_haar_matrix (bit-exact).
#define HT_TS 64
#define HT_FT 64
#define HT_THREADS 256
// one-CTA counting sort (small / medium lists): counts and cursors live in shared memory
__global__ void __launch_bounds__(1024)
k_tile_sort(const int* __restrict__ col, const int* __restrict__ row, int n, int W, int H, int tiles_x, int n_tiles,
            int* __restrict__ start /* [n_tiles + 1] */, int* __restrict__ order /* [n] */)
{
    extern __shared__ int s_cnt[];                    // [n_tiles] counts, then cursors
    __shared__ int s_part[32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < n_tiles; t += nt) s_cnt[t] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const int tx = clampi(col[i], 0, W - 1) / HT_TS, ty = clampi(row[i], 0, H - 1) / HT_TS;
        atomicAdd(&s_cnt[ty * tiles_x + tx], 1);
    }
    __syncthreads();
    int base = 0;
    for (int t0 = 0; t0 < n_tiles; t0 += nt) {
        const int t = t0 + tid;
        const int v = t < n_tiles ? s_cnt[t] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (nt >> 5) ? s_part[lane] : 0, wi = w;
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
            s_part[lane] = wi - w;
        }
        __syncthreads();
        const int excl = base + s_part[warp] + inc - v;
        if (t < n_tiles) { start[t] = excl; s_cnt[t] = excl; }
        __syncthreads();
        if (tid == nt - 1) s_part[0] = excl + v;
        __syncthreads();
        base = s_part[0];
        __syncthreads();
    }
    if (tid == 0) start[n_tiles] = n;
    for (int i = tid; i < n; i += nt) {
        const int tx = clampi(col[i], 0, W - 1) / HT_TS, ty = clampi(row[i], 0, H - 1) / HT_TS;
        order[atomicAdd(&s_cnt[ty * tiles_x + tx], 1)] = i;
    }
}

__device__ __forceinline__ int tile_of(int c, int r, int W, int H, int tiles_x)
{
    return (clampi(r, 0, H - 1) / HT_TS) * tiles_x + clampi(c, 0, W - 1) / HT_TS;
}
// counting sort of the boxes by tile: count (grid over boxes) -> exclusive scan (one CTA) -> scatter (grid over boxes).
// The order inside a tile is arbitrary (atomics); every box is computed independently, so results do not depend on it.
__global__ void __launch_bounds__(256)
k_tile_count(const int* __restrict__ col, const int* __restrict__ row, int n, int W, int H, int tiles_x, int* __restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&cnt[tile_of(col[i], row[i], W, H, tiles_x)], 1);
}
__global__ void __launch_bounds__(1024)
k_tile_scan(int* __restrict__ cnt /* in: counts, out: cursors */, int n_tiles, int n, int* __restrict__ start /* [n_tiles + 1] */)
{
    __shared__ int s_part[32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    int base = 0;
    for (int t0 = 0; t0 < n_tiles; t0 += nt) {
        const int t = t0 + tid;
        const int v = t < n_tiles ? cnt[t] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (nt >> 5) ? s_part[lane] : 0, wi = w;
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
            s_part[lane] = wi - w;
        }
        __syncthreads();
        const int excl = base + s_part[warp] + inc - v;
        if (t < n_tiles) { start[t] = excl; cnt[t] = excl; }
        __syncthreads();
        if (tid == nt - 1) s_part[0] = excl + v;
        __syncthreads();
        base = s_part[0];
        __syncthreads();
    }
    if (tid == 0) start[n_tiles] = n;
}
__global__ void __launch_bounds__(256)
k_tile_scatter(const int* __restrict__ col, const int* __restrict__ row, int n, int W, int H, int tiles_x, int* __restrict__ cursor,
               int* __restrict__ order)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[atomicAdd(&cursor[tile_of(col[i], row[i], W, H, tiles_x)], 1)] = i;
}

__global__ void __launch_bounds__(HT_THREADS)
k_haar_matrix_tiled(const float* __restrict__ ii, int iip, int W, int H,
                    const int4* __restrict__ rects, const float* __restrict__ wts, const int* __restrict__ nrect, int n_haar,
                    const int* __restrict__ col, const int* __restrict__ row, const float* __restrict__ sx,
                    const float* __restrict__ sy, const int* __restrict__ start, const int* __restrict__ order, int tiles_x,
                    int ext_w, int ext_h, int mxw, int mxh, int ft, float* __restrict__ out, int ld)
{
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ int4 s_r[HT_FT * MCT_MAX_RECTS];
    __shared__ float s_w[HT_FT * MCT_MAX_RECTS];
    __shared__ int s_n[HT_FT];
    __shared__ __align__(8) uint64_t s_bar;
    const int tile = blockIdx.x;
    const int b0 = start[tile], b1 = start[tile + 1];
    if (b0 >= b1) return;                                           // empty tile (block-uniform)
    const int tid = threadIdx.x, lane = tid & 31;
    const int f0 = blockIdx.y * ft, nf = min(ft, n_haar - f0);        // ft <= HT_FT features per CTA
    if (tid == 0) { mbar_init(&s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int t = tid; t < nf * MCT_MAX_RECTS; t += blockDim.x) {
        s_r[t] = rects[(size_t)f0 * MCT_MAX_RECTS + t];
        s_w[t] = wts[(size_t)f0 * MCT_MAX_RECTS + t];
    }
    for (int t = tid; t < nf; t += blockDim.x) s_n[t] = nrect[f0 + t];
    __syncthreads();
    // region of this tile in the integral image: positions [tx*TS, tx*TS + TS) + box extent, clipped to the table
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int bx0 = tx * HT_TS, by0 = ty * HT_TS;
    const int bx1 = min(bx0 + HT_TS - 1 + ext_w, W), by1 = min(by0 + HT_TS - 1 + ext_h, H);
    const int ox = ((bx0 + 3) & ~3) - 3;                            // 16-byte aligned source rows (3-float lead of the table)
    const int RH = by1 - by0 + 1;
    const int RWc = (bx1 - ox + 1 + 3) & ~3;
    const int RWp = (RWc & 4) ? RWc : RWc + 4;
    float* region = (float*)smraw;
    if (tid < 32) {
        if (lane == 0) mbar_expect_tx(&s_bar, (unsigned)(RWc * RH * 4));
        __syncwarp();
        for (int y = lane; y < RH; y += 32)
            bulk_g2s(region + (size_t)y * RWp, ii + (size_t)(by0 + y) * iip + ox, (unsigned)(RWc * 4), &s_bar);
    }
    mbar_wait(&s_bar, 0);
    const int nbox = b1 - b0;
    if (nbox * 2 >= (int)blockDim.x) {
        // dense tile: one thread per box, its position loaded once for the whole feature chunk
        for (int j = b0 + tid; j < b1; j += blockDim.x) {
            const int i = order[j];
            const int c = col[i], r = row[i];
            const float fx = sx[i], fy = sy[i];
            // round(a*s) + round(b*s) <= (a+b)*s + 1: does every rect of this box stay inside the staged region?
            // (corner coordinates are clamped to the table, so a region that reaches the table edge always suffices)
            const int ex = (int)ceilf((float)mxw * fx) + 1, ey = (int)ceilf((float)mxh * fy) + 1;
            const bool fits = (c + ex <= bx1 || bx1 >= W) && (r + ey <= by1 || by1 >= H);
            for (int f = 0; f < nf; f++) {
                float v;
                if (fits) v = haar_eval_region(region, RWp, ox, by0, W, H, s_r + f * MCT_MAX_RECTS, s_w + f * MCT_MAX_RECTS, s_n[f], c, r, fx, fy);
                else v = haar_eval_dev(ii, iip, W, H, s_r + f * MCT_MAX_RECTS, s_w + f * MCT_MAX_RECTS, s_n[f], c, r, fx, fy);
                out[(size_t)(f0 + f) * ld + i] = v;
            }
        }
    } else {
        // sparse tile: work items = (feature, box) pairs, boxes fastest, so that every thread is busy
        const int items = nbox * nf;
        for (int item = tid; item < items; item += blockDim.x) {
            const int f = item / nbox, jb = item - f * nbox;
            const int i = order[b0 + jb];
            const int c = col[i], r = row[i];
            const float fx = sx[i], fy = sy[i];
            const int ex = (int)ceilf((float)mxw * fx) + 1, ey = (int)ceilf((float)mxh * fy) + 1;
            const bool fits = (c + ex <= bx1 || bx1 >= W) && (r + ey <= by1 || by1 >= H);
            float v;
            if (fits) v = haar_eval_region(region, RWp, ox, by0, W, H, s_r + f * MCT_MAX_RECTS, s_w + f * MCT_MAX_RECTS, s_n[f], c, r, fx, fy);
            else v = haar_eval_dev(ii, iip, W, H, s_r + f * MCT_MAX_RECTS, s_w + f * MCT_MAX_RECTS, s_n[f], c, r, fx, fy);
            out[(size_t)(f0 + f) * ld + i] = v;
        }
    }
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
#define MDCH_TILE 8
#define MDCH_WARPS 8
#define MDCH_CHUNK 1024
__device__ __forceinline__ void
mdch_body(const uint16_t* __restrict__ binimg, int W, int H, int nb, int w0, int h0,
       const int* __restrict__ col, const int* __restrict__ row, const float* __restrict__ sx,
       const float* __restrict__ sy, const int* __restrict__ valid, int n,
       const int* __restrict__ outbins, int n_out, float* __restrict__ out, int ld)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const int dim = nb * nb * nb;
    uint32_t* hist_all = (uint32_t*)smraw;                         // [MDCH_WARPS][dim]
    float* acc_all = (float*)(hist_all + MDCH_WARPS * dim);        // [MDCH_WARPS][dim] f32 accumulators
    float* tile = acc_all + MDCH_WARPS * dim;                      // [n_out][MDCH_TILE+1]
    uint16_t* buf_all = (uint16_t*)(tile + (size_t)n_out * (MDCH_TILE + 1));   // [MDCH_WARPS][MDCH_CHUNK]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* hist = hist_all + warp * dim;
    float* acc = acc_all + warp * dim;
    uint16_t* buf = buf_all + warp * MDCH_CHUNK;
    const int i0 = blockIdx.x * MDCH_TILE;
    if (i0 >= n) return;                                            // block-uniform (batched launches use the largest grid)

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
            // lanes stride over the flattened segment in chunks of MDCH_CHUNK pixels: the joint bins of a chunk are first
            // fetched into a per-warp shared buffer with independent loads (8 in flight per lane; they come from L2, and
            // issued one per iteration in front of the atomics they were what bounded this kernel), then binned.
            const unsigned magic = 0xFFFFFFFFu / (unsigned)cols + 1u;       // p / cols == umulhi(p, magic) for p < 65536
            for (int pc = 0; pc < npix; pc += MDCH_CHUNK) {
#pragma unroll 8
                for (int k = 0; k < MDCH_CHUNK / 32; k++) {
                    const int q = lane + 32 * k, p = pc + q;
                    if (p < npix) {
                        const int r = (int)__umulhi((unsigned)p, magic), c = p - r * cols;
                        const int yy = clampi(r0 + rs + r, 0, H - 1), xx = clampi(c0 + c, 0, W - 1);
                        buf[q] = binimg[(size_t)yy * W + xx];
                    }
                }
                __syncwarp();
                for (int k = 0; k < MDCH_CHUNK / 32; k++) {
                    const int q = lane + 32 * k, p = pc + q;
                    if (p < npix) {
                        const int r = (int)__umulhi((unsigned)p, magic), c = p - r * cols;
                        const unsigned bin = buf[q];
                        const double dx = (double)__fsub_rn(cX, (float)(c0 + c));
                        const double dy = (double)__fsub_rn(cY, (float)(r0 + rs + r));
                        const float wd = (float)(dx * dx * ivx + dy * dy * ivy);
                        const float wgt = expf(-wd);
                        atomicAdd(&hist[bin], __float2uint_rn(__fmul_rn(wgt, scale)));
                    }
                }
                __syncwarp();
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

__global__ void __launch_bounds__(MDCH_WARPS * 32)
k_mdch(const uint16_t* __restrict__ binimg, int W, int H, int nb, int w0, int h0,
       const int* __restrict__ col, const int* __restrict__ row, const float* __restrict__ sx,
       const float* __restrict__ sy, const int* __restrict__ valid, int n,
       const int* __restrict__ outbins, int n_out, float* __restrict__ out, int ld)
{
    mdch_body(binimg, W, H, nb, w0, h0, col, row, sx, sy, valid, n, outbins, n_out, out, ld);
}
// batched form: blockIdx.y = object; all nb^3 bins into rows [n_haar, n_haar + nb^3) of the training matrix
__global__ void __launch_bounds__(MDCH_WARPS * 32)
k_mdch_batch(const TrainDesc* __restrict__ descs, const uint16_t* __restrict__ binimg, int W, int H)
{
    const TrainDesc& d = descs[blockIdx.y];
    if (!(d.feature_type == MCT_FEAT_MDCH || d.feature_type == MCT_FEAT_HAAR_MDCH) || d.mdch_fast) return;
    mdch_body(binimg, W, H, d.nb, d.w0, d.h0, d.col, d.row, d.sx, d.sy, nullptr, d.P + d.Nn, nullptr, d.n_color,
              d.featT + (size_t)d.n_haar * d.ldT, d.ldT);
}

// ------------------------------------------------------------------------------------------
// MDCH rows of the TRAINING boxes (UpdateClassifier): all nb^3 bins of ~335 boxes per object.
// The training boxes of an object come in two groups (positives in a disc of radius ~10 px, negatives in a ring), every box of a
// group has the same size and the boxes overlap almost completely.  Instead of walking every box's pixels (1.07 M atomic
// histogram updates per object in k_mdch_batch), the group's bounding region (a few thousand pixels) is sorted by joint
// bin once, and then box k's bin-b value is the sum over the region's pixels of bin b of the box's weight at that pixel:
//     U[b][k] = sum_{p : bin(p) = b} E[offset(p) + base(k)]
// where E is the box's Gaussian weight table (fixed point, the same quantisation as mdch_body / k_mdch_sel) embedded in
// a zero field large enough that every (pixel, box) pair of the region lands inside it -- no bounds test, no atomics in
// the inner loop, lanes = boxes read consecutive table words.  Integer sums: identical to the per-box histograms.
//   k_mdch_train_prep    (group, object): table total, counting sort of the region's pixels by bin
//   k_mdch_train_main    (chunk, group, object): thread = box, walks its chunk of the sorted list, flushes per bin
//   k_mdch_train_finish  (bins, object): u32 -> normalised f32 rows, accumulators zeroed for the next frame
// counting sort with one private histogram per warp (no cross-warp contention, no match.any): [warp][dim] counts ->
// per-(bin, warp) start offsets -> scatter.  The region is cut into MTP_PARTS row-major parts, one CTA each; a part's
// sorted pixels occupy the same list range as its unsorted ones, so the list is a concatenation of sorted parts (the walk
// only needs runs of equal bins, not a global order).
#define MTP_WARPS 8
#define MTP_PARTS 4
__global__ void __launch_bounds__(MTP_WARPS * 32) k_mdch_train_prep(const TrainDesc* __restrict__ descs, const uint16_t* __restrict__ binimg, int W, int H)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ unsigned s_part[32];
    __shared__ unsigned s_base[1024];
    const TrainDesc& d = descs[blockIdx.z];
    if (!d.mdch_fast) return;
    const MdchGroup g = d.grp[blockIdx.y];
    if (g.n <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
    const int dim = d.n_color;
    const int part_len = (g.npx + MTP_PARTS - 1) / MTP_PARTS;
    const int q_begin = blockIdx.x * part_len, q_end = min(g.npx, q_begin + part_len), qn = q_end - q_begin;
    unsigned* hist = (unsigned*)smraw;                               // [MTP_WARPS][dim]
    uint16_t* sbin = (uint16_t*)(hist + MTP_WARPS * dim);            // [part_len] joint bins of this part
    for (int q = tid; q < MTP_WARPS * dim; q += nt) hist[q] = 0u;
    // the part's bins into shared memory, eight independent loads in flight per thread (pixels outside the frame take the
    // clamped pixel, as the per-box kernels do)
    for (int p0 = tid; p0 < qn; p0 += 8 * nt) {
        uint16_t v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int p = q_begin + min(p0 + q * nt, qn - 1);
            const int y = p / g.RW, x = p - y * g.RW;
            v[q] = binimg[(size_t)clampi(g.y0 + y, 0, H - 1) * W + clampi(g.x0 + x, 0, W - 1)];
        }
#pragma unroll
        for (int q = 0; q < 8; q++) if (p0 + q * nt < qn) sbin[p0 + q * nt] = v[q];
    }
    // fixed-point total of the weight table (part 0 only)
    if (blockIdx.x == 0) {
        const float scale = (float)(1u << g.shift);
        const float hx = __fdiv_rn((float)g.cols, 2.0f), hy = __fdiv_rn((float)g.rows, 2.0f);
        const double ivx = 1.0 / (double)__fmul_rn(2.0f, __fmul_rn(hx, hx));
        const double ivy = 1.0 / (double)__fmul_rn(2.0f, __fmul_rn(hy, hy));
        unsigned part = 0;
        for (int p = tid; p < g.rows * g.cols; p += nt) {
            const int r = p / g.cols, c = p - r * g.cols;
            const double dx = (double)__fsub_rn(hx, (float)c), dy = (double)__fsub_rn(hy, (float)r);
            const float wd = (float)(dx * dx * ivx + dy * dy * ivy);
            part += __float2uint_rn(__fmul_rn(expf(-wd), scale));
        }
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_part[warp] = part;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) { unsigned t = 0; for (int q = 0; q < (nt >> 5); q++) t += s_part[q]; d.mtot[blockIdx.y] = t; }
    if (qn <= 0) return;
    // each warp owns a contiguous slice of the part
    const int per_w = (qn + MTP_WARPS - 1) / MTP_WARPS;
    const int w_begin = warp * per_w, w_end = min(qn, w_begin + per_w);
    unsigned* myh = hist + warp * dim;
    for (int p = w_begin + lane; p < w_end; p += 32) atomicAdd(&myh[sbin[p]], 1u);
    __syncthreads();
    // per bin: exclusive scan over the warps, then the bins' bases (threads cover the bins in rounds of blockDim)
    unsigned carry = 0;
    for (int b0 = 0; b0 < dim; b0 += nt) {
        const int bq = b0 + tid;
        unsigned tot = 0;
        if (bq < dim)
            for (int w = 0; w < MTP_WARPS; w++) { const unsigned c = hist[w * dim + bq]; hist[w * dim + bq] = tot; tot += c; }
        unsigned inc = tot;
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __syncthreads();
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        unsigned wb = carry, all = 0;
        for (int w = 0; w < (nt >> 5); w++) { if (w < warp) wb += s_part[w]; all += s_part[w]; }
        if (bq < dim) s_base[bq] = wb + inc - tot;
        carry += all;
    }
    __syncthreads();
    uint32_t* list = d.mlist + g.list_off + q_begin;
    for (int p = w_begin + lane; p < w_end; p += 32) {
        const unsigned bin = sbin[p];
        const unsigned pos = s_base[bin] + atomicAdd(&myh[bin], 1u);
        const int pp = q_begin + p;
        const int y = pp / g.RW, x = pp - y * g.RW;
        list[pos] = (bin << 22) | ((unsigned)y << 11) | (unsigned)x;
    }
}

// thread = (box, slice of the CTA's list chunk); lanes of a warp are consecutive boxes of the same slice.
// The chunk's runs of equal bins are found once per CTA; a thread then sums a run's table words in a register with no
// bin test in the loop, and adds the sum to the (bin, box) accumulator with one global integer add.
// EXT: the box's weight table sits in a zero field [2 RH - rows][2 RW - cols] so that every (pixel, box) pair of the region
// indexes it without a bounds test (used when the field is small: the positives' disc); otherwise bounds-tested lookups.
#define MTR_THREADS 384
struct MtrRun { unsigned bin; int start, end; int pad; };
template <bool EXT>
__device__ __forceinline__ void mdch_train_walk(const TrainDesc& d, const MdchGroup& g, const uint32_t* __restrict__ tab, const uint32_t* __restrict__ sl,
                                                const MtrRun* __restrict__ runs, int nruns, int npx)
{
    const int tid = threadIdx.x;
    const int B = min((g.n + 31) & ~31, (int)blockDim.x);         // boxes per pass
    const int S = blockDim.x / B;                                  // list slices walked side by side
    const int kk = tid % B, s = tid / B;
    if (s >= S) return;
    const int len = (npx + S - 1) / S;
    const int p0 = s * len, p1 = min(npx, p0 + len);
    if (p0 >= p1) return;
    const unsigned ldT = (unsigned)d.ldT;
    const int EW = 2 * g.RW - g.cols;
    for (int k0 = 0; k0 < g.n; k0 += B) {
        const int k = k0 + kk;
        if (k >= g.n) continue;
        const int i = g.first + k;
        const int cxk = d.col[i] - g.x0, cyk = d.row[i] - g.y0;              // box corner in region coordinates
        const int base = (g.RH - g.rows - cyk) * EW + (g.RW - g.cols - cxk);
        uint32_t* acc_col = d.macc + i;
        for (int r = 0; r < nruns; r++) {
            const MtrRun rn = runs[r];
            const int a = max(rn.start, p0), e = min(rn.end, p1);           // warp-uniform
            if (a >= e) continue;
            unsigned acc = 0;
            if (EXT) {
                int p = a;
                for (; p + 4 <= e; p += 4) {
                    const unsigned w0 = sl[p], w1 = sl[p + 1], w2 = sl[p + 2], w3 = sl[p + 3];
                    acc += tab[(int)w0 + base] + tab[(int)w1 + base] + tab[(int)w2 + base] + tab[(int)w3 + base];
                }
                for (; p < e; p++) acc += tab[(int)sl[p] + base];
            } else {
                for (int p = a; p < e; p++) {
                    const unsigned w = sl[p];
                    const unsigned dx = (w & 0x7ffu) - (unsigned)cxk, dy = (w >> 11) - (unsigned)cyk;
                    if (dx < (unsigned)g.cols && dy < (unsigned)g.rows) acc += tab[dy * g.cols + dx];
                }
            }
            if (acc) atomicAdd(acc_col + (size_t)rn.bin * ldT, acc);
        }
    }
}
__global__ void __launch_bounds__(MTR_THREADS) k_mdch_train_main(const TrainDesc* __restrict__ descs, int smem_tab_words, int smem_chunk_words)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ int s_nruns;
    const TrainDesc& d = descs[blockIdx.z];
    if (!d.mdch_fast) return;
    const MdchGroup g = d.grp[blockIdx.y];
    const int c_begin = blockIdx.x * g.chunk_px;
    if (g.n <= 0 || c_begin >= g.npx) return;
    const int c_end = min(g.npx, c_begin + g.chunk_px), npx = c_end - c_begin;
    const int tid = threadIdx.x;
    uint32_t* tab = (uint32_t*)smraw;                            // weight table (plain or embedded in the zero field)
    uint32_t* sl = tab + smem_tab_words;                         // the chunk's table offsets (or packed (y, x))
    MtrRun* runs = (MtrRun*)(sl + smem_chunk_words);             // [<= npx]
    const int EW = 2 * g.RW - g.cols, EH = 2 * g.RH - g.rows;
    const bool ext = g.EW != 0;                                  // decided with the plan (host: mdch_group_plan; device: k_train_gen_default)
    if (tid == 0) s_nruns = 0;
    if (ext) for (int p = tid; p < EW * EH; p += blockDim.x) tab[p] = 0u;
    __syncthreads();
    // the chunk's list: table offsets into sl, bins into sb
    uint16_t* sb = (uint16_t*)(runs + smem_chunk_words);         // [npx]
    const uint32_t* gl = d.mlist + g.list_off + c_begin;
    for (int p = tid; p < npx; p += blockDim.x) {
        const unsigned w = gl[p];
        const unsigned y = (w >> 11) & 0x7ffu, x = w & 0x7ffu;
        sl[p] = ext ? (y * (unsigned)EW + x) : ((y << 11) | x);
        sb[p] = (uint16_t)(w >> 22);
    }
    __syncthreads();
    // runs of equal bins (a bin may occur in several runs when the chunk straddles sorted parts): every thread takes a
    // contiguous stretch, counts the run starts in it, a block scan numbers them
    {
        __shared__ int s_w[MTR_THREADS / 32];
        const int per = (npx + (int)blockDim.x - 1) / (int)blockDim.x;
        const int a0 = min(npx, tid * per), a1 = min(npx, a0 + per);
        int cnt = 0;
        for (int p = a0; p < a1; p++) cnt += (p == 0 || sb[p - 1] != sb[p]) ? 1 : 0;
        int inc = cnt;
        const int lane = tid & 31, warp = tid >> 5;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        int wb = 0, all = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { if (w < warp) wb += s_w[w]; all += s_w[w]; }
        int idx = wb + inc - cnt - 1;                            // run that contains the element in front of a0
        for (int p = a0; p < a1; p++) {
            const unsigned bin = sb[p];
            if (p == 0 || sb[p - 1] != bin) { idx++; runs[idx].bin = bin; runs[idx].start = p; }
            if (p == npx - 1 || sb[p + 1] != bin) runs[idx].end = p + 1;
        }
        if (tid == 0) s_nruns = all;
    }
    {
        const float scale = (float)(1u << g.shift);
        const float hx = __fdiv_rn((float)g.cols, 2.0f), hy = __fdiv_rn((float)g.rows, 2.0f);
        const double ivx = 1.0 / (double)__fmul_rn(2.0f, __fmul_rn(hx, hx));
        const double ivy = 1.0 / (double)__fmul_rn(2.0f, __fmul_rn(hy, hy));
        for (int p = tid; p < g.rows * g.cols; p += blockDim.x) {
            const int r = p / g.cols, c = p - r * g.cols;
            const double dx = (double)__fsub_rn(hx, (float)c), dy = (double)__fsub_rn(hy, (float)r);
            const float wd = (float)(dx * dx * ivx + dy * dy * ivy);
            const uint32_t wq = __float2uint_rn(__fmul_rn(expf(-wd), scale));
            if (ext) tab[(r + g.RH - g.rows) * EW + c + g.RW - g.cols] = wq;
            else tab[p] = wq;
        }
    }
    __syncthreads();
    if (ext) mdch_train_walk<true>(d, g, tab, sl, runs, s_nruns, npx);
    else mdch_train_walk<false>(d, g, tab, sl, runs, s_nruns, npx);
}

// u32 sums -> normalised f32 rows; the accumulators are left zero for the next frame.  Flat over (bin, box quad), the
// four elements of a thread are independent.
__global__ void __launch_bounds__(256) k_mdch_train_finish(const TrainDesc* __restrict__ descs)
{
    const TrainDesc& d = descs[blockIdx.y];
    if (!d.mdch_fast) return;
    const int n = d.P + d.Nn, first1 = d.grp[1].first, ld = d.ldT;
    const int nq = (n + 3) >> 2;                                  // quads per row
    const int u = blockIdx.x * 256 + threadIdx.x;
    if (u >= d.n_color * nq) return;
    const int b = u / nq, i = (u - b * nq) * 4;
    const float is0 = __fdiv_rn(1.0f, (float)(1u << d.grp[0].shift)), is1 = __fdiv_rn(1.0f, (float)(1u << d.grp[1].shift));
    const float S0 = __fmul_rn((float)d.mtot[0], is0), S1 = __fmul_rn((float)d.mtot[1], is1);
    uint32_t* src = d.macc + (size_t)b * ld + i;                  // ld is a multiple of 4, the buffer 16-byte aligned
    const uint4 t = *reinterpret_cast<const uint4*>(src);
    if (t.x | t.y | t.z | t.w) *reinterpret_cast<uint4*>(src) = make_uint4(0u, 0u, 0u, 0u);     // (most bins are empty: nothing to clear)
    const unsigned tv[4] = {t.x, t.y, t.z, t.w};
    float* dst = d.featT + (size_t)(d.n_haar + b) * ld + i;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (i + q >= n) break;
        dst[q] = (i + q >= first1) ? __fdiv_rn(__fmul_rn((float)tv[q], is1), S1) : __fdiv_rn(__fmul_rn((float)tv[q], is0), S0);
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
    // larger lists: bin the boxes into tiles and gather from shared memory
    const int ext_w = clf->haar_mxw + 2, ext_h = clf->haar_mxh + 2;  // rect extent at scale 1 (+ rounding slack); larger boxes fall back per box
    const int tiles_x = ceil_div(ctx->W, HT_TS), tiles_y = ceil_div(ctx->H, HT_TS), n_tiles = tiles_x * tiles_y;
    const size_t reg_b = (size_t)(HT_TS + ext_w + 12) * (HT_TS + ext_h + 1) * 4;
    // (the tiled path pays three tiny sort launches and one region staging per (tile, feature chunk): it wins once a
    //  tile holds a few dozen boxes; MCT_HAAR_TILED=0/1 in the environment forces the choice for tests and sweeps)
    static int s_force = -2;
    if (s_force == -2) { const char* e = getenv("MCT_HAAR_TILED"); s_force = e ? atoi(e) : -1; }
    // a (tile, chunk) CTA does 16 * HT_FT gathers per box against one staged region: worth it from about one gather per staged float
    const int ft = clf->n_haar <= 512 ? 32 : HT_FT;                 // smaller chunks keep enough CTAs in flight when F is small
    const bool want = s_force >= 0 ? (s_force != 0 && n >= 2) : (n >= 1024 && (size_t)(n / n_tiles) * 16 * ft >= reg_b / 4);
    if (want && reg_b <= 200 * 1024 && n_tiles <= 8192 && ((size_t)ctx->ii_pitch & 3) == 0) {
        const size_t need = (size_t)2 * n_tiles + 1 + (size_t)n;
        if (need > clf->tsort_cap) {
            MCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (clf->d_tsort) cudaFree(clf->d_tsort);
            MCT_CUDA(ctx, cudaMalloc(&clf->d_tsort, (need + 1024) * sizeof(int)));
            clf->tsort_cap = need + 1024;
        }
        int* d_start = clf->d_tsort; int* d_cnt = clf->d_tsort + n_tiles + 1; int* d_order = d_cnt + n_tiles;
        if (n <= 32768 && (size_t)n_tiles * sizeof(int) <= 40 * 1024) {
            MCT_LAUNCH(ctx, "k_tile_sort", k_tile_sort<<<1, 1024, (size_t)n_tiles * sizeof(int), ctx->stream>>>(d_col, d_row, n, ctx->W, ctx->H, tiles_x,
                                                                                                         n_tiles, d_start, d_order));
        } else {
            MCT_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, (size_t)n_tiles * sizeof(int), ctx->stream));
            MCT_LAUNCH(ctx, "k_tile_sort", k_tile_count<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_col, d_row, n, ctx->W, ctx->H, tiles_x, d_cnt));
            MCT_LAUNCH(ctx, "k_tile_sort", k_tile_scan<<<1, 1024, 0, ctx->stream>>>(d_cnt, n_tiles, n, d_start));
            MCT_LAUNCH(ctx, "k_tile_sort", k_tile_scatter<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_col, d_row, n, ctx->W, ctx->H, tiles_x, d_cnt, d_order));
        }
        static size_t s_attr[MCT_MAX_DEV];
        if (int rc_ = mct_smem_attr(ctx, k_haar_matrix_tiled, s_attr, reg_b, 0)) return rc_;
        dim3 g(n_tiles, ceil_div(clf->n_haar, ft));
        MCT_LAUNCH(ctx, "k_haar_matrix_tiled", k_haar_matrix_tiled<<<g, HT_THREADS, reg_b, ctx->stream>>>(
            ctx->ii_base + MCT_II_LEAD, ctx->ii_pitch, ctx->W, ctx->H, clf->d_rects, clf->d_wts, clf->d_nrect, clf->n_haar,
            d_col, d_row, d_sx, d_sy, d_start, d_order, tiles_x, ext_w, ext_h, clf->haar_mxw, clf->haar_mxh, ft, d_out, ld));
        return MCT_OK;
    }
    dim3 g(ceil_div(n, 256), ceil_div(clf->n_haar, HAAR_FT));
    MCT_LAUNCH(ctx, "k_haar_matrix", k_haar_matrix<<<g, 256, 0, ctx->stream>>>(ctx->ii_base + MCT_II_LEAD, ctx->ii_pitch, ctx->W, ctx->H, clf->d_rects,
                                              clf->d_wts, clf->d_nrect, clf->n_haar, d_col, d_row, d_sx, d_sy, n, d_out, ld));
    return MCT_OK;
}

static size_t mdch_smem(int dim, int n_out)
{
    return (size_t)MDCH_WARPS * dim * (sizeof(uint32_t) + sizeof(float)) + (size_t)n_out * (MDCH_TILE + 1) * sizeof(float) +
           (size_t)MDCH_WARPS * MDCH_CHUNK * sizeof(uint16_t);
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
    static size_t s_attr[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_mdch, s_attr, smem, 0)) return rc_;
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

// training features of all objects: Haar rows + MDCH rows, two launches for the whole camera
int launch_train_features(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj, const TrainLaunchBounds* lb, bool weak_too)
{
    ctx->weak_split_done = 0;
    // (MILEnsemble's retain step sits between the rows and the weak update and reads all rows: no split with such objects)
    for (int o = 0; o < n_obj && weak_too; o++) if (h[o].strong_type == MCT_STRONG_MILENSEMBLE) weak_too = false;
    if (getenv("MCT_NO_WEAK_SPLIT")) weak_too = false;
    int n_max = lb ? lb->n_max : 0, nh_max = 0, any_mdch = 0, dim_max = 0, nb = 0;
    for (int o = 0; o < n_obj; o++) {
        if (!lb) n_max = max(n_max, h[o].P + h[o].Nn);
        nh_max = max(nh_max, h[o].n_haar);
        if (h[o].feature_type == MCT_FEAT_MDCH || h[o].feature_type == MCT_FEAT_HAAR_MDCH) { any_mdch = 1; dim_max = max(dim_max, h[o].n_color); nb = h[o].nb; }
    }
    if (n_max <= 0) return MCT_OK;
    if (any_mdch) { int rc = launch_bin_image(ctx, nb); if (rc) return rc; }     // (in front of the fork: both streams read it)
    bool forked = false;
    if (nh_max > 0) {
        dim3 g(ceil_div(n_max, 256), ceil_div(nh_max, HAAR_FT), n_obj);
        // the Haar rows and the MDCH rows of the training boxes are independent: with both feature families present the Haar matrix
        // runs on a side stream beside the three MDCH kernels and joins in front of the weak-classifier update
        if (any_mdch && ctx->train_stream && !getenv("MCT_TRAIN_SERIAL")) {
            MCT_CUDA(ctx, cudaEventRecord(ctx->ev_tfork, ctx->stream));
            MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->train_stream, ctx->ev_tfork, 0));
            MCT_LAUNCH_ON(ctx, ctx->train_stream, "k_haar_matrix_batch",
                          k_haar_matrix_batch<<<g, 256, 0, ctx->train_stream>>>(d, ctx->ii_base + MCT_II_LEAD, ctx->ii_pitch, ctx->W, ctx->H));
            forked = true;
        } else {
            MCT_LAUNCH(ctx, "k_haar_matrix_batch", k_haar_matrix_batch<<<g, 256, 0, ctx->stream>>>(d, ctx->ii_base + MCT_II_LEAD, ctx->ii_pitch, ctx->W, ctx->H));
        }
    }
    if (any_mdch) {
        // objects whose two groups qualified on the host take the pixel-list kernels; the others the generic per-box kernel
        int n_fast = 0, n_slow = 0, nch = 1, tab_words = 0, chunk_words = 0, npx_max = 0, n_train_max = 0;
        if (lb) {
            // sizes and plans are decided by k_train_gen_default: both kernel families are launched from the bounds, every
            // object takes one of them (TrainDesc::mdch_fast on the device)
            n_fast = lb->nch > 0 ? 1 : 0; n_slow = 1;
            nch = max(lb->nch, 1); tab_words = lb->tab_words; chunk_words = lb->chunk_words; npx_max = lb->npx_max; n_train_max = lb->n_train_max;
        }
        for (int o = 0; o < n_obj && !lb; o++) {
            if (!(h[o].feature_type == MCT_FEAT_MDCH || h[o].feature_type == MCT_FEAT_HAAR_MDCH)) continue;
            if (h[o].mdch_fast) {
                n_fast++;
                n_train_max = max(n_train_max, h[o].P + h[o].Nn);
                for (int gi = 0; gi < 2; gi++) {
                    const MdchGroup& g = h[o].grp[gi];
                    nch = max(nch, g.nchunk);
                    const int ext = (2 * g.RW - g.cols) * (2 * g.RH - g.rows);
                    tab_words = max(tab_words, ext <= MTR_EXT_WORDS ? ext : g.rows * g.cols);
                    chunk_words = max(chunk_words, g.chunk_px);
                    npx_max = max(npx_max, g.npx);
                }
            } else n_slow++;
        }
        if (n_fast) {
            const size_t smem_main = (size_t)(tab_words + chunk_words) * 4 + (size_t)chunk_words * sizeof(MtrRun) + (size_t)chunk_words * 2 + 16;
            const size_t smem_prep = (size_t)MTP_WARPS * dim_max * 4 + (size_t)ceil_div(npx_max, MTP_PARTS) * 2 + 16;
            static size_t s_attr_m[MCT_MAX_DEV], s_attr_p[MCT_MAX_DEV];
            if (int rc_ = mct_smem_attr(ctx, k_mdch_train_main, s_attr_m, smem_main, 48 * 1024)) return rc_;
            if (int rc_ = mct_smem_attr(ctx, k_mdch_train_prep, s_attr_p, smem_prep, 48 * 1024)) return rc_;
            MCT_LAUNCH(ctx, "k_mdch_train_prep", k_mdch_train_prep<<<dim3(MTP_PARTS, 2, n_obj), MTP_WARPS * 32, smem_prep, ctx->stream>>>(d, ctx->binimg, ctx->W, ctx->H));
            MCT_LAUNCH(ctx, "k_mdch_train_main", k_mdch_train_main<<<dim3(nch, 2, n_obj), MTR_THREADS, smem_main, ctx->stream>>>(d, tab_words, chunk_words));
            MCT_LAUNCH(ctx, "k_mdch_train_finish", k_mdch_train_finish<<<dim3(ceil_div(dim_max * ceil_div(n_train_max, 4), 256), n_obj), 256, 0, ctx->stream>>>(d));
        }
        if (n_slow) {
            size_t smem = mdch_smem(dim_max, dim_max);
            static size_t s_attr[MCT_MAX_DEV];
            if (int rc_ = mct_smem_attr(ctx, k_mdch_batch, s_attr, smem, 0)) return rc_;
            dim3 g(ceil_div(n_max, MDCH_TILE), n_obj);
            // device-decided plans: the generic kernel is launched for the objects whose plan left the launch bounds -- normally
            // none, every CTA leaves at once -- so it rides on the side stream behind the Haar rows, off the critical path
            if (lb && forked) MCT_LAUNCH_ON(ctx, ctx->train_stream, "k_mdch_batch", k_mdch_batch<<<g, MDCH_WARPS * 32, smem, ctx->train_stream>>>(d, ctx->binimg, ctx->W, ctx->H));
            else MCT_LAUNCH(ctx, "k_mdch_batch", k_mdch_batch<<<g, MDCH_WARPS * 32, smem, ctx->stream>>>(d, ctx->binimg, ctx->W, ctx->H));
        }
    }
    if (forked && weak_too) {
        int rc;
        if ((rc = launch_weak_update_part(ctx, d, h, n_obj, 1, ctx->train_stream))) return rc;      // Haar part behind the Haar rows, beside the colour kernels
    }
    if (forked) {
        MCT_CUDA(ctx, cudaEventRecord(ctx->ev_tjoin, ctx->train_stream));
        MCT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_tjoin, 0));
    }
    if (forked && weak_too) {
        // colour part behind the join: the generic colour kernel (objects whose plan left the launch bounds) rides on the side stream
        int rc;
        if ((rc = launch_weak_update_part(ctx, d, h, n_obj, 2, ctx->stream))) return rc;
        ctx->weak_split_done = 1;
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

// ==========================================================================================
// MDCH for Classify: only the SELECTED colour bins are read by the strong classifier
// (MILBoostClassifier.cpp:351-361), so the test-set pass keeps, per box, one fixed-point accumulator per
// selected bin ("slot"; slot 0 collects every unselected bin).
//
//   k_build_colormap : selector list -> bin->slot map, colour feature -> featN row      (once per Update)
//   k_mdch_sel       : persistent CTAs, each owning a contiguous chunk of ONE object's test boxes:
//       1. bounding box of the chunk's boxes; the joint-bin image of that region is staged ONCE into shared
//          memory as u8 slots (slot map applied while staging: 4 pixels per thread, one packed STS);
//       2. the Gaussian weights depend on the offset inside the box only (cX - x = cols/2 - c exactly), so one
//          fixed-point LUT per CTA serves every box of the common size (MultiDimensionalColorHistogram.cpp:
//          66-69, 87-89, 107-110; weights in [e^-1, 1], scale 2^20 -> <= 2.6e-6 relative quantisation);
//       3. T = cols/4 lanes per box, each lane owning a 4-pixel column strip and a PRIVATE histogram
//          hist[slot][lane] (bank == lane: conflict-free plain LDS/IADD/STS -- shared atomics cost 2 cycles per
//          lane on sm_100a, and warp match/redux aggregation 29 cycles per 32 pixels; both were measured slower,
//          see tools/mdch_lab.cu).  The four read-modify-writes of a row are issued as 4 loads, a register
//          forwarding network for equal slots, and 4 stores: one LDS latency per 4 pixels instead of four;
//       4. the T partial histograms of a box are summed (skewed order: no bank conflicts) and normalised:
//          normalizeVec (Public.h:270-275) divides by the sum over all nb^3 bins, which for a box inside the
//          frame is the LUT total.
//     Integer accumulation is order independent, so the result is run-to-run bit-stable.
//   k_mdch_big       : generic warp-per-box kernel for the boxes k_mdch_sel flags (other size than the chunk's
//                      common one, crossing the frame, > 128 columns, region larger than shared memory, or more
//                      than 255 selected bins).
// Shared-memory-resident gather kernel: per box rows*cols slot bytes + LUT words + private RMW traffic.
// ==========================================================================================
__device__ __forceinline__ void build_colormap_body(const int* __restrict__ sel, const int* __restrict__ nsel, int n_haar, int n_color, int mdch,
                                 uint16_t* __restrict__ slotmap, int* __restrict__ colorrow, int* __restrict__ outbins,
                                 int* __restrict__ nout, const int4* __restrict__ rects, const float* __restrict__ wts,
                                 const int* __restrict__ nrect, const float* __restrict__ weak, int F, int K,
                                 SelEntry* __restrict__ tab, int* __restrict__ hdr, const float* __restrict__ alpha = nullptr,
                                 float* __restrict__ rowst = nullptr)
{
    __shared__ int s_ext[2];
    // the first-occurrence walk over the selector list runs out of shared memory (it was a chain of dependent global loads and
    // stores on one thread: 0.45 us per selector, 350 us per launch with the 102 selectors of BASELINE configs[3])
    constexpr int CM_SEL = 1024, CM_BINS = 1024;
    __shared__ int s_sel[CM_SEL];
    __shared__ unsigned short s_slot[CM_BINS];
    __shared__ unsigned short s_row[CM_BINS];
    const int ns = *nsel;
    if (threadIdx.x == 0) { s_ext[0] = 0; s_ext[1] = 0; }
    if (mdch && ns <= CM_SEL && n_color <= CM_BINS) {
        for (int b = threadIdx.x; b < n_color; b += blockDim.x) { s_slot[b] = 0; s_row[b] = 0; }
        for (int q = threadIdx.x; q < ns; q += blockDim.x) s_sel[q] = sel[q];
        __syncthreads();
        if (threadIdx.x == 0) {
            int n = 0;
            for (int q = 0; q < ns; q++) {
                const int f = s_sel[q];
                if (f < n_haar) continue;
                const int b = f - n_haar;
                if (s_slot[b] == 0) { s_sel[n] = b; s_row[b] = (unsigned short)n; n++; s_slot[b] = (unsigned short)n; }   // (n <= q: s_sel[n] was read)
            }
            s_ext[0] = n;                                          // (parked here until the rect extents are collected below)
        }
        __syncthreads();
        const int n = s_ext[0];
        __syncthreads();
        if (threadIdx.x == 0) { *nout = n; s_ext[0] = 0; }
        for (int b = threadIdx.x; b < n_color; b += blockDim.x) { slotmap[b] = s_slot[b]; colorrow[b] = s_row[b]; }
        for (int q = threadIdx.x; q < n; q += blockDim.x) outbins[q] = s_sel[q];
    } else if (mdch) {
        for (int b = threadIdx.x; b < n_color; b += blockDim.x) { slotmap[b] = 0; colorrow[b] = 0; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int n = 0;
            for (int s = 0; s < ns; s++) {
                const int f = sel[s];
                if (f < n_haar) continue;
                const int b = f - n_haar;
                if (slotmap[b] == 0) { outbins[n] = b; colorrow[b] = n; n++; slotmap[b] = (uint16_t)n; }
            }
            *nout = n;
        }
    }
    __syncthreads();
    // the prepared table read by k_classify_staged (entries beyond nsel are zero)
    int mxw = 0, mxh = 0;
    for (int s = threadIdx.x; s < K; s += blockDim.x) {
        SelEntry e;
        memset(&e, 0, sizeof(e));
        if (s < ns) {
            const int f = sel[s];
            for (int q = 0; q < 8; q++) e.st[q] = weak[(size_t)q * F + f];
            if (alpha) e.st[2] = alpha[s];                    // AdaBoost: the vote weight rides in the (unused) sigma0 slot
            if (f < n_haar) {
                e.nrect = nrect[f];
                for (int k = 0; k < MCT_MAX_RECTS; k++) {
                    e.r[k] = rects[(size_t)f * MCT_MAX_RECTS + k];
                    e.w[k] = wts[(size_t)f * MCT_MAX_RECTS + k];
                    if (k < e.nrect) { mxw = max(mxw, e.r[k].x + e.r[k].z); mxh = max(mxh, e.r[k].y + e.r[k].w); }
                }
            } else {
                e.row = colorrow[f - n_haar];
                // weak state by colour row, for the response evaluation fused into k_mdch_sel (a bin selected twice has one state)
                if (rowst && mdch) for (int q = 0; q < 8; q++) rowst[(size_t)e.row * 8 + q] = e.st[q];
            }
        }
        tab[s] = e;
    }
    if (mxw > 0) { atomicMax(&s_ext[0], mxw); atomicMax(&s_ext[1], mxh); }
    __syncthreads();
    if (threadIdx.x == 0) { hdr[0] = ns < K ? ns : K; hdr[1] = s_ext[0]; hdr[2] = s_ext[1]; hdr[3] = 0; }
}
__global__ void k_build_colormap(const int* __restrict__ sel, const int* __restrict__ nsel, int n_haar, int n_color, int mdch,
                                 uint16_t* __restrict__ slotmap, int* __restrict__ colorrow, int* __restrict__ outbins,
                                 int* __restrict__ nout, const int4* __restrict__ rects, const float* __restrict__ wts,
                                 const int* __restrict__ nrect, const float* __restrict__ weak, int F, int K,
                                 SelEntry* __restrict__ tab, int* __restrict__ hdr, const float* __restrict__ alpha, float* __restrict__ rowst)
{
    build_colormap_body(sel, nsel, n_haar, n_color, mdch, slotmap, colorrow, outbins, nout, rects, wts, nrect, weak, F, K, tab, hdr, alpha, rowst);
}
__global__ void k_build_colormap_batch(const TrainDesc* __restrict__ descs)
{
    const TrainDesc& d = descs[blockIdx.x];
    if (d.P < 1 || d.Nn < 1) return;              // empty training set decided on the device (k_train_gen_default): classifier left as it is
    const int mdch = (d.feature_type == MCT_FEAT_MDCH || d.feature_type == MCT_FEAT_HAAR_MDCH) ? 1 : 0;
    build_colormap_body(d.sel, d.nsel, d.n_haar, d.n_color, mdch, d.slotmap, d.colorrow, d.outbins, d.nout, d.rects, d.wts, d.nrect,
                        d.weak, d.F, d.K, d.seltab, d.selhdr, d.strong_type == MCT_STRONG_ADABOOST ? d.alpha : nullptr, d.rowst);
}
int launch_build_colormap_batch(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj)
{
    (void)h;
    MCT_LAUNCH(ctx, "k_build_colormap_batch", k_build_colormap_batch<<<n_obj, 256, 0, ctx->stream>>>(d));
    return MCT_OK;
}
int launch_build_colormap(mct_clf* clf)
{
    mct_ctx* ctx = clf->ctx;
    const int mdch = (clf->p.feature_type == MCT_FEAT_MDCH || clf->p.feature_type == MCT_FEAT_HAAR_MDCH) ? 1 : 0;
    MCT_LAUNCH(ctx, "k_build_colormap", k_build_colormap<<<1, 256, 0, ctx->stream>>>(clf->d_sel, clf->d_nsel, clf->n_haar, clf->n_color, mdch,
                                                                                  clf->d_slotmap, clf->d_colorrow, clf->d_outbins, clf->d_nout,
                                                                                  clf->d_rects, clf->d_wts, clf->d_nrect, clf->d_weak, clf->F, clf->K,
                                                                                  clf->d_seltab, clf->d_selhdr,
                                                                                  clf->p.strong_type == MCT_STRONG_ADABOOST ? clf->d_alpha : nullptr, clf->d_rowst));
    return MCT_OK;
}

// generic warp-per-box pass over the boxes flagged by k_mdch_sel (segmented fixed point, all nb^3 bins in
// shared memory), writing the selected bins only.  `first` / `stride` / `nwarps` describe the calling kernel's warp layout.
// Everything another CTA of the same launch may have written (box geometry from the fused predict prologue, the flags) is
// read with ld.cg: L1 is not coherent across SMs.
__device__ __forceinline__ void mdch_big_body(const ObjDesc& d, const uint16_t* __restrict__ binimg, int W, int H,
                                              unsigned char* smraw, int nwarps, int first, int stride)
{
    const int nb = d.nb, dim = nb * nb * nb;
    uint32_t* hist_all = (uint32_t*)smraw;
    float* acc_all = (float*)(hist_all + nwarps * dim);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* hist = hist_all + warp * dim;
    float* acc = acc_all + warp * dim;
    const int n_out = *d.nout;
    for (int i = first; i < d.N; i += stride) {
        if (!__ldcg(d.big + i)) continue;
        const int c0 = __ldcg(d.col + i), r0 = __ldcg(d.row + i);
        const int rows = __float2int_rn(__fmul_rn((float)d.ch0, __ldcg(d.sy + i)));
        const int cols = __float2int_rn(__fmul_rn((float)d.cw0, __ldcg(d.sx + i)));
        const float hx = __fdiv_rn((float)cols, 2.0f), hy = __fdiv_rn((float)rows, 2.0f);
        const float cX = __fadd_rn((float)c0, hx), cY = __fadd_rn((float)r0, hy);
        const double ivx = 1.0 / (double)__fmul_rn(2.0f, __fmul_rn(hx, hx));
        const double ivy = 1.0 / (double)__fmul_rn(2.0f, __fmul_rn(hy, hy));
        for (int b = lane; b < dim; b += 32) { hist[b] = 0u; acc[b] = 0.0f; }
        __syncwarp();
        const int rows_per_seg = max(1, MCT_MDCH_SEG_PIX / max(cols, 1));
        for (int rs = 0; rs < rows; rs += rows_per_seg) {
            const int re = min(rows, rs + rows_per_seg);
            const int npix = (re - rs) * cols;
            int shift = __clz(max(npix, 1)); if (shift > MCT_FIX_SHIFT_MAX) shift = MCT_FIX_SHIFT_MAX;
            const float scale = (float)(1u << shift);
            int c = lane, r = rs;
            while (c >= cols && r < re) { c -= cols; r++; }
            while (r < re) {
                const int yy = clampi(r0 + r, 0, H - 1), xx = clampi(c0 + c, 0, W - 1);
                const unsigned bin = binimg[(size_t)yy * W + xx];
                const double dx = (double)__fsub_rn(cX, (float)(c0 + c)), dy = (double)__fsub_rn(cY, (float)(r0 + r));
                const float wd = (float)(dx * dx * ivx + dy * dy * ivy);
                atomicAdd(&hist[bin], __float2uint_rn(__fmul_rn(expf(-wd), scale)));
                c += 32;
                while (c >= cols && r < re) { c -= cols; r++; }
            }
            __syncwarp();
            const float inv = __fdiv_rn(1.0f, scale);
            for (int b = lane; b < dim; b += 32) { acc[b] = __fadd_rn(acc[b], __fmul_rn((float)hist[b], inv)); hist[b] = 0u; }
            __syncwarp();
        }
        float s = 0.0f;
        for (int b = lane; b < dim; b += 32) s += acc[b];
        s = warp_sum_f(s);
        for (int o = lane; o < n_out; o += 32) {
            float x = __fdiv_rn(acc[d.outbins[o]], s);
            if (d.color_resp) {
                const float* st = d.rowst + (size_t)o * 8;
                x = weak_classify_q_dev(d.weak_type, x, st[0], st[1], st[4], st[5], st[6], st[7]);
            }
            d.featN[(size_t)o * d.ldN + i] = x;
        }
        __syncwarp();
    }
}

#define MDS_WARPS 16
#define MDS_SMEM (226 * 1024)
#define MDS_MIN_WARPS 8           // fewer private histograms than this in the single-group plan -> row-band plan
// the gather over this CTA's chunk [i_begin, i_end) of the object's boxes
__device__ __forceinline__ void mdch_sel_chunk(const ObjDesc& d, const uint16_t* __restrict__ binimg, int W, int H, int smem_bytes,
                                               unsigned char* smraw, int i_begin, int i_end)
{
    __shared__ int s_bb[4], s_ref, s_rows, s_cols, s_nord, s_gend;
    __shared__ uint32_t s_tot[MDS_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_out = *d.nout, n_slots = n_out + 1, dim = d.n_color_rows;

    // ---- 1. common box size of the chunk (first usable box) and bounding box of the conforming boxes
    if (tid == 0) { s_ref = 0x7fffffff; s_bb[0] = 1 << 30; s_bb[1] = 1 << 30; s_bb[2] = -(1 << 30); s_bb[3] = -(1 << 30); }
    __syncthreads();
    // per-thread view of its boxes is recomputed where needed (registers are scarce in the main loop)
    for (int i = i_begin + tid; i < i_end; i += blockDim.x) {
        if (d.valid != nullptr && d.valid[i] == 0) continue;
        const int rows = __float2int_rn(__fmul_rn((float)d.ch0, d.sy[i]));
        const int cols = __float2int_rn(__fmul_rn((float)d.cw0, d.sx[i]));
        const int c0 = d.col[i], r0 = d.row[i];
        if (rows > 0 && cols > 0 && cols <= 128 && c0 >= 0 && r0 >= 0 && c0 + cols <= W && r0 + rows <= H) atomicMin(&s_ref, i);
    }
    __syncthreads();
    const int ref = s_ref;
    if (tid == 0 && ref != 0x7fffffff) {
        s_rows = __float2int_rn(__fmul_rn((float)d.ch0, d.sy[ref]));
        s_cols = __float2int_rn(__fmul_rn((float)d.cw0, d.sx[ref]));
    }
    __syncthreads();
    const int rows = ref != 0x7fffffff ? s_rows : 0, cols = ref != 0x7fffffff ? s_cols : 0;
    const int colsp = (cols + 3) & ~3, T = max(colsp >> 2, 1), BPW = 32 / T;
    {
        int x0 = 1 << 30, y0 = 1 << 30, x1 = -(1 << 30), y1 = -(1 << 30);
        for (int i = i_begin + tid; i < i_end; i += blockDim.x) {
            if (d.valid != nullptr && d.valid[i] == 0) continue;
            const int br = __float2int_rn(__fmul_rn((float)d.ch0, d.sy[i]));
            const int bc = __float2int_rn(__fmul_rn((float)d.cw0, d.sx[i]));
            const int c0 = d.col[i], r0 = d.row[i];
            if (br == rows && bc == cols && c0 >= 0 && r0 >= 0 && c0 + cols <= W && r0 + rows <= H) {
                x0 = min(x0, c0); y0 = min(y0, r0); x1 = max(x1, c0); y1 = max(y1, r0);
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
            x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
        if (lane == 0 && x1 >= x0) { atomicMin(&s_bb[0], x0); atomicMin(&s_bb[1], y0); atomicMax(&s_bb[2], x1); atomicMax(&s_bb[3], y1); }
    }
    __syncthreads();
    // ---- 2. shared-memory plan
    //   single group : LUT | slot map | region of the whole chunk | per-warp private histograms (what is left, <= MDS_WARPS)
    //   row bands    : LUT | slot map | row-sorted box list | row counters | region budget | histograms of `wm` warps
    // The second plan takes over when the chunk's bounding region leaves fewer than MDS_MIN_WARPS histograms (particles spread
    // over the frame, e.g. a tracker that lost its object): the conforming boxes are counting-sorted by their top row and cut
    // greedily into groups whose bounding region fits the budget; each group is staged and gathered like a small chunk.
    const int n_c = i_end - i_begin;
    const size_t lut_b = (size_t)(rows + 1) * colsp * 4;
    uint32_t* lut = (uint32_t*)smraw;                                         // [(rows + 1)][colsp], last row zero
    uint16_t* smap = (uint16_t*)(smraw + lut_b);                              // [dim]
    const size_t map_b = ((size_t)dim * 2 + 15) & ~(size_t)15;
    const size_t hist_b = (size_t)n_slots * 128;
    const bool usable = ref != 0x7fffffff && s_bb[2] >= s_bb[0] && n_slots <= 256 && T <= 32 && (W & 3) == 0;
    int warps_used = 0;
    uint8_t* region = smraw + lut_b + map_b;
    uint32_t* hist_all = nullptr;
    bool fast = false, banded = false;
    uint16_t *ord_i = nullptr, *ord_c = nullptr, *ord_r = nullptr;
    int* rowcnt = nullptr;
    size_t band_budget = 0;
    if (usable) {
        const int bx0 = s_bb[0] & ~3, by0 = s_bb[1];
        const int RW = s_bb[2] + colsp - bx0, RH = s_bb[3] + rows - by0 + 1;  // +1 row: the row prefetch stays inside
        const int RWp = ((RW + 3) & ~3) + 4;
        const size_t reg_b = ((size_t)RWp * RH + 15) & ~(size_t)15;
        if (lut_b + map_b + reg_b + hist_b <= (size_t)smem_bytes) {
            warps_used = min(MDS_WARPS, (int)(((size_t)smem_bytes - lut_b - map_b - reg_b) / hist_b));
            hist_all = (uint32_t*)(smraw + lut_b + map_b + reg_b);
            fast = true;
        }
        if ((!fast || warps_used < MDS_MIN_WARPS) && n_c < 65536 && W < 65536 && H < 65536) {
            const size_t ord_b = (((size_t)n_c * 6) + 15) & ~(size_t)15, cnt_b = (((size_t)H + 2) * 4 + 15) & ~(size_t)15;
            const size_t one_box = (size_t)(colsp + 8) * (rows + 1) + 16;
            for (int wm = MDS_MIN_WARPS; wm >= 1 && !banded; wm >>= 1) {
                const size_t fixed = lut_b + map_b + ord_b + cnt_b + (size_t)wm * hist_b;
                if (wm <= warps_used) break;                                   // the single group is at least as good
                if (fixed + 4 * one_box <= (size_t)smem_bytes) {
                    banded = true; fast = true;
                    band_budget = ((size_t)smem_bytes - fixed) & ~(size_t)15;
                    warps_used = wm;
                    ord_i = (uint16_t*)(smraw + lut_b + map_b); ord_c = ord_i + n_c; ord_r = ord_c + n_c;
                    rowcnt = (int*)(smraw + lut_b + map_b + ord_b);
                    region = smraw + lut_b + map_b + ord_b + cnt_b;
                    hist_all = (uint32_t*)(region + band_budget);
                }
            }
        }
    }
    if (!fast) {
        // everything in this chunk goes to the generic kernel
        for (int i = i_begin + tid; i < i_end; i += blockDim.x) {
            const bool v = d.valid == nullptr || d.valid[i] != 0;
            d.big[i] = v ? 1 : 0;
            if (v) atomicAdd(d.nbig, 1);
            else for (int o = 0; o < n_out; o++) d.featN[(size_t)o * d.ldN + i] = 0.0f;
        }
        return;
    }
    // ---- 3. LUT, slot map
    int shift = __clz(rows * cols); if (shift > MCT_FIX_SHIFT_MAX) shift = MCT_FIX_SHIFT_MAX;
    {
        const float scale = (float)(1u << shift);
        const float hx = __fdiv_rn((float)cols, 2.0f), hy = __fdiv_rn((float)rows, 2.0f);
        const double ivx = 1.0 / (double)__fmul_rn(2.0f, __fmul_rn(hx, hx));
        const double ivy = 1.0 / (double)__fmul_rn(2.0f, __fmul_rn(hy, hy));
        uint32_t part = 0;
        for (int p = tid; p < (rows + 1) * colsp; p += blockDim.x) {
            const int r = p / colsp, c = p - r * colsp;
            uint32_t w = 0;
            if (c < cols && r < rows) {
                const double dx = (double)__fsub_rn(hx, (float)c), dy = (double)__fsub_rn(hy, (float)r);
                const float wd = (float)(dx * dx * ivx + dy * dy * ivy);
                w = __float2uint_rn(__fmul_rn(expf(-wd), scale));
            }
            // four planes [k][r][T]: plane k holds the weights of column 4 j + k, so that the lanes of a box (and of the other boxes
            // of the warp, which read the same addresses) load consecutive words -- one wavefront per load; the former [r][colsp]
            // layout read 16 bytes per lane and cost 6 wavefronts per row (40-word rows wrap around the 32 banks)
            lut[(size_t)(c & 3) * (rows + 1) * T + r * T + (c >> 2)] = w;
            part += w;
        }
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_tot[warp] = part;
    }
    for (int b = tid; b < dim; b += blockDim.x) smap[b] = d.slotmap[b];
    if (banded) {
        // counting sort of the conforming boxes by top row; everything else is flagged for the generic pass right away
        for (int r = tid; r < H + 2; r += blockDim.x) rowcnt[r] = 0;
        if (tid == 0) s_nord = 0;
        __syncthreads();
        for (int i = i_begin + tid; i < i_end; i += blockDim.x) {
            const bool v = d.valid == nullptr || d.valid[i] != 0;
            bool okb = false;
            int r0 = 0;
            if (v) {
                const int br = __float2int_rn(__fmul_rn((float)d.ch0, d.sy[i]));
                const int bc = __float2int_rn(__fmul_rn((float)d.cw0, d.sx[i]));
                const int c0 = d.col[i]; r0 = d.row[i];
                okb = br == rows && bc == cols && c0 >= 0 && r0 >= 0 && c0 + cols <= W && r0 + rows <= H;
            }
            if (okb) { atomicAdd(&rowcnt[r0 + 1], 1); atomicAdd(&s_nord, 1); }
            else {
                d.big[i] = v ? 1 : 0;
                if (v) atomicAdd(d.nbig, 1);
                for (int o = 0; o < n_out; o++) d.featN[(size_t)o * d.ldN + i] = 0.0f;
            }
        }
        __syncthreads();
        if (warp == 0) {                                                      // rowcnt[r] := first list position of row r
            int carry = 0;
            for (int r0 = 0; r0 < H + 2; r0 += 32) {
                const int r = r0 + lane;
                int v = r < H + 2 ? rowcnt[r] : 0;
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
                if (r < H + 2) rowcnt[r] = carry + v;
                carry += __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncthreads();
        for (int i = i_begin + tid; i < i_end; i += blockDim.x) {
            if (d.valid != nullptr && d.valid[i] == 0) continue;
            const int br = __float2int_rn(__fmul_rn((float)d.ch0, d.sy[i]));
            const int bc = __float2int_rn(__fmul_rn((float)d.cw0, d.sx[i]));
            const int c0 = d.col[i], r0 = d.row[i];
            if (br == rows && bc == cols && c0 >= 0 && r0 >= 0 && c0 + cols <= W && r0 + rows <= H) {
                const int pos = atomicAdd(&rowcnt[r0], 1);
                ord_i[pos] = (uint16_t)(i - i_begin); ord_c[pos] = (uint16_t)c0; ord_r[pos] = (uint16_t)r0;
            }
        }
    }
    __syncthreads();
    uint32_t total = 0;
    for (int q = 0; q < MDS_WARPS; q++) total += s_tot[q];
    const float inv_scale = __fdiv_rn(1.0f, (float)(1u << shift));
    const float S = __fmul_rn((float)total, inv_scale);
    const int bw = lane / T, j = lane - bw * T;
    // ---- 4. per group: stage the region, then rounds of warps_used * BPW boxes
    int g_begin = banded ? 0 : i_begin;
    const int g_stop = banded ? s_nord : i_end;
    while (g_begin < g_stop) {
        int g_end = i_end;
        if (banded) {
            if (tid == 0) {
                // greedy cut of the row-sorted list: the largest prefix whose bounding region fits the budget
                int minc = ord_c[g_begin], maxc = minc;
                const int rfirst = ord_r[g_begin];
                int rlast = rfirst, k = g_begin + 1;
                for (; k < g_stop; k++) {
                    const int c = ord_c[k], r = ord_r[k];
                    const int nmin = min(minc, c), nmax = max(maxc, c);
                    const int rwp = ((nmax + colsp - (nmin & ~3) + 3) & ~3) + 4;
                    if ((size_t)rwp * (r + rows - rfirst + 1) + 16 > band_budget) break;
                    minc = nmin; maxc = nmax; rlast = r;
                }
                s_bb[0] = minc; s_bb[1] = rfirst; s_bb[2] = maxc; s_bb[3] = rlast; s_gend = k;
            }
            __syncthreads();
            g_end = s_gend;
        }
        const int bx0 = s_bb[0] & ~3, by0 = s_bb[1];
        const int RW = s_bb[2] + colsp - bx0, RH = s_bb[3] + rows - by0 + 1;
        const int RWp = ((RW + 3) & ~3) + 4;
        const int rstep = RWp >> 2;
        {
            const int qw = RWp >> 2;                                          // 4-pixel groups per region row
            uint32_t* r32 = reinterpret_cast<uint32_t*>(region);
            for (int g = tid; g < qw * RH; g += blockDim.x) {
                const int y = g / qw, xg = g - y * qw;
                const int yy = min(by0 + y, H - 1), xx = min(bx0 + 4 * xg, W - 4);
                const uint2 px = *reinterpret_cast<const uint2*>(binimg + (size_t)yy * W + xx);
                r32[g] = (uint32_t)smap[px.x & 0xffffu] | ((uint32_t)smap[px.x >> 16] << 8) |
                         ((uint32_t)smap[px.y & 0xffffu] << 16) | ((uint32_t)smap[px.y >> 16] << 24);
            }
        }
        __syncthreads();
        if (warp < warps_used) {
            uint32_t* hist = hist_all + (size_t)warp * n_slots * 32 + lane;   // element s at hist[s * 32]
            for (int base = g_begin + warp * BPW; base < g_end; base += warps_used * BPW) {
                for (int s = 0; s < n_slots; s++) hist[s * 32] = 0u;
                __syncwarp();
                const int k = base + bw;
                const bool mine = bw < BPW && k < g_end;
                const int i = mine ? (banded ? i_begin + (int)ord_i[k] : k) : 0;
                bool ok = false, valid = false;
                if (mine) {
                    valid = d.valid == nullptr || d.valid[i] != 0;
                    if (valid) {
                        const int br = __float2int_rn(__fmul_rn((float)d.ch0, d.sy[i]));
                        const int bc = __float2int_rn(__fmul_rn((float)d.cw0, d.sx[i]));
                        const int c0 = d.col[i], r0 = d.row[i];
                        ok = br == rows && bc == cols && c0 >= 0 && r0 >= 0 && c0 + cols <= W && r0 + rows <= H;
                        if (ok) {
                            const int addr = (r0 - by0) * RWp + (c0 - bx0) + 4 * j;
                            const uint32_t* lp = lut + j;
                            const int PS = (rows + 1) * T;                    // plane stride
                            const uint32_t* wp = reinterpret_cast<const uint32_t*>(region + (addr & ~3));
                            const int sh = (addr & 3) * 8;
                            uint32_t v = __funnelshift_r(wp[0], wp[1], sh);
                            uint4 w = make_uint4(lp[0], lp[PS], lp[2 * PS], lp[3 * PS]);
                            for (int r = 0; r < rows; r++) {
                                wp += rstep; lp += T;
                                const uint32_t vn = __funnelshift_r(wp[0], wp[1], sh);    // next row (row `rows` is padding)
                                const uint4 wn = make_uint4(lp[0], lp[PS], lp[2 * PS], lp[3 * PS]);
                                const uint32_t s0 = v & 0xffu, s1 = (v >> 8) & 0xffu, s2 = (v >> 16) & 0xffu, s3 = v >> 24;
                                const uint32_t h0 = hist[s0 * 32], h1 = hist[s1 * 32], h2 = hist[s2 * 32], h3 = hist[s3 * 32];
                                const uint32_t n0 = h0 + w.x;
                                const uint32_t n1 = (s1 == s0 ? n0 : h1) + w.y;
                                uint32_t t = (s2 == s0 ? n0 : h2); t = (s2 == s1 ? n1 : t);
                                const uint32_t n2 = t + w.z;
                                t = (s3 == s0 ? n0 : h3); t = (s3 == s1 ? n1 : t); t = (s3 == s2 ? n2 : t);
                                const uint32_t n3 = t + w.w;
                                hist[s0 * 32] = n0; hist[s1 * 32] = n1; hist[s2 * 32] = n2; hist[s3 * 32] = n3;
                                v = vn; w = wn;
                            }
                        }
                    }
                }
                __syncwarp();
                if (mine) {
                    if (j == 0) {
                        const int flag = (valid && !ok) ? 1 : 0;
                        d.big[i] = flag;
                        if (flag) atomicAdd(d.nbig, 1);
                    }
                    const uint32_t* hb = hist_all + (size_t)warp * n_slots * 32 + bw * T;
                    float* dst = d.featN + i;
                    if (!d.color_resp) {
                        for (int s = 1 + j; s < n_slots; s += T) {
                            uint32_t t = 0;
                            int qq = j;                                       // skewed start: the lanes of a box hit distinct banks
                            for (int k2 = 0; k2 < T; k2++) { t += hb[s * 32 + qq]; qq = (qq + 1 == T) ? 0 : qq + 1; }
                            dst[(size_t)(s - 1) * d.ldN] = ok ? __fdiv_rn(__fmul_rn((float)t, inv_scale), S) : 0.0f;
                        }
                    } else {
                        // fused ClassifyF of the selector that owns the bin: the bin value never leaves the SM, featN receives the
                        // response the classify kernel adds in selector order.  Two slots per step: their exp / log chains overlap.
                        const float4* rst = reinterpret_cast<const float4*>(d.rowst);
                        for (int s = 1 + j; s < n_slots; s += 2 * T) {
                            const int s2 = s + T;
                            const bool two = s2 < n_slots;
                            uint32_t t = 0, t2 = 0;
                            int qq = j;
                            for (int k2 = 0; k2 < T; k2++) {
                                t += hb[s * 32 + qq];
                                if (two) t2 += hb[s2 * 32 + qq];
                                qq = (qq + 1 == T) ? 0 : qq + 1;
                            }
                            const float4 a0 = __ldg(rst + 2 * (s - 1)), a1 = __ldg(rst + 2 * (s - 1) + 1);
                            const float4 b0 = __ldg(rst + 2 * (two ? s2 - 1 : s - 1)), b1 = __ldg(rst + 2 * (two ? s2 - 1 : s - 1) + 1);
                            const float x = __fdiv_rn(__fmul_rn((float)t, inv_scale), S), x2 = __fdiv_rn(__fmul_rn((float)t2, inv_scale), S);
                            const float r1 = weak_classify_q_dev(d.weak_type, x, a0.x, a0.y, a1.x, a1.y, a1.z, a1.w);
                            const float r2 = weak_classify_q_dev(d.weak_type, x2, b0.x, b0.y, b1.x, b1.y, b1.z, b1.w);
                            dst[(size_t)(s - 1) * d.ldN] = ok ? r1 : 0.0f;
                            if (two) dst[(size_t)(s2 - 1) * d.ldN] = ok ? r2 : 0.0f;
                        }
                    }
                }
                __syncwarp();
            }
        }
        g_begin = g_end;
        if (banded) __syncthreads();                                          // the next group rewrites s_bb and the region
    }
}

// Per-frame kernel of the colour features.  One persistent CTA per SM, the SMs split evenly over the objects, each CTA owns a
// contiguous chunk of the object's particles.
//   prologue (pre = 1 / 2): PredictWithBrownianMotion + GenerateTestSampleSet / GenerateTestSampleSet alone for the CTA's own
//       particles -- the per-particle steps need nothing from other particles, so the predict kernel's launch folds away;
//   gather: mdch_sel_chunk;
//   tail: the CTA that finishes an object last (arrival counter, no waiting) runs the generic pass over the boxes the fast
//       path flagged -- normally none, and the separate always-launched fallback kernel folds away as well.
__global__ void __launch_bounds__(MDS_WARPS * 32, 1)
k_mdch_sel(const ObjDesc* __restrict__ descs, const uint16_t* __restrict__ binimg, int W, int H, int smem_bytes, const StepArgs sa, int pre)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ ObjDesc s_desc;
    __shared__ int s_ticket;
    pdl_trigger();                            // k_classify_staged may take the SMs this kernel's CTAs leave and run up to its pdl_wait()
    const ObjDesc& d = load_desc(descs + blockIdx.y, &s_desc);
    const int tid = threadIdx.x;
    const int N = d.N;
    const int chunk = (N + gridDim.x - 1) / gridDim.x;
    const int i_begin = min(N, (int)blockIdx.x * chunk), i_end = min(N, i_begin + chunk);
    if (pre) {
        if (blockIdx.x == 0 && tid == 0 && pre == 1) { d.st->filter_state = MCT_PF_PREDICTED; d.st->time_index += 1; }
        int ok = 0;
        for (int i = i_begin + tid; i < i_end; i += blockDim.x) {
            if (pre == 1) {
                ok |= refine_valid_one(d, i) ? 1 : 0;
                predict_one(d, i, sa.noise_base + d.noise_off);
            }
            testset_one(d, i, W, H);
        }
        if (pre == 1) {
            const int any = __syncthreads_or(ok);
            if (tid == 0) atomicOr(&d.st->refine, MCT_REFINE_PENDING | (any ? MCT_REFINE_VALID : 0));
        }
        __syncthreads();                      // the chunk's boxes are read back by other threads of this CTA
    }
    if (d.slotmap == nullptr) return;         // object without colour features: nothing to gather, nobody runs a tail
    if (i_begin < i_end) mdch_sel_chunk(d, binimg, W, H, smem_bytes, smraw, i_begin, i_end);
    // ---- tail
    __syncthreads();
    if (tid == 0) { __threadfence(); s_ticket = atomicAdd(d.nbig + 1, 1); }
    __syncthreads();
    if (s_ticket != (int)gridDim.x - 1) return;
    if (tid == 0) d.nbig[1] = 0;              // ready for the next launch
    __threadfence();
    if (__ldcg(d.nbig) == 0) return;
    const int warp = tid >> 5;
    mdch_big_body(d, binimg, W, H, smraw, MDS_WARPS, warp, MDS_WARPS);
    __syncthreads();
    if (tid == 0) *d.nbig = 0;                // counted again from zero by the next launch
}

int launch_mdch_selected(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int max_slots, const StepArgs& sa, int pre)
{
    (void)max_slots;
    if (n_obj <= 0 || n_max <= 0) return MCT_OK;
    static size_t s_attr[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_mdch_sel, s_attr, MDS_SMEM, 0)) return rc_;
    const int s_sms = ctx->sms;
    // one persistent CTA per SM, the SMs split evenly over the objects; each CTA owns a contiguous chunk of boxes
    int per_obj = s_sms / n_obj;
    if (per_obj < 1) per_obj = 1;
    if (per_obj > ceil_div(n_max, 3)) per_obj = ceil_div(n_max, 3);
    dim3 g1(per_obj, n_obj);
    MCT_LAUNCH(ctx, "k_mdch_sel", k_mdch_sel<<<g1, MDS_WARPS * 32, MDS_SMEM, ctx->stream>>>(d_desc, ctx->binimg, ctx->W, ctx->H, MDS_SMEM, sa, pre));
    return MCT_OK;
}
