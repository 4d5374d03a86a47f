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

struct SelEntry {
    int4  r[MCT_MAX_RECTS];
    float w[MCT_MAX_RECTS];
    float st[8];
    int   nrect;      // >0: Haar feature; 0: colour feature, `row` indexes featN
    int   row;
};

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
        Hs = __fadd_rn(Hs, weak_classify_dev(d.weak_type, x, e.st[0], e.st[1], e.st[4], e.st[5], e.st[6], e.st[7]));
    }
    if (!log_ratio) Hs = sigmoid_dev(Hs);
    d.H[i] = Hs;
}

int launch_classify(mct_ctx* ctx, const ObjDesc* d_desc, int n_obj, int n_max, int K_max, int from_matrix, int log_ratio)
{
    if (n_obj <= 0 || n_max <= 0) return MCT_OK;
    size_t smem = (size_t)(K_max > 0 ? K_max : 1) * sizeof(SelEntry);
    static size_t s_attr = 48 * 1024;
    if (smem > s_attr) {
        MCT_CUDA(ctx, cudaFuncSetAttribute(k_classify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s_attr = smem;
    }
    dim3 g(ceil_div(n_max, 256), n_obj);
    MCT_LAUNCH(ctx, "k_classify", k_classify<<<g, 256, smem, ctx->stream>>>(d_desc, ctx->ii_base ? ctx->ii_base + MCT_II_LEAD : nullptr, ctx->ii_pitch,
                                              ctx->W, ctx->H, from_matrix, log_ratio));
    return MCT_OK;
}
