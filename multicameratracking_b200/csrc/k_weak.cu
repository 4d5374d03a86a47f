// k_weak.cu -- per-feature weak-classifier training + predictions (hot loop C):
//   for f in F: weak[f]->Update(pos, neg); pp[f] = ClassifySetF(pos); pn[f] = ClassifySetF(neg)
//   (MILBoostClassifier.cpp:53-59, MILAnyBoostClassifier.cpp:50-56).
// One warp per feature; means/variances are f64 shuffle reductions of the f32 feature row, like the
// IPP f64-accumulator calls the reference makes (Matrix.h:674-730).  Features rows are [F][ld] with the
// P positives first, then the Nn negatives; predictions use the same layout.
#include "mct_internal.cuh"

__device__ __forceinline__ float lerp_lr(float lr, float old, float nw)
{
    // m_learningRate*old + (1-m_learningRate)*new, every op in f32
    return __fadd_rn(__fmul_rn(lr, old), __fmul_rn(__fsub_rn(1.0f, lr), nw));
}
// Matrix::Mean (Matrix.h:674-683)
__device__ __forceinline__ double row_sum_d(const float* x, int n, int lane)
{
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += (double)x[i];
    return warp_sum_d(s);
}
// ((X - mu).Sqr()).Mean()   (OnlineStumpsWeakClassifier.cpp:118,124)
__device__ __forceinline__ float mean_sqdev(const float* x, int n, float mu, int lane)
{
    const float nm = __fmul_rn(mu, -1.0f);
    double s = 0.0;
    for (int i = lane; i < n; i += 32) { float d = __fadd_rn(x[i], nm); s += (double)__fmul_rn(d, d); }
    s = warp_sum_d(s);
    return (float)(s / (double)n);
}
// Matrix::Var (Matrix.h:685-694): population std-dev squared, in f64
__device__ __forceinline__ float row_var(const float* x, int n, double sum, int lane)
{
    const double m = sum / (double)n;
    double q = 0.0;
    for (int i = lane; i < n; i += 32) { double d = (double)x[i] - m; q += d * d; }
    q = warp_sum_d(q);
    const double sd = sqrt(q / (double)n);
    return (float)(sd * sd);
}
// Matrix::MeanW / VarW with normalised weights (Matrix.h:696-719); w == NULL -> 1.0f
__device__ __forceinline__ float row_meanw(const float* x, const float* w, float k, int n, int lane)
{
    double s = 0.0;
    for (int i = lane; i < n; i += 32) { float wn = __fmul_rn(w ? w[i] : 1.0f, k); s += (double)__fmul_rn(x[i], wn); }
    return (float)warp_sum_d(s);
}
__device__ __forceinline__ float row_varw(const float* x, const float* w, float k, int n, float mu, int lane)
{
    const float nm = __fmul_rn(mu, -1.0f);
    double s = 0.0;
    for (int i = lane; i < n; i += 32) {
        float wn = __fmul_rn(w ? w[i] : 1.0f, k);
        float d = __fadd_rn(x[i], nm);
        s += (double)__fmul_rn(__fmul_rn(d, d), wn);
    }
    return (float)warp_sum_d(s);
}
// Matrix::normalize scale (Matrix.h:540-544): (float)(1.0 / (Sum + 1e-6))
__device__ __forceinline__ float norm_scale(const float* w, int n, int lane)
{
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += (double)(w ? w[i] : 1.0f);
    s = warp_sum_d(s);
    return (float)(1.0 / (s + 1e-6));
}

__device__ __forceinline__ void
weak_update_body(const float* __restrict__ feat, int ld, int F, int P, int Nn, int weak_type, float lr,
              const float* __restrict__ tw, float* __restrict__ weak, int* __restrict__ trained,
              float* __restrict__ pred, const int* __restrict__ keep, int f_first = 0, int f_end = 0x7fffffff)
{
    const int lane = threadIdx.x & 31;
    const int f = f_first + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= F || f >= f_end) return;
    const float* xp = feat + (size_t)f * ld;
    const float* xn = xp + P;
    float mu0 = weak[0 * (size_t)F + f], mu1 = weak[1 * (size_t)F + f];
    float sig0 = weak[2 * (size_t)F + f], sig1 = weak[3 * (size_t)F + f];
    float n0 = 0, n1 = 0, e0 = 0, e1 = 0;
    int was_trained = trained[f];
    // MILEnsemble (MILEnsembleClassifier.cpp:56-72): a retained weak classifier keeps its state and is only evaluated;
    // every other one is re-initialised (mu = 0, sig = 1, untrained) and trained from scratch on this set
    const bool kept = keep != nullptr && keep[f] != 0;
    if (keep != nullptr && !kept) { mu0 = 0.0f; mu1 = 0.0f; sig0 = 1.0f; sig1 = 1.0f; was_trained = 0; }

    if (kept) {
        // state as stored, derived rows included: the classifier a fresh MILEnsemble retains first (K copies of feature 0,
        // StrongClassifierBase.cpp:91) was never trained; its m_n / m_e are uninitialised memory in the reference
        // (OnlineStumpsWeakClassifier.cpp:53-60), zero here and in the oracle
        n0 = weak[4 * (size_t)F + f]; n1 = weak[5 * (size_t)F + f]; e0 = weak[6 * (size_t)F + f]; e1 = weak[7 * (size_t)F + f];
    } else if (weak_type == MCT_WEAK_STUMP) {
        // OnlineStumpsWeakClassifier::Update (OnlineStumpsWeakClassifier.cpp:99-151)
        const double sp = row_sum_d(xp, P, lane), sn = row_sum_d(xn, Nn, lane);
        const float pm = P > 0 ? (float)(sp / (double)P) : 0.0f;
        const float nm = Nn > 0 ? (float)(sn / (double)Nn) : 0.0f;
        if (was_trained) {
            if (P > 0) { mu1 = lerp_lr(lr, mu1, pm); sig1 = lerp_lr(lr, sig1, mean_sqdev(xp, P, mu1, lane)); }
            if (Nn > 0) { mu0 = lerp_lr(lr, mu0, nm); sig0 = lerp_lr(lr, sig0, mean_sqdev(xn, Nn, mu0, lane)); }
        } else {
            if (P > 0) { mu1 = pm; sig1 = __fadd_rn(row_var(xp, P, sp, lane), 1e-9f); }
            if (Nn > 0) { mu0 = nm; sig0 = __fadd_rn(row_var(xn, Nn, sn, lane), 1e-9f); }
        }
    } else if (weak_type == MCT_WEAK_WSTUMP) {
        // WeightedStumpsWeakClassifier::Update (WeightedStumpsWeakClassifier.cpp:60-118)
        const float* wp = tw; const float* wn = tw ? tw + P : nullptr;
        const float kp = norm_scale(wp, P, lane), kn = norm_scale(wn, Nn, lane);
        const float posmu = P > 0 ? row_meanw(xp, wp, kp, P, lane) : 0.0f;
        const float negmu = Nn > 0 ? row_meanw(xn, wn, kn, Nn, lane) : 0.0f;
        if (was_trained) {
            if (P > 0) { mu1 = lerp_lr(lr, mu1, posmu); sig1 = lerp_lr(lr, sig1, row_varw(xp, wp, kp, P, mu1, lane)); }
            if (Nn > 0) { mu0 = lerp_lr(lr, mu0, negmu); sig0 = lerp_lr(lr, sig0, row_varw(xn, wn, kn, Nn, mu0, lane)); }
        } else {
            mu1 = posmu; mu0 = negmu;
            if (Nn > 0) sig0 = __fadd_rn(row_varw(xn, wn, kn, Nn, negmu, lane), 1e-9f);
            if (P > 0) sig1 = __fadd_rn(row_varw(xp, wp, kp, P, posmu, lane), 1e-9f);
        }
    } else {
        // PerceptronWeakClassifier::Update (PerceptronWeakClassifier.cpp:98-156): order dependent ->
        // one lane runs it; rows 0,1 of the state hold m_weightList[0], [1].
        float w0 = 0.0f; const float w1 = 0.5f;
        if (lane == 0) {
            unsigned it = 0;
            const unsigned total = (unsigned)(P + Nn);
            const int maxAllowed = P < Nn ? P : Nn;
            int prevErr = maxAllowed;
            while (it < 100) {
                unsigned err = 0;
                for (int i = 0; i < P; i++) {
                    double fv = (double)xp[i];
                    int result = (fv * (double)w0) > (double)w1 ? 1 : 0;
                    int e = 1 - result;
                    if (e != 0) { w0 = (float)((double)w0 + 0.1 * (double)e * fv); err++; }
                }
                for (int i = 0; i < Nn; i++) {
                    double fv = (double)xn[i];
                    int result = (fv * (double)w0) > (double)w1 ? 1 : 0;
                    int e = 0 - result;
                    if (e != 0) { w0 = (float)((double)w0 + 0.1 * (double)e * fv); err++; }
                }
                float improvement = __fdiv_rn((float)((unsigned)prevErr - err), (float)total);
                if (err < (unsigned)maxAllowed && (double)improvement < 0.01) break;
                prevErr = (int)err;
                it++;
            }
        }
        w0 = __shfl_sync(0xffffffffu, w0, 0);
        mu0 = w0; mu1 = w1;
    }
    if (weak_type != MCT_WEAK_PERCEPTRON && !kept) {
        n0 = __fdiv_rn(1.0f, __fsqrt_rn(sig0));          // 1.0f / pow(sig, 0.5f)
        n1 = __fdiv_rn(1.0f, __fsqrt_rn(sig1));
        e1 = __fdiv_rn(-1.0f, __fmul_rn(2.0f, sig1));    // the reference's +1e-99f is 0 in f32
        e0 = __fdiv_rn(-1.0f, __fmul_rn(2.0f, sig0));
    }
    if (lane == 0) {
        weak[0 * (size_t)F + f] = mu0; weak[1 * (size_t)F + f] = mu1;
        weak[2 * (size_t)F + f] = sig0; weak[3 * (size_t)F + f] = sig1;
        weak[4 * (size_t)F + f] = n0; weak[5 * (size_t)F + f] = n1;
        weak[6 * (size_t)F + f] = e0; weak[7 * (size_t)F + f] = e1;
        trained[f] = 1;
    }
    // predictions (ClassifySetF) for positives then negatives.  A feature row that holds one value for every sample -- most of
    // the nb^3 joint colour bins are empty in every training box -- has one prediction: evaluated once instead of P + Nn times
    // (the kernel is bound by instruction issue, and the f64 log of the response is most of a prediction's ~150 instructions)
    float* pr = pred + (size_t)f * ld;
    const float x0 = xp[0];
    bool same = true;
    for (int i = lane; i < P + Nn; i += 32) same = same && (xp[i] == x0);
    if (__all_sync(0xffffffffu, same)) {
        const float v = weak_classify_q_dev(weak_type, x0, mu0, mu1, n0, n1, e0, e1);
        for (int i = lane; i < P + Nn; i += 32) pr[i] = v;
    } else {
        for (int i = lane; i < P + Nn; i += 32)
            pr[i] = weak_classify_q_dev(weak_type, xp[i], mu0, mu1, n0, n1, e0, e1);
    }
}

__global__ void __launch_bounds__(256)
k_weak_update(const float* __restrict__ feat, int ld, int F, int P, int Nn, int weak_type, float lr,
              const float* __restrict__ tw, float* __restrict__ weak, int* __restrict__ trained,
              float* __restrict__ pred, const int* __restrict__ keep)
{
    weak_update_body(feat, ld, F, P, Nn, weak_type, lr, tw, weak, trained, pred, keep);
}
// batched form: blockIdx.y = object
// part: 0 = every weak classifier, 1 = the Haar features' [0, n_haar), 2 = the colour features' [n_haar, F).  The two feature
// families' training rows come from two streams (launch_train_features); the weak classifiers of a family only need their own
// rows, so each part rides behind its rows and the Haar part (the expensive one: no constant rows) runs beside the colour kernels.
__global__ void __launch_bounds__(256) k_weak_update_batch(const TrainDesc* __restrict__ descs, int part)
{
    const TrainDesc& d = descs[blockIdx.y];
    if (d.P < 1 || d.Nn < 1) return;              // empty training set decided on the device (k_train_gen_default): classifier left as it is
    const int f_first = part == 2 ? d.n_haar : 0, f_end = part == 1 ? d.n_haar : d.F;
    weak_update_body(d.featT, d.ldT, d.F, d.P, d.Nn, d.weak_type, d.lr, d.tw, d.weak, d.trained, d.pred, d.keep, f_first, f_end);
}
int launch_weak_update_part(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj, int part, cudaStream_t st)
{
    int rows = 0;
    for (int o = 0; o < n_obj; o++) rows = max(rows, part == 1 ? h[o].n_haar : (part == 2 ? h[o].F - h[o].n_haar : h[o].F));
    if (rows <= 0) return MCT_OK;
    dim3 g(ceil_div(rows, 8), n_obj);
    MCT_LAUNCH_ON(ctx, st, "k_weak_update_batch", k_weak_update_batch<<<g, 256, 0, st>>>(d, part));
    return MCT_OK;
}
int launch_weak_update_batch(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj)
{
    if (ctx->weak_split_done) { ctx->weak_split_done = 0; return MCT_OK; }     // already queued behind the feature rows, in two parts
    return launch_weak_update_part(ctx, d, h, n_obj, 0, ctx->stream);
}

int launch_weak_update(mct_clf* clf, int P, int Nn, int has_weights)
{
    mct_ctx* ctx = clf->ctx;
    MCT_LAUNCH(ctx, "k_weak_update", k_weak_update<<<ceil_div(clf->F, 8), 256, 0, ctx->stream>>>(clf->d_featT, clf->ldT, clf->F, P, Nn, clf->p.weak_type,
                                                               clf->p.learning_rate, has_weights ? clf->d_tw : nullptr,
                                                               clf->d_weak, clf->d_trained, clf->d_pred,
                                                               clf->p.strong_type == MCT_STRONG_MILENSEMBLE ? clf->d_keep : nullptr));
    return MCT_OK;
}
