// k_milselect.cu -- greedy weak-learner selection (hot loop D).
//
// k_milboost_select: MILBoostClassifier::Update selection loop (MILBoostClassifier.cpp:61-131).
//   K sequential rounds, each scoring all F candidates over the P positives / Nn negatives:
//       Lp  = prod_i (1 - sigmoid(Hp_i + pp[f][i]))                       f32 product, in order
//       nll = (float)-log(1 - Lp + 1e-5)/P + sum_j (float)-logf(1e-5f + 1 - sigmoid(Hn_j + pn[f][j]))/Nn
//   then argmin over the not-yet-selected valid candidates, stop when the NLL stops improving.
//   The whole loop runs inside ONE kernel on a thread-block cluster (one cluster per classifier):
//   the candidates are partitioned over the CTAs of the cluster; per round every CTA
//     phase 1  evaluates the sigmoid/log terms for its candidates on all threads (coalesced over the
//              samples) into a transposed shared-memory tile,
//     phase 2  lets one thread per candidate run the product / sum chains *in reference order* (so the
//              bag likelihood is bit-identical to a sequential evaluation given the same terms),
//     then publishes its best candidate into every peer's shared memory (DSMEM), one cluster.sync()
//     later every CTA holds all candidates' bests and takes the same argmin (ties -> lowest index,
//     the stable-sort order; the reference's std::sort tie order is unspecified, Public.h:223).
//   SFU/latency bound (exp, rcp, log per term), not a bandwidth kernel; the prediction matrix
//   [F][P+Nn] (a few hundred KB) stays L2 resident.
//
// k_anyboost_select: MILAnyBoostClassifier::Update selection loop (MILAnyBoostClassifier.cpp:78-222),
//   one CTA per classifier (one sigmoid pass + F dot products per round).
#include "mct_internal.cuh"
#include <cooperative_groups.h>
#include <math_constants.h>
namespace cg = cooperative_groups;

#define SEL_THREADS 1024
#define SEL_MAX_CLUSTER 16

struct SelSlot { float nll; int f; };

// Order of the candidates = the reference's sort_order (Public.h:211-228) with its two unspecified points pinned: exactly tied
// scores keep index order (a stable sort), and NaN scores -- the appearance fuser's classifiers (learning rate 0) get sigma = 0
// for colour bins that are empty in a class, hence n = inf, e = -inf and NaN responses -- rank LAST (std::sort on them is
// undefined behaviour in the reference; oracle/ref_shim/boost/shared_ptr.hpp gives the reference build the same order).
// f == 0x7fffffff marks "no candidate yet".
__device__ __forceinline__ bool better_min(float v, int f, float bv, int bf)
{
    if (f == 0x7fffffff) return false;
    if (bf == 0x7fffffff) return true;
    const bool vn = v != v, bn = bv != bv;
    if (vn | bn) return bn && (!vn || f < bf);
    return (v < bv) || (v == bv && f < bf);
}

// warp-wide arg-min in the order of better_min with two redux instructions instead of a five-step shuffle butterfly: the score
// is mapped to a u32 whose unsigned order is the scores' order (NaN last, -0 == +0), ties and the "no candidate" mark
// (f == 0x7fffffff, which loses against everything) are settled by the index; the score's own bits come from the winning lane.
__device__ __forceinline__ unsigned score_key(float v, int f)
{
    if (f == 0x7fffffff || v != v) return 0xffffffffu;
    if (v == 0.0f) return 0x80000000u;
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void warp_argmin(float& v, int& f)
{
    const unsigned key = score_key(v, f);
    const unsigned m = __reduce_min_sync(0xffffffffu, key);
    const unsigned fm = __reduce_min_sync(0xffffffffu, key == m ? (unsigned)f : 0x7fffffffu);
    const unsigned who = __ballot_sync(0xffffffffu, key == m && (unsigned)f == fm);
    const int src = who ? __ffs(who) - 1 : 0;
    v = __shfl_sync(0xffffffffu, v, src);
    f = (int)fm;
}

// terms per pass and candidate: MIL_RN negatives + up to MIL_PC positives, kept per candidate in one padded row of
// MIL_RS floats (MIL_RS % 32 == 4: the float4 reads of 8 neighbouring candidates cover the 32 banks once)
#define MIL_RN 32
#define MIL_PC 128
#define MIL_ROWS (MIL_RN + MIL_PC)
#define MIL_RS (MIL_ROWS + 4)
#define MIL_FC 128

#ifdef MCT_MIL_DEBUG
// the read of barrier-protected state in front of the clock makes the stamp see the barrier's release, not its issue
#define MIL_STAMP(k) { int v_ = *(volatile int*)&s_nact[0]; if (v_ == 0x7fffffff) dbg_pass += 1000; long long c_ = clock64(); dbg_t[k] += c_ - dbg_c; dbg_c = c_; }
#else
#define MIL_STAMP(k)
#endif

template <bool ROWS_SMEM>
__device__ __forceinline__ void
milboost_select_body(const float* __restrict__ pred, int ld, int F, int P, int Nn, int K,
                  const float* __restrict__ weak, int weak_type, int* __restrict__ sel, int* __restrict__ nsel,
                  int FC, int PC, int RN)
{
    // RN negatives per pass (a multiple of 32, MIL_RN for the trackers' own classifiers): the appearance fuser's pooled sets hold
    // the negatives of every camera (8 x 30), and with 32 per pass a round was eight barrier-separated passes over rows that
    // come from L2 (29 us per round); the pass partition does not touch the order of the chains
    const int RS = RN + MIL_PC + 4;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int M = P + Nn;
    const int FL = (F + C - 1) / C;
    extern __shared__ __align__(16) unsigned char smraw[];
    float* scratch = (float*)smraw;                          // [FC][MIL_RS] terms of the current pass, one row per list entry
    int* act = (int*)(scratch + (size_t)FC * RS);        // [2][FC] compacted candidate lists (slot | need_pos << 16)
    float* Hs = (float*)(act + 2 * FC);                      // [M]
    float* rows = Hs + M;                                    // [FL][M] this CTA's candidate rows (ROWS_SMEM)
    unsigned char* selflag = (unsigned char*)(rows + (ROWS_SMEM ? (size_t)FL * M : 0));   // [F] 1 selected, 2 invalid
    __shared__ SelSlot slots[2][SEL_MAX_CLUSTER];
    __shared__ SelSlot s_cand[MIL_FC];
    __shared__ int s_nact[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    // candidate l of this CTA is feature rank + l * C: every CTA gets the same mix of feature families (with contiguous ranges the
    // CTAs that owned the Haar features carried 48 live candidates per round and the ones with colour bins, many of them empty
    // in both classes and therefore invalid, a dozen: the cluster barrier waited for the former)
    const int nl = F > rank ? (F - rank + C - 1) / C : 0;

    for (int i = tid; i < M; i += blockDim.x) Hs[i] = 0.0f;
    for (int f = tid; f < F; f += blockDim.x) {
        unsigned char fl = 0;
        if (weak_type != MCT_WEAK_PERCEPTRON && weak[0 * (size_t)F + f] == weak[1 * (size_t)F + f]) fl = 2;   // IsValidWeakClassifier: mu0 != mu1
        selflag[f] = fl;
    }
    if (ROWS_SMEM)
        for (int l = warp; l < nl; l += nwarps) {
            const float* pr = pred + (size_t)(rank + l * C) * ld;
            float* dst = rows + l * M;
            for (int i = lane; i < M; i += 32) dst[i] = pr[i];
        }
    if (tid == 0) { s_nact[0] = 0; s_nact[1] = 0; }
    __syncthreads();

    // (float)-log(1 - like + 1e-5) of a saturated bag (1 - like == 1.0f): the f64 log leaves the per-round critical path
    const float posl_sat = (float)(-log((double)1.0f + 1e-5));
    const float fP = (float)P, fN = (float)Nn;
#ifdef MCT_MIL_DEBUG
    long long dbg_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long dbg_c = 0; int dbg_pass = 0;
#endif
    double prev = 1000000.0;
    int n_sel = 0;
    int g = 0;                  // running pass counter: list / counter parity continues across chunks and rounds
    for (int s = 0; s < K; s++) {
        float bv = CUDART_INF_F; int bf = 0x7fffffff;
#ifdef MCT_MIL_DEBUG
        dbg_c = clock64();
#endif
        for (int fc = 0; fc < nl; fc += FC) {
            const int nfc = min(FC, nl - fc);
            // The bag likelihood is the ordered f32 product of (1 - sigmoid) over the positives: it never grows, and once
            // 1 - like rounds to 1.0f the later factors cannot change the NLL any more, so the positives are evaluated
            // PC at a time and a candidate leaves the pass list as soon as its product saturates (bit-identical to the
            // full chain).  The negatives' sum always runs to the end.
            // Both counters are zero here (the previous list ran empty), and no thread can still be reading counter g & 1
            // of an earlier pass: the last read was of the other parity.
            float like = 1.0f, sn = 0.0f;
            bool active = false, sat = false;
            int col = 0;
            if (tid < ((nfc + 31) & ~31)) {                     // whole warps: the list slots are handed out by ballot
                active = tid < nfc && !selflag[rank + (fc + tid) * C];
                const unsigned bal = __ballot_sync(0xffffffffu, active);
                int base = 0;
                if (lane == 0 && bal) base = atomicAdd(&s_nact[g & 1], __popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (active) {
                    col = base + __popc(bal & ((1u << lane) - 1u));
                    act[(g & 1) * FC + col] = tid | (1 << 16);
                }
            }
            for (int it = 0;; it++, g++) {
                __syncthreads();
                const int nact = s_nact[g & 1];
                if (nact == 0) { g++; break; }
#ifdef MCT_MIL_DEBUG
                dbg_pass++;
#endif
                MIL_STAMP(4)
                // the first pass takes PC positives (most bags saturate inside it); a bag that did not saturate is unlikely to do so
                // soon, so the later passes take MIL_PC at a time: fewer passes (barriers) when the classes do not separate
                const int pcur = it == 0 ? PC : MIL_PC;
                const int n0 = it * RN, p0 = it == 0 ? 0 : PC + (it - 1) * MIL_PC;
                const int cntN = max(0, min(RN, Nn - n0)), cntP = max(0, min(pcur, P - p0));
                const int nbN = (cntN + 31) >> 5, nbP = (cntP + 31) >> 5, nblk = nbN + nbP;
                const unsigned magic = 65536u / (unsigned)nblk + 1u;        // u / nblk exactly for u < 65536 / nblk (u < 128 * 12 here)
                const int* al = act + (g & 1) * FC;
                // phase 1: terms, one warp per (list entry, 32 samples); a block's tail is filled with the chain's neutral
                // element so that phase 2 runs whole blocks.  Three units per trip: their exp / reciprocal chains are
                // independent and interleave (the pass is bound by the latency of those chains, not by issue slots).
                const int nunits = nact * nblk;
                for (int u0 = warp; u0 < nunits; u0 += 3 * nwarps) {
                    float xv[3]; int kind[3]; float* dp[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const int u = u0 + j * nwarps;
                        kind[j] = 0; xv[j] = 0.0f; dp[j] = scratch;
                        if (u < nunits) {
                            const int a = (int)(((unsigned)u * magic) >> 16), bb = u - a * nblk;
                            const int e = al[a];
                            const int slot = e & 0xffff;
                            const float* pr = ROWS_SMEM ? rows + (fc + slot) * M : pred + (size_t)(rank + (fc + slot) * C) * ld;
                            float* dst = scratch + a * RS;
                            if (bb < nbN) {
                                const int r = bb * 32 + lane;
                                kind[j] = (r < cntN) ? 1 : 3;                 // 3: tail of a negative block -> 0
                                dp[j] = dst + r;
                                if (r < cntN) { const int i = P + n0 + r; xv[j] = __fadd_rn(Hs[i], pr[i]); }
                            } else if (e >> 16) {
                                const int r = (bb - nbN) * 32 + lane;
                                kind[j] = (r < cntP) ? 2 : 4;                 // 4: tail of a positive block -> 1
                                dp[j] = dst + RN + r;
                                if (r < cntP) { const int i = p0 + r; xv[j] = __fadd_rn(Hs[i], pr[i]); }
                            }
                        }
                    }
                    float sg[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) sg[j] = sigmoid_dev(xv[j]);
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        if (kind[j] == 0) continue;
                        float t;
                        if (kind[j] == 1) t = -logf(__fsub_rn(__fadd_rn(1e-5f, 1.0f), sg[j]));
                        else if (kind[j] == 2) t = __fsub_rn(1.0f, sg[j]);
                        else t = (kind[j] == 3) ? 0.0f : 1.0f;
                        *dp[j] = t;
                    }
                }
                MIL_STAMP(5)
                __syncthreads();
                MIL_STAMP(6)
                if (tid == 0) s_nact[g & 1] = 0;                // becomes the next-list counter of pass g + 1
                // phase 2: ordered chains, one thread per candidate; the sum and the product are independent chains and
                // advance together
                if (tid < ((nfc + 31) & ~31)) {
                    bool again = false, more_p = false;
                    if (active) {
                        const float4* n4 = reinterpret_cast<const float4*>(scratch + col * RS);
                        const float4* p4 = n4 + RN / 4;
                        const bool doP = !sat && nbP > 0;
                        if (nbN && doP) {
#pragma unroll
                            for (int k = 0; k < 8; k++) {
                                const float4 x = n4[k], y = p4[k];
                                sn = __fadd_rn(sn, x.x); like = __fmul_rn(like, y.x);
                                sn = __fadd_rn(sn, x.y); like = __fmul_rn(like, y.y);
                                sn = __fadd_rn(sn, x.z); like = __fmul_rn(like, y.z);
                                sn = __fadd_rn(sn, x.w); like = __fmul_rn(like, y.w);
                            }
                        } else if (nbN) {
#pragma unroll
                            for (int k = 0; k < 8; k++) {
                                const float4 x = n4[k];
                                sn = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sn, x.x), x.y), x.z), x.w);
                            }
                        } else if (doP) {
#pragma unroll
                            for (int k = 0; k < 8; k++) {
                                const float4 y = p4[k];
                                like = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(like, y.x), y.y), y.z), y.w);
                            }
                        }
                        for (int k = 8; k < nbN * 8; k++) {           // (more than 32 negatives in the pass)
                            const float4 x = n4[k];
                            sn = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sn, x.x), x.y), x.z), x.w);
                        }
                        if (doP) {
                            for (int k = 8; k < nbP * 8; k++) {
                                const float4 y = p4[k];
                                like = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(like, y.x), y.y), y.z), y.w);
                            }
                            sat = __fsub_rn(1.0f, like) == 1.0f;
                        }
                        // A NaN in either chain stays a NaN to the end and makes the candidate's score a NaN whatever the remaining
                        // terms are: the candidate is finished.  (The appearance fuser's global classifiers have learning rate 0: a
                        // colour bin that is empty in a class has sigma = 0 and NaN responses.  Such a bag never saturates, and
                        // every NaN candidate walked all ~2500 pooled positives in passes of 128, every round: 29 us per round.)
                        const bool dead = (like != like) || (sn != sn);
                        more_p = !sat && !dead && (p0 + pcur < P);
                        again = !dead && ((n0 + RN < Nn) || more_p);
                        if (!again) {
                            active = false;
                            const int f = rank + (fc + tid) * C;
                            const float posl = sat ? posl_sat : (float)(-log((double)__fsub_rn(1.0f, like) + 1e-5));
                            const float nll = dead ? CUDART_NAN_F : __fadd_rn(__fdiv_rn(posl, fP), __fdiv_rn(sn, fN));
                            if (better_min(nll, f, bv, bf)) { bv = nll; bf = f; }
                        }
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, again);
                    if (bal) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&s_nact[(g + 1) & 1], __popc(bal));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (again) {
                            col = base + __popc(bal & ((1u << lane) - 1u));
                            act[((g + 1) & 1) * FC + col] = tid | ((int)more_p << 16);
                        }
                    }
                }
                MIL_STAMP(7)
            }
        }
        MIL_STAMP(0)
        // CTA argmin: the owning threads park their best in shared memory, warp 0 reduces and publishes
        // (one candidate per thread: the warps that own candidates reduce their registers, warp 0 combines the few warp bests)
        const int ncw = (min(FC, FL) + 31) >> 5;
        if (warp < ncw) {
            float v = bv; int f = bf;
            warp_argmin(v, f);
            if (lane == 0) { s_cand[warp].nll = v; s_cand[warp].f = f; }
        }
        __syncthreads();
        const int par = s & 1;
        if (warp == 0) {
            float v = CUDART_INF_F; int f = 0x7fffffff;
            if (lane < ncw) { v = s_cand[lane].nll; f = s_cand[lane].f; }
            if (ncw > 1) warp_argmin(v, f);
            else { v = __shfl_sync(0xffffffffu, v, 0); f = __shfl_sync(0xffffffffu, f, 0); }
            if (lane < C) {
                // publish into peer `lane`'s slot table (distributed shared memory)
                SelSlot* remote = cluster.map_shared_rank(&slots[0][0], lane);
                remote[par * SEL_MAX_CLUSTER + rank].nll = v;
                remote[par * SEL_MAX_CLUSTER + rank].f = f;
            }
        }
        MIL_STAMP(1)
        cluster.sync();
        MIL_STAMP(2)
        // every thread reduces the C published bests the same way (2 slots per lane, then a butterfly)
        float gv = CUDART_INF_F; int gf = 0x7fffffff;
        if (lane < C) { gv = slots[par][lane].nll; gf = slots[par][lane].f; }
        warp_argmin(gv, gf);
        if (gf == 0x7fffffff) break;                               // nothing eligible (reference: UB read)
        if ((prev - (double)gv) < 1e-100) break;                   // :108-112 no improvement -> pop & stop
        prev = (double)gv;
        if (rank == 0 && tid == 0) sel[n_sel] = gf;
        n_sel++;
        // the winner's row comes from L2: all CTAs of the cluster read the same 4 M bytes (a DSMEM read of the owner's copy
        // queues 16 readers on one SM's shared-memory port and measured slower)
        const float* pb = pred + (size_t)gf * ld;
        for (int i = tid; i < M; i += blockDim.x) Hs[i] = __fadd_rn(Hs[i], __ldg(pb + i));
        if (tid == 0) selflag[gf] = 1;
        __syncthreads();
        MIL_STAMP(3)
    }
#ifdef MCT_MIL_DEBUG
    if (tid == 0 && rank == 0 && blockIdx.y == 0)
        printf("mil dbg: rounds %d passes %d | passes %lld argmin %lld csync %lld hs %lld clk | A %lld ph1 %lld B %lld ph2 %lld\n", n_sel, dbg_pass,
               dbg_t[0], dbg_t[1], dbg_t[2], dbg_t[3], dbg_t[4], dbg_t[5], dbg_t[6], dbg_t[7]);
#endif
    if (rank == 0 && tid == 0) *nsel = n_sel;
    cluster.sync();          // no CTA may exit while a peer can still write its slots
}
__global__ void __launch_bounds__(SEL_THREADS)
k_milboost_select(const float* __restrict__ pred, int ld, int F, int P, int Nn, int K,
                  const float* __restrict__ weak, int weak_type, int* __restrict__ sel, int* __restrict__ nsel,
                  int FC, int rows_in_smem, int PC, int RN)
{
    if (rows_in_smem) milboost_select_body<true>(pred, ld, F, P, Nn, K, weak, weak_type, sel, nsel, FC, PC, RN);
    else milboost_select_body<false>(pred, ld, F, P, Nn, K, weak, weak_type, sel, nsel, FC, PC, RN);
}
// batched form: one cluster per object (blockIdx.y); objects of another strong-classifier type leave at once
// (the whole cluster takes the same branch)
__global__ void __launch_bounds__(SEL_THREADS)
k_milboost_select_batch(const TrainDesc* __restrict__ descs, int FC, int FLmax, int Mmax, int PC, int RN)
{
    const TrainDesc& d = descs[blockIdx.y];
    if (d.strong_type != MCT_STRONG_MILBOOST) return;
    if (d.P < 1 || d.Nn < 1) return;              // empty training set decided on the device (k_train_gen_default): classifier left as it is
    // the shared-memory plan was sized for (FLmax, Mmax): an object's rows fit iff its own FL * M does
    const int FL = (d.F + (int)gridDim.x - 1) / (int)gridDim.x;
    const int rows_in_smem = FLmax > 0 && (size_t)FL * (d.P + d.Nn) <= (size_t)FLmax * Mmax && min(FC, FL) <= FC;
    const int fc = min(FC, max(FL, 1));
    if (rows_in_smem) milboost_select_body<true>(d.pred, d.ldT, d.F, d.P, d.Nn, d.K, d.weak, d.weak_type, d.sel, d.nsel, fc, PC, RN);
    else milboost_select_body<false>(d.pred, d.ldT, d.F, d.P, d.Nn, d.K, d.weak, d.weak_type, d.sel, d.nsel, fc, PC, RN);
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool better_max(float v, int f, float bv, int bf)
{
    if (f == 0x7fffffff) return false;
    if (bf == 0x7fffffff) return true;
    const bool vn = v != v, bn = bv != bv;
    if (vn | bn) return bn && (!vn || f < bf);            // NaN scores last in the descending order too (sort_order_des)
    return (v > bv) || (v == bv && f < bf);
}
// block argmax over (value desc, index asc); returns through shared
__device__ void block_argmax(float v, int f, float* out_v, int* out_f, SelSlot* sh)
{
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int of = __shfl_xor_sync(0xffffffffu, f, o);
        if (better_max(ov, of, v, f)) { v = ov; f = of; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sh[warp].nll = v; sh[warp].f = f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++)
            if (better_max(sh[w].nll, sh[w].f, sh[0].nll, sh[0].f)) sh[0] = sh[w];
    }
    __syncthreads();
    *out_v = sh[0].nll; *out_f = sh[0].f;
    __syncthreads();
}

// MILAnyBoost on a thread-block cluster (round 2; one CTA per classifier before: every round re-read the F x M prediction
// matrix from L2 with 32 distinct lines per warp load and ran the F ordered objective chains on one SM, ~20 us per round).
// The F candidates are split over the C CTAs of the cluster; a CTA stages its candidates' prediction rows ONCE in shared memory
// (row pitch odd: lane = candidate reads conflict-free, the weight is a broadcast) -- the rows do not change across the rounds.
// Every CTA keeps its own copy of the hypothesis H and repeats the M-long sigmoid pass and the ordered bag chain (identical
// arithmetic on identical data: all CTAs take the same decisions without talking); per round the CTAs exchange only their
// best (objective, index) pair through distributed shared memory across ONE cluster barrier.  Same expressions and the same
// order per candidate as before; (value desc, index asc) is a total order, so the reduction tree does not matter.
#define ANY_THREADS 256
#define ANY_MAX_CLUSTER 8
template <bool ROWS_SMEM>
__device__ __forceinline__ void
anyboost_select_body(const float* __restrict__ pred, int ld, int F, int P, int Nn, int K,
                  int* __restrict__ sel, int* __restrict__ nsel)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int M = P + Nn, Mp = M | 1;
    const int FL = (F + C - 1) / C;
    const int f_begin = rank * FL, f_end = min(F, f_begin + FL), nf = max(0, f_end - f_begin);
    extern __shared__ __align__(16) unsigned char smraw[];
    float* Hs = (float*)smraw;              // [M]  hypothesis
    float* pi = Hs + M;                     // [M]  instance probabilities
    float* tq = pi + M;                     // [M]  chain terms, later instance weights
    float* obj = tq + M;                    // [FL] last objective values of this CTA's candidates
    float* rows = obj + FL;                 // [FL][Mp] their prediction rows (ROWS_SMEM)
    unsigned char* selflag = (unsigned char*)(rows + (ROWS_SMEM ? (size_t)FL * Mp : 0));   // [F]
    __shared__ SelSlot sh[ANY_THREADS / 32];
    __shared__ SelSlot slots[2][ANY_MAX_CLUSTER];
    __shared__ float s_pbag, s_nll;
    __shared__ int s_action;                // 0 continue normally, 1 special retry, 2 pop & stop
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

    for (int i = tid; i < M; i += nt) Hs[i] = 0.0f;
    for (int f = tid; f < F; f += nt) selflag[f] = 0;
    for (int f = tid; f < FL; f += nt) obj[f] = 0.0f;
    if (ROWS_SMEM)
        for (int f = warp; f < nf; f += nwarps) {
            const float* pr = pred + (size_t)(f_begin + f) * ld;
            float* dst = rows + (size_t)f * Mp;
            for (int i = lane; i < M; i += 32) dst[i] = pr[i];
        }
    __syncthreads();
    int xc = 0;                             // exchange counter (slot parity)
    // cluster-wide argmax of the CTAs' (v, f): identical result in every thread of every CTA
    auto exchange = [&](float v, int f, float* gv, int* gf) {
        const int par = xc & 1; xc++;
        if (tid < C) {
            SelSlot* remote = cluster.map_shared_rank(&slots[0][0], tid);
            remote[par * ANY_MAX_CLUSTER + rank].nll = v;
            remote[par * ANY_MAX_CLUSTER + rank].f = f;
        }
        cluster.sync();
        float bv = slots[par][0].nll; int bf = slots[par][0].f;
        for (int q = 1; q < C; q++)
            if (better_max(slots[par][q].nll, slots[par][q].f, bv, bf)) { bv = slots[par][q].nll; bf = slots[par][q].f; }
        *gv = bv; *gf = bf;
    };
    double prev = 1000000.0;
    int n_sel = 0;
    float cur_v = 0.0f; int cur_f = -1;     // position k in the last sorted order
    int have_order = 0;
    int last_sel = -1;                      // m_selectorList.back()
    // `s` mirrors selectedFeatureIndex; the retry branch re-enters the same round
    for (int s = 0; s < K; s++) {
        for (int i = tid; i < M; i += nt) {
            const float p = sigmoid_dev(Hs[i]);
            pi[i] = p;
            if (i < P) tq[i] = __fsub_rn(1.0f, p);
            else tq[i] = __fdiv_rn(-logf(__fsub_rn(__fadd_rn(1e-5f, 1.0f), p)), (float)Nn);
        }
        __syncthreads();
        if (tid == 0) {
            float like = 1.0f;
            for (int i = 0; i < P; i++) like = __fmul_rn(like, tq[i]);
            const float y = __fdiv_rn(1.0f, (float)P);
            const float pbag = __fsub_rn(1.0f, (float)pow((double)like, (double)y));   // 1 - powf(like, 1/(float)P)
            float nll = -logf(__fadd_rn(pbag, 1e-5f));
            for (int j = P; j < M; j++) nll = __fadd_rn(nll, tq[j]);
            s_pbag = pbag; s_nll = nll;
            int action = 0;
            if ((prev - (double)nll) < 1e-100) action = (n_sel <= 2) ? 1 : 2;
            s_action = action;
        }
        __syncthreads();
        const int action = s_action;
        const float pbag = s_pbag;
        if (action == 2) { n_sel--; break; }
        if (action == 1) {
            // :104-150 special retry: drop the last pick, take the next one in the previous order
            if (n_sel == 0 || !have_order) break;
            const float* pl = pred + (size_t)last_sel * ld;
            for (int i = tid; i < M; i += nt) Hs[i] = __fsub_rn(Hs[i], pl[i]);
            if (tid == 0) selflag[last_sel] = 0;
            __syncthreads();
            n_sel--;
            float bv = -CUDART_INF_F; int bf = 0x7fffffff;
            for (int fl = tid; fl < nf; fl += nt) {
                const int f = f_begin + fl;
                const float v = obj[fl];
                const bool after = (v < cur_v) || (v == cur_v && f > cur_f);
                if (after && !selflag[f] && better_max(v, f, bv, bf)) { bv = v; bf = f; }
            }
            float lv; int lf, gf; float gv;
            block_argmax(bv, bf, &lv, &lf, sh);
            exchange(lv, lf, &gv, &gf);
            if (gf == 0x7fffffff) break;                          // reference: abortError
            if (tid == 0) { if (rank == 0) sel[n_sel] = gf; selflag[gf] = 1; }
            n_sel++;
            last_sel = gf; cur_v = gv; cur_f = gf;
            const float* pb = pred + (size_t)gf * ld;
            __syncthreads();
            for (int i = tid; i < M; i += nt) Hs[i] = __fadd_rn(Hs[i], pb[i]);
            __syncthreads();
            s = n_sel - 1;                                        // selectedFeatureIndex-- ; continue -> ++
            continue;
        }
        prev = (double)s_nll;
        // instance weights (:163-173)
        const float invP = __fdiv_rn(1.0f, (float)P), invN = __fdiv_rn(1.0f, (float)Nn);
        __syncthreads();
        for (int i = tid; i < M; i += nt) {
            if (i < P) tq[i] = __fdiv_rn(__fmul_rn(__fmul_rn(invP, __fsub_rn(1.0f, pbag)), pi[i]), pbag);
            else tq[i] = __fmul_rn(-invN, pi[i]);
        }
        __syncthreads();
        // objective: ordered f32 chain per candidate (:176-189)
        float bv = -CUDART_INF_F; int bf = 0x7fffffff;
        for (int fl = tid; fl < nf; fl += nt) {
            const int f = f_begin + fl;
            float o = 0.0f;
            if (ROWS_SMEM) {
                const float* pr = rows + (size_t)fl * Mp;
#pragma unroll 8
                for (int i = 0; i < M; i++) o = __fadd_rn(o, __fmul_rn(tq[i], pr[i]));
            } else {
                const float* pr = pred + (size_t)f * ld;
                for (int i = 0; i < M; i++) o = __fadd_rn(o, __fmul_rn(tq[i], pr[i]));
            }
            obj[fl] = o;
            if (!selflag[f] && better_max(o, f, bv, bf)) { bv = o; bf = f; }
        }
        have_order = 1;
        float lv; int lf, gf; float gv;
        block_argmax(bv, bf, &lv, &lf, sh);
        exchange(lv, lf, &gv, &gf);
        if (gf == 0x7fffffff) break;                              // all selected (:203-207)
        if (tid == 0) { if (rank == 0) sel[n_sel] = gf; selflag[gf] = 1; }
        n_sel++;
        last_sel = gf; cur_v = gv; cur_f = gf;
        const float* pb = pred + (size_t)gf * ld;
        __syncthreads();
        for (int i = tid; i < M; i += nt) Hs[i] = __fadd_rn(Hs[i], pb[i]);
        __syncthreads();
    }
    if (rank == 0 && tid == 0) *nsel = n_sel < 0 ? 0 : n_sel;
    cluster.sync();          // no CTA may exit while a peer can still write its slots
}
// shared-memory plan: H, pi, tq [M] + obj [FL] + rows [FL][M | 1] (when they fit) + flags [F]
static size_t any_smem_plan(int F, int M, int C, int* rows_in_smem)
{
    const int FL = ceil_div(F, C);
    const size_t base = (size_t)(3 * M + FL) * sizeof(float) + (size_t)round_up(F, 16);
    const size_t rows = (size_t)FL * (M | 1) * sizeof(float);
    *rows_in_smem = base + rows <= 200 * 1024 ? 1 : 0;
    return base + (*rows_in_smem ? rows : 0);
}
static int any_cluster_width(int F)
{
    int C = 1;
    while (C < ANY_MAX_CLUSTER && ceil_div(F, C) > 64) C <<= 1;
    return C;
}
__global__ void __launch_bounds__(ANY_THREADS)
k_anyboost_select(const float* __restrict__ pred, int ld, int F, int P, int Nn, int K,
                  int* __restrict__ sel, int* __restrict__ nsel, int rows_in_smem)
{
    if (rows_in_smem) anyboost_select_body<true>(pred, ld, F, P, Nn, K, sel, nsel);
    else anyboost_select_body<false>(pred, ld, F, P, Nn, K, sel, nsel);
}
// batched form: one cluster per object (blockIdx.y); the plan was sized for (F_max, M_max): an object's rows fit iff its own do
__global__ void __launch_bounds__(ANY_THREADS) k_anyboost_select_batch(const TrainDesc* __restrict__ descs, int rows_words)
{
    const TrainDesc& d = descs[blockIdx.y];
    if (d.strong_type != MCT_STRONG_MILANYBOOST) return;
    if (d.P < 1 || d.Nn < 1) return;              // empty training set decided on the device (k_train_gen_default): classifier left as it is
    const int FL = (d.F + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool fit = rows_words > 0 && (size_t)FL * ((d.P + d.Nn) | 1) <= (size_t)rows_words;
    if (fit) anyboost_select_body<true>(d.pred, d.ldT, d.F, d.P, d.Nn, d.K, d.sel, d.nsel);
    else anyboost_select_body<false>(d.pred, d.ldT, d.F, d.P, d.Nn, d.K, d.sel, d.nsel);
}


// shared-memory plan of the MILBoost selection kernel: terms[FC][MIL_RS] + lists[2][FC] + H[M] + (rows[FL][M]) + flags[F]
// positives per pass (<= MIL_PC); MCT_MIL_PC overrides for experiments
static int mil_pc()
{
    static int pc = 0;
    if (!pc) { const char* e = getenv("MCT_MIL_PC"); pc = e ? atoi(e) : 32; if (pc < 1 || pc > MIL_PC) pc = 32; }
    return pc;
}
// negatives per pass: one pass for sets of up to 256 negatives as long as the term rows of a candidate chunk stay under ~64 KB
static int mil_rn(int Nn, int FL)
{
    const int FC = FL < MIL_FC ? (FL < 1 ? 1 : FL) : MIL_FC;
    int RN = MIL_RN;
    while (RN < 256 && RN < Nn && (size_t)FC * (2 * RN + MIL_PC + 4) * 4 <= 64 * 1024) RN *= 2;
    return RN;
}
static size_t mil_smem_plan(int F, int M, int FL, int* FC_out, int* rows_out, int RN = MIL_RN)
{
    int FC = FL < MIL_FC ? FL : MIL_FC;
    if (FC < 1) FC = 1;
    size_t base = (size_t)M * 4 + (size_t)(RN + MIL_PC + 4) * FC * 4 + (size_t)2 * FC * 4 + (size_t)F;
    size_t rows = (size_t)FL * M * 4;
    int in_smem = base + rows <= (size_t)200 * 1024;
    *FC_out = FC; *rows_out = in_smem;
    size_t smem = base + (in_smem ? rows : 0);
    return (smem + 15) & ~(size_t)15;
}
// ------------------------------------------------------------------------------------------
// k_adaboost_select: AdaBoostClassifier::Update selection loop (AdaBoostClassifier.cpp:62-171), online boosting with running
// TP / FP / TN / FN weights per (selector, weak classifier).  One CTA per classifier, one thread per weak classifier:
//   prologue  the weak classifiers' boolean labels for the P + Nn training samples, packed 32 per word in shared memory;
//   round s   every thread adds the sample weights to its four counters IN SAMPLE ORDER (the reference accumulates
//             straight onto the running counters in f32), forms err = (FP+FN)/(FP+FN+TP+TN); block arg-min over the not yet
//             selected classifiers (ties -> lowest index; std::sort's tie order is unspecified in the reference);
//             alpha = clamp(0.5f * logf((1-err)/(err+1e-5f)), 0, 10); sample weights *= 1/(2-2err) (correct) or 1/(2err).
#define ADA_THREADS 1024
__device__ __forceinline__ void
adaboost_select_body(const float* __restrict__ feat, int ld, int F, int P, int Nn, int K, const float* __restrict__ weak, int weak_type,
                     float* __restrict__ cnt, float* __restrict__ alpha, int* __restrict__ sel, int* __restrict__ nsel)
{
    const int M = P + Nn, W32 = (M + 31) >> 5;
    extern __shared__ __align__(16) unsigned char smraw[];
    float* lam = (float*)smraw;                          // [M] sample weights (positives then negatives)
    unsigned* bits = (unsigned*)(lam + M);               // [F][W32] labels
    unsigned char* selflag = (unsigned char*)(bits + (size_t)F * W32);
    __shared__ SelSlot sh[ADA_THREADS / 32];
    __shared__ float s_corw, s_incorw;
    __shared__ int s_best;
    const int tid = threadIdx.x, nt = blockDim.x;
    const float lp = __fdiv_rn(0.5f, (float)P), ln = __fdiv_rn(0.5f, (float)Nn);
    for (int i = tid; i < M; i += nt) lam[i] = i < P ? lp : ln;
    for (int f = tid; f < F; f += nt) selflag[f] = 0;
    // labels: one warp per weak classifier, a ballot per 32 samples
    {
        const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
        for (int f = warp; f < F; f += nw) {
            const float mu0 = weak[0 * (size_t)F + f], mu1 = weak[1 * (size_t)F + f];
            const float n0 = weak[4 * (size_t)F + f], n1 = weak[5 * (size_t)F + f], e0 = weak[6 * (size_t)F + f], e1 = weak[7 * (size_t)F + f];
            const float* x = feat + (size_t)f * ld;
            for (int w = 0; w < W32; w++) {
                const int i = w * 32 + lane;
                const bool b = i < M && weak_classify_bool_dev(weak_type, x[i], mu0, mu1, n0, n1, e0, e1);
                const unsigned bal = __ballot_sync(0xffffffffu, b);
                if (lane == 0) bits[(size_t)f * W32 + w] = bal;
            }
        }
    }
    __syncthreads();
    const size_t KF = (size_t)K * F;
    for (int s = 0; s < K; s++) {
        float bv = CUDART_INF_F; int bf = 0x7fffffff;
        for (int f = tid; f < F; f += nt) {
            float* c0 = cnt + (size_t)s * F + f;
            float tp = c0[0], fp = c0[KF], tn = c0[2 * KF], fn = c0[3 * KF];
            const unsigned* bw = bits + (size_t)f * W32;
            for (int i = 0; i < P; i++) {
                const float l = lam[i];
                if ((bw[i >> 5] >> (i & 31)) & 1u) tp = __fadd_rn(tp, l); else fp = __fadd_rn(fp, l);
            }
            for (int i = P; i < M; i++) {
                const float l = lam[i];
                if ((bw[i >> 5] >> (i & 31)) & 1u) fn = __fadd_rn(fn, l); else tn = __fadd_rn(tn, l);
            }
            c0[0] = tp; c0[KF] = fp; c0[2 * KF] = tn; c0[3 * KF] = fn;
            const float num = __fadd_rn(fp, fn);
            const float err = __fdiv_rn(num, __fadd_rn(__fadd_rn(num, tp), tn));
            if (!selflag[f] && better_min(err, f, bv, bf)) { bv = err; bf = f; }
        }
        // block arg-min
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int of = __shfl_xor_sync(0xffffffffu, bf, o);
            if (better_min(ov, of, bv, bf)) { bv = ov; bf = of; }
        }
        if ((tid & 31) == 0) { sh[tid >> 5].nll = bv; sh[tid >> 5].f = bf; }
        __syncthreads();
        if (tid == 0) {
            float v = sh[0].nll; int f = sh[0].f;
            for (int w = 1; w < (nt >> 5); w++) if (better_min(sh[w].nll, sh[w].f, v, f)) { v = sh[w].nll; f = sh[w].f; }
            s_best = f;
            if (f != 0x7fffffff) {
                sel[s] = f; selflag[f] = 1;
                const float minerr = v;
                float a = __fmul_rn(0.5f, logf(__fdiv_rn(__fsub_rn(1.0f, minerr), __fadd_rn(minerr, 0.00001f))));
                a = fmaxf(0.0f, fminf(a, 10.0f));
                alpha[s] = a;
                s_corw = __fdiv_rn(1.0f, __fsub_rn(2.0f, __fmul_rn(2.0f, minerr)));
                s_incorw = __fdiv_rn(1.0f, __fmul_rn(2.0f, minerr));
            }
        }
        __syncthreads();
        const int best = s_best;
        if (best == 0x7fffffff) { if (tid == 0) *nsel = s; return; }       // every weak classifier already selected (K > F cannot happen)
        const unsigned* bb = bits + (size_t)best * W32;
        const float corw = s_corw, incorw = s_incorw;
        for (int i = tid; i < M; i += nt) {
            const bool lab = (bb[i >> 5] >> (i & 31)) & 1u;
            const bool correct = i < P ? lab : !lab;
            lam[i] = __fmul_rn(lam[i], correct ? corw : incorw);
        }
        __syncthreads();
    }
    if (tid == 0) *nsel = K;
}
__global__ void __launch_bounds__(ADA_THREADS) k_adaboost_select_batch(const TrainDesc* __restrict__ descs)
{
    const TrainDesc& d = descs[blockIdx.x];
    if (d.strong_type != MCT_STRONG_ADABOOST) return;
    if (d.P < 1 || d.Nn < 1) return;              // empty training set decided on the device (k_train_gen_default): classifier left as it is
    adaboost_select_body(d.featT, d.ldT, d.F, d.P, d.Nn, d.K, d.weak, d.weak_type, d.ada_cnt, d.alpha, d.sel, d.nsel);
}
__global__ void __launch_bounds__(ADA_THREADS) k_adaboost_select(const float* __restrict__ feat, int ld, int F, int P, int Nn, int K,
                                                                 const float* __restrict__ weak, int weak_type, float* __restrict__ cnt,
                                                                 float* __restrict__ alpha, int* __restrict__ sel, int* __restrict__ nsel)
{
    adaboost_select_body(feat, ld, F, P, Nn, K, weak, weak_type, cnt, alpha, sel, nsel);
}
static size_t ada_smem(int F, int M) { return (size_t)M * 4 + (size_t)F * ((M + 31) / 32) * 4 + (size_t)F + 16; }

// ------------------------------------------------------------------------------------------
// k_ensemble_retain / k_ensemble_select: MILEnsembleClassifier::Update (MILEnsembleClassifier.cpp:14-157) and
// RetainBestPerformingWeakClassifiers (:206-337).  Same greedy bag-likelihood selection as MILBoost with three differences
// that matter for the bits: (1) the previous frame's selectors are re-ranked on the new samples with their OLD weak state
// and the best pct% of them are kept (f32 product chain, :270); (2) the main loop accumulates the positive bag product and
// the negative sum in DOUBLE (:88, :100) and has no IsValidWeakClassifier test (:114); (3) sort_order (Public.h:213-228)
// sorts the values in place, so `negLogLikelihoodList[order[k]]` (:131, :321) is the value of RANK order[k], not the chosen
// candidate's -- that value is what the next round's stopping test sees.  It is reproduced with an exact rank sort
// (ties -> lowest index, like the oracle; std::sort's tie order is unspecified).
// One thread-block cluster per classifier, candidates split over its CTAs; per round
//   terms   all threads: 1 - sigmoid(H + p) / -logf(1e-5f + 1 - sigmoid(H + p)) of the CTA's candidates into shared rows,
//   chains  one thread per candidate, in sample order (f32 or f64 as the reference has it) -> nll into a global row,
//   cluster barrier, every CTA reads all nll, ranks its own candidates (warp per candidate) into the sorted row and
//   takes the block arg-min over the unselected candidates, cluster barrier, every CTA applies the same decision.
// Rows written by other CTAs of the cluster during the kernel are read with ld.cg (L2), never through L1 / the nc path.
#define ENS_THREADS 1024
#define ENS_MAX_CLUSTER 8
template <bool RETAIN>
__device__ __forceinline__ void
ensemble_body(const float* __restrict__ feat, float* pred, int ld, int F, int P, int Nn, int K, const float* __restrict__ weak, int weak_type,
              float pct, int* sel, int* nsel, int* keep, float* ensH, float* gscr, int FC)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int M = P + Nn, MS = M | 1;
    extern __shared__ __align__(16) unsigned char smraw[];
    float* Hs = (float*)smraw;                       // [M]
    float* nllS = Hs + M;                            // [F] all candidates' nll of the round
    int* cand = (int*)(nllS + F);                    // [F] candidate -> weak classifier index
    float* terms = (float*)(cand + F);               // [FC][MS]
    unsigned char* flag = (unsigned char*)(terms + (size_t)FC * MS);   // [F] candidate already taken
    __shared__ SelSlot sh[ENS_THREADS / 32];
    __shared__ int s_best, s_stop;
    float* gnll = gscr; float* gsorted = gscr + F;

    const int n0 = *nsel;                            // previous selectors (RETAIN) / retained selectors (select)
    int NC, rounds_end, n_sel;
    if (RETAIN) {
        NC = n0;
        for (int q = tid; q < NC; q += nt) { cand[q] = sel[q]; flag[q] = 0; }
        for (int i = tid; i < M; i += nt) Hs[i] = 0.0f;
        n_sel = 0;
        rounds_end = (int)__fmul_rn(pct, (float)n0);                         // :214
        if (rounds_end > NC) rounds_end = NC;
    } else {
        NC = F;
        for (int q = tid; q < NC; q += nt) { cand[q] = q; flag[q] = 0; }
        for (int i = tid; i < M; i += nt) Hs[i] = ensH[i];
        __syncthreads();
        for (int q = tid; q < n0; q += nt) flag[sel[q]] = 1;
        n_sel = n0;
        rounds_end = K;
    }
    cluster.sync();                                  // everybody holds its copy of sel / nsel before CTA 0 rewrites them
    if (RETAIN) {
        if (rank == 0) {
            for (int f = tid; f < F; f += nt) keep[f] = 0;
            if (n0 == 0 || rounds_end == 0) {
                for (int i = tid; i < M; i += nt) ensH[i] = 0.0f;
                if (tid == 0) *nsel = 0;
            }
        }
        if (n0 == 0 || rounds_end == 0) return;
    }
    const int FL = (NC + C - 1) / C;
    const int q0 = min(NC, rank * FL), q1 = min(NC, q0 + FL);
    if (RETAIN) {
        // predictions of the previous selectors on the new samples, old weak state (:251-257)
        for (int idx = tid; idx < (q1 - q0) * M; idx += nt) {
            const int c = idx / M, i = idx - c * M, f = cand[q0 + c];
            const float mu0 = weak[0 * (size_t)F + f], mu1 = weak[1 * (size_t)F + f];
            const float a0 = weak[4 * (size_t)F + f], a1 = weak[5 * (size_t)F + f], e0 = weak[6 * (size_t)F + f], e1 = weak[7 * (size_t)F + f];
            pred[(size_t)f * ld + i] = weak_classify_q_dev(weak_type, feat[(size_t)f * ld + i], mu0, mu1, a0, a1, e0, e1);
        }
        __threadfence();
        cluster.sync();
    }
    double prev = 1000000.0;
    const float fP = (float)P, fN = (float)Nn;
    for (int s = n_sel; s < rounds_end; s++) {
        for (int c0 = q0; c0 < q1; c0 += FC) {
            const int nc = min(FC, q1 - c0);
            for (int idx = tid; idx < nc * M; idx += nt) {
                const int c = idx / M, i = idx - c * M;
                const float x = __fadd_rn(Hs[i], __ldcg(pred + (size_t)cand[c0 + c] * ld + i));
                const float sg = sigmoid_dev(x);
                terms[(size_t)c * MS + i] = i < P ? __fsub_rn(1.0f, sg) : -logf(__fsub_rn(__fadd_rn(1e-5f, 1.0f), sg));
            }
            __syncthreads();
            if (tid < nc) {
                const float* row = terms + (size_t)tid * MS;
                float posl, negl;
                if (RETAIN) {
                    float like = 1.0f;
                    for (int i = 0; i < P; i++) like = __fmul_rn(like, row[i]);
                    posl = (float)(-log((double)__fsub_rn(1.0f, like) + 1e-5));
                    float a = 0.0f;
                    for (int j = P; j < M; j++) a = __fadd_rn(a, row[j]);
                    negl = a;
                } else {
                    double like = 1.0;
                    for (int i = 0; i < P; i++) like = __dmul_rn(like, (double)row[i]);
                    posl = (float)(-log(__dadd_rn(__dsub_rn(1.0, like), 1e-5)));
                    double a = 0.0;
                    for (int j = P; j < M; j++) a = __dadd_rn(a, (double)row[j]);
                    negl = (float)a;
                }
                gnll[c0 + tid] = __fadd_rn(__fdiv_rn(posl, fP), __fdiv_rn(negl, fN));
            }
            __syncthreads();
        }
        __threadfence();
        cluster.sync();
        for (int q = tid; q < NC; q += nt) nllS[q] = __ldcg(gnll + q);
        __syncthreads();
        // rank of the CTA's own candidates in the stable ascending order -> sorted row
        for (int q = q0 + warp; q < q1; q += nw) {
            const float v = nllS[q];
            int cnt = 0;
            for (int u = lane; u < NC; u += 32) { const float o = nllS[u]; cnt += (o < v || (o == v && u < q)) ? 1 : 0; }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0) gsorted[cnt] = v;
        }
        // first candidate of the order that is not taken yet
        float bv = CUDART_INF_F; int bf = 0x7fffffff;
        for (int q = tid; q < NC; q += nt) { const float v = nllS[q]; if (!flag[q] && better_min(v, q, bv, bf)) { bv = v; bf = q; } }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int of = __shfl_xor_sync(0xffffffffu, bf, o);
            if (better_min(ov, of, bv, bf)) { bv = ov; bf = of; }
        }
        if (lane == 0) { sh[warp].nll = bv; sh[warp].f = bf; }
        __threadfence();
        cluster.sync();
        if (tid == 0) {
            float v = sh[0].nll; int f = sh[0].f;
            for (int w = 1; w < nw; w++) if (better_min(sh[w].nll, sh[w].f, v, f)) { v = sh[w].nll; f = sh[w].f; }
            int stop = 0;
            if (f == 0x7fffffff) stop = 1;                                       // every candidate taken (:304-307)
            else if ((prev - (double)v) < 1e-100) stop = 1;                      // :122-127 / :309-314 push then pop
            s_best = f; s_stop = stop;
        }
        __syncthreads();
        if (s_stop) break;
        const int best = s_best, bw = cand[best];
        prev = (double)__ldcg(gsorted + best);                                   // negLogLikelihoodList[order[k]] after the in-place sort
        for (int i = tid; i < M; i += nt) Hs[i] = __fadd_rn(Hs[i], __ldcg(pred + (size_t)bw * ld + i));
        if (RETAIN) {
            // "not already included" is a test on the FEATURE (count(m_selectorList, previouslySelected[order[k]]), :296): the list
            // a new classifier starts with is K copies of feature 0 (StrongClassifierBase.cpp:91), so all its slots go at once
            for (int q = tid; q < NC; q += nt) if (cand[q] == bw) flag[q] = 1;
        }
        if (tid == 0) { flag[best] = 1; if (rank == 0) { sel[n_sel] = bw; if (RETAIN) keep[bw] = 1; } }
        n_sel++;
        __syncthreads();
    }
    if (rank == 0) {
        if (RETAIN) for (int i = tid; i < M; i += nt) ensH[i] = Hs[i];
        if (tid == 0) *nsel = n_sel;
    }
}
__global__ void __launch_bounds__(ENS_THREADS) k_ensemble_retain_batch(const TrainDesc* __restrict__ descs, int FC)
{
    const TrainDesc& d = descs[blockIdx.y];
    if (d.strong_type != MCT_STRONG_MILENSEMBLE) return;
    if (d.P < 1 || d.Nn < 1) return;              // empty training set decided on the device (k_train_gen_default): classifier left as it is
    ensemble_body<true>(d.featT, d.pred, d.ldT, d.F, d.P, d.Nn, d.K, d.weak, d.weak_type, d.pct_retained, d.sel, d.nsel, d.keep, d.ensH, d.ensH + d.ldT, FC);
}
__global__ void __launch_bounds__(ENS_THREADS) k_ensemble_select_batch(const TrainDesc* __restrict__ descs, int FC)
{
    const TrainDesc& d = descs[blockIdx.y];
    if (d.strong_type != MCT_STRONG_MILENSEMBLE) return;
    if (d.P < 1 || d.Nn < 1) return;              // empty training set decided on the device (k_train_gen_default): classifier left as it is
    ensemble_body<false>(d.featT, d.pred, d.ldT, d.F, d.P, d.Nn, d.K, d.weak, d.weak_type, d.pct_retained, d.sel, d.nsel, d.keep, d.ensH, d.ensH + d.ldT, FC);
}
__global__ void __launch_bounds__(ENS_THREADS)
k_ensemble_retain(const float* __restrict__ feat, float* pred, int ld, int F, int P, int Nn, int K, const float* __restrict__ weak, int weak_type,
                  float pct, int* sel, int* nsel, int* keep, float* ensH, int FC)
{
    ensemble_body<true>(feat, pred, ld, F, P, Nn, K, weak, weak_type, pct, sel, nsel, keep, ensH, ensH + ld, FC);
}
__global__ void __launch_bounds__(ENS_THREADS)
k_ensemble_select(const float* __restrict__ feat, float* pred, int ld, int F, int P, int Nn, int K, const float* __restrict__ weak, int weak_type,
                  float pct, int* sel, int* nsel, int* keep, float* ensH, int FC)
{
    ensemble_body<false>(feat, pred, ld, F, P, Nn, K, weak, weak_type, pct, sel, nsel, keep, ensH, ensH + ld, FC);
}
// cluster width and candidate chunk of the ensemble kernels: NC candidates (K for the retain step, F for the selection)
static size_t ens_plan(int F, int M, int NC, int* C_out, int* FC_out)
{
    int C = 1;
    while (C < ENS_MAX_CLUSTER && NC / (2 * C) >= 24) C <<= 1;
    const int FL = ceil_div(NC, C), MS = M | 1;
    const size_t base = (size_t)M * 4 + (size_t)F * 8 + (size_t)F;
    int FC = (int)(((size_t)160 * 1024) / ((size_t)MS * 4));
    if (FC > FL) FC = FL;
    if (FC > ENS_THREADS) FC = ENS_THREADS;
    if (FC < 1) FC = 1;
    *C_out = C; *FC_out = FC;
    return (base + (size_t)FC * MS * 4 + 15) & ~(size_t)15;
}
template <typename KernT, typename... Args>
static int ens_launch(mct_ctx* ctx, const char* name, KernT kern, size_t* attr_state, dim3 grid, int C, size_t smem, Args... args)
{
    if (smem > 220 * 1024) return mct_fail(ctx, MCT_E_UNSUPPORTED, "training set too large for the MILEnsemble kernels%s");
    if (int rc_ = mct_smem_attr(ctx, kern, attr_state, smem, 0)) return rc_;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(ENS_THREADS, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    mct_prof_begin(ctx, name);
    MCT_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, args...));
    mct_prof_end(ctx);
    MCT_KERNEL_CHECK(ctx);
    return MCT_OK;
}
template <typename KernT, typename... Args>
static int any_launch(mct_ctx* ctx, const char* name, KernT kern, size_t* attr_state, dim3 grid, int C, size_t smem, Args... args)
{
    if (int rc_ = mct_smem_attr(ctx, kern, attr_state, smem, 0)) return rc_;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(ANY_THREADS, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    mct_prof_begin(ctx, name);
    MCT_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, args...));
    mct_prof_end(ctx);
    MCT_KERNEL_CHECK(ctx);
    return MCT_OK;
}
int launch_ensemble_retain(mct_clf* clf, int P, int Nn)
{
    if (clf->p.strong_type != MCT_STRONG_MILENSEMBLE) return MCT_OK;
    int C, FC;
    const size_t smem = ens_plan(clf->F, P + Nn, clf->K, &C, &FC);
    static size_t s_attr[MCT_MAX_DEV];
    return ens_launch(clf->ctx, "k_ensemble_retain", k_ensemble_retain, s_attr, dim3(C, 1, 1), C, smem, (const float*)clf->d_featT, clf->d_pred, clf->ldT,
                      clf->F, P, Nn, clf->K, (const float*)clf->d_weak, clf->p.weak_type, clf->p.pct_retained / 100.0f, clf->d_sel, clf->d_nsel,
                      clf->d_keep, clf->d_ensH, FC);
}
static int launch_ensemble_select(mct_clf* clf, int P, int Nn)
{
    int C, FC;
    const size_t smem = ens_plan(clf->F, P + Nn, clf->F, &C, &FC);
    static size_t s_attr[MCT_MAX_DEV];
    return ens_launch(clf->ctx, "k_ensemble_select", k_ensemble_select, s_attr, dim3(C, 1, 1), C, smem, (const float*)clf->d_featT, clf->d_pred, clf->ldT,
                      clf->F, P, Nn, clf->K, (const float*)clf->d_weak, clf->p.weak_type, clf->p.pct_retained / 100.0f, clf->d_sel, clf->d_nsel,
                      clf->d_keep, clf->d_ensH, FC);
}
// batched forms: the plan covers the largest ensemble classifier of the camera
static int ens_batch_plan(const TrainDesc* h, int n_obj, int retain, int* C_out, int* FC_out, size_t* smem_out)
{
    int any = 0, F = 0, M = 0, NC = 0;
    for (int o = 0; o < n_obj; o++) {
        if (h[o].strong_type != MCT_STRONG_MILENSEMBLE) continue;
        any = 1; F = max(F, h[o].F); M = max(M, h[o].P + h[o].Nn); NC = max(NC, retain ? h[o].K : h[o].F);
    }
    if (any) *smem_out = ens_plan(F, M, NC, C_out, FC_out);
    return any;
}
int launch_ensemble_retain_batch(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj)
{
    int C, FC; size_t smem;
    if (!ens_batch_plan(h, n_obj, 1, &C, &FC, &smem)) return MCT_OK;
    static size_t s_attr[MCT_MAX_DEV];
    return ens_launch(ctx, "k_ensemble_retain_batch", k_ensemble_retain_batch, s_attr, dim3(C, n_obj, 1), C, smem, d, FC);
}
static int launch_ensemble_select_batch(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj)
{
    int C, FC; size_t smem;
    if (!ens_batch_plan(h, n_obj, 0, &C, &FC, &smem)) return MCT_OK;
    static size_t s_attr[MCT_MAX_DEV];
    return ens_launch(ctx, "k_ensemble_select_batch", k_ensemble_select_batch, s_attr, dim3(C, n_obj, 1), C, smem, d, FC);
}

// ------------------------------------------------------------------------------------------
int launch_mil_select(mct_clf* clf, int P, int Nn)
{
    mct_ctx* ctx = clf->ctx;
    const int F = clf->F, M = P + Nn;
    if (clf->p.strong_type == MCT_STRONG_MILENSEMBLE) return launch_ensemble_select(clf, P, Nn);
    if (clf->p.strong_type == MCT_STRONG_ADABOOST) {
        const size_t smem = ada_smem(F, M);
        if (smem > 200 * 1024) return mct_fail(ctx, MCT_E_UNSUPPORTED, "training set too large for the AdaBoost kernel%s (M=%d)", "", M);
        static size_t s_attr_a[MCT_MAX_DEV];
        if (int rc_ = mct_smem_attr(ctx, k_adaboost_select, s_attr_a, smem, 32 * 1024)) return rc_;
        MCT_LAUNCH(ctx, "k_adaboost_select", k_adaboost_select<<<1, ADA_THREADS, smem, ctx->stream>>>(clf->d_featT, clf->ldT, F, P, Nn, clf->K, clf->d_weak,
                                                                                           clf->p.weak_type, clf->d_ada_cnt, clf->d_alpha, clf->d_sel, clf->d_nsel));
        return MCT_OK;
    }
    if (clf->p.strong_type == MCT_STRONG_MILANYBOOST) {
        const int C = any_cluster_width(F);
        int rows_in_smem;
        const size_t smem = any_smem_plan(F, M, C, &rows_in_smem);
        if (smem > 200 * 1024) return mct_fail(ctx, MCT_E_UNSUPPORTED, "training set too large for the AnyBoost kernel%s (M=%d)", "", M);
        static size_t s_attr[MCT_MAX_DEV];
        return any_launch(ctx, "k_anyboost_select", k_anyboost_select, s_attr, dim3(C, 1, 1), C, smem, (const float*)clf->d_pred, clf->ldT, F, P, Nn, clf->K,
                          clf->d_sel, clf->d_nsel, rows_in_smem);
    }
    // cluster width: enough CTAs that each owns <= ~48 candidates, capped at 16 (non-portable size)
    int C = clf->cluster;
    if (C <= 0) { C = 1; while (C < SEL_MAX_CLUSTER && ceil_div(F, C) > 48) C <<= 1; }
    if (C > SEL_MAX_CLUSTER) C = SEL_MAX_CLUSTER;
    const int FL = ceil_div(F, C);
    int FC, rows_in_smem;
    const int RN = mil_rn(Nn, FL);
    const size_t smem = mil_smem_plan(F, M, FL, &FC, &rows_in_smem, RN);
    if (smem > 220 * 1024) return mct_fail(ctx, MCT_E_UNSUPPORTED, "training set too large for the MIL selection kernel%s (M=%d)", "", M);
    static size_t s_attr[MCT_MAX_DEV];
    static int s_np[MCT_MAX_DEV];
    if (int rc_ = mct_smem_attr(ctx, k_milboost_select, s_attr, smem, 0)) return rc_;
    if (C > 8 && !s_np[ctx->device]) {
        MCT_CUDA(ctx, cudaFuncSetAttribute(k_milboost_select, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        s_np[ctx->device] = 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C, 1, 1);
    cfg.blockDim = dim3(SEL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const float* pred = clf->d_pred; const float* weak = clf->d_weak;
    int ld = clf->ldT, K = clf->K, wt = clf->p.weak_type;
    int* sel = clf->d_sel; int* nsel = clf->d_nsel;
    mct_prof_begin(ctx, "k_milboost_select");
    MCT_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_milboost_select, pred, ld, F, P, Nn, K, weak, wt, sel, nsel, FC, rows_in_smem, mil_pc(), RN));
    mct_prof_end(ctx);
    MCT_KERNEL_CHECK(ctx);
    return MCT_OK;
}

// ------------------------------------------------------------------------------------------
// Batched selection for all objects of a camera: one launch per strong-classifier type.
int launch_mil_select_batch(mct_ctx* ctx, const TrainDesc* d, const TrainDesc* h, int n_obj, const int* cluster_req)
{
    int any_mil = 0, any_any = 0, any_ada = 0, F_max = 0, M_max = 0, F_any = 0, M_any = 0, creq = 0;
    size_t smem_ada = 0;
    { int rc = launch_ensemble_select_batch(ctx, d, h, n_obj); if (rc) return rc; }
    for (int o = 0; o < n_obj; o++) {
        const int M = h[o].P + h[o].Nn;
        if (h[o].strong_type == MCT_STRONG_MILENSEMBLE) continue;
        if (h[o].strong_type == MCT_STRONG_ADABOOST) { any_ada = 1; smem_ada = max(smem_ada, ada_smem(h[o].F, M)); continue; }
        if (h[o].strong_type == MCT_STRONG_MILANYBOOST) { any_any = 1; F_any = max(F_any, h[o].F); M_any = max(M_any, M); }
        else { any_mil = 1; F_max = max(F_max, h[o].F); M_max = max(M_max, M); if (cluster_req && cluster_req[o] > creq) creq = cluster_req[o]; }
    }
    if (any_ada) {
        if (smem_ada > 200 * 1024) return mct_fail(ctx, MCT_E_UNSUPPORTED, "training set too large for the AdaBoost kernel%s");
        static size_t s_attr_a[MCT_MAX_DEV];
        if (int rc_ = mct_smem_attr(ctx, k_adaboost_select_batch, s_attr_a, smem_ada, 32 * 1024)) return rc_;
        MCT_LAUNCH(ctx, "k_adaboost_select_batch", k_adaboost_select_batch<<<n_obj, ADA_THREADS, smem_ada, ctx->stream>>>(d));
    }
    if (any_any) {
        const int C = any_cluster_width(F_any);
        int rows_in_smem;
        const size_t smem = any_smem_plan(F_any, M_any, C, &rows_in_smem);
        if (smem > 200 * 1024) return mct_fail(ctx, MCT_E_UNSUPPORTED, "training set too large for the AnyBoost kernel%s (M=%d)", "", M_any);
        const int rows_words = rows_in_smem ? ceil_div(F_any, C) * (M_any | 1) : 0;
        static size_t s_attr[MCT_MAX_DEV];
        if (int rc_ = any_launch(ctx, "k_anyboost_select_batch", k_anyboost_select_batch, s_attr, dim3(C, n_obj, 1), C, smem, d, rows_words)) return rc_;
    }
    if (any_mil) {
        const int F = F_max, M = M_max;
        int C = creq;
        if (C <= 0) { C = 1; while (C < SEL_MAX_CLUSTER && ceil_div(F, C) > 48) C <<= 1; }
        if (C > SEL_MAX_CLUSTER) C = SEL_MAX_CLUSTER;
        const int FL = ceil_div(F, C);
        int FC, rows_in_smem;
        int Nn_mil = 1;
        for (int o = 0; o < n_obj; o++) if (h[o].strong_type == MCT_STRONG_MILBOOST) Nn_mil = max(Nn_mil, h[o].Nn);
        const int RN = mil_rn(Nn_mil, FL);
        const size_t smem = mil_smem_plan(F, M, FL, &FC, &rows_in_smem, RN);
        if (smem > 220 * 1024) return mct_fail(ctx, MCT_E_UNSUPPORTED, "training set too large for the MIL selection kernel%s (M=%d)", "", M);
        const int FLmax = rows_in_smem ? FL : 0, Mmax = M;
        static size_t s_attr[MCT_MAX_DEV];
        static int s_np[MCT_MAX_DEV];
        if (int rc_ = mct_smem_attr(ctx, k_milboost_select_batch, s_attr, smem, 0)) return rc_;
        if (C > 8 && !s_np[ctx->device]) {
            MCT_CUDA(ctx, cudaFuncSetAttribute(k_milboost_select_batch, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            s_np[ctx->device] = 1;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C, n_obj, 1);
        cfg.blockDim = dim3(SEL_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        mct_prof_begin(ctx, "k_milboost_select_batch");
        MCT_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_milboost_select_batch, d, FC, FLmax, Mmax, mil_pc(), RN));
        mct_prof_end(ctx);
        MCT_KERNEL_CHECK(ctx);
    }
    return MCT_OK;
}
