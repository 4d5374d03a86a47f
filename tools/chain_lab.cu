// chain_lab.cu -- cycles per element of the ordered f32 chains of k_pf_update (one warp, the other warps parked at a barrier).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/chain_lab tools/chain_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define PIDX(e) ((e) + ((e) >> 6))

template <int CH> __device__ __forceinline__ float sum_sel(const float* w, int n, int lane)
{
    float carry = 0.0f;
    for (int t0 = 0; t0 < n; t0 += 32 * CH) {
        float r[CH];
        const int e0 = t0 + lane * CH;
#pragma unroll
        for (int k = 0; k < CH; k++) r[k] = (e0 + k < n) ? w[PIDX(e0 + k)] : 0.0f;
        for (int L = 0; L < 32; L++) {
            const bool mine = lane == L;
            float c = carry;
#pragma unroll
            for (int k = 0; k < CH; k++) c = __fadd_rn(c, mine ? r[k] : 0.0f);
            carry = __shfl_sync(0xffffffffu, c, L);
        }
    }
    return carry;
}
template <int CH> __device__ __forceinline__ float sum_br(const float* w, int n, int lane)
{
    float carry = 0.0f;
    for (int t0 = 0; t0 < n; t0 += 32 * CH) {
        float r[CH];
        const int e0 = t0 + lane * CH;
#pragma unroll
        for (int k = 0; k < CH; k++) r[k] = (e0 + k < n) ? w[PIDX(e0 + k)] : 0.0f;
        for (int L = 0; L < 32; L++) {
            float c = carry;
            if (lane == L) {
#pragma unroll
                for (int k = 0; k < CH; k++) c = __fadd_rn(c, r[k]);
            }
            carry = __shfl_sync(0xffffffffu, c, L);
        }
    }
    return carry;
}
template <int CH> __device__ __forceinline__ float cdf_sel(const float* w, float* cdf, int n, int lane)
{
    float carry = 0.0f;
    for (int t0 = 0; t0 < n; t0 += 32 * CH) {
        float r[CH];
        const int e0 = t0 + lane * CH;
#pragma unroll
        for (int k = 0; k < CH; k++) r[k] = (e0 + k < n) ? w[PIDX(e0 + k)] : 0.0f;
        for (int L = 0; L < 32; L++) {
            const bool mine = lane == L;
            float c = carry;
#pragma unroll
            for (int k = 0; k < CH; k++) { c = __fadd_rn(c, mine ? r[k] : 0.0f); r[k] = mine ? c : r[k]; }
            carry = __shfl_sync(0xffffffffu, c, L);
        }
#pragma unroll
        for (int k = 0; k < CH; k++) if (e0 + k < n) cdf[PIDX(e0 + k)] = r[k];
    }
    return carry;
}
// in-place prefix inside a divergent branch: r[k] += r[k-1]
template <int CH> __device__ __forceinline__ float cdf_br(const float* w, float* cdf, int n, int lane)
{
    float carry = 0.0f;
    for (int t0 = 0; t0 < n; t0 += 32 * CH) {
        float r[CH];
        const int e0 = t0 + lane * CH;
#pragma unroll
        for (int k = 0; k < CH; k++) r[k] = (e0 + k < n) ? w[PIDX(e0 + k)] : 0.0f;
        for (int L = 0; L < 32; L++) {
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
        for (int k = 0; k < CH; k++) if (e0 + k < n) cdf[PIDX(e0 + k)] = r[k];
    }
    return carry;
}
// one thread, operands from shared memory, software-pipelined by the compiler (unroll 16)
__device__ __forceinline__ float sum_one(const float* w, int n)
{
    float c = 0.0f;
#pragma unroll 16
    for (int i = 0; i < n; i++) c = __fadd_rn(c, w[PIDX(i)]);
    return c;
}

template <int V, int CH>
__global__ void __launch_bounds__(512, 1) k_lab(const float* __restrict__ in, int n, float* out, long long* cyc, float* cdf_out)
{
    __shared__ float sw[2048 + 64], cdf[2048 + 64];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sw[PIDX(i)] = in[i];
    __syncthreads();
    float r = 0.0f;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        t0 = clock64();
        if (V == 0) r = sum_sel<CH>(sw, n, lane);
        if (V == 1) r = sum_br<CH>(sw, n, lane);
        if (V == 2) r = cdf_sel<CH>(sw, cdf, n, lane);
        if (V == 3) r = cdf_br<CH>(sw, cdf, n, lane);
        if (V == 4) { if (lane == 0) r = sum_one(sw, n); r = __shfl_sync(0xffffffffu, r, 0); }
        t1 = clock64();
    }
    __syncthreads();
    if (threadIdx.x == 0) { out[blockIdx.x] = r; cyc[blockIdx.x] = t1 - t0; }
    if (V >= 2 && V <= 3) for (int i = threadIdx.x; i < n; i += blockDim.x) cdf_out[i] = cdf[PIDX(i)];
}

template <int V, int CH> void run(const char* name, const float* d_in, int n, float ref, const float* ref_cdf)
{
    float* d_out; long long* d_cyc; float* d_cdf;
    cudaMalloc(&d_out, 64); cudaMalloc(&d_cyc, 64); cudaMalloc(&d_cdf, 4 * 2048);
    long long best = 1ll << 60; float r = 0;
    for (int it = 0; it < 5; it++) {
        k_lab<V, CH><<<1, 512>>>(d_in, n, d_out, d_cyc, d_cdf);
        long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&r, d_out, 4, cudaMemcpyDeviceToHost);
        if (c < best) best = c;
    }
    int bad = 0;
    if (V >= 2 && V <= 3) { float h[2048]; cudaMemcpy(h, d_cdf, 4 * n, cudaMemcpyDeviceToHost); for (int i = 0; i < n; i++) bad += h[i] != ref_cdf[i]; }
    printf("%-10s CH=%2d n=%d  %7lld cycles  %.2f cyc/elt  result %s (cdf mismatches %d)\n", name, CH, n, best, (double)best / n, r == ref ? "exact" : "DIFFERENT", bad);
    cudaFree(d_out); cudaFree(d_cyc); cudaFree(d_cdf);
}

int main()
{
    for (int n : {1000, 2000}) {
        float h[2048], cdf[2048]; srand(7);
        for (int i = 0; i < n; i++) h[i] = (float)rand() / RAND_MAX / n;
        float ref = 0; for (int i = 0; i < n; i++) { ref += h[i]; cdf[i] = ref; }
        float* d_in; cudaMalloc(&d_in, 4 * 2048); cudaMemcpy(d_in, h, 4 * n, cudaMemcpyHostToDevice);
        run<0, 64>("sum_sel", d_in, n, ref, cdf); run<0, 32>("sum_sel", d_in, n, ref, cdf); run<0, 16>("sum_sel", d_in, n, ref, cdf);
        run<1, 64>("sum_br", d_in, n, ref, cdf); run<1, 32>("sum_br", d_in, n, ref, cdf); run<1, 16>("sum_br", d_in, n, ref, cdf);
        run<2, 64>("cdf_sel", d_in, n, ref, cdf); run<2, 32>("cdf_sel", d_in, n, ref, cdf);
        run<3, 64>("cdf_br", d_in, n, ref, cdf); run<3, 32>("cdf_br", d_in, n, ref, cdf); run<3, 16>("cdf_br", d_in, n, ref, cdf);
        run<4, 1>("sum_one", d_in, n, ref, cdf);
        cudaFree(d_in);
    }
    return 0;
}
