// peaks_lab.cu -- measured on-chip bandwidth ceilings for the gather kernels ("the builder must measure an L2
// read-bandwidth peak with its own microbenchmark"; MEASURED_PEAKS.json only has the HBM copy peak).
//   l2_read     every CTA streams a buffer that fits L2 (default 32 MiB, read 64 times inside one launch) with 16-byte loads;
//   smem_read   every thread reads 16-byte words of its CTA's shared memory, conflict-free, from a resident tile;
//   hbm_read    the same streaming kernel over a buffer larger than L2 (1 GiB), for reference against MEASURED_PEAKS.json.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks_lab tools/peaks_lab.cu
// Run  : tools/peaks_lab > gpurun_out/peaks.json   (one JSON object)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(512) k_stream_read(const uint4* __restrict__ buf, size_t n16, int passes, unsigned* sink)
{
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        // four independent loads in flight per thread
        for (; i + 3 * stride < n16; i += 4 * stride) {
            const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
            acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
        }
        for (; i < n16; i += stride) { const uint4 a = __ldcg(buf + i); acc ^= a.x ^ a.y ^ a.z ^ a.w; }
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(1024) k_smem_read(int iters, unsigned* sink)
{
    extern __shared__ uint4 tile[];
    const int n16 = 64 * 1024 / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) tile[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    unsigned acc = 0;
    int j = threadIdx.x;
    const unsigned base = (unsigned)__cvta_generic_to_shared(tile);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            unsigned x, y, z, w;
            // ld.volatile: ptxas hoists a plain ld.shared of unchanged shared memory out of the loop
            asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(base + (unsigned)j * 16u));
            acc += x ^ y ^ z ^ w;
            j = (j + 1024) & (n16 - 1);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char** argv)
{
    size_t l2_mib = argc > 1 ? (size_t)atoi(argv[1]) : 32;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    unsigned* sink; CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    uint4* buf; const size_t big = (size_t)1 << 30;
    CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 1, big));

    // ---- L2 read
    double l2_best = 0;
    {
        const size_t bytes = l2_mib << 20, n16 = bytes / 16;
        const int passes = 64, grid = sms * 4;
        k_stream_read<<<grid, 512>>>(buf, n16, 4, sink);            // warm L2
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaEventRecord(e0));
            k_stream_read<<<grid, 512>>>(buf, n16, passes, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            const double gbs = (double)bytes * passes / (time_ms(e0, e1) * 1e-3) / 1e9;
            if (gbs > l2_best) l2_best = gbs;
        }
    }
    // ---- HBM read (buffer >> L2)
    double hbm_best = 0;
    {
        const size_t n16 = big / 16;
        const int grid = sms * 8;
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaEventRecord(e0));
            k_stream_read<<<grid, 512>>>(buf, n16, 1, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            const double gbs = (double)big / (time_ms(e0, e1) * 1e-3) / 1e9;
            if (gbs > hbm_best) hbm_best = gbs;
        }
    }
    // ---- shared-memory read
    double sm_best = 0;
    {
        const int iters = 4096, grid = sms * 2;
        CK(cudaFuncSetAttribute(k_smem_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        k_smem_read<<<grid, 1024, 64 * 1024>>>(16, sink);
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaEventRecord(e0));
            k_smem_read<<<grid, 1024, 64 * 1024>>>(iters, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            const double bytes = (double)grid * 1024 * iters * 4 * 16;
            const double gbs = bytes / (time_ms(e0, e1) * 1e-3) / 1e9;
            if (gbs > sm_best) sm_best = gbs;
        }
    }
    CK(cudaDeviceSynchronize());
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_read_gbs\": %.1f, \"l2_buffer_mib\": %zu, \"hbm_read_gbs\": %.1f, \"smem_read_gbs\": %.1f, "
           "\"how\": \"tools/peaks_lab.cu: 16-byte loads, best of 5, CUDA events; l2 = 64 passes over an L2-resident buffer in one launch; "
           "smem = conflict-free 16-byte reads of a 64 KiB tile, 2 CTAs of 1024 threads per SM\"}\n",
           prop.name, sms, l2_best, l2_mib, hbm_best, sm_best);
    return 0;
}
