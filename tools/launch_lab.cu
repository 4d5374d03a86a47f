// launch_lab.cu -- what a launch of 144 CTAs x 512 threads costs as a function of the dynamic shared memory it asks for
// (back-to-back launches on one stream, CUDA events around 50 of them).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k_empty(float* out, int touch)
{
    __shared__ float st[512];
    if (touch) { st[threadIdx.x] = 1.0f; __syncthreads(); if (threadIdx.x == 0 && st[5] == 7.0f) out[0] = 1.0f; }
}
int main()
{
    float* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024); printf("attr: %s\n", cudaGetErrorString(cudaGetLastError()));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int grid : {8, 144, 148, 296}) for (int kb : {0, 16, 48, 100, 160, 200, 224}) {
        if (grid == 296 && kb > 100) continue;
        for (int i = 0; i < 10; i++) k_empty<<<grid, 512, kb * 1024>>>(d, 1);
        cudaEventRecord(a);
        for (int i = 0; i < 50; i++) k_empty<<<grid, 512, kb * 1024>>>(d, 1);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaError_t le = cudaGetLastError(); if (le != cudaSuccess) printf("launch err %s\n", cudaGetErrorString(le));
        float ms = -1; cudaError_t e = cudaEventElapsedTime(&ms, a, b); if (e != cudaSuccess) printf("err %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(cudaGetLastError()));
        printf("grid %3d  dyn smem %3d KB : %.3f ms total, %.2f us per launch\n", grid, kb, ms, ms * 1000.0f / 50.0f);
    }
    // alternating carve-outs: 224 KB kernel followed by a 16 KB one
    cudaEventRecord(a);
    for (int i = 0; i < 50; i++) { k_empty<<<144, 512, 224 * 1024>>>(d, 1); k_empty<<<8, 512, 50 * 1024>>>(d, 1); }
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("alternating 224 KB x144 / 50 KB x8 : %.2f us per pair\n", ms * 1000 / 50);
    return 0;
}
