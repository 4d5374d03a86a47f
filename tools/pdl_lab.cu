// pdl_lab.cu -- what programmatic dependent launch takes off a launch boundary: a chain of dependent kernels (each reads what the
// previous one wrote) launched back to back, plain against cudaLaunchAttributeProgrammaticStreamSerialization with the trigger at
// the top of the producer and griddepcontrol.wait in front of the consumer's first dependent load.
#include <cstdio>
#include <cuda_runtime.h>
template <bool PDL>
__global__ void __launch_bounds__(512, 1) k_link(const float* in, float* out, int work)
{
    if (PDL) asm volatile("griddepcontrol.launch_dependents;");
    __shared__ float st[512];
    st[threadIdx.x] = (float)threadIdx.x;                  // prologue that does not depend on the producer
    __syncthreads();
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    float v = in[blockIdx.x * 512 + threadIdx.x] + st[(threadIdx.x + 1) & 511];
    for (int i = 0; i < work; i++) v = v * 1.0001f + 0.5f;
    out[blockIdx.x * 512 + threadIdx.x] = v;
}
template <bool PDL>
static void launch(int grid, int smem_kb, const float* in, float* out, int work, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = (size_t)smem_kb * 1024; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = PDL ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_link<PDL>, in, out, work);
}
int main()
{
    float *a, *b; cudaMalloc(&a, 296 * 512 * 4); cudaMalloc(&b, 296 * 512 * 4); cudaMemset(a, 0, 296 * 512 * 4); cudaMemset(b, 0, 296 * 512 * 4);
    cudaFuncSetAttribute(k_link<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_link<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int work : {0, 2000, 20000}) for (int grid : {8, 144}) for (int kb : {0, 200}) {
        float ms[2];
        for (int pdl = 0; pdl < 2; pdl++) {
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0, st);
                for (int i = 0; i < 60; i++) {
                    if (pdl) launch<true>(grid, kb, (i & 1) ? b : a, (i & 1) ? a : b, work, st);
                    else launch<false>(grid, kb, (i & 1) ? b : a, (i & 1) ? a : b, work, st);
                }
                cudaEventRecord(e1, st); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms[pdl], e0, e1);
            }
        }
        printf("work %5d grid %3d smem %3d KB : plain %.2f us per kernel, PDL %.2f us  (%s)\n", work, grid, kb, ms[0] * 1000 / 60, ms[1] * 1000 / 60,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
