// prep_lab.cu -- stand-alone timing of k_frame_prep (not product code): includes the product kernel source directly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -DMCT_PREP_DEBUG -o tools/prep_lab tools/prep_lab.cu
#include "../multicameratracking_b200/csrc/k_integral.cu"
#include <vector>
char g_mct_err[512];
std::mutex g_mct_attr_mu;
void mct_prof_begin(mct_ctx*, const char*) {}
void mct_prof_end(mct_ctx*) {}
void mct_prof_begin_on(mct_ctx*, const char*, cudaStream_t) {}
void mct_prof_end_on(mct_ctx*, cudaStream_t) {}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
int main(int argc, char** argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 640, H = argc > 2 ? atoi(argv[2]) : 480, POOL = 96;
    mct_ctx ctx; memset(&ctx, 0, sizeof(ctx));
    CK(cudaStreamCreate(&ctx.stream));
    cudaDeviceGetAttribute(&ctx.sms, cudaDevAttrMultiProcessorCount, 0);
    ctx.W = W; ctx.H = H; ctx.pitch = W; ctx.binimg_nb = argc > 3 ? atoi(argv[3]) : 8; ctx.binimg_frame = -1;
    ctx.ii_pitch = (W + 1 + 3) & ~3;
    CK(cudaMalloc(&ctx.ii_base, ((size_t)(H + 1) * ctx.ii_pitch + 16) * 4));
    uint8_t* pool; CK(cudaMalloc(&pool, (size_t)POOL * 4 * W * H));
    std::vector<uint8_t> h((size_t)POOL * 4 * W * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 24);
    CK(cudaMemcpy(pool, h.data(), h.size(), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto setf = [&](int t) { uint8_t* b = pool + (size_t)(t % POOL) * 4 * W * H; ctx.gray = b; ctx.c0 = b + W * H; ctx.c1 = b + 2 * W * H; ctx.c2 = b + 3 * W * H; ctx.frame_id++; };
    int wb = 0;
    auto prep = [&]() { PrepTarget t = {ctx.gray, ctx.c0, ctx.c1, ctx.c2, W, H, W, ctx.ii_base, ctx.ii_pitch, ctx.binimg}; return launch_frame_prep(&ctx, t, ctx.stream, &wb); };
    if (ctx.binimg_nb) CK(cudaMalloc(&ctx.binimg, (size_t)W * H * 2));
    for (int t = 0; t < 5; t++) { setf(t); if (prep()) { printf("launch failed: %s\n", ctx.err); return 1; } }
    CK(cudaStreamSynchronize(ctx.stream));
    const int n = 200;
    CK(cudaEventRecord(e0, ctx.stream));
    for (int t = 0; t < n; t++) { setf(5 + t); prep(); }
    CK(cudaEventRecord(e1, ctx.stream)); CK(cudaStreamSynchronize(ctx.stream));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%dx%d nb=%d: %.2f us per frame_prep (back-to-back launches, %d-frame pool)\n", W, H, ctx.binimg_nb, 1e3 * ms / n, POOL);
#ifdef MCT_PREP_DEBUG
    unsigned long long hd[256 * 8];
    CK(cudaMemcpyFromSymbol(hd, g_prep_dbg, sizeof(hd)));
    unsigned long long t0 = ~0ull;
    for (int b = 0; b < 130; b++) if (hd[b * 8] && hd[b * 8] < t0) t0 = hd[b * 8];
    for (int k = 1; k < 5; k++) {
        double mx = 0, sum = 0; int arg = -1, cnt = 0;
        for (int b = 0; b < 256; b++) {
            if (!hd[b * 8] || !hd[b * 8 + k] || hd[b * 8 + k] < t0 || (k < 4 && !hd[b * 8 + 1])) continue;
            const double v = (double)(hd[b * 8 + k] - t0) * 1e-3;
            if (v > 1e6) continue;
            sum += v; cnt++;
            if (v > mx) { mx = v; arg = b; }
        }
        printf("stamp %d: mean %.2f max %.2f us (ticket %d) over %d CTAs\n", k, cnt ? sum / cnt : 0.0, mx, arg, cnt);
    }
    for (int b = 0; b < 130; b += 10) { printf("ticket %3d:", b); for (int k = 0; k < 6; k++) printf(" %7.2f", (double)(hd[b * 8 + k] - t0) * 1e-3); printf(" us\n"); }
#endif
    return 0;
}
