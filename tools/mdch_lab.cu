// mdch_lab.cu -- design-space microbenchmark for the MDCH selected-bin histogram kernel (not product code).
// Compares per-box accumulation strategies on cfg2-shaped data: 8 objects x 2000 boxes of 40x80 px,
// sigma = 20 px scatter, 640x480 joint-bin image, ~100 selected bins.  All variants must agree bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mdch_lab tools/mdch_lab.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <string.h>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Args {
    const uint16_t* binimg; int W, H;
    const uint16_t* slotmap;   // [n_obj][512]
    const int* col; const int* row;   // [n_obj][N]
    int N, rows, cols, n_slots;
    uint32_t* out;             // [n_obj][n_slots][N]
};

__device__ __forceinline__ void build_lut(uint32_t* lut, int rows, int cols)
{
    const int np = rows * cols;
    int shift = __clz(np); if (shift > 20) shift = 20;
    const float scale = (float)(1u << shift);
    const float hx = (float)cols / 2.0f, hy = (float)rows / 2.0f;
    const double ivx = 1.0 / (double)(2.0f * hx * hx), ivy = 1.0 / (double)(2.0f * hy * hy);
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        const int r = p / cols, c = p - r * cols;
        const double dx = (double)(hx - (float)c), dy = (double)(hy - (float)r);
        const float wd = (float)(dx * dx * ivx + dy * dy * ivy);
        lut[p] = __float2uint_rn(expf(-wd) * scale);
    }
}

// ---------------------------------------------------------------- V0: thread per box, private smem hist [slot][lane]
__global__ void __launch_bounds__(64) v0_thread_per_box(Args a)
{
    extern __shared__ uint32_t sm[];
    uint32_t* lut = sm; uint16_t* smap = (uint16_t*)(lut + a.rows * a.cols);
    uint32_t* hist_all = (uint32_t*)(smap + 512);
    const int o = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    build_lut(lut, a.rows, a.cols);
    for (int b = tid; b < 512; b += blockDim.x) smap[b] = a.slotmap[o * 512 + b];
    uint32_t* hist = hist_all + warp * a.n_slots * 32 + lane;
    for (int s = 0; s < a.n_slots; s++) hist[s * 32] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + tid;
    if (i >= a.N) return;
    const int c0 = a.col[o * a.N + i], r0 = a.row[o * a.N + i];
    const uint16_t* img = a.binimg + (size_t)r0 * a.W + c0;
    const uint32_t* lr = lut;
    for (int r = 0; r < a.rows; r++) {
        for (int c = 0; c < a.cols; c++) hist[(unsigned)smap[img[c]] * 32] += lr[c];
        img += a.W; lr += a.cols;
    }
    for (int s = 0; s < a.n_slots; s++) a.out[((size_t)o * a.n_slots + s) * a.N + i] = hist[s * 32];
}

// ---------------------------------------------------------------- warp per box family
// MODE 0: all-lane smem atomics; 1: match_any + redux + leader atomic; 2: match_any + redux + leader plain RMW
//      3: leader loop (ballot/shfl) + redux
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) v_warp_per_box(Args a)
{
    extern __shared__ uint32_t sm[];
    uint32_t* lut = sm; uint16_t* smap = (uint16_t*)(lut + a.rows * a.cols);
    const int nsp = (a.n_slots + 31) & ~31;
    uint32_t* hist_all = (uint32_t*)(smap + 512);
    const int o = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    build_lut(lut, a.rows, a.cols);
    for (int b = tid; b < 512; b += blockDim.x) smap[b] = a.slotmap[o * 512 + b];
    __syncthreads();
    uint32_t* hist = hist_all + warp * nsp;
    const int npix = a.rows * a.cols, cols = a.cols;
    for (int i = blockIdx.x * WARPS + warp; i < a.N; i += gridDim.x * WARPS) {
        for (int s = lane; s < nsp; s += 32) hist[s] = 0;
        __syncwarp();
        const int c0 = a.col[o * a.N + i], r0 = a.row[o * a.N + i];
        const uint16_t* img = a.binimg + (size_t)r0 * a.W + c0;
        int c = lane, roff = 0;
        while (c >= cols) { c -= cols; roff += a.W; }
        for (int p0 = 0; p0 < npix; p0 += 32) {
            const int p = p0 + lane;
            const bool act = p < npix;
            unsigned s = 0; uint32_t w = 0;
            if (act) { s = smap[img[roff + c]]; w = lut[p]; }
            if (MODE == 0) {
                if (act) atomicAdd(&hist[s], w);
            } else if (MODE == 1 || MODE == 2) {
                const unsigned am = __ballot_sync(0xffffffffu, act);
                if (act) {
                    const unsigned m = __match_any_sync(am, s);
                    const uint32_t sum = __reduce_add_sync(m, w);
                    if ((m & ((1u << lane) - 1)) == 0) {
                        if (MODE == 1) atomicAdd(&hist[s], sum); else hist[s] += sum;
                    }
                }
            } else {
                unsigned todo = __ballot_sync(0xffffffffu, act);
                while (todo) {
                    const int leader = __ffs(todo) - 1;
                    const unsigned ls = __shfl_sync(0xffffffffu, s, leader);
                    const bool mine = act && (s == ls);
                    const uint32_t sum = __reduce_add_sync(0xffffffffu, mine ? w : 0u);
                    if (lane == leader) hist[ls] += sum;
                    todo &= ~__ballot_sync(0xffffffffu, mine);
                }
            }
            c += 32;
            while (c >= cols) { c -= cols; roff += a.W; }
        }
        __syncwarp();
        for (int s = lane; s < a.n_slots; s += 32) a.out[((size_t)o * a.n_slots + s) * a.N + i] = hist[s];
        __syncwarp();
    }
}

// ---------------------------------------------------------------- V5: lanes own rows-strips with T lanes per box, private hists
// T lanes per box (T = 8): lane j of the box handles rows j, j+T, ...; private hist [slot][T lanes] -> layout [slot][32] per warp
template <int T>
__global__ void __launch_bounds__(64) v5_sub_warp(Args a)
{
    extern __shared__ uint32_t sm[];
    uint32_t* lut = sm; uint16_t* smap = (uint16_t*)(lut + a.rows * a.cols);
    uint32_t* hist_all = (uint32_t*)(smap + 512);
    const int o = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    build_lut(lut, a.rows, a.cols);
    for (int b = tid; b < 512; b += blockDim.x) smap[b] = a.slotmap[o * 512 + b];
    uint32_t* hist = hist_all + warp * a.n_slots * 32 + lane;
    for (int s = 0; s < a.n_slots; s++) hist[s * 32] = 0;
    __syncthreads();
    const int BPW = 32 / T;                                  // boxes per warp
    const int i = (blockIdx.x * 2 + warp) * BPW + lane / T;
    const int j = lane % T;
    if (i < a.N) {
        const int c0 = a.col[o * a.N + i], r0 = a.row[o * a.N + i];
        for (int r = j; r < a.rows; r += T) {
            const uint16_t* img = a.binimg + (size_t)(r0 + r) * a.W + c0;
            const uint32_t* lr = lut + r * a.cols;
            for (int c = 0; c < a.cols; c++) hist[(unsigned)smap[img[c]] * 32] += lr[c];
        }
    }
    __syncwarp();
    // reduce the T partials of each box: lanes of the box stride over slots
    if (i < a.N) {
        const uint32_t* hb = hist_all + warp * a.n_slots * 32 + (lane / T) * T;
        for (int s = j; s < a.n_slots; s += T) {
            uint32_t t = 0;
            for (int q = 0; q < T; q++) t += hb[s * 32 + q];
            a.out[((size_t)o * a.n_slots + s) * a.N + i] = t;
        }
    }
}


// ---------------------------------------------------------------- V6: T lanes per box, 4 columns per lane, private hists,
// u8 slot region of the object staged in shared memory (slot map applied while staging), uint4 LUT rows.
struct BBox { int x0, y0, x1, y1; };
template <int T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) v6_strips(Args a, const BBox* __restrict__ bbox, const uint32_t* __restrict__ glut)
{
    extern __shared__ uint32_t sm[];
    const int o = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BBox bb = bbox[o];
    const int RW = bb.x1 - bb.x0, RH = bb.y1 - bb.y0;
    const int RWp = ((RW + 3) & ~3) + 4;                      // bytes per region row (word aligned, +1 word slack)
    uint32_t* lut = sm;                                       // [rows*cols]
    uint16_t* smap = (uint16_t*)(lut + a.rows * a.cols);      // [512]
    uint32_t* hist_all = (uint32_t*)(smap + 512);             // [WARPS][n_slots][32]
    uint8_t* region = (uint8_t*)(hist_all + (size_t)WARPS * a.n_slots * 32);
    for (int b = tid; b < 512; b += blockDim.x) smap[b] = a.slotmap[o * 512 + b];
    for (int p = tid; p < a.rows * a.cols; p += blockDim.x) lut[p] = glut[p];
    uint32_t* hist = hist_all + (size_t)warp * a.n_slots * 32 + lane;
    for (int s = 0; s < a.n_slots; s++) hist[s * 32] = 0;
    __syncthreads();
    for (int y = warp; y < RH; y += WARPS) {
        const uint16_t* src = a.binimg + (size_t)(bb.y0 + y) * a.W + bb.x0;
        for (int x = lane; x < RWp; x += 32) region[y * RWp + x] = x < RW ? (uint8_t)smap[src[x]] : 0;
    }
    __syncthreads();
    const int BPW = 32 / T;
    const int bw = lane / T, j = lane - bw * T;
    const int i = (blockIdx.x * WARPS + warp) * BPW + bw;
    const bool ok = bw < BPW && i < a.N;
    if (ok) {
        const int c0 = a.col[o * a.N + i] - bb.x0, r0 = a.row[o * a.N + i] - bb.y0;
        int addr = r0 * RWp + c0 + 4 * j;
        const uint4* l4 = reinterpret_cast<const uint4*>(lut) + j;
        const int q = a.cols >> 2;
        #pragma unroll 2
        for (int r = 0; r < a.rows; r++) {
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(region + (addr & ~3));
            const uint32_t v = __funnelshift_r(wp[0], wp[1], (addr & 3) * 8);
            const uint4 w = l4[r * q];
            hist[(v & 0xffu) * 32] += w.x;
            hist[((v >> 8) & 0xffu) * 32] += w.y;
            hist[((v >> 16) & 0xffu) * 32] += w.z;
            hist[(v >> 24) * 32] += w.w;
            addr += RWp;
        }
    }
    __syncwarp();
    if (ok) {
        const uint32_t* hb = hist_all + (size_t)warp * a.n_slots * 32 + bw * T;
        for (int s = j; s < a.n_slots; s += T) {
            uint32_t t = 0;
            #pragma unroll
            for (int qq = 0; qq < T; qq++) t += hb[s * 32 + qq];
            a.out[((size_t)o * a.n_slots + s) * a.N + i] = t;
        }
    }
}

// ---------------------------------------------------------------- V7: persistent CTAs per object chunk; T = cols/4 lanes per box,
// private hists, 4-way load-then-forward RMW batches (one LDS latency per 4 pixels), region bbox computed per CTA.
__device__ unsigned long long g_dbg[64];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define STAMP(k) do { if (blockIdx.x == 3 && blockIdx.y == 2 && threadIdx.x == 0) g_dbg[k] = gtime(); } while (0)
template <int WARPS, int FWD>
__global__ void __launch_bounds__(WARPS * 32, 1) v7_forward(Args a, int ctas_per_obj, int region_cap)
{
    extern __shared__ uint32_t sm[];
    __shared__ int s_bb[4];
    const int o = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows = a.rows, cols = a.cols, colsp = (cols + 3) & ~3, T = colsp >> 2, BPW = 32 / T;
    const int chunk = (a.N + ctas_per_obj - 1) / ctas_per_obj;
    const int i_begin = blockIdx.x * chunk, i_end = min(a.N, i_begin + chunk);
    uint32_t* lut = sm;                                       // [rows][colsp]
    uint16_t* smap = (uint16_t*)(lut + (rows + 1) * colsp);   // [512]
    uint32_t* hist_all = (uint32_t*)(smap + 512);             // [WARPS][n_slots][32]
    uint8_t* region = (uint8_t*)(hist_all + (size_t)WARPS * a.n_slots * 32);
    STAMP(0);
    if (tid == 0) { s_bb[0] = 1 << 30; s_bb[1] = 1 << 30; s_bb[2] = -(1 << 30); s_bb[3] = -(1 << 30); }
    for (int b = tid; b < 512; b += blockDim.x) smap[b] = a.slotmap[o * 512 + b];
    __syncthreads();
    {
        int x0 = 1 << 30, y0 = 1 << 30, x1 = -(1 << 30), y1 = -(1 << 30);
        for (int i = i_begin + tid; i < i_end; i += blockDim.x) {
            const int c = a.col[o * a.N + i], r = a.row[o * a.N + i];
            x0 = min(x0, c); y0 = min(y0, r); x1 = max(x1, c); y1 = max(y1, r);
        }
        for (int d = 16; d > 0; d >>= 1) {
            x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, d)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, d));
            x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, d)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, d));
        }
        if (lane == 0) { atomicMin(&s_bb[0], x0); atomicMin(&s_bb[1], y0); atomicMax(&s_bb[2], x1); atomicMax(&s_bb[3], y1); }
    }
    STAMP(1);
    {   // LUT, zero padded to colsp
        const int np = rows * cols;
        int shift = __clz(np); if (shift > 20) shift = 20;
        const float scale = (float)(1u << shift);
        const float hx = (float)cols / 2.0f, hy = (float)rows / 2.0f;
        const double ivx = 1.0 / (double)(2.0f * hx * hx), ivy = 1.0 / (double)(2.0f * hy * hy);
        for (int p = tid; p < (rows + 1) * colsp; p += blockDim.x) {
            const int r = p / colsp, c = p - r * colsp;
            const double dx = (double)(hx - (float)c), dy = (double)(hy - (float)r);
            const float wd = (float)(dx * dx * ivx + dy * dy * ivy);
            lut[p] = (c < cols && r < rows) ? __float2uint_rn(expf(-wd) * scale) : 0u;
        }
    }
    __syncthreads();
    STAMP(2);
    const int bx0 = s_bb[0] & ~3, by0 = s_bb[1];
    const int RW = s_bb[2] + colsp - bx0, RH = s_bb[3] + rows - by0 + 1;   // +1 row: the prefetch of row `rows` stays inside
    const int RWp = ((RW + 3) & ~3) + 4;
    if (RWp * RH > region_cap) { if (tid == 0 && blockIdx.x == 0) printf("region too large %d x %d\n", RW, RH); return; }
    {
        const int qw = RWp >> 2;                               // 4-pixel groups per region row (W % 4 == 0 assumed)
        uint32_t* r32 = reinterpret_cast<uint32_t*>(region);
#pragma unroll 4
        for (int g = tid; g < qw * RH; g += blockDim.x) {
            const int y = g / qw, xg = g - y * qw;
            const int yy = min(by0 + y, a.H - 1), xx = min(bx0 + 4 * xg, a.W - 4);
            const uint2 px = *reinterpret_cast<const uint2*>(a.binimg + (size_t)yy * a.W + xx);
            r32[g] = (uint32_t)smap[px.x & 0xffffu] | ((uint32_t)smap[px.x >> 16] << 8) |
                     ((uint32_t)smap[px.y & 0xffffu] << 16) | ((uint32_t)smap[px.y >> 16] << 24);
        }
    }
    __syncthreads();
    STAMP(3);
    int rnd = 0;
    const int bw = lane / T, j = lane - bw * T;
    uint32_t* hist = hist_all + (size_t)warp * a.n_slots * 32 + lane;
    const int q = colsp >> 2;
    for (int base = i_begin + warp * BPW; base < i_end; base += WARPS * BPW) {
        STAMP(4 + 3 * rnd);
        for (int s = 0; s < a.n_slots; s++) hist[s * 32] = 0;
        __syncwarp();
        const int i = base + bw;
        const bool ok = bw < BPW && i < i_end;
        if (ok) {
            const int c0 = a.col[o * a.N + i] - bx0, r0 = a.row[o * a.N + i] - by0;
            int addr = r0 * RWp + c0 + 4 * j;
            const uint4* l4 = reinterpret_cast<const uint4*>(lut) + j;
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(region + (addr & ~3));
            const int sh = (addr & 3) * 8;
            uint32_t v = __funnelshift_r(wp[0], wp[1], sh);
            uint4 w = l4[0];
            const int rstep = RWp >> 2;
            for (int r = 0; r < rows; r++) {
                // prefetch the next row (row `rows` reads one row past the box: still inside the padded region or the hist area)
                wp += rstep;
                const uint32_t vn = __funnelshift_r(wp[0], wp[1], sh);
                const uint4 wn = l4[(r + 1) * q];
                const uint32_t s0 = v & 0xffu, s1 = (v >> 8) & 0xffu, s2 = (v >> 16) & 0xffu, s3 = v >> 24;
                if (FWD == 4) {
                    const uint32_t h0 = hist[s0 * 32], h1 = hist[s1 * 32], h2 = hist[s2 * 32], h3 = hist[s3 * 32];
                    const uint32_t n0 = h0 + w.x;
                    const uint32_t n1 = (s1 == s0 ? n0 : h1) + w.y;
                    uint32_t t = (s2 == s0 ? n0 : h2); t = (s2 == s1 ? n1 : t);
                    const uint32_t n2 = t + w.z;
                    t = (s3 == s0 ? n0 : h3); t = (s3 == s1 ? n1 : t); t = (s3 == s2 ? n2 : t);
                    const uint32_t n3 = t + w.w;
                    hist[s0 * 32] = n0; hist[s1 * 32] = n1; hist[s2 * 32] = n2; hist[s3 * 32] = n3;
                } else if (FWD == 2) {
                    {
                        const uint32_t h0 = hist[s0 * 32], h1 = hist[s1 * 32];
                        const uint32_t n0 = h0 + w.x;
                        const uint32_t n1 = (s1 == s0 ? n0 : h1) + w.y;
                        hist[s0 * 32] = n0; hist[s1 * 32] = n1;
                    }
                    {
                        const uint32_t h2 = hist[s2 * 32], h3 = hist[s3 * 32];
                        const uint32_t n2 = h2 + w.z;
                        const uint32_t n3 = (s3 == s2 ? n2 : h3) + w.w;
                        hist[s2 * 32] = n2; hist[s3 * 32] = n3;
                    }
                } else {
                    hist[s0 * 32] += w.x; hist[s1 * 32] += w.y; hist[s2 * 32] += w.z; hist[s3 * 32] += w.w;
                }
                v = vn; w = wn;
            }
        }
        __syncwarp();
        STAMP(5 + 3 * rnd);
        if (ok) {
            const uint32_t* hb = hist_all + (size_t)warp * a.n_slots * 32 + bw * T;
            for (int s = j; s < a.n_slots; s += 2 * T) {
                const int s2 = s + T;
                const bool has2 = s2 < a.n_slots;
                const uint32_t* r0p = hb + s * 32;
                const uint32_t* r1p = hb + (has2 ? s2 : s) * 32;
                uint32_t t0 = 0, t1 = 0;
                int qq = j;                                   // skewed start: lanes of a box hit distinct banks
#pragma unroll 2
                for (int k = 0; k < T; k++) { t0 += r0p[qq]; t1 += r1p[qq]; qq = (qq + 1 == T) ? 0 : qq + 1; }
                a.out[((size_t)o * a.n_slots + s) * a.N + i] = t0;
                if (has2) a.out[((size_t)o * a.n_slots + s2) * a.N + i] = t1;
            }
        }
        __syncwarp();
        STAMP(6 + 3 * rnd); rnd++;
    }
    STAMP(20);
}

// ---------------------------------------------------------------- microbench: match_any / redux throughput
__global__ void mb_match(const unsigned* in, unsigned* out, int iters, int mode)
{
    unsigned v = in[threadIdx.x & 31] , acc = 0;
    for (int it = 0; it < iters; it++) {
        if (mode == 0) { unsigned m = __match_any_sync(0xffffffffu, v + (acc & 1)); acc += m; }
        else if (mode == 1) { acc += __reduce_add_sync(0xffffffffu, v + acc); }
        else { unsigned m = __match_any_sync(0xffffffffu, v + (acc & 1)); acc += __reduce_add_sync(m, v); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv)
{
    const int W = 640, H = 480, NOBJ = 8, N = 2000, rows = 80, cols = 40;
    int n_sel = argc > 1 ? atoi(argv[1]) : 100;
    float noise_amp = argc > 2 ? atof(argv[2]) : 8.0f;
    const int n_slots = n_sel + 1;
    srand(5);
    std::vector<uint16_t> bin(W * H), smap(NOBJ * 512, 0);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        int ch[3];
        for (int k = 0; k < 3; k++) {
            float v = 128 + 90 * sinf(0.013f * x * (k + 1) + 0.7f * k) * cosf(0.017f * y + k) + noise_amp * (2.0f * rand() / RAND_MAX - 1.0f);
            if (((x / 40 + y / 80) & 1) && k == 1) v += 40;
            int iv = (int)v; iv = iv < 0 ? 0 : (iv > 255 ? 255 : iv);
            ch[k] = iv / 32;
        }
        bin[y * W + x] = ch[0] + 8 * ch[1] + 64 * ch[2];
    }
    for (int o = 0; o < NOBJ; o++) {
        int n = 0;
        while (n < n_sel) { int b = rand() % 512; if (!smap[o * 512 + b]) smap[o * 512 + b] = ++n; }
    }
    std::vector<int> col(NOBJ * N), row(NOBJ * N);
    for (int o = 0; o < NOBJ; o++) {
        float cx = 100 + 60 * o, cy = 120 + 25 * (o % 4);
        for (int i = 0; i < N; i++) {
            float u1 = (rand() + 1.0f) / (RAND_MAX + 2.0f), u2 = (rand() + 1.0f) / (RAND_MAX + 2.0f);
            float g1 = sqrtf(-2 * logf(u1)) * cosf(6.2831853f * u2), g2 = sqrtf(-2 * logf(u1)) * sinf(6.2831853f * u2);
            int x = (int)lrintf(cx + 20 * g1), y = (int)lrintf(cy + 20 * g2);
            x = x < 1 ? 1 : (x > W - cols - 2 ? W - cols - 2 : x); y = y < 1 ? 1 : (y > H - rows - 2 ? H - rows - 2 : y);
            col[o * N + i] = x; row[o * N + i] = y;
        }
    }
    if (argc > 3 && atoi(argv[3])) {   // sort boxes by (row, col) per object
        for (int o = 0; o < NOBJ; o++) {
            std::vector<std::pair<int,int>> v(N);
            for (int i = 0; i < N; i++) v[i] = {row[o * N + i], col[o * N + i]};
            std::sort(v.begin(), v.end());
            for (int i = 0; i < N; i++) { row[o * N + i] = v[i].first; col[o * N + i] = v[i].second; }
        }
        printf("boxes sorted by (row, col)\n");
    }
    std::vector<int> hb(NOBJ * 4);
    for (int o = 0; o < NOBJ; o++) {
        int x0 = 1 << 30, y0 = 1 << 30, x1 = 0, y1 = 0;
        for (int i = 0; i < N; i++) { x0 = std::min(x0, col[o*N+i]); y0 = std::min(y0, row[o*N+i]); x1 = std::max(x1, col[o*N+i] + cols); y1 = std::max(y1, row[o*N+i] + rows); }
        hb[o*4] = x0; hb[o*4+1] = y0; hb[o*4+2] = x1; hb[o*4+3] = y1;
        if (o == 0) printf("object 0 region %d x %d\n", x1 - x0, y1 - y0);
    }
    Args a; a.W = W; a.H = H; a.N = N; a.rows = rows; a.cols = cols; a.n_slots = n_slots;
    uint16_t *d_bin, *d_smap; int *d_col, *d_row; uint32_t *d_out, *d_ref;
    size_t outb = (size_t)NOBJ * n_slots * N * 4;
    CK(cudaMalloc(&d_bin, bin.size() * 2)); CK(cudaMalloc(&d_smap, smap.size() * 2));
    CK(cudaMalloc(&d_col, col.size() * 4)); CK(cudaMalloc(&d_row, row.size() * 4));
    CK(cudaMalloc(&d_out, outb)); CK(cudaMalloc(&d_ref, outb));
    CK(cudaMemcpy(d_bin, bin.data(), bin.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_smap, smap.data(), smap.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_col, col.data(), col.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_row, row.data(), row.size() * 4, cudaMemcpyHostToDevice));
    a.binimg = d_bin; a.slotmap = d_smap; a.col = d_col; a.row = d_row;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<uint32_t> ref(outb / 4), got(outb / 4);

    auto run = [&](const char* name, auto launch, bool is_ref) {
        a.out = is_ref ? d_ref : d_out;
        CK(cudaMemset(a.out, 0xff, outb));
        for (int i = 0; i < 3; i++) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 10; i++) launch();
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        cudaError_t e = cudaGetLastError();
        const char* ok = "ref";
        if (is_ref) CK(cudaMemcpy(ref.data(), d_ref, outb, cudaMemcpyDeviceToHost));
        else { CK(cudaMemcpy(got.data(), d_out, outb, cudaMemcpyDeviceToHost)); ok = memcmp(ref.data(), got.data(), outb) ? "MISMATCH" : "match"; }
        printf("%-34s %8.1f us/launch  %s %s\n", name, 1e3 * ms / 10, ok, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    const size_t base = (size_t)rows * cols * 4 + 512 * 2;
    {
        size_t smem = base + 2 * n_slots * 32 * 4;
        CK(cudaFuncSetAttribute(v0_thread_per_box, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        run("v0 thread/box private hist", [&] { v0_thread_per_box<<<dim3((N + 63) / 64, NOBJ), 64, smem>>>(a); }, true);
    }
    const int nsp = (n_slots + 31) & ~31;
#define RUNW(MODE, WARPS, GX, label) { size_t smem = base + (size_t)WARPS * nsp * 4; \
        CK(cudaFuncSetAttribute(v_warp_per_box<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        run(label, [&] { v_warp_per_box<MODE, WARPS><<<dim3(GX, NOBJ), WARPS * 32, smem>>>(a); }, false); }
    RUNW(0, 8, 250, "w/box all-lane ATOMS      8w g250");
    RUNW(1, 8, 250, "w/box match+redux+ATOMS   8w g250");
    RUNW(2, 8, 250, "w/box match+redux+RMW     8w g250");
    RUNW(3, 8, 250, "w/box leaderloop+redux    8w g250");
    RUNW(1, 8, 74, "w/box match+redux+ATOMS   8w g74");
    RUNW(2, 8, 74, "w/box match+redux+RMW     8w g74");
    RUNW(1, 16, 37, "w/box match+redux+ATOMS  16w g37");
    RUNW(2, 16, 37, "w/box match+redux+RMW    16w g37");
    RUNW(2, 4, 148, "w/box match+redux+RMW     4w g148");
    RUNW(2, 32, 18, "w/box match+redux+RMW    32w g18");
    {
        size_t smem = base + 2 * n_slots * 32 * 4;
        CK(cudaFuncSetAttribute(v5_sub_warp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        run("v5 8 lanes/box private hists", [&] { v5_sub_warp<8><<<dim3((N + 7) / 8, NOBJ), 64, smem>>>(a); }, false);
        CK(cudaFuncSetAttribute(v5_sub_warp<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        run("v5 4 lanes/box private hists", [&] { v5_sub_warp<4><<<dim3((N + 15) / 16, NOBJ), 64, smem>>>(a); }, false);
    }
    {
        BBox* d_bb; CK(cudaMalloc(&d_bb, NOBJ * 16)); CK(cudaMemcpy(d_bb, hb.data(), NOBJ * 16, cudaMemcpyHostToDevice));
        // LUT on the host with the same formula
        std::vector<uint32_t> hl(rows * cols);
        { int np = rows * cols; int shift = __builtin_clz(np); if (shift > 20) shift = 20; float scale = (float)(1u << shift);
          float hx = cols / 2.0f, hy = rows / 2.0f; double ivx = 1.0 / (double)(2.0f * hx * hx), ivy = 1.0 / (double)(2.0f * hy * hy);
          for (int p = 0; p < np; p++) { int r = p / cols, c = p % cols; double dx = hx - (float)c, dy = hy - (float)r;
            float wd = (float)(dx * dx * ivx + dy * dy * ivy); hl[p] = (uint32_t)lrintf(expf(-wd) * scale); } }
        uint32_t* d_l; CK(cudaMalloc(&d_l, hl.size() * 4)); CK(cudaMemcpy(d_l, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice));
        int maxreg = 0;
        for (int o = 0; o < NOBJ; o++) { int RW = hb[o*4+2] - hb[o*4], RH = hb[o*4+3] - hb[o*4+1]; maxreg = std::max(maxreg, (((RW + 3) & ~3) + 4) * RH); }
#define RUN6(T, WARPS, label) { size_t smem = base + (size_t)WARPS * n_slots * 32 * 4 + maxreg + 16; \
        if (smem <= 227 * 1024) { CK(cudaFuncSetAttribute(v6_strips<T, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        const int bpc = WARPS * (32 / T); \
        run(label, [&] { v6_strips<T, WARPS><<<dim3((N + bpc - 1) / bpc, NOBJ), WARPS * 32, smem>>>(a, d_bb, d_l); }, false); \
        printf("   smem %zu B, grid %d x %d\n", smem, (N + bpc - 1) / bpc, NOBJ); } else printf("%s: smem %zu too large\n", label, smem); }
#define RUN7(WARPS, CPO, FWD, label) { size_t fixed = (size_t)(rows + 1) * ((cols + 3) & ~3) * 4 + 1024 + (size_t)WARPS * n_slots * 32 * 4; \
        size_t smem = fixed + maxreg + 8192; if (smem > 227 * 1024 - 64) smem = 227 * 1024 - 64; int cap = (int)(smem - fixed); \
        CK((cudaFuncSetAttribute(v7_forward<WARPS, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))); \
        run(label, [&] { v7_forward<WARPS, FWD><<<dim3(CPO, NOBJ), WARPS * 32, smem>>>(a, CPO, cap); }, false); \
        printf("   region cap %d B\n", cap); }
        RUN7(13, 18, 2, "v7 fwd2 13w 18 cta/obj");
        { unsigned long long hd[64]; CK(cudaMemcpyFromSymbol(hd, g_dbg, sizeof(hd))); printf("   stamps (us from kernel start of CTA(3,2) warp0):"); for (int k = 1; k <= 20; k++) if (hd[k]) printf(" [%d]%.2f", k, (double)(hd[k] - hd[0]) * 1e-3); printf("\n"); }
        RUN7(13, 18, 1, "v7 fwd1 13w 18 cta/obj");
        RUN7(16, 18, 2, "v7 fwd2 16w 18 cta/obj");
        RUN7(16, 18, 1, "v7 fwd1 16w 18 cta/obj");
        RUN7(12, 18, 4, "v7 fwd4 12w 18 cta/obj");
        RUN7(12, 14, 4, "v7 fwd4 12w 14 cta/obj");
        RUN7(8, 18, 4, "v7 fwd4  8w 18 cta/obj");
        RUN7(13, 18, 4, "v7 fwd4 13w 18 cta/obj");
        RUN7(13, 17, 4, "v7 fwd4 13w 17 cta/obj");
        RUN7(7, 37, 4, "v7 fwd4  7w 37 cta/obj");
        RUN7(6, 37, 4, "v7 fwd4  6w 37 cta/obj");
        RUN6(10, 4, "v6 10 lanes/box 4 cols  4w");
        RUN6(10, 8, "v6 10 lanes/box 4 cols  8w");
        RUN6(10, 12, "v6 10 lanes/box 4 cols 12w");
    }
    // distinct-slot statistics
    {
        double tot = 0; long cnt = 0;
        for (int i = 0; i < 200; i++) {
            int c0 = col[i], r0 = row[i];
            for (int p0 = 0; p0 < rows * cols; p0 += 32) {
                int seen[32], ns = 0;
                for (int l = 0; l < 32; l++) { int p = p0 + l, r = p / cols, c = p % cols; int s = smap[bin[(r0 + r) * W + c0 + c]];
                    bool f = false; for (int q = 0; q < ns; q++) f |= seen[q] == s; if (!f) seen[ns++] = s; }
                tot += ns; cnt++;
            }
        }
        printf("mean distinct slots per 32-pixel group: %.2f\n", tot / cnt);
    }
    // match / redux throughput: 148*8 CTAs x 256 threads
    {
        unsigned hin[32]; for (int i = 0; i < 32; i++) hin[i] = i / 8;
        unsigned *din, *dout; CK(cudaMalloc(&din, 128)); CK(cudaMalloc(&dout, 148 * 8 * 256 * 4));
        CK(cudaMemcpy(din, hin, 128, cudaMemcpyHostToDevice));
        const char* nm[3] = {"match_any", "redux.add(full)", "match_any+redux(mask)"};
        for (int mode = 0; mode < 3; mode++) {
            const int iters = 4096;
            mb_match<<<148 * 8, 256>>>(din, dout, 16, mode);
            CK(cudaEventRecord(e0));
            mb_match<<<148 * 8, 256>>>(din, dout, iters, mode);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            double warp_ops = 148.0 * 8 * 8 * iters;
            printf("%-24s %.2f ns per warp-op per SM  (%.1f clk @1.9GHz per SM)\n", nm[mode], ms * 1e6 / (warp_ops / 148), ms * 1e6 / (warp_ops / 148) * 1.9);
        }
    }
    return 0;
}
