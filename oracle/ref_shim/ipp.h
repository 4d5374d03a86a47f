/* ipp.h shim: the Intel IPP entry points the reference calls (Matrix.h, Matrix.cpp), restated in plain C++.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Intel IPP is closed source and not in this image; what is
 * stated here is the documented semantics of each call.  Stated assumptions about IPP-internal arithmetic:
 *   - ippiIntegral_8u32f_C1R: exact integer prefix sums, rounded once to f32 (IPP's accumulation order is
 *     not published);
 *   - ippiSum_32f / ippiMean_32f / ippiMean_StdDev_32f: f64 accumulators in row-major order (the results are
 *     Ipp64f; the hint "Fast" permits any order);
 *   - the 8u "Sfs" variants: integer result, scaled by 2^-scaleFactor with round-half-even, saturated.
 * Everything else (copy/set/elementwise f32) is fully determined by the documentation. */
#ifndef MCT_SHIM_IPP_H
#define MCT_SHIM_IPP_H
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>

typedef unsigned char Ipp8u;
typedef float Ipp32f;
typedef int Ipp32s;
typedef double Ipp64f;
typedef int IppStatus;
enum { ippStsNoErr = 0 };
typedef enum { ippAlgHintNone, ippAlgHintFast, ippAlgHintAccurate } IppHintAlgorithm;
typedef struct { int width, height; } IppiSize;
typedef struct { int x, y, width, height; } IppiRect;
typedef struct { int x, y; } IppiPoint;
enum { IPPI_INTER_NN = 1, IPPI_INTER_LINEAR = 2, IPPI_INTER_CUBIC = 4, IPPI_INTER_SUPER = 8, IPPI_INTER_LANCZOS = 16 };
typedef enum { ippAxsHorizontal, ippAxsVertical, ippAxsBoth } IppiAxis;

static inline int mct_shim_step(int bytes) { return (bytes + 63) & ~63; }
static inline Ipp8u* ippiMalloc_8u_C1(int w, int h, int* step) {
    *step = mct_shim_step(w);
    void* p = NULL;
    if (posix_memalign(&p, 64, (size_t)(*step) * (h > 0 ? h : 1))) return NULL;
    return (Ipp8u*)p;
}
static inline Ipp32f* ippiMalloc_32f_C1(int w, int h, int* step) {
    *step = mct_shim_step(w * 4);
    void* p = NULL;
    if (posix_memalign(&p, 64, (size_t)(*step) * (h > 0 ? h : 1))) return NULL;
    return (Ipp32f*)p;
}
static inline void ippiFree(void* p) { free(p); }
static inline void ippsFree(void* p) { free(p); }

#define MCT_ROW8(p, step, y)  ((Ipp8u*)(p) + (size_t)(y) * (step))
#define MCT_ROW8C(p, step, y) ((const Ipp8u*)(p) + (size_t)(y) * (step))
#define MCT_ROWF(p, step, y)  ((Ipp32f*)((char*)(p) + (size_t)(y) * (step)))
#define MCT_ROWFC(p, step, y) ((const Ipp32f*)((const char*)(p) + (size_t)(y) * (step)))

static inline IppStatus ippiCopy_8u_C1R(const Ipp8u* s, int ss, Ipp8u* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) memcpy(MCT_ROW8(d, ds, y), MCT_ROW8C(s, ss, y), r.width);
    return ippStsNoErr;
}
static inline IppStatus ippiCopy_32f_C1R(const Ipp32f* s, int ss, Ipp32f* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) memcpy(MCT_ROWF(d, ds, y), MCT_ROWFC(s, ss, y), (size_t)r.width * 4);
    return ippStsNoErr;
}
static inline IppStatus ippiSet_8u_C1R(Ipp8u v, Ipp8u* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) memset(MCT_ROW8(d, ds, y), v, r.width);
    return ippStsNoErr;
}
static inline IppStatus ippiSet_32f_C1R(Ipp32f v, Ipp32f* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) { Ipp32f* p = MCT_ROWF(d, ds, y); for (int x = 0; x < r.width; x++) p[x] = v; }
    return ippStsNoErr;
}

/* dst is (width+1) x (height+1); first row and first column = val (Matrix.cpp:42 passes 0) */
static inline IppStatus ippiIntegral_8u32f_C1R(const Ipp8u* s, int ss, Ipp32f* d, int ds, IppiSize r, Ipp32f val) {
    uint64_t* col = (uint64_t*)calloc((size_t)r.width + 1, sizeof(uint64_t));
    Ipp32f* p0 = MCT_ROWF(d, ds, 0);
    for (int x = 0; x <= r.width; x++) p0[x] = val;
    for (int y = 0; y < r.height; y++) {
        const Ipp8u* sp = MCT_ROW8C(s, ss, y);
        Ipp32f* dp = MCT_ROWF(d, ds, y + 1);
        uint64_t run = 0;
        dp[0] = val;
        for (int x = 0; x < r.width; x++) {
            run += sp[x];
            col[x + 1] += run;
            dp[x + 1] = (Ipp32f)((double)col[x + 1] + (double)val);
        }
    }
    free(col);
    return ippStsNoErr;
}

static inline int mct_shim_sfs(long long v, int sf) {      /* v * 2^-sf, round half even, saturate to u8 */
    if (sf > 0) {
        long long half = 1LL << (sf - 1), q = v >> sf, rem = v & ((1LL << sf) - 1);
        if (rem > half || (rem == half && (q & 1))) q++;
        v = q;
    } else if (sf < 0) v <<= -sf;
    return v < 0 ? 0 : v > 255 ? 255 : (int)v;
}
#define MCT_BIN8(NAME, EXPR)                                                                                   \
    static inline IppStatus NAME(const Ipp8u* a, int as, const Ipp8u* b, int bs, Ipp8u* d, int ds, IppiSize r, \
                                 int sf) {                                                                     \
        for (int y = 0; y < r.height; y++)                                                                     \
            for (int x = 0; x < r.width; x++) {                                                                \
                long long A = MCT_ROW8C(a, as, y)[x], B = MCT_ROW8C(b, bs, y)[x];                              \
                MCT_ROW8(d, ds, y)[x] = (Ipp8u)mct_shim_sfs(EXPR, sf);                                         \
            }                                                                                                  \
        return ippStsNoErr;                                                                                    \
    }
MCT_BIN8(ippiAdd_8u_C1RSfs, A + B)
MCT_BIN8(ippiMul_8u_C1RSfs, A * B)
#define MCT_BINC8(NAME, EXPR)                                                                         \
    static inline IppStatus NAME(const Ipp8u* a, int as, Ipp8u c, Ipp8u* d, int ds, IppiSize r, int sf) { \
        for (int y = 0; y < r.height; y++)                                                            \
            for (int x = 0; x < r.width; x++) {                                                       \
                long long A = MCT_ROW8C(a, as, y)[x], B = c;                                          \
                MCT_ROW8(d, ds, y)[x] = (Ipp8u)mct_shim_sfs(EXPR, sf);                                \
            }                                                                                         \
        return ippStsNoErr;                                                                           \
    }
MCT_BINC8(ippiAddC_8u_C1RSfs, A + B)
MCT_BINC8(ippiMulC_8u_C1RSfs, A * B)
static inline IppStatus ippiSqr_8u_C1RSfs(const Ipp8u* a, int as, Ipp8u* d, int ds, IppiSize r, int sf) {
    for (int y = 0; y < r.height; y++)
        for (int x = 0; x < r.width; x++) { long long A = MCT_ROW8C(a, as, y)[x]; MCT_ROW8(d, ds, y)[x] = (Ipp8u)mct_shim_sfs(A * A, sf); }
    return ippStsNoErr;
}
static inline IppStatus ippiExp_8u_C1RSfs(const Ipp8u* a, int as, Ipp8u* d, int ds, IppiSize r, int sf) {
    for (int y = 0; y < r.height; y++)
        for (int x = 0; x < r.width; x++) {
            double v = std::exp((double)MCT_ROW8C(a, as, y)[x]) * std::ldexp(1.0, -sf);
            v = std::nearbyint(v);
            MCT_ROW8(d, ds, y)[x] = (Ipp8u)(v > 255 ? 255 : v);
        }
    return ippStsNoErr;
}

/* f32 elementwise: one rounded f32 operation per element */
static inline IppStatus ippiAdd_32f_C1R(const Ipp32f* a, int as, const Ipp32f* b, int bs, Ipp32f* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) { const Ipp32f *A = MCT_ROWFC(a, as, y), *B = MCT_ROWFC(b, bs, y); Ipp32f* D = MCT_ROWF(d, ds, y);
        for (int x = 0; x < r.width; x++) D[x] = A[x] + B[x]; }
    return ippStsNoErr;
}
static inline IppStatus ippiMul_32f_C1R(const Ipp32f* a, int as, const Ipp32f* b, int bs, Ipp32f* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) { const Ipp32f *A = MCT_ROWFC(a, as, y), *B = MCT_ROWFC(b, bs, y); Ipp32f* D = MCT_ROWF(d, ds, y);
        for (int x = 0; x < r.width; x++) D[x] = A[x] * B[x]; }
    return ippStsNoErr;
}
static inline IppStatus ippiAddC_32f_C1R(const Ipp32f* a, int as, Ipp32f c, Ipp32f* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) { const Ipp32f* A = MCT_ROWFC(a, as, y); Ipp32f* D = MCT_ROWF(d, ds, y);
        for (int x = 0; x < r.width; x++) D[x] = A[x] + c; }
    return ippStsNoErr;
}
static inline IppStatus ippiMulC_32f_C1R(const Ipp32f* a, int as, Ipp32f c, Ipp32f* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) { const Ipp32f* A = MCT_ROWFC(a, as, y); Ipp32f* D = MCT_ROWF(d, ds, y);
        for (int x = 0; x < r.width; x++) D[x] = A[x] * c; }
    return ippStsNoErr;
}
static inline IppStatus ippiSqr_32f_C1R(const Ipp32f* a, int as, Ipp32f* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) { const Ipp32f* A = MCT_ROWFC(a, as, y); Ipp32f* D = MCT_ROWF(d, ds, y);
        for (int x = 0; x < r.width; x++) D[x] = A[x] * A[x]; }
    return ippStsNoErr;
}
static inline IppStatus ippiExp_32f_C1R(const Ipp32f* a, int as, Ipp32f* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) { const Ipp32f* A = MCT_ROWFC(a, as, y); Ipp32f* D = MCT_ROWF(d, ds, y);
        for (int x = 0; x < r.width; x++) D[x] = expf(A[x]); }
    return ippStsNoErr;
}
static inline IppStatus ippiTranspose_8u_C1R(const Ipp8u* s, int ss, Ipp8u* d, int ds, IppiSize r) {
    for (int y = 0; y < r.height; y++) for (int x = 0; x < r.width; x++) MCT_ROW8(d, ds, x)[y] = MCT_ROW8C(s, ss, y)[x];
    return ippStsNoErr;
}

/* reductions */
#define MCT_MINMAX(NAME, T, ROW, CMP)                                                    \
    static inline IppStatus NAME(const T* s, int ss, IppiSize r, T* out) {               \
        T m = ROW(s, ss, 0)[0];                                                          \
        for (int y = 0; y < r.height; y++) for (int x = 0; x < r.width; x++) { T v = ROW(s, ss, y)[x]; if (v CMP m) m = v; } \
        *out = m; return ippStsNoErr;                                                    \
    }
MCT_MINMAX(ippiMax_8u_C1R, Ipp8u, MCT_ROW8C, >)
MCT_MINMAX(ippiMin_8u_C1R, Ipp8u, MCT_ROW8C, <)
MCT_MINMAX(ippiMax_32f_C1R, Ipp32f, MCT_ROWFC, >)
MCT_MINMAX(ippiMin_32f_C1R, Ipp32f, MCT_ROWFC, <)
#define MCT_MINMAXI(NAME, T, ROW, CMP)                                                   \
    static inline IppStatus NAME(const T* s, int ss, IppiSize r, T* out, int* ix, int* iy) { \
        T m = ROW(s, ss, 0)[0]; *ix = 0; *iy = 0;                                        \
        for (int y = 0; y < r.height; y++) for (int x = 0; x < r.width; x++) { T v = ROW(s, ss, y)[x]; if (v CMP m) { m = v; *ix = x; *iy = y; } } \
        *out = m; return ippStsNoErr;                                                    \
    }
MCT_MINMAXI(ippiMaxIndx_8u_C1R, Ipp8u, MCT_ROW8C, >)
MCT_MINMAXI(ippiMinIndx_8u_C1R, Ipp8u, MCT_ROW8C, <)
MCT_MINMAXI(ippiMaxIndx_32f_C1R, Ipp32f, MCT_ROWFC, >)
MCT_MINMAXI(ippiMinIndx_32f_C1R, Ipp32f, MCT_ROWFC, <)

static inline IppStatus ippiSum_8u_C1R(const Ipp8u* s, int ss, IppiSize r, Ipp64f* sum) {
    double a = 0; for (int y = 0; y < r.height; y++) for (int x = 0; x < r.width; x++) a += MCT_ROW8C(s, ss, y)[x];
    *sum = a; return ippStsNoErr;
}
static inline IppStatus ippiSum_32f_C1R(const Ipp32f* s, int ss, IppiSize r, Ipp64f* sum, IppHintAlgorithm) {
    double a = 0; for (int y = 0; y < r.height; y++) for (int x = 0; x < r.width; x++) a += (double)MCT_ROWFC(s, ss, y)[x];
    *sum = a; return ippStsNoErr;
}
static inline IppStatus ippiMean_8u_C1R(const Ipp8u* s, int ss, IppiSize r, Ipp64f* mean) {
    double a; ippiSum_8u_C1R(s, ss, r, &a); *mean = a / ((double)r.width * r.height); return ippStsNoErr;
}
static inline IppStatus ippiMean_32f_C1R(const Ipp32f* s, int ss, IppiSize r, Ipp64f* mean, IppHintAlgorithm h) {
    double a; ippiSum_32f_C1R(s, ss, r, &a, h); *mean = a / ((double)r.width * r.height); return ippStsNoErr;
}
/* population standard deviation about the mean, f64 */
static inline IppStatus ippiMean_StdDev_32f_C1R(const Ipp32f* s, int ss, IppiSize r, Ipp64f* mean, Ipp64f* sd) {
    double n = (double)r.width * r.height, a = 0, q = 0;
    for (int y = 0; y < r.height; y++) for (int x = 0; x < r.width; x++) a += (double)MCT_ROWFC(s, ss, y)[x];
    a /= n;
    for (int y = 0; y < r.height; y++) for (int x = 0; x < r.width; x++) { double dlt = (double)MCT_ROWFC(s, ss, y)[x] - a; q += dlt * dlt; }
    if (mean) *mean = a;
    if (sd) *sd = std::sqrt(q / n);
    return ippStsNoErr;
}
static inline IppStatus ippiMean_StdDev_8u_C1R(const Ipp8u* s, int ss, IppiSize r, Ipp64f* mean, Ipp64f* sd) {
    double n = (double)r.width * r.height, a = 0, q = 0;
    for (int y = 0; y < r.height; y++) for (int x = 0; x < r.width; x++) a += MCT_ROW8C(s, ss, y)[x];
    a /= n;
    for (int y = 0; y < r.height; y++) for (int x = 0; x < r.width; x++) { double dlt = MCT_ROW8C(s, ss, y)[x] - a; q += dlt * dlt; }
    if (mean) *mean = a;
    if (sd) *sd = std::sqrt(q / n);
    return ippStsNoErr;
}

/* Geometry / filtering entry points of Matrix.cpp's display and warping helpers: not on the appearance-scoring
 * path, never called by the harness.  Declared so the untouched sources compile; calling one aborts. */
void mct_shim_not_on_path(const char* what);
#define MCT_OFFPATH(NAME, ...) static inline IppStatus NAME(__VA_ARGS__) { mct_shim_not_on_path(#NAME); return -1; }
MCT_OFFPATH(ippiResizeGetBufSize, IppiRect, IppiRect, int, int, int*)
MCT_OFFPATH(ippiResizeSqrPixel_8u_C1R, const Ipp8u*, IppiSize, int, IppiRect, Ipp8u*, int, IppiRect, double, double, double, double, int, Ipp8u*)
MCT_OFFPATH(ippiResize_8u_C1R, const Ipp8u*, IppiSize, int, IppiRect, Ipp8u*, int, IppiSize, double, double, int)
MCT_OFFPATH(ippiGetAffineTransform, IppiRect, const double[4][2], double[2][3])
MCT_OFFPATH(ippiWarpAffine_8u_C1R, const Ipp8u*, IppiSize, int, IppiRect, Ipp8u*, int, IppiRect, const double[2][3], int)
MCT_OFFPATH(ippiFilterRow_8u_C1R, const Ipp8u*, int, Ipp8u*, int, IppiSize, const Ipp32s*, int, int, int)
MCT_OFFPATH(ippiFilterColumn_8u_C1R, const Ipp8u*, int, Ipp8u*, int, IppiSize, const int*, int, int, int)
#endif
