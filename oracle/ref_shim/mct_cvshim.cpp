/* Implementation side of the OpenCV shim (see mct_cvshim.h).  TEST INFRASTRUCTURE ONLY.
 *
 * Restated from OpenCV 2.3.1's published behaviour, pinned against cv2 4.13 goldens where noted:
 *   cvCvtColor 8u BGR2GRAY / BGR2HSV      fixed point, tests/golden/cv2_gray.json, cv2_hsv.json
 *   cvInvert / cvDet (n <= 3)             closed forms in f64, result stored in the matrix type  (cv2_fuser.json)
 *   cvGEMM family on small matrices       f64 accumulation, one rounding to the matrix type      (cv2_fuser.json)
 *   cvSVD                                 one-sided Jacobi in f64 (only the null direction is read by the caller)
 *   cvKalmanPredict / cvKalmanCorrect     the textbook sequence OpenCV runs, gain by cvSolve(temp3, temp2, CV_SVD)
 */
#include "mct_cvshim.h"
#include <cstdio>
#include <vector>
#include <algorithm>

void mct_shim_not_on_path(const char* what)
{
    fprintf(stderr, "oracle/_ref shim: %s is outside the appearance-scoring path and has no implementation\n", what);
    abort();
}

extern "C" { uint64_t* mct_shim_rng_override = 0; }

/* ---- injected noise rows (cv::randn / cv::randu) ---- */
static const float* g_noise = 0;
static size_t g_noise_n = 0, g_noise_pos = 0;
extern "C" void mct_shim_set_noise(const float* rows, size_t n) { g_noise = rows; g_noise_n = n; g_noise_pos = 0; }
extern "C" size_t mct_shim_noise_consumed() { return g_noise_pos; }
void mct_shim_fill_noise(float* dst, int n, int, double, double)
{
    if (!g_noise || g_noise_pos + (size_t)n > g_noise_n) {
        fprintf(stderr, "oracle/_ref shim: cv::randn/randu called without injected noise (%zu + %d > %zu)\n", g_noise_pos, n, g_noise_n);
        abort();
    }
    memcpy(dst, g_noise + g_noise_pos, (size_t)n * sizeof(float));
    g_noise_pos += (size_t)n;
}

/* ---- IplImage ---- */
IplImage* cvInitImageHeader(IplImage* img, CvSize size, int depth, int channels, int origin, int align)
{
    memset(img, 0, sizeof(*img));
    img->nSize = sizeof(*img); img->nChannels = channels; img->depth = depth; img->origin = origin; img->align = align;
    img->width = size.width; img->height = size.height;
    int bpp = (depth & 255) / 8;
    img->widthStep = (size.width * channels * bpp + align - 1) & ~(align - 1);
    img->imageSize = img->widthStep * size.height;
    return img;
}
IplImage* cvCreateImageHeader(CvSize size, int depth, int channels)
{
    IplImage* img = (IplImage*)malloc(sizeof(IplImage));
    return cvInitImageHeader(img, size, depth, channels, 0, 4);
}
void cvCreateData(CvArr* arr)
{
    IplImage* img = (IplImage*)arr;
    img->imageData = img->imageDataOrigin = (char*)calloc((size_t)img->imageSize + 16, 1);
}
IplImage* cvCreateImage(CvSize size, int depth, int channels)
{
    IplImage* img = cvCreateImageHeader(size, depth, channels);
    cvCreateData(img);
    return img;
}
void cvReleaseImageHeader(IplImage** img) { if (img && *img) { free(*img); *img = 0; } }
void cvReleaseImage(IplImage** img)
{
    if (img && *img) { free((*img)->imageDataOrigin); free(*img); *img = 0; }
}
void cvCopy(const CvArr* src, CvArr* dst, const CvArr*)
{
    const IplImage* s = (const IplImage*)src; IplImage* d = (IplImage*)dst;
    for (int y = 0; y < s->height; y++) memcpy(d->imageData + (size_t)y * d->widthStep, s->imageData + (size_t)y * s->widthStep,
                                               (size_t)s->width * s->nChannels * ((s->depth & 255) / 8));
}
void cvEqualizeHist(const CvArr*, CvArr*) { mct_shim_not_on_path("cvEqualizeHist"); }

static inline int sat_round_div(double a, double b)
{
    double q = a / b;
    int v = _mm_cvtsd_si32(_mm_set_sd(q));
    return v;
}
static void cvt_row(const uchar* s, uchar* d, int n, int code)
{
    enum { HSV_SHIFT = 12 };
    static int sdiv[256], hdiv[256], init = 0;
    if (!init) {
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; i++) {
            sdiv[i] = sat_round_div((double)(255 << HSV_SHIFT), 1.0 * i);
            hdiv[i] = sat_round_div((double)(180 << HSV_SHIFT), 6.0 * i);
        }
        init = 1;
    }
    for (int x = 0; x < n; x++) {
        int c0 = s[3 * x], c1 = s[3 * x + 1], c2 = s[3 * x + 2];
        int b = c0, g = c1, r = c2;
        if (code == CV_RGB2GRAY || code == CV_RGB2HSV) { r = c0; b = c2; }
        if (code == CV_BGR2GRAY || code == CV_RGB2GRAY) { d[x] = (uchar)((r * 4899 + g * 9617 + b * 1868 + 8192) >> 14); continue; }
        int v = std::max(b, std::max(g, r)), vmin = std::min(b, std::min(g, r));
        int diff = v - vmin;
        int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
        int sat = (diff * sdiv[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT;
        int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
        h = (h * hdiv[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT;
        h += h < 0 ? 180 : 0;
        d[3 * x] = (uchar)h; d[3 * x + 1] = (uchar)sat; d[3 * x + 2] = (uchar)v;
    }
}
void cvCvtColor(const CvArr* src, CvArr* dst, int code)
{
    const IplImage* s = (const IplImage*)src; IplImage* d = (IplImage*)dst;
    for (int y = 0; y < s->height; y++)
        cvt_row((const uchar*)s->imageData + (size_t)y * s->widthStep, (uchar*)d->imageData + (size_t)y * d->widthStep, s->width, code);
}
namespace cv {
void cvtColor(const Mat& src, Mat& dst, int code)
{
    bool gray = (code == CV_BGR2GRAY || code == CV_RGB2GRAY);
    dst.create(src.rows, src.cols, gray ? CV_8UC1 : CV_8UC3);
    for (int y = 0; y < src.rows; y++) cvt_row(src.data + src.step * y, dst.data + dst.step * y, src.cols, code);
}
}

/* ---- CvMat ---- */
static int elem_size(int type) { return CV_MAT_DEPTH(type) == CV_64F ? 8 : CV_MAT_DEPTH(type) == CV_32F ? 4 : 1; }
CvMat* cvCreateMat(int rows, int cols, int type)
{
    CvMat* m = (CvMat*)calloc(1, sizeof(CvMat));
    m->type = type; m->rows = rows; m->cols = cols; m->step = cols * elem_size(type) * CV_MAT_CN(type);
    m->data.ptr = (uchar*)calloc((size_t)m->step * rows + 16, 1);
    m->refcount = (int*)m;          /* marks "owns data" */
    return m;
}
void cvReleaseMat(CvMat** m)
{
    if (m && *m) { if ((*m)->refcount) free((*m)->data.ptr); free(*m); *m = 0; }
}
CvMat cvMat(int rows, int cols, int type, void* data)
{
    CvMat m; memset(&m, 0, sizeof(m));
    m.type = type; m.rows = rows; m.cols = cols; m.step = cols * elem_size(type); m.data.ptr = (uchar*)data;
    return m;
}
CvMat* cvCloneMat(const CvMat* s)
{
    CvMat* m = cvCreateMat(s->rows, s->cols, s->type);
    for (int r = 0; r < s->rows; r++) memcpy(m->data.ptr + (size_t)m->step * r, s->data.ptr + (size_t)s->step * r, (size_t)m->step);
    return m;
}
static inline double G(const CvMat* m, int r, int c) { return cvmGet(m, r, c); }
static inline void S(CvMat* m, int r, int c, double v) { cvmSet(m, r, c, v); }

double cvDet(const CvArr* arr)
{
    const CvMat* m = (const CvMat*)arr;
    if (m->rows == 1) return G(m, 0, 0);
    if (m->rows == 2) return G(m, 0, 0) * G(m, 1, 1) - G(m, 0, 1) * G(m, 1, 0);
    if (m->rows == 3)
        return G(m, 0, 0) * (G(m, 1, 1) * G(m, 2, 2) - G(m, 1, 2) * G(m, 2, 1)) - G(m, 0, 1) * (G(m, 1, 0) * G(m, 2, 2) - G(m, 1, 2) * G(m, 2, 0)) +
               G(m, 0, 2) * (G(m, 1, 0) * G(m, 2, 1) - G(m, 1, 1) * G(m, 2, 0));
    mct_shim_not_on_path("cvDet n > 3");
    return 0;
}
double cvInvert(const CvArr* sa, CvArr* da, int)
{
    const CvMat* s = (const CvMat*)sa; CvMat* d = (CvMat*)da;
    int n = s->rows;
    if (n == 1) { double v = G(s, 0, 0); if (v == 0) { S(d, 0, 0, 0); return 0; } S(d, 0, 0, 1.0 / v); return v; }
    if (n == 2) {
        double det = G(s, 0, 0) * G(s, 1, 1) - G(s, 0, 1) * G(s, 1, 0);
        if (det == 0) { for (int i = 0; i < 4; i++) S(d, i / 2, i % 2, 0); return 0; }
        double id = 1.0 / det;
        double t00 = G(s, 1, 1) * id, t01 = -G(s, 0, 1) * id, t10 = -G(s, 1, 0) * id, t11 = G(s, 0, 0) * id;
        S(d, 0, 0, t00); S(d, 0, 1, t01); S(d, 1, 0, t10); S(d, 1, 1, t11);
        return det;
    }
    if (n == 3) {
        double H[3][3];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) H[r][c] = G(s, r, c);
        double det = cvDet(s);
        if (det == 0) { for (int i = 0; i < 9; i++) S(d, i / 3, i % 3, 0); return 0; }
        double id = 1.0 / det, t[3][3];
        t[0][0] = (H[1][1] * H[2][2] - H[1][2] * H[2][1]) * id;
        t[0][1] = (H[0][2] * H[2][1] - H[0][1] * H[2][2]) * id;
        t[0][2] = (H[0][1] * H[1][2] - H[0][2] * H[1][1]) * id;
        t[1][0] = (H[1][2] * H[2][0] - H[1][0] * H[2][2]) * id;
        t[1][1] = (H[0][0] * H[2][2] - H[0][2] * H[2][0]) * id;
        t[1][2] = (H[0][2] * H[1][0] - H[0][0] * H[1][2]) * id;
        t[2][0] = (H[1][0] * H[2][1] - H[1][1] * H[2][0]) * id;
        t[2][1] = (H[0][1] * H[2][0] - H[0][0] * H[2][1]) * id;
        t[2][2] = (H[0][0] * H[1][1] - H[0][1] * H[1][0]) * id;
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) S(d, r, c, t[r][c]);
        return det;
    }
    mct_shim_not_on_path("cvInvert n > 3");
    return 0;
}

/* dst = alpha * op(A) * op(B) + beta * C, f64 accumulation, one rounding at the store */
void cvGEMM(const CvArr* aa, const CvArr* ba, double alpha, const CvArr* ca, double beta, CvArr* da, int tABC)
{
    const CvMat *A = (const CvMat*)aa, *B = (const CvMat*)ba, *C = (const CvMat*)ca; CvMat* D = (CvMat*)da;
    bool tA = tABC & 1, tB = tABC & 2;
    int M = tA ? A->cols : A->rows, K = tA ? A->rows : A->cols, N = tB ? B->rows : B->cols;
    std::vector<double> out((size_t)M * N);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            double acc = 0;
            for (int k = 0; k < K; k++) acc += (tA ? G(A, k, i) : G(A, i, k)) * (tB ? G(B, j, k) : G(B, k, j));
            acc *= alpha;
            if (C) acc += beta * G(C, i, j);
            out[(size_t)i * N + j] = acc;
        }
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) S(D, i, j, out[(size_t)i * N + j]);
}
void cvAdd(const CvArr* a, const CvArr* b, CvArr* d, const CvArr*)
{
    const CvMat *A = (const CvMat*)a, *B = (const CvMat*)b; CvMat* D = (CvMat*)d;
    for (int r = 0; r < A->rows; r++) for (int c = 0; c < A->cols; c++) {
        if (CV_MAT_DEPTH(D->type) == CV_32F) S(D, r, c, (float)G(A, r, c) + (float)G(B, r, c)); else S(D, r, c, G(A, r, c) + G(B, r, c));
    }
}
void cvSub(const CvArr* a, const CvArr* b, CvArr* d, const CvArr*)
{
    const CvMat *A = (const CvMat*)a, *B = (const CvMat*)b; CvMat* D = (CvMat*)d;
    for (int r = 0; r < A->rows; r++) for (int c = 0; c < A->cols; c++) {
        if (CV_MAT_DEPTH(D->type) == CV_32F) S(D, r, c, (float)G(A, r, c) - (float)G(B, r, c)); else S(D, r, c, G(A, r, c) - G(B, r, c));
    }
}
void cvTranspose(const CvArr* a, CvArr* d)
{
    const CvMat* A = (const CvMat*)a; CvMat* D = (CvMat*)d;
    std::vector<double> t((size_t)A->rows * A->cols);
    for (int r = 0; r < A->rows; r++) for (int c = 0; c < A->cols; c++) t[(size_t)c * A->rows + r] = G(A, r, c);
    for (int r = 0; r < A->cols; r++) for (int c = 0; c < A->rows; c++) S(D, r, c, t[(size_t)r * A->rows + c]);
}
void cvSetIdentity(CvArr* a, CvScalar v)
{
    CvMat* A = (CvMat*)a;
    for (int r = 0; r < A->rows; r++) for (int c = 0; c < A->cols; c++) S(A, r, c, r == c ? v.val[0] : 0.0);
}
void cvSetZero(CvArr* a)
{
    CvMat* A = (CvMat*)a;
    for (int r = 0; r < A->rows; r++) memset(A->data.ptr + (size_t)A->step * r, 0, (size_t)A->cols * elem_size(A->type));
}
void cvCopyMat(const CvMat* s, CvMat* d)
{
    for (int r = 0; r < s->rows; r++) for (int c = 0; c < s->cols; c++) S(d, r, c, G(s, r, c));
}

/* A (m x n, n <= 8) = U diag(W) V^T; one-sided Jacobi on the columns in f64; singular values descending.  Only W and V
 * are produced with full accuracy; U is returned for m >= n. */
void cvSVD(CvArr* aa, CvArr* wa, CvArr* ua, CvArr* va, int flags)
{
    CvMat* A = (CvMat*)aa; CvMat* W = (CvMat*)wa; CvMat* U = (CvMat*)ua; CvMat* V = (CvMat*)va;
    int m = A->rows, n = A->cols, mp = std::max(m, n);
    std::vector<double> a((size_t)mp * n, 0.0), v((size_t)n * n, 0.0);
    for (int r = 0; r < m; r++) for (int c = 0; c < n; c++) a[(size_t)r * n + c] = G(A, r, c);
    for (int i = 0; i < n; i++) v[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        bool changed = false;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int r = 0; r < mp; r++) { double x = a[(size_t)r * n + p], y = a[(size_t)r * n + q]; alpha += x * x; beta += y * y; gamma += x * y; }
                if (std::fabs(gamma) <= 2.220446049250313e-16 * std::sqrt(alpha * beta) || gamma == 0) continue;
                changed = true;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < mp; r++) { double x = a[(size_t)r * n + p], y = a[(size_t)r * n + q]; a[(size_t)r * n + p] = c * x - s * y; a[(size_t)r * n + q] = s * x + c * y; }
                for (int r = 0; r < n; r++) { double x = v[(size_t)r * n + p], y = v[(size_t)r * n + q]; v[(size_t)r * n + p] = c * x - s * y; v[(size_t)r * n + q] = s * x + c * y; }
            }
        if (!changed) break;
    }
    std::vector<double> w(n);
    std::vector<int> ord(n);
    for (int c = 0; c < n; c++) { double s = 0; for (int r = 0; r < mp; r++) s += a[(size_t)r * n + c] * a[(size_t)r * n + c]; w[c] = std::sqrt(s); ord[c] = c; }
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return w[x] > w[y]; });
    if (W) for (int i = 0; i < std::min(n, W->rows * W->cols); i++) {
        if (W->rows == W->cols && W->rows > 1) S(W, i, i, w[ord[i]]); else if (W->cols == 1) S(W, i, 0, w[ord[i]]); else S(W, 0, i, w[ord[i]]);
    }
    if (V) for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) {
        if (flags & 4 /* CV_SVD_V_T */) S(V, c, r, v[(size_t)r * n + ord[c]]); else S(V, r, c, v[(size_t)r * n + ord[c]]);
    }
    if (U && U->rows == m) for (int r = 0; r < m; r++) for (int c = 0; c < std::min(n, U->cols); c++) {
        double s = w[ord[c]];
        S(U, r, c, s > 0 ? a[(size_t)r * n + ord[c]] / s : 0.0);
    }
}

/* ---- Kalman (cvCreateKalman(DP, MP, CP); all CV_32F) ---- */
CvKalman* cvCreateKalman(int DP, int MP, int CP)
{
    CvKalman* k = (CvKalman*)calloc(1, sizeof(CvKalman));
    k->DP = DP; k->MP = MP; k->CP = CP;
    k->state_pre = cvCreateMat(DP, 1, CV_32FC1);
    k->state_post = cvCreateMat(DP, 1, CV_32FC1);
    k->transition_matrix = cvCreateMat(DP, DP, CV_32FC1); cvSetIdentity(k->transition_matrix, cvRealScalar(1));
    k->process_noise_cov = cvCreateMat(DP, DP, CV_32FC1); cvSetIdentity(k->process_noise_cov, cvRealScalar(1));
    k->measurement_matrix = cvCreateMat(MP, DP, CV_32FC1);
    k->measurement_noise_cov = cvCreateMat(MP, MP, CV_32FC1); cvSetIdentity(k->measurement_noise_cov, cvRealScalar(1));
    k->error_cov_pre = cvCreateMat(DP, DP, CV_32FC1);
    k->error_cov_post = cvCreateMat(DP, DP, CV_32FC1);
    k->gain = cvCreateMat(DP, MP, CV_32FC1);
    if (CP > 0) k->control_matrix = cvCreateMat(DP, CP, CV_32FC1);
    k->temp1 = cvCreateMat(DP, DP, CV_32FC1);
    k->temp2 = cvCreateMat(MP, DP, CV_32FC1);
    k->temp3 = cvCreateMat(MP, MP, CV_32FC1);
    k->temp4 = cvCreateMat(MP, DP, CV_32FC1);
    k->temp5 = cvCreateMat(MP, 1, CV_32FC1);
    return k;
}
void cvReleaseKalman(CvKalman** pk)
{
    if (!pk || !*pk) return;
    CvKalman* k = *pk;
    CvMat** ms[] = {&k->state_pre, &k->state_post, &k->transition_matrix, &k->process_noise_cov, &k->measurement_matrix,
                    &k->measurement_noise_cov, &k->error_cov_pre, &k->error_cov_post, &k->gain, &k->control_matrix,
                    &k->temp1, &k->temp2, &k->temp3, &k->temp4, &k->temp5};
    for (auto m : ms) if (*m) cvReleaseMat(m);
    free(k); *pk = 0;
}
const CvMat* cvKalmanPredict(CvKalman* k, const CvMat* control)
{
    cvGEMM(k->transition_matrix, k->state_post, 1, 0, 0, k->state_pre, 0);
    if (control && k->CP > 0) cvGEMM(k->control_matrix, control, 1, k->state_pre, 1, k->state_pre, 0);
    cvGEMM(k->transition_matrix, k->error_cov_post, 1, 0, 0, k->temp1, 0);
    cvGEMM(k->temp1, k->transition_matrix, 1, k->process_noise_cov, 1, k->error_cov_pre, 2 /* B^T */);
    cvCopyMat(k->state_pre, k->state_post);          /* OpenCV >= 2.3 keeps post = pre after a predict */
    cvCopyMat(k->error_cov_pre, k->error_cov_post);
    return k->state_pre;
}
const CvMat* cvKalmanCorrect(CvKalman* k, const CvMat* z)
{
    int MP = k->MP, DP = k->DP;
    cvGEMM(k->measurement_matrix, k->error_cov_pre, 1, 0, 0, k->temp2, 0);                           /* H P' */
    cvGEMM(k->temp2, k->measurement_matrix, 1, k->measurement_noise_cov, 1, k->temp3, 2);            /* H P' H^T + R */
    /* temp4 = inv(temp3) temp2, solved in f64 (cvSolve CV_SVD on a 2x2 SPD system) */
    if (MP == 2) {
        double a = G(k->temp3, 0, 0), b = G(k->temp3, 0, 1), c = G(k->temp3, 1, 0), d = G(k->temp3, 1, 1), det = a * d - b * c;
        for (int j = 0; j < DP; j++) {
            double r0 = G(k->temp2, 0, j), r1 = G(k->temp2, 1, j);
            S(k->temp4, 0, j, (d * r0 - b * r1) / det);
            S(k->temp4, 1, j, (a * r1 - c * r0) / det);
        }
    } else if (MP == 1) {
        for (int j = 0; j < DP; j++) S(k->temp4, 0, j, G(k->temp2, 0, j) / G(k->temp3, 0, 0));
    } else mct_shim_not_on_path("cvKalmanCorrect MP > 2");
    cvTranspose(k->temp4, k->gain);                                                                  /* K */
    cvGEMM(k->measurement_matrix, k->state_pre, -1, z, 1, k->temp5, 0);                              /* z - H x' */
    cvGEMM(k->gain, k->temp5, 1, k->state_pre, 1, k->state_post, 0);                                 /* x' + K (.) */
    cvGEMM(k->gain, k->temp2, -1, k->error_cov_pre, 1, k->error_cov_post, 0);                        /* P' - K H P' */
    return k->state_post;
}
