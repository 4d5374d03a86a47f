/* OpenCV 2.3.1 C / C++ API shim for building the UNMODIFIED reference sources (oracle/_ref).
 *
 * TEST INFRASTRUCTURE ONLY.  OpenCV's C headers are not in this image.  What the appearance-scoring path really
 * uses is restated here from OpenCV's published behaviour and pinned against cv2 4.13 in tests/golden/:
 *   cvRNG / cvRandInt / cvRandReal  multiply-with-carry, A = 4164903690 (cxcore types_c.h)      [cv2_rng.json]
 *   cvRound                         SSE2 cvtsd2si = round half to even                           [cv2_round.json]
 *   cvInvert / cvDet                small f32 / f64 matrices (closed forms for n <= 3, as OpenCV has them)
 *   cvCvtColor BGR2GRAY / BGR2HSV   8-bit fixed point formulas                                   [cv2_gray/hsv.json]
 *   cv::randn / cv::randu           NOT restated: OpenCV's ziggurat stream on theRNG() is outside the bit-exact
 *                                   claim; the harness injects the noise rows through
 *                                   mct_shim_set_noise(), the same rows the oracle and the GPU path receive.
 * Display, video IO, the face detector and drawing are declared so the sources compile; they do nothing (display,
 * drawing, cvWaitKey) or abort when called (IO). */
#ifndef MCT_CVSHIM_H
#define MCT_CVSHIM_H
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <emmintrin.h>
#include <memory>

void mct_shim_not_on_path(const char* what);

/* ---- basic types ---- */
typedef uint64_t CvRNG;
typedef unsigned char uchar;
typedef struct CvPoint { int x, y; } CvPoint;
typedef struct CvPoint2D32f { float x, y; } CvPoint2D32f;
typedef struct CvSize { int width, height; } CvSize;
typedef struct CvRect { int x, y, width, height; } CvRect;
typedef struct CvScalar { double val[4]; } CvScalar;
typedef struct CvFont { int dummy; } CvFont;
typedef struct CvCapture CvCapture;
typedef struct CvVideoWriter CvVideoWriter;
typedef struct CvMemStorage { int dummy; } CvMemStorage;
typedef struct CvSeq { int total; } CvSeq;
typedef struct CvHaarClassifierCascade { int dummy; } CvHaarClassifierCascade;
typedef void CvArr;

static inline CvPoint cvPoint(int x, int y) { CvPoint p = {x, y}; return p; }
static inline CvSize cvSize(int w, int h) { CvSize s = {w, h}; return s; }
static inline CvRect cvRect(int x, int y, int w, int h) { CvRect r = {x, y, w, h}; return r; }
static inline CvScalar cvScalar(double a, double b = 0, double c = 0, double d = 0) { CvScalar s = {{a, b, c, d}}; return s; }
static inline CvScalar cvScalarAll(double a) { return cvScalar(a, a, a, a); }
static inline CvScalar cvRealScalar(double a) { return cvScalar(a); }
#define CV_RGB(r, g, b) cvScalar((b), (g), (r), 0)
#define CV_AA 16
#define CV_FONT_HERSHEY_SIMPLEX 0
#define CV_WINDOW_AUTOSIZE 1
#define CV_FOURCC(a, b, c, d) (((a) & 255) + (((b) & 255) << 8) + (((c) & 255) << 16) + (((d) & 255) << 24))
#define CV_BGR2GRAY 6
#define CV_RGB2GRAY 7
#define CV_BGR2HSV 40
#define CV_RGB2HSV 41
#define CV_HAAR_DO_CANNY_PRUNING 1
#define CV_PI 3.1415926535897932384626433832795
#define IPL_DEPTH_8U 8
#define IPL_DEPTH_32F 32
#define IPL_ORIGIN_TL 0
#define CV_LU 0
#define CV_SVD 1
#define CV_SVD_SYM 2

/* ---- RNG (types_c.h: CV_RNG_COEFF) ---- */
static inline CvRNG cvRNG(int64_t seed = -1) { CvRNG r = seed ? (uint64_t)seed : (uint64_t)(int64_t)-1; return r; }
/* Stream multiplexing for the multi-camera harness: the reference draws everything from ONE global stream (Public.h:156),
 * which serialises cameras and objects.  The B200 path gives every (camera, object) pair its own stream so that cameras
 * can run on different GPUs; ref_harness.cpp points this override at the stream of the (camera, object) whose reference
 * method it is about to call.  NULL (default) = the reference's own global stream, untouched. */
#ifdef __cplusplus
extern "C" {
#endif
extern uint64_t* mct_shim_rng_override;
#ifdef __cplusplus
}
#endif
static inline unsigned cvRandInt(CvRNG* rng) {
    CvRNG* r = mct_shim_rng_override ? (CvRNG*)mct_shim_rng_override : rng;
    uint64_t t = *r;
    t = (uint64_t)(unsigned)t * 4164903690U + (t >> 32);
    *r = t;
    return (unsigned)t;
}
static inline double cvRandReal(CvRNG* rng) { return cvRandInt(rng) * 2.3283064365386962890625e-10; }
static inline int cvRound(double v) { return _mm_cvtsd_si32(_mm_set_sd(v)); }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

/* ---- IplImage ---- */
typedef struct _IplImage {
    int nSize, ID, nChannels, alphaChannel, depth;
    char colorModel[4], channelSeq[4];
    int dataOrder, origin, align, width, height;
    void *roi, *maskROI, *imageId, *tileInfo;
    int imageSize;
    char* imageData;
    int widthStep;
    int BorderMode[4], BorderConst[4];
    char* imageDataOrigin;
} IplImage;

IplImage* cvCreateImage(CvSize size, int depth, int channels);
IplImage* cvCreateImageHeader(CvSize size, int depth, int channels);
IplImage* cvInitImageHeader(IplImage* img, CvSize size, int depth, int channels, int origin = 0, int align = 4);
void cvCreateData(CvArr* arr);
void cvReleaseImage(IplImage** img);
void cvReleaseImageHeader(IplImage** img);
void cvCopy(const CvArr* src, CvArr* dst, const CvArr* mask = 0);
void cvCvtColor(const CvArr* src, CvArr* dst, int code);          /* IplImage 8u: BGR2GRAY, BGR2HSV */
void cvEqualizeHist(const CvArr*, CvArr*);

/* ---- CvMat ---- */
#define CV_32F 5
#define CV_64F 6
#define CV_8U 0
#define CV_32FC1 5
#define CV_64FC1 6
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> 3) & 63) + 1)
typedef struct CvMat {
    int type, step;
    int* refcount; int hdr_refcount;
    union { uchar* ptr; short* s; int* i; float* fl; double* db; } data;
    int rows, cols;
} CvMat;
CvMat* cvCreateMat(int rows, int cols, int type);
void cvReleaseMat(CvMat** m);
CvMat* cvCloneMat(const CvMat* m);
double cvDet(const CvArr* m);
double cvInvert(const CvArr* src, CvArr* dst, int method = CV_LU);
#define cvInv cvInvert
static inline double cvmGet(const CvMat* m, int r, int c) {
    return CV_MAT_DEPTH(m->type) == CV_32F ? ((const float*)(m->data.ptr + (size_t)m->step * r))[c]
                                           : ((const double*)(m->data.ptr + (size_t)m->step * r))[c];
}
static inline void cvmSet(CvMat* m, int r, int c, double v) {
    if (CV_MAT_DEPTH(m->type) == CV_32F) ((float*)(m->data.ptr + (size_t)m->step * r))[c] = (float)v;
    else ((double*)(m->data.ptr + (size_t)m->step * r))[c] = v;
}


CvMat cvMat(int rows, int cols, int type, void* data = 0);
void cvGEMM(const CvArr* A, const CvArr* B, double alpha, const CvArr* C, double beta, CvArr* D, int tABC = 0);
#define cvMatMulAdd(A, B, C, D) cvGEMM((A), (B), 1., (C), 1., (D), 0)
#define cvMatMul(A, B, D) cvMatMulAdd((A), (B), NULL, (D))
void cvAdd(const CvArr* a, const CvArr* b, CvArr* dst, const CvArr* mask = 0);
void cvSub(const CvArr* a, const CvArr* b, CvArr* dst, const CvArr* mask = 0);
void cvTranspose(const CvArr* src, CvArr* dst);
void cvSetIdentity(CvArr* mat, CvScalar value = cvRealScalar(1));
void cvSetZero(CvArr* mat);
#define cvZero cvSetZero
void cvSVD(CvArr* A, CvArr* W, CvArr* U = 0, CvArr* V = 0, int flags = 0);

/* ---- Kalman filter (C API of the video module) ---- */
typedef struct CvKalman {
    int MP, DP, CP;
    CvMat *state_pre, *state_post, *transition_matrix, *control_matrix, *measurement_matrix, *process_noise_cov,
          *measurement_noise_cov, *error_cov_pre, *gain, *error_cov_post, *temp1, *temp2, *temp3, *temp4, *temp5;
} CvKalman;
CvKalman* cvCreateKalman(int dynam_params, int measure_params, int control_params = 0);
void cvReleaseKalman(CvKalman** kalman);
const CvMat* cvKalmanPredict(CvKalman* kalman, const CvMat* control = 0);
const CvMat* cvKalmanCorrect(CvKalman* kalman, const CvMat* measurement);

/* ---- EM (ml.h): only referenced by the geometric fuser's GMM modes, which CameraNetwork.cpp:398 never selects ---- */
#define CV_TERMCRIT_ITER 1
#define CV_TERMCRIT_EPS 2
typedef struct CvTermCriteria { int type, max_iter; double epsilon; } CvTermCriteria;
#ifdef __cplusplus
struct CvEMParams {
    int nclusters, cov_mat_type, start_step;
    const CvMat *probs, *weights, *means; const CvMat** covs;
    CvTermCriteria term_crit;
    CvEMParams() : nclusters(10), cov_mat_type(1), start_step(0), probs(0), weights(0), means(0), covs(0) { term_crit.type = 3; term_crit.max_iter = 100; term_crit.epsilon = 1e-6; }
};
class CvEM {
public:
    enum { COV_MAT_SPHERICAL = 0, COV_MAT_DIAGONAL = 1, COV_MAT_GENERIC = 2 };
    enum { START_E_STEP = 1, START_M_STEP = 2, START_AUTO_STEP = 0 };
    CvEM() {}
    bool train(const CvMat*, const CvMat* = 0, CvEMParams = CvEMParams(), CvMat* = 0) { mct_shim_not_on_path("CvEM::train"); return false; }
    const CvMat* get_means() const { return 0; }
    const CvMat** get_covs() const { return 0; }
    const CvMat* get_weights() const { return 0; }
    const CvMat* get_probs() const { return 0; }
    int get_nclusters() const { return 0; }
};
#endif

/* ---- display / IO / detector: compile-only ---- */
static inline int cvWaitKey(int = 0) { return -1; }
static inline int cvNamedWindow(const char*, int = 1) { return 0; }
static inline void cvShowImage(const char*, const CvArr*) {}
static inline void cvResizeWindow(const char*, int, int) {}
static inline void cvDestroyAllWindows() {}
static inline void cvDestroyWindow(const char*) {}
static inline void cvInitFont(CvFont*, int, double, double, double = 0, int = 1, int = 8) {}
static inline void cvPutText(CvArr*, const char*, CvPoint, const CvFont*, CvScalar) {}
static inline void cvLine(CvArr*, CvPoint, CvPoint, CvScalar, int = 1, int = 8, int = 0) {}
static inline void cvDrawRect(CvArr*, CvPoint, CvPoint, CvScalar, int = 1, int = 8, int = 0) {}
static inline void cvRectangle(CvArr*, CvPoint, CvPoint, CvScalar, int = 1, int = 8, int = 0) {}
static inline void cvCircle(CvArr*, CvPoint, int, CvScalar, int = 1, int = 8, int = 0) {}
static inline void cvEllipse(CvArr*, CvPoint, CvSize, double, double, double, CvScalar, int = 1, int = 8, int = 0) {}
static inline IplImage* cvLoadImage(const char*, int = 1) { mct_shim_not_on_path("cvLoadImage"); return 0; }
static inline int cvSaveImage(const char*, const CvArr*, const int* = 0) { return 0; }
static inline CvCapture* cvCaptureFromFile(const char*) { mct_shim_not_on_path("cvCaptureFromFile"); return 0; }
static inline CvCapture* cvCaptureFromCAM(int) { mct_shim_not_on_path("cvCaptureFromCAM"); return 0; }
static inline CvCapture* cvCreateCameraCapture(int) { mct_shim_not_on_path("cvCreateCameraCapture"); return 0; }
static inline IplImage* cvQueryFrame(CvCapture*) { mct_shim_not_on_path("cvQueryFrame"); return 0; }
static inline void cvReleaseCapture(CvCapture**) {}
static inline CvVideoWriter* cvCreateVideoWriter(const char*, int, double, CvSize, int = 1) { return 0; }
static inline int cvWriteFrame(CvVideoWriter*, const IplImage*) { return 0; }
static inline void cvReleaseVideoWriter(CvVideoWriter**) {}
static inline CvMemStorage* cvCreateMemStorage(int = 0) { return 0; }
static inline void cvClearMemStorage(CvMemStorage*) {}
static inline void* cvLoad(const char*, CvMemStorage* = 0, const char* = 0, const char** = 0) { return 0; }
static inline CvSeq* cvHaarDetectObjects(const CvArr*, CvHaarClassifierCascade*, CvMemStorage*, double = 1.1, int = 3,
                                         int = 0, CvSize = cvSize(0, 0), CvSize = cvSize(0, 0)) { return 0; }
static inline char* cvGetSeqElem(const CvSeq*, int) { return 0; }

#ifdef __cplusplus
/* CameraNetwork.cpp:111 streams a std::ofstream into a stream; C++03 did that through the stream's void* conversion */
#include <fstream>
namespace std { inline ostream& operator<<(ostream& os, const ofstream& f) { return os << (const void*)(f.fail() ? 0 : &f); } }
/* ---- the little of the C++ API the path touches ---- */
void mct_shim_fill_noise(float* dst, int n, int kind, double a, double b);   /* kind 0 = randn(mean a, sd b), 1 = randu[a,b) */
namespace cv {
typedef std::string String;
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
                double operator[](int i) const { return val[i]; } };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int X, int Y, int W, int H) : x(X), y(Y), width(W), height(H) {} };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
template <class T, int N> struct Vec { T val[N]; T& operator[](int i) { return val[i]; } const T& operator[](int i) const { return val[i]; } };
typedef Vec<uchar, 3> Vec3b;
/* dense 2-D matrix: owning (shared buffer) or a view into an IplImage / a parent (ROI) */
struct Mat {
    int rows, cols, flags;          /* flags = type */
    size_t step;
    uchar* data;
    std::shared_ptr<std::vector<uchar> > buf;
    Mat() : rows(0), cols(0), flags(0), step(0), data(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), flags(0), step(0), data(0) { create(r, c, type); }
    Mat(const IplImage* img, bool = false)
        : rows(img->height), cols(img->width), flags(img->nChannels == 3 ? CV_8UC3 : CV_8UC1), step(img->widthStep), data((uchar*)img->imageData) {}
    static int esz(int type) { static const int d[8] = {1, 1, 2, 2, 4, 4, 8, 0}; return d[CV_MAT_DEPTH(type)] * CV_MAT_CN(type); }
    void create(int r, int c, int type) {
        rows = r; cols = c; flags = type; step = (size_t)c * esz(type);
        buf.reset(new std::vector<uchar>(step * (size_t)r + 16)); data = buf->data();
    }
    int type() const { return flags; }
    int channels() const { return CV_MAT_CN(flags); }
    bool empty() const { return data == 0 || rows * cols == 0; }
    Mat operator()(const Rect& r) const { Mat m(*this); m.rows = r.height; m.cols = r.width; m.data = data + step * r.y + (size_t)r.x * esz(flags); return m; }
    template <class T> T& at(int r, int c) { return ((T*)(data + step * r))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + step * r))[c]; }
    template <class T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + step * i); }
    template <class T> T* ptr(int r = 0) { return (T*)(data + step * r); }
    operator IplImage() const {
        IplImage im; memset(&im, 0, sizeof(im)); im.nSize = sizeof(im); im.nChannels = channels(); im.depth = IPL_DEPTH_8U;
        im.width = cols; im.height = rows; im.widthStep = (int)step; im.imageData = (char*)data; im.imageSize = (int)(step * rows);
        return im;
    }
};
static inline void randn(Mat& m, const Scalar& mean, const Scalar& sd) { mct_shim_fill_noise((float*)m.data, m.rows * m.cols, 0, mean[0], sd[0]); }
static inline void randu(Mat& m, const Scalar& lo, const Scalar& hi) { mct_shim_fill_noise((float*)m.data, m.rows * m.cols, 1, lo[0], hi[0]); }
void cvtColor(const Mat& src, Mat& dst, int code);               /* 8UC3 BGR2HSV / BGR2GRAY / RGB2GRAY */
static inline void imshow(const std::string&, const Mat&) {}
struct VideoCapture {
    VideoCapture() {}
    VideoCapture(const std::string&) { mct_shim_not_on_path("cv::VideoCapture"); }
    bool isOpened() const { return false; }
    bool grab() { return false; }
    bool retrieve(Mat&, int = 0) { return false; }
    void release() {}
};
}  // namespace cv
#endif
#endif
