/*
 * mct_oracle.c -- CPU ORACLE (test infrastructure only; see mct_oracle.h header comment).
 *
 * Plain-C restatement of the reference's appearance-scoring hot path.  Each function cites the
 * reference file:line (relative to /root/reference/src) whose arithmetic it follows, including
 * the float/double promotion of every intermediate (C++11 overload rules: pow(float,int) is
 * double; exp/log of a float argument are the float overloads).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off so no FMA contraction changes f32 results).
 * PARITY: pinned row by row to the reference's own code (oracle/_ref, tests/test_ref_pins_oracle.py) -- see header.
 */
#include "mct_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* RNG: Public.h:156 (cvRNG(1)), Public.cpp:10-23.  OpenCV cvRNG = multiply-with-carry,        */
/* cvRandInt: state = (uint32)state*4164903690 + (state>>32); cvRandReal = uint32 * 2^-32.     */
/* ------------------------------------------------------------------------------------------ */
void orc_rng_seed(orc_rng* r, int64_t seed) { r->state = seed ? (uint64_t)seed : (uint64_t)(int64_t)-1; }
uint32_t orc_rng_next(orc_rng* r)
{
    uint64_t t = r->state;
    t = (uint64_t)(uint32_t)t * 4164903690U + (t >> 32);
    r->state = t;
    return (uint32_t)t;
}
/* Public.cpp:15-18: cvRandInt(&rng) % (max-min+1) + min  (unsigned arithmetic) */
int orc_randint(orc_rng* r, int mn, int mx) { return (int)(orc_rng_next(r) % (unsigned)(mx - mn + 1) + (unsigned)mn); }
/* Public.cpp:20-23 */
float orc_randfloat(orc_rng* r) { return (float)(orc_rng_next(r) * 2.3283064365386962890625e-10); }
/* cvRound(double): SSE2 cvtsd2si = round-half-to-even in the default rounding mode */
int orc_cvround(double v) { return (int)lrint(v); }

/* ------------------------------------------------------------------------------------------ */
/* Integral image: Matrix.cpp:33-47 -> ippiIntegral_8u32f_C1R(src, dst (W+1)x(H+1) f32, val=0) */
/* IPP's f32 accumulation order is unobtainable; restated as exact u32 prefix sums rounded     */
/* once to f32 ("hard parts").                                               */
/* ------------------------------------------------------------------------------------------ */
void orc_integral(const uint8_t* img, int W, int H, int pitch, float* ii)
{
    int W1 = W + 1;
    uint32_t* acc = (uint32_t*)calloc((size_t)W1, sizeof(uint32_t)); /* previous row of exact sums */
    for (int x = 0; x < W1; x++) ii[x] = 0.0f;
    for (int y = 0; y < H; y++) {
        uint32_t run = 0;
        float* out = ii + (size_t)(y + 1) * W1;
        out[0] = 0.0f;
        for (int x = 0; x < W; x++) {
            run += img[(size_t)y * pitch + x];
            acc[x + 1] += run;
            out[x + 1] = (float)acc[x + 1];
        }
    }
    free(acc);
}

/* Matrix.cpp:49-67: return br + tl - tr - bl  (f32, left to right) */
float orc_sum_rect(const float* ii, int W, int x, int y, int w, int h)
{
    int W1 = W + 1;
    int maxy = (y + h) * W1, maxx = x + w, yy = y * W1;
    float tl = ii[yy + x], tr = ii[yy + maxx], br = ii[maxy + maxx], bl = ii[maxy + x];
    float s = br + tl;
    s = s - tr;
    s = s - bl;
    return s;
}

/* ------------------------------------------------------------------------------------------ */
/* Frame pre-processing (Camera.cpp:449-494; Matrix.cpp:70-100, 518-580)                        */
/* ------------------------------------------------------------------------------------------ */
static int orc_sat_round_div(double num, double den) { return (int)lrint(num / den); }   /* saturate_cast<int>(double) = cvRound */
void orc_bgr_to_planes(const uint8_t* bgr, int W, int H, int pitch_bytes, int use_hsv,
                       uint8_t* gray, uint8_t* c0, uint8_t* c1, uint8_t* c2)
{
    enum { HSV_SHIFT = 12 };
    static int sdiv[256], hdiv[256], init = 0;
    if (!init) {
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; i++) {
            sdiv[i] = orc_sat_round_div((double)(255 << HSV_SHIFT), 1.0 * i);
            hdiv[i] = orc_sat_round_div((double)(180 << HSV_SHIFT), 6.0 * i);
        }
        init = 1;
    }
    for (int y = 0; y < H; y++) {
        const uint8_t* p = bgr + (size_t)y * pitch_bytes;
        for (int x = 0; x < W; x++) {
            const int b = p[3 * x], g = p[3 * x + 1], r = p[3 * x + 2];
            const size_t o = (size_t)y * W + x;
            gray[o] = (uint8_t)((r * 4899 + g * 9617 + b * 1868 + 8192) >> 14);
            if (!use_hsv) { c0[o] = (uint8_t)r; c1[o] = (uint8_t)g; c2[o] = (uint8_t)b; continue; }
            int v = b > g ? b : g; if (r > v) v = r;
            int vmin = b < g ? b : g; if (r < vmin) vmin = r;
            const int diff = v - vmin;
            const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
            const int s = (diff * sdiv[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT;
            int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
            h = (h * hdiv[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT;
            h += h < 0 ? 180 : 0;
            c0[o] = (uint8_t)h; c1[o] = (uint8_t)s; c2[o] = (uint8_t)v;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Haar features                                                                               */
/* ------------------------------------------------------------------------------------------ */
/* HaarFeature.cpp:25-69 (draw order), HaarFeatureVector.cpp:24-29 (feature loop) */
orc_haar_table* orc_haar_generate(orc_rng* r, int F, int w0, int h0)
{
    orc_haar_table* t = (orc_haar_table*)calloc(1, sizeof(*t));
    t->F = F;
    t->nrect = (int*)calloc((size_t)F, sizeof(int));
    t->rect = (int*)calloc((size_t)F * ORC_MAX_RECTS * 4, sizeof(int));
    t->weight = (float*)calloc((size_t)F * ORC_MAX_RECTS, sizeof(float));
    t->channel = (int*)calloc((size_t)F, sizeof(int));
    for (int f = 0; f < F; f++) {
        int n = orc_randint(r, 2, 6); /* DefaultParameters.h:14-15 */
        t->nrect[f] = n;
        for (int k = 0; k < n; k++) {
            int* rc = t->rect + ((size_t)f * ORC_MAX_RECTS + k) * 4;
            t->weight[f * ORC_MAX_RECTS + k] = orc_randfloat(r) * 2 - 1;
            rc[0] = orc_randint(r, 0, w0 - 3);
            rc[1] = orc_randint(r, 0, h0 - 3);
            rc[2] = orc_randint(r, 1, w0 - rc[0] - 2);
            rc[3] = orc_randint(r, 1, h0 - rc[1] - 2);
        }
        /* HaarFeature.cpp:68: m_useChannels[randint(0, nch-1)] with nch = 1 -> one draw, channel 0 */
        (void)orc_randint(r, 0, 0);
        t->channel[f] = 0;
    }
    return t;
}
void orc_haar_free(orc_haar_table* t)
{
    if (!t) return;
    free(t->nrect); free(t->rect); free(t->weight); free(t->channel); free(t);
}
static orc_haar_table* haar_clone(const orc_haar_table* s)
{
    orc_haar_table* t = (orc_haar_table*)calloc(1, sizeof(*t));
    int F = s->F;
    t->F = F;
    t->nrect = (int*)malloc((size_t)F * sizeof(int)); memcpy(t->nrect, s->nrect, (size_t)F * sizeof(int));
    t->rect = (int*)malloc((size_t)F * 24 * sizeof(int)); memcpy(t->rect, s->rect, (size_t)F * 24 * sizeof(int));
    t->weight = (float*)malloc((size_t)F * 6 * sizeof(float)); memcpy(t->weight, s->weight, (size_t)F * 6 * sizeof(float));
    t->channel = (int*)malloc((size_t)F * sizeof(int)); memcpy(t->channel, s->channel, (size_t)F * sizeof(int));
    return t;
}

/* HaarFeature.cpp:111-141 */
static float haar_one(const orc_haar_table* t, int f, const float* ii, int W, int col, int row, float sx, float sy)
{
    float sum = 0.0f;
    for (int k = 0; k < t->nrect[f]; k++) {
        const int* rc = t->rect + ((size_t)f * ORC_MAX_RECTS + k) * 4;
        int x = orc_cvround((double)((float)rc[0] * sx)) + col;
        int y = orc_cvround((double)((float)rc[1] * sy)) + row;
        int h = orc_cvround((double)((float)rc[3] * sy));
        int w = orc_cvround((double)((float)rc[2] * sx));
        float p = t->weight[f * ORC_MAX_RECTS + k] * orc_sum_rect(ii, W, x, y, w, h);
        sum = sum + p;
    }
    float d = sx * sy;
    return sum / d;
}

static int g_threads = 1;
void orc_clf_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* HaarFeatureVector.cpp:48-79: `#pragma omp parallel for` over features, inner loop over samples */
void orc_haar_compute(const orc_haar_table* t, const float* ii, int W, const orc_boxes* b, float* out, int ld)
{
#ifdef _OPENMP
#pragma omp parallel for num_threads(g_threads) schedule(static)
#endif
    for (int f = 0; f < t->F; f++)
        for (int i = 0; i < b->n; i++)
            out[(size_t)f * ld + i] = haar_one(t, f, ii, W, b->col[i], b->row[i], b->sx[i], b->sy[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* MultiDimensionalColorHistogram.cpp:32-127 (COLOR_NUM_PARTS = 1, FeatureParameters.h:5;      */
/* m_shouldWeightFromCenter = true, MultiDimensionalColorHistogram.h).  FV wrapper:            */
/* MultiDimensionalColorHistogramFeatureVector.cpp:52-90 (omp parallel for over samples).      */
/* ------------------------------------------------------------------------------------------ */
static void mdch_one(const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, int pitch, int nb,
                     int w0, int h0, int col, int row, float sx, float sy, float* hist /* nb^3 */)
{
    int dim = nb * nb * nb;
    for (int i = 0; i < dim; i++) hist[i] = 0.0f;
    float numberOfRows = (float)orc_cvround((double)((float)h0 * sy));
    int partEnd = orc_cvround((double)(numberOfRows * 100 / 100));     /* :47-48, accum = 100 */
    float numberOfColumns = (float)orc_cvround((double)((float)w0 * sx));
    float varianceX = (float)pow((double)(numberOfColumns / 2), 2.0);  /* :68 pow(float,int) -> double */
    float sampleCenterX = (float)col + numberOfColumns / 2;            /* :69 */
    float binWidth = (float)(256 / nb);                                /* :74 integer division */
    float partNumberOfRows = (float)(partEnd - 0);                     /* :87 */
    float varianceY = (float)pow((double)(partNumberOfRows / 2), 2.0);
    float sampleCenterY = (float)(row + 0) + partNumberOfRows / 2;
    for (unsigned rowIndex = (unsigned)row; rowIndex < (unsigned)(row + partEnd); rowIndex++) {
        for (unsigned columnIndex = (unsigned)col; (float)columnIndex < ((float)col + numberOfColumns); columnIndex++) {
            unsigned rPixel = c0[(size_t)rowIndex * pitch + columnIndex];
            unsigned gPixel = c1[(size_t)rowIndex * pitch + columnIndex];
            unsigned bPixel = c2[(size_t)rowIndex * pitch + columnIndex];
            unsigned rBin = (unsigned)floorf((float)rPixel / binWidth);
            unsigned gBin = (unsigned)floorf((float)gPixel / binWidth);
            unsigned bBin = (unsigned)floorf((float)bPixel / binWidth);
            unsigned featureIndex = rBin + nb * gBin + nb * nb * bBin;
            /* :107-108: pow(float,int) -> double; double / float -> double; sum stored to float */
            double dx = (double)(sampleCenterX - (float)columnIndex);
            double dy = (double)(sampleCenterY - (float)rowIndex);
            float weightedDistanceFromCenter =
                (float)(dx * dx / (double)(2.0f * varianceX) + dy * dy / (double)(2.0f * varianceY));
            float featureWeight = expf(-1 * weightedDistanceFromCenter);  /* :110 exp(float) */
            hist[featureIndex] = hist[featureIndex] + featureWeight;
        }
    }
    /* :118 ::normalizeVec (Public.h:270-275): f32 sequential sum then divide */
    float sum = 0;
    for (int i = 0; i < dim; i++) sum += hist[i];
    for (int i = 0; i < dim; i++) hist[i] /= sum;
}

void orc_mdch_compute(const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, int pitch,
                      int nb, int w0, int h0, const orc_boxes* b, float* out, int ld)
{
    int dim = nb * nb * nb;
#ifdef _OPENMP
#pragma omp parallel for num_threads(g_threads) schedule(static)
#endif
    for (int i = 0; i < b->n; i++) {
        float* hist = (float*)malloc((size_t)dim * sizeof(float));
        mdch_one(c0, c1, c2, pitch, nb, w0, h0, b->col[i], b->row[i], b->sx[i], b->sy[i], hist);
        for (int k = 0; k < dim; k++) out[(size_t)k * ld + i] = hist[k];
        free(hist);
    }
}

/* CultureColorHistogram.cpp:25-111.  mode 0 ("reference_exact"): the distance is always taken to
   HC[0] (:88-90) with a strict '>' so every pixel lands in bin 0; CC_Histogram[0] is never
   initialised (:72-76) -- treated as 0.  mode 1 ("intended"): nearest hue centroid, first minimum. */
void orc_cch_compute(const uint8_t* hue, int pitch, int w0, int h0, const orc_boxes* b, int mode, float* out, int ld)
{
    static const int HC[11] = {5, 206, 181, 160, 81, 46, 28, 20, 110, 121, 133};
    for (int i = 0; i < b->n; i++) {
        unsigned sh = (unsigned)orc_cvround((double)((float)h0 * b->sy[i]));
        unsigned sw = (unsigned)orc_cvround((double)((float)w0 * b->sx[i]));
        int partEnd = (int)(sh * 100 / 100);
        int hist[11] = {0};
        for (int r = 0; r < partEnd; r++)
            for (unsigned c = 0; c < sw; c++) {
                int H = hue[(size_t)(r + b->row[i]) * pitch + (c + b->col[i])];
                int minimumDist = 999999, indicator = 0;
                for (int k = 0; k < 11; k++) {
                    int ref = mode == 0 ? HC[0] : HC[k];
                    int d = (int)sqrt((double)(H - ref) * (double)(H - ref));
                    if (minimumDist > d) { minimumDist = d; indicator = k; }
                }
                hist[indicator]++;
            }
        for (int k = 0; k < 11; k++) out[(size_t)k * ld + i] = (float)hist[k];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Strong classifier                                                                           */
/* ------------------------------------------------------------------------------------------ */
struct orc_clf {
    int feature_type, n_haar, nb, w0, h0, weak_type, strong_type, K, F;
    float lr;
    orc_haar_table* haar;
    float *mu0, *mu1, *sig0, *sig1, *n0, *n1, *e0, *e1;   /* stumps */
    float *pw0, *pw1;                                     /* perceptron m_weightList[0], [1] */
    int* trained;
    int* sel; int n_sel;
    int cch_mode;
    float *cTP, *cFP, *cTN, *cFN;                         /* AdaBoost: [K][F] running counts, initial value 1 (AdaBoostClassifier.cpp:27-36) */
    float* alpha; float sum_alpha;                        /* [K] */
    float pct_retained;                                   /* MILEnsemble: m_percentageOfRetainedWeakClassifiers (= pct / 100.0f, ClassifierParameters.h:52) */
};

/* StrongClassifierBase.cpp:8-129 (K clamp :20-23; weak clf ctor: lr 0.85 then SetLearningRate) */
orc_clf* orc_clf_create(int feature_type, int n_haar, int nb, int w0, int h0, int weak_type,
                        int strong_type, int K, float lr, const orc_haar_table* haar)
{
    orc_clf* c = (orc_clf*)calloc(1, sizeof(*c));
    c->feature_type = feature_type; c->nb = nb; c->w0 = w0; c->h0 = h0;
    c->weak_type = weak_type; c->strong_type = strong_type; c->lr = lr;
    int ncol = nb * nb * nb;
    switch (feature_type) {
        case ORC_FEAT_HAAR: c->n_haar = n_haar; c->F = n_haar; break;
        case ORC_FEAT_CCH: c->n_haar = 0; c->F = 11; break;
        case ORC_FEAT_MDCH: c->n_haar = 0; c->F = ncol; break;
        default: c->n_haar = n_haar; c->F = n_haar + ncol; break;
    }
    if (K > c->F || K < 1) K = c->F / 2;
    c->K = K;
    if (c->n_haar > 0 && haar) c->haar = haar_clone(haar);
    int F = c->F;
    c->mu0 = (float*)calloc(F, 4); c->mu1 = (float*)calloc(F, 4);
    c->sig0 = (float*)calloc(F, 4); c->sig1 = (float*)calloc(F, 4);
    c->n0 = (float*)calloc(F, 4); c->n1 = (float*)calloc(F, 4);
    c->e0 = (float*)calloc(F, 4); c->e1 = (float*)calloc(F, 4);
    c->pw0 = (float*)calloc(F, 4); c->pw1 = (float*)calloc(F, 4);
    c->trained = (int*)calloc(F, sizeof(int));
    c->sel = (int*)calloc(F, sizeof(int));
    if (strong_type == ORC_STRONG_ADABOOST) {
        size_t n = (size_t)c->K * F;
        c->cTP = (float*)malloc(n * 4); c->cFP = (float*)malloc(n * 4); c->cTN = (float*)malloc(n * 4); c->cFN = (float*)malloc(n * 4);
        for (size_t i = 0; i < n; i++) { c->cTP[i] = 1.0f; c->cFP[i] = 1.0f; c->cTN[i] = 1.0f; c->cFN[i] = 1.0f; }
        c->alpha = (float*)calloc((size_t)c->K, 4);
    }
    for (int f = 0; f < F; f++) { c->sig0[f] = 1; c->sig1[f] = 1; }  /* OnlineStumps::Initialize :53-60 */
    /* StrongClassifierBase.cpp:91: m_selectorList.resize(K, 0) -- the list starts as K ZEROS, not empty.  MILBoost / MILAnyBoost /
       AdaBoost clear it at the top of Update; MILEnsemble does not: its first RetainBestPerformingWeakClassifiers sees K copies
       of feature 0, retains it once (untrained), and MILEnsemble::Update then never trains feature 0 while it stays retained
       (pinned against the reference build, tests/test_ref_pins_oracle.py).  Perceptron::Initialize leaves the weights (0, 0). */
    c->n_sel = strong_type == ORC_STRONG_MILENSEMBLE ? c->K : 0;
    return c;
}
void orc_clf_free(orc_clf* c)
{
    if (!c) return;
    orc_haar_free(c->haar);
    free(c->mu0); free(c->mu1); free(c->sig0); free(c->sig1); free(c->n0); free(c->n1); free(c->e0); free(c->e1);
    free(c->pw0); free(c->pw1); free(c->trained); free(c->sel);
    free(c->cTP); free(c->cFP); free(c->cTN); free(c->cFN); free(c->alpha); free(c);
}
int orc_clf_num_features(const orc_clf* c) { return c->F; }
/* MILEnsembleClassifierParameters (ClassifierParameters.h:105-117): the percentage is stored divided by 100.0f (:52) */
void orc_clf_set_retained(orc_clf* c, float percentage) { c->pct_retained = percentage / 100.0f; }

/* FeatureVector::Compute by type (HaarAndColorHistogramFeatureVector.cpp:33-44: Haar rows then MDCH rows) */
void orc_clf_features(const orc_clf* c, const orc_frame* f, const orc_boxes* b, float* out, int ld)
{
    if (b->n == 0) return;
    if (c->n_haar > 0) orc_haar_compute(c->haar, f->ii, f->W, b, out, ld);
    if (c->feature_type == ORC_FEAT_MDCH || c->feature_type == ORC_FEAT_HAAR_MDCH)
        orc_mdch_compute(f->c0, f->c1, f->c2, f->pitch, c->nb, c->w0, c->h0, b, out + (size_t)c->n_haar * ld, ld);
    if (c->feature_type == ORC_FEAT_CCH)
        orc_cch_compute(f->c0, f->pitch, c->w0, c->h0, b, c->cch_mode, out, ld);
}

/* Matrix.h:674-683 Mean: ippiMean_32f (f64 accumulate) -> (float) */
static float mean_f(const float* x, int n)
{
    double s = 0; for (int i = 0; i < n; i++) s += x[i];
    return (float)(s / n);
}
/* Matrix.h:685-694 Var: ippiMean_StdDev_32f (population), returns (float)(std*std) */
static float var_f(const float* x, int n)
{
    double s = 0; for (int i = 0; i < n; i++) s += x[i];
    double m = s / n, q = 0;
    for (int i = 0; i < n; i++) { double d = x[i] - m; q += d * d; }
    double sd = sqrt(q / n);
    return (float)(sd * sd);
}
/* (X - mu).Sqr().Mean(): Matrix.h:520-527 (AddC with -mu, f32), :558 Sqr f32, Mean f64-accumulate */
static float mean_sqdev_f(const float* x, int n, float mu)
{
    double s = 0;
    float nm = mu * -1;
    for (int i = 0; i < n; i++) { float d = x[i] + nm; float q = d * d; s += q; }
    return (float)(s / n);
}

/* OnlineStumpsWeakClassifier.cpp:99-151 */
static void stump_update(orc_clf* c, int f, const float* xp, int P, const float* xn, int Nn)
{
    float lr = c->lr;
    float pm = 0.0f, nm = 0.0f;
    if (P > 0) pm = mean_f(xp, P);
    if (Nn > 0) nm = mean_f(xn, Nn);
    if (c->trained[f]) {
        if (P > 0) {
            c->mu1[f] = (lr * c->mu1[f] + (1 - lr) * pm);
            c->sig1[f] = (lr * c->sig1[f] + (1 - lr) * mean_sqdev_f(xp, P, c->mu1[f]));
        }
        if (Nn > 0) {
            c->mu0[f] = (lr * c->mu0[f] + (1 - lr) * nm);
            c->sig0[f] = (lr * c->sig0[f] + (1 - lr) * mean_sqdev_f(xn, Nn, c->mu0[f]));
        }
    } else {
        c->trained[f] = 1;
        if (P > 0) { c->mu1[f] = pm; c->sig1[f] = var_f(xp, P) + 1e-9f; }
        if (Nn > 0) { c->mu0[f] = nm; c->sig0[f] = var_f(xn, Nn) + 1e-9f; }
    }
    c->n0[f] = 1.0f / powf(c->sig0[f], 0.5f);
    c->n1[f] = 1.0f / powf(c->sig1[f], 0.5f);
    c->e1[f] = -1.0f / (2.0f * c->sig1[f] + 0.0f /* 1e-99f underflows to 0 */);
    c->e0[f] = -1.0f / (2.0f * c->sig0[f] + 0.0f);
}

/* WeightedStumpsWeakClassifier.cpp:50-118; Matrix.h:540-544 normalize, :696-719 VarW/MeanW.
   NULL weight lists dereference in the reference (:71); restated as uniform 1.0. */
static float meanw_f(const float* x, const float* wn, int n)
{
    double s = 0; for (int i = 0; i < n; i++) { float p = x[i] * wn[i]; s += p; }
    return (float)s;
}
static float varw_f(const float* x, const float* wn, int n, float mu)
{
    double s = 0;
    float nm = mu * -1;
    for (int i = 0; i < n; i++) { float d = x[i] + nm; float q = d * d; float p = q * wn[i]; s += p; }
    return (float)s;
}
static void normalize_w(const float* w, int n, float* wn)
{
    double s = 0;
    for (int i = 0; i < n; i++) s += w ? w[i] : 1.0f;
    float k = (float)(1.0 / (s + 1e-6));
    for (int i = 0; i < n; i++) wn[i] = (w ? w[i] : 1.0f) * k;
}
static void wstump_update(orc_clf* c, int f, const float* xp, int P, const float* xn, int Nn,
                          const float* pwn, const float* nwn /* already normalised */)
{
    float lr = c->lr;
    float posmu = 0.0f, negmu = 0.0f;
    if (P > 0) posmu = meanw_f(xp, pwn, P);
    if (Nn > 0) negmu = meanw_f(xn, nwn, Nn);
    if (c->trained[f]) {
        if (P > 0) {
            c->mu1[f] = (lr * c->mu1[f] + (1 - lr) * posmu);
            c->sig1[f] = (lr * c->sig1[f] + (1 - lr) * varw_f(xp, pwn, P, c->mu1[f]));
        }
        if (Nn > 0) {
            c->mu0[f] = (lr * c->mu0[f] + (1 - lr) * negmu);
            c->sig0[f] = (lr * c->sig0[f] + (1 - lr) * varw_f(xn, nwn, Nn, c->mu0[f]));
        }
    } else {
        c->trained[f] = 1;
        c->mu1[f] = posmu; c->mu0[f] = negmu;
        if (Nn > 0) c->sig0[f] = varw_f(xn, nwn, Nn, negmu) + 1e-9f;
        if (P > 0) c->sig1[f] = varw_f(xp, pwn, P, posmu) + 1e-9f;
    }
    c->n0[f] = 1.0f / powf(c->sig0[f], 0.5f);
    c->n1[f] = 1.0f / powf(c->sig1[f], 0.5f);
    c->e1[f] = -1.0f / (2.0f * c->sig1[f]);
    c->e0[f] = -1.0f / (2.0f * c->sig0[f]);
}

/* PerceptronWeakClassifier.cpp:98-156 */
static void perceptron_update(orc_clf* c, int f, const float* xp, int P, const float* xn, int Nn)
{
    unsigned numberOfIterations = 0;
    float w0 = 0.0f, w1 = (float)0.5;
    unsigned total = (unsigned)(P + Nn);
    int error;
    int maxAllowed = P < Nn ? P : Nn;
    int previousErrorCount = maxAllowed;
    while (numberOfIterations < 100) {
        unsigned errorCount = 0;
        for (int i = 0; i < P; i++) {
            double fv = xp[i];
            int result = (fv * w0) > w1 ? 1 : 0;
            error = 1 - result;
            if (error != 0) { w0 = (float)(w0 + 0.1 * error * fv); errorCount++; }
        }
        for (int i = 0; i < Nn; i++) {
            double fv = xn[i];
            int result = (fv * w0) > w1 ? 1 : 0;
            error = 0 - result;
            if (error != 0) { w0 = (float)(w0 + 0.1 * error * fv); errorCount++; }
        }
        /* :142 static_cast<float>(int - unsigned) : unsigned wrap-around quirk */
        float improvement = ((float)((unsigned)previousErrorCount - errorCount) / total);
        if (errorCount < (unsigned)maxAllowed && improvement < 0.01) break;
        previousErrorCount = (int)errorCount;
        numberOfIterations++;
    }
    c->pw0[f] = w0; c->pw1[f] = w1;
}

/* OnlineStumps::ClassifyF :83-90 / WeightedStumps::ClassifyF :50-57 / Perceptron::ClassifyF :78-89 */
static float weak_classify(const orc_clf* c, int f, float xx)
{
    if (c->weak_type == ORC_WEAK_PERCEPTRON) {
        double fv = xx;
        double response = fv * c->pw0[f] > c->pw1[f] ? 1 : -1;
        return (float)response;
    }
    float d0 = xx - c->mu0[f], d1 = xx - c->mu1[f];
    float a0 = d0 * d0; a0 = a0 * c->e0[f];
    float a1 = d1 * d1; a1 = a1 * c->e1[f];
    double p0 = (double)(expf(a0) * c->n0[f]);
    double p1 = (double)(expf(a1) * c->n1[f]);
    return (float)(log(1e-5 + p1) - log(1e-5 + p0));
}
/* IsValidWeakClassifier: OnlineStumps :25-36 (mu0 != mu1); WeightedStumps header; Perceptron: true */
static int weak_valid(const orc_clf* c, int f)
{
    if (c->weak_type == ORC_WEAK_PERCEPTRON) return 1;
    return c->mu0[f] != c->mu1[f];
}
/* Public.h:169-172 sigmoid(float): 1.0f/(1.0f+exp(-rate*x)) with the float exp overload */
static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-1 * x)); }

static int in_list(const int* l, int n, int v) { for (int i = 0; i < n; i++) if (l[i] == v) return 1; return 0; }
/* Order of sort_order / sort_order_des (Public.h:211-245) with its two unspecified points pinned the way oracle/_ref is built
   (ref_shim/boost/shared_ptr.hpp): exactly tied scores keep index order, NaN scores rank last.  NaN scores are real: the
   appearance fuser's classifiers (learning rate 0) get sigma = 0 for colour bins that are empty in a class on their second
   Update, hence n = inf, e = -inf and NaN responses; std::sort on them is undefined behaviour in the reference. */
static int lt_nanlast(float a, float b) { int an = a != a, bn = b != b; if (an || bn) return !an && bn; return a < b; }
static int gt_nanlast(float a, float b) { int an = a != a, bn = b != b; if (an || bn) return !an && bn; return a > b; }

/* MILBoostClassifier.cpp:61-131 (selection).  pp [F][P], pn [F][Nn].
   sort_order (Public.h:213-228) is std::sort on values only; ties -> smallest index (stable). */
static void milboost_select(orc_clf* c, const float* pp, int P, const float* pn, int Nn)
{
    int F = c->F;
    float* Hp = (float*)calloc((size_t)(P > 0 ? P : 1), 4);
    float* Hn = (float*)calloc((size_t)(Nn > 0 ? Nn : 1), 4);
    float* nll = (float*)malloc((size_t)F * 4);
    c->n_sel = 0;
    double previousNLL = 1000000;
    for (int s = 0; s < c->K; s++) {
#ifdef _OPENMP
#pragma omp parallel for num_threads(g_threads) schedule(static)
#endif
        for (int w = 0; w < F; w++) {
            float like = 1.0f;
            for (int i = 0; i < P; i++) like *= (1 - sigmoidf_(Hp[i] + pp[(size_t)w * P + i]));
            float posl = (float)-log(1 - like + 1e-5);            /* :83 double log */
            like = 0.0f;
            for (int j = 0; j < Nn; j++)
                like += (float)-logf(1e-5f + 1 - sigmoidf_(Hn[j] + pn[(size_t)w * Nn + j]));   /* :88 float log */
            nll[w] = posl / (float)P + like / (float)Nn;          /* :92 float / size_t */
        }
        int best = -1;
        for (int w = 0; w < F; w++) {
            if (in_list(c->sel, c->n_sel, w) || !weak_valid(c, w)) continue;
            if (best < 0 || lt_nanlast(nll[w], nll[best])) best = w;
        }
        if (best < 0) break;                                      /* reference: out-of-range read (UB) */
        if ((previousNLL - nll[best]) < 1e-100) break;            /* :108-112 push then pop */
        c->sel[c->n_sel++] = best;
        previousNLL = nll[best];
        for (int i = 0; i < P; i++) Hp[i] += pp[(size_t)best * P + i];
        for (int j = 0; j < Nn; j++) Hn[j] += pn[(size_t)best * Nn + j];
    }
    free(Hp); free(Hn); free(nll);
}

/* stable descending order of obj[] (sort_order_des, Public.h:230-245; ties -> smallest index) */
typedef struct { float v; int i; } vi_t;
static int cmp_desc(const void* a, const void* b)
{
    const vi_t* x = (const vi_t*)a; const vi_t* y = (const vi_t*)b;
    if (gt_nanlast(x->v, y->v)) return -1;
    if (gt_nanlast(y->v, x->v)) return 1;
    return x->i - y->i;
}
static int cmp_asc(const void* a, const void* b)
{
    const vi_t* x = (const vi_t*)a; const vi_t* y = (const vi_t*)b;
    if (lt_nanlast(x->v, y->v)) return -1;
    if (lt_nanlast(y->v, x->v)) return 1;
    return x->i - y->i;
}

/* MILAnyBoostClassifier.cpp:58-222 */
static void milanyboost_select(orc_clf* c, const float* pp, int P, const float* pn, int Nn)
{
    int F = c->F;
    float* Hp = (float*)calloc((size_t)(P > 0 ? P : 1), 4);
    float* Hn = (float*)calloc((size_t)(Nn > 0 ? Nn : 1), 4);
    float* pip = (float*)calloc((size_t)(P > 0 ? P : 1), 4);
    float* pin = (float*)calloc((size_t)(Nn > 0 ? Nn : 1), 4);
    float* wp = (float*)calloc((size_t)(P > 0 ? P : 1), 4);
    float* wn = (float*)calloc((size_t)(Nn > 0 ? Nn : 1), 4);
    vi_t* obj = (vi_t*)malloc((size_t)F * sizeof(vi_t));
    int* order = (int*)malloc((size_t)F * sizeof(int));
    int n_order = 0;
    int k = 0;
    c->n_sel = 0;
    double previousNLL = 1000000;
    for (int s = 0; s < c->K; s++) {
        float like = 1.0f;
        for (int i = 0; i < P; i++) { pip[i] = sigmoidf_(Hp[i]); like *= (1 - pip[i]); }
        float pbag = 1 - powf(like, 1 / (float)P);                         /* :90 */
        float nll = -logf(pbag + 1e-5f);                                   /* :92 */
        for (int j = 0; j < Nn; j++) {
            pin[j] = sigmoidf_(Hn[j]);
            nll += -logf(1e-5f + 1 - pin[j]) / (float)Nn;                  /* :98 */
        }
        if ((previousNLL - nll) < 1e-100) {
            if (c->n_sel <= 2) {                                           /* :104-150 "special" retry */
                s--;
                if (s < 0 || c->n_sel == 0) break;                         /* unreachable: previous = 1e6 at s=0 */
                int last = c->sel[s];
                for (int i = 0; i < P; i++) Hp[i] -= pp[(size_t)last * P + i];
                for (int j = 0; j < Nn; j++) Hn[j] -= pn[(size_t)last * Nn + j];
                c->n_sel--;
                k++;
                for (; k < n_order; k++)
                    if (!in_list(c->sel, c->n_sel, order[k])) { c->sel[c->n_sel++] = order[k]; break; }
                if (k >= n_order) break;                                    /* reference: abortError */
                int nw = c->sel[s];
                for (int i = 0; i < P; i++) Hp[i] += pp[(size_t)nw * P + i];
                for (int j = 0; j < Nn; j++) Hn[j] += pn[(size_t)nw * Nn + j];
                continue;
            } else {
                c->n_sel--;                                                /* :152-155 pop_back; break */
                break;
            }
        } else {
            previousNLL = nll;
        }
        for (int i = 0; i < P; i++) wp[i] = (1 / (float)P) * (1 - pbag) * pip[i] / pbag;   /* :163-167 */
        for (int j = 0; j < Nn; j++) wn[j] = -(1 / (float)Nn) * pin[j];                    /* :169-173 */
#ifdef _OPENMP
#pragma omp parallel for num_threads(g_threads) schedule(static)
#endif
        for (int w = 0; w < F; w++) {
            float o = 0;
            for (int i = 0; i < P; i++) { float t = wp[i] * pp[(size_t)w * P + i]; o += t; }
            for (int j = 0; j < Nn; j++) { float t = wn[j] * pn[(size_t)w * Nn + j]; o += t; }
            obj[w].v = o; obj[w].i = w;
        }
        qsort(obj, (size_t)F, sizeof(vi_t), cmp_desc);
        for (int w = 0; w < F; w++) order[w] = obj[w].i;
        n_order = F;
        for (k = 0; k < n_order; k++)
            if (!in_list(c->sel, c->n_sel, order[k])) { c->sel[c->n_sel++] = order[k]; break; }
        if (k == n_order) break;
        int nw = c->sel[s];
        for (int i = 0; i < P; i++) Hp[i] += pp[(size_t)nw * P + i];
        for (int j = 0; j < Nn; j++) Hn[j] += pn[(size_t)nw * Nn + j];
    }
    free(Hp); free(Hn); free(pip); free(pin); free(wp); free(wn); free(obj); free(order);
}

/* MILBoostClassifier.cpp:15-157 / MILAnyBoostClassifier.cpp:14-254 from feature matrices */
/* WeakClassifier::Classify (bool): OnlineStumps :68-75 compares the two class likelihoods, WeightedStumps (header :25) and
   Perceptron (:61-68) take the sign of ClassifyF */
static int weak_classify_bool(const orc_clf* c, int f, float xx)
{
    if (c->weak_type == ORC_WEAK_STUMP) {
        float d0 = xx - c->mu0[f], d1 = xx - c->mu1[f];
        float a0 = d0 * d0; a0 = a0 * c->e0[f];
        float a1 = d1 * d1; a1 = a1 * c->e1[f];
        double p0 = (double)(expf(a0) * c->n0[f]);
        double p1 = (double)(expf(a1) * c->n1[f]);
        return p1 > p0;
    }
    return weak_classify(c, f, xx) > 0;
}
/* AdaBoostClassifier::Update selection loop (AdaBoostClassifier.cpp:62-171): online boosting with running TP/FP/TN/FN
   weights per (selector, weak classifier).  bp [F][P], bn [F][Nn] = ClassifySet labels.  sort_order sorts errs in place,
   so errs[k] IS the error of order[k]; ties in std::sort are unspecified in the reference, index order here. */
static void adaboost_select(orc_clf* c, const unsigned char* bp, int P, const unsigned char* bn, int Nn)
{
    int F = c->F, K = c->K;
    float* poslam = (float*)malloc((size_t)(P > 0 ? P : 1) * 4);
    float* neglam = (float*)malloc((size_t)(Nn > 0 ? Nn : 1) * 4);
    for (int i = 0; i < P; i++) poslam[i] = .5f / (float)P;
    for (int j = 0; j < Nn; j++) neglam[j] = .5f / (float)Nn;
    vi_t* errs = (vi_t*)malloc((size_t)F * sizeof(vi_t));
    c->sum_alpha = 0.0f; c->n_sel = 0;
    for (int s = 0; s < K; s++) {
        float *TP = c->cTP + (size_t)s * F, *FP = c->cFP + (size_t)s * F, *TN = c->cTN + (size_t)s * F, *FN = c->cFN + (size_t)s * F;
        for (int w = 0; w < F; w++) {
            for (int i = 0; i < P; i++) { if (bp[(size_t)w * P + i]) TP[w] += poslam[i]; else FP[w] += poslam[i]; }
            for (int j = 0; j < Nn; j++) { if (!bn[(size_t)w * Nn + j]) TN[w] += neglam[j]; else FN[w] += neglam[j]; }
            errs[w].v = (FP[w] + FN[w]) / (FP[w] + FN[w] + TP[w] + TN[w]);
            errs[w].i = w;
        }
        qsort(errs, (size_t)F, sizeof(vi_t), cmp_asc);
        float minerr = 0; int bestind = 0, found = 0;
        for (int k = 0; k < F; k++) {
            if (!in_list(c->sel, c->n_sel, errs[k].i)) { c->sel[c->n_sel++] = errs[k].i; minerr = errs[k].v; bestind = errs[k].i; found = 1; break; }
        }
        (void)found;
        float a = 0.5f * logf((1 - minerr) / (minerr + 0.00001f));
        if (a > (float)10) a = (float)10;
        if (a < (float)0) a = (float)0;
        c->alpha[s] = a;
        c->sum_alpha += a;
        float corw = 1 / (2 - 2 * minerr), incorw = 1 / (2 * minerr);
        for (int j = 0; j < P; j++) poslam[j] *= (bp[(size_t)bestind * P + j] == 1) ? corw : incorw;
        for (int j = 0; j < Nn; j++) neglam[j] *= (bn[(size_t)bestind * Nn + j] == 0) ? corw : incorw;
    }
    free(poslam); free(neglam); free(errs);
}

/* ---- MILEnsembleClassifier (MILEnsembleClassifier.cpp:14-337) ------------------------------------------------
   sort_order (Public.h:213-228) sorts the VALUES in place as well, so after it `negLogLikelihoodList[k]` is the k-th
   smallest value (the chosen one's) while `negLogLikelihoodList[order[k]]` (:131, :321) is the value of RANK order[k] --
   a quirk that feeds the stopping test of the following rounds; restated as written.  Ties -> lowest index. */
static void sort_order_f(const float* v, int n, float* sorted, int* order)
{
    vi_t* a = (vi_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(vi_t));
    for (int i = 0; i < n; i++) { a[i].v = v[i]; a[i].i = i; }
    qsort(a, (size_t)n, sizeof(vi_t), cmp_asc);
    for (int i = 0; i < n; i++) { sorted[i] = a[i].v; order[i] = a[i].i; }
    free(a);
}
/* RetainBestPerformingWeakClassifiers (:206-337).  f32 positive chain (:270), predictions from the weak state of the
   previous frame on the new samples; Hp / Hn come back advanced by the retained selectors. */
static void ensemble_retain(orc_clf* c, const float* fpos, int P, const float* fneg, int Nn, float* Hp, float* Hn)
{
    const int n_prev = c->n_sel;
    const int n_retain = (int)(c->pct_retained * (float)n_prev);             /* :214 float * size_t */
    if (n_prev == 0) return;
    if (n_retain == 0) { c->n_sel = 0; return; }
    int* prev = (int*)malloc((size_t)n_prev * sizeof(int));
    for (int i = 0; i < n_prev; i++) prev[i] = c->sel[i];
    c->n_sel = 0;
    float* pp = (float*)malloc((size_t)n_prev * P * 4);
    float* pn = (float*)malloc((size_t)n_prev * Nn * 4);
    for (int q = 0; q < n_prev; q++) {
        int f = prev[q];
        for (int i = 0; i < P; i++) pp[(size_t)q * P + i] = weak_classify(c, f, fpos[(size_t)f * P + i]);
        for (int j = 0; j < Nn; j++) pn[(size_t)q * Nn + j] = weak_classify(c, f, fneg[(size_t)f * Nn + j]);
    }
    float* nll = (float*)malloc((size_t)n_prev * 4);
    float* sorted = (float*)malloc((size_t)n_prev * 4);
    int* order = (int*)malloc((size_t)n_prev * sizeof(int));
    double previousNLL = 1000000;
    for (int s = 0; s < n_retain; s++) {
        for (int q = 0; q < n_prev; q++) {
            float like = 1.0f;
            for (int i = 0; i < P; i++) like *= (1 - sigmoidf_(Hp[i] + pp[(size_t)q * P + i]));
            float posl = (float)-log(1 - like + 1e-5);
            like = 0.0f;
            for (int j = 0; j < Nn; j++) like += (float)-logf(1e-5f + 1 - sigmoidf_(Hn[j] + pn[(size_t)q * Nn + j]));
            nll[q] = posl / (float)P + like / (float)Nn;
        }
        sort_order_f(nll, n_prev, sorted, order);
        int k = 0;
        for (; k < n_prev; k++) if (!in_list(c->sel, c->n_sel, prev[order[k]])) break;
        if (k == n_prev) break;                                               /* :304-307 */
        if ((previousNLL - (double)sorted[k]) < 1e-100) break;                /* :309-314 push then pop */
        c->sel[c->n_sel++] = prev[order[k]];
        previousNLL = sorted[order[k]];                                       /* :318 value at rank order[k] */
        const int q = order[k];
        for (int i = 0; i < P; i++) Hp[i] += pp[(size_t)q * P + i];
        for (int j = 0; j < Nn; j++) Hn[j] += pn[(size_t)q * Nn + j];
    }
    free(prev); free(pp); free(pn); free(nll); free(sorted); free(order);
}
/* MILEnsembleClassifier::Update (:14-157).  Retained weak classifiers keep their state and are only evaluated (:59-65);
   every other one is re-initialised and trained from scratch on this frame's samples (:67-72).  The positive bag
   likelihood and the negative sum are DOUBLE accumulators here (:88, :100), unlike MILBoost and the retain step. */
static void milensemble_update(orc_clf* c, const float* fpos, int P, const float* fneg, int Nn,
                               const float* pwn, const float* nwn)
{
    int F = c->F;
    float* Hp = (float*)calloc((size_t)(P > 0 ? P : 1), 4);
    float* Hn = (float*)calloc((size_t)(Nn > 0 ? Nn : 1), 4);
    ensemble_retain(c, fpos, P, fneg, Nn, Hp, Hn);
    float* pp = (float*)malloc((size_t)F * (P > 0 ? P : 1) * 4);
    float* pn = (float*)malloc((size_t)F * (Nn > 0 ? Nn : 1) * 4);
#ifdef _OPENMP
#pragma omp parallel for num_threads(g_threads) schedule(static)
#endif
    for (int f = 0; f < F; f++) {
        const float* xp = fpos + (size_t)f * P;
        const float* xn = fneg + (size_t)f * Nn;
        if (!in_list(c->sel, c->n_sel, f)) {
            c->mu0[f] = 0; c->mu1[f] = 0; c->sig0[f] = 1; c->sig1[f] = 1; c->trained[f] = 0;     /* Initialize() */
            if (c->weak_type == ORC_WEAK_STUMP) stump_update(c, f, xp, P, xn, Nn);
            else if (c->weak_type == ORC_WEAK_WSTUMP) wstump_update(c, f, xp, P, xn, Nn, pwn, nwn);
            else perceptron_update(c, f, xp, P, xn, Nn);
        }
        for (int i = 0; i < P; i++) pp[(size_t)f * P + i] = weak_classify(c, f, xp[i]);
        for (int j = 0; j < Nn; j++) pn[(size_t)f * Nn + j] = weak_classify(c, f, xn[j]);
    }
    float* nll = (float*)malloc((size_t)F * 4);
    float* sorted = (float*)malloc((size_t)F * 4);
    int* order = (int*)malloc((size_t)F * sizeof(int));
    double previousNLL = 1000000;
    for (int s = c->n_sel; s < c->K; s++) {
#ifdef _OPENMP
#pragma omp parallel for num_threads(g_threads) schedule(static)
#endif
        for (int w = 0; w < F; w++) {
            double like = 1.0f;
            for (int i = 0; i < P; i++) like *= (1 - sigmoidf_(Hp[i] + pp[(size_t)w * P + i]));
            float posl = (float)-log(1 - like + 1e-5);
            like = 0.0f;
            for (int j = 0; j < Nn; j++) like += (float)-logf(1e-5f + 1 - sigmoidf_(Hn[j] + pn[(size_t)w * Nn + j]));
            float negl = (float)like;
            nll[w] = posl / (float)P + negl / (float)Nn;
        }
        sort_order_f(nll, F, sorted, order);
        int k = 0;
        for (; k < F; k++) if (!in_list(c->sel, c->n_sel, order[k])) break;   /* no IsValidWeakClassifier test here (:114) */
        if (k == F) break;                                                    /* reference: out-of-range read (UB) */
        if ((previousNLL - (double)sorted[k]) < 1e-100) break;                /* :122-127 */
        const int best = order[k];
        c->sel[c->n_sel++] = best;
        previousNLL = sorted[order[k]];                                       /* :131 value at rank order[k] */
        for (int i = 0; i < P; i++) Hp[i] += pp[(size_t)best * P + i];
        for (int j = 0; j < Nn; j++) Hn[j] += pn[(size_t)best * Nn + j];
    }
    free(Hp); free(Hn); free(pp); free(pn); free(nll); free(sorted); free(order);
}

void orc_clf_update_from_features(orc_clf* c, const float* fpos, int P, const float* fneg, int Nn,
                                  const float* wpos, const float* wneg)
{
    int F = c->F;
    float* pp = (float*)malloc((size_t)F * (P > 0 ? P : 1) * 4);
    float* pn = (float*)malloc((size_t)F * (Nn > 0 ? Nn : 1) * 4);
    float* pwn = (float*)malloc((size_t)(P > 0 ? P : 1) * 4);
    float* nwn = (float*)malloc((size_t)(Nn > 0 ? Nn : 1) * 4);
    if (c->weak_type == ORC_WEAK_WSTUMP) { normalize_w(wpos, P, pwn); normalize_w(wneg, Nn, nwn); }
    if (c->strong_type == ORC_STRONG_MILENSEMBLE) {
        milensemble_update(c, fpos, P, fneg, Nn, pwn, nwn);
        free(pp); free(pn); free(pwn); free(nwn);
        return;
    }
#ifdef _OPENMP
#pragma omp parallel for num_threads(g_threads) schedule(static)
#endif
    for (int f = 0; f < F; f++) {
        const float* xp = fpos + (size_t)f * P;
        const float* xn = fneg + (size_t)f * Nn;
        if (c->weak_type == ORC_WEAK_STUMP) stump_update(c, f, xp, P, xn, Nn);
        else if (c->weak_type == ORC_WEAK_WSTUMP) wstump_update(c, f, xp, P, xn, Nn, pwn, nwn);
        else perceptron_update(c, f, xp, P, xn, Nn);
        for (int i = 0; i < P; i++) pp[(size_t)f * P + i] = weak_classify(c, f, xp[i]);
        for (int j = 0; j < Nn; j++) pn[(size_t)f * Nn + j] = weak_classify(c, f, xn[j]);
    }
    if (c->strong_type == ORC_STRONG_ADABOOST) {
        unsigned char* bp = (unsigned char*)malloc((size_t)F * (P > 0 ? P : 1));
        unsigned char* bn = (unsigned char*)malloc((size_t)F * (Nn > 0 ? Nn : 1));
        for (int f = 0; f < F; f++) {
            for (int i = 0; i < P; i++) bp[(size_t)f * P + i] = (unsigned char)weak_classify_bool(c, f, fpos[(size_t)f * P + i]);
            for (int j = 0; j < Nn; j++) bn[(size_t)f * Nn + j] = (unsigned char)weak_classify_bool(c, f, fneg[(size_t)f * Nn + j]);
        }
        adaboost_select(c, bp, P, bn, Nn);
        free(bp); free(bn);
    }
    else if (c->strong_type == ORC_STRONG_MILANYBOOST) milanyboost_select(c, pp, P, pn, Nn);
    else milboost_select(c, pp, P, pn, Nn);
    free(pp); free(pn); free(pwn); free(nwn);
}

void orc_clf_update(orc_clf* c, const orc_frame* f, const orc_boxes* pos, const orc_boxes* neg,
                    const float* wpos, const float* wneg)
{
    int F = c->F, P = pos->n, Nn = neg->n;
    float* fp = (float*)malloc((size_t)F * (P > 0 ? P : 1) * 4);
    float* fn = (float*)malloc((size_t)F * (Nn > 0 ? Nn : 1) * 4);
    orc_clf_features(c, f, pos, fp, P);
    orc_clf_features(c, f, neg, fn, Nn);
    orc_clf_update_from_features(c, fp, P, fn, Nn, wpos, wneg);
    free(fp); free(fn);
}

/* MILBoostClassifier.cpp:339-374 (identical in MILAnyBoostClassifier.cpp:262-296) */
void orc_clf_classify_from_features(const orc_clf* c, const float* feat, int n, int ld, int log_ratio, float* out)
{
    for (int i = 0; i < n; i++) out[i] = 0.0f;
    if (c->strong_type == ORC_STRONG_ADABOOST) {                   /* AdaBoostClassifier::Classify :181-224 */
        for (int s = 0; s < c->n_sel; s++) {
            int f = c->sel[s];
            for (int i = 0; i < n; i++) out[i] += weak_classify_bool(c, f, feat[(size_t)f * ld + i]) ? c->alpha[s] : -c->alpha[s];
        }
        if (!log_ratio) for (int i = 0; i < n; i++) out[i] = sigmoidf_(2 * out[i]);
        return;
    }
    for (int s = 0; s < c->n_sel; s++) {
        int f = c->sel[s];
        for (int i = 0; i < n; i++) out[i] += weak_classify(c, f, feat[(size_t)f * ld + i]);
    }
    if (!log_ratio) for (int i = 0; i < n; i++) out[i] = sigmoidf_(out[i]);
}
void orc_clf_classify(const orc_clf* c, const orc_frame* f, const orc_boxes* b, int log_ratio, float* out)
{
    if (b->n == 0) return;
    float* feat = (float*)malloc((size_t)c->F * b->n * 4);
    orc_clf_features(c, f, b, feat, b->n);      /* the reference computes ALL F (:345-348) */
    orc_clf_classify_from_features(c, feat, b->n, b->n, log_ratio, out);
    free(feat);
}
int orc_clf_get_alphas(const orc_clf* c, float* a)
{
    if (!c->alpha) return 0;
    for (int i = 0; i < c->n_sel; i++) a[i] = c->alpha[i];
    return c->n_sel;
}
int orc_clf_get_selectors(const orc_clf* c, int* idx)
{
    for (int i = 0; i < c->n_sel; i++) idx[i] = c->sel[i];
    return c->n_sel;
}
void orc_clf_get_weak_state(const orc_clf* c, float* o)
{
    int F = c->F;
    if (c->weak_type == ORC_WEAK_PERCEPTRON) {
        memset(o, 0, (size_t)8 * F * 4);
        memcpy(o, c->pw0, (size_t)F * 4); memcpy(o + F, c->pw1, (size_t)F * 4);
        return;
    }
    memcpy(o + 0 * F, c->mu0, (size_t)F * 4); memcpy(o + 1 * F, c->mu1, (size_t)F * 4);
    memcpy(o + 2 * F, c->sig0, (size_t)F * 4); memcpy(o + 3 * F, c->sig1, (size_t)F * 4);
    memcpy(o + 4 * F, c->n0, (size_t)F * 4); memcpy(o + 5 * F, c->n1, (size_t)F * 4);
    memcpy(o + 6 * F, c->e0, (size_t)F * 4); memcpy(o + 7 * F, c->e1, (size_t)F * 4);
}

/* ------------------------------------------------------------------------------------------ */
/* ParticleFilterTracker.cpp:541-566: min/max (first occurrence, Public.h:256-268), then        */
/* float(pow((H - min) / range, exponent)) evaluated in double; 1 if range == 0.               */
/* ------------------------------------------------------------------------------------------ */
void orc_likelihood_map(float* h, int n, int exponent)
{
    if (n <= 0) return;
    int bi = 0, wi = 0;
    for (int i = 1; i < n; i++) { if (h[i] > h[bi]) bi = i; if (h[i] < h[wi]) wi = i; }
    double mx = h[bi], mn = h[wi];
    double range = mx - mn;
    for (int i = 0; i < n; i++) {
        if (range != 0) h[i] = (float)pow((h[i] - mn) / range, (double)exponent);
        else h[i] = 1;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* ParticleFilter                                                                              */
/* ------------------------------------------------------------------------------------------ */
#define ST(p, i, j) ((p)->state[(size_t)(i) * 5 + (j)])
#define RS(p, i, j) ((p)->resampled[(size_t)(i) * 5 + (j)])

orc_pf* orc_pf_create(int N, float img_w, float img_h)
{
    orc_pf* p = (orc_pf*)calloc(1, sizeof(*p));
    p->N = N; p->img_w = img_w; p->img_h = img_h;
    p->state = (float*)calloc((size_t)N * 5, 4);
    p->resampled = (float*)calloc((size_t)N * 5, 4);
    p->resample_idx = (int*)calloc((size_t)N, sizeof(int));
    p->uniq = (float*)calloc((size_t)N * 5, 4);
    p->filter_state = ORC_PF_UNINIT;
    return p;
}
void orc_pf_free(orc_pf* p)
{
    if (!p) return;
    free(p->state); free(p->resampled); free(p->resample_idx); free(p->uniq); free(p);
}
/* ParticleFilter.cpp:15-75 (PF_INITIALIZATION_BOUNDARY = 0, ParticleFilter.h:11) */
void orc_pf_initialize(orc_pf* p, float cx, float cy, float sx, float sy, float msc, float maxs, float mins)
{
    p->max_scale_change = msc; p->max_scale = maxs; p->min_scale = mins;
    for (int i = 0; i < p->N; i++) {
        RS(p, i, 0) = cx; RS(p, i, 1) = cy; RS(p, i, 2) = sx; RS(p, i, 3) = sy; RS(p, i, 4) = 1;
    }
    memcpy(p->state, p->resampled, (size_t)p->N * 5 * 4);
    p->filter_state = ORC_PF_INIT;
    p->time_index = 0; p->uniq_valid = 0;
}
/* ParticleFilter.cpp:187-226: only m_particlesResampledMatrix is touched */
static void pf_force_refinement(orc_pf* p, float width, float height)
{
    for (int r = 0; r < p->N; r++) {
        float leftX = RS(p, r, 0) - RS(p, r, 2) * width / 2;
        float topY = RS(p, r, 1) - RS(p, r, 3) * height / 2;
        float footX = RS(p, r, 0);
        float footY = RS(p, r, 1) + RS(p, r, 3) * height / 2;
        if (leftX < 0) RS(p, r, 0) = footX > 0.0f ? footX : 0.0f;
        if (leftX >= p->img_w) RS(p, r, 0) = footX < p->img_w - 1 ? footX : p->img_w - 1;
        if (topY < 0) { float v = footY - RS(p, r, 2) * height / 2; RS(p, r, 1) = v > 0.0f ? v : 0.0f; }
        if (topY >= p->img_h) { float v = footY - RS(p, r, 2) * height / 2; RS(p, r, 1) = v < p->img_h - 1 ? v : p->img_h - 1; }
        RS(p, r, 4) = 1;
    }
}
/* ParticleFilter.cpp:83-112 */
static void pf_check_refinement(orc_pf* p, float width, float height)
{
    int valid = 0;
    for (int r = 0; r < p->N; r++) {
        float leftX = RS(p, r, 0) - RS(p, r, 2) * width / 2;
        float topY = RS(p, r, 1) - RS(p, r, 3) * height / 2;
        if (leftX < 0 || leftX >= p->img_w || topY < 0 || topY >= p->img_h) continue;
        valid++;
    }
    if (valid == 0) pf_force_refinement(p, width, height);
}
static float scale_step(float s, float n, float msc, float maxs, float mins)
{
    if (msc > fabsf(n)) s = s + n;
    else if (n < 0) s = s - msc;
    else s = s + msc;
    if (s > maxs) s = maxs;
    if (s < mins) s = mins;
    return s;
}
/* ParticleFilter.cpp:236-341; noise rows = the four cv::randn(1xN, 0, sigma) draws (x, y, sx, sy),
   generated by the caller (cv::theRNG() stream is not part of the parity claim) */
void orc_pf_predict(orc_pf* p, const float* nz, float w0, float h0)
{
    int N = p->N;
    for (int i = 0; i < N; i++) { float v = ST(p, i, 0) + ST(p, i, 2) * nz[0 * N + i]; ST(p, i, 0) = (float)orc_cvround((double)v); }
    for (int i = 0; i < N; i++) { float v = ST(p, i, 1) + ST(p, i, 3) * nz[1 * N + i]; ST(p, i, 1) = (float)orc_cvround((double)v); }
    for (int i = 0; i < N; i++) ST(p, i, 2) = scale_step(ST(p, i, 2), nz[2 * N + i], p->max_scale_change, p->max_scale, p->min_scale);
    for (int i = 0; i < N; i++) ST(p, i, 3) = scale_step(ST(p, i, 3), nz[3 * N + i], p->max_scale_change, p->max_scale, p->min_scale);
    p->filter_state = ORC_PF_PREDICTED;
    p->uniq_valid = 0;
    p->time_index++;
    pf_check_refinement(p, w0, h0);
}
/* ParticleFilter.cpp:641-671 */
void orc_pf_update_particle_weight(orc_pf* p, int i, float prob, int force_reset)
{
    if (!force_reset) ST(p, i, 4) = prob * ST(p, i, 4);
    else ST(p, i, 4) = prob;
    if (ST(p, i, 2) / ST(p, i, 3) > 1.2) ST(p, i, 3) = (float)(0.9 * ST(p, i, 2));
    else if (ST(p, i, 3) / ST(p, i, 2) > 1.2) ST(p, i, 2) = (float)(0.9 * ST(p, i, 3));
    p->uniq_valid = 0;
}
/* ParticleFilter.cpp:987-1011 */
void orc_pf_normalize(orc_pf* p)
{
    float total = 0.0f;
    for (int i = 0; i < p->N; i++) total += ST(p, i, 4);
    for (int i = 0; i < p->N; i++) {
        if (total != 0) ST(p, i, 4) = ST(p, i, 4) / total;
        else ST(p, i, 4) = (float)1e-2;
    }
}
/* ParticleFilter.cpp:920-978 */
void orc_pf_resample_u01(orc_pf* p, int force, float u01)
{
    int N = p->N;
    orc_pf_normalize(p);
    float* cdf = (float*)malloc((size_t)N * 4);
    cdf[0] = 0.0f;                                   /* w[0] excluded: quirk */
    for (int i = 1; i < N; i++) cdf[i] = cdf[i - 1] + ST(p, i, 4);
    float seed = (1 / (float)N) * u01;
    int idx = 0;
    for (int i = 0; i < N; i++) {
        float thr = seed + (1 / (float)N) * (float)i;
        while (idx < N - 1 && thr > cdf[idx]) {
            idx++;
            if ((idx + 1) == N) break;
        }
        p->resample_idx[i] = idx;
        for (int j = 0; j < 4; j++) RS(p, i, j) = ST(p, idx, j);
        RS(p, i, 4) = 1;
    }
    if (force) {
        memcpy(p->state, p->resampled, (size_t)N * 5 * 4);
        p->filter_state = ORC_PF_RESAMPLED;
    }
    free(cdf);
}
void orc_pf_resample(orc_pf* p, int force, orc_rng* r) { orc_pf_resample_u01(p, force, orc_randfloat(r)); }
/* ParticleFilter.cpp:845-910, WEIGHT_BASED (ParticleFilter.h:48 default) */
int orc_pf_resample_if_required(orc_pf* p, orc_rng* r)
{
    float nEffInv = 0;
    float nThreshold = (float)(p->N * 0.01);
    for (int i = 0; i < p->N; i++) {
        double w = ST(p, i, 4);
        nEffInv = (float)(nEffInv + w * w);           /* pow(float,2) is double; float + double -> double -> float */
    }
    p->last_neff_inv = nEffInv;
    p->last_resampled = 0;
    if (nEffInv == 0) return 0;
    if ((1 / nEffInv) < nThreshold) { orc_pf_resample(p, 1, r); p->last_resampled = 1; return 1; }
    return 0;
}
/* ParticleFilter.cpp:680-717 (shouldUpdateOnlyNonZeroWeights = true is the only path callers take) */
int orc_pf_update_all_weights(orc_pf* p, const float* prob, int only_nonzero, int force_reset, int resample, orc_rng* r)
{
    if (!only_nonzero) {
        for (int i = 0; i < p->N; i++) { orc_pf_update_particle_weight(p, i, prob[i], force_reset); i++; } /* :692 double increment */
    } else {
        int p2 = 0;
        for (int i = 0; i < p->N; i++)
            if (ST(p, i, 4) != 0) orc_pf_update_particle_weight(p, i, prob[p2++], force_reset);
    }
    orc_pf_normalize(p);
    if (resample) return orc_pf_resample_if_required(p, r);
    return 0;
}
/* ParticleFilter.cpp:728-815 */
void orc_pf_find_ordered_unique(orc_pf* p)
{
    if (p->uniq_valid) return;
    int N = p->N;
    vi_t* v = (vi_t*)malloc((size_t)N * sizeof(vi_t));
    if (p->filter_state == ORC_PF_PREDICTED) {
        for (int i = 0; i < N; i++) { v[i].v = ST(p, i, 4); v[i].i = i; }
        qsort(v, (size_t)N, sizeof(vi_t), cmp_desc);
        for (int i = 0; i < N; i++) for (int j = 0; j < 5; j++) p->uniq[(size_t)i * 5 + j] = ST(p, v[i].i, j);
        p->n_uniq = N;
    } else if (p->filter_state == ORC_PF_RESAMPLED) {
        for (int i = 0; i < N; i++) { v[i].v = (float)p->resample_idx[i]; v[i].i = i; }
        qsort(v, (size_t)N, sizeof(vi_t), cmp_asc);
        int* uidx = (int*)malloc((size_t)N * sizeof(int));
        vi_t* uw = (vi_t*)malloc((size_t)N * sizeof(vi_t));
        int it = 0;
        uidx[0] = v[0].i; uw[0].v = ST(p, v[0].i, 4); uw[0].i = 0;
        for (int q = 1; q < N; q++) {
            if (v[q].v == v[q - 1].v) uw[it].v += ST(p, v[q].i, 4);
            else { it++; uidx[it] = v[q].i; uw[it].v = ST(p, v[q].i, 4); uw[it].i = it; }
        }
        qsort(uw, (size_t)(it + 1), sizeof(vi_t), cmp_desc);
        /* :799-806: states are taken in index order, weights in sorted order (mismatch kept) */
        for (int q = 0; q <= it; q++) {
            for (int j = 0; j < 4; j++) p->uniq[(size_t)q * 5 + j] = ST(p, uidx[q], j);
            p->uniq[(size_t)q * 5 + 4] = uw[q].v;
        }
        p->n_uniq = it + 1;
        free(uidx); free(uw);
    } else {
        p->n_uniq = 0;
    }
    p->uniq_valid = 1;
    free(v);
}
/* ParticleFilter.cpp:601-633 */
void orc_pf_get_average(const orc_pf* p, float* out)
{
    for (int j = 0; j < 4; j++) out[j] = 0;
    float total = 0.0f;
    for (int i = 0; i < p->N; i++) {
        float w = ST(p, i, 4);
        for (int j = 0; j < 4; j++) out[j] = out[j] + ST(p, i, j) * w;
        total = total + w;
    }
    for (int j = 0; j < 4; j++) out[j] = out[j] / total;
}
/* ParticleFilter.cpp:529-549 (weight slot of the result is left unset in the reference) */
void orc_pf_get_highest_weight(const orc_pf* p, float* out)
{
    int b = 0;
    for (int i = 1; i < p->N; i++) if (ST(p, i, 4) > ST(p, b, 4)) b = i;
    for (int j = 0; j < 4; j++) out[j] = ST(p, b, j);
    out[4] = 0;
}
/* ParticleFilter.cpp:557-593: closestDistance is never updated -> last top-weight tie wins */
void orc_pf_get_highest_unique_close(orc_pf* p, const float* closest, float* out)
{
    orc_pf_find_ordered_unique(p);
    (void)closest;
    float hw = p->uniq[4];
    for (int j = 0; j < p->N && j < p->n_uniq; j++) {
        if (hw != p->uniq[(size_t)j * 5 + 4]) break;
        for (int q = 0; q < 5; q++) out[q] = p->uniq[(size_t)j * 5 + q];
    }
}
/* ParticleFilterTracker.cpp:1247-1305 */
int orc_pf_generate_test_set(orc_pf* p, float w0, float h0, int fw, int fh, int* col, int* row, float* sx, float* sy)
{
    int n = 0;
    for (int i = 0; i < p->N; i++) {
        float widthFloat = ST(p, i, 2) * w0;
        float heightFloat = ST(p, i, 3) * h0;
        float leftX = (float)orc_cvround((double)(ST(p, i, 0) - widthFloat / 2));
        float topY = (float)orc_cvround((double)(ST(p, i, 1) - heightFloat / 2));
        float width = (float)orc_cvround((double)widthFloat);
        float height = (float)orc_cvround((double)heightFloat);
        if (leftX <= 0 || leftX + width >= fw - 1 || topY <= 0 || topY + height >= fh - 1)
            orc_pf_update_particle_weight(p, i, 0, 0);
        else if (ST(p, i, 4) != 0) {
            col[n] = (int)leftX; row[n] = (int)topY; sx[n] = ST(p, i, 2); sy[n] = ST(p, i, 3); n++;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* SampleSet samplers                                                                          */
/* ------------------------------------------------------------------------------------------ */
/* SampleSet.cpp:331-361: keep iff randfloat() <= (max+1)/count, in order; truncate to max */
static int thin_uniform(orc_rng* r, int* col, int* row, int n, int max_n)
{
    float prob = ((float)(max_n + 1)) / (float)(unsigned)n;
    if (prob < 1) {
        int k = 0;
        for (int i = 0; i < n; i++) {
            if (orc_randfloat(r) > prob) continue;
            col[k] = col[i]; row[k] = row[i]; k++;
        }
        n = k < max_n ? k : max_n;
    }
    return n;
}
/* SampleSet.cpp:82-168 */
int orc_sample_ring(orc_rng* r, int img_w, int img_h, int x, int y, int w, int h,
                    float outer_r, float inner_r, int max_n, float sx, float sy, int* col, int* row, int cap)
{
    int scaledW = orc_cvround((double)((float)w * sx));
    int scaledH = orc_cvround((double)((float)h * sy));
    int nrows = img_h - scaledH - 1, ncols = img_w - scaledW - 1;
    float o2 = outer_r * outer_r, i2 = inner_r * inner_r;
    int minrow = 0 > y - (int)outer_r ? 0 : y - (int)outer_r;
    int maxrow = nrows - 1 < y + (int)outer_r ? nrows - 1 : y + (int)outer_r;
    int mincol = 0 > x - (int)outer_r ? 0 : x - (int)outer_r;
    int maxcol = ncols - 1 < x + (int)outer_r ? ncols - 1 : x + (int)outer_r;
    int n = 0;
    for (int rr = minrow; rr <= maxrow; rr++)
        for (int cc = mincol; cc <= maxcol; cc++) {
            int d = (y - rr) * (y - rr) + (x - cc) * (x - cc);
            if ((float)d < o2 && (float)d >= i2) { if (n < cap) { col[n] = cc; row[n] = rr; } n++; }
        }
    if (n > cap) n = cap;
    return thin_uniform(r, col, row, n, max_n);
}
/* SampleSet.cpp:180-260 */
int orc_sample_rect(orc_rng* r, int img_w, int img_h, int x, int y, int w, int h,
                    float maxdx, float maxdy, float mindx, float mindy, int max_n, float sx, float sy,
                    int* col, int* row, int cap)
{
    int scaledW = orc_cvround((double)((float)w * sx));
    int scaledH = orc_cvround((double)((float)h * sy));
    int nrows = img_h - scaledH - 1, ncols = img_w - scaledW - 1;
    int minrow = 0 > y - (int)maxdy ? 0 : y - (int)maxdy;
    int maxrow = nrows - 1 < y + (int)maxdy ? nrows - 1 : y + (int)maxdy;
    int mincol = 0 > x - (int)maxdx ? 0 : x - (int)maxdx;
    int maxcol = ncols - 1 < x + (int)maxdx ? ncols - 1 : x + (int)maxdx;
    int n = 0;
    for (int rr = minrow; rr <= maxrow; rr++)
        for (int cc = mincol; cc <= maxcol; cc++) {
            int dx = abs(x - cc), dy = abs(y - rr);
            if ((float)dx >= mindx || (float)dy >= mindy) { if (n < cap) { col[n] = cc; row[n] = rr; } n++; }
        }
    if (n > cap) n = cap;
    return thin_uniform(r, col, row, n, max_n);
}

/* ------------------------------------------------------------------------------------------ */
/* ParticleFilterTracker frame loop                                                            */
/* ------------------------------------------------------------------------------------------ */
#define TRAIN_CAP 16384
struct orc_tracker {
    orc_tracker_params tp;
    orc_clf* clf;
    orc_pf* pf;
    orc_rng* rng;
    float cur[6];                 /* m_currentStateList: leftX, topY, w, h, scaleX, scaleY */
    int last_valid;
    int ok;
    int *pcol, *prow, *ncol, *nrow; float *psx, *psy, *nsx, *nsy; int P, Nn;
    int *tcol, *trow; float *tsx, *tsy, *H;
};

static void fill_scale(float* a, int n, float v) { for (int i = 0; i < n; i++) a[i] = v; }

/* ParticleFilterTracker.cpp:101-231 */
orc_tracker* orc_tracker_create(const orc_tracker_params* tp, orc_clf* clf, const orc_frame* f0,
                                const float init[4], orc_rng* r)
{
    orc_tracker* t = (orc_tracker*)calloc(1, sizeof(*t));
    t->tp = *tp; t->clf = clf; t->rng = r;
    int N = tp->n_particles;
    t->pcol = (int*)malloc(TRAIN_CAP * 4); t->prow = (int*)malloc(TRAIN_CAP * 4);
    t->ncol = (int*)malloc(TRAIN_CAP * 4); t->nrow = (int*)malloc(TRAIN_CAP * 4);
    t->psx = (float*)malloc(TRAIN_CAP * 4); t->psy = (float*)malloc(TRAIN_CAP * 4);
    t->nsx = (float*)malloc(TRAIN_CAP * 4); t->nsy = (float*)malloc(TRAIN_CAP * 4);
    t->tcol = (int*)malloc((size_t)N * 4); t->trow = (int*)malloc((size_t)N * 4);
    t->tsx = (float*)malloc((size_t)N * 4); t->tsy = (float*)malloc((size_t)N * 4);
    t->H = (float*)malloc((size_t)N * 4);
    for (int i = 0; i < 4; i++) t->cur[i] = init[i];
    t->cur[4] = 1; t->cur[5] = 1;
    /* :155-176 initial positive disc and negative ring */
    t->P = orc_sample_ring(r, f0->W, f0->H, (int)(unsigned)t->cur[0], (int)(unsigned)t->cur[1],
                           (int)(unsigned)t->cur[2], (int)(unsigned)t->cur[3],
                           tp->init_pos_radius_train, 0, 1000000, 1, 1, t->pcol, t->prow, TRAIN_CAP);
    t->Nn = orc_sample_ring(r, f0->W, f0->H, (int)(unsigned)t->cur[0], (int)(unsigned)t->cur[1],
                            (int)(unsigned)t->cur[2], (int)(unsigned)t->cur[3],
                            2.0f * tp->search_win, 1.5f * tp->init_pos_radius_train, tp->init_n_neg_train,
                            1, 1, t->ncol, t->nrow, TRAIN_CAP);
    fill_scale(t->psx, t->P, 1); fill_scale(t->psy, t->P, 1);
    fill_scale(t->nsx, t->Nn, 1); fill_scale(t->nsy, t->Nn, 1);
    t->ok = !(t->P < 1 || t->Nn < 1);
    if (t->ok) {
        orc_boxes pb = {t->P, t->pcol, t->prow, t->psx, t->psy};
        orc_boxes nb = {t->Nn, t->ncol, t->nrow, t->nsx, t->nsy};
        orc_clf_update(clf, f0, &pb, &nb, NULL, NULL);             /* :184 */
    }
    t->pf = orc_pf_create(N, (float)f0->W, (float)f0->H);
    orc_pf_initialize(t->pf, t->cur[0] + t->cur[2] / 2, t->cur[1] + t->cur[3] / 2, t->cur[4], t->cur[5],
                      0.5f, 10.0f, 0.1f);                           /* :214-218, ParticleFilter.h:72-74 defaults */
    return t;
}
void orc_tracker_free(orc_tracker* t)
{
    if (!t) return;
    orc_pf_free(t->pf);
    free(t->pcol); free(t->prow); free(t->ncol); free(t->nrow); free(t->psx); free(t->psy); free(t->nsx); free(t->nsy);
    free(t->tcol); free(t->trow); free(t->tsx); free(t->tsy); free(t->H); free(t);
}
int orc_tracker_last_valid(const orc_tracker* t) { return t->last_valid; }
orc_pf* orc_tracker_pf(orc_tracker* t) { return t->pf; }
const float* orc_tracker_state(const orc_tracker* t) { return t->cur; }
int orc_tracker_last_train(const orc_tracker* t, int which, int* col, int* row, float* sx, float* sy, int cap)
{
    int n = which == 0 ? t->P : t->Nn;
    if (n > cap) n = cap;
    const int* c = which == 0 ? t->pcol : t->ncol; const int* r = which == 0 ? t->prow : t->nrow;
    const float* x = which == 0 ? t->psx : t->nsx; const float* y = which == 0 ? t->psy : t->nsy;
    for (int i = 0; i < n; i++) { col[i] = c[i]; row[i] = r[i]; sx[i] = x[i]; sy[i] = y[i]; }
    return n;
}

/* TrackObjectOnTheGivenFrame :468-639 */
static void tracker_score(orc_tracker* t, const orc_frame* f, const float* noise)
{
    orc_pf* pf = t->pf;
    orc_pf_predict(pf, noise, t->cur[2], t->cur[3]);
    int n = orc_pf_generate_test_set(pf, t->cur[2], t->cur[3], f->W, f->H, t->tcol, t->trow, t->tsx, t->tsy);
    if (n == 0) {
        pf_force_refinement(pf, t->cur[2], t->cur[3]);
        n = orc_pf_generate_test_set(pf, t->cur[2], t->cur[3], f->W, f->H, t->tcol, t->trow, t->tsx, t->tsy);
    }
    t->last_valid = n;
    if (n == 0) { t->ok = 0; return; }      /* reference: ASSERT_TRUE abort (:513) */
    orc_boxes tb = {n, t->tcol, t->trow, t->tsx, t->tsy};
    orc_clf_classify(t->clf, f, &tb, 1, t->H);                       /* :525 log ratio */
    orc_likelihood_map(t->H, n, 20);                                  /* :541-566 */
    orc_pf_update_all_weights(pf, t->H, 1, 0, 1, t->rng);             /* :605 */
}
/* StoreObjectState :286-393 */
static void tracker_store(orc_tracker* t, int out_box[4])
{
    float a[5];
    if (t->tp.trajectory_option == 0) {
        float prev[2] = {t->cur[0] + t->cur[2] * t->cur[4] / 2, t->cur[1] + t->cur[3] * t->cur[5] / 2};
        orc_pf_get_highest_unique_close(t->pf, prev, a);
    } else {
        orc_pf_get_average(t->pf, a);
    }
    t->cur[0] = a[0] - t->cur[2] * a[2] / 2;
    t->cur[1] = a[1] - t->cur[3] * a[3] / 2;
    t->cur[4] = a[2]; t->cur[5] = a[3];
    out_box[0] = orc_cvround((double)t->cur[0]);
    out_box[1] = orc_cvround((double)t->cur[1]);
    out_box[2] = orc_cvround((double)(t->cur[2] * t->cur[4]));
    out_box[3] = orc_cvround((double)(t->cur[3] * t->cur[5]));
}
/* GeneratePositiveTrainingSampleSet (ParticleFilterTracker.cpp:697-906): the four strategies of
   PFTracker_Positive_Sample_Strategy (TrackerParameters.h:42-51).  Boxes carry the particle's own scales. */
static int box_in_frame(int leftX, int topY, int width, int height, int fw, int fh)
{
    return leftX > 0 && topY > 0 && leftX + width < fw - 1 && topY + height < fh - 1;
}
static void particle_box(const orc_tracker* t, const float* a, int* leftX, int* topY, int* width, int* height)
{
    float widthFloat = a[2] * t->cur[2], heightFloat = a[3] * t->cur[3];
    *leftX = orc_cvround((double)(a[0] - widthFloat / 2));
    *topY = orc_cvround((double)(a[1] - heightFloat / 2));
    *width = orc_cvround((double)widthFloat);
    *height = orc_cvround((double)heightFloat);
}
static void positive_ring(orc_tracker* t, const orc_frame* f)
{
    t->P = orc_sample_ring(t->rng, f->W, f->H, (int)t->cur[0], (int)t->cur[1], (int)t->cur[2], (int)t->cur[3],
                           t->tp.pos_radius_train, 0, t->tp.max_pos_train, t->cur[4], t->cur[5],
                           t->pcol, t->prow, TRAIN_CAP);
    fill_scale(t->psx, t->P, t->cur[4]); fill_scale(t->psy, t->P, t->cur[5]);
}
static void tracker_gen_positive(orc_tracker* t, const orc_frame* f)
{
    const int strategy = t->tp.pos_strategy, maxpos = t->tp.max_num_pos_examples;
    int lx, ty, w, h;
    t->P = 0;
    if (strategy == 0) { positive_ring(t, f); return; }                         /* SAMPLE_POS_SIMPLETRACKER :727-741 */
    if (strategy == 1 || strategy == 2) {                                       /* GREEDY :743-795, SEMI_GREEDY :797-843 */
        orc_pf_find_ordered_unique(t->pf);
        int n = t->pf->n_uniq < maxpos ? t->pf->n_uniq : maxpos;
        float ccx = t->cur[0] + t->cur[2] * t->cur[4] / 2, ccy = t->cur[1] + t->cur[3] * t->cur[5] / 2;
        for (int q = 0; q < n && t->P < TRAIN_CAP; q++) {
            const float* a = t->pf->uniq + (size_t)q * 5;
            particle_box(t, a, &lx, &ty, &w, &h);
            float dist = (float)sqrt(pow((double)(ccx - a[0]), 2) + pow((double)(ccy - a[1]), 2));
            if (a[4] != 0 && box_in_frame(lx, ty, w, h, f->W, f->H) && (strategy == 1 || dist < 8)) {
                t->pcol[t->P] = lx; t->prow[t->P] = ty; t->psx[t->P] = a[2]; t->psy[t->P] = a[3]; t->P++;
            }
        }
        if (strategy == 1 && t->P == 0) positive_ring(t, f);                    /* :779-793 fall back to the disc */
        return;
    }
    /* SAMPLE_POS_PARTICLE_RANDOM :845-892: one randfloat() per in-frame resampled particle */
    int N = t->pf->N, i = 0;
    float prob = (float)maxpos / N;
    for (int q = 0; q < N; q++) {
        const float* a = t->pf->resampled + (size_t)q * 5;
        particle_box(t, a, &lx, &ty, &w, &h);
        if (box_in_frame(lx, ty, w, h, f->W, f->H)) {
            if (orc_randfloat(t->rng) < prob) {
                if (i < TRAIN_CAP) { t->pcol[i] = lx; t->prow[i] = ty; t->psx[i] = a[2]; t->psy[i] = a[3]; }
                i++;
            }
        }
    }
    t->P = i < maxpos ? i : maxpos;
}
/* GenerateNegativeTrainingSampleSet (:913-1004) */
static void tracker_gen_negative(orc_tracker* t, const orc_frame* f)
{
    float maxs = t->pf->max_scale;
    float nsx = t->cur[4] < maxs ? t->cur[4] : maxs, nsy = t->cur[5] < maxs ? t->cur[5] : maxs;
    if (t->tp.neg_strategy == 0) {                                              /* SAMPLE_NEG_SIMPLETRACKER :921-937 */
        t->Nn = orc_sample_ring(t->rng, f->W, f->H, orc_cvround((double)t->cur[0]), orc_cvround((double)t->cur[1]),
                                orc_cvround((double)t->cur[2]), orc_cvround((double)t->cur[3]),
                                1.5f * t->tp.search_win, t->tp.pos_radius_train + 15, t->tp.n_neg_train, nsx, nsy,
                                t->ncol, t->nrow, TRAIN_CAP);
    } else {                                                                    /* SAMPLE_NEG_PARTICLE_RANDOM :939-995 */
        float mindx = 0.0f, mindy = 0.0f;
        float ccx = t->cur[0] + t->cur[2] * t->cur[4] / 2, ccy = t->cur[1] + t->cur[3] * t->cur[5] / 2;
        orc_pf_find_ordered_unique(t->pf);
        for (int q = 0; q < t->pf->n_uniq; q++) {
            const float* a = t->pf->uniq + (size_t)q * 5;
            if (fabsf(a[0] - ccx) > mindx) mindx = fabsf(a[0] - ccx);
            if (fabsf(a[1] - ccy) > mindy) mindy = fabsf(a[1] - ccy);
        }
        float lo = (float)t->tp.pos_radius_train + 5;
        if (mindx < lo) mindx = lo;
        if (mindy < lo) mindy = lo;
        if (mindx > 15.0f) mindx = 15.0f;
        if (mindy > 15.0f) mindy = 15.0f;
        float maxd = 1.5f * t->tp.search_win;
        t->Nn = orc_sample_rect(t->rng, f->W, f->H, orc_cvround((double)t->cur[0]), orc_cvround((double)t->cur[1]),
                                orc_cvround((double)t->cur[2]), orc_cvround((double)t->cur[3]),
                                maxd, maxd, mindx, mindy, t->tp.n_neg_train, nsx, nsy, t->ncol, t->nrow, TRAIN_CAP);
    }
    fill_scale(t->nsx, t->Nn, nsx); fill_scale(t->nsy, t->Nn, nsy);
}
/* UpdateClassifier :1013-1032 -> GenerateTrainingSampleSet :647-688 */
static void tracker_update(orc_tracker* t, const orc_frame* f)
{
    float a[5];
    if (t->tp.trajectory_option == 0) {
        orc_pf_find_ordered_unique(t->pf);
        for (int q = 0; q < 5; q++) a[q] = t->pf->uniq[q];          /* GetOrderedUniqueParticles(0) */
    } else {
        orc_pf_get_average(t->pf, a);
    }
    float widthFloat = a[2] * t->cur[2], heightFloat = a[3] * t->cur[3];
    float l = a[0] - widthFloat / 2, tp_ = a[1] - heightFloat / 2;
    t->cur[0] = l > 0.0f ? l : 0.0f;
    t->cur[1] = tp_ > 0.0f ? tp_ : 0.0f;
    t->cur[4] = a[2]; t->cur[5] = a[3];
    tracker_gen_positive(t, f);
    if (t->P == 0) pf_force_refinement(t->pf, t->cur[2], t->cur[3]);  /* :898-901 */
    tracker_gen_negative(t, f);
    if (t->P < 1 || t->Nn < 1) { t->ok = 0; return; }
    orc_boxes pb = {t->P, t->pcol, t->prow, t->psx, t->psy};
    orc_boxes nb = {t->Nn, t->ncol, t->nrow, t->nsx, t->nsy};
    orc_clf_update(t->clf, f, &pb, &nb, NULL, NULL);                  /* :1024 */
}
/* TrackObjectAndSaveState :239-258 */
void orc_tracker_track(orc_tracker* t, const orc_frame* f, const float* noise, int out_box[4], int phases)
{
    if (phases & 1) { tracker_score(t, f, noise); tracker_store(t, out_box); }
    if (phases & 2) tracker_update(t, f);
}

/* ------------------------------------------------------------------------------------------------
 * SimpleTracker (the reference's default local tracker, SimpleTracker.cpp): exhaustive search window.
 * Restated for the default strategies: positives = disc of radius pos_radius_train (or the single current box when the
 * radius is 1, :375-386), negatives = ring [pos_radius_train + 5, 1.5 * search_win) (m_negSampleStrategy 1, :436-447). */
struct orc_simple {
    orc_tracker_params tp;
    orc_clf* clf;
    orc_rng* rng;
    float cur[4];                 /* m_currentStateList: leftX, topY, w, h */
    int ok, last_n, last_best;
    int *col, *row; float *sx, *sy, *H;
    int *pcol, *prow, *ncol, *nrow; float* one;
    int P, Nn;
};
#define SIMPLE_CAP 16384
static void simple_update(orc_simple* t, const orc_frame* f)
{
    if (t->tp.pos_radius_train == 1) {                                     /* :375-386 */
        t->pcol[0] = (int)t->cur[0]; t->prow[0] = (int)t->cur[1]; t->P = 1;
    } else {
        t->P = orc_sample_ring(t->rng, f->W, f->H, (int)t->cur[0], (int)t->cur[1], (int)t->cur[2], (int)t->cur[3],
                               t->tp.pos_radius_train, 0, t->tp.max_pos_train, 1, 1, t->pcol, t->prow, SIMPLE_CAP);
    }
    t->Nn = orc_sample_ring(t->rng, f->W, f->H, (int)t->cur[0], (int)t->cur[1], (int)t->cur[2], (int)t->cur[3],
                            1.5f * t->tp.search_win, t->tp.pos_radius_train + 5, t->tp.n_neg_train, 1, 1,
                            t->ncol, t->nrow, SIMPLE_CAP);
    orc_boxes pb = {t->P, t->pcol, t->prow, t->one, t->one};
    orc_boxes nb = {t->Nn, t->ncol, t->nrow, t->one, t->one};
    orc_clf_update(t->clf, f, &pb, &nb, NULL, NULL);                       /* :262 (no size check in the reference) */
}
/* SimpleTracker::InitializeTrackerWithParameters :35-108 */
orc_simple* orc_simple_create(const orc_tracker_params* tp, orc_clf* clf, const orc_frame* f0, const float init[4], orc_rng* r)
{
    orc_simple* t = (orc_simple*)calloc(1, sizeof(*t));
    t->tp = *tp; t->clf = clf; t->rng = r;
    t->col = (int*)malloc(SIMPLE_CAP * 4); t->row = (int*)malloc(SIMPLE_CAP * 4);
    t->sx = (float*)malloc(SIMPLE_CAP * 4); t->sy = (float*)malloc(SIMPLE_CAP * 4); t->H = (float*)malloc(SIMPLE_CAP * 4);
    t->pcol = (int*)malloc(SIMPLE_CAP * 4); t->prow = (int*)malloc(SIMPLE_CAP * 4);
    t->ncol = (int*)malloc(SIMPLE_CAP * 4); t->nrow = (int*)malloc(SIMPLE_CAP * 4);
    t->one = (float*)malloc(SIMPLE_CAP * 4);
    for (int i = 0; i < SIMPLE_CAP; i++) { t->one[i] = 1.0f; t->sx[i] = 1.0f; t->sy[i] = 1.0f; }
    for (int i = 0; i < 4; i++) t->cur[i] = init[i];
    t->P = orc_sample_ring(r, f0->W, f0->H, (int)(unsigned)t->cur[0], (int)(unsigned)t->cur[1], (int)(unsigned)t->cur[2],
                           (int)(unsigned)t->cur[3], tp->init_pos_radius_train, 0, 1000000, 1, 1, t->pcol, t->prow, SIMPLE_CAP);
    t->Nn = orc_sample_ring(r, f0->W, f0->H, (int)(unsigned)t->cur[0], (int)(unsigned)t->cur[1], (int)(unsigned)t->cur[2],
                            (int)(unsigned)t->cur[3], 2.0f * tp->search_win, 1.5f * tp->init_pos_radius_train,
                            tp->init_n_neg_train, 1, 1, t->ncol, t->nrow, SIMPLE_CAP);
    t->ok = !(t->P < 1 || t->Nn < 1);
    if (t->ok) {
        orc_boxes pb = {t->P, t->pcol, t->prow, t->one, t->one};
        orc_boxes nb = {t->Nn, t->ncol, t->nrow, t->one, t->one};
        orc_clf_update(clf, f0, &pb, &nb, NULL, NULL);
    }
    return t;
}
void orc_simple_free(orc_simple* t)
{
    if (!t) return;
    free(t->col); free(t->row); free(t->sx); free(t->sy); free(t->H);
    free(t->pcol); free(t->prow); free(t->ncol); free(t->nrow); free(t->one); free(t);
}
/* TrackObjectAndSaveState :116-150 -> TrackObjectOnTheGivenFrame :280-342.  Returns the best response; out_state = the
   m_states row (x, y, w, h as floats). */
double orc_simple_track(orc_simple* t, const orc_frame* f, float out_state[4])
{
    /* GenerateTestSampleSet :472-493: every position within search_win of the current one */
    int n = orc_sample_ring(t->rng, f->W, f->H, (int)t->cur[0], (int)t->cur[1], (int)t->cur[2], (int)t->cur[3],
                            (float)t->tp.search_win, 0, 1000000, 1, 1, t->col, t->row, SIMPLE_CAP);
    t->last_n = n;
    if (n < 1) { t->ok = 0; return 0; }
    orc_boxes tb = {n, t->col, t->row, t->sx, t->sy};
    orc_clf_classify(t->clf, f, &tb, 1, t->H);                            /* m_shouldNotUseSigmoid = true (DefaultParameters.h:12) */
    int best = 0;                                                         /* max_idx: first maximum (Public.h:256-268) */
    for (int i = 1; i < n; i++) if (t->H[i] > t->H[best]) best = i;
    t->last_best = best;
    double resp = t->H[best];
    t->cur[1] = (float)t->row[best];
    t->cur[0] = (float)t->col[best];
    simple_update(t, f);
    for (int s = 0; s < 4; s++) out_state[s] = t->cur[s];
    return resp;
}
int orc_simple_last_n(const orc_simple* t) { return t->last_n; }
int orc_simple_last_best(const orc_simple* t) { return t->last_best; }
