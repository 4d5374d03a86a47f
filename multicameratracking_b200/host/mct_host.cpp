// mct_host.cpp -- host-side orchestration behind the reference's interface names (see mct_host.h).
// Everything numeric on the scoring path is delegated to libmct_b200.so; what stays here is the RNG-order
// sensitive host logic the reference also runs on the host: Haar table generation (HaarFeature.cpp:25-69),
// training-box enumeration + Bernoulli thinning (SampleSet.cpp:82-361) and the tracker's per-frame
// sequencing (ParticleFilterTracker.cpp).
#include "mct_host.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>

static void chk(mct_ctx* ctx, int rc)
{
    if (rc != MCT_OK) throw MctHostError(rc, std::string(mct_last_error(ctx)));
}

// ================================================================================ Public
namespace Public {
static RngStream g_default(1);              // static CvRNG rng_state = cvRNG(1)   (Public.h:156)
// per host thread: one thread per camera / GPU may drive its trackers concurrently with the others (threading)
static thread_local RngStream* g_cur = nullptr;

uint32_t RngStream::Next()
{
    uint64_t t = state;
    t = (uint64_t)(uint32_t)t * 4164903690U + (t >> 32);
    state = t;
    return (uint32_t)t;
}
uint32_t RngStream::Peek() const
{
    uint64_t t = (uint64_t)(uint32_t)state * 4164903690U + (state >> 32);
    return (uint32_t)t;
}
void SetCurrentRng(RngStream* s) { g_cur = s; }
RngStream* CurrentRng() { return g_cur ? g_cur : &g_default; }
void randinitalize(const int init) { CurrentRng()->Seed(init); }
int randint(const int mn, const int mx) { return (int)(CurrentRng()->Next() % (unsigned)(mx - mn + 1) + (unsigned)mn); }
float randfloat() { return (float)(CurrentRng()->Next() * 2.3283064365386962890625e-10); }
float peekfloat() { return (float)(CurrentRng()->Peek() * 2.3283064365386962890625e-10); }
int cvRound(double v) { return (int)std::lrint(v); }
}  // namespace Public
using Public::cvRound;

// ================================================================================ Matrixu
void Matrixu::SetFrame(const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, int cols, int rows,
                       int pitch, int deviceMode)
{
    chk(m_ctx, mct_frame_set(m_ctx, gray, c0, c1, c2, cols, rows, pitch, deviceMode));
    m_rows = rows; m_cols = cols; m_iiInit = true;
}
void Matrixu::SetFrameBGR(const uint8_t* bgr, int cols, int rows, int pitchBytes, bool useHSV, int deviceMode)
{
    chk(m_ctx, mct_frame_set_bgr(m_ctx, bgr, cols, rows, pitchBytes, useHSV ? 1 : 0, deviceMode));
    m_rows = rows; m_cols = cols; m_iiInit = true;
}
void Matrixu::PrefetchFrameBGR(const uint8_t* bgr, int cols, int rows, int pitchBytes, bool useHSV)
{
    chk(m_ctx, mct_frame_prefetch_bgr(m_ctx, bgr, cols, rows, pitchBytes, useHSV ? 1 : 0));
}
void Matrixu::PrefetchFrame(const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, int cols, int rows, int pitch)
{
    chk(m_ctx, mct_frame_prefetch(m_ctx, gray, c0, c1, c2, cols, rows, pitch));
}
void Matrixu::FlushPrefetch() { chk(m_ctx, mct_frame_prefetch_flush(m_ctx)); }
float Matrixu::sumRect(int x, int y, int w, int h) const
{
    int r[4] = {x, y, w, h};
    float out = 0;
    chk(m_ctx, mct_sum_rects(m_ctx, r, 1, &out));
    return out;
}

// ================================================================================ SampleSet
namespace Classifier {
void SampleSet::PushBackSample(Matrixu* gray, int x, int y, int width, int height, float weight, Matrixu* rgb,
                               Matrixu* hsv, float scaleX, float scaleY)
{
    Sample s;
    s.m_pImgGray = gray; s.m_pImgColor = rgb; s.m_pImgHSV = hsv;
    s.m_row = y; s.m_col = x; s.m_width = width; s.m_height = height;
    s.m_weight = weight; s.m_scaleX = scaleX; s.m_scaleY = scaleY;
    m_sampleList.push_back(s);
}

static void frame_limits(Matrixu* gray, Matrixu* rgb, Matrixu* hsv, int width, int height, float sx, float sy,
                         int& numberOfRows, int& numberOfColumns)
{
    Matrixu* m = gray ? gray : (rgb ? rgb : hsv);
    if (!m) throw MctHostError(MCT_E_ARG, "SampleImage: no image given");
    int scaledWidth = cvRound((double)((float)width * sx));
    int scaledHeight = cvRound((double)((float)height * sy));
    numberOfRows = m->rows() - scaledHeight - 1;
    numberOfColumns = m->cols() - scaledWidth - 1;
}

// The candidates of both region forms are runs of consecutive columns per row, so the larger set is never materialised:
// pass 1 adds up the run lengths (the keep probability of SelectSamplesUniformlyFromLargerSet needs the count), pass 2 walks
// the runs in the reference's row-major order, draws one randfloat() per candidate when the set is thinned, and stores the
// survivors only.  Same samples, same order and same RNG draws as enumerate-then-erase (SampleSet.cpp:82-168, :180-262, :331-361).
struct ColumnRun { int row, c0, c1; };
static int isqrt_floor(int v)
{
    int r = (int)std::sqrt((double)v);
    while (r * r > v) r--;
    while ((r + 1) * (r + 1) <= v) r++;
    return r;
}
void SampleSet::TakeRuns(const std::vector<ColumnRun>& runs, int maximumNumberOfSamples, Matrixu* gray, int width, int height,
                         Matrixu* rgb, Matrixu* hsv, float scaleX, float scaleY)
{
    size_t n = 0;
    for (const ColumnRun& q : runs) n += (size_t)(q.c1 - q.c0 + 1);
    m_sampleList.clear();
    const float probability = ((float)(maximumNumberOfSamples + 1)) / (float)(uint)n;
    const bool thin = probability < 1;
    Sample proto;                                   // everything but the position is common to the set
    proto.m_pImgGray = gray; proto.m_pImgColor = rgb; proto.m_pImgHSV = hsv;
    proto.m_width = width; proto.m_height = height; proto.m_weight = 1.0f; proto.m_scaleX = scaleX; proto.m_scaleY = scaleY;
    if (!thin) {
        m_sampleList.resize(n, proto);
        size_t k = 0;
        for (const ColumnRun& q : runs)
            for (int c = q.c0; c <= q.c1; c++, k++) { m_sampleList[k].m_col = c; m_sampleList[k].m_row = q.row; }
        return;
    }
    // one randfloat() per candidate, in order (Public.cpp:15-23); the stream's state lives in a register during the walk
    Public::RngStream* rng = Public::CurrentRng();
    uint64_t st = rng->state;
    m_sampleList.reserve((size_t)maximumNumberOfSamples + 64);
    for (const ColumnRun& q : runs)
        for (int c = q.c0; c <= q.c1; c++) {
            st = (uint64_t)(uint32_t)st * 4164903690U + (st >> 32);
            if ((float)((uint32_t)st * 2.3283064365386962890625e-10) > probability) continue;
            proto.m_col = c; proto.m_row = q.row;
            m_sampleList.push_back(proto);
        }
    rng->state = st;
    if (m_sampleList.size() > (size_t)maximumNumberOfSamples) m_sampleList.resize((size_t)maximumNumberOfSamples);
}

void SampleSet::SampleImage(Matrixu* gray, int x, int y, int width, int height, float outerCircleRadius,
                            float innerCircleRadius, int maximumNumberOfSamples, Matrixu* rgb, Matrixu* hsv,
                            float scaleX, float scaleY)
{
    if (!(outerCircleRadius > innerCircleRadius)) throw MctHostError(MCT_E_ARG, "SampleImage: outer radius <= inner radius");
    int nrows, ncols;
    frame_limits(gray, rgb, hsv, width, height, scaleX, scaleY, nrows, ncols);
    const float outerSq = outerCircleRadius * outerCircleRadius, innerSq = innerCircleRadius * innerCircleRadius;
    const int minrow = std::max(0, y - (int)outerCircleRadius), maxrow = std::min(nrows - 1, y + (int)outerCircleRadius);
    const int mincol = std::max(0, x - (int)outerCircleRadius), maxcol = std::min(ncols - 1, x + (int)outerCircleRadius);
    // (float)dist < outerSq && (float)dist >= innerSq for an integer dist (exact in f32 here)  <=>  dLo <= dist <= dHi
    const int dHi = (int)std::ceil(outerSq) - 1, dLo = (int)std::ceil(innerSq);
    std::vector<ColumnRun> runs;
    for (int r = minrow; r <= maxrow; r++) {
        const int dy2 = (y - r) * (y - r);
        if (dHi - dy2 < 0) continue;
        const int b = isqrt_floor(dHi - dy2);                                   // |x - c| <= b
        int a = 0;                                                              // |x - c| >= a
        if (dLo - dy2 > 0) { a = isqrt_floor(dLo - dy2 - 1) + 1; }
        if (a > b) continue;
        if (a == 0) {
            const int c0 = std::max(mincol, x - b), c1 = std::min(maxcol, x + b);
            if (c0 <= c1) runs.push_back({r, c0, c1});
        } else {
            int c0 = std::max(mincol, x - b), c1 = std::min(maxcol, x - a);
            if (c0 <= c1) runs.push_back({r, c0, c1});
            c0 = std::max(mincol, x + a); c1 = std::min(maxcol, x + b);
            if (c0 <= c1) runs.push_back({r, c0, c1});
        }
    }
    TakeRuns(runs, maximumNumberOfSamples, gray, width, height, rgb, hsv, scaleX, scaleY);
}

void SampleSet::SampleImage(Matrixu* gray, int x, int y, int width, int height, float maximumDistanceX,
                            float maximumDistanceY, float minimumDistanceX, float minimumDistanceY,
                            int maximumNumberOfSamples, Matrixu* rgb, Matrixu* hsv, float scaleX, float scaleY)
{
    if (!(maximumDistanceY > minimumDistanceY) || !(maximumDistanceX > minimumDistanceX))
        throw MctHostError(MCT_E_ARG, "SampleImage: max distance <= min distance");
    int nrows, ncols;
    frame_limits(gray, rgb, hsv, width, height, scaleX, scaleY, nrows, ncols);
    const int minrow = std::max(0, y - (int)maximumDistanceY), maxrow = std::min(nrows - 1, y + (int)maximumDistanceY);
    const int mincol = std::max(0, x - (int)maximumDistanceX), maxcol = std::min(ncols - 1, x + (int)maximumDistanceX);
    // (float)dx >= minimumDistanceX for an integer dx >= 0  <=>  dx >= aX
    const int aX = minimumDistanceX > 0 ? (int)std::ceil(minimumDistanceX) : 0;
    std::vector<ColumnRun> runs;
    for (int r = minrow; r <= maxrow; r++) {
        const int dy = std::abs(y - r);
        if ((float)dy >= minimumDistanceY || aX == 0) {
            if (mincol <= maxcol) runs.push_back({r, mincol, maxcol});
            continue;
        }
        int c0 = mincol, c1 = std::min(maxcol, x - aX);
        if (c0 <= c1) runs.push_back({r, c0, c1});
        c0 = std::max(mincol, x + aX); c1 = maxcol;
        if (c0 <= c1) runs.push_back({r, c0, c1});
    }
    TakeRuns(runs, maximumNumberOfSamples, gray, width, height, rgb, hsv, scaleX, scaleY);
}

void SampleSet::SampleImage(Matrixu* gray, uint numberOfSamples, int w, int h, Matrixu* rgb, Matrixu* hsv, float scaleX,
                            float scaleY)
{
    int nrows, ncols;
    frame_limits(gray, rgb, hsv, w, h, scaleX, scaleY, nrows, ncols);
    m_sampleList.clear();
    for (uint i = 0; i < numberOfSamples; i++) {
        const int c = Public::randint(0, ncols);
        const int r = Public::randint(0, nrows);
        PushBackSample(gray, c, r, w, h, 1.0f, rgb, hsv, scaleX, scaleY);
    }
}

void SampleSet::SelectSamplesUniformlyFromLargerSet(int maximumNumberOfSamples)
{
    const uint numberOfSamples = (uint)Size();
    const float probability = ((float)(maximumNumberOfSamples + 1)) / (float)numberOfSamples;
    if (probability < 1) {
        // one draw per candidate, in order; survivors keep their relative order
        size_t kept = 0;
        for (size_t i = 0; i < m_sampleList.size(); i++)
            if (!(Public::randfloat() > probability)) m_sampleList[kept++] = m_sampleList[i];
        m_sampleList.resize(std::min(kept, (size_t)maximumNumberOfSamples));
    }
}

void SampleSet::ResizeFeatures(size_t numFeatures)
{
    m_numFeatures = numFeatures;
    m_featureMatrix.assign(numFeatures * Size(), 0.0f);
}

void SampleSet::ToBoxes(vectori& col, vectori& row, vectorf& sx, vectorf& sy) const
{
    const size_t n = Size();
    col.resize(n); row.resize(n); sx.resize(n); sy.resize(n);
    for (size_t i = 0; i < n; i++) {
        col[i] = m_sampleList[i].m_col; row[i] = m_sampleList[i].m_row;
        sx[i] = m_sampleList[i].m_scaleX; sy[i] = m_sampleList[i].m_scaleY;
    }
}
}  // namespace Classifier

// ================================================================================ Features
namespace Features {
void HaarFeature::Generate(const FeatureParameters& p)
{
    const int n = Public::randint((int)p.m_minimumNumberOfRectangles, (int)p.m_maximumNumberOfRectangles);
    m_rects.resize((size_t)n * 4);
    m_weights.resize(n);
    for (int k = 0; k < n; k++) {
        m_weights[k] = Public::randfloat() * 2 - 1;
        const int x = Public::randint(0, (int)(p.m_width - 3));
        const int y = Public::randint(0, (int)(p.m_height - 3));
        const int w = Public::randint(1, (int)(p.m_width - x - 2));
        const int h = Public::randint(1, (int)(p.m_height - y - 2));
        m_rects[4 * k + 0] = x; m_rects[4 * k + 1] = y; m_rects[4 * k + 2] = w; m_rects[4 * k + 3] = h;
    }
    // one more draw selects the channel among the single enabled one (HaarFeature.cpp:68)
    m_channel = (uint)Public::randint(0, 0);
}

void FeatureVector::Generate(FeatureParametersPtr p)
{
    m_params = p;
    m_haar.clear();
    const uint nh = p->GetHaarFeatureDimension();
    m_haar.resize(nh);
    for (uint f = 0; f < nh; f++) m_haar[f].Generate(*p);
}

void FeatureVector::Compute(Classifier::SampleSet& set, bool shouldResizeFeatureMatrix)
{
    if (!m_clf) throw MctHostError(MCT_E_STATE, "FeatureVector::Compute: not bound to a classifier");
    if (set.Size() == 0) return;
    if (shouldResizeFeatureMatrix || set.NumFeatures() != GetNumberOfFeatures()) set.ResizeFeatures(GetNumberOfFeatures());
    vectori col, row; vectorf sx, sy;
    set.ToBoxes(col, row, sx, sy);
    mct_boxes b = {(int)set.Size(), col.data(), row.data(), sx.data(), sy.data()};
    int rc = mct_features_compute(m_clf, &b, set.FeatureData());
    if (rc) throw MctHostError(rc, "mct_features_compute failed");
}
}  // namespace Features

// ================================================================================ strong classifier
namespace Classifier {
StrongClassifierBase::StrongClassifierBase(mct_ctx* ctx, StrongClassifierParametersBasePtr p)
    : m_ctx(ctx), m_clf(nullptr), m_strongClassifierParametersBasePtr(p)
{
    if (!p || !p->m_featureParametersPtr) throw MctHostError(MCT_E_ARG, "classifier parameters missing");
    // K clamp (StrongClassifierBase.cpp:20-23)
    if (p->m_numberOfSelectedWeakClassifiers > p->m_totalNumberOfWeakClassifiers || p->m_numberOfSelectedWeakClassifiers < 1)
        p->m_numberOfSelectedWeakClassifiers = p->m_totalNumberOfWeakClassifiers / 2;
    // InitializeFeatureVector: Generate consumes the RNG exactly like the reference
    m_featureVectorPtr = Features::FeatureVectorPtr(new Features::FeatureVector());
    m_featureVectorPtr->Generate(p->m_featureParametersPtr);
    const Features::FeatureParameters& fp = *p->m_featureParametersPtr;
    mct_clf_params cp;
    memset(&cp, 0, sizeof(cp));
    cp.feature_type = (int)fp.type; cp.n_haar = (int)fp.GetHaarFeatureDimension(); cp.n_bins = (int)fp.m_numberOfBins;
    cp.w0 = (int)fp.m_width; cp.h0 = (int)fp.m_height;
    cp.weak_type = (int)p->m_weakClassifierType;
    cp.strong_type = (int)p->GetClassifierType();
    cp.n_selected = p->m_numberOfSelectedWeakClassifiers;
    cp.learning_rate = p->m_learningRate;
    cp.cch_mode = fp.m_cultureColorMode;
    cp.pct_retained = p->m_percentageOfRetainedWeakClassifiers * 100.0f;
    const auto& hf = m_featureVectorPtr->HaarFeatures();
    std::vector<int> nrect(hf.size()), rects(hf.size() * MCT_MAX_RECTS * 4, 0);
    std::vector<float> wts(hf.size() * MCT_MAX_RECTS, 0.0f);
    for (size_t f = 0; f < hf.size(); f++) {
        nrect[f] = (int)hf[f].m_weights.size();
        for (int k = 0; k < nrect[f]; k++) {
            for (int q = 0; q < 4; q++) rects[(f * MCT_MAX_RECTS + k) * 4 + q] = hf[f].m_rects[4 * k + q];
            wts[f * MCT_MAX_RECTS + k] = hf[f].m_weights[k];
        }
    }
    mct_haar_table ht = {(int)hf.size(), nrect.data(), rects.data(), wts.data()};
    chk(ctx, mct_classifier_create(ctx, &cp, hf.empty() ? nullptr : &ht, &m_clf));
    m_featureVectorPtr->Bind(m_clf);
}
StrongClassifierBase::~StrongClassifierBase() { if (m_clf) mct_classifier_destroy(m_clf); }

void StrongClassifierBase::Update(SampleSet& pos, SampleSet& neg)
{
    vectori pc, pr, nc, nr; vectorf psx, psy, nsx, nsy;
    pos.ToBoxes(pc, pr, psx, psy); neg.ToBoxes(nc, nr, nsx, nsy);
    mct_boxes pb = {(int)pos.Size(), pc.data(), pr.data(), psx.data(), psy.data()};
    mct_boxes nb = {(int)neg.Size(), nc.data(), nr.data(), nsx.data(), nsy.data()};
    chk(m_ctx, mct_classifier_update(m_clf, &pb, &nb, nullptr, nullptr));
}
vectorf StrongClassifierBase::Classify(SampleSet& set, bool isLogRatioEnabled)
{
    vectorf out(set.Size());
    if (set.Size() == 0) return out;
    vectori col, row; vectorf sx, sy;
    set.ToBoxes(col, row, sx, sy);
    mct_boxes b = {(int)set.Size(), col.data(), row.data(), sx.data(), sy.data()};
    chk(m_ctx, mct_classifier_classify(m_clf, &b, isLogRatioEnabled ? 1 : 0, out.data()));
    return out;
}
vectori StrongClassifierBase::GetSelectorList()
{
    vectori s(GetNumberOfFeatures());
    int n = 0;
    chk(m_ctx, mct_classifier_get_selectors(m_clf, s.data(), &n));
    s.resize(n);
    return s;
}

vectorf AdaBoostClassifier::GetAlphaList()
{
    vectorf a(4096); int n = 0;
    int rc = mct_classifier_get_alphas(m_clf, a.data(), &n);
    if (rc) throw MctHostError(rc, mct_last_error(m_ctx));
    a.resize(n);
    return a;
}

StrongClassifierBasePtr StrongClassifierFactory::CreateAndInitializeClassifier(StrongClassifierParametersBasePtr p, mct_ctx* ctx)
{
    if (!p) throw MctHostError(MCT_E_ARG, "CreateAndInitializeClassifier: NULL parameters");
    switch (p->GetClassifierType()) {
        case ONLINE_STOCHASTIC_BOOST_MIL: return StrongClassifierBasePtr(new MILBoostClassifier(ctx, p));
        case ONLINE_ANY_BOOST_MIL: return StrongClassifierBasePtr(new MILAnyBoostClassifier(ctx, p));
        case ONLINE_ADABOOST: return StrongClassifierBasePtr(new AdaBoostClassifier(ctx, p));
        case ONLINE_ENSEMBLE_BOOST_MIL: return StrongClassifierBasePtr(new MILEnsembleClassifier(ctx, p));
        default: throw MctHostError(MCT_E_UNSUPPORTED, "unknown strong classifier type");
    }
}
}  // namespace Classifier

// ================================================================================ ParticleFilter
namespace MultipleCameraTracking {
ParticleFilter::ParticleFilter(mct_ctx* ctx, float imageWidth, float imageHeight)
    : m_ctx(ctx), m_pf(nullptr), m_imageWidth(imageWidth), m_imageHeight(imageHeight), m_maxScale(10), m_numberOfParticles(0),
      m_summaryValid(false), m_hostStateValid(false)
{
    memset(&m_summary, 0, sizeof(m_summary));
}
ParticleFilter::~ParticleFilter() { if (m_pf) mct_pf_destroy(m_pf); }

void ParticleFilter::Initialize(int numberOfParticles, float centerX, float centerY, float scaleX, float scaleY,
                                float maximumScaleChangeBetweenFrame, float maxScale, float minScale)
{
    if (m_pf) throw MctHostError(MCT_E_STATE, "ParticleFilter::Initialize called twice");
    m_numberOfParticles = numberOfParticles; m_maxScale = maxScale;
    chk(m_ctx, mct_pf_create(m_ctx, numberOfParticles, m_imageWidth, m_imageHeight, &m_pf));
    chk(m_ctx, mct_pf_initialize(m_pf, centerX, centerY, scaleX, scaleY, maximumScaleChangeBetweenFrame, maxScale, minScale));
}
void ParticleFilter::PredictWithBrownianMotion(const float* noise4xN, float width, float height)
{
    chk(m_ctx, mct_pf_predict(m_pf, noise4xN, 0, width, height));
    m_summaryValid = false; m_hostStateValid = false;
}
void ParticleFilter::UpdateAllParticlesWeight(vectorf& probability, bool onlyNonZero, bool forceReset, bool shouldResample)
{
    int resampled = 0;
    const float u01 = Public::peekfloat();
    chk(m_ctx, mct_pf_update_all_weights(m_pf, probability.data(), (int)probability.size(), onlyNonZero, forceReset,
                                         shouldResample, u01, &resampled));
    if (resampled) (void)Public::randfloat();       // the draw ResampleParticles consumed (ParticleFilter.cpp:936)
    m_summaryValid = false; m_hostStateValid = false;
}
void ParticleFilter::ResampleParticles(bool force)
{
    chk(m_ctx, mct_pf_resample(m_pf, force ? 1 : 0, Public::randfloat()));
    m_summaryValid = false; m_hostStateValid = false;
}
const mct_pf_summary& ParticleFilter::Summary()
{
    if (!m_summaryValid) { chk(m_ctx, mct_pf_get_summary(m_pf, &m_summary)); m_summaryValid = true; }
    return m_summary;
}
void ParticleFilter::FetchState()
{
    if (m_hostStateValid) return;
    m_hostState.resize((size_t)m_numberOfParticles * 5); m_hostResampled.resize((size_t)m_numberOfParticles * 5);
    chk(m_ctx, mct_pf_get_state(m_pf, m_hostState.data()));
    chk(m_ctx, mct_pf_get_resampled(m_pf, m_hostResampled.data()));
    m_hostStateValid = true;
}
void ParticleFilter::GetParticle(int index, vectorf& particle)
{
    FetchState();
    particle.assign(m_hostState.begin() + (size_t)index * 5, m_hostState.begin() + (size_t)index * 5 + 5);
}
void ParticleFilter::GetResampledParticle(int index, vectorf& particle)
{
    FetchState();
    particle.assign(m_hostResampled.begin() + (size_t)index * 5, m_hostResampled.begin() + (size_t)index * 5 + 5);
}
void ParticleFilter::GetOrderedUniqueParticles(int index, vectorf& particle)
{
    if (index != 0) throw MctHostError(MCT_E_UNSUPPORTED, "GetOrderedUniqueParticles: only row 0 is read on the hot path");
    const mct_pf_summary& s = Summary();
    particle.assign(s.ordered_first, s.ordered_first + 5);
}
void ParticleFilter::GetHighestOrderedUniqueParticleCloseToTheGivenState(vectorf& particle, vectorf&)
{
    const mct_pf_summary& s = Summary();
    particle.assign(s.highest_close, s.highest_close + 5);
}
void ParticleFilter::GetAverageofAllParticles(vectorf& particle)
{
    const mct_pf_summary& s = Summary();
    particle.assign(s.average, s.average + 4);
}
int ParticleFilter::GetNumberOfOrderedUniqueParticles() { return Summary().n_unique; }

// ================================================================================ ParticleFilterTracker
ParticleFilterTracker::ParticleFilterTracker(mct_ctx* ctx, int64_t rngSeed)
    : m_ctx(ctx), m_rng(rngSeed), m_currentStateList(6, 0.0f), m_noise(nullptr), m_isInitialized(false)
{
}

// InitializeTracker (ParticleFilterTracker.cpp:101-231)
void ParticleFilterTracker::InitializeTrackerWithParameters(Matrixu* pColor, Matrixu* pGray, int, uint videoLength,
                                                            TrackerParametersPtr tp,
                                                            Classifier::StrongClassifierParametersBasePtr cp, Matrixu* pHSV)
{
    PullRng();
    Public::SetCurrentRng(&m_rng);
    m_params = tp;
    m_isColorCulture = cp->m_featureParametersPtr && cp->m_featureParametersPtr->GetFeatureType() == Features::CULTURE_COLOR_HISTOGRAM;
    m_states.assign((size_t)videoLength * 4, 0);
    m_strongClassifierBasePtr = Classifier::StrongClassifierFactory::CreateAndInitializeClassifier(cp, m_ctx);
    for (int i = 0; i < 4; i++) m_currentStateList[i] = tp->m_initState[i];
    m_currentStateList[4] = 1; m_currentStateList[5] = 1;
    m_positiveSampleSet.SampleImage(pGray, (int)(uint)m_currentStateList[0], (int)(uint)m_currentStateList[1],
                                    (int)(uint)m_currentStateList[2], (int)(uint)m_currentStateList[3],
                                    tp->m_init_posTrainRadius, 0, 1000000, pColor, pHSV);
    m_negativeSampleSet.SampleImage(pGray, (int)(uint)m_currentStateList[0], (int)(uint)m_currentStateList[1],
                                    (int)(uint)m_currentStateList[2], (int)(uint)m_currentStateList[3],
                                    2.0f * tp->m_searchWindSize, 1.5f * tp->m_init_posTrainRadius, (int)tp->m_init_negNumTrain,
                                    pColor, pHSV);
    if (m_positiveSampleSet.Size() < 1 || m_negativeSampleSet.Size() < 1)
        throw MctHostError(MCT_E_EMPTY, "InitializeTracker: empty initial training set");
    m_strongClassifierBasePtr->Update(m_positiveSampleSet, m_negativeSampleSet);
    m_positiveSampleSet.Clear(); m_negativeSampleSet.Clear();
    Matrixu* fr = pGray ? pGray : pColor;
    m_particleFilterPtr = ParticleFilterPtr(new ParticleFilter(m_ctx, (float)fr->cols(), (float)fr->rows()));
    m_particleFilterPtr->Initialize(tp->m_numberOfParticles, m_currentStateList[0] + m_currentStateList[2] / 2,
                                    m_currentStateList[1] + m_currentStateList[3] / 2, m_currentStateList[4],
                                    m_currentStateList[5]);
    m_isInitialized = true;
}

void ParticleFilterTracker::PullRng()
{
    if (!m_rngOnDevice) return;
    unsigned long long st = 0;
    chk(m_ctx, mct_pf_rng_get(m_particleFilterPtr->Handle(), &st));
    m_rng.state = st;
    m_rngOnDevice = false;
}
void ParticleFilterTracker::PushRng()
{
    if (m_rngOnDevice) return;
    chk(m_ctx, mct_pf_rng_set(m_particleFilterPtr->Handle(), m_rng.state));
    m_rngOnDevice = true;
}
static int g_deviceTrainingSets = -1;
void ParticleFilterTracker::SetDeviceTrainingSets(bool on) { g_deviceTrainingSets = on ? 1 : 0; }
bool ParticleFilterTracker::DeviceTrainingSets()
{
    if (g_deviceTrainingSets < 0) g_deviceTrainingSets = getenv("MCT_HOST_TRAINING_SETS") ? 0 : 1;
    return g_deviceTrainingSets != 0;
}

// StoreObjectState (ParticleFilterTracker.cpp:286-393)
void ParticleFilterTracker::StoreObjectState(int frameind)
{
    vectorf a;
    if (m_params->m_outputTrajectoryOption == PARTICLE_HIGHEST_WEIGHT) {
        vectorf prev(2);
        prev[0] = m_currentStateList[0] + m_currentStateList[2] * m_currentStateList[4] / 2;
        prev[1] = m_currentStateList[1] + m_currentStateList[3] * m_currentStateList[5] / 2;
        m_particleFilterPtr->GetHighestOrderedUniqueParticleCloseToTheGivenState(a, prev);
    } else {
        m_particleFilterPtr->GetAverageofAllParticles(a);
    }
    m_currentStateList[0] = a[0] - m_currentStateList[2] * a[2] / 2;
    m_currentStateList[1] = a[1] - m_currentStateList[3] * a[3] / 2;
    m_currentStateList[4] = a[2]; m_currentStateList[5] = a[3];
    if ((size_t)frameind * 4 + 3 < m_states.size()) {
        m_states[(size_t)frameind * 4 + 0] = cvRound((double)m_currentStateList[0]);
        m_states[(size_t)frameind * 4 + 1] = cvRound((double)m_currentStateList[1]);
        m_states[(size_t)frameind * 4 + 2] = cvRound((double)(m_currentStateList[2] * m_currentStateList[4]));
        m_states[(size_t)frameind * 4 + 3] = cvRound((double)(m_currentStateList[3] * m_currentStateList[5]));
    }
}

// GenerateTrainingSampleSet (ParticleFilterTracker.cpp:647-688): current state from the particles (:655-672) ...
void ParticleFilterTracker::PrepareTrainingState()
{
    vectorf a;
    if (m_params->m_outputTrajectoryOption == PARTICLE_HIGHEST_WEIGHT) m_particleFilterPtr->GetOrderedUniqueParticles(0, a);
    else m_particleFilterPtr->GetAverageofAllParticles(a);
    const float widthFloat = a[2] * m_currentStateList[2], heightFloat = a[3] * m_currentStateList[3];
    m_currentStateList[0] = std::max(a[0] - widthFloat / 2, 0.0f);
    m_currentStateList[1] = std::max(a[1] - heightFloat / 2, 0.0f);
    m_currentStateList[4] = a[2]; m_currentStateList[5] = a[3];
}
void ParticleFilterTracker::FillTrainGen(mct_train_gen& g, Matrixu* pColor, Matrixu* pGray) const
{
    memset(&g, 0, sizeof(g));
    g.strategy = m_params->m_positiveSampleStrategy; g.max_pos = m_params->m_maxNumPositiveExamples;
    for (int i = 0; i < 6; i++) g.cur[i] = m_currentStateList[i];
    Matrixu* m = pGray ? pGray : pColor;
    if (!m) throw MctHostError(MCT_E_ARG, "GenerateTrainingSampleSet: no image given");
    g.frame_w = m->cols(); g.frame_h = m->rows();
    g.rng_state = const_cast<ParticleFilterTracker*>(this)->Rng().state;
}
// ... then GeneratePositiveTrainingSampleSet (:697-906) and GenerateNegativeTrainingSampleSet (:913-1004).  The strategies that
// walk the particles (GREEDY, SEMI_GREEDY, PARTICLE_RANDOM positives; the distances of PARTICLE_RANDOM negatives) were evaluated
// on the device (`readout` + box arrays); the disc / ring / rectangle enumerations stay here with the tracker's RNG stream.
void ParticleFilterTracker::BuildTrainingSets(Matrixu* pColor, Matrixu* pGray, Matrixu* pHSV, const mct_train_gen_out* readout,
                                              const int* col, const int* row, const float* sx, const float* sy)
{
    PullRng();
    Public::SetCurrentRng(&m_rng);
    const int ps = m_params->m_positiveSampleStrategy, ns = m_params->m_negativeSampleStrategy;
    if (ps < 0 || ps > 3 || ns < 0 || ns > 1) throw MctHostError(MCT_E_ARG, "unknown training-sample strategy");
    if ((ps != 0 || ns != 0) && !readout) throw MctHostError(MCT_E_STATE, "BuildTrainingSets: particle read-out missing");
    m_positiveSampleSet.Clear();
    if (ps != 0) {
        for (int i = 0; i < readout->n_pos; i++)
            m_positiveSampleSet.PushBackSample(pGray, col[i], row[i], Public::cvRound((double)m_currentStateList[2]),
                                               Public::cvRound((double)m_currentStateList[3]), 1, pColor, pHSV, sx[i], sy[i]);
        m_rng.state = readout->rng_state;                              // the randfloat() draws of PARTICLE_RANDOM (:867)
    }
    if (ps == 0 || (ps == 1 && m_positiveSampleSet.Size() == 0))       // SIMPLETRACKER (:727-741) / GREEDY's fall-back (:779-793)
        m_positiveSampleSet.SampleImage(pGray, (int)m_currentStateList[0], (int)m_currentStateList[1], (int)m_currentStateList[2],
                                        (int)m_currentStateList[3], m_params->m_posRadiusTrain, 0,
                                        (int)m_params->m_maximumNumberOfPositiveTrainingSamples, pColor, pHSV,
                                        m_currentStateList[4], m_currentStateList[5]);
    const float maxs = m_particleFilterPtr->GetMaxScale();
    const float nsx = std::min(m_currentStateList[4], maxs), nsy = std::min(m_currentStateList[5], maxs);
    m_negativeSampleSet.Clear();
    if (ns == 0) {
        m_negativeSampleSet.SampleImage(pGray, cvRound((double)m_currentStateList[0]), cvRound((double)m_currentStateList[1]),
                                        cvRound((double)m_currentStateList[2]), cvRound((double)m_currentStateList[3]),
                                        1.5f * m_params->m_searchWindSize, m_params->m_posRadiusTrain + 15,
                                        (int)m_params->m_numberOfNegativeTrainingSamples, pColor, pHSV, nsx, nsy);
    } else {
        // SAMPLE_NEG_PARTICLE_RANDOM (:939-995): keep away from the spread of the unique particles, within [radius + 5, 15]
        float minimumDistanceX = std::max(readout->max_dx, (float)m_params->m_posRadiusTrain + 5);
        float minimumDistanceY = std::max(readout->max_dy, (float)m_params->m_posRadiusTrain + 5);
        minimumDistanceX = std::min(minimumDistanceX, 15.0f); minimumDistanceY = std::min(minimumDistanceY, 15.0f);
        const float maximumDistance = 1.5f * m_params->m_searchWindSize;
        m_negativeSampleSet.SampleImage(pGray, cvRound((double)m_currentStateList[0]), cvRound((double)m_currentStateList[1]),
                                        cvRound((double)m_currentStateList[2]), cvRound((double)m_currentStateList[3]),
                                        maximumDistance, maximumDistance, minimumDistanceX, minimumDistanceY,
                                        (int)m_params->m_numberOfNegativeTrainingSamples, pColor, pHSV, nsx, nsy);
    }
    if (m_positiveSampleSet.Size() < 1 || m_negativeSampleSet.Size() < 1)
        throw MctHostError(MCT_E_EMPTY, "GenerateTrainingSampleSet: empty training set");
}
void ParticleFilterTracker::GenerateTrainingSampleSet(Matrixu* pColor, Matrixu* pGray, Matrixu* pHSV)
{
    PrepareTrainingState();
    if (!NeedsParticleReadout()) { BuildTrainingSets(pColor, pGray, pHSV, nullptr, nullptr, nullptr, nullptr, nullptr); return; }
    mct_train_gen g; mct_train_gen_out go;
    FillTrainGen(g, pColor, pGray);
    const int cap = std::max(1, g.max_pos);
    vectori col(cap), row(cap); vectorf sx(cap), sy(cap);
    mct_pf* pf = m_particleFilterPtr->Handle();
    chk(m_ctx, mct_pf_training_boxes(m_ctx, 1, &pf, &g, cap, col.data(), row.data(), sx.data(), sy.data(), &go));
    BuildTrainingSets(pColor, pGray, pHSV, &go, col.data(), row.data(), sx.data(), sy.data());
}

// UpdateClassifier (ParticleFilterTracker.cpp:1013-1032)
void ParticleFilterTracker::UpdateClassifier(Matrixu* pColor, Matrixu* pGray, Matrixu* pHSV)
{
    GenerateTrainingSampleSet(pColor, pGray, pHSV);
    m_strongClassifierBasePtr->Update(m_positiveSampleSet, m_negativeSampleSet);
}

// TrackObjectAndSaveState (ParticleFilterTracker.cpp:239-258)
void ParticleFilterTracker::TrackObjectAndSaveState(int frameind, Matrixu* pColor, Matrixu* pGray, Matrixu* pHSV)
{
    std::vector<ParticleFilterTracker*> one(1, this);
    TrackCameraObjects(m_ctx, one, frameind, pColor, pGray, pHSV, m_noise, 3);
}

void ParticleFilterTracker::ScoreCameraObjectsAsync(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, const float* noiseDev)
{
    const int n = (int)trackers.size();
    std::vector<mct_pf*> pfs(n); std::vector<mct_clf*> clfs(n);
    std::vector<mct_track_params> tp(n); std::vector<float> u01(n);
    for (int o = 0; o < n; o++) {
        ParticleFilterTracker* t = trackers[o];
        pfs[o] = t->m_particleFilterPtr->Handle(); clfs[o] = t->m_strongClassifierBasePtr->Handle();
        tp[o].w0 = t->m_currentStateList[2]; tp[o].h0 = t->m_currentStateList[3];
        tp[o].exponent = 20; tp[o].force_reset = 0; tp[o].resample = 1;
        u01[o] = (float)(t->Rng().Next() * 2.3283064365386962890625e-10);
    }
    chk(ctx, mct_track_score_async(ctx, n, pfs.data(), clfs.data(), tp.data(), noiseDev, 1, u01.data()));
}

void ParticleFilterTracker::IssueScore(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, const float* noise)
{
    const int n = (int)trackers.size();
    if (n == 0) return;
    if (!noise) throw MctHostError(MCT_E_ARG, "IssueScore: prediction noise not provided");
    std::vector<mct_pf*> pfs(n); std::vector<mct_clf*> clfs(n); std::vector<mct_track_params> tp(n);
    for (int o = 0; o < n; o++) {
        ParticleFilterTracker* t = trackers[o];
        if (!t->m_isInitialized) throw MctHostError(MCT_E_STATE, "tracker not initialised");
        t->PushRng();
        pfs[o] = t->m_particleFilterPtr->Handle(); clfs[o] = t->m_strongClassifierBasePtr->Handle();
        tp[o].w0 = t->m_currentStateList[2]; tp[o].h0 = t->m_currentStateList[3];
        tp[o].exponent = 20; tp[o].force_reset = 0; tp[o].resample = 1;
    }
    chk(ctx, mct_track_score_async(ctx, n, pfs.data(), clfs.data(), tp.data(), noise, 0, nullptr));
}
void ParticleFilterTracker::FinishScore(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, int frameind)
{
    const int n = (int)trackers.size();
    if (n == 0) return;
    std::vector<mct_pf_summary> summ(n);
    chk(ctx, mct_track_collect(ctx, n, summ.data()));
    for (int o = 0; o < n; o++) {
        trackers[o]->m_particleFilterPtr->AdoptSummary(summ[o]);
        trackers[o]->StoreObjectState(frameind);
    }
}

void ParticleFilterTracker::EnsureSummaries(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers)
{
    std::vector<mct_pf*> pfs; std::vector<ParticleFilterTracker*> who;
    for (ParticleFilterTracker* t : trackers)
        if (t->m_isInitialized && !t->m_particleFilterPtr->SummaryValid()) { pfs.push_back(t->m_particleFilterPtr->Handle()); who.push_back(t); }
    if (pfs.empty()) return;
    std::vector<mct_pf_summary> summ(pfs.size());
    chk(ctx, mct_pf_get_summary_many(ctx, (int)pfs.size(), pfs.data(), summ.data()));
    for (size_t k = 0; k < who.size(); k++) who[k]->m_particleFilterPtr->AdoptSummary(summ[k]);
}

void ParticleFilterTracker::TrackCameraObjects(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, int frameind,
                                               Matrixu* pColor, Matrixu* pGray, Matrixu* pHSV, const float* noise, int phases)
{
    const int n = (int)trackers.size();
    if (n == 0) return;
    if (phases == 3 && DeviceTrainingSets()) {
        // Whole frame on the device: scoring, the default-strategy training sets (k_train_gen_default) and UpdateClassifier are
        // queued back to back; the host waits for the read-outs only (the update kernels of this frame run while the caller
        // prepares the next one).  The trackers' streams live on the device meanwhile.
        bool eligible = noise != nullptr;
        for (ParticleFilterTracker* t : trackers)
            eligible = eligible && t->m_isInitialized && !t->NeedsParticleReadout() && !t->m_isColorCulture;
        if (eligible) {
            std::vector<mct_pf*> pfs(n); std::vector<mct_clf*> clfs(n);
            std::vector<mct_track_params> tp(n); std::vector<mct_train_default> gen(n);
            std::vector<mct_pf_summary> summ(n);
            for (int o = 0; o < n; o++) {
                ParticleFilterTracker* t = trackers[o];
                t->PushRng();
                pfs[o] = t->m_particleFilterPtr->Handle(); clfs[o] = t->m_strongClassifierBasePtr->Handle();
                tp[o].w0 = t->m_currentStateList[2]; tp[o].h0 = t->m_currentStateList[3];
                tp[o].exponent = 20; tp[o].force_reset = 0; tp[o].resample = 1;
                mct_train_default& g = gen[o];
                memset(&g, 0, sizeof(g));
                g.w0 = t->m_currentStateList[2]; g.h0 = t->m_currentStateList[3];
                g.trajectory_option = (int)t->m_params->m_outputTrajectoryOption;
                g.pos_radius = t->m_params->m_posRadiusTrain; g.max_pos = (int)t->m_params->m_maximumNumberOfPositiveTrainingSamples;
                g.neg_outer = 1.5f * t->m_params->m_searchWindSize; g.neg_inner = t->m_params->m_posRadiusTrain + 15;
                g.max_neg = (int)t->m_params->m_numberOfNegativeTrainingSamples;
                g.scale_hint_x = t->m_currentStateList[4]; g.scale_hint_y = t->m_currentStateList[5];
            }
            chk(ctx, mct_track_score_async(ctx, n, pfs.data(), clfs.data(), tp.data(), noise, 0, nullptr));
            chk(ctx, mct_track_update_default(ctx, n, pfs.data(), clfs.data(), gen.data()));
            chk(ctx, mct_track_collect(ctx, n, summ.data()));
            for (int o = 0; o < n; o++) {
                ParticleFilterTracker* t = trackers[o];
                t->m_particleFilterPtr->AdoptSummary(summ[o]);
                t->StoreObjectState(frameind);
                t->PrepareTrainingState();                  // m_currentStateList as GenerateTrainingSampleSet leaves it (:655-672)
                t->m_positiveSampleSet.Clear(); t->m_negativeSampleSet.Clear();
            }
            return;
        }
    }
    if (phases & 1) {
        if (!noise) throw MctHostError(MCT_E_ARG, "TrackCameraObjects: prediction noise not provided");
        std::vector<mct_pf*> pfs(n); std::vector<mct_clf*> clfs(n);
        std::vector<mct_track_params> tp(n); std::vector<float> u01(n);
        std::vector<mct_pf_summary> summ(n);
        for (int o = 0; o < n; o++) {
            ParticleFilterTracker* t = trackers[o];
            if (!t->m_isInitialized) throw MctHostError(MCT_E_STATE, "tracker not initialised");
            pfs[o] = t->m_particleFilterPtr->Handle(); clfs[o] = t->m_strongClassifierBasePtr->Handle();
            tp[o].w0 = t->m_currentStateList[2]; tp[o].h0 = t->m_currentStateList[3];
            tp[o].exponent = 20; tp[o].force_reset = 0; tp[o].resample = 1;
            u01[o] = (float)(t->Rng().Peek() * 2.3283064365386962890625e-10);
        }
        // TrackObjectOnTheGivenFrame for every object of the camera in one batched device pass
        chk(ctx, mct_track_score(ctx, n, pfs.data(), clfs.data(), tp.data(), noise, 0, u01.data(), summ.data()));
        for (int o = 0; o < n; o++) {
            ParticleFilterTracker* t = trackers[o];
            if (summ[o].resampled) (void)t->Rng().Next();        // the randfloat() inside ResampleParticles
            t->m_particleFilterPtr->AdoptSummary(summ[o]);
            if (!(phases & 4)) t->StoreObjectState(frameind);    // bit 2: TrackObjectWithoutSaveState (fusion branch, Object.cpp:433-450)
        }
    }
    if (phases & 8) {
        // StoreAllParticleFilterTrackerState (Camera.cpp:556-574) after the fusers changed the particles on the device
        // (one read-out launch and one synchronisation for all objects of the camera)
        std::vector<mct_pf*> pfs(n); std::vector<mct_pf_summary> summ(n);
        for (int o = 0; o < n; o++) pfs[o] = trackers[o]->m_particleFilterPtr->Handle();
        chk(ctx, mct_pf_get_summary_many(ctx, n, pfs.data(), summ.data()));
        for (int o = 0; o < n; o++) {
            trackers[o]->m_particleFilterPtr->AdoptSummary(summ[o]);
            trackers[o]->StoreObjectState(frameind);
        }
    }
    if (phases & 2) {
        // UpdateClassifier: host enumerates the training boxes (RNG order), device does the rest, asynchronously
        std::vector<mct_clf*> clfs(n);
        std::vector<vectori> pc(n), pr(n), nc(n), nr(n);
        std::vector<vectorf> psx(n), psy(n), nsx(n), nsy(n);
        std::vector<mct_boxes> pb(n), nb(n);
        // MCT_HOST_TIMING=1: host time of the training-set generation and of the mct_track_update call (stderr, every 64 frames)
        static const bool timing = getenv("MCT_HOST_TIMING") != nullptr;
        static double t_gen = 0, t_upd = 0; static int t_calls = 0;
        const auto tp0 = std::chrono::steady_clock::now();
        EnsureSummaries(ctx, trackers);
        // particle-reading strategies: one device pass for all objects of the camera that ask for it
        std::vector<int> rd; std::vector<mct_pf*> rpf; std::vector<mct_train_gen> rg;
        int cap = 1;
        for (int o = 0; o < n; o++) {
            ParticleFilterTracker* t = trackers[o];
            t->PrepareTrainingState();
            if (!t->NeedsParticleReadout()) continue;
            mct_train_gen g; t->FillTrainGen(g, pColor, pGray);
            rd.push_back(o); rpf.push_back(t->m_particleFilterPtr->Handle()); rg.push_back(g);
            cap = std::max(cap, g.max_pos);
        }
        std::vector<mct_train_gen_out> ro(rd.size());
        vectori rcol(rd.size() * cap), rrow(rd.size() * cap); vectorf rsx(rd.size() * cap), rsy(rd.size() * cap);
        if (!rd.empty())
            chk(ctx, mct_pf_training_boxes(ctx, (int)rd.size(), rpf.data(), rg.data(), cap, rcol.data(), rrow.data(), rsx.data(), rsy.data(), ro.data()));
        size_t k = 0;
        for (int o = 0; o < n; o++) {
            ParticleFilterTracker* t = trackers[o];
            if (k < rd.size() && rd[k] == o) {
                t->BuildTrainingSets(pColor, pGray, pHSV, &ro[k], rcol.data() + k * cap, rrow.data() + k * cap, rsx.data() + k * cap, rsy.data() + k * cap);
                k++;
            } else {
                t->BuildTrainingSets(pColor, pGray, pHSV, nullptr, nullptr, nullptr, nullptr, nullptr);
            }
            t->m_positiveSampleSet.ToBoxes(pc[o], pr[o], psx[o], psy[o]);
            t->m_negativeSampleSet.ToBoxes(nc[o], nr[o], nsx[o], nsy[o]);
            pb[o] = {(int)pc[o].size(), pc[o].data(), pr[o].data(), psx[o].data(), psy[o].data()};
            nb[o] = {(int)nc[o].size(), nc[o].data(), nr[o].data(), nsx[o].data(), nsy[o].data()};
            clfs[o] = t->m_strongClassifierBasePtr->Handle();
        }
        const auto tp1 = std::chrono::steady_clock::now();
        chk(ctx, mct_track_update(ctx, n, clfs.data(), pb.data(), nb.data()));
        if (timing) {
            const auto tp2 = std::chrono::steady_clock::now();
            t_gen += std::chrono::duration<double, std::micro>(tp1 - tp0).count();
            t_upd += std::chrono::duration<double, std::micro>(tp2 - tp1).count();
            if (++t_calls == 64) {
                fprintf(stderr, "UpdateClassifier host side: training sets %.1f us, mct_track_update (plan + copies + launches) %.1f us per frame\n",
                        t_gen / t_calls, t_upd / t_calls);
                t_gen = t_upd = 0; t_calls = 0;
            }
        }
    }
}

// ================================================================================ SimpleTracker
SimpleTracker::SimpleTracker(mct_ctx* ctx, int64_t rngSeed)
    : m_ctx(ctx), m_rng(rngSeed), m_currentStateList(4, 0.0f), m_lastResponse(0), m_lastTestSetSize(0), m_isInitialized(false)
{
}

// SimpleTracker.cpp:35-108
void SimpleTracker::InitializeTrackerWithParameters(Matrixu* pColor, Matrixu* pGray, int, uint videoLength, TrackerParametersPtr tp,
                                                    Classifier::StrongClassifierParametersBasePtr cp, Matrixu* pHSV)
{
    Public::SetCurrentRng(&m_rng);
    m_params = std::dynamic_pointer_cast<SimpleTrackerParameters>(tp);
    if (!m_params) { m_params = SimpleTrackerParametersPtr(new SimpleTrackerParameters()); static_cast<ParticleFilterTrackerParameters&>(*m_params) = *tp; }
    m_states.assign((size_t)videoLength * 4, 0.0f);
    m_strongClassifierBasePtr = Classifier::StrongClassifierFactory::CreateAndInitializeClassifier(cp, m_ctx);
    for (int i = 0; i < 4; i++) m_currentStateList[i] = tp->m_initState[i];
    Classifier::SampleSet pos, neg;
    pos.SampleImage(pGray, (int)(uint)m_currentStateList[0], (int)(uint)m_currentStateList[1], (int)(uint)m_currentStateList[2],
                    (int)(uint)m_currentStateList[3], tp->m_init_posTrainRadius, 0, 1000000, pColor, pHSV);
    neg.SampleImage(pGray, (int)(uint)m_currentStateList[0], (int)(uint)m_currentStateList[1], (int)(uint)m_currentStateList[2],
                    (int)(uint)m_currentStateList[3], 2.0f * tp->m_searchWindSize, 1.5f * tp->m_init_posTrainRadius,
                    (int)tp->m_init_negNumTrain, pColor, pHSV);
    if (pos.Size() < 1 || neg.Size() < 1) throw MctHostError(MCT_E_EMPTY, "SimpleTracker: empty initial training set");
    m_strongClassifierBasePtr->Update(pos, neg);
    m_isInitialized = true;
}

// :472-493
void SimpleTracker::GenerateTestSampleSet(Matrixu* pColor, Matrixu* pGray, Matrixu* pHSV)
{
    Public::SetCurrentRng(&m_rng);
    m_testSampleSet.Clear();
    m_testSampleSet.SampleImage(pGray, (int)m_currentStateList[0], (int)m_currentStateList[1], (int)m_currentStateList[2],
                                (int)m_currentStateList[3], (float)m_params->m_searchWindSize, 0, 1000000, pColor, pHSV);
}

// :350-455
void SimpleTracker::GenerateTrainingSampleSet(Matrixu* pColor, Matrixu* pGray, Matrixu* pHSV)
{
    Public::SetCurrentRng(&m_rng);
    m_positiveSampleSet.Clear();
    if (m_params->m_posRadiusTrain == 1) {
        Classifier::Sample smp;
        smp.m_col = (int)m_currentStateList[0]; smp.m_row = (int)m_currentStateList[1];
        smp.m_width = (int)m_currentStateList[2]; smp.m_height = (int)m_currentStateList[3];
        smp.m_scaleX = 1; smp.m_scaleY = 1; smp.m_weight = 1.0f;
        m_positiveSampleSet.PushBackSample(smp);
    } else {
        m_positiveSampleSet.SampleImage(pGray, (int)m_currentStateList[0], (int)m_currentStateList[1], (int)m_currentStateList[2],
                                        (int)m_currentStateList[3], m_params->m_posRadiusTrain, 0,
                                        (int)m_params->m_maximumNumberOfPositiveTrainingSamples, pColor, pHSV);
    }
    m_negativeSampleSet.Clear();
    if (m_params->m_negSampleStrategy == 0)
        m_negativeSampleSet.SampleImage(pGray, m_params->m_numberOfNegativeTrainingSamples, (int)m_currentStateList[2],
                                        (int)m_currentStateList[3], pColor, pHSV);
    else
        m_negativeSampleSet.SampleImage(pGray, (int)m_currentStateList[0], (int)m_currentStateList[1], (int)m_currentStateList[2],
                                        (int)m_currentStateList[3], 1.5f * m_params->m_searchWindSize, m_params->m_posRadiusTrain + 5,
                                        (int)m_params->m_numberOfNegativeTrainingSamples, pColor, pHSV);
}

// :258-270
void SimpleTracker::UpdateClassifier(Matrixu* pColor, Matrixu* pGray, Matrixu* pHSV)
{
    GenerateTrainingSampleSet(pColor, pGray, pHSV);
    m_strongClassifierBasePtr->Update(m_positiveSampleSet, m_negativeSampleSet);
}

void SimpleTracker::TrackObjectAndSaveState(int frameind, Matrixu* pColor, Matrixu* pGray, Matrixu* pHSV)
{
    std::vector<SimpleTracker*> one(1, this);
    TrackCameraObjects(m_ctx, one, frameind, pColor, pGray, pHSV, nullptr);
}

void SimpleTracker::TrackCameraObjects(mct_ctx* ctx, std::vector<SimpleTracker*>& trackers, int frameind, Matrixu* pColor,
                                       Matrixu* pGray, Matrixu* pHSV, double* responses)
{
    const int n = (int)trackers.size();
    if (n == 0) return;
    // 1. search windows (host enumeration, each tracker's own RNG stream), scored in one batched device pass
    std::vector<mct_clf*> clfs(n);
    std::vector<vectori> tc(n), tr(n); std::vector<vectorf> tsx(n), tsy(n);
    std::vector<mct_boxes> tb(n);
    for (int o = 0; o < n; o++) {
        SimpleTracker* t = trackers[o];
        if (!t->m_isInitialized) throw MctHostError(MCT_E_STATE, "tracker not initialised");
        t->GenerateTestSampleSet(pColor, pGray, pHSV);
        t->m_testSampleSet.ToBoxes(tc[o], tr[o], tsx[o], tsy[o]);
        t->m_lastTestSetSize = (int)tc[o].size();
        tb[o] = {(int)tc[o].size(), tc[o].data(), tr[o].data(), tsx[o].data(), tsy[o].data()};
        clfs[o] = t->m_strongClassifierBasePtr->Handle();
    }
    std::vector<int> best(n); std::vector<float> resp(n);
    chk(ctx, mct_classify_argmax(ctx, n, clfs.data(), tb.data(), 1 /* m_shouldNotUseSigmoid */, best.data(), resp.data()));
    // 2. new position = the best box (:316-323), then UpdateClassifier for all objects in one batched pass
    std::vector<vectori> pc(n), pr(n), nc(n), nr(n);
    std::vector<vectorf> psx(n), psy(n), nsx(n), nsy(n);
    std::vector<mct_boxes> pb(n), nb(n);
    for (int o = 0; o < n; o++) {
        SimpleTracker* t = trackers[o];
        if (best[o] < 0) throw MctHostError(MCT_E_EMPTY, "SimpleTracker: no response");
        t->m_lastResponse = resp[o];
        if (responses) responses[o] = resp[o];
        t->m_currentStateList[1] = (float)tr[o][best[o]];
        t->m_currentStateList[0] = (float)tc[o][best[o]];
        t->m_testSampleSet.Clear();
        t->GenerateTrainingSampleSet(pColor, pGray, pHSV);
        t->m_positiveSampleSet.ToBoxes(pc[o], pr[o], psx[o], psy[o]);
        t->m_negativeSampleSet.ToBoxes(nc[o], nr[o], nsx[o], nsy[o]);
        pb[o] = {(int)pc[o].size(), pc[o].data(), pr[o].data(), psx[o].data(), psy[o].data()};
        nb[o] = {(int)nc[o].size(), nc[o].data(), nr[o].data(), nsx[o].data(), nsy[o].data()};
    }
    chk(ctx, mct_track_update(ctx, n, clfs.data(), pb.data(), nb.data()));
    for (int o = 0; o < n; o++) {
        SimpleTracker* t = trackers[o];
        if ((size_t)frameind * 4 + 3 < t->m_states.size())
            for (int s = 0; s < 4; s++) t->m_states[(size_t)frameind * 4 + s] = t->m_currentStateList[s];
    }
}
}  // namespace MultipleCameraTracking
