// mct_host.h -- C++ host side of the B200 path: the reference's interface names over the C-ABI.
//
// The reference's seam for this path is a set of C++ classes handed around by boost::shared_ptr
// (SURVEY.md 8b).  This header keeps those names, argument meanings and call order so that the
// reference's Camera/Object/CameraNetwork code can be pointed at it:
//
//   reference (src/)                                   here
//   -------------------------------------------------  ---------------------------------------------
//   Public.{h,cpp} randint/randfloat, cvRound          Public::  (per-stream MWC, same draw order)
//   Matrixu (planes + initII/FreeII)                   Matrixu   (handle on the mct_ctx frame)
//   Classifier::Sample / SampleSet                     Classifier::Sample / SampleSet (SoA boxes)
//   Features::*Parameters, HaarFeature(Vector), ...    Features:: (Generate on host, Compute on device)
//   Classifier::StrongClassifierBase / Factory         same names; Update/Classify forward to mct_classifier_*
//   MultipleCameraTracking::ParticleFilter             same name; state lives on the device
//   MultipleCameraTracking::Tracker / ParticleFilterTracker   same names; + TrackCameraObjects (batched)
//
// Errors: the reference prints and exit(0)s (Public.h:288-308); here every failure throws
// MctHostError carrying the C-ABI code and message.  No method computes on the CPU: features,
// classifier updates/responses and particle weights only ever come from libmct_b200.so.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mct_b200.h"

typedef std::vector<float> vectorf;
typedef std::vector<int> vectori;
typedef unsigned int uint;

struct MctHostError : public std::runtime_error {
    int code;
    MctHostError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// ------------------------------------------------------------------------------------------------
namespace Public {
// OpenCV cvRNG multiply-with-carry stream (Public.h:156, Public.cpp:10-23).  The reference has one global
// stream; sharding by camera/object needs one stream per (camera, object), so the "global" functions act
// on the stream made current by the tracker (single-camera single-object order is unchanged).
struct RngStream {
    uint64_t state;
    explicit RngStream(int64_t seed = 1) { Seed(seed); }
    void Seed(int64_t seed) { state = seed ? (uint64_t)seed : (uint64_t)(int64_t)-1; }
    uint32_t Next();
    uint32_t Peek() const;
};
void        SetCurrentRng(RngStream* s);
RngStream*  CurrentRng();
void        randinitalize(const int init);
int         randint(const int min = 0, const int max = 5);
float       randfloat();
float       peekfloat();                  // value the next randfloat() will return (used for the resampler u01)
int         cvRound(double v);            // round-half-to-even
}  // namespace Public

// ------------------------------------------------------------------------------------------------
// Matrixu: the frame planes of one camera + initII()/FreeII() (Matrix.h, Matrix.cpp:33-67).  The pixels live
// on the GPU; this object is the handle Samples point at.
class Matrixu {
public:
    Matrixu(mct_ctx* ctx = nullptr) : m_ctx(ctx), m_rows(0), m_cols(0), m_iiInit(false) {}
    // size-only matrix without a device frame (host-side sample enumeration: SampleSet::SampleImage reads rows() / cols() only)
    Matrixu(int rows, int cols) : m_ctx(nullptr), m_rows(rows), m_cols(cols), m_iiInit(false) {}
    // upload gray (+ optional colour planes) and build the integral image  (Camera.cpp:460-491)
    // deviceMode: 0 = host planes (copied), 1 = device planes ordered after the work already enqueued on the ctx stream,
    //             2 = device planes that are already complete (lets the frame preparation overlap the previous frame)
    void SetFrame(const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, int cols, int rows,
                  int pitch, int deviceMode = 0);
    // the packed BGR frame the reference's Camera reads (Camera.cpp:449-494): split, conv2BW and (useHSV) conv2HSV run on the device
    void SetFrameBGR(const uint8_t* bgr, int cols, int rows, int pitchBytes, bool useHSV = false, int deviceMode = 0);
    void PrefetchFrameBGR(const uint8_t* bgr, int cols, int rows, int pitchBytes, bool useHSV = false);
    // start the upload of the NEXT frame while the current one is tracked; SetFrame with the same pointers adopts it
    void PrefetchFrame(const uint8_t* gray, const uint8_t* c0, const uint8_t* c1, const uint8_t* c2, int cols, int rows, int pitch);
    // issue the announced frame's upload and preparation now (a caller that queues the next frame while this one is tracked)
    void FlushPrefetch();
    void initII() {}                      // done inside SetFrame; kept for call-site compatibility
    void FreeII() {}
    bool isInitII() const { return m_iiInit; }
    int rows() const { return m_rows; }
    int cols() const { return m_cols; }
    float sumRect(int x, int y, int w, int h) const;      // Matrix.cpp:49-67 (probe; one D2H per call)
    mct_ctx* ctx() const { return m_ctx; }
private:
    mct_ctx* m_ctx; int m_rows, m_cols; bool m_iiInit;
};

// ------------------------------------------------------------------------------------------------
namespace Classifier {
// Sample.h:14-52
struct Sample {
    Matrixu* m_pImgGray; Matrixu* m_pImgColor; Matrixu* m_pImgHSV;
    int m_row, m_col, m_width, m_height;
    float m_weight, m_scaleX, m_scaleY;
    int m_cameraID;
    Sample() : m_pImgGray(nullptr), m_pImgColor(nullptr), m_pImgHSV(nullptr), m_row(0), m_col(0), m_width(0),
               m_height(0), m_weight(1.0f), m_scaleX(1), m_scaleY(1), m_cameraID(0) {}
};

struct ColumnRun;          // {row, first column, last column}: a run of candidate samples (mct_host.cpp)
// SampleSet.h:13-94.  Feature matrix is feature-major [F][N] (m_featureMatrix[ftr](sample)).
class SampleSet {
public:
    size_t Size() const { return m_sampleList.size(); }
    void Resize(size_t n) { m_sampleList.resize(n); }
    void Clear() { m_featureMatrix.clear(); m_numFeatures = 0; m_sampleList.clear(); }
    // samples of the column runs, thinned like SelectSamplesUniformlyFromLargerSet (mct_host.cpp)
    void TakeRuns(const std::vector<ColumnRun>& runs, int maximumNumberOfSamples, Matrixu* gray, int width, int height,
                  Matrixu* rgb, Matrixu* hsv, float scaleX, float scaleY);
    Sample& operator[](int i) { return m_sampleList[i]; }
    const Sample& operator[](int i) const { return m_sampleList[i]; }
    void PushBackSample(const Sample& s) { m_sampleList.push_back(s); }
    void PushBackSample(Matrixu* gray, int x, int y, int width = 0, int height = 0, float weight = 1.0f,
                        Matrixu* rgb = nullptr, Matrixu* hsv = nullptr, float scaleX = 1.0f, float scaleY = 1.0f);
    // ring sampler (SampleSet.cpp:82-168)
    void SampleImage(Matrixu* gray, int x, int y, int width, int height, float outerCircleRadius,
                     float innerCircleRadius = 0, int maximumNumberOfSamples = 1000000, Matrixu* rgb = nullptr,
                     Matrixu* hsv = nullptr, float scaleX = 1, float scaleY = 1);
    // two-rectangle sampler (SampleSet.cpp:180-260)
    void SampleImage(Matrixu* gray, int x, int y, int width, int height, float maximumDistanceX, float maximumDistanceY,
                     float minimumDistanceX, float minimumDistanceY, int maximumNumberOfSamples = 1000000,
                     Matrixu* rgb = nullptr, Matrixu* hsv = nullptr, float scaleX = 1, float scaleY = 1);
    // uniform random sampler (SampleSet.cpp:269-323)
    void SampleImage(Matrixu* gray, uint numberOfSamples, int w, int h, Matrixu* rgb = nullptr, Matrixu* hsv = nullptr,
                     float scaleX = 1, float scaleY = 1);
    // feature matrix
    void ResizeFeatures(size_t numFeatures);
    float& GetFeatureValue(int sample, int ftr) { return m_featureMatrix[(size_t)ftr * Size() + sample]; }
    float GetFeatureValue(int sample, int ftr) const { return m_featureMatrix[(size_t)ftr * Size() + sample]; }
    bool IsFeatureComputed() const { return !m_featureMatrix.empty() && !m_sampleList.empty(); }
    float* FeatureData() { return m_featureMatrix.data(); }
    size_t NumFeatures() const { return m_numFeatures; }
    // SoA view for the C-ABI
    void ToBoxes(vectori& col, vectori& row, vectorf& sx, vectorf& sy) const;
private:
    void SelectSamplesUniformlyFromLargerSet(int maximumNumberOfSamples);   // SampleSet.cpp:331-361
    std::vector<Sample> m_sampleList;
    std::vector<float> m_featureMatrix;
    size_t m_numFeatures = 0;
};
}  // namespace Classifier

// ------------------------------------------------------------------------------------------------
namespace Features {
enum FeatureType { HAAR_LIKE = 0, CULTURE_COLOR_HISTOGRAM = 1, MULTI_DIMENSIONAL_COLOR_HISTOGRAM = 2, HAAR_COLOR_HISTOGRAM = 3 };

// FeatureParameters.h (flattened: one struct, the type decides which fields matter)
struct FeatureParameters {
    FeatureType type = HAAR_LIKE;
    uint m_width = 0, m_height = 0;
    uint m_featureDimensionHaar = 0;                 // Tracker_Feature_Parameter
    uint m_minimumNumberOfRectangles = 2, m_maximumNumberOfRectangles = 6;   // DefaultParameters.h:14-15
    uint m_numberOfBins = 8; bool m_useHSVColorSpace = false;
    int  m_cultureColorMode = 0;                     // 0 reference_exact, 1 intended
    FeatureType GetFeatureType() const { return type; }
    uint GetHaarFeatureDimension() const { return (type == HAAR_LIKE || type == HAAR_COLOR_HISTOGRAM) ? m_featureDimensionHaar : 0; }
    uint GetColorFeatureDimension() const {
        if (type == CULTURE_COLOR_HISTOGRAM) return 11;
        if (type == MULTI_DIMENSIONAL_COLOR_HISTOGRAM || type == HAAR_COLOR_HISTOGRAM) return m_numberOfBins * m_numberOfBins * m_numberOfBins;
        return 0;
    }
    uint GetFeatureDimension() const { return GetHaarFeatureDimension() + GetColorFeatureDimension(); }
};
typedef std::shared_ptr<FeatureParameters> FeatureParametersPtr;

// HaarFeature.h: rectangles + weights, generated on the host in the reference's RNG draw order
struct HaarFeature {
    std::vector<int> m_rects;          // [n][4] x,y,w,h
    std::vector<float> m_weights;
    uint m_channel = 0;
    void Generate(const FeatureParameters& p);      // HaarFeature.cpp:25-69
};

// FeatureVector.h:28-46.  Generate() builds the host-side table; Compute() fills SampleSet's feature matrix
// on the device through the owning classifier handle.
class FeatureVector {
public:
    void Generate(FeatureParametersPtr p);
    uint GetNumberOfFeatures() const { return m_params ? m_params->GetFeatureDimension() : 0; }
    void Compute(Classifier::SampleSet& set, bool shouldResizeFeatureMatrix = true);
    const std::vector<HaarFeature>& HaarFeatures() const { return m_haar; }
    void Bind(mct_clf* clf) { m_clf = clf; }
    FeatureParametersPtr Params() const { return m_params; }
private:
    FeatureParametersPtr m_params;
    std::vector<HaarFeature> m_haar;
    mct_clf* m_clf = nullptr;
};
typedef std::shared_ptr<FeatureVector> FeatureVectorPtr;
}  // namespace Features

// ------------------------------------------------------------------------------------------------
namespace Classifier {
enum WeakClassifierType { STUMP, WEIGHTED_STUMP, PERCEPTRON };
enum StrongClassifierType { ONLINE_ADABOOST = 0, ONLINE_STOCHASTIC_BOOST_MIL = 2, ONLINE_ENSEMBLE_BOOST_MIL = 3, ONLINE_ANY_BOOST_MIL = 4 };

// ClassifierParameters.h:38-65
struct StrongClassifierParametersBase {
    StrongClassifierParametersBase(StrongClassifierType t, int numberOfSelectedWeakClassifiers, int totalNumberOfWeakClassifiers,
                                   float learningRate = 0.85f, float percentageOfRetainedWeakClassifiers = 0)
        : m_type(t), m_weakClassifierType(STUMP), m_learningRate(learningRate),
          m_numberOfSelectedWeakClassifiers(numberOfSelectedWeakClassifiers),
          m_totalNumberOfWeakClassifiers(totalNumberOfWeakClassifiers),
          m_percentageOfRetainedWeakClassifiers(percentageOfRetainedWeakClassifiers / 100.0f) {}    // ClassifierParameters.h:52
    StrongClassifierType GetClassifierType() const { return m_type; }
    StrongClassifierType m_type;
    Features::FeatureParametersPtr m_featureParametersPtr;
    WeakClassifierType m_weakClassifierType;
    float m_learningRate;
    int m_numberOfSelectedWeakClassifiers, m_totalNumberOfWeakClassifiers;
    float m_percentageOfRetainedWeakClassifiers;
};
typedef std::shared_ptr<StrongClassifierParametersBase> StrongClassifierParametersBasePtr;
struct MILBoostClassifierParameters : StrongClassifierParametersBase {
    MILBoostClassifierParameters(int k, int f) : StrongClassifierParametersBase(ONLINE_STOCHASTIC_BOOST_MIL, k, f) {}
};
struct AdaBoostClassifierParameters : StrongClassifierParametersBase {          // ClassifierParameters.h:68-85
    AdaBoostClassifierParameters(int k, int f) : StrongClassifierParametersBase(ONLINE_ADABOOST, k, f) {}
};
struct MILEnsembleClassifierParameters : StrongClassifierParametersBase {       // ClassifierParameters.h:105-117
    MILEnsembleClassifierParameters(int k, int f, float percentageOfRetainedWeakClassifiers)
        : StrongClassifierParametersBase(ONLINE_ENSEMBLE_BOOST_MIL, k, f, 0.85f, percentageOfRetainedWeakClassifiers) {}
};
struct MILAnyBoostClassifierParameters : StrongClassifierParametersBase {
    MILAnyBoostClassifierParameters(int k, int f) : StrongClassifierParametersBase(ONLINE_ANY_BOOST_MIL, k, f) {}
};

// StrongClassifierBase.h:37-66
class StrongClassifierBase {
public:
    StrongClassifierBase(mct_ctx* ctx, StrongClassifierParametersBasePtr p);
    virtual ~StrongClassifierBase();
    virtual void Update(SampleSet& positiveSampleSet, SampleSet& negativeSampleSet);
    virtual vectorf Classify(SampleSet& sampleSet, bool isLogRatioEnabled = true);
    int GetNumberOfFeatures() { return (int)m_featureVectorPtr->GetNumberOfFeatures(); }
    vectori GetSelectorList();
    mct_clf* Handle() const { return m_clf; }
    Features::FeatureVectorPtr GetFeatureVector() const { return m_featureVectorPtr; }
protected:
    mct_ctx* m_ctx;
    mct_clf* m_clf;
    StrongClassifierParametersBasePtr m_strongClassifierParametersBasePtr;
    Features::FeatureVectorPtr m_featureVectorPtr;
};
typedef std::shared_ptr<StrongClassifierBase> StrongClassifierBasePtr;
class MILBoostClassifier : public StrongClassifierBase { public: using StrongClassifierBase::StrongClassifierBase; };
class MILAnyBoostClassifier : public StrongClassifierBase { public: using StrongClassifierBase::StrongClassifierBase; };
// MILEnsembleClassifier.h: retained selectors re-ranked on the new samples, every other weak classifier trained from scratch
class MILEnsembleClassifier : public StrongClassifierBase { public: using StrongClassifierBase::StrongClassifierBase; };
// AdaBoostClassifier.h: online boosting (Oza / Grabner); GetAlphaList = m_alphaList of the selected weak classifiers
class AdaBoostClassifier : public StrongClassifierBase {
public:
    using StrongClassifierBase::StrongClassifierBase;
    vectorf GetAlphaList();
};

// StrongClassifierFactory.h:9 (+ the device context the classifier lives on)
struct StrongClassifierFactory {
    static StrongClassifierBasePtr CreateAndInitializeClassifier(StrongClassifierParametersBasePtr p, mct_ctx* ctx);
};
}  // namespace Classifier

// ------------------------------------------------------------------------------------------------
namespace MultipleCameraTracking {
// ParticleFilter.h:28-156
class ParticleFilter {
public:
    enum State { UNINTIALIZED = 0, INITIALIZED = 1, PREDICTED = 2, RESAMPLED = 3 };
    ParticleFilter(mct_ctx* ctx, float imageWidth, float imageHeight);
    ~ParticleFilter();
    void Initialize(int numberOfParticles, float centerX, float centerY, float scaleX, float scaleY,
                    float maximumScaleChangeBetweenFrame = 0.5f, float maxScale = 10, float minScale = 0.1f);
    // noise4xN: the four cv::randn(1xN, 0, sigma) rows the reference draws inside (ParticleFilter.cpp:251-301)
    void PredictWithBrownianMotion(const float* noise4xN, float width, float height);
    void UpdateAllParticlesWeight(vectorf& probability, bool shouldUpdateOnlyNonZeroWeights = true,
                                  bool shouldForceReset = false, bool shouldResample = true);
    void ResampleParticles(bool shouldForceFilterStateChange = false);
    void GetParticle(int index, vectorf& particle);
    void GetResampledParticle(int index, vectorf& particle);
    void GetOrderedUniqueParticles(int index, vectorf& particle);          // index 0 only on the fast path
    void GetHighestOrderedUniqueParticleCloseToTheGivenState(vectorf& particle, vectorf& closestState);
    void GetAverageofAllParticles(vectorf& particle);
    int GetNumberOfOrderedUniqueParticles();
    int GetNumberOfParticles() const { return m_numberOfParticles; }
    float GetMaxScale() const { return m_maxScale; }
    mct_pf* Handle() const { return m_pf; }
    // fast path: adopt the read-out the fused scoring call already produced
    void AdoptSummary(const mct_pf_summary& s) { m_summary = s; m_summaryValid = true; m_hostStateValid = false; }
    // the particles were changed on the device behind this object's back (fusers): read-outs are fetched again
    void Invalidate() { m_summaryValid = false; m_hostStateValid = false; }
    const mct_pf_summary& Summary();
    bool SummaryValid() const { return m_summaryValid; }
private:
    void FetchState();
    mct_ctx* m_ctx; mct_pf* m_pf;
    float m_imageWidth, m_imageHeight, m_maxScale;
    int m_numberOfParticles;
    mct_pf_summary m_summary; bool m_summaryValid;
    std::vector<float> m_hostState, m_hostResampled; bool m_hostStateValid;
};
typedef std::shared_ptr<ParticleFilter> ParticleFilterPtr;

// TrackerParameters.h:135-196 (only the fields the hot path reads)
enum PFTracker_TrajectoryOption { PARTICLE_HIGHEST_WEIGHT = 0, PARTICLE_AVERAGE = 1 };
struct ParticleFilterTrackerParameters {
    virtual ~ParticleFilterTrackerParameters() {}
    uint m_numberOfNegativeTrainingSamples = 15, m_init_negNumTrain = 1000;
    float m_posRadiusTrain = 1, m_init_posTrainRadius = 3;
    uint m_maximumNumberOfPositiveTrainingSamples = 100000;
    vectorf m_initState = vectorf(4, 0.0f);
    bool m_shouldNotUseSigmoid = true;
    uint m_searchWindSize = 30;
    PFTracker_TrajectoryOption m_outputTrajectoryOption = PARTICLE_AVERAGE;
    int m_numberOfParticles = 50;
    float m_standardDeviationX = 5, m_standardDeviationY = 5, m_standardDeviationScaleX = 0, m_standardDeviationScaleY = 0;
    // PFTracker_Positive_Sample_Strategy: 0 SIMPLETRACKER, 1 GREEDY, 2 SEMI_GREEDY, 3 PARTICLE_RANDOM;
    // PFTracker_Negative_Sample_Strategy: 0 SIMPLETRACKER, 1 PARTICLE_RANDOM (TrackerParameters.h:42-58; config.cfg defaults 0 / 0)
    int m_positiveSampleStrategy = 0, m_negativeSampleStrategy = 0;
    int m_maxNumPositiveExamples = 30;
};
typedef std::shared_ptr<ParticleFilterTrackerParameters> TrackerParametersPtr;

// Tracker.h:21-82 (display / replay / face-detector members are out of scope)
class Tracker {
public:
    virtual ~Tracker() {}
    virtual void InitializeTrackerWithParameters(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, int frameInd,
                                                 uint videoLength, TrackerParametersPtr trackerParametersPtr,
                                                 Classifier::StrongClassifierParametersBasePtr clfparamsPtr,
                                                 Matrixu* pFrameImageHSV = nullptr) = 0;
    virtual void TrackObjectAndSaveState(int frameind, Matrixu* pFrameImageColor, Matrixu* pFrameImageGray,
                                         Matrixu* pFrameImageHSV = nullptr) = 0;
    virtual void GenerateTrainingSampleSet(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV = nullptr) = 0;
    virtual void GenerateTestSampleSet(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV) = 0;
    virtual void GetTrainingSampleSets(Classifier::SampleSet& positiveSampleSet, Classifier::SampleSet& negativeSampleSet) = 0;
    virtual void UpdateClassifier(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV = nullptr) = 0;
    virtual const vectorf& GetCurrentTrackerState() const = 0;
    virtual void SaveStates() = 0;
};
typedef std::shared_ptr<Tracker> TrackerPtr;

// ParticleFilterTracker.h
class ParticleFilterTracker : public Tracker {
public:
    ParticleFilterTracker(mct_ctx* ctx, int64_t rngSeed = 1);
    void InitializeTrackerWithParameters(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, int frameInd, uint videoLength,
                                         TrackerParametersPtr trackerParametersPtr,
                                         Classifier::StrongClassifierParametersBasePtr clfparamsPtr,
                                         Matrixu* pFrameImageHSV = nullptr) override;
    // one object: predict -> test set -> Classify -> weights/resample -> StoreObjectState -> UpdateClassifier
    void TrackObjectAndSaveState(int frameind, Matrixu* pFrameImageColor, Matrixu* pFrameImageGray,
                                 Matrixu* pFrameImageHSV = nullptr) override;
    void GenerateTrainingSampleSet(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV = nullptr) override;
    // the three steps of GenerateTrainingSampleSet, split so that TrackCameraObjects can read the particles of all objects of a
    // camera in ONE device pass (mct_pf_training_boxes): state from the particles (:655-672), the request for the particle
    // strategies, and the sample sets themselves (:697-1004)
    void PrepareTrainingState();
    bool NeedsParticleReadout() const { return m_params->m_positiveSampleStrategy != 0 || m_params->m_negativeSampleStrategy != 0; }
    void FillTrainGen(mct_train_gen& g, Matrixu* pFrameImageColor, Matrixu* pFrameImageGray) const;
    void BuildTrainingSets(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV, const mct_train_gen_out* readout,
                           const int* col, const int* row, const float* sx, const float* sy);
    void GenerateTestSampleSet(Matrixu*, Matrixu*, Matrixu*) override {}   // fused into the device scoring call
    void GetTrainingSampleSets(Classifier::SampleSet& pos, Classifier::SampleSet& neg) override { pos = m_positiveSampleSet; neg = m_negativeSampleSet; }
    void UpdateClassifier(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV = nullptr) override;
    const vectorf& GetCurrentTrackerState() const override { return m_currentStateList; }
    void SaveStates() override {}
    // the Gaussian draws of the next PredictWithBrownianMotion (the host owns that stream)
    void SetPredictionNoise(const float* noise4xN) { m_noise = noise4xN; }
    void StoreObjectState(int frameind);                       // ParticleFilterTracker.cpp:286-393
    const std::vector<int>& States() const { return m_states; }   // [frame][4] x,y,w,h
    bool IsInitialized() const { return m_isInitialized; }
    Classifier::StrongClassifierBasePtr GetClassifier() const { return m_strongClassifierBasePtr; }
    ParticleFilterPtr GetParticleFilter() const { return m_particleFilterPtr; }
    // the tracker's stream; when the last frames ran with the stream on the device (TrackCameraObjects, whole frame on the
    // device) its state is fetched back first (one synchronisation) and host-side draws continue from there
    Public::RngStream& Rng() { PullRng(); return m_rng; }
    void PullRng();
    void PushRng();
    // whole frame on the device: default-strategy training sets generated by k_train_gen_default behind the frame's scoring
    // kernels (no host round trip).  On by default for the configurations it covers; MCT_HOST_TRAINING_SETS=1 in the
    // environment or SetDeviceTrainingSets(false) keeps the host enumeration.
    static void SetDeviceTrainingSets(bool on);
    static bool DeviceTrainingSets();
    const Classifier::SampleSet& PositiveSet() const { return m_positiveSampleSet; }
    const Classifier::SampleSet& NegativeSet() const { return m_negativeSampleSet; }

    // B200-native batched form of Camera::TrackCameraFrame's object loop (Camera.cpp:334-397): all objects of a
    // camera are scored in one set of launches (phase bit0), then their classifiers are updated (bit1).
    static void TrackCameraObjects(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, int frameind,
                                   Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV,
                                   const float* noise /* [n][4][N] */, int phases = 3);
    // resident-input variant of the score path: noise already on the device, no host synchronisation, the
    // read-out lands in the ctx's pinned buffer (mct_track_collect).  The resampler draw is consumed
    // unconditionally, so this form is for throughput runs, not for parity runs.
    static void ScoreCameraObjectsAsync(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, const float* noiseDev);
    // Streaming form of the score path (TrackObjectOnTheGivenFrame + StoreObjectState for a video whose next frame is known):
    // IssueScore queues the frame that is current on the ctx without waiting (the resampler's draw comes from the trackers'
    // device-resident streams, so nothing of the previous frame's read-out is needed), FinishScore waits for the OLDEST queued
    // frame's read-outs and stores the objects' states for it.  At most two frames in flight.
    static void IssueScore(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, const float* noise /* host [n][4][N] */);
    static void FinishScore(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers, int frameind);
    // the read-outs of all trackers whose particles changed on the device since the last one: ONE launch and ONE synchronisation
    // instead of one of each per object (in front of loops that read the particle filters object by object)
    static void EnsureSummaries(mct_ctx* ctx, std::vector<ParticleFilterTracker*>& trackers);
private:
    mct_ctx* m_ctx;
    Public::RngStream m_rng;
    bool m_rngOnDevice = false;            // m_rng is stale: the stream lives in the particle filter's device state
    bool m_isColorCulture = false;         // culture colour features: training rows take the per-object path
    TrackerParametersPtr m_params;
    Classifier::StrongClassifierBasePtr m_strongClassifierBasePtr;
    ParticleFilterPtr m_particleFilterPtr;
    vectorf m_currentStateList;            // leftX, topY, w, h, scaleX, scaleY
    Classifier::SampleSet m_positiveSampleSet, m_negativeSampleSet;
    std::vector<int> m_states;
    const float* m_noise;
    bool m_isInitialized;
};

// SimpleTracker.h: the reference's default local tracker (Local_Tracker_Type 0): exhaustive search window around the
// current position, the classifier's maximum response decides (SimpleTracker.cpp:280-342).
struct SimpleTrackerParameters : public ParticleFilterTrackerParameters {
    uint m_negSampleStrategy = 1;          // TrackerParameters.cpp:81; 0 = negatives from all over the image (:425-433)
};
typedef std::shared_ptr<SimpleTrackerParameters> SimpleTrackerParametersPtr;
class SimpleTracker : public Tracker {
public:
    SimpleTracker(mct_ctx* ctx, int64_t rngSeed = 1);
    void InitializeTrackerWithParameters(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, int frameInd, uint videoLength,
                                         TrackerParametersPtr trackerParametersPtr,
                                         Classifier::StrongClassifierParametersBasePtr clfparamsPtr,
                                         Matrixu* pFrameImageHSV = nullptr) override;
    void TrackObjectAndSaveState(int frameind, Matrixu* pFrameImageColor, Matrixu* pFrameImageGray,
                                 Matrixu* pFrameImageHSV = nullptr) override;
    void GenerateTrainingSampleSet(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV = nullptr) override;
    void GenerateTestSampleSet(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV) override;
    void GetTrainingSampleSets(Classifier::SampleSet& pos, Classifier::SampleSet& neg) override { pos = m_positiveSampleSet; neg = m_negativeSampleSet; }
    void UpdateClassifier(Matrixu* pFrameImageColor, Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV = nullptr) override;
    const vectorf& GetCurrentTrackerState() const override { return m_currentStateList; }
    void SaveStates() override {}
    const std::vector<float>& States() const { return m_states; }      // [frame][4] x, y, w, h (SimpleTracker.cpp:145-148)
    Classifier::StrongClassifierBasePtr GetClassifier() const { return m_strongClassifierBasePtr; }
    Public::RngStream& Rng() { return m_rng; }
    double LastResponse() const { return m_lastResponse; }
    int LastTestSetSize() const { return m_lastTestSetSize; }
    // all objects of a camera in one batched device pass per phase (search windows scored together, classifiers updated
    // together); responses[o] = the value TrackObjectOnTheGivenFrame returns
    static void TrackCameraObjects(mct_ctx* ctx, std::vector<SimpleTracker*>& trackers, int frameind, Matrixu* pFrameImageColor,
                                   Matrixu* pFrameImageGray, Matrixu* pFrameImageHSV, double* responses = nullptr);
private:
    mct_ctx* m_ctx;
    Public::RngStream m_rng;
    SimpleTrackerParametersPtr m_params;
    Classifier::StrongClassifierBasePtr m_strongClassifierBasePtr;
    vectorf m_currentStateList;            // leftX, topY, w, h
    Classifier::SampleSet m_positiveSampleSet, m_negativeSampleSet, m_testSampleSet;
    std::vector<float> m_states;
    double m_lastResponse; int m_lastTestSetSize;
    bool m_isInitialized;
};
}  // namespace MultipleCameraTracking
