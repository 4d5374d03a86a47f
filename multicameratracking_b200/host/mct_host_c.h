// mct_host_c.h -- the per-camera state behind the flat C driver (mct_host_c.cpp) and the camera network (mct_network.cpp).
#pragma once
#include <string>
#include <vector>

#include "mct_host.h"
#include "mct_fuser.h"

struct mcth_object_params {
    int feature_type, n_haar, n_bins, weak_type, strong_type;
    int percent_selected;             // Percentage_Of_Weak_Classifiers_Selected (Object.cpp:235-236)
    int n_particles;
    float sd_x, sd_y, sd_sx, sd_sy;
    float pos_radius_train, init_pos_radius_train;
    int n_neg_train, init_n_neg_train, search_win;
    int trajectory_option;
    int64_t rng_seed;
    float init_xywh[4];
    float percent_retained;           // Percentage_Of_Weak_Classifiers_Retained (MILEnsemble, Object.cpp:254-262)
    int pos_strategy, neg_strategy;   // PfTracker_Positive / Negative_Sample_Strategy (Object.cpp:316-342)
    int max_num_pos_examples;         // PfTracker_Max_Num_Positive_Examples (Object.cpp:313)
};

struct mcth_camera {
    mct_ctx* ctx;
    Matrixu frame;
    std::vector<MultipleCameraTracking::ParticleFilterTracker*> trackers;
    std::vector<mcth_object_params> params;
    std::vector<MultipleCameraTracking::SimpleTracker*> simple;                 // Local_Tracker_Type 0 objects (their own list and params)
    std::vector<mcth_object_params> simple_params;
    int n_frames;
    std::string err;
    int stream_pending = 0;           // streaming score path (mcth_camera_step, phases bit 4): a frame is queued ahead, identified
    const void* stream_gray = nullptr; const void* stream_noise = nullptr;   // by the host buffers it was announced with
};

