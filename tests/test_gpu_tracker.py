"""End-to-end parity: the C++ host mirror (ParticleFilterTracker over the C-ABI) against the oracle's
restatement of the reference tracker loop on the same synthetic sequence, same RNG streams and same
prediction noise.  Bar (north_star): tracked boxes within 1 px."""
import ctypes as C

import numpy as np
import pytest

from multicameratracking_b200 import capi, synth
from multicameratracking_b200 import host as mhost
from oracle import oracle_py

pytestmark = pytest.mark.gpu


def _oracle_tracker(oracle, seq, o, f0, p, seed):
    rng = oracle.rng(seed)
    t = oracle.haar_generate(rng, p["n_haar"], seq.w0, seq.h0) if p["feature_type"] in (0, 3) else None
    F = {0: p["n_haar"], 1: 11, 2: 512, 3: p["n_haar"] + 512}[p["feature_type"]]
    K = int(F * p["percent_selected"] / 100)
    clf = oracle.clf_create(p["feature_type"], p["n_haar"], p["n_bins"], seq.w0, seq.h0, p["weak_type"], p["strong_type"], K, 0.85, t,
                            pct_retained=p.get("percent_retained", 0.0))
    tp = oracle_py.TrackerParams(p["n_particles"], p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"], p["pos_radius_train"],
                                 p["init_pos_radius_train"], p["n_neg_train"], p["init_n_neg_train"], 100000,
                                 p["search_win"], p["trajectory_option"], p.get("pos_strategy", 0), p.get("neg_strategy", 0),
                                 p.get("max_num_pos_examples", 30))
    x, y, w, h = seq.box(0, o)
    init = np.array([x, y, w, h], np.float32)
    trk = oracle.lib.orc_tracker_create(C.byref(tp), clf, C.byref(f0), init.ctypes.data_as(C.POINTER(C.c_float)), C.byref(rng))
    return trk, clf, rng


def _oracle_last_train(oracle, trk, which, cap=8192):
    col = np.zeros(cap, np.int32); row = np.zeros(cap, np.int32); sx = np.zeros(cap, np.float32); sy = np.zeros(cap, np.float32)
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
    n = oracle.lib.orc_tracker_last_train(trk, which, col.ctypes.data_as(ip), row.ctypes.data_as(ip), sx.ctypes.data_as(fp), sy.ctypes.data_as(fp), cap)
    return col[:n], row[:n], sx[:n], sy[:n]


@pytest.mark.parametrize("cfg", ["cfg1_haar_milboost", "haarmdch_wstump_2obj", "mdch_anyboost_avg", "cfg2_full_size_8obj_2000",
                                 "haar_ensemble_2obj", "greedy_pos_random_neg", "semi_greedy_pos", "particle_random_pos"])
def test_tracker_sequence_matches_oracle(oracle, cfg):
    _tracker_sequence(oracle, cfg)


@pytest.mark.parametrize("cfg", ["cfg1_haar_milboost", "haarmdch_wstump_2obj", "mdch_anyboost_avg"])
def test_tracker_sequence_matches_oracle_with_host_training_sets(oracle, cfg):
    """the host enumeration of the default-strategy training sets (the path of round 1) stays covered: same sequences, with the
    device-side generation switched off"""
    mhost.set_device_training_sets(False)
    try:
        _tracker_sequence(oracle, cfg)
    finally:
        mhost.set_device_training_sets(True)


def test_device_training_sets_with_a_wrong_scale_hint_take_the_generic_kernel(oracle, monkeypatch):
    """mct_track_update_default sizes its launches from the last known state scale; a plan that leaves those bounds must cost
    time only: with the hint halved the pixel-list plans of the colour rows do not fit, every object takes the generic per-box
    kernel (mdch_fast == 0 in the read-out), and boxes, stream states and selector lists still equal the oracle's"""
    monkeypatch.setenv("MCT_TEST_HINT_SCALE", "0.5")
    fast = _tracker_sequence(oracle, "haarmdch_wstump_2obj", want_fast_flags=True)
    assert fast and not any(fast), fast


def _tracker_sequence(oracle, cfg, want_fast_flags=False):
    n_frames = 12
    fast_flags = []
    device_sets = mhost.device_training_sets() and cfg not in ("greedy_pos_random_neg", "semi_greedy_pos", "particle_random_pos")
    if cfg == "cfg1_haar_milboost":
        n_obj, over = 1, dict()
    elif cfg == "cfg2_full_size_8obj_2000":
        # BASELINE configs[1] at its full size: 8 objects, Haar(250) + MDCH(512) -> 152 selected, WeightedStumps + MILBoost,
        # 2000 particles per object (the oracle needs ~1 s per frame for it)
        n_frames = 6
        n_obj, over = 8, dict(feature_type=3, weak_type=1, n_particles=2000, n_haar=250)
    elif cfg == "haarmdch_wstump_2obj":
        n_obj, over = 2, dict(feature_type=3, weak_type=1, n_particles=600, n_haar=100)
    elif cfg == "greedy_pos_random_neg":
        # PfTracker_Positive_Sample_Strategy 1 (ordered unique particles) + Negative 1 (rectangle ring sized by the particle spread)
        n_obj, over = 2, dict(n_particles=600, n_haar=100, pos_strategy=1, neg_strategy=1, max_num_pos_examples=60)
    elif cfg == "semi_greedy_pos":
        n_obj, over = 1, dict(n_particles=600, n_haar=100, pos_strategy=2, max_num_pos_examples=300, sd_x=4.0, sd_y=4.0)
    elif cfg == "particle_random_pos":
        # one randfloat() per in-frame resampled particle: the device jumps the tracker's RNG stream ahead
        n_obj, over = 2, dict(n_particles=800, n_haar=100, pos_strategy=3, neg_strategy=1, max_num_pos_examples=50, trajectory_option=1)
    elif cfg == "haar_ensemble_2obj":
        # Tracker_Strong_Classifier_Type 3 (MILEnsemble), half of the previous selectors re-ranked and retained every frame
        n_obj, over = 2, dict(strong_type=3, n_particles=500, n_haar=120, percent_retained=50.0)
    else:
        n_obj, over = 1, dict(feature_type=2, strong_type=4, n_particles=500, trajectory_option=1)
    seq = synth.Sequence(640, 480, n_obj, n_frames=n_frames)
    p = dict(mhost.DEFAULTS); p.update(over)
    cam = mhost.Camera(0, n_frames)
    gray, r, g, b = seq.frame(0)
    cam.set_frame(gray, r, g, b)
    f0 = oracle.frame(gray, r, g, b)
    otr = []
    for o in range(n_obj):
        cam.add_object(seq.box(0, o), rng_seed=o + 1, **over)
        otr.append(_oracle_tracker(oracle, seq, o, f0, p, o + 1))
    cam.init_trackers()
    for o in range(n_obj):
        oclf = otr[o][1]
        assert cam.selectors(o).tolist() == oracle.clf_selectors(oclf).tolist()
        assert cam.rng_state(o) == otr[o][2].state
    N = p["n_particles"]
    worst = 0
    for t in range(1, n_frames):
        gray, r, g, b = seq.frame(t)
        cam.set_frame(gray, r, g, b)
        of = oracle.frame(gray, r, g, b)
        nz = np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), obj=o, t=t) for o in range(n_obj)])
        boxes = cam.track(t, nz, 3)
        cam.sync()
        for o in range(n_obj):
            ob = np.zeros(4, np.int32)
            oracle.lib.orc_tracker_track(otr[o][0], C.byref(of), nz[o].ctypes.data_as(C.POINTER(C.c_float)),
                                         ob.ctypes.data_as(C.POINTER(C.c_int)), 3)
            d = np.abs(boxes[o] - ob).max()
            worst = max(worst, int(d))
            assert d <= 1, (cfg, t, o, boxes[o], ob)
            # the RNG streams stay in lock-step (same resample decisions, same thinning draws)
            assert cam.rng_state(o) == otr[o][2].state, (cfg, t, o)
            assert cam.selectors(o).tolist() == oracle.clf_selectors(otr[o][1]).tolist(), (cfg, t, o)
            if device_sets:
                # the boxes k_train_gen_default wrote (positives, then negatives) against the oracle's SampleImage + thinning
                col, row, sx, sy, P, Nn = capi.track_update_default_boxes(cam.capi_context(), o)
                if want_fast_flags:
                    cctx = cam.capi_context()
                    outs = (capi.TrainDefaultOut * n_obj)()
                    cctx.check(cctx.L.mct_track_update_default_collect(cctx.h, n_obj, outs))
                    fast_flags.append(int(outs[o].mdch_fast))
                oc, orw, osx, osy = _oracle_last_train(oracle, otr[o][0], 0)
                nc, nr, nsx, nsy = _oracle_last_train(oracle, otr[o][0], 1)
                assert (P, Nn) == (len(oc), len(nc)), (cfg, t, o, P, Nn, len(oc), len(nc))
                np.testing.assert_array_equal(col, np.concatenate([oc, nc])); np.testing.assert_array_equal(row, np.concatenate([orw, nr]))
                # scales = the state's scale as the read-out of this frame reports it (exactly), which agrees with the oracle's to
                # 1e-6: PARTICLE_AVERAGE is a f64 tree against the reference's f32 chain, and PARTICLE_HIGHEST_WEIGHT picks between
                # particles of the same pixel cell whose weights agree to the response tolerance (their scales differ by an ulp)
                st = cam.state(o)
                np.testing.assert_array_equal(sx[:P], np.full(P, st[4], np.float32)); np.testing.assert_array_equal(sy[:P], np.full(P, st[5], np.float32))
                np.testing.assert_allclose(sx, np.concatenate([osx, nsx]), rtol=1e-6, atol=0); np.testing.assert_allclose(sy, np.concatenate([osy, nsy]), rtol=1e-6, atol=0)
            # sanity only: the tracker (reference algorithm) stays in the neighbourhood of the synthetic object
            gx, gy, gw, gh = seq.box(t, o)
            assert abs(boxes[o][0] - gx) <= 60 and abs(boxes[o][1] - gy) <= 60
    for trk, clf, _ in otr:
        oracle.lib.orc_tracker_free(trk); oracle.lib.orc_clf_free(clf)
    cam.close()
    return fast_flags

def test_single_call_frame_step_equals_separate_calls():
    """mcth_camera_step (set frame + announce the next frame and noise + track, one call per video frame) gives the same boxes,
    selectors and RNG states as set_frame / track"""
    n_frames, n_obj = 8, 2
    seq = synth.Sequence(640, 480, n_obj, n_frames=n_frames)
    over = dict(feature_type=3, weak_type=1, n_particles=500, n_haar=60)
    p = dict(mhost.DEFAULTS); p.update(over)
    cams = []
    for _ in range(2):
        cam = mhost.Camera(0, n_frames)
        cam.set_frame(*seq.frame(0))
        for o in range(n_obj):
            cam.add_object(seq.box(0, o), rng_seed=o + 1, **over)
        cam.init_trackers()
        cams.append(cam)
    frames = [tuple(np.ascontiguousarray(x) for x in seq.frame(t)) for t in range(n_frames)]
    noise = [np.stack([synth.noise(p["n_particles"], (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), obj=o, t=t) for o in range(n_obj)])
             for t in range(n_frames)]
    for t in range(1, n_frames):
        cams[0].set_frame(*frames[t])
        a = cams[0].track(t, noise[t], 3)
        nxt = t + 1 < n_frames
        b = cams[1].step(t, frames[t], frames[t + 1] if nxt else None, noise[t], noise[t + 1] if nxt else None, 3)
        np.testing.assert_array_equal(a, b)
        for o in range(n_obj):
            assert cams[0].rng_state(o) == cams[1].rng_state(o)
            assert cams[0].selectors(o).tolist() == cams[1].selectors(o).tolist()
    for cam in cams:
        cam.close()



def test_streaming_score_steps_equal_synchronous_steps():
    """mcth_camera_step with phases = 1 | 16 queues frame t + 1 before it returns frame t's boxes (the trackers' streams are
    device-resident, so nothing of frame t's read-out is needed for that): same boxes per frame, same particles, same stream states
    as the synchronous calls -- over a sequence with resampling and non-resampling frames, and across a run that is closed and
    reopened"""
    n_frames, n_obj = 10, 3
    seq = synth.Sequence(640, 480, n_obj, n_frames=n_frames)
    over = dict(feature_type=3, weak_type=1, n_particles=700, n_haar=80)
    p = dict(mhost.DEFAULTS); p.update(over)
    cams = []
    for _ in range(2):
        cam = mhost.Camera(0, n_frames)
        cam.set_frame(*seq.frame(0))
        for o in range(n_obj):
            cam.add_object(seq.box(0, o), rng_seed=o + 1, **over)
        cam.init_trackers()
        cams.append(cam)
    frames = [tuple(np.ascontiguousarray(x) for x in seq.frame(t)) for t in range(n_frames)]
    noise = [np.stack([synth.noise(p["n_particles"], (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), obj=o, t=t) for o in range(n_obj)])
             for t in range(n_frames)]
    kinds = set()
    for t in range(1, n_frames):
        cams[0].set_frame(*frames[t])
        a = cams[0].track(t, noise[t], 1)
        # the streaming run is closed after frame 5 (nothing announced) and reopened at frame 6
        nxt = t + 1 < n_frames and t != 5
        b = cams[1].step(t, frames[t], frames[t + 1] if nxt else None, noise[t], noise[t + 1] if nxt else None, 1 | 16)
        np.testing.assert_array_equal(a, b)
        pa = cams[0].particles(0)
        kinds.add(bool((pa[:, 4] == 1).all()))
    for o in range(n_obj):
        np.testing.assert_array_equal(cams[0].particles(o), cams[1].particles(o))
        assert cams[0].rng_state(o) == cams[1].rng_state(o)
    assert len(kinds) == 2 or True in kinds
    for cam in cams:
        cam.close()


def test_free_running_score_path_equals_track_up_to_the_unconditional_draw():
    """ScoreCameraObjectsAsync (the free-running, device-resident leg bench.py times as `value`) draws the resampler's randfloat()
    for every frame without waiting for the read-out; TrackCameraObjects draws it only when the frame resampled.  That draw is
    the ONLY difference: with one extra randfloat() consumed on the synchronous side after every frame that did not resample,
    the two paths agree bit for bit (particles and RNG state) over a sequence with both kinds of frames."""
    import torch
    n_frames, n_obj = 7, 3
    seq = synth.Sequence(640, 480, n_obj, n_frames=n_frames)
    over = dict(feature_type=3, weak_type=1, n_particles=700, n_haar=80)
    p = dict(mhost.DEFAULTS); p.update(over)
    cams = []
    for _ in range(2):
        cam = mhost.Camera(0, n_frames)
        cam.set_frame(*seq.frame(0))
        for o in range(n_obj):
            cam.add_object(seq.box(0, o), rng_seed=o + 1, **over)
        cam.init_trackers()
        cams.append(cam)
    N = p["n_particles"]
    kinds = set()
    for t in range(1, n_frames):
        fr = [np.ascontiguousarray(x) for x in seq.frame(t)]
        nz = np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), obj=o, t=t) for o in range(n_obj)])
        cams[0].set_frame(*fr)
        cams[0].track(t, nz, 1)
        cams[1].set_frame(*fr)
        nzd = torch.from_numpy(nz).cuda()
        torch.cuda.synchronize()
        cams[1]._chk(cams[1].L.mcth_camera_track_resident(cams[1].h, t, C.c_void_p(nzd.data_ptr())))
        cams[1].sync()
        for o in range(n_obj):
            a, b = cams[0].particles(o), cams[1].particles(o)
            np.testing.assert_array_equal(a, b)
            resampled = bool((a[:, 4] == 1).all())
            kinds.add(resampled)
            if not resampled:
                cams[0].randfloat(o)                     # the draw the free-running path spent on this frame anyway
            assert cams[0].rng_state(o) == cams[1].rng_state(o), (t, o)
    assert True in kinds, "no frame resampled: the test does not cover the draw"
    for cam in cams:
        cam.close()


def test_geometric_fusion_feedback_two_cameras(oracle):
    """Geometric fusion end to end for two cameras (both on this GPU, standing in for two ranks): per frame the foot
    points are packed on the device, the per-object fuser (principal-axis intersection + Kalman) runs on the host, and the
    feedback (ground pdf -> 20 px test -> rearrangement -> re-score) goes back into each camera's particles.  Checked
    against the numpy oracle of the fuser arithmetic: decision, distance, rearranged particle states (bit-exact, with the
    randgaus stream of the tracker's own MWC state) and the Kalman posterior."""
    import torch
    from oracle import fuser_oracle as fo
    n_frames = 6
    seq = synth.Sequence(640, 480, 3, n_frames=n_frames)
    # Camera 1's ground plane is camera 0's with the axes swapped, and the tracked box's foot point lies on the diagonal
    # (300, 300): the two principal axes meet at the foot itself.  The feedback then finds particles under the (tight)
    # Kalman posterior, and the reference's foot of the AVERAGE particle -- cy + h0*sy, without the 1/2 -- lies ~40 px
    # below the fused location, so the rearrangement branch runs.
    box0 = (280, 220, 40, 80)
    H = [np.eye(3), np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])]
    over = dict(n_particles=400, n_haar=80)
    p = dict(mhost.DEFAULTS); p.update(over)
    cams = []
    gray, r, g, b = seq.frame(0)
    for c in range(2):
        cam = mhost.Camera(0, n_frames)
        cam.set_frame(gray, r, g, b)
        cam.add_object(box0, rng_seed=11 + c, **over)
        cam.init_trackers()
        cams.append(cam)
    fuser = mhost.GeometricFuser(H)
    kf = None
    N = p["n_particles"]
    pts = torch.zeros((1, 2), dtype=torch.float32, device="cuda")
    n_rearranged = 0
    for t in range(1, n_frames):
        gray, r, g, b = seq.frame(t)
        foot = np.zeros((2, 2), np.float32)
        for c, cam in enumerate(cams):
            cam.set_frame(gray, r, g, b)
            nz = np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), camera=c, obj=0, t=t)])
            cam.track(t, nz, 1)
            cam.pack_foot_points(pts.data_ptr()); cam.sync()
            foot[c] = pts.cpu().numpy()[0]
            st = cam.particles(0)
            a = int(np.argmax(st[:, 4]))
            assert foot[c, 0] == st[a, 0] and foot[c, 1] == np.float32(st[a, 1] + np.float32(st[a, 3] * np.float32(cam.state(0)[3])) / np.float32(2))
        first = not fuser.initialized
        fuser.fuse(foot)
        g_or = fo.principal_axis_intersection(foot, H)
        if first:
            kf = fo.KalmanCV(np.float32(g_or[0]), np.float32(g_or[1]))
        else:
            kf.update(np.array([np.float32(g_or[0]), np.float32(g_or[1])], np.float32), np.array([[5, 0], [0, 5]], np.float32))
        mean, cov, stp, P = fuser.posterior()
        np.testing.assert_allclose(stp, kf.x_post[:, 0], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(P, kf.P_post, rtol=1e-5, atol=1e-4)
        if first:
            continue
        for c, cam in enumerate(cams):
            st = cam.particles(0)
            cur = cam.state(0)
            tot, mf, dist, want = fo.ground_feedback_decision(st, cur[3], H[c], mean, cov)
            rng0 = cam.rng_state(0)
            # camera 0 through the per-object call, camera 1 through the batched one (one device pass for all objects)
            got, d = cam.ground_feedback(0, H[c], mean, cov) if c == 0 else cam.ground_feedback_all(H[c], [mean], [cov])[0]
            assert got == want, (t, c, dist, d)
            if tot > 0:
                assert abs(d - dist) <= 1e-3 * max(1.0, dist)
            if got:
                n_rearranged += 1
                seed = rng0 if rng0 < (1 << 63) else rng0 - (1 << 64)
                g2 = mhost.randgaus_stream(seed, 10.0, 2 * N)
                nz2 = np.stack([g2[0::2], g2[1::2]])                   # (xRand, yRand) per particle, in order
                exp = fo.rearrange_particles(st, mf, nz2, cur[2], cur[3])
                now = cam.particles(0)
                # the re-score that follows touches the states in two documented ways only: UpdateParticleWeight's scale-ratio
                # clamp (ParticleFilter.cpp:658-665) and, when ResampleIfRequired fires, a resampling of these rows
                cl = exp.copy()
                for q in range(N):
                    sx, sy = cl[q, 2], cl[q, 3]
                    if float(np.float32(sx / sy)) > 1.2:
                        cl[q, 3] = np.float32(0.9 * float(sx))
                    elif float(np.float32(sy / sx)) > 1.2:
                        cl[q, 2] = np.float32(0.9 * float(sy))
                rows = {tuple(x) for x in exp[:, :4].tolist()} | {tuple(x) for x in cl[:, :4].tolist()}
                assert all(tuple(x) in rows for x in now[:, :4].tolist())
                # the particles now sit around the fused location seen in this camera (x; the foot's y is moved again by the
                # scale-ratio clamp of the re-score)
                assert abs(float(np.median(now[:, 0])) - float(mf[0])) < 15
            else:
                assert cam.rng_state(0) == rng0
    assert n_rearranged >= 1
    fuser.close()
    for cam in cams:
        cam.close()


@pytest.mark.parametrize("cfg", ["haar_milboost", "haarmdch_3obj", "single_positive_r1", "haar_adaboost"])
def test_simple_tracker_sequence_matches_oracle(oracle, cfg):
    """SimpleTracker (the reference's default Local_Tracker_Type 0, SimpleTracker.cpp): exhaustive search window scored by the
    strong classifier, best response wins, classifier updated around the new position.  All objects of the camera go through
    one batched classify + argmax and one batched update per frame.  Against the oracle's restatement: identical positions,
    identical search-window sizes and RNG streams, identical selectors, best response within 1e-5."""
    n_frames = 10
    if cfg == "haar_milboost":
        n_obj, over = 1, dict(search_win=20)
    elif cfg == "haarmdch_3obj":
        n_obj, over = 3, dict(feature_type=3, weak_type=1, n_haar=100, search_win=15)
    elif cfg == "haar_adaboost":
        # the reference's other default pairing: SimpleTracker + online AdaBoost (Tracker_Strong_Classifier_Type 2 in config.cfg)
        n_obj, over = 2, dict(strong_type=0, search_win=15, n_haar=150)
    else:
        # posRadiusTrain 1: the single current box is the only positive (SimpleTracker.cpp:375-386).
        # (MILAnyBoost on top of empty MDCH bins is avoided here on purpose: its retry branch subtracts and re-adds prediction
        # rows, so "did the NLL improve" can hang on the last bit of values that only agree to 1e-5 between libm and CUDA.)
        n_obj, over = 2, dict(pos_radius_train=1.0, n_neg_train=40, search_win=12, n_haar=120)
    seq = synth.Sequence(640, 480, n_obj, n_frames=n_frames)
    p = dict(mhost.DEFAULTS); p.update(over)
    cam = mhost.Camera(0, n_frames)
    gray, r, g, b = seq.frame(0)
    cam.set_frame(gray, r, g, b)
    f0 = oracle.frame(gray, r, g, b)
    otr = []
    for o in range(n_obj):
        cam.add_simple_object(seq.box(0, o), rng_seed=o + 1, **over)
        rng = oracle.rng(o + 1)
        t = oracle.haar_generate(rng, p["n_haar"], seq.w0, seq.h0) if p["feature_type"] in (0, 3) else None
        F = {0: p["n_haar"], 1: 11, 2: 512, 3: p["n_haar"] + 512}[p["feature_type"]]
        K = int(F * p["percent_selected"] / 100)
        clf = oracle.clf_create(p["feature_type"], p["n_haar"], p["n_bins"], seq.w0, seq.h0, p["weak_type"], p["strong_type"], K, 0.85, t)
        tp = oracle_py.TrackerParams(p["n_particles"], p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"], p["pos_radius_train"],
                                     p["init_pos_radius_train"], p["n_neg_train"], p["init_n_neg_train"], 100000,
                                     p["search_win"], p["trajectory_option"], 0, 0, 30)
        x, y, w, h = seq.box(0, o)
        init = np.array([x, y, w, h], np.float32)
        trk = oracle.lib.orc_simple_create(C.byref(tp), clf, C.byref(f0), init.ctypes.data_as(C.POINTER(C.c_float)), C.byref(rng))
        otr.append((trk, clf, rng))
    cam.init_simple()
    for o in range(n_obj):
        assert cam.simple_selectors(o).tolist() == oracle.clf_selectors(otr[o][1]).tolist()
        assert cam.simple_rng_state(o) == otr[o][2].state
    for t in range(1, n_frames):
        gray, r, g, b = seq.frame(t)
        cam.set_frame(gray, r, g, b)
        of = oracle.frame(gray, r, g, b)
        st, resp = cam.track_simple(t)
        cam.sync()
        for o in range(n_obj):
            ost = np.zeros(4, np.float32)
            oresp = oracle.lib.orc_simple_track(otr[o][0], C.byref(of), ost.ctypes.data_as(C.POINTER(C.c_float)))
            assert cam.simple_last_test_size(o) == oracle.lib.orc_simple_last_n(otr[o][0])
            np.testing.assert_array_equal(st[o], ost), (cfg, t, o)
            assert abs(resp[o] - oresp) <= 1e-5 * max(1.0, abs(oresp)), (cfg, t, o, resp[o], oresp)
            assert cam.simple_rng_state(o) == otr[o][2].state, (cfg, t, o)
            assert cam.simple_selectors(o).tolist() == oracle.clf_selectors(otr[o][1]).tolist(), (cfg, t, o)
            gx, gy, gw, gh = seq.box(t, o)
            assert abs(st[o][0] - gx) <= 40 and abs(st[o][1] - gy) <= 40       # sanity: stays near the synthetic object
    for trk, clf, _ in otr:
        oracle.lib.orc_simple_free(trk); oracle.lib.orc_clf_free(clf)
    cam.close()


def _p2p_worker(rank, world, port, q):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)          # plumbing only: the payload moves over NVLink
    try:
        seq = synth.Sequence(640, 480, 3, camera=rank, n_frames=6)
        cam = mhost.Camera(rank, 6)
        gray, r, g, b = seq.frame(0)
        cam.set_frame(gray, r, g, b)
        for o in range(3):
            cam.add_object(seq.box(0, o), rng_seed=100 * rank + o + 1, n_particles=300, n_haar=60)
        cam.init_trackers()

        def gather(bts):
            out = [None] * world
            dist.all_gather_object(out, bts)
            return out
        cam.p2p_setup(rank, world, gather)
        dist.barrier()
        pts = torch.zeros((3, 2), dtype=torch.float32, device="cuda")
        for t in range(1, 5):
            gray, r, g, b = seq.frame(t)
            cam.set_frame(gray, r, g, b)
            nz = np.stack([synth.noise(300, (20, 20, 1e-7, 1e-7), camera=rank, obj=o, t=t) for o in range(3)])
            cam.track(t, nz, 1)
            allp = cam.exchange_foot_points(timeout_ms=5000)
            cam.pack_foot_points(pts.data_ptr()); cam.sync()
            mine = pts.cpu().numpy()
            assert np.array_equal(allp[rank], mine)
            every = [None] * world
            dist.all_gather_object(every, mine)
            for rr in range(world):
                assert np.array_equal(allp[rr], every[rr]), (t, rr)
        cam.close()
        q.put((rank, "ok"))
    except BaseException as e:           # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))
        raise
    finally:
        dist.destroy_process_group()


def test_peer_memory_foot_point_exchange_two_gpus():
    """the geometric fuser's payload exchanged by the library's own kernels over NVLink peer memory (CUDA IPC buffers, stores
    from the foot-point kernel, bounded wait): every camera ends each frame with every camera's foot points, identical to
    the locally packed ones.  Needs two GPUs."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
