"""Multi-camera parity (BASELINE configs[2] / [3]): the product's CameraNetwork (host/mct_network.cpp over the C-ABI, all
cameras of the network on this GPU) against THE REFERENCE'S OWN CODE running the same schedule (oracle/_ref: Object,
ParticleFilterTracker, GeometryBasedInformationFuser, AppearanceBasedInformationFuser compiled from /root/reference/src and
driven by ref_harness.cpp's restatement of CameraNetwork::TrackObjectsOnCurrentFrame, CameraNetwork.cpp:537-613).

Same synthetic multi-view scene, same prediction noise, one MWC stream per (camera, object) pair on both sides.  Bar: tracked
boxes within 1 px, selector lists (local and global classifiers) and every pair's RNG state equal after every frame, the
geometric fuser's Kalman state within 1e-4 relative."""
import numpy as np
import pytest

from multicameratracking_b200 import host as mhost
from multicameratracking_b200 import synth
from oracle.ref_py import Ref, RefNetwork

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built")

CASES = {
    # name: (cameras, objects, frames, W, H, per-object overrides, network settings)
    "geometric_2x2_haar": (2, 2, 6, 640, 480, dict(n_particles=400, n_haar=80), dict(geometric_fusion=1)),
    "appearance_2x2_mdch": (2, 2, 6, 640, 480, dict(n_particles=300, feature_type=2, strong_type=2), dict(appearance_fusion=2, af_refresh=1)),
    "both_3x2_haarmdch_refresh2": (3, 2, 7, 640, 480, dict(n_particles=300, feature_type=3, n_haar=60), dict(geometric_fusion=1, appearance_fusion=2, af_refresh=2)),
    # BASELINE configs[2] at its full shape: 4 cameras x 8 objects, Haar(250) + MILBoost, 1000 particles, geometric fuser
    "cfg2_full_4x8_geometric": (4, 8, 6, 640, 480, dict(), dict(geometric_fusion=1)),
}


def _noise(scene, N, p, t):
    return np.stack([np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), camera=c, obj=o, t=t) for o in range(scene.O)])
                     for c in range(scene.C)])


@needs_ref
@pytest.mark.parametrize("case", list(CASES))
def test_camera_network_matches_reference(case):
    Cn, On, T, W, H, over, netkw = CASES[case]
    ref = Ref()
    scene = synth.MultiViewScene(Cn, On, W, H, n_frames=T)
    p = dict(mhost.DEFAULTS); p.update(over)
    N = p["n_particles"]
    init = np.array([[scene.box(c, 0, o) for o in range(On)] for c in range(Cn)], np.float32)
    seeds = np.array([[1000 * c + o + 1 for o in range(On)] for c in range(Cn)])
    rkw = dict(feature_type=p["feature_type"], n_haar=p["n_haar"], n_bins=p["n_bins"], strong_type=p["strong_type"], weak_type=p["weak_type"],
               pct_selected=p["percent_selected"], n_particles=N, sd_x=p["sd_x"], sd_y=p["sd_y"], sd_sx=p["sd_sx"], sd_sy=p["sd_sy"],
               pos_radius_train=int(p["pos_radius_train"]), init_pos_radius_train=int(p["init_pos_radius_train"]), n_neg_train=p["n_neg_train"],
               init_n_neg_train=p["init_n_neg_train"], search_win=p["search_win"], trajectory_option=p["trajectory_option"])
    rkw.update(netkw)
    rnet = RefNetwork(ref, Cn, On, T, scene.homographies, init, seeds, **rkw)
    rnet.init([scene.frame(c, 0) for c in range(Cn)])
    cams = []
    for c in range(Cn):
        cam = mhost.Camera(0, T)
        cam.set_frame(*scene.frame(c, 0))
        for o in range(On):
            cam.add_object(scene.box(c, 0, o), rng_seed=int(seeds[c, o]), **over)
        cams.append(cam)
    net = mhost.Network(cams, scene.homographies, **netkw)
    net.init()
    F = {0: p["n_haar"], 1: 11, 2: 512, 3: p["n_haar"] + 512}[p["feature_type"]]
    for c in range(Cn):
        for o in range(On):
            assert cams[c].selectors(o).tolist() == rnet.selectors(c, o).tolist(), (case, "init", c, o)
            assert cams[c].rng_state(o) == rnet.rng_state(c, o), (case, "init", c, o)
    if netkw.get("geometric_fusion"):
        for o in range(On):
            np.testing.assert_allclose(net.kalman(o)[0], rnet.kalman(o)[0], rtol=1e-5, atol=1e-3)
    worst = 0
    for t in range(1, T):
        planes = [scene.frame(c, t) for c in range(Cn)]
        nz = _noise(scene, N, p, t)
        rnet.track(t, planes, nz)
        for c in range(Cn):
            cams[c].set_frame(*planes[c])
        net.track(t, [nz[c] for c in range(Cn)])
        for c in range(Cn):
            cams[c].sync()
            for o in range(On):
                rb = rnet.box(c, o, t)
                gb = np.array(cams[c].boxes(o, t), np.float32)
                d = float(np.abs(rb - gb).max())
                worst = max(worst, d)
                assert d <= 1, (case, t, c, o, rb, gb)
                assert cams[c].rng_state(o) == rnet.rng_state(c, o), (case, t, c, o, "RNG streams diverged")
                assert cams[c].selectors(o).tolist() == rnet.selectors(c, o).tolist(), (case, t, c, o)
                if netkw.get("appearance_fusion") and (t + 1) % netkw.get("af_refresh", 1) == 0:
                    assert net.global_selectors(c, o).tolist() == rnet.selectors(c, o, True).tolist(), (case, t, c, o, "global classifier")
        if netkw.get("geometric_fusion"):
            for o in range(On):
                a, Pa = net.kalman(o); b, Pb = rnet.kalman(o)
                np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-3, err_msg="%s frame %d object %d Kalman state" % (case, t, o))
                np.testing.assert_allclose(Pa, Pb, rtol=1e-4, atol=1e-4)
    net.close(); rnet.close()
    for cam in cams:
        cam.close()


# ---------------------------------------------------------------------------------------------------------------------
# Full-shape configurations against REFERENCE-GENERATED golden files (tests/golden/ref_*.json, made by
# tests/golden/make_ref_golden.py from the reference build in the development container): nothing of the reference is
# needed on the GPU box, the synthetic inputs are regenerated from their seeds.
import importlib.util
import json
import os

_G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden_module():
    spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(_G, "make_ref_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["cfg2_full", "cfg3_full"])
def test_camera_network_full_shape_matches_reference_golden(name):
    path = os.path.join(_G, "ref_%s.json" % name)
    if not os.path.exists(path):
        pytest.skip("golden file %s not generated" % path)
    gold = json.load(open(path))
    mg = _golden_module()
    cs = mg.CASES[name]
    Cn, On, T = cs["cameras"], cs["objects"], cs["frames"]
    p = dict(mg.OBJ_DEFAULTS); p.update(cs["obj"])
    N = p["n_particles"]
    scene = synth.MultiViewScene(Cn, On, cs["W"], cs["H"], n_frames=T, camouflage=cs["camouflage"])
    seeds = mg.seeds_for(Cn, On)
    cams = []
    for c in range(Cn):
        cam = mhost.Camera(0, T)
        cam.set_frame(*scene.frame(c, 0))
        for o in range(On):
            cam.add_object(scene.box(c, 0, o), rng_seed=int(seeds[c, o]), **cs["obj"])
        cams.append(cam)
    net = mhost.Network(cams, scene.homographies, **cs["net"])
    net.init()
    pairs = [(c, o) for c in range(Cn) for o in range(On)]
    for k, (c, o) in enumerate(pairs):
        assert cams[c].selectors(o).tolist() == gold["init"]["selectors"][k], (name, "init", c, o)
        assert str(cams[c].rng_state(o)) == gold["init"]["rng"][k], (name, "init", c, o)
    n_global = n_global_equal = 0
    for fr in gold["frames"]:
        t = fr["t"]
        for c in range(Cn):
            cams[c].set_frame(*scene.frame(c, t))
        nz = mg.noise_for(scene, N, p, t)
        net.track(t, [nz[c] for c in range(Cn)])
        for c in range(Cn):
            cams[c].sync()
        for k, (c, o) in enumerate(pairs):
            gb = np.array(cams[c].boxes(o, t)); rb = np.array(fr["boxes"][k])
            assert np.abs(gb - rb).max() <= 1, (name, t, c, o, gb, rb)
            assert str(cams[c].rng_state(o)) == fr["rng"][k], (name, t, c, o, "RNG streams diverged")
            assert cams[c].selectors(o).tolist() == fr["selectors"][k], (name, t, c, o)
            if "global_selectors" in fr:
                # The global classifier ranks ~100 colour bins on pooled MDCH rows; the rows carry the feature tolerance of the
                # contract (1e-5: fixed-point accumulation here, sequential f32 in the reference), so deep in the list a near tie
                # may resolve differently.  Bar: the leading selectors -- the ones that carry the response -- are identical.
                gs, rs = net.global_selectors(c, o).tolist(), fr["global_selectors"][k]
                lead = min(8, len(rs))
                assert gs[:lead] == rs[:lead], (name, t, c, o, "global classifier", gs[:lead], rs[:lead])
                n_global += 1; n_global_equal += int(gs == rs)
        if "kalman" in fr:
            for o in range(On):
                a, Pa = net.kalman(o)
                np.testing.assert_allclose(a, np.array(fr["kalman"][o], np.float32), rtol=1e-4, atol=1e-3)
                np.testing.assert_allclose(Pa.ravel(), np.array(fr["kalman_P"][o], np.float32), rtol=1e-4, atol=1e-4)
    if n_global:
        print("%s: %d of %d global selector lists identical over their whole length" % (name, n_global_equal, n_global))
        assert n_global_equal >= 0.8 * n_global
    net.close()
    for cam in cams:
        cam.close()
