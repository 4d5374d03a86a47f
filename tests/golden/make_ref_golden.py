"""Reference-generated golden vectors for the multi-camera configurations at their FULL shape (BASELINE configs[2], [3]).

Runs the reference's own code (oracle/_ref: /root/reference/src compiled against oracle/ref_shim, multi-camera schedule of
ref_harness.cpp) on the synthetic multi-view scene and stores, per frame, every (camera, object) pair's tracked box, RNG
state, local selector list, the appearance fuser's global selector lists and the geometric fuser's Kalman state.  The inputs
are synthetic and seeded (multicameratracking_b200/synth.py), so tests/test_gpu_network.py regenerates them on the GPU box and
compares the CUDA path with these files -- no reference code is needed there.

    python tests/golden/make_ref_golden.py [case ...]        (needs /root/reference; cfg3_full takes ~10 minutes)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from multicameratracking_b200 import synth            # noqa: E402
from oracle.ref_py import Ref, RefNetwork             # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# per-object settings follow multicameratracking_b200.host.DEFAULTS (config.cfg values) unless overridden
CASES = {
    # BASELINE configs[2]: 4 cameras x 8 objects, Haar(250) -> 50, MILBoost + stumps, 1000 particles, geometric fuser
    "cfg2_full": dict(cameras=4, objects=8, frames=10, W=640, H=480, camouflage=None,
                      obj=dict(), net=dict(geometric_fusion=1)),
    # BASELINE configs[3]: 8 cameras x 16 objects, 1280x720, MDCH(512) -> 102, MILAnyBoost, 1000 particles, appearance fuser
    # (MILBoost over MDCH, refresh every frame).  camouflage: see synth.Sequence (MILAnyBoost aborts on separable samples)
    "cfg3_full": dict(cameras=8, objects=16, frames=6, W=1280, H=720, camouflage=30,
                      obj=dict(feature_type=2, strong_type=4), net=dict(appearance_fusion=2, af_refresh=1)),
}
OBJ_DEFAULTS = dict(feature_type=0, n_haar=250, n_bins=8, weak_type=0, strong_type=2, percent_selected=20, n_particles=1000,
                    sd_x=20.0, sd_y=20.0, sd_sx=1e-7, sd_sy=1e-7, pos_radius_train=10.0, init_pos_radius_train=7.0,
                    n_neg_train=30, init_n_neg_train=30, search_win=20, trajectory_option=0)


def seeds_for(Cn, On):
    return np.array([[1000 * c + o + 1 for o in range(On)] for c in range(Cn)])


def noise_for(scene, N, p, t):
    return np.stack([np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), camera=c, obj=o, t=t) for o in range(scene.O)])
                     for c in range(scene.C)])


def ref_kwargs(p, net):
    kw = dict(feature_type=p["feature_type"], n_haar=p["n_haar"], n_bins=p["n_bins"], strong_type=p["strong_type"], weak_type=p["weak_type"],
              pct_selected=p["percent_selected"], n_particles=p["n_particles"], sd_x=p["sd_x"], sd_y=p["sd_y"], sd_sx=p["sd_sx"], sd_sy=p["sd_sy"],
              pos_radius_train=int(p["pos_radius_train"]), init_pos_radius_train=int(p["init_pos_radius_train"]), n_neg_train=p["n_neg_train"],
              init_n_neg_train=p["init_n_neg_train"], search_win=p["search_win"], trajectory_option=p["trajectory_option"])
    kw.update(net)
    return kw


def run(name):
    cs = CASES[name]
    Cn, On, T = cs["cameras"], cs["objects"], cs["frames"]
    p = dict(OBJ_DEFAULTS); p.update(cs["obj"])
    scene = synth.MultiViewScene(Cn, On, cs["W"], cs["H"], n_frames=T, camouflage=cs["camouflage"])
    init = np.array([[scene.box(c, 0, o) for o in range(On)] for c in range(Cn)], np.float32)
    ref = Ref()
    t0 = time.time()
    net = RefNetwork(ref, Cn, On, T, scene.homographies, init, seeds_for(Cn, On), **ref_kwargs(p, cs["net"]))
    net.init([scene.frame(c, 0) for c in range(Cn)])
    out = {"case": name, "settings": {k: v for k, v in cs.items()}, "frames": []}
    pairs = [(c, o) for c in range(Cn) for o in range(On)]
    out["init"] = {"selectors": [net.selectors(c, o).tolist() for c, o in pairs], "rng": [str(net.rng_state(c, o)) for c, o in pairs]}
    if cs["net"].get("geometric_fusion"):
        out["init"]["kalman"] = [net.kalman(o)[0].astype(float).tolist() for o in range(On)]
    for t in range(1, T):
        net.track(t, [scene.frame(c, t) for c in range(Cn)], noise_for(scene, p["n_particles"], p, t))
        fr = {"t": t, "boxes": [net.box(c, o, t).astype(int).tolist() for c, o in pairs],
              "rng": [str(net.rng_state(c, o)) for c, o in pairs],
              "selectors": [net.selectors(c, o).tolist() for c, o in pairs]}
        if cs["net"].get("appearance_fusion"):
            fr["global_selectors"] = [net.selectors(c, o, True).tolist() for c, o in pairs]
        if cs["net"].get("geometric_fusion"):
            fr["kalman"] = [net.kalman(o)[0].astype(float).tolist() for o in range(On)]
            fr["kalman_P"] = [net.kalman(o)[1].astype(float).ravel().tolist() for o in range(On)]
        out["frames"].append(fr)
        print("%s frame %d done after %.0f s" % (name, t, time.time() - t0), flush=True)
    net.close()
    with open(os.path.join(HERE, "ref_%s.json" % name), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        run(nm)
