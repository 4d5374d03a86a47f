"""Multi-camera tracking with both fusers over NCCL: one camera per rank per GPU (BASELINE configs[2] / [3] in small).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/fusion_nccl.py
         [--frames 12] [--objects 4] [--particles 1000] [--refresh 3] [--feature 3]

Per frame and rank: track the camera's objects (score path), geometric fusion (device-side foot-point pack -> NCCL
all-gather -> per-object principal-axis intersection + Kalman on every rank -> feedback into the own particles), classifier
update; every `--refresh` frames the appearance fuser pools the training feature rows of all cameras (ragged NCCL all-gather)
into each rank's global classifier and re-weights the particles with it.  Rank 0 prints one JSON line with the timings
(CUDA events around the fusion steps, max over ranks) and the cross-rank consistency checks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multicameratracking_b200 import capi, synth                     # noqa: E402
from multicameratracking_b200 import host as mhost                    # noqa: E402
from multicameratracking_b200.network import AppearanceFusion, CameraNetworkExchange, GeometricFusion   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--objects", type=int, default=4)
    ap.add_argument("--particles", type=int, default=1000)
    ap.add_argument("--refresh", type=int, default=3)
    ap.add_argument("--feature", type=int, default=3, help="0 Haar, 2 MDCH, 3 Haar+MDCH")
    ap.add_argument("--p2p", action="store_true", help="foot points through the peer-memory kernels instead of the NCCL all-gather")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    dev = torch.device("cuda", local)
    W, H, n_obj, N = 640, 480, args.objects, args.particles
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    seq = synth.Sequence(W, H, n_obj, camera=rank, n_frames=args.frames + 1)
    cam = mhost.Camera(local, args.frames + 1, stream=stream.cuda_stream)
    gray, r, g, b = seq.frame(0)
    cam.set_frame(gray, r, g, b)
    over = dict(feature_type=args.feature, n_particles=N, n_haar=100, weak_type=1)
    for o in range(n_obj):
        cam.add_object(seq.box(0, o), rng_seed=1000 * rank + o + 1, **over)
    cam.init_trackers()
    p = dict(mhost.DEFAULTS); p.update(over)

    ex = CameraNetworkExchange(n_obj)
    # image -> ground homographies: camera r looks at the ground plane rotated by r * 360 / world degrees
    Hs = []
    for c in range(world):
        a = 2 * np.pi * c / max(world, 2) + 0.3
        Hs.append(np.array([[np.cos(a), -np.sin(a), 40.0 * c], [np.sin(a), np.cos(a), -25.0 * c], [1e-4 * c, -5e-5 * c, 1.0]]))
    gf = GeometricFusion(ex, cam, Hs, p2p=args.p2p)

    # appearance fuser: one global classifier per object, the SAME feature table on every rank (seeded by the object id)
    # (colour-histogram features: the table is the same on every rank by construction; cfg 3 uses MDCH + MILAnyBoost)
    ctx = cam.capi_context()
    gclf, af = [], []
    for o in range(n_obj):
        clf = capi.StrongClassifier(ctx, capi.FEAT_MDCH, seq.w0, seq.h0, n_bins=8, weak_type=capi.WEAK_STUMP,
                                    strong_type=capi.STRONG_MILANYBOOST, n_selected=102, learning_rate=0.0)
        gclf.append(clf); af.append(AppearanceFusion(ex, ctx, clf))

    n_rearr = 0
    post_ok = True
    fuse_ms, geo_ms, learn_ms, geo_parts = [], [], [], []
    n_learn_all = 0
    t0 = time.time()
    for t in range(1, args.frames + 1):
        gray, r, g, b = seq.frame(t)
        cam.set_frame(gray, r, g, b)
        nz = np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), camera=rank, obj=o, t=t) for o in range(n_obj)])
        cam.track(t, nz, 1 | 4)                                # TrackObjectWithoutSaveState: score path, state not stored yet
        # appearance re-weighting with the global model (Object.cpp:433-450)
        if t > 2 and t % args.refresh == 0 and all(a.trained for a in af):
            cam.sync()
            for o in range(n_obj):
                cur = cam.state(o)
                pf = cam.particle_filter(o)
                tt = time.perf_counter()
                af[o].fuse(pf, float(cur[2]), float(cur[3]), cam.randfloat(o), cam.randfloat(o))
                fuse_ms.append(round(1e3 * (time.perf_counter() - tt), 2))
        # geometric fusion + feedback (CameraNetwork.cpp:94-268)
        cam.sync(); tt = time.perf_counter()
        allp, fb = gf.step(device=dev)
        cam.sync()
        if t > 1:                                              # (frame 1 carries NCCL's lazy communicator set-up)
            geo_ms.append(1e3 * (time.perf_counter() - tt))
            geo_parts.append(dict(gf.last_ms))
        n_rearr += sum(1 for x in fb if x[0])
        # every rank holds the same Kalman posterior (computed redundantly from the all-gathered foot points)
        post = torch.tensor(np.concatenate([np.concatenate([f.posterior()[2], f.posterior()[3].ravel()]) for f in gf.fusers]), device=dev)
        allpost = [torch.empty_like(post) for _ in range(world)]
        dist.all_gather(allpost, post)
        post_ok = post_ok and all(torch.equal(allpost[0], x) for x in allpost)
        cam.track(t, nz, 8 | 2)                                # StoreAllParticleFilterTrackerState, then UpdateClassifier
        cam.sync()
        # LearnGlobalAppearanceModel every refresh frames (CameraNetwork.cpp:276-320)
        if (t + 1) % args.refresh == 0:
            cam.sync(); tt = time.perf_counter()
            for o in range(n_obj):
                cur = cam.state(o)
                pc, pr = cam.last_train(o, 0); nc, nr = cam.last_train(o, 1)
                one = lambda n, v: np.full(n, v, np.float32)   # noqa: E731
                pos = capi.make_boxes(pc, pr, one(len(pc), cur[4]), one(len(pc), cur[5]))
                neg = capi.make_boxes(nc, nr, one(len(nc), min(cur[4], 10.0)), one(len(nc), min(cur[5], 10.0)))
                af[o].learn(pos, neg, lambda o=o: cam.randfloat(o))
            cam.sync()
            if n_learn_all > 0:                                # (first refresh: first ragged all-gather shapes)
                learn_ms.append(1e3 * (time.perf_counter() - tt))
            n_learn_all += 1
    wall = time.time() - t0
    red = lambda v: ex.max_over_ranks(v, device=dev)          # noqa: E731
    med = lambda v: float(np.median(v)) if len(v) else 0.0    # noqa: E731
    out = {"what": "multi-camera tracking with geometric + appearance fusion over NCCL, one camera per GPU",
           "n_gpus": world, "foot_point_exchange": "peer-memory kernels (NVLink stores)" if args.p2p else "NCCL all-gather", "frames": args.frames, "objects_per_camera": n_obj, "particles_per_object": N, "feature_type": args.feature,
           "timing": "host wall clock around each step with the camera's stream drained before and after, median, max over ranks",
           "geometric_fusion_ms_per_frame": red(med(geo_ms)),
           "geometric_fusion_stages_ms_rank0": {k: round(med([g[k] for g in geo_parts]), 3) for k in (geo_parts[0] if geo_parts else {})}, "rearranged_objects_total": ex.sum_over_ranks(n_rearr, device=dev),
           "appearance_learn_ms_per_refresh_all_objects": red(med(learn_ms)), "appearance_learn_refreshes": len(learn_ms),
           "appearance_fuse_ms_per_object": red(med(fuse_ms[n_obj:])), "kalman_posterior_identical_on_all_ranks": bool(post_ok),
           "selectors_global_obj0": gclf[0].selectors()[:8].tolist(), "wall_s": red(wall)}
    if rank == 0:
        print(json.dumps(out))
    gf.close()
    for c in gclf:
        c.close()
    cam.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
