"""Multi-process parity of the camera network: one camera per rank per GPU (torchrun), exchange over NCCL (TorchExchange), against
the reference build (oracle/_ref) running the same network on rank 0's host.  Checks boxes (<= 1 px), RNG states and selector lists
of EVERY camera after every frame (the other ranks' results are all-gathered).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/network_parity_mp.py [--mode geo|app|both]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multicameratracking_b200 import host as mhost, synth            # noqa: E402
from multicameratracking_b200.network import TorchExchange            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="both")
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--objects", type=int, default=3)
    ap.add_argument("--particles", type=int, default=400)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Cn, On, T, N = world, args.objects, args.frames, args.particles
    netkw = {"geo": dict(geometric_fusion=1), "app": dict(appearance_fusion=2, af_refresh=1),
             "both": dict(geometric_fusion=1, appearance_fusion=2, af_refresh=2)}[args.mode]
    over = dict(n_particles=N, feature_type=3 if args.mode != "geo" else 0, n_haar=60 if args.mode != "geo" else 80)
    if args.mode == "geo":
        On = 2          # (with 3 objects the reference itself aborts in frame 4 of this scene: its feedback inflates a box until the sampler fails)
    scene = synth.MultiViewScene(Cn, On, 640, 480, n_frames=T)
    p = dict(mhost.DEFAULTS); p.update(over)
    seeds = np.array([[1000 * c + o + 1 for o in range(On)] for c in range(Cn)])
    cam = mhost.Camera(local, T)
    cam.set_frame(*scene.frame(rank, 0))
    for o in range(On):
        cam.add_object(scene.box(rank, 0, o), rng_seed=int(seeds[rank, o]), **over)
    ex = TorchExchange()
    net = mhost.Network([cam], scene.homographies, camera_ids=[rank], n_cameras=Cn, exchange=ex.as_tuple(), **netkw)
    net.init()
    ex.bind_rows(net)
    rnet = None
    if rank == 0:
        from oracle.ref_py import Ref, RefNetwork
        init = np.array([[scene.box(c, 0, o) for o in range(On)] for c in range(Cn)], np.float32)
        rkw = dict(feature_type=p["feature_type"], n_haar=p["n_haar"], n_particles=N)
        rkw.update(netkw)
        rnet = RefNetwork(Ref(), Cn, On, T, scene.homographies, init, seeds, **rkw)
        rnet.init([scene.frame(c, 0) for c in range(Cn)])
    ok, worst = True, 0
    for t in range(1, T):
        nz = np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), camera=rank, obj=o, t=t) for o in range(On)])
        cam.set_frame(*scene.frame(rank, t))
        net.track(t, [nz])
        cam.sync()
        mine = {"boxes": [cam.boxes(o, t) for o in range(On)], "rng": [str(cam.rng_state(o)) for o in range(On)],
                "sel": [cam.selectors(o).tolist() for o in range(On)]}
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        if rank == 0:
            allnz = np.stack([np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), camera=c, obj=o, t=t) for o in range(On)]) for c in range(Cn)])
            rnet.track(t, [scene.frame(c, t) for c in range(Cn)], allnz)
            for c in range(Cn):
                for o in range(On):
                    d = float(np.abs(np.array(allr[c]["boxes"][o]) - rnet.box(c, o, t)).max())
                    worst = max(worst, d)
                    good = d <= 1 and allr[c]["rng"][o] == str(rnet.rng_state(c, o)) and allr[c]["sel"][o] == rnet.selectors(c, o).tolist()
                    if not good:
                        ok = False
                        print("MISMATCH frame %d camera %d object %d: box diff %.0f" % (t, c, o, d), flush=True)
    if rank == 0:
        print(json.dumps({"what": "multi-process camera network vs the reference build", "mode": args.mode, "n_gpus": world, "frames": T - 1,
                          "objects_per_camera": On, "particles": N, "worst_box_diff_px": worst, "rng_and_selectors_equal": ok, "ok": ok}))
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, 0)
    net.close(); cam.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() > 0.5 else 1)


if __name__ == "__main__":
    main()
