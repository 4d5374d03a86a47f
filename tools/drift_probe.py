#!/usr/bin/env python
"""Diagnostic for bench.py's legs: runs the cfg-2 camera for many frames and prints, every `every` frames, the host-timed
step time, the mean distance of the tracked boxes from the synthetic ground truth and the particles' spread.
  python tools/drift_probe.py [score|full] [frames]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multicameratracking_b200 import host as mhost, synth

mode = sys.argv[1] if len(sys.argv) > 1 else "score"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
phases = 1 if mode == "score" else 3
seq, frames = bench.make_sequence(0)
cam = mhost.Camera(0, n_frames=bench.POOL + 8)
cam.set_frame(*frames[0])
for o in range(bench.N_OBJ):
    cam.add_object(seq.box(0, o), feature_type=bench.FEAT, n_haar=bench.N_HAAR, n_bins=bench.N_BINS, weak_type=bench.WEAK,
                   strong_type=bench.STRONG, percent_selected=bench.PCT, n_particles=bench.N_PART, sd_x=bench.SD[0], sd_y=bench.SD[1],
                   sd_sx=bench.SD[2], sd_sy=bench.SD[3], rng_seed=o + 1)
cam.init_trackers()
cam.sync()
every = 20
acc = 0.0
for i in range(n):
    t = bench.pingpong(i, bench.POOL)
    cam.set_frame(*frames[t])
    nz = np.stack([synth.noise(bench.N_PART, bench.SD, obj=o, t=i) for o in range(bench.N_OBJ)])
    cam.sync()
    t0 = time.perf_counter()
    boxes = cam.track(t, nz, phases)
    cam.sync()
    acc += time.perf_counter() - t0
    if (i + 1) % every == 0:
        err = [np.hypot(boxes[o][0] - seq.box(t, o)[0], boxes[o][1] - seq.box(t, o)[1]) for o in range(bench.N_OBJ)]
        spread = []
        for o in range(bench.N_OBJ):
            p = cam.particles(o)
            spread.append(float(np.std(p[:, 0]) + np.std(p[:, 1])))
        print("frame %4d  step %.0f us  box error mean %.1f max %.1f px  particle spread mean %.1f max %.1f" %
              (i + 1, 1e6 * acc / every, np.mean(err), np.max(err), np.mean(spread), np.max(spread)), flush=True)
        acc = 0.0
