#!/usr/bin/env python
"""BASELINE.json configs[4]: scoring-kernel sweep (roofline characterisation of the Haar rectangle-sum gather).
N particles x F Haar features x frame size on one GPU; boxes uniformly random over the frame (default_rng(5)), Haar table
from the reference's Generate with MWC seed 1 (via the oracle, which is only used to GENERATE the table here).
Per point: k_haar_matrix device time (CUDA events around the launch, mct_profile_*), algorithmic bytes
B_haar = N * sum_f(R_f) * 16 + N * F * 4 and the fraction of the measured HBM copy peak.
usage: python tools/sweep_haar.py [--quick] > profiles/<tag>_sweep_haar.json"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multicameratracking_b200 import capi, synth
from oracle.oracle_py import Oracle

quick = "--quick" in sys.argv
orc = Oracle()
ctx = capi.Context(0)
L = capi.load()
peak = 6545.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
res = []
sizes = [(640, 480), (1280, 720), (1920, 1080)]
Ns = [1000, 4000, 16000, 64000]
Fs = [250, 1000, 2000]
if quick:
    sizes, Ns, Fs = sizes[:2], [4000, 64000], [250, 2000]
for (W, H) in sizes:
    seq = synth.Sequence(W, H, 1)
    w0, h0 = seq.w0, seq.h0
    gray = seq.frame(0)[0]
    ctx.frame_set(gray)
    for F in Fs:
        t = orc.haar_generate(orc.rng(1), F, w0, h0)
        arrs = orc.haar_arrays(t)
        sumR = int(np.asarray(arrs[0]).sum())
        clf = capi.StrongClassifier(ctx, capi.FEAT_HAAR, w0, h0, n_haar=F, n_selected=max(1, F // 5), haar=arrs)
        for N in Ns:
            rng = np.random.default_rng(5)
            col = rng.integers(1, W - w0 - 2, N).astype(np.int32); row = rng.integers(1, H - h0 - 2, N).astype(np.int32)
            sx = np.ones(N, np.float32); sy = np.ones(N, np.float32)
            bx = capi.make_boxes(col, row, sx, sy)
            clf.features(bx)                       # warm-up (allocations)
            capi.profile_enable(L, ctx.h, True)
            reps = 3
            for _ in range(reps):
                clf.features(bx)
            prof = capi.profile_read(L, ctx.h)
            capi.profile_enable(L, ctx.h, False)
            if "k_haar_matrix_tiled" in prof:       # tile sort + tiled gather
                ms, n = prof["k_haar_matrix_tiled"]
                us = 1e3 * (ms + prof["k_tile_sort"][0]) / n
                kern, sort_us = "k_tile_sort + k_haar_matrix_tiled", round(1e3 * prof["k_tile_sort"][0] / n, 2)
            else:
                ms, n = prof["k_haar_matrix"]
                us = 1e3 * ms / n
                kern, sort_us = "k_haar_matrix", 0.0
            alg = N * sumR * 16 + N * F * 4
            gbs = alg / (us * 1e-6) / 1e9
            res.append({"frame": [W, H], "N": N, "F": F, "sum_rects": sumR, "kernels": kern, "us": round(us, 2), "of_which_sort_us": sort_us, "algorithmic_bytes": alg,
                        "achieved_gbs": round(gbs, 1), "frac_of_measured_hbm_peak": round(gbs / peak, 3),
                        "rect_sums_per_s": round(N * sumR / (us * 1e-6) / 1e9, 2)})
            print(json.dumps(res[-1]), file=sys.stderr)
        clf.close()
print(json.dumps({"kernel": "Haar feature matrix (FeatureVector::Compute)", "peak_gbs": peak, "unit_rect_sums_per_s": "G/s", "points": res}, indent=1))
