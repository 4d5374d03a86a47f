import ctypes as C, json, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
# run the bench's resident loop for a few steps, then read the stamps of the last k_pf_update launch
import bench
from multicameratracking_b200 import capi
L = capi.load()
import torch
args = type("A", (), dict(steps=7, warmup=3, gpus=1, no_cpu_baseline=True, impl="ours", e2e_bgr=False))()
bench.run_ours(args)
out = (C.c_ulonglong * 64)()
L.mct_debug_pf_stamps.argtypes = [C.POINTER(C.c_ulonglong)]
print("rc", L.mct_debug_pf_stamps(out))
names = {0: "start", 1: "after count/min/max", 2: "after weight update", 3: "after chain1", 4: "after normalise+s2", 5: "after decision",
         6: "after chain2", 7: "after divide+cdf chain", 8: "after search+gather", 10: "after resample/copy", 11: "summary: before average",
         12: "after average", 13: "after run weights", 15: "end"}
names[6] = "resample: before normalise+cdf"
names.update({16: "after chain2", 17: "after divide", 18: "after search", 19: "after gather+barrier", 12: "summary: after warp reduce", 13: "summary: after barrier"})
for base, lab in ((0, "no resample"), (32, "RESAMPLED")):
    print("----", lab)
    t0 = out[base]
    for k in sorted(names, key=lambda k: out[base + k]):
        if out[base + k] and out[base + k] >= t0:
            print("%-32s %8.2f us" % (names[k], (out[base + k] - t0) * 1e-3))
