import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
# build with MCT_NVCC_EXTRA=-DMCT_CLS_DEBUG; runs the bench's legs, then reads the per-CTA stamps of the last k_classify_staged launch
import bench
from multicameratracking_b200 import capi
L = capi.load()
args = type("A", (), dict(steps=7, warmup=3, gpus=1, no_cpu_baseline=True, impl="ours", e2e_bgr=False, config=1))()
if len(sys.argv) > 1:
    bench.configure(int(sys.argv[1]))
bench.run_ours(args)
out = (C.c_ulonglong * (256 * 8))()
L.mct_debug_cls_stamps.argtypes = [C.POINTER(C.c_ulonglong)]
print("rc", L.mct_debug_cls_stamps(out))
a = np.array(out[:], dtype=np.uint64).reshape(256, 8)
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
names = ["start", "desc", "bbox scan", "region loads issued", "table + region in smem", "own boxes done", "cta done"]
rel = (a[:, :7].astype(np.int64) - np.int64(t0)) * 1e-3
print("CTAs", len(a), " staged", int((a[:, 7] & 1).sum()), " partial", int(((a[:, 7] >> 1) & 1).sum()), " nsel", sorted(set(((a[:, 7] >> 8) & 0xffff).tolist())),
      " region floats mean %.0f max %d" % ((a[:, 7] >> 24).mean(), (a[:, 7] >> 24).max()))
for k, n in enumerate(names):
    print("%-24s min %7.2f  mean %7.2f  max %7.2f us" % (n, rel[:, k].min(), rel[:, k].mean(), rel[:, k].max()))
d = np.diff(rel, axis=1)
for k in range(6):
    print("phase %-22s mean %6.2f  max %6.2f us" % (names[k + 1], d[:, k].mean(), d[:, k].max()))
# with -DMCT_PF_DEBUG as well: where the next kernel of the frame (k_pf_update, CTA 0) starts relative to this kernel's CTAs
try:
    pf = (C.c_ulonglong * 64)()
    L.mct_debug_pf_stamps.argtypes = [C.POINTER(C.c_ulonglong)]
    L.mct_debug_pf_stamps(pf)
    for base in (0, 32):
        if pf[base]:
            print("k_pf_update CTA 0 (%s): start %+.2f us, end %+.2f us relative to the first classify CTA start" % (
                "no resample" if base == 0 else "resampled", (int(pf[base]) - int(t0)) * 1e-3, (int(pf[base + 15]) - int(t0)) * 1e-3))
except Exception as e:
    print("no pf stamps", e)
