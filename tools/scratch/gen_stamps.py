import ctypes as C, sys
sys.path.insert(0, '/root/repo')
exec(open('/root/repo/tools/scratch/run_frames.py').read())
from multicameratracking_b200 import capi
L = capi.load()
out = (C.c_ulonglong * 64)()
L.mct_debug_pf_stamps.argtypes = [C.POINTER(C.c_ulonglong)]
print("rc", L.mct_debug_pf_stamps(out))
t0 = out[0]
for k in sorted(range(32), key=lambda k: out[k]):
    if out[k] >= t0 and out[k] - t0 < 10**8: print(k, "%.2f us" % ((out[k] - t0) * 1e-3))
