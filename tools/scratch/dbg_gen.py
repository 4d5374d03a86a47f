import ctypes as C, numpy as np, sys
sys.path.insert(0, '/root/repo')
from multicameratracking_b200 import capi, synth
from multicameratracking_b200 import host as mhost
from oracle import oracle_py
from oracle.oracle_py import Oracle
sys.path.insert(0, '/root/repo/tests')
import test_gpu_tracker as T
oracle = Oracle()
n_frames, n_obj = 6, 8
over = dict(feature_type=3, weak_type=1, n_particles=2000, n_haar=250)
seq = synth.Sequence(640, 480, n_obj, n_frames=n_frames)
p = dict(mhost.DEFAULTS); p.update(over)
cam = mhost.Camera(0, n_frames)
gray, r, g, b = seq.frame(0)
cam.set_frame(gray, r, g, b)
f0 = oracle.frame(gray, r, g, b)
otr = []
for o in range(n_obj):
    cam.add_object(seq.box(0, o), rng_seed=o + 1, **over)
    otr.append(T._oracle_tracker(oracle, seq, o, f0, p, o + 1))
cam.init_trackers()
N = p["n_particles"]
for t in range(1, 2):
    gray, r, g, b = seq.frame(t)
    cam.set_frame(gray, r, g, b)
    of = oracle.frame(gray, r, g, b)
    nz = np.stack([synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), obj=o, t=t) for o in range(n_obj)])
    boxes = cam.track(t, nz, 3)
    cam.sync()
    for o in range(n_obj):
        ob = np.zeros(4, np.int32)
        oracle.lib.orc_tracker_track(otr[o][0], C.byref(of), nz[o].ctypes.data_as(C.POINTER(C.c_float)), ob.ctypes.data_as(C.POINTER(C.c_int)), 3)
        col, row, sx, sy, P, Nn = capi.track_update_default_boxes(cam.capi_context(), o)
        oc, orw, osx, osy = T._oracle_last_train(oracle, otr[o][0], 0)
        st = cam.state(o)
        ost = np.ctypeslib.as_array(oracle.lib.orc_tracker_state(otr[o][0]), shape=(6,)) if hasattr(oracle.lib.orc_tracker_state, 'restype') and oracle.lib.orc_tracker_state.restype is not None else None
        pf = cam.particle_filter(o)
        summ = pf.summary() if hasattr(pf, 'summary') else None
        if o == 6:
            opf = oracle.lib.orc_tracker_pf(otr[o][0]) if False else None
            from oracle.oracle_py import Pf
            oracle.lib.orc_tracker_pf.restype = C.POINTER(Pf)
            opf = oracle.lib.orc_tracker_pf(otr[o][0])
            ost = oracle.pf_state(opf); gst = cam.particles(o)
            print('state equal', np.array_equal(ost, gst), 'maxdiff', np.abs(ost-gst).max(), 'oracle fs', opf.contents.filter_state, 'last_resampled', opf.contents.last_resampled, 'n_uniq', opf.contents.n_uniq)
            print('gpu summ resampled', summ.resampled, 'fs', summ.filter_state, 'n_unique', summ.n_unique, 'first', list(summ.ordered_first), 'close', list(summ.highest_close))
            u = np.ctypeslib.as_array(opf.contents.uniq, shape=(5*3,))
            print('oracle uniq rows 0-2', u)
            print('gpu state row0', gst[0], 'oracle row0', ost[0])
            ridx = oracle.pf_resample_idx(opf)
            print('oracle residx[:5]', ridx[:5])
            w = np.where(np.abs(gst[:,3]-0.99999994)<1e-9)[0]
            print('particles with sy=0.99999994 and cx==first cx:', [(i, gst[i]) for i in w if gst[i,0]==summ.ordered_first[0]][:5])
        print(t, o, 'dev sx %.9g sy %.9g | host state4 %.9g %.9g | oracle psx %.9g psy %.9g' % (sx[0], sy[0], st[4], st[5], osx[0], osy[0]),
              'summ first', None if summ is None else [float(x) for x in summ.ordered_first][:4])
