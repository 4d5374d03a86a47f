"""ctypes binding of oracle/_ref/libmct_ref.so = the UNMODIFIED reference sources behind ref_harness.cpp.

TEST INFRASTRUCTURE ONLY (tests/, tests/golden/make_ref_golden.py, bench.py --impl reference).  The product package never
imports this module.  `Ref.available()` is False when the library has not been built and /root/reference is absent
(build_ref.py needs the reference sources; the built library travels to the GPU box).
"""
import ctypes as C
import os

import numpy as np

from . import build_ref

_c_int_p = C.POINTER(C.c_int)
_c_float_p = C.POINTER(C.c_float)
_c_u8_p = C.POINTER(C.c_uint8)


def _fp(a):
    return None if a is None else a.ctypes.data_as(_c_float_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_c_int_p)


def _up(a):
    return None if a is None else a.ctypes.data_as(_c_u8_p)


class TrackerParams(C.Structure):
    _fields_ = [("n_particles", C.c_int), ("sd_x", C.c_float), ("sd_y", C.c_float), ("sd_sx", C.c_float),
                ("sd_sy", C.c_float), ("pos_radius_train", C.c_float), ("init_pos_radius_train", C.c_float),
                ("n_neg_train", C.c_int), ("init_n_neg_train", C.c_int), ("max_pos_train", C.c_int),
                ("search_win", C.c_int), ("trajectory_option", C.c_int), ("pos_strategy", C.c_int),
                ("neg_strategy", C.c_int), ("max_num_pos_examples", C.c_int)]


class RefAbort(RuntimeError):
    """the reference hit one of its own assertions (abortError, Public.h:288-308: it would have exit(0)ed)"""


class Ref:
    _lib = None

    @staticmethod
    def available():
        try:
            return build_ref.build() is not None
        except Exception:
            return False

    def __init__(self):
        if Ref._lib is None:
            path = build_ref.build()
            if path is None:
                raise RuntimeError("oracle/_ref/libmct_ref.so is not built and the reference sources are absent")
            Ref._lib = C.CDLL(path)
            self._proto(Ref._lib)
        self.lib = Ref._lib
        self.lib.ref_set_log(0)

    @staticmethod
    def _proto(L):
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.ref_seed.argtypes = [i]
        L.ref_randfloat.restype = f
        L.ref_randint.argtypes = [i, i]; L.ref_randint.restype = i
        L.ref_cvround.argtypes = [C.c_double]; L.ref_cvround.restype = i
        L.ref_set_log.argtypes = [i]
        L.ref_frame_create.argtypes = [_c_u8_p] * 4 + [i, i, i]; L.ref_frame_create.restype = vp
        L.ref_frame_free.argtypes = [vp]
        L.ref_integral.argtypes = [vp, _c_float_p]
        L.ref_sum_rect.argtypes = [vp, i, i, i, i]; L.ref_sum_rect.restype = f
        L.ref_bgr_to_planes.argtypes = [_c_u8_p, i, i, i, _c_u8_p, _c_u8_p]
        L.ref_clf_create.argtypes = [i] * 8 + [f, i, f]; L.ref_clf_create.restype = vp
        L.ref_clf_free.argtypes = [vp]
        L.ref_clf_num_features.argtypes = [vp]; L.ref_clf_num_features.restype = i
        L.ref_clf_num_selected.argtypes = [vp]; L.ref_clf_num_selected.restype = i
        L.ref_clf_haar_table.argtypes = [vp, _c_int_p, _c_int_p, _c_float_p]; L.ref_clf_haar_table.restype = i
        box = [i, _c_int_p, _c_int_p, _c_float_p, _c_float_p]
        L.ref_clf_features.argtypes = [vp, vp] + box + [_c_float_p, i]
        L.ref_clf_update.argtypes = [vp, vp] + box + box; L.ref_clf_update.restype = i
        L.ref_clf_classify.argtypes = [vp, vp] + box + [i, _c_float_p]
        L.ref_clf_get_selectors.argtypes = [vp, _c_int_p]; L.ref_clf_get_selectors.restype = i
        L.ref_clf_get_alphas.argtypes = [vp, _c_float_p]; L.ref_clf_get_alphas.restype = i
        L.ref_clf_get_weak_state.argtypes = [vp, _c_float_p]
        L.ref_weak_run.argtypes = [i, f, i, _c_float_p, i, _c_float_p, i, _c_float_p, _c_float_p, _c_float_p, _c_float_p, i,
                                   _c_float_p, _c_int_p]
        L.ref_sample_ring.argtypes = [vp, i, i, i, i, f, f, i, f, f, _c_int_p, _c_int_p, i]; L.ref_sample_ring.restype = i
        L.ref_sample_rect.argtypes = [vp, i, i, i, i, f, f, f, f, i, f, f, _c_int_p, _c_int_p, i]; L.ref_sample_rect.restype = i
        L.ref_pf_create.argtypes = [i] + [f] * 9; L.ref_pf_create.restype = vp
        L.ref_pf_free.argtypes = [vp]
        L.ref_pf_get.argtypes = [vp, _c_float_p, _c_float_p, _c_int_p, _c_int_p]
        L.ref_pf_set.argtypes = [vp, _c_float_p, _c_float_p, i]
        L.ref_pf_predict.argtypes = [vp, _c_float_p] + [f] * 6
        L.ref_pf_update_all_weights.argtypes = [vp, _c_float_p, i, i, i, i]
        L.ref_pf_update_particle_weight.argtypes = [vp, i, f, i]
        L.ref_pf_normalize.argtypes = [vp]
        L.ref_pf_resample.argtypes = [vp, i]
        L.ref_pf_resample_if_required.argtypes = [vp]
        L.ref_pf_get_average.argtypes = [vp, _c_float_p]
        L.ref_pf_get_highest_weight.argtypes = [vp, _c_float_p]
        L.ref_pf_ordered_unique.argtypes = [vp, _c_float_p, i]; L.ref_pf_ordered_unique.restype = i
        L.ref_pf_get_highest_unique_close.argtypes = [vp, _c_float_p, _c_float_p]
        L.ref_pf_rearrange.argtypes = [vp, _c_float_p, i, f, f]
        L.ref_tracker_create.argtypes = [i, C.POINTER(TrackerParams)] + [i] * 6 + [f, i, f, vp, _c_float_p, i]
        L.ref_tracker_create.restype = vp
        L.ref_tracker_free.argtypes = [vp]
        L.ref_tracker_track.argtypes = [vp, vp, _c_float_p, _c_int_p, i]; L.ref_tracker_track.restype = i
        L.ref_tracker_state.argtypes = [vp, _c_float_p]
        L.ref_tracker_states_row.argtypes = [vp, i, _c_float_p]
        L.ref_tracker_pf.argtypes = [vp]; L.ref_tracker_pf.restype = vp
        L.ref_tracker_selectors.argtypes = [vp, _c_int_p]; L.ref_tracker_selectors.restype = i
        L.ref_tracker_weak_state.argtypes = [vp, _c_float_p]
        L.ref_tracker_test_set_size.argtypes = [vp]; L.ref_tracker_test_set_size.restype = i

    # ---- RNG ----
    def seed(self, s):
        self.lib.ref_seed(s)

    def randfloat(self):
        return self.lib.ref_randfloat()

    def randint(self, a, b):
        return self.lib.ref_randint(a, b)

    # ---- frames ----
    def frame(self, gray, c0=None, c1=None, c2=None):
        gray = np.ascontiguousarray(gray, np.uint8)
        H, W = gray.shape
        pl = [None if p is None else np.ascontiguousarray(p, np.uint8) for p in (c0, c1, c2)]
        return self.lib.ref_frame_create(_up(gray), _up(pl[0]), _up(pl[1]), _up(pl[2]), W, H, W), (W, H)

    def frame_free(self, fr):
        self.lib.ref_frame_free(fr[0])

    def integral(self, fr):
        W, H = fr[1]
        out = np.empty((H + 1, W + 1), np.float32)
        self.lib.ref_integral(fr[0], _fp(out))
        return out

    def sum_rect(self, fr, x, y, w, h):
        return self.lib.ref_sum_rect(fr[0], x, y, w, h)

    def bgr_to_planes(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        H, W = bgr.shape[:2]
        gray = np.empty((H, W), np.uint8); rgb = np.empty((3, H, W), np.uint8)
        self.lib.ref_bgr_to_planes(_up(bgr), W, H, 3 * W, _up(gray), _up(rgb))
        return gray, rgb[0], rgb[1], rgb[2]

    # ---- classifier ----
    @staticmethod
    def _box(col, row, sx, sy):
        col = np.ascontiguousarray(col, np.int32); row = np.ascontiguousarray(row, np.int32)
        sx = np.ascontiguousarray(sx, np.float32); sy = np.ascontiguousarray(sy, np.float32)
        return (len(col), _ip(col), _ip(row), _fp(sx), _fp(sy)), (col, row, sx, sy)

    def clf_create(self, feature_type, n_haar, nb, w0, h0, weak_type, strong_type, K, lr, use_hsv=0, pct_retained=0.0):
        return self.lib.ref_clf_create(feature_type, n_haar, nb, w0, h0, weak_type, strong_type, K, lr, use_hsv, pct_retained)

    def clf_haar_table(self, clf, F):
        nrect = np.zeros(F, np.int32); rect = np.zeros((F, 6, 4), np.int32); wt = np.zeros((F, 6), np.float32)
        n = self.lib.ref_clf_haar_table(clf, _ip(nrect), _ip(rect), _fp(wt))
        return nrect[:n], rect[:n], wt[:n]

    def clf_features(self, clf, fr, col, row, sx, sy):
        F = self.lib.ref_clf_num_features(clf)
        b, keep = self._box(col, row, sx, sy)
        out = np.zeros((F, max(b[0], 1)), np.float32)
        self.lib.ref_clf_features(clf, fr[0], *b, _fp(out), out.shape[1])
        return out[:, :b[0]]

    def clf_update(self, clf, fr, pos, neg):
        bp, k1 = self._box(*pos); bn, k2 = self._box(*neg)
        if self.lib.ref_clf_update(clf, fr[0], *bp, *bn):
            raise RefAbort("the reference aborted inside StrongClassifierBase::Update (message on stderr)")

    def clf_classify(self, clf, fr, col, row, sx, sy, log_ratio=True):
        b, keep = self._box(col, row, sx, sy)
        out = np.zeros(max(b[0], 1), np.float32)
        self.lib.ref_clf_classify(clf, fr[0], *b, int(log_ratio), _fp(out))
        return out[:b[0]]

    def clf_selectors(self, clf):
        idx = np.zeros(self.lib.ref_clf_num_features(clf), np.int32)
        n = self.lib.ref_clf_get_selectors(clf, _ip(idx))
        return idx[:n].copy()

    def clf_alphas(self, clf):
        a = np.zeros(4096, np.float32)
        n = self.lib.ref_clf_get_alphas(clf, _fp(a))
        return a[:n].copy()

    def clf_weak_state(self, clf):
        F = self.lib.ref_clf_num_features(clf)
        out = np.zeros((8, F), np.float32)
        self.lib.ref_clf_get_weak_state(clf, _fp(out))
        return out

    def weak_run(self, weak_type, lr, n_updates, xpos, xneg, wpos, wneg, xtest):
        xpos = np.ascontiguousarray(xpos, np.float32); xneg = np.ascontiguousarray(xneg, np.float32)
        xtest = np.ascontiguousarray(xtest, np.float32)
        wp = None if wpos is None else np.ascontiguousarray(wpos, np.float32)
        wn = None if wneg is None else np.ascontiguousarray(wneg, np.float32)
        st = np.zeros(8, np.float32); out = np.zeros(len(xtest), np.float32); valid = C.c_int(0)
        self.lib.ref_weak_run(weak_type, lr, n_updates, _fp(xpos), len(xpos), _fp(xneg), len(xneg), _fp(wp), _fp(wn), _fp(st),
                              _fp(xtest), len(xtest), _fp(out), C.byref(valid))
        return st, out, valid.value

    # ---- samplers ----
    def sample_ring(self, fr, x, y, w, h, outer, inner, max_n, sx=1.0, sy=1.0, cap=65536):
        col = np.zeros(cap, np.int32); row = np.zeros(cap, np.int32)
        n = self.lib.ref_sample_ring(fr[0], x, y, w, h, outer, inner, max_n, sx, sy, _ip(col), _ip(row), cap)
        return col[:n].copy(), row[:n].copy()

    def sample_rect(self, fr, x, y, w, h, maxdx, maxdy, mindx, mindy, max_n, sx=1.0, sy=1.0, cap=65536):
        col = np.zeros(cap, np.int32); row = np.zeros(cap, np.int32)
        n = self.lib.ref_sample_rect(fr[0], x, y, w, h, maxdx, maxdy, mindx, mindy, max_n, sx, sy, _ip(col), _ip(row), cap)
        return col[:n].copy(), row[:n].copy()

    # ---- particle filter ----
    def pf_create(self, N, img_w, img_h, cx, cy, sx=1.0, sy=1.0, max_scale_change=0.5, max_scale=10.0, min_scale=0.1):
        return self.lib.ref_pf_create(N, img_w, img_h, cx, cy, sx, sy, max_scale_change, max_scale, min_scale), N

    def pf_get(self, pf):
        N = pf[1]
        st = np.zeros((N, 5), np.float32); rs = np.zeros((N, 5), np.float32); idx = np.zeros(N, np.int32); fs = C.c_int(0)
        self.lib.ref_pf_get(pf[0], _fp(st), _fp(rs), _ip(idx), C.byref(fs))
        return st, rs, idx, fs.value

    def pf_set(self, pf, state=None, resampled=None, filter_state=-1):
        st = None if state is None else np.ascontiguousarray(state, np.float32)
        rs = None if resampled is None else np.ascontiguousarray(resampled, np.float32)
        self.lib.ref_pf_set(pf[0], _fp(st), _fp(rs), filter_state)

    def pf_predict(self, pf, noise4xN, sd, w0, h0):
        nz = np.ascontiguousarray(noise4xN, np.float32)
        self.lib.ref_pf_predict(pf[0], _fp(nz), sd[0], sd[1], sd[2], sd[3], w0, h0)

    def pf_update_all_weights(self, pf, prob, only_nonzero=True, force_reset=False, resample=True):
        p = np.ascontiguousarray(prob, np.float32)
        self.lib.ref_pf_update_all_weights(pf[0], _fp(p), len(p), int(only_nonzero), int(force_reset), int(resample))

    def pf_average(self, pf):
        o = np.zeros(4, np.float32); self.lib.ref_pf_get_average(pf[0], _fp(o)); return o

    def pf_highest(self, pf):
        o = np.zeros(5, np.float32); self.lib.ref_pf_get_highest_weight(pf[0], _fp(o)); return o

    def pf_ordered_unique(self, pf):
        o = np.zeros((pf[1], 5), np.float32)
        n = self.lib.ref_pf_ordered_unique(pf[0], _fp(o), pf[1])
        return o[:n].copy()

    def pf_highest_unique_close(self, pf, closest):
        c = np.ascontiguousarray(closest, np.float32); o = np.zeros(5, np.float32)
        self.lib.ref_pf_get_highest_unique_close(pf[0], _fp(c), _fp(o)); return o

    # ---- tracker ----
    def tracker_create(self, particle_filter, tp, feature_type, n_haar, nb, weak_type, strong_type, K, lr, fr, init_xywh, n_frames,
                       use_hsv=0, pct_retained=0.0):
        init = np.ascontiguousarray(init_xywh, np.float32)
        t = self.lib.ref_tracker_create(int(particle_filter), C.byref(tp), feature_type, n_haar, nb, weak_type, strong_type, K, lr,
                                        use_hsv, pct_retained, fr[0], _fp(init), n_frames)
        if not t:
            raise RefAbort("the reference aborted inside InitializeTrackerWithParameters (message on stderr)")
        return t, tp.n_particles

    def tracker_track(self, t, fr, noise4xN, phases=3):
        nz = None if noise4xN is None else np.ascontiguousarray(noise4xN, np.float32)
        box = np.zeros(4, np.int32)
        if self.lib.ref_tracker_track(t[0], fr[0], _fp(nz), _ip(box), phases):
            raise RefAbort("the reference aborted inside TrackObjectAndSaveState (message on stderr)")
        return box

    def tracker_state(self, t):
        o = np.zeros(6, np.float32); self.lib.ref_tracker_state(t[0], _fp(o)); return o

    def tracker_pf(self, t):
        return self.lib.ref_tracker_pf(t[0]), t[1]

    def tracker_selectors(self, t, F):
        idx = np.zeros(F, np.int32)
        n = self.lib.ref_tracker_selectors(t[0], _ip(idx))
        return idx[:n].copy()

    def tracker_weak_state(self, t, F):
        out = np.zeros((8, F), np.float32)
        self.lib.ref_tracker_weak_state(t[0], _fp(out))
        return out


class NetCfg(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n_cameras", "n_objects", "n_frames", "feature_type", "n_haar", "n_bins", "use_hsv", "strong_type",
                                       "weak_type", "pct_selected", "pct_retained", "n_particles")] + \
               [(k, C.c_float) for k in ("sd_x", "sd_y", "sd_sx", "sd_sy")] + \
               [(k, C.c_int) for k in ("pos_radius_train", "init_pos_radius_train", "n_neg_train", "init_n_neg_train", "search_win",
                                       "neg_sample_strategy", "trajectory_option", "pos_strategy", "neg_strategy", "max_num_pos_examples",
                                       "geometric_fusion", "appearance_fusion", "af_strong", "af_weak", "af_pct_selected", "af_pct_retained",
                                       "af_refresh", "shared_stream")]


NET_DEFAULTS = dict(feature_type=0, n_haar=250, n_bins=8, use_hsv=0, strong_type=2, weak_type=0, pct_selected=20, pct_retained=0,
                    n_particles=1000, sd_x=20.0, sd_y=20.0, sd_sx=1e-7, sd_sy=1e-7, pos_radius_train=10, init_pos_radius_train=7,
                    n_neg_train=30, init_n_neg_train=30, search_win=20, neg_sample_strategy=1, trajectory_option=0, pos_strategy=0,
                    neg_strategy=0, max_num_pos_examples=30, geometric_fusion=0, appearance_fusion=0, af_strong=1, af_weak=0,
                    af_pct_selected=20, af_pct_retained=0, af_refresh=1, shared_stream=0)


class RefNetwork:
    """The reference's multi-camera frame loop (CameraNetwork.cpp:537-613) over its own Object / tracker / fuser classes, one MWC
    stream per (camera, object) pair (ref_harness.cpp, second half)."""

    def __init__(self, ref, n_cameras, n_objects, n_frames, homographies, init_xywh, seeds, **kw):
        L = ref.lib
        vp, i = C.c_void_p, C.c_int
        L.ref_net_create.argtypes = [C.POINTER(NetCfg), C.POINTER(C.c_double), _c_float_p, C.POINTER(C.c_int64)]; L.ref_net_create.restype = vp
        L.ref_net_free.argtypes = [vp]
        L.ref_net_init.argtypes = [vp, C.POINTER(vp)]
        L.ref_net_track.argtypes = [vp, i, C.POINTER(vp), _c_float_p]; L.ref_net_track.restype = i
        L.ref_net_init.restype = i
        L.ref_net_box.argtypes = [vp, i, i, i, _c_float_p]
        L.ref_net_pf.argtypes = [vp, i, i]; L.ref_net_pf.restype = vp
        L.ref_net_state.argtypes = [vp, i, i, _c_float_p]
        L.ref_net_selectors.argtypes = [vp, i, i, i, _c_int_p]; L.ref_net_selectors.restype = i
        L.ref_net_weak_state.argtypes = [vp, i, i, i, _c_float_p]
        L.ref_net_kalman.argtypes = [vp, i, _c_float_p, _c_float_p]; L.ref_net_kalman.restype = i
        L.ref_net_rng_state.argtypes = [vp, i, i]; L.ref_net_rng_state.restype = C.c_uint64
        L.ref_net_last_train.argtypes = [vp, i, i, i, _c_int_p, _c_int_p, i]; L.ref_net_last_train.restype = i
        self.ref, self.L = ref, L
        p = dict(NET_DEFAULTS); p.update(kw)
        cfg = NetCfg(n_cameras=n_cameras, n_objects=n_objects, n_frames=n_frames, **p)
        self.cfg, self.C, self.O, self.N = cfg, n_cameras, n_objects, p["n_particles"]
        H = None if homographies is None else np.ascontiguousarray(homographies, np.float64).reshape(n_cameras, 9)
        init = np.ascontiguousarray(init_xywh, np.float32).reshape(n_cameras, n_objects, 4)
        sd = np.ascontiguousarray(seeds, np.int64).reshape(n_cameras, n_objects)
        self.h = L.ref_net_create(C.byref(cfg), None if H is None else H.ctypes.data_as(C.POINTER(C.c_double)), _fp(init),
                                  sd.ctypes.data_as(C.POINTER(C.c_int64)))
        if not self.h:
            raise RefAbort("the reference aborted while the objects were created (message on stderr)")
        self.frames = None

    def _set_frames(self, planes):
        """planes: per camera (gray, c0, c1, c2)"""
        new = [self.ref.frame(*p) for p in planes]
        arr = (C.c_void_p * self.C)(*[f[0] for f in new])
        return new, arr

    def init(self, planes):
        self.frames, arr = self._set_frames(planes)
        if self.L.ref_net_init(self.h, arr):
            raise RefAbort("the reference aborted while initialising the trackers (message on stderr)")

    def track(self, frame_index, planes, noise):
        """noise: float32 [C][O][4][N]"""
        new, arr = self._set_frames(planes)
        nz = np.ascontiguousarray(noise, np.float32)
        assert nz.shape == (self.C, self.O, 4, self.N)
        rc = self.L.ref_net_track(self.h, frame_index, arr, _fp(nz))
        if rc:
            raise RefAbort("the reference aborted in frame %d (message on stderr)" % frame_index)
        for f in self.frames:
            self.ref.frame_free(f)
        self.frames = new

    def box(self, c, o, frame):
        out = np.zeros(4, np.float32); self.L.ref_net_box(self.h, c, o, frame, _fp(out)); return out

    def particles(self, c, o):
        return self.ref.pf_get((self.L.ref_net_pf(self.h, c, o), self.N))

    def state(self, c, o):
        out = np.zeros(6, np.float32); self.L.ref_net_state(self.h, c, o, _fp(out)); return out

    def selectors(self, c, o, global_clf=False):
        idx = np.zeros(4096, np.int32)
        n = self.L.ref_net_selectors(self.h, c, o, int(global_clf), _ip(idx))
        return idx[:n].copy()

    def weak_state(self, c, o, F, global_clf=False):
        out = np.zeros((8, F), np.float32); self.L.ref_net_weak_state(self.h, c, o, int(global_clf), _fp(out)); return out

    def kalman(self, o):
        st = np.zeros(4, np.float32); P = np.zeros(16, np.float32)
        self.L.ref_net_kalman(self.h, o, _fp(st), _fp(P))
        return st, P.reshape(4, 4)

    def rng_state(self, c, o):
        return int(self.L.ref_net_rng_state(self.h, c, o))

    def last_train(self, c, o, which, cap=65536):
        col = np.zeros(cap, np.int32); row = np.zeros(cap, np.int32)
        n = self.L.ref_net_last_train(self.h, c, o, which, _ip(col), _ip(row), cap)
        return col[:n].copy(), row[:n].copy()

    def close(self):
        if self.h:
            self.L.ref_net_free(self.h); self.h = None
        if self.frames:
            for f in self.frames:
                self.ref.frame_free(f)
            self.frames = None
