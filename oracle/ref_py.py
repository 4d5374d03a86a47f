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
        L.ref_clf_update.argtypes = [vp, vp] + box + box
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
        L.ref_tracker_track.argtypes = [vp, vp, _c_float_p, _c_int_p, i]
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
        self.lib.ref_clf_update(clf, fr[0], *bp, *bn)

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
        return t, tp.n_particles

    def tracker_track(self, t, fr, noise4xN, phases=3):
        nz = None if noise4xN is None else np.ascontiguousarray(noise4xN, np.float32)
        box = np.zeros(4, np.int32)
        self.lib.ref_tracker_track(t[0], fr[0], _fp(nz), _ip(box), phases)
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
