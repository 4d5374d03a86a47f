"""ctypes binding of the CPU oracle (oracle/libmct_oracle*.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_int_p = C.POINTER(C.c_int)
_c_float_p = C.POINTER(C.c_float)
_c_u8_p = C.POINTER(C.c_uint8)


def build(force=False):
    so = os.path.join(_HERE, "libmct_oracle.so")
    omp = os.path.join(_HERE, "libmct_oracle_omp.so")
    src = max(os.path.getmtime(os.path.join(_HERE, "mct_oracle.c")), os.path.getmtime(os.path.join(_HERE, "mct_oracle.h")))
    if force or not os.path.exists(so) or os.path.getmtime(so) < src:
        subprocess.run(["make", "-C", _HERE, "-B", "libmct_oracle.so"], check=True, capture_output=True)
    if force or not os.path.exists(omp) or os.path.getmtime(omp) < src:
        subprocess.run(["make", "-C", _HERE, "-B", "libmct_oracle_omp.so"], check=False, capture_output=True)
    return so


class Rng(C.Structure):
    _fields_ = [("state", C.c_uint64)]


class Boxes(C.Structure):
    _fields_ = [("n", C.c_int), ("col", _c_int_p), ("row", _c_int_p), ("sx", _c_float_p), ("sy", _c_float_p)]


class HaarTable(C.Structure):
    _fields_ = [("F", C.c_int), ("nrect", _c_int_p), ("rect", _c_int_p), ("weight", _c_float_p), ("channel", _c_int_p)]


class Frame(C.Structure):
    _fields_ = [("gray", _c_u8_p), ("c0", _c_u8_p), ("c1", _c_u8_p), ("c2", _c_u8_p),
                ("W", C.c_int), ("H", C.c_int), ("pitch", C.c_int), ("ii", _c_float_p)]


class TrackerParams(C.Structure):
    _fields_ = [("n_particles", C.c_int), ("sd_x", C.c_float), ("sd_y", C.c_float), ("sd_sx", C.c_float),
                ("sd_sy", C.c_float), ("pos_radius_train", C.c_float), ("init_pos_radius_train", C.c_float),
                ("n_neg_train", C.c_int), ("init_n_neg_train", C.c_int), ("max_pos_train", C.c_int),
                ("search_win", C.c_int), ("trajectory_option", C.c_int), ("pos_strategy", C.c_int),
                ("neg_strategy", C.c_int), ("max_num_pos_examples", C.c_int)]


class Pf(C.Structure):
    _fields_ = [("N", C.c_int), ("img_w", C.c_float), ("img_h", C.c_float), ("max_scale_change", C.c_float),
                ("max_scale", C.c_float), ("min_scale", C.c_float), ("state", _c_float_p), ("resampled", _c_float_p),
                ("resample_idx", _c_int_p), ("filter_state", C.c_int), ("time_index", C.c_int),
                ("uniq", _c_float_p), ("n_uniq", C.c_int), ("uniq_valid", C.c_int), ("last_resampled", C.c_int),
                ("last_neff_inv", C.c_float)]


def _fp(a):
    return a.ctypes.data_as(_c_float_p)


def _ip(a):
    return a.ctypes.data_as(_c_int_p)


def _up(a):
    return None if a is None else a.ctypes.data_as(_c_u8_p)


class Oracle:
    """Thin wrapper; `omp=True` loads the pragma-honouring build."""

    def __init__(self, omp=False):
        build()
        name = "libmct_oracle_omp.so" if omp else "libmct_oracle.so"
        self.lib = L = C.CDLL(os.path.join(_HERE, name))
        self.omp = omp
        L.orc_rng_seed.argtypes = [C.POINTER(Rng), C.c_int64]
        L.orc_rng_next.argtypes = [C.POINTER(Rng)]; L.orc_rng_next.restype = C.c_uint32
        L.orc_randint.argtypes = [C.POINTER(Rng), C.c_int, C.c_int]; L.orc_randint.restype = C.c_int
        L.orc_randfloat.argtypes = [C.POINTER(Rng)]; L.orc_randfloat.restype = C.c_float
        L.orc_cvround.argtypes = [C.c_double]; L.orc_cvround.restype = C.c_int
        L.orc_integral.argtypes = [_c_u8_p, C.c_int, C.c_int, C.c_int, _c_float_p]
        L.orc_sum_rect.argtypes = [_c_float_p, C.c_int] + [C.c_int] * 4; L.orc_sum_rect.restype = C.c_float
        L.orc_haar_generate.argtypes = [C.POINTER(Rng), C.c_int, C.c_int, C.c_int]
        L.orc_haar_generate.restype = C.POINTER(HaarTable)
        L.orc_haar_free.argtypes = [C.POINTER(HaarTable)]
        L.orc_haar_compute.argtypes = [C.POINTER(HaarTable), _c_float_p, C.c_int, C.POINTER(Boxes), _c_float_p, C.c_int]
        L.orc_mdch_compute.argtypes = [_c_u8_p, _c_u8_p, _c_u8_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(Boxes), _c_float_p, C.c_int]
        L.orc_cch_compute.argtypes = [_c_u8_p, C.c_int, C.c_int, C.c_int, C.POINTER(Boxes), C.c_int, _c_float_p, C.c_int]
        L.orc_clf_create.argtypes = [C.c_int] * 8 + [C.c_float, C.POINTER(HaarTable)]
        L.orc_clf_create.restype = C.c_void_p
        L.orc_clf_free.argtypes = [C.c_void_p]
        L.orc_clf_set_retained.argtypes = [C.c_void_p, C.c_float]; L.orc_clf_set_retained.restype = None
        L.orc_clf_num_features.argtypes = [C.c_void_p]; L.orc_clf_num_features.restype = C.c_int
        L.orc_clf_features.argtypes = [C.c_void_p, C.POINTER(Frame), C.POINTER(Boxes), _c_float_p, C.c_int]
        L.orc_clf_update.argtypes = [C.c_void_p, C.POINTER(Frame), C.POINTER(Boxes), C.POINTER(Boxes), _c_float_p, _c_float_p]
        L.orc_clf_update_from_features.argtypes = [C.c_void_p, _c_float_p, C.c_int, _c_float_p, C.c_int, _c_float_p, _c_float_p]
        L.orc_clf_classify.argtypes = [C.c_void_p, C.POINTER(Frame), C.POINTER(Boxes), C.c_int, _c_float_p]
        L.orc_clf_classify_from_features.argtypes = [C.c_void_p, _c_float_p, C.c_int, C.c_int, C.c_int, _c_float_p]
        L.orc_clf_get_selectors.argtypes = [C.c_void_p, _c_int_p]; L.orc_clf_get_selectors.restype = C.c_int
        L.orc_clf_get_weak_state.argtypes = [C.c_void_p, _c_float_p]
        L.orc_clf_get_alphas.argtypes = [C.c_void_p, _c_float_p]; L.orc_clf_get_alphas.restype = C.c_int
        L.orc_clf_set_threads.argtypes = [C.c_int]
        L.orc_likelihood_map.argtypes = [_c_float_p, C.c_int, C.c_int]
        L.orc_pf_create.argtypes = [C.c_int, C.c_float, C.c_float]; L.orc_pf_create.restype = C.POINTER(Pf)
        L.orc_pf_free.argtypes = [C.POINTER(Pf)]
        L.orc_pf_initialize.argtypes = [C.POINTER(Pf)] + [C.c_float] * 7
        L.orc_pf_predict.argtypes = [C.POINTER(Pf), _c_float_p, C.c_float, C.c_float]
        L.orc_pf_update_all_weights.argtypes = [C.POINTER(Pf), _c_float_p, C.c_int, C.c_int, C.c_int, C.POINTER(Rng)]
        L.orc_pf_update_all_weights.restype = C.c_int
        L.orc_pf_normalize.argtypes = [C.POINTER(Pf)]
        L.orc_pf_resample_if_required.argtypes = [C.POINTER(Pf), C.POINTER(Rng)]; L.orc_pf_resample_if_required.restype = C.c_int
        L.orc_pf_resample_u01.argtypes = [C.POINTER(Pf), C.c_int, C.c_float]
        L.orc_pf_find_ordered_unique.argtypes = [C.POINTER(Pf)]
        L.orc_pf_get_average.argtypes = [C.POINTER(Pf), _c_float_p]
        L.orc_pf_get_highest_weight.argtypes = [C.POINTER(Pf), _c_float_p]
        L.orc_pf_get_highest_unique_close.argtypes = [C.POINTER(Pf), _c_float_p, _c_float_p]
        L.orc_pf_generate_test_set.argtypes = [C.POINTER(Pf), C.c_float, C.c_float, C.c_int, C.c_int,
                                               _c_int_p, _c_int_p, _c_float_p, _c_float_p]
        L.orc_pf_generate_test_set.restype = C.c_int
        L.orc_sample_ring.argtypes = [C.POINTER(Rng)] + [C.c_int] * 6 + [C.c_float, C.c_float, C.c_int, C.c_float, C.c_float,
                                                                      _c_int_p, _c_int_p, C.c_int]
        L.orc_sample_ring.restype = C.c_int
        L.orc_sample_rect.argtypes = [C.POINTER(Rng)] + [C.c_int] * 6 + [C.c_float] * 4 + [C.c_int, C.c_float, C.c_float,
                                                                                      _c_int_p, _c_int_p, C.c_int]
        L.orc_sample_rect.restype = C.c_int
        L.orc_tracker_create.argtypes = [C.POINTER(TrackerParams), C.c_void_p, C.POINTER(Frame), _c_float_p, C.POINTER(Rng)]
        L.orc_tracker_create.restype = C.c_void_p
        L.orc_tracker_free.argtypes = [C.c_void_p]
        L.orc_simple_create.argtypes = [C.POINTER(TrackerParams), C.c_void_p, C.POINTER(Frame), _c_float_p, C.POINTER(Rng)]
        L.orc_simple_create.restype = C.c_void_p
        L.orc_simple_free.argtypes = [C.c_void_p]
        L.orc_simple_track.argtypes = [C.c_void_p, C.POINTER(Frame), _c_float_p]; L.orc_simple_track.restype = C.c_double
        L.orc_simple_last_n.argtypes = [C.c_void_p]; L.orc_simple_last_n.restype = C.c_int
        L.orc_simple_last_best.argtypes = [C.c_void_p]; L.orc_simple_last_best.restype = C.c_int
        L.orc_tracker_track.argtypes = [C.c_void_p, C.POINTER(Frame), _c_float_p, _c_int_p, C.c_int]
        L.orc_tracker_last_valid.argtypes = [C.c_void_p]; L.orc_tracker_last_valid.restype = C.c_int
        L.orc_tracker_pf.argtypes = [C.c_void_p]; L.orc_tracker_pf.restype = C.POINTER(Pf)
        L.orc_tracker_state.argtypes = [C.c_void_p]; L.orc_tracker_state.restype = _c_float_p
        L.orc_tracker_last_train.argtypes = [C.c_void_p, C.c_int, _c_int_p, _c_int_p, _c_float_p, _c_float_p, C.c_int]
        L.orc_tracker_last_train.restype = C.c_int

    # ---- helpers -------------------------------------------------------------------------
    def rng(self, seed=1):
        r = Rng()
        self.lib.orc_rng_seed(C.byref(r), seed)
        return r

    def randfloat(self, r):
        return self.lib.orc_randfloat(C.byref(r))

    def randint(self, r, a, b):
        return self.lib.orc_randint(C.byref(r), a, b)

    def bgr_to_planes(self, bgr, use_hsv=False):
        """packed BGR [H][W][3] uint8 -> (gray, c0, c1, c2) planes: R,G,B or H,S,V"""
        bgr = np.ascontiguousarray(bgr, np.uint8)
        H, W = bgr.shape[:2]
        out = [np.empty((H, W), np.uint8) for _ in range(4)]
        self.lib.orc_bgr_to_planes(_up(bgr), W, H, 3 * W, int(use_hsv), *[_up(o) for o in out])
        return out

    def integral(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        H, W = img.shape
        out = np.empty((H + 1, W + 1), np.float32)
        self.lib.orc_integral(_up(img), W, H, W, _fp(out))
        return out

    def sum_rect(self, ii, x, y, w, h):
        return self.lib.orc_sum_rect(_fp(ii), ii.shape[1] - 1, x, y, w, h)

    @staticmethod
    def boxes(col, row, sx, sy):
        col = np.ascontiguousarray(col, np.int32); row = np.ascontiguousarray(row, np.int32)
        sx = np.ascontiguousarray(sx, np.float32); sy = np.ascontiguousarray(sy, np.float32)
        b = Boxes(len(col), _ip(col), _ip(row), _fp(sx), _fp(sy))
        b._keep = (col, row, sx, sy)
        return b

    def haar_generate(self, r, F, w0, h0):
        return self.lib.orc_haar_generate(C.byref(r), F, w0, h0)

    @staticmethod
    def haar_arrays(t):
        F = t.contents.F
        nrect = np.ctypeslib.as_array(t.contents.nrect, (F,)).copy()
        rect = np.ctypeslib.as_array(t.contents.rect, (F, 6, 4)).copy()
        weight = np.ctypeslib.as_array(t.contents.weight, (F, 6)).copy()
        return nrect.astype(np.int32), rect.astype(np.int32), weight.astype(np.float32)

    def frame(self, gray, c0=None, c1=None, c2=None, ii=None):
        gray = np.ascontiguousarray(gray, np.uint8)
        H, W = gray.shape
        if ii is None:
            ii = self.integral(gray)
        planes = [None if p is None else np.ascontiguousarray(p, np.uint8) for p in (c0, c1, c2)]
        f = Frame(_up(gray), _up(planes[0]), _up(planes[1]), _up(planes[2]), W, H, W, _fp(ii))
        f._keep = (gray, planes, ii)
        return f

    def clf_create(self, feature_type, n_haar, nb, w0, h0, weak_type, strong_type, K, lr, haar, pct_retained=0.0):
        c = self.lib.orc_clf_create(feature_type, n_haar, nb, w0, h0, weak_type, strong_type, K, lr, haar)
        if pct_retained:
            self.lib.orc_clf_set_retained(c, pct_retained)
        return c

    def clf_features(self, clf, frame, boxes):
        F = self.lib.orc_clf_num_features(clf)
        out = np.zeros((F, max(boxes.n, 1)), np.float32)
        self.lib.orc_clf_features(clf, C.byref(frame), C.byref(boxes), _fp(out), out.shape[1])
        return out[:, :boxes.n]

    def clf_update(self, clf, frame, pos, neg, wpos=None, wneg=None):
        wp = None if wpos is None else _fp(np.ascontiguousarray(wpos, np.float32))
        wn = None if wneg is None else _fp(np.ascontiguousarray(wneg, np.float32))
        self.lib.orc_clf_update(clf, C.byref(frame), C.byref(pos), C.byref(neg), wp, wn)

    def clf_update_from_features(self, clf, fpos, fneg, wpos=None, wneg=None):
        fpos = np.ascontiguousarray(fpos, np.float32); fneg = np.ascontiguousarray(fneg, np.float32)
        wp = None if wpos is None else _fp(np.ascontiguousarray(wpos, np.float32))
        wn = None if wneg is None else _fp(np.ascontiguousarray(wneg, np.float32))
        self.lib.orc_clf_update_from_features(clf, _fp(fpos), fpos.shape[1], _fp(fneg), fneg.shape[1], wp, wn)

    def clf_classify(self, clf, frame, boxes, log_ratio=True):
        out = np.zeros(max(boxes.n, 1), np.float32)
        self.lib.orc_clf_classify(clf, C.byref(frame), C.byref(boxes), int(log_ratio), _fp(out))
        return out[:boxes.n]

    def clf_classify_from_features(self, clf, feat, log_ratio=True):
        feat = np.ascontiguousarray(feat, np.float32)
        out = np.zeros(feat.shape[1], np.float32)
        self.lib.orc_clf_classify_from_features(clf, _fp(feat), feat.shape[1], feat.shape[1], int(log_ratio), _fp(out))
        return out

    def clf_selectors(self, clf):
        F = self.lib.orc_clf_num_features(clf)
        idx = np.zeros(F, np.int32)
        n = self.lib.orc_clf_get_selectors(clf, _ip(idx))
        return idx[:n].copy()

    def clf_alphas(self, clf):
        a = np.zeros(4096, np.float32)
        n = self.lib.orc_clf_get_alphas(clf, _fp(a))
        return a[:n].copy()

    def clf_weak_state(self, clf):
        F = self.lib.orc_clf_num_features(clf)
        out = np.zeros((8, F), np.float32)
        self.lib.orc_clf_get_weak_state(clf, _fp(out))
        return out

    def likelihood_map(self, h, exponent=20):
        h = np.array(h, np.float32)
        self.lib.orc_likelihood_map(_fp(h), len(h), exponent)
        return h

    # particle filter
    def pf_create(self, N, w, h):
        return self.lib.orc_pf_create(N, w, h)

    @staticmethod
    def pf_state(pf):
        return np.ctypeslib.as_array(pf.contents.state, (pf.contents.N, 5))

    @staticmethod
    def pf_resampled(pf):
        return np.ctypeslib.as_array(pf.contents.resampled, (pf.contents.N, 5))

    @staticmethod
    def pf_resample_idx(pf):
        return np.ctypeslib.as_array(pf.contents.resample_idx, (pf.contents.N,))

    def pf_test_set(self, pf, w0, h0, fw, fh):
        N = pf.contents.N
        col = np.zeros(N, np.int32); row = np.zeros(N, np.int32)
        sx = np.zeros(N, np.float32); sy = np.zeros(N, np.float32)
        n = self.lib.orc_pf_generate_test_set(pf, w0, h0, fw, fh, _ip(col), _ip(row), _fp(sx), _fp(sy))
        return col[:n], row[:n], sx[:n], sy[:n]

    def sample_ring(self, r, img_w, img_h, x, y, w, h, outer, inner, max_n, sx=1.0, sy=1.0, cap=16384):
        col = np.zeros(cap, np.int32); row = np.zeros(cap, np.int32)
        n = self.lib.orc_sample_ring(C.byref(r), img_w, img_h, x, y, w, h, outer, inner, max_n, sx, sy, _ip(col), _ip(row), cap)
        return col[:n].copy(), row[:n].copy()

    def sample_rect(self, r, img_w, img_h, x, y, w, h, maxdx, maxdy, mindx, mindy, max_n, sx=1.0, sy=1.0, cap=16384):
        col = np.zeros(cap, np.int32); row = np.zeros(cap, np.int32)
        n = self.lib.orc_sample_rect(C.byref(r), img_w, img_h, x, y, w, h, maxdx, maxdy, mindx, mindy, max_n, sx, sy,
                                     _ip(col), _ip(row), cap)
        return col[:n].copy(), row[:n].copy()
