"""ctypes binding of the C-ABI in include/mct_b200.h (libmct_b200.so).

This is the only way Python reaches the GPU path; there is no Python/NumPy/torch fallback.  If the
shared library is missing it is built with nvcc; if that fails, or no CUDA device is present, every
compute call raises MctError.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_c_int_p = C.POINTER(C.c_int)
_c_float_p = C.POINTER(C.c_float)
_c_u8_p = C.POINTER(C.c_uint8)

OK, E_ARG, E_CUDA, E_STATE, E_EMPTY, E_NOMEM, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
FEAT_HAAR, FEAT_CCH, FEAT_MDCH, FEAT_HAAR_MDCH = 0, 1, 2, 3
WEAK_STUMP, WEAK_WSTUMP, WEAK_PERCEPTRON = 0, 1, 2
STRONG_ADABOOST, STRONG_MILBOOST, STRONG_MILENSEMBLE, STRONG_MILANYBOOST = 0, 2, 3, 4
PF_UNINIT, PF_INIT, PF_PREDICTED, PF_RESAMPLED = 0, 1, 2, 3

# every symbol include/mct_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "mct_ctx_create", "mct_ctx_destroy", "mct_sync", "mct_last_error", "mct_version", "mct_launch_count",
    "mct_ctx_stream", "mct_profile_enable", "mct_profile_read", "mct_scored_particles", "mct_frame_set", "mct_frame_prefetch", "mct_frame_set_bgr", "mct_frame_prefetch_bgr", "mct_frame_get_planes", "mct_frame_get_integral", "mct_sum_rects", "mct_classifier_create",
    "mct_classifier_destroy", "mct_classifier_num_features", "mct_features_compute", "mct_classifier_update",
    "mct_classifier_update_from_features", "mct_classifier_classify", "mct_classifier_classify_from_features",
    "mct_classifier_get_selectors", "mct_classifier_get_weak_state", "mct_classifier_get_predictions",
    "mct_classifier_set_cluster", "mct_pf_create", "mct_pf_destroy", "mct_pf_initialize", "mct_pf_predict",
    "mct_pf_update_all_weights", "mct_pf_resample", "mct_pf_get_state", "mct_pf_set_state", "mct_pf_get_resampled",
    "mct_pf_get_resample_indices", "mct_pf_get_summary", "mct_track_score", "mct_noise_prefetch", "mct_track_score_async",
    "mct_track_collect", "mct_track_update", "mct_pf_get_last_response", "mct_fuser_pack_foot_points",
    "mct_pf_ground_pdf", "mct_pf_rearrange_ground", "mct_classify_argmax", "mct_fuser_ground_pdf", "mct_p2p_create", "mct_p2p_open", "mct_p2p_destroy",
    "mct_fuser_exchange_foot_points", "mct_classifier_get_alphas", "mct_features_compute_dev", "mct_classifier_update_from_features_dev",
    "mct_pf_training_boxes", "mct_sample_regions",
    # called by the C++ host schedule (host/mct_network.cpp), listed so that load() checks they are exported
    "mct_classifier_update_from_gathered", "mct_dev_alloc", "mct_dev_copy", "mct_dev_free", "mct_fuser_foot_points",
    "mct_pf_resample_many", "mct_track_update_from_gathered", "mct_train_features_many_dev",
    "mct_pf_get_summary_many", "mct_pf_rearrange_ground_many",
    # whole frame on the device
    "mct_pf_rng_set", "mct_pf_rng_get", "mct_track_update_default", "mct_track_update_default_collect",
    "mct_track_update_default_boxes", "mct_frame_prefetch_flush", "mct_abi_struct_sizes",
]


class MctError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mct error %d: %s" % (code, msg))
        self.code = code


class Boxes(C.Structure):
    _fields_ = [("n", C.c_int), ("col", _c_int_p), ("row", _c_int_p), ("sx", _c_float_p), ("sy", _c_float_p)]


class ClfParams(C.Structure):
    _fields_ = [("feature_type", C.c_int), ("n_haar", C.c_int), ("n_bins", C.c_int), ("w0", C.c_int), ("h0", C.c_int),
                ("weak_type", C.c_int), ("strong_type", C.c_int), ("n_selected", C.c_int), ("learning_rate", C.c_float),
                ("cch_mode", C.c_int), ("max_train", C.c_int), ("max_test", C.c_int), ("pct_retained", C.c_float)]


class TrainDefault(C.Structure):
    _fields_ = [("w0", C.c_float), ("h0", C.c_float), ("trajectory_option", C.c_int), ("pos_radius", C.c_float), ("max_pos", C.c_int),
                ("neg_outer", C.c_float), ("neg_inner", C.c_float), ("max_neg", C.c_int), ("scale_hint_x", C.c_float),
                ("scale_hint_y", C.c_float)]


class TrainDefaultOut(C.Structure):
    _fields_ = [("P", C.c_int), ("Nn", C.c_int), ("status", C.c_int), ("mdch_fast", C.c_int), ("rng_state", C.c_uint64),
                ("cur", C.c_float * 6), ("n_draws", C.c_int), ("pad", C.c_int)]


class TrainGen(C.Structure):
    _fields_ = [("strategy", C.c_int), ("max_pos", C.c_int), ("cur", C.c_float * 6), ("frame_w", C.c_int), ("frame_h", C.c_int),
                ("rng_state", C.c_uint64)]


class TrainGenOut(C.Structure):
    _fields_ = [("n_pos", C.c_int), ("n_unique", C.c_int), ("max_dx", C.c_float), ("max_dy", C.c_float), ("rng_state", C.c_uint64),
                ("n_draws", C.c_int), ("pad", C.c_int)]


class SampleRegion(C.Structure):
    _fields_ = [("kind", C.c_int), ("x", C.c_int), ("y", C.c_int), ("width", C.c_int), ("height", C.c_int), ("p", C.c_float * 4),
                ("max_n", C.c_int), ("scale_x", C.c_float), ("scale_y", C.c_float), ("frame_w", C.c_int), ("frame_h", C.c_int),
                ("rng_state", C.c_uint64)]


class SampleRegionOut(C.Structure):
    _fields_ = [("n", C.c_int), ("n_candidates", C.c_int), ("n_draws", C.c_int), ("pad", C.c_int), ("rng_state", C.c_uint64)]


class HaarTable(C.Structure):
    _fields_ = [("n_features", C.c_int), ("nrect", _c_int_p), ("rects", _c_int_p), ("weights", _c_float_p)]


class GroundPdfResult(C.Structure):
    _fields_ = [("total_weight", C.c_float), ("average", C.c_float * 4), ("usable_covariance", C.c_int)]


class PfSummary(C.Structure):
    _fields_ = [("n_valid", C.c_int), ("resampled", C.c_int), ("filter_state", C.c_int), ("n_unique", C.c_int),
                ("neff_inv", C.c_float), ("max_response", C.c_float), ("average", C.c_float * 4),
                ("ordered_first", C.c_float * 5), ("highest_close", C.c_float * 5)]


class TrackParams(C.Structure):
    _fields_ = [("w0", C.c_float), ("h0", C.c_float), ("exponent", C.c_int), ("force_reset", C.c_int),
                ("resample", C.c_int)]


_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (building first if needed) libmct_b200.so.  Raises if it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    L = C.CDLL(path)
    vp = C.c_void_p
    L.mct_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.mct_ctx_destroy.argtypes = [vp]
    L.mct_sync.argtypes = [vp]
    L.mct_last_error.argtypes = [vp]; L.mct_last_error.restype = C.c_char_p
    L.mct_version.restype = C.c_char_p
    L.mct_launch_count.argtypes = [vp]; L.mct_launch_count.restype = C.c_longlong
    L.mct_ctx_stream.argtypes = [vp]; L.mct_ctx_stream.restype = vp
    L.mct_profile_enable.argtypes = [vp, C.c_int]
    L.mct_profile_read.argtypes = [vp, C.c_int, C.c_char_p, _c_float_p, _c_int_p, _c_int_p]
    L.mct_scored_particles.argtypes = [vp, C.POINTER(C.c_ulonglong), C.c_int]
    L.mct_frame_set.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mct_frame_prefetch.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.mct_frame_set_bgr.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mct_frame_prefetch_bgr.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mct_frame_get_planes.argtypes = [vp, vp, vp, vp, vp]
    L.mct_frame_get_integral.argtypes = [vp, _c_float_p]
    L.mct_sum_rects.argtypes = [vp, _c_int_p, C.c_int, _c_float_p]
    L.mct_classifier_create.argtypes = [vp, C.POINTER(ClfParams), C.POINTER(HaarTable), C.POINTER(vp)]
    L.mct_classifier_destroy.argtypes = [vp]
    L.mct_classifier_num_features.argtypes = [vp]
    L.mct_features_compute.argtypes = [vp, C.POINTER(Boxes), _c_float_p]
    L.mct_classifier_update.argtypes = [vp, C.POINTER(Boxes), C.POINTER(Boxes), _c_float_p, _c_float_p]
    L.mct_classifier_update_from_features.argtypes = [vp, _c_float_p, C.c_int, _c_float_p, C.c_int, _c_float_p, _c_float_p]
    L.mct_classifier_classify.argtypes = [vp, C.POINTER(Boxes), C.c_int, _c_float_p]
    L.mct_classifier_classify_from_features.argtypes = [vp, _c_float_p, C.c_int, C.c_int, _c_float_p]
    L.mct_classifier_get_selectors.argtypes = [vp, _c_int_p, _c_int_p]
    L.mct_classifier_get_weak_state.argtypes = [vp, _c_float_p]
    L.mct_classifier_get_predictions.argtypes = [vp, _c_float_p, _c_int_p, _c_int_p]
    L.mct_classifier_set_cluster.argtypes = [vp, C.c_int]
    L.mct_pf_create.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.POINTER(vp)]
    L.mct_pf_destroy.argtypes = [vp]
    L.mct_pf_initialize.argtypes = [vp] + [C.c_float] * 7
    L.mct_pf_predict.argtypes = [vp, vp, C.c_int, C.c_float, C.c_float]
    L.mct_pf_update_all_weights.argtypes = [vp, _c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _c_int_p]
    L.mct_pf_resample.argtypes = [vp, C.c_int, C.c_float]
    L.mct_pf_get_state.argtypes = [vp, _c_float_p]
    L.mct_pf_set_state.argtypes = [vp, _c_float_p, C.c_int]
    L.mct_pf_get_resampled.argtypes = [vp, _c_float_p]
    L.mct_pf_get_resample_indices.argtypes = [vp, _c_int_p]
    L.mct_pf_get_summary.argtypes = [vp, C.POINTER(PfSummary)]
    L.mct_track_score.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(TrackParams), vp, C.c_int,
                                  _c_float_p, C.POINTER(PfSummary)]
    L.mct_track_score_async.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(TrackParams), vp, C.c_int,
                                        _c_float_p]
    L.mct_noise_prefetch.argtypes = [vp, vp, C.c_size_t]
    L.mct_track_collect.argtypes = [vp, C.c_int, C.POINTER(PfSummary)]
    L.mct_track_update.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(Boxes), C.POINTER(Boxes)]
    L.mct_pf_get_last_response.argtypes = [vp, _c_float_p, _c_int_p]
    L.mct_fuser_pack_foot_points.argtypes = [vp, C.c_int, C.POINTER(vp), _c_float_p, vp]
    L.mct_pf_ground_pdf.argtypes = [vp, C.POINTER(C.c_double), _c_float_p, _c_float_p, C.c_float, C.POINTER(GroundPdfResult), _c_float_p]
    L.mct_pf_rearrange_ground.argtypes = [vp, C.c_float, C.c_float, _c_float_p, C.c_float, C.c_float]
    L.mct_fuser_ground_pdf.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_double), _c_float_p, _c_float_p, _c_float_p, C.POINTER(GroundPdfResult)]
    L.mct_p2p_create.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.mct_p2p_open.argtypes = [vp, C.c_char_p]
    L.mct_p2p_destroy.argtypes = [vp]
    L.mct_fuser_exchange_foot_points.argtypes = [vp, C.c_int, C.POINTER(vp), _c_float_p, _c_float_p, C.c_int]
    L.mct_features_compute_dev.argtypes = [vp, C.POINTER(Boxes), vp, C.c_int]
    L.mct_classifier_update_from_features_dev.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int]
    L.mct_classifier_get_alphas.argtypes = [vp, _c_float_p, _c_int_p]
    L.mct_classify_argmax.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(Boxes), C.c_int, _c_int_p, _c_float_p]
    L.mct_pf_training_boxes.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(TrainGen), C.c_int, _c_int_p, _c_int_p, _c_float_p, _c_float_p,
                                        C.POINTER(TrainGenOut)]
    L.mct_sample_regions.argtypes = [vp, C.c_int, C.POINTER(SampleRegion), C.c_int, _c_int_p, _c_int_p, C.POINTER(SampleRegionOut)]
    L.mct_pf_rng_set.argtypes = [vp, C.c_uint64]
    L.mct_pf_rng_get.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mct_track_update_default.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(TrainDefault)]
    L.mct_track_update_default_collect.argtypes = [vp, C.c_int, C.POINTER(TrainDefaultOut)]
    L.mct_track_update_default_boxes.argtypes = [vp, C.c_int, C.c_int, _c_int_p, _c_int_p, _c_float_p, _c_float_p,
                                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(_c_float_p)


def _ip(a):
    return a.ctypes.data_as(_c_int_p)


def make_boxes(col, row, sx, sy):
    col = np.ascontiguousarray(col, np.int32); row = np.ascontiguousarray(row, np.int32)
    sx = np.ascontiguousarray(sx, np.float32); sy = np.ascontiguousarray(sy, np.float32)
    if not (len(col) == len(row) == len(sx) == len(sy)):
        raise ValueError("box arrays differ in length")
    b = Boxes(len(col), _ip(col), _ip(row), _fp(sx), _fp(sy))
    b._keep = (col, row, sx, sy)
    return b


def profile_enable(L, h, on):
    L.mct_profile_enable(h, int(on))


def profile_read(L, h, max_kernels=64):
    """-> {kernel name: (total_ms, launches)} since the last read (synchronises the stream)"""
    names = C.create_string_buffer(48 * max_kernels)
    ms = np.zeros(max_kernels, np.float32); cnt = np.zeros(max_kernels, np.int32); n = C.c_int(0)
    rc = L.mct_profile_read(h, max_kernels, names, _fp(ms), _ip(cnt), C.byref(n))
    if rc:
        raise MctError(rc, "mct_profile_read")
    out = {}
    for i in range(n.value):
        nm = names.raw[48 * i:48 * (i + 1)].split(b"\0")[0].decode()
        out[nm] = (float(ms[i]), int(cnt[i]))
    return out


def scored_particles(L, h, reset=False):
    v = C.c_ulonglong(0)
    rc = L.mct_scored_particles(h, C.byref(v), int(reset))
    if rc:
        raise MctError(rc, "mct_scored_particles")
    return int(v.value)


class Context:
    """One camera == one GPU == one CUDA stream (mct_ctx)."""

    def __init__(self, device=0, stream=None):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.mct_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc:
            raise MctError(rc, self.L.mct_last_error(None).decode())
        self.h = h
        self.W = self.H = 0
        self._keep = None

    @classmethod
    def from_handle(cls, h):
        """non-owning view of an existing mct_ctx (e.g. the one a host.Camera created)"""
        self = cls.__new__(cls)
        self.L = load()
        self.h = h
        self.W = self.H = 0
        self._keep = None
        self._borrowed = True
        return self

    def check(self, rc):
        if rc:
            raise MctError(rc, self.L.mct_last_error(self.h).decode())

    def close(self):
        if self.h and not getattr(self, "_borrowed", False):
            self.L.mct_ctx_destroy(self.h)
        self.h = None

    def sync(self):
        self.check(self.L.mct_sync(self.h))

    @property
    def launches(self):
        return int(self.L.mct_launch_count(self.h))

    def frame_set(self, gray, c0=None, c1=None, c2=None):
        """Matrixu planes + initII() (Matrix.cpp:33-47).  numpy uint8 [H][W] host arrays."""
        gray = np.ascontiguousarray(gray, np.uint8)
        H, W = gray.shape
        planes = [None if p is None else np.ascontiguousarray(p, np.uint8) for p in (c0, c1, c2)]
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self._keep = (gray, planes)
        self.check(self.L.mct_frame_set(self.h, ptr(gray), ptr(planes[0]), ptr(planes[1]), ptr(planes[2]), W, H, W, 0))
        self.W, self.H = W, H

    def frame_set_bgr(self, bgr, use_hsv=False):
        """packed BGR uint8 [H][W][3] (the frame the reference's Camera reads): planes, gray and HSV are made on the device"""
        bgr = np.ascontiguousarray(bgr, np.uint8)
        H, W = bgr.shape[:2]
        self._keep = (bgr,)
        self.check(self.L.mct_frame_set_bgr(self.h, bgr.ctypes.data_as(C.c_void_p), W, H, 3 * W, int(use_hsv), 0))
        self.W, self.H = W, H

    def frame_prefetch_bgr(self, bgr, use_hsv=False):
        assert bgr.dtype == np.uint8 and bgr.flags["C_CONTIGUOUS"]
        H, W = bgr.shape[:2]
        self._keep_next = (bgr,)
        self.check(self.L.mct_frame_prefetch_bgr(self.h, bgr.ctypes.data_as(C.c_void_p), W, H, 3 * W, int(use_hsv)))

    def frame_planes(self):
        """device planes gray, c0, c1, c2 of the current frame as host arrays (tests)"""
        out = [np.empty((self.H, self.W), np.uint8) for _ in range(4)]
        self.check(self.L.mct_frame_get_planes(self.h, *[o.ctypes.data_as(C.c_void_p) for o in out]))
        return out

    def frame_prefetch(self, gray, c0=None, c1=None, c2=None):
        """Start the upload of the NEXT frame on the copy stream; frame_set with the same arrays adopts it."""
        H, W = gray.shape
        for a in (gray, c0, c1, c2):
            assert a is None or (a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"])
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self._keep_next = (gray, c0, c1, c2)
        self.check(self.L.mct_frame_prefetch(self.h, ptr(gray), ptr(c0), ptr(c1), ptr(c2), W, H, W))

    def frame_set_device(self, W, H, pitch, gray_ptr, c0_ptr=0, c1_ptr=0, c2_ptr=0):
        """Adopt device-resident planes (e.g. torch uint8 tensors' data_ptr()) without a copy."""
        vp = lambda p: C.c_void_p(p) if p else None
        self.check(self.L.mct_frame_set(self.h, vp(gray_ptr), vp(c0_ptr), vp(c1_ptr), vp(c2_ptr), W, H, pitch, 1))
        self.W, self.H = W, H

    def integral(self):
        out = np.empty((self.H + 1, self.W + 1), np.float32)
        self.check(self.L.mct_frame_get_integral(self.h, _fp(out)))
        return out

    def sum_rects(self, rects_xywh):
        r = np.ascontiguousarray(rects_xywh, np.int32).reshape(-1, 4)
        out = np.empty(len(r), np.float32)
        self.check(self.L.mct_sum_rects(self.h, _ip(r), len(r), _fp(out)))
        return out


class StrongClassifier:
    """StrongClassifierBase behind StrongClassifierFactory::CreateAndInitializeClassifier."""

    def __init__(self, ctx, feature_type, w0, h0, n_haar=0, n_bins=8, weak_type=WEAK_STUMP, strong_type=STRONG_MILBOOST,
                 n_selected=0, learning_rate=0.85, haar=None, cch_mode=0, max_train=4096, pct_retained=0.0):
        self.ctx = ctx
        L = ctx.L
        p = ClfParams(feature_type, n_haar, n_bins, w0, h0, weak_type, strong_type, n_selected, learning_rate, cch_mode,
                      max_train, 0, pct_retained)
        ht = None
        if haar is not None:
            nrect, rects, weights = haar
            nrect = np.ascontiguousarray(nrect, np.int32); rects = np.ascontiguousarray(rects, np.int32)
            weights = np.ascontiguousarray(weights, np.float32)
            self._keep = (nrect, rects, weights)
            ht = HaarTable(len(nrect), _ip(nrect), _ip(rects), _fp(weights))
        h = C.c_void_p()
        ctx.check(L.mct_classifier_create(ctx.h, C.byref(p), C.byref(ht) if ht is not None else None, C.byref(h)))
        self.h = h
        self.F = L.mct_classifier_num_features(h)

    def close(self):
        if self.h and not getattr(self, "_borrowed", False):
            self.ctx.L.mct_classifier_destroy(self.h)
        self.h = None

    @classmethod
    def from_handle(cls, ctx, h):
        """non-owning view of an existing mct_clf"""
        self = cls.__new__(cls)
        self.ctx = ctx; self.h = h
        self.F = ctx.L.mct_classifier_num_features(h)
        self._borrowed = True
        return self

    def GetNumberOfFeatures(self):
        return self.F

    def set_cluster(self, c):
        self.ctx.check(self.ctx.L.mct_classifier_set_cluster(self.h, c))

    def features(self, boxes):
        out = np.zeros((self.F, max(boxes.n, 1)), np.float32)
        self.ctx.check(self.ctx.L.mct_features_compute(self.h, C.byref(boxes), _fp(out)))
        return out[:, :boxes.n] if boxes.n else out[:, :0]

    def features_dev(self, boxes, dev_ptr, ld):
        """feature rows [F][n] written to device memory at dev_ptr (leading dimension ld floats)"""
        self.ctx.check(self.ctx.L.mct_features_compute_dev(self.h, C.byref(boxes), C.c_void_p(dev_ptr), int(ld)))

    def update_from_features_dev(self, pos_ptr, ld_pos, P, neg_ptr, ld_neg, Nn):
        self.ctx.check(self.ctx.L.mct_classifier_update_from_features_dev(self.h, C.c_void_p(pos_ptr), int(ld_pos), int(P),
                                                                          C.c_void_p(neg_ptr), int(ld_neg), int(Nn)))

    def Update(self, pos, neg, wpos=None, wneg=None):
        wp = None if wpos is None else _fp(np.ascontiguousarray(wpos, np.float32))
        wn = None if wneg is None else _fp(np.ascontiguousarray(wneg, np.float32))
        self.ctx.check(self.ctx.L.mct_classifier_update(self.h, C.byref(pos), C.byref(neg), wp, wn))

    def update_from_features(self, fpos, fneg, wpos=None, wneg=None):
        fpos = np.ascontiguousarray(fpos, np.float32); fneg = np.ascontiguousarray(fneg, np.float32)
        wp = None if wpos is None else _fp(np.ascontiguousarray(wpos, np.float32))
        wn = None if wneg is None else _fp(np.ascontiguousarray(wneg, np.float32))
        self.ctx.check(self.ctx.L.mct_classifier_update_from_features(self.h, _fp(fpos), fpos.shape[1], _fp(fneg),
                                                                      fneg.shape[1], wp, wn))

    def Classify(self, boxes, isLogRatioEnabled=True):
        out = np.zeros(max(boxes.n, 1), np.float32)
        self.ctx.check(self.ctx.L.mct_classifier_classify(self.h, C.byref(boxes), int(isLogRatioEnabled), _fp(out)))
        return out[:boxes.n]

    def classify_from_features(self, feat, log_ratio=True):
        feat = np.ascontiguousarray(feat, np.float32)
        out = np.zeros(max(feat.shape[1], 1), np.float32)
        self.ctx.check(self.ctx.L.mct_classifier_classify_from_features(self.h, _fp(feat), feat.shape[1], int(log_ratio),
                                                                        _fp(out)))
        return out[:feat.shape[1]]

    def alphas(self):
        """AdaBoost voting weights of the selected weak classifiers (empty for the MIL classifiers)"""
        a = np.zeros(4096, np.float32); n = C.c_int(0)
        self.ctx.check(self.ctx.L.mct_classifier_get_alphas(self.h, _fp(a), C.byref(n)))
        return a[:n.value].copy()

    def selectors(self):
        idx = np.zeros(self.F, np.int32)
        n = C.c_int(0)
        self.ctx.check(self.ctx.L.mct_classifier_get_selectors(self.h, _ip(idx), C.byref(n)))
        return idx[:n.value].copy()

    def weak_state(self):
        out = np.zeros((8, self.F), np.float32)
        self.ctx.check(self.ctx.L.mct_classifier_get_weak_state(self.h, _fp(out)))
        return out

    def predictions(self, max_m=4096):
        out = np.zeros((self.F, max_m), np.float32)
        P = C.c_int(0); Nn = C.c_int(0)
        # the call writes rows of length P+Nn; read sizes first through a dry buffer of the right size
        tmp = np.zeros(self.F * max_m, np.float32)
        self.ctx.check(self.ctx.L.mct_classifier_get_predictions(self.h, _fp(tmp), C.byref(P), C.byref(Nn)))
        M = P.value + Nn.value
        return tmp[:self.F * M].reshape(self.F, M).copy(), P.value, Nn.value


class ParticleFilter:
    """ParticleFilter (ParticleFilter.h:28-156) with device-resident state."""

    def __init__(self, ctx, n_particles, image_w, image_h):
        self.ctx = ctx
        self.N = n_particles
        h = C.c_void_p()
        ctx.check(ctx.L.mct_pf_create(ctx.h, n_particles, float(image_w), float(image_h), C.byref(h)))
        self.h = h

    @classmethod
    def from_handle(cls, ctx, h, n_particles):
        """non-owning view of an existing mct_pf"""
        self = cls.__new__(cls)
        self.ctx = ctx; self.N = n_particles; self.h = h
        self._borrowed = True
        return self

    def ground_pdf(self, H, mean, cov, h0, want_weights=False):
        """mct_pf_ground_pdf -> (total_weight, average[4], usable_covariance, weights or None)"""
        Hd = np.ascontiguousarray(H, np.float64).reshape(9)
        m = np.ascontiguousarray(mean, np.float32).reshape(2); c = np.ascontiguousarray(cov, np.float32).reshape(4)
        res = GroundPdfResult()
        w = np.zeros(self.N, np.float32) if want_weights else None
        self.ctx.check(self.ctx.L.mct_pf_ground_pdf(self.h, Hd.ctypes.data_as(C.POINTER(C.c_double)), _fp(m), _fp(c), float(h0),
                                                    C.byref(res), _fp(w) if want_weights else None))
        return float(res.total_weight), np.array(list(res.average), np.float32), int(res.usable_covariance), w

    def rearrange_ground(self, mean_foot, noise2xN, w0, h0):
        """mct_pf_rearrange_ground (RearrangeParticlesBasedOnGroundLocation + CheckForParticleRefinement)"""
        nz = np.ascontiguousarray(noise2xN, np.float32)
        assert nz.shape == (2, self.N)
        self.ctx.check(self.ctx.L.mct_pf_rearrange_ground(self.h, float(mean_foot[0]), float(mean_foot[1]), _fp(nz), float(w0), float(h0)))

    def close(self):
        if self.h and not getattr(self, "_borrowed", False):
            self.ctx.L.mct_pf_destroy(self.h)
            self.h = None

    def Initialize(self, cx, cy, sx=1.0, sy=1.0, max_scale_change=0.5, max_scale=10.0, min_scale=0.1):
        self.ctx.check(self.ctx.L.mct_pf_initialize(self.h, cx, cy, sx, sy, max_scale_change, max_scale, min_scale))

    def PredictWithBrownianMotion(self, noise4xN, w0, h0):
        nz = np.ascontiguousarray(noise4xN, np.float32)
        assert nz.size == 4 * self.N
        self.ctx.check(self.ctx.L.mct_pf_predict(self.h, nz.ctypes.data_as(C.c_void_p), 0, w0, h0))

    def UpdateAllParticlesWeight(self, prob, only_nonzero=True, force_reset=False, resample=True, u01=0.0):
        prob = np.ascontiguousarray(prob, np.float32)
        res = C.c_int(0)
        self.ctx.check(self.ctx.L.mct_pf_update_all_weights(self.h, _fp(prob), len(prob), int(only_nonzero),
                                                            int(force_reset), int(resample), u01, C.byref(res)))
        return bool(res.value)

    def ResampleParticles(self, force=False, u01=0.0):
        self.ctx.check(self.ctx.L.mct_pf_resample(self.h, int(force), u01))

    def state(self):
        out = np.zeros((self.N, 5), np.float32)
        self.ctx.check(self.ctx.L.mct_pf_get_state(self.h, _fp(out)))
        return out

    def set_state(self, m, filter_state=PF_PREDICTED):
        m = np.ascontiguousarray(m, np.float32)
        assert m.shape == (self.N, 5)
        self.ctx.check(self.ctx.L.mct_pf_set_state(self.h, _fp(m), filter_state))

    def resampled(self):
        out = np.zeros((self.N, 5), np.float32)
        self.ctx.check(self.ctx.L.mct_pf_get_resampled(self.h, _fp(out)))
        return out

    def resample_indices(self):
        out = np.zeros(self.N, np.int32)
        self.ctx.check(self.ctx.L.mct_pf_get_resample_indices(self.h, _ip(out)))
        return out

    def summary(self):
        s = PfSummary()
        self.ctx.check(self.ctx.L.mct_pf_get_summary(self.h, C.byref(s)))
        return s

    def last_response(self):
        H = np.zeros(self.N, np.float32); v = np.zeros(self.N, np.int32)
        self.ctx.check(self.ctx.L.mct_pf_get_last_response(self.h, _fp(H), _ip(v)))
        return H, v


def track_score(ctx, pfs, clfs, params, noise, u01, noise_device_ptr=None):
    """mct_track_score for the objects of one camera.  params: list of (w0, h0, exponent, force_reset, resample)."""
    n = len(pfs)
    vp = C.c_void_p
    pa = (vp * n)(*[p.h for p in pfs]); ca = (vp * n)(*[c.h for c in clfs])
    tp = (TrackParams * n)(*[TrackParams(*p) for p in params])
    u = np.ascontiguousarray(u01, np.float32)
    out = (PfSummary * n)()
    if noise_device_ptr is not None:
        rc = ctx.L.mct_track_score(ctx.h, n, pa, ca, tp, vp(noise_device_ptr), 1, _fp(u), out)
    elif noise is None:            # re-score variant: no prediction (UpdateParticleWeightsForRearrangedParticles)
        rc = ctx.L.mct_track_score(ctx.h, n, pa, ca, tp, None, 0, _fp(u), out)
    else:
        nz = np.ascontiguousarray(noise, np.float32)
        rc = ctx.L.mct_track_score(ctx.h, n, pa, ca, tp, nz.ctypes.data_as(vp), 0, _fp(u), out)
    ctx.check(rc)
    return list(out)


def abi_struct_layouts():
    """(name, ctypes mirror) of every public struct in the order mct_abi_struct_sizes reports them"""
    return [("mct_boxes", Boxes), ("mct_clf_params", ClfParams), ("mct_haar_table", HaarTable), ("mct_pf_summary", PfSummary),
            ("mct_track_params", TrackParams), ("mct_ground_pdf_result", GroundPdfResult), ("mct_train_gen", TrainGen),
            ("mct_train_gen_out", TrainGenOut), ("mct_sample_region", SampleRegion), ("mct_sample_region_out", SampleRegionOut),
            ("mct_train_default", TrainDefault), ("mct_train_default_out", TrainDefaultOut)]


def pf_rng_set(pf, state):
    pf.ctx.check(pf.ctx.L.mct_pf_rng_set(pf.h, C.c_uint64(int(state))))


def pf_rng_get(pf):
    st = C.c_uint64(0)
    pf.ctx.check(pf.ctx.L.mct_pf_rng_get(pf.h, C.byref(st)))
    return int(st.value)


def track_score_dev_rng(ctx, pfs, clfs, params, noise):
    """mct_track_score with u01 == NULL: the resampler draw comes from every object's device-resident stream (pf_rng_set)."""
    n = len(pfs)
    vp = C.c_void_p
    pa = (vp * n)(*[p.h for p in pfs]); ca = (vp * n)(*[c.h for c in clfs])
    tp = (TrackParams * n)(*[TrackParams(*p) for p in params])
    out = (PfSummary * n)()
    nz = np.ascontiguousarray(noise, np.float32)
    ctx.check(ctx.L.mct_track_score(ctx.h, n, pa, ca, tp, nz.ctypes.data_as(vp), 0, None, out))
    return list(out)


def track_update_default(ctx, pfs, clfs, gens, collect=True):
    """mct_track_update_default (+ _collect): gens = [TrainDefault(...)] per object; returns the read-outs when collected."""
    n = len(pfs)
    vp = C.c_void_p
    pa = (vp * n)(*[p.h for p in pfs]); ca = (vp * n)(*[c.h for c in clfs])
    ga = (TrainDefault * n)(*gens)
    ctx.check(ctx.L.mct_track_update_default(ctx.h, n, pa, ca, ga))
    if not collect:
        return None
    out = (TrainDefaultOut * n)()
    ctx.check(ctx.L.mct_track_update_default_collect(ctx.h, n, out))
    return list(out)


def track_update_default_boxes(ctx, obj, cap=8192):
    """boxes of object `obj` generated by the last track_update_default: (col, row, sx, sy, P, Nn)"""
    col = np.zeros(cap, np.int32); row = np.zeros(cap, np.int32); sx = np.zeros(cap, np.float32); sy = np.zeros(cap, np.float32)
    P = C.c_int(0); Nn = C.c_int(0)
    ctx.check(ctx.L.mct_track_update_default_boxes(ctx.h, obj, cap, _ip(col), _ip(row), _fp(sx), _fp(sy), C.byref(P), C.byref(Nn)))
    m = min(P.value + Nn.value, cap)
    return col[:m], row[:m], sx[:m], sy[:m], P.value, Nn.value


def classify_argmax(ctx, clfs, boxes_list, log_ratio=True):
    """mct_classify_argmax: (best index, best response) per object over its search window"""
    n = len(clfs)
    vp = C.c_void_p
    ca = (vp * n)(*[c.h for c in clfs])
    ba = (Boxes * n)(*boxes_list)
    idx = np.zeros(n, np.int32); val = np.zeros(n, np.float32)
    ctx.check(ctx.L.mct_classify_argmax(ctx.h, n, ca, ba, int(log_ratio), _ip(idx), _fp(val)))
    return idx, val


def pack_foot_points(ctx, pfs, h0, dev_out_ptr):
    """mct_fuser_pack_foot_points: [n_obj][2] float32 on the device at dev_out_ptr (e.g. a torch tensor's data_ptr())."""
    n = len(pfs)
    vp = C.c_void_p
    pa = (vp * n)(*[p.h for p in pfs])
    hh = np.ascontiguousarray(h0, np.float32)
    ctx.check(ctx.L.mct_fuser_pack_foot_points(ctx.h, n, pa, _fp(hh), vp(dev_out_ptr)))


def training_boxes(ctx, pfs, gens, cap):
    """mct_pf_training_boxes: gens = [(strategy, max_pos, cur[6], frame_w, frame_h, rng_state)]; returns per object
    (col, row, sx, sy, TrainGenOut)"""
    n = len(pfs)
    vp = C.c_void_p
    pa = (vp * n)(*[p.h for p in pfs])
    ga = (TrainGen * n)()
    for o, (st, mp, cur, fw, fh, rs) in enumerate(gens):
        ga[o].strategy = st; ga[o].max_pos = mp; ga[o].frame_w = fw; ga[o].frame_h = fh; ga[o].rng_state = rs
        for i in range(6):
            ga[o].cur[i] = float(cur[i])
    col = np.zeros((n, cap), np.int32); row = np.zeros((n, cap), np.int32)
    sx = np.zeros((n, cap), np.float32); sy = np.zeros((n, cap), np.float32)
    out = (TrainGenOut * n)()
    ctx.check(ctx.L.mct_pf_training_boxes(ctx.h, n, pa, ga, cap, _ip(col), _ip(row), _fp(sx), _fp(sy), out))
    res = []
    for o in range(n):
        k = out[o].n_pos
        res.append((col[o, :k].copy(), row[o, :k].copy(), sx[o, :k].copy(), sy[o, :k].copy(), out[o]))
    return res


def sample_regions(ctx, regions, cap):
    """mct_sample_regions: regions = [dict(kind, x, y, width, height, p, max_n, scale_x, scale_y, frame_w, frame_h, rng_state)];
    returns per region (col, row, SampleRegionOut)"""
    n = len(regions)
    ra = (SampleRegion * n)()
    for i, g in enumerate(regions):
        for k, v in g.items():
            if k == "p":
                for j in range(4):
                    ra[i].p[j] = float(v[j]) if j < len(v) else 0.0
            else:
                setattr(ra[i], k, v)
    col = np.zeros((n, cap), np.int32); row = np.zeros((n, cap), np.int32)
    out = (SampleRegionOut * n)()
    ctx.check(ctx.L.mct_sample_regions(ctx.h, n, ra, cap, _ip(col), _ip(row), out))
    return [(col[i, :out[i].n].copy(), row[i, :out[i].n].copy(), out[i]) for i in range(n)]


def track_update(ctx, clfs, pos_list, neg_list):
    n = len(clfs)
    vp = C.c_void_p
    ca = (vp * n)(*[c.h for c in clfs])
    pa = (Boxes * n)(*pos_list); na = (Boxes * n)(*neg_list)
    ctx.check(ctx.L.mct_track_update(ctx.h, n, ca, pa, na))
