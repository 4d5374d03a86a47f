"""ctypes binding of libmct_host.so: the C++ host mirror of the reference's Tracker / StrongClassifierFactory /
FeatureVector / SampleSet / ParticleFilter interfaces (mct_host.h), driven per camera."""
import ctypes as C

import numpy as np

from . import build_host

_lib = None


class ObjectParams(C.Structure):
    _fields_ = [("feature_type", C.c_int), ("n_haar", C.c_int), ("n_bins", C.c_int), ("weak_type", C.c_int),
                ("strong_type", C.c_int), ("percent_selected", C.c_int), ("n_particles", C.c_int),
                ("sd_x", C.c_float), ("sd_y", C.c_float), ("sd_sx", C.c_float), ("sd_sy", C.c_float),
                ("pos_radius_train", C.c_float), ("init_pos_radius_train", C.c_float),
                ("n_neg_train", C.c_int), ("init_n_neg_train", C.c_int), ("search_win", C.c_int),
                ("trajectory_option", C.c_int), ("rng_seed", C.c_int64), ("init_xywh", C.c_float * 4),
                ("percent_retained", C.c_float), ("pos_strategy", C.c_int), ("neg_strategy", C.c_int),
                ("max_num_pos_examples", C.c_int)]


def load():
    global _lib
    if _lib is None:
        L = C.CDLL(build_host.build())
        vp = C.c_void_p
        L.mcth_last_error.argtypes = [vp]; L.mcth_last_error.restype = C.c_char_p
        L.mcth_camera_create.argtypes = [C.c_int, vp, C.c_int, C.POINTER(vp)]
        L.mcth_camera_destroy.argtypes = [vp]
        L.mcth_camera_ctx.argtypes = [vp]; L.mcth_camera_ctx.restype = vp
        L.mcth_camera_set_frame.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.mcth_camera_prefetch_frame.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int]
        L.mcth_camera_set_frame_bgr.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.mcth_camera_prefetch_frame_bgr.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.mcth_camera_prefetch_noise.argtypes = [vp, vp, C.c_size_t]
        L.mcth_camera_add_object.argtypes = [vp, C.POINTER(ObjectParams)]
        L.mcth_camera_init_trackers.argtypes = [vp]
        L.mcth_camera_track.argtypes = [vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
        L.mcth_camera_step.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_size_t, C.c_int, C.POINTER(C.c_int)]
        L.mcth_camera_track_resident.argtypes = [vp, C.c_int, vp]
        L.mcth_camera_pack_foot_points.argtypes = [vp, vp]
        dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
        L.mcth_camera_ground_feedback.argtypes = [vp, C.c_int, dp, fp, fp, C.POINTER(C.c_int), fp]
        L.mcth_camera_ground_feedback_all.argtypes = [vp, dp, fp, fp, C.POINTER(C.c_int), fp]
        L.mcth_fuser_create.argtypes = [C.c_int, dp, C.POINTER(vp)]
        L.mcth_fuser_destroy.argtypes = [vp]
        L.mcth_fuser_last_error.argtypes = [vp]; L.mcth_fuser_last_error.restype = C.c_char_p
        L.mcth_fuser_initialize.argtypes = [vp, fp]
        L.mcth_fuser_fuse.argtypes = [vp, fp]
        L.mcth_fuser_get.argtypes = [vp, fp, fp, fp, fp]
        L.mcth_fuser_fuse_many.argtypes = [C.POINTER(vp), C.c_int, fp, C.c_int, fp, fp, C.POINTER(C.c_int)]
        L.mcth_fuser_intersection.argtypes = [C.c_int, dp, fp, dp]
        L.mcth_randgaus.argtypes = [C.c_int64, C.c_float, C.c_int, fp]
        L.mcth_camera_add_simple_object.argtypes = [vp, C.POINTER(ObjectParams)]
        L.mcth_camera_init_simple.argtypes = [vp]
        L.mcth_camera_track_simple.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double)]
        L.mcth_simple_selectors.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
        L.mcth_simple_rng_state.argtypes = [vp, C.c_int]; L.mcth_simple_rng_state.restype = C.c_uint64
        L.mcth_simple_last_test_size.argtypes = [vp, C.c_int]
        L.mcth_simple_classifier.argtypes = [vp, C.c_int]; L.mcth_simple_classifier.restype = vp
        L.mcth_camera_exchange_foot_points.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
        L.mcth_object_handles.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]
        L.mcth_object_randfloat.argtypes = [vp, C.c_int]; L.mcth_object_randfloat.restype = C.c_float
        L.mcth_camera_sync.argtypes = [vp]
        L.mcth_camera_launches.argtypes = [vp]; L.mcth_camera_launches.restype = C.c_longlong
        L.mcth_object_selectors.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
        L.mcth_object_state.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
        L.mcth_object_particles.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
        L.mcth_object_last_train.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        L.mcth_object_rng_state.argtypes = [vp, C.c_int]; L.mcth_object_rng_state.restype = C.c_uint64
        L.mcth_object_box.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.mcth_set_device_training_sets.argtypes = [C.c_int]; L.mcth_set_device_training_sets.restype = None
        _lib = L
    return _lib


def set_device_training_sets(on):
    """Whole frame on the device (k_train_gen_default behind the scoring kernels, the trackers' streams device-resident) on / off
    for this process; off = the host enumerates the default-strategy training sets (ParticleFilterTracker::BuildTrainingSets)."""
    load().mcth_set_device_training_sets(1 if on else 0)


def device_training_sets():
    return bool(load().mcth_get_device_training_sets())


class HostError(RuntimeError):
    pass


# config.cfg values
DEFAULTS = dict(feature_type=0, n_haar=250, n_bins=8, weak_type=0, strong_type=2, percent_selected=20, n_particles=1000,
                sd_x=20.0, sd_y=20.0, sd_sx=1e-7, sd_sy=1e-7, pos_radius_train=10.0, init_pos_radius_train=7.0,
                n_neg_train=30, init_n_neg_train=30, search_win=20, trajectory_option=0, rng_seed=1, percent_retained=0.0,
                pos_strategy=0, neg_strategy=0, max_num_pos_examples=30)


class Camera:
    """One camera on one GPU: frame + the ParticleFilterTrackers of its objects."""

    def __init__(self, device=0, n_frames=100, stream=None):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.mcth_camera_create(device, C.c_void_p(stream) if stream else None, n_frames, C.byref(h))
        if rc:
            raise HostError("mcth_camera_create failed (%d): no CUDA device? (there is no CPU fallback)" % rc)
        self.h = h
        self.n_obj = 0
        self.N = []
        self._keep = None

    def _chk(self, rc):
        if rc:
            raise HostError("host error %d: %s" % (rc, self.L.mcth_last_error(self.h).decode()))

    def close(self):
        if self.h:
            self.L.mcth_camera_destroy(self.h)
            self.h = None

    def add_object(self, init_xywh, **kw):
        p = dict(DEFAULTS); p.update(kw)
        op = ObjectParams()
        for k, v in p.items():
            setattr(op, k, v)
        op.init_xywh = (C.c_float * 4)(*[float(x) for x in init_xywh])
        self._chk(self.L.mcth_camera_add_object(self.h, C.byref(op)))
        self.n_obj += 1
        self.N.append(p["n_particles"])

    # ---- SimpleTracker objects (Local_Tracker_Type 0, the reference's default)
    def add_simple_object(self, init_xywh, **kw):
        p = dict(DEFAULTS); p.update(kw)
        op = ObjectParams()
        for k, v in p.items():
            setattr(op, k, v)
        op.init_xywh = (C.c_float * 4)(*[float(x) for x in init_xywh])
        self._chk(self.L.mcth_camera_add_simple_object(self.h, C.byref(op)))
        self.n_simple = getattr(self, "n_simple", 0) + 1

    def init_simple(self):
        self._chk(self.L.mcth_camera_init_simple(self.h))

    def track_simple(self, frameind):
        """one frame for all SimpleTracker objects: returns (states [n][4] float32 x, y, w, h; best responses [n])"""
        n = getattr(self, "n_simple", 0)
        st = np.zeros((n, 4), np.float32); rs = np.zeros(n, np.float64)
        self._chk(self.L.mcth_camera_track_simple(self.h, frameind, st.ctypes.data_as(C.POINTER(C.c_float)),
                                                  rs.ctypes.data_as(C.POINTER(C.c_double))))
        return st, rs

    def simple_selectors(self, o):
        idx = np.zeros(4096, np.int32)
        n = self.L.mcth_simple_selectors(self.h, o, idx.ctypes.data_as(C.POINTER(C.c_int)))
        if n < 0:
            self._chk(n)
        return idx[:n].copy()

    def simple_classifier(self, o):
        """non-owning capi.StrongClassifier of SimpleTracker object o (introspection)"""
        from .. import capi
        return capi.StrongClassifier.from_handle(self.capi_context(), C.c_void_p(self.L.mcth_simple_classifier(self.h, o)))

    def simple_rng_state(self, o):
        return int(self.L.mcth_simple_rng_state(self.h, o))

    def simple_last_test_size(self, o):
        return int(self.L.mcth_simple_last_test_size(self.h, o))

    def set_frame(self, gray, c0=None, c1=None, c2=None):
        gray = np.ascontiguousarray(gray, np.uint8)
        H, W = gray.shape
        pl = [None if p is None else np.ascontiguousarray(p, np.uint8) for p in (c0, c1, c2)]
        self._keep = (gray, pl)
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self._chk(self.L.mcth_camera_set_frame(self.h, ptr(gray), ptr(pl[0]), ptr(pl[1]), ptr(pl[2]), W, H, W, 0))

    def set_frame_device(self, W, H, pitch, gray_ptr, c0_ptr=0, c1_ptr=0, c2_ptr=0, ready=False):
        """device-resident planes; ready=True: the planes are already complete (no ordering against the ctx stream
        needed), which lets the frame preparation overlap the previous frame's kernels"""
        vp = lambda p: C.c_void_p(p) if p else None
        self._chk(self.L.mcth_camera_set_frame(self.h, vp(gray_ptr), vp(c0_ptr), vp(c1_ptr), vp(c2_ptr), W, H, pitch, 2 if ready else 1))

    def init_trackers(self):
        self._chk(self.L.mcth_camera_init_trackers(self.h))

    def track(self, frameind, noise, phases=3):
        """noise: [n_obj][4][N] float32.  phases: bit 0 score + store the state, bit 1 UpdateClassifier, bit 2 (with bit 0)
        score WITHOUT storing the state (fusion branch), bit 3 store the state from the particles as they are now (after the
        fusers).  Returns [n_obj][4] boxes (x, y, w, h) of the stored state."""
        nz = np.ascontiguousarray(noise, np.float32)
        out = np.zeros((self.n_obj, 4), np.int32)
        self._chk(self.L.mcth_camera_track(self.h, frameind, nz.ctypes.data_as(C.c_void_p), phases,
                                           out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def step(self, frameind, planes, next_planes, noise, next_noise=None, phases=3):
        """mcth_camera_step: make `planes` (gray, c0, c1, c2 uint8 arrays) the current frame, announce the next frame and its noise
        (uploaded while this frame is tracked), track.  The arrays must stay alive until the next call."""
        pl = [np.ascontiguousarray(p, np.uint8) for p in planes]
        H, W = pl[0].shape
        pa = (C.c_void_p * 4)(*[p.ctypes.data for p in pl])
        na = None
        if next_planes is not None:
            npl = [np.ascontiguousarray(p, np.uint8) for p in next_planes]
            na = (C.c_void_p * 4)(*[p.ctypes.data for p in npl])
            self._keep_next = npl
        nz = np.ascontiguousarray(noise, np.float32)
        nn = np.ascontiguousarray(next_noise, np.float32) if next_noise is not None else None
        self._keep_step = (pl, nz, nn)
        out = np.zeros((self.n_obj, 4), np.int32)
        self._chk(self.L.mcth_camera_step(self.h, frameind, pa, na, W, H, W, nz.ctypes.data_as(C.c_void_p),
                                          nn.ctypes.data_as(C.c_void_p) if nn is not None else None, nn.size if nn is not None else 0,
                                          phases, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def pack_foot_points(self, dev_out_ptr):
        """geometric-fuser payload [n_obj][2] float32 written on the device (torch tensor data_ptr())"""
        self._chk(self.L.mcth_camera_pack_foot_points(self.h, C.c_void_p(dev_out_ptr)))

    def ground_feedback(self, o, H, mean, cov):
        """UpdateParticlesWithGroundPDF for object o (ParticleFilterTracker.cpp:1040-1125): the fused ground-plane estimate
        (Kalman mean / covariance) is projected into this camera; when the particles' mean foot point is 20 px or more
        away they are rearranged around it and re-scored.  Returns (rearranged, distance)."""
        Hd = np.ascontiguousarray(H, np.float64).reshape(9)
        m = np.ascontiguousarray(mean, np.float32).reshape(2); cv = np.ascontiguousarray(cov, np.float32).reshape(4)
        r = C.c_int(0); d = C.c_float(0)
        fp = C.POINTER(C.c_float)
        self._chk(self.L.mcth_camera_ground_feedback(self.h, o, Hd.ctypes.data_as(C.POINTER(C.c_double)), m.ctypes.data_as(fp),
                                                     cv.ctypes.data_as(fp), C.byref(r), C.byref(d)))
        return bool(r.value), float(d.value)

    def capi_context(self):
        """non-owning capi.Context on this camera's mct_ctx (same stream, same frame)"""
        from .. import capi
        return capi.Context.from_handle(C.c_void_p(self.L.mcth_camera_ctx(self.h)))

    def particle_filter(self, o):
        """non-owning capi.ParticleFilter of object o"""
        from .. import capi
        pf = C.c_void_p()
        self._chk(self.L.mcth_object_handles(self.h, o, C.byref(pf), None))
        return capi.ParticleFilter.from_handle(self.capi_context(), pf, self.N[o])

    def randfloat(self, o):
        return float(self.L.mcth_object_randfloat(self.h, o))

    def p2p_setup(self, rank, world, allgather_bytes):
        """peer-memory exchange set-up: allgather_bytes(b: bytes) -> list of every rank's bytes (e.g. through
        torch.distributed.all_gather_object); the caller puts a barrier behind it"""
        from .. import capi
        ctx = self.capi_context()
        buf = C.create_string_buffer(64)
        ctx.check(ctx.L.mct_p2p_create(ctx.h, world, rank, buf))
        allh = allgather_bytes(bytes(buf.raw))
        assert len(allh) == world and all(len(x) == 64 for x in allh)
        ctx.check(ctx.L.mct_p2p_open(ctx.h, b"".join(allh)))
        self._p2p_world = world

    def exchange_foot_points(self, timeout_ms=2000):
        """foot points of all cameras [world][n_obj][2] through the peer-memory kernels (every camera calls this once per frame)"""
        out = np.zeros((self._p2p_world, self.n_obj, 2), np.float32)
        self._chk(self.L.mcth_camera_exchange_foot_points(self.h, out.ctypes.data_as(C.POINTER(C.c_float)), int(timeout_ms)))
        return out

    def ground_feedback_all(self, H, means, covs):
        """ground_feedback for every object of the camera with ONE device pass for the pdfs: means [n][2], covs [n][2][2].
        Returns [(rearranged, distance)] per object."""
        n = self.n_obj
        Hd = np.ascontiguousarray(H, np.float64).reshape(9)
        m = np.ascontiguousarray(means, np.float32).reshape(n, 2); cv = np.ascontiguousarray(covs, np.float32).reshape(n, 4)
        r = np.zeros(n, np.int32); d = np.zeros(n, np.float32)
        fp = C.POINTER(C.c_float)
        self._chk(self.L.mcth_camera_ground_feedback_all(self.h, Hd.ctypes.data_as(C.POINTER(C.c_double)), m.ctypes.data_as(fp),
                                                         cv.ctypes.data_as(fp), r.ctypes.data_as(C.POINTER(C.c_int)), d.ctypes.data_as(fp)))
        return [(bool(r[o]), float(d[o])) for o in range(n)]

    def sync(self):
        self._chk(self.L.mcth_camera_sync(self.h))

    @property
    def launches(self):
        return int(self.L.mcth_camera_launches(self.h))

    def selectors(self, o):
        idx = np.zeros(4096, np.int32)
        n = self.L.mcth_object_selectors(self.h, o, idx.ctypes.data_as(C.POINTER(C.c_int)))
        if n < 0:
            self._chk(n)
        return idx[:n].copy()

    def state(self, o):
        s = np.zeros(6, np.float32)
        self.L.mcth_object_state(self.h, o, s.ctypes.data_as(C.POINTER(C.c_float)))
        return s

    def particles(self, o):
        m = np.zeros((self.N[o], 5), np.float32)
        self._chk(self.L.mcth_object_particles(self.h, o, m.ctypes.data_as(C.POINTER(C.c_float))))
        return m

    def last_train(self, o, which):
        col = np.zeros(16384, np.int32); row = np.zeros(16384, np.int32)
        n = self.L.mcth_object_last_train(self.h, o, which, col.ctypes.data_as(C.POINTER(C.c_int)),
                                          row.ctypes.data_as(C.POINTER(C.c_int)), 16384)
        return col[:n].copy(), row[:n].copy()

    def boxes(self, o, frameind):
        """m_states row (x, y, w, h) stored for a frame"""
        out = (C.c_int * 4)()
        self._chk(self.L.mcth_object_box(self.h, o, frameind, out))
        return list(out)

    def rng_state(self, o):
        return int(self.L.mcth_object_rng_state(self.h, o))


class GeometricFuser:
    """GeometryBasedInformationFuser of ONE object (principal-axis mode): every rank holds the same instance and feeds it
    the same all-gathered foot points, so no broadcast of the Kalman state is needed."""

    def __init__(self, homographies):
        self.L = load()
        Hs = np.ascontiguousarray(homographies, np.float64).reshape(-1, 9)
        self.n_cameras = Hs.shape[0]
        h = C.c_void_p()
        rc = self.L.mcth_fuser_create(self.n_cameras, Hs.ctypes.data_as(C.POINTER(C.c_double)), C.byref(h))
        if rc:
            raise HostError("mcth_fuser_create failed (%d): at least two cameras are needed" % rc)
        self.h = h
        self.initialized = False

    def _chk(self, rc):
        if rc:
            raise HostError("fuser error %d: %s" % (rc, self.L.mcth_fuser_last_error(self.h).decode()))

    def _foot(self, foot):
        f = np.ascontiguousarray(foot, np.float32)
        if f.shape != (self.n_cameras, 2):
            raise ValueError("foot points must be float32 [%d][2]" % self.n_cameras)
        return f

    def initialize(self, foot):
        f = self._foot(foot)
        self._chk(self.L.mcth_fuser_initialize(self.h, f.ctypes.data_as(C.POINTER(C.c_float))))
        self.initialized = True

    def fuse(self, foot):
        """FuseInformation: InitializeKalman on the first call, predict + correct afterwards"""
        if not self.initialized:
            return self.initialize(foot)
        f = self._foot(foot)
        self._chk(self.L.mcth_fuser_fuse(self.h, f.ctypes.data_as(C.POINTER(C.c_float))))

    def posterior(self):
        """-> (mean[2], cov[2][2], state[4], P[4][4]) float32"""
        fp = C.POINTER(C.c_float)
        m = np.zeros(2, np.float32); cv = np.zeros(4, np.float32); st = np.zeros(4, np.float32); P = np.zeros(16, np.float32)
        self._chk(self.L.mcth_fuser_get(self.h, m.ctypes.data_as(fp), cv.ctypes.data_as(fp), st.ctypes.data_as(fp), P.ctypes.data_as(fp)))
        return m, cv.reshape(2, 2), st, P.reshape(4, 4)

    def close(self):
        if self.h:
            self.L.mcth_fuser_destroy(self.h)
            self.h = None


def fuse_many(fusers, allp):
    """one frame for the fusers of all objects in one call: allp float32 [n_cameras][n_obj][2].
    Returns (means [n_obj][2], covs [n_obj][2][2], fresh [n_obj] bool: initialised by this frame)"""
    L = load()
    n = len(fusers)
    a = np.ascontiguousarray(allp, np.float32)
    assert a.shape[1:] == (n, 2)
    hs = (C.c_void_p * n)(*[f.h for f in fusers])
    means = np.zeros((n, 2), np.float32); covs = np.zeros((n, 4), np.float32); fresh = np.zeros(n, np.int32)
    fp = C.POINTER(C.c_float)
    rc = L.mcth_fuser_fuse_many(hs, n, a.ctypes.data_as(fp), a.shape[0], means.ctypes.data_as(fp), covs.ctypes.data_as(fp),
                                fresh.ctypes.data_as(C.POINTER(C.c_int)))
    if rc:
        raise HostError("fuser error %d" % rc)
    for f, fr in zip(fusers, fresh):
        f.initialized = True
    return means, covs.reshape(n, 2, 2), fresh.astype(bool)


def principal_axis_intersection(homographies, foot):
    L = load()
    Hs = np.ascontiguousarray(homographies, np.float64).reshape(-1, 9)
    f = np.ascontiguousarray(foot, np.float32)
    out = np.zeros(2, np.float64)
    if L.mcth_fuser_intersection(Hs.shape[0], Hs.ctypes.data_as(C.POINTER(C.c_double)), f.ctypes.data_as(C.POINTER(C.c_float)),
                                 out.ctypes.data_as(C.POINTER(C.c_double))):
        raise HostError("principal-axis intersection failed (needs >= 2 cameras)")
    return out


def randgaus_stream(seed, sigma, n):
    out = np.zeros(n, np.float32)
    load().mcth_randgaus(seed, float(sigma), n, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


class NetworkParams(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("geometric_fusion", "appearance_fusion", "af_strong", "af_weak", "af_pct_selected",
                                       "af_pct_retained", "af_refresh")]


_AG_HOST = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
_AG_ROWS = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_void_p)


class Exchange(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("user", C.c_void_p), ("allgather_host", _AG_HOST), ("alltoall_rows", _AG_ROWS)]


class Network:
    """CameraNetwork (mct_network.h): the reference's per-frame schedule over the cameras of a network, fusers included.

      cameras        this process' Camera objects (objects added, trackers NOT yet initialised, frame 0 set)
      camera_ids     their indices in the whole network (None: 0 .. len - 1)
      n_cameras      cameras in the whole network (None: len(cameras), i.e. everything lives in this process)
      homographies   [n_cameras][3][3] image -> ground plane (geometric fuser)
      exchange       (rank, world, allgather_host(send_ptr, recv_ptr, nbytes), alltoall_rows(slot_bytes, stream_ptr)) when the
                     network spans processes (one camera per process); see network.TorchExchange
    """

    def __init__(self, cameras, homographies=None, camera_ids=None, n_cameras=None, exchange=None, geometric_fusion=0,
                 appearance_fusion=0, af_strong=1, af_weak=0, af_pct_selected=20, af_pct_retained=0, af_refresh=1):
        self.L = L = load()
        vp = C.c_void_p
        L.mcth_network_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_double),
                                          C.POINTER(NetworkParams), C.POINTER(Exchange), C.POINTER(vp)]
        L.mcth_network_destroy.argtypes = [vp]
        L.mcth_network_last_error.argtypes = [vp]; L.mcth_network_last_error.restype = C.c_char_p
        L.mcth_network_init.argtypes = [vp]
        L.mcth_network_track.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.mcth_network_row_bytes.argtypes = [vp]; L.mcth_network_row_bytes.restype = C.c_size_t
        L.mcth_network_set_row_buffers.argtypes = [vp, vp, vp]
        L.mcth_network_kalman.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.mcth_network_global_selectors.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.mcth_network_global_classifier.argtypes = [vp, C.c_int, C.c_int]; L.mcth_network_global_classifier.restype = vp
        L.mcth_network_last_rearranged.argtypes = [vp, C.c_int, C.c_int]
        L.mcth_network_stage_ms.argtypes = [vp, C.POINTER(C.c_double)]
        self.cameras = list(cameras)
        n_local = len(self.cameras)
        ids = list(range(n_local)) if camera_ids is None else list(camera_ids)
        nc = n_local if n_cameras is None else int(n_cameras)
        self.n_cameras = nc
        Hp = None
        if homographies is not None:
            self._H = np.ascontiguousarray(homographies, np.float64).reshape(nc, 9)
            Hp = self._H.ctypes.data_as(C.POINTER(C.c_double))
        p = NetworkParams(geometric_fusion, appearance_fusion, af_strong, af_weak, af_pct_selected, af_pct_retained, af_refresh)
        self._ex = None
        exp = None
        if exchange is not None:
            rank, world, ag_host, ag_rows = exchange
            self._cb = (_AG_HOST(lambda user, s, r, n: ag_host(s, r, n)), _AG_ROWS(lambda user, n, st: ag_rows(n, st)))
            self._ex = Exchange(rank, world, None, self._cb[0], self._cb[1])
            exp = C.byref(self._ex)
        h = vp()
        cams = (vp * n_local)(*[c.h for c in self.cameras])
        rc = L.mcth_network_create(cams, (C.c_int * n_local)(*ids), n_local, nc, Hp, C.byref(p), exp, C.byref(h))
        self.h = h
        if rc:
            raise HostError("mcth_network_create failed (%d): %s" % (rc, L.mcth_network_last_error(h).decode()))

    def _chk(self, rc):
        if rc:
            raise HostError("network error %d: %s" % (rc, self.L.mcth_network_last_error(self.h).decode()))

    def init(self):
        """trackers of every local camera (frame 0 already set), global classifiers, geometric fusers"""
        self._chk(self.L.mcth_network_init(self.h))

    def row_bytes(self):
        return int(self.L.mcth_network_row_bytes(self.h))

    def set_row_buffers(self, send_ptr, recv_ptr):
        self._chk(self.L.mcth_network_set_row_buffers(self.h, C.c_void_p(send_ptr), C.c_void_p(recv_ptr)))

    def track(self, frameind, noise):
        """noise: per local camera a float32 array [n_obj][4][N].  The cameras' frames are already set."""
        keep = [np.ascontiguousarray(z, np.float32) for z in noise]
        ptrs = (C.c_void_p * len(keep))(*[z.ctypes.data for z in keep])
        self._chk(self.L.mcth_network_track(self.h, frameind, ptrs))

    def kalman(self, o):
        st = np.zeros(4, np.float32); P = np.zeros(16, np.float32)
        if self.L.mcth_network_kalman(self.h, o, st.ctypes.data_as(C.POINTER(C.c_float)), P.ctypes.data_as(C.POINTER(C.c_float))):
            raise HostError("no geometric fuser")
        return st, P.reshape(4, 4)

    def global_selectors(self, local_cam, o):
        idx = np.zeros(4096, np.int32)
        n = self.L.mcth_network_global_selectors(self.h, local_cam, o, idx.ctypes.data_as(C.POINTER(C.c_int)))
        if n < 0:
            raise HostError(self.L.mcth_network_last_error(self.h).decode())
        return idx[:n].copy()

    def last_rearranged(self, local_cam, o):
        return int(self.L.mcth_network_last_rearranged(self.h, local_cam, o))

    def stage_ms(self):
        out = (C.c_double * 8)()
        self.L.mcth_network_stage_ms(self.h, out)
        return dict(zip(("score", "reweight", "geo_exchange", "geo_fusers", "geo_feedback", "update", "learn", "store"), list(out)))

    def close(self):
        if self.h:
            self.L.mcth_network_destroy(self.h); self.h = None
