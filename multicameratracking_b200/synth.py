"""Synthetic camera sequences: smooth low-frequency background + uniform noise, plus one
textured rectangle per object with a distinct mean colour moving along x(t) = x0 + v t + A sin(w t).
Planes are R, G, B (Matrix.cpp:92-94) and gray = (R*4899 + G*9617 + B*1868 + 8192) >> 14 (cv BGR2GRAY).
"""
import numpy as np

OBJ_WH = {(640, 480): (40, 80), (1280, 720): (60, 120), (1920, 1080): (90, 180)}


def gray_from_rgb(r, g, b):
    return ((r.astype(np.int32) * 4899 + g.astype(np.int32) * 9617 + b.astype(np.int32) * 1868 + 8192) >> 14).astype(np.uint8)


class Sequence:
    def __init__(self, W=640, H=480, n_objects=1, camera=0, n_frames=100, obj_wh=None, camouflage=None):
        """camouflage: None = objects with their own mean colour (clearly separable from the background); a number d = the
        object is the local background shifted by a per-object colour offset of magnitude <= d (plus its texture): colour
        histograms of object and surroundings overlap, as in real footage.  MILAnyBoost needs that: on perfectly separable
        samples its first weak classifier saturates the bag likelihood, no second one can improve it, and the reference
        aborts with "Could not find best 2 weak classifiers" (MILAnyBoostClassifier.cpp:102-131)."""
        self.W, self.H, self.n_objects, self.n_frames = W, H, n_objects, n_frames
        self.camouflage = camouflage
        self.w0, self.h0 = obj_wh if obj_wh else OBJ_WH.get((W, H), (40, 80))
        rng = np.random.default_rng(1000 + camera)
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        bg = []
        for _ in range(3):
            a = rng.uniform(60, 160); fx, fy = rng.uniform(0.004, 0.02, 2); ph = rng.uniform(0, 6.28, 2)
            bg.append(a + 40 * np.sin(fx * xx + ph[0]) * np.cos(fy * yy + ph[1]))
        self.bg = np.stack(bg)
        self.colors = rng.uniform(30, 225, (n_objects, 3)).astype(np.float32)
        if camouflage is not None:
            self.offsets = (np.random.default_rng(77 + camera).uniform(-1, 1, (n_objects, 3)) * float(camouflage)).astype(np.float32)
        self.tex = rng.uniform(-25, 25, (n_objects, 3, self.h0, self.w0)).astype(np.float32)
        m = 2 * 20 + 8
        self.x0 = rng.uniform(m, W - self.w0 - m, n_objects)
        self.y0 = rng.uniform(m, H - self.h0 - m, n_objects)
        self.v = rng.uniform(-1.5, 1.5, (n_objects, 2))
        self.A = rng.uniform(2, 10, (n_objects, 2))
        self.om = rng.uniform(0.05, 0.2, (n_objects, 2))
        self.camera = camera

    def box(self, t, o):
        m = 2 * 20 + 8
        x = self.x0[o] + self.v[o, 0] * t + self.A[o, 0] * np.sin(self.om[o, 0] * t)
        y = self.y0[o] + self.v[o, 1] * t + self.A[o, 1] * np.sin(self.om[o, 1] * t)
        x = float(np.clip(x, m, self.W - self.w0 - m)); y = float(np.clip(y, m, self.H - self.h0 - m))
        return int(round(x)), int(round(y)), self.w0, self.h0

    def frame(self, t):
        """returns gray, r, g, b (uint8 [H][W])"""
        rng = np.random.default_rng(100000 * (self.camera + 1) + t)
        img = self.bg + rng.uniform(-8, 8, self.bg.shape).astype(np.float32)
        for o in range(self.n_objects):
            x, y, w, h = self.box(t, o)
            if self.camouflage is None:
                img[:, y:y + h, x:x + w] = self.colors[o][:, None, None] + self.tex[o]
            else:
                img[:, y:y + h, x:x + w] += self.offsets[o][:, None, None] + self.tex[o]
        img = np.clip(img, 0, 255).astype(np.uint8)
        r, g, b = img[0], img[1], img[2]
        return gray_from_rgb(r, g, b), np.ascontiguousarray(r), np.ascontiguousarray(g), np.ascontiguousarray(b)


def noise(n_particles, sd, camera=0, obj=0, t=0):
    """The four cv::randn(1xN, 0, sigma) rows of PredictWithBrownianMotion (x, y, scaleX, scaleY)."""
    rng = np.random.default_rng(2000 + camera * 64 + obj + 7919 * t)
    out = np.empty((4, n_particles), np.float32)
    for k in range(4):
        out[k] = (rng.standard_normal(n_particles) * sd[k]).astype(np.float32)
    return out


class MultiViewScene:
    """C cameras looking at the same O objects on one ground plane (BASELINE configs[2] / [3]).  The objects move on the
    ground (world coordinates); camera c sees the FOOT of object o at  centre + s * R(theta_c) (P - world_centre)  and the
    object's box stands on that foot point (foot = bottom centre).  The image -> ground homography of camera c is the
    inverse of that affine map, so the cameras' principal axes (vertical image lines through the foot points, mapped to the
    ground) meet at the object's ground position: a geometrically consistent input for GeometryBasedInformationFuser."""

    def __init__(self, n_cameras, n_objects, W=640, H=480, n_frames=100, obj_wh=None, seed=7, camouflage=None):
        self.C, self.O, self.W, self.H = n_cameras, n_objects, W, H
        self.seqs = [Sequence(W, H, n_objects, camera=c, n_frames=n_frames, obj_wh=obj_wh, camouflage=camouflage) for c in range(n_cameras)]
        w0, h0 = self.seqs[0].w0, self.seqs[0].h0
        rng = np.random.default_rng(seed)
        half = 0.28 * min(W, H)                           # world square [-half, half]^2 around the origin
        self.P0 = rng.uniform(-0.8 * half, 0.8 * half, (n_objects, 2))
        self.v = rng.uniform(-1.0, 1.0, (n_objects, 2))
        self.A = rng.uniform(2, 8, (n_objects, 2))
        self.om = rng.uniform(0.05, 0.2, (n_objects, 2))
        self.half = half
        self.theta = [0.35 + c * np.pi / max(n_cameras, 2) for c in range(n_cameras)]
        self.scale = 1.0
        self.centre = np.array([W / 2.0, (H + h0) / 2.0])
        self.homographies = []
        for c in range(n_cameras):
            M = self._affine(c)
            self.homographies.append(np.linalg.inv(M))
            self.seqs[c].box = (lambda t, o, c=c: self.box(c, t, o))

    def _affine(self, c):
        th, s = self.theta[c], self.scale
        R = s * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        M = np.eye(3)
        M[:2, :2] = R
        M[:2, 2] = self.centre
        return M                                           # ground (x, y, 1) -> image foot point

    def ground(self, t, o):
        p = self.P0[o] + self.v[o] * t + self.A[o] * np.sin(self.om[o] * t)
        return np.clip(p, -self.half, self.half)

    def box(self, c, t, o):
        w0, h0 = self.seqs[c].w0, self.seqs[c].h0
        g = self.ground(t, o)
        f = self._affine(c) @ np.array([g[0], g[1], 1.0])
        x = int(round(float(np.clip(f[0] - w0 / 2.0, 4, self.W - w0 - 5))))
        y = int(round(float(np.clip(f[1] - h0, 4, self.H - h0 - 5))))
        return x, y, w0, h0

    def frame(self, c, t):
        return self.seqs[c].frame(t)


def haar_table(seed, F, w0, h0):
    """HaarFeature::Generate (HaarFeature.cpp:25-69) for F features of a w0 x h0 blob, drawn from OpenCV's multiply-with-carry
    stream seeded like randinitalize(seed) (Public.cpp:10-23), in the reference's draw order.  Returns (nrect [F], rects
    [F][6][4], weights [F][6]) as capi.StrongClassifier(haar=...) takes them."""
    state = [seed if seed else (1 << 64) - 1]

    def nxt():
        t = state[0]
        t = (t & 0xffffffff) * 4164903690 + (t >> 32)
        state[0] = t & ((1 << 64) - 1)
        return state[0] & 0xffffffff

    randint = lambda a, b: nxt() % (b - a + 1) + a                      # noqa: E731
    randfloat = lambda: np.float32(nxt() * 2.3283064365386962890625e-10)  # noqa: E731
    nrect = np.zeros(F, np.int32); rects = np.zeros((F, 6, 4), np.int32); wts = np.zeros((F, 6), np.float32)
    for f in range(F):
        n = randint(2, 6)
        nrect[f] = n
        for k in range(n):
            wts[f, k] = np.float32(randfloat() * np.float32(2) - np.float32(1))
            x = randint(0, w0 - 3); y = randint(0, h0 - 3)
            rects[f, k] = (x, y, randint(1, w0 - x - 2), randint(1, h0 - y - 2))
        randint(0, 0)                                                   # the channel draw (one channel)
    return nrect, rects, wts
