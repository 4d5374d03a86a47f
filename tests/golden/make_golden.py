"""Generates the committed golden fixtures under tests/golden/ (run once, here, in the build container).

  cv2_rng.json   uint32 stream of OpenCV's RNG (the generator behind cvRNG/cvRandInt/cvRandReal that
                 Public.cpp:10-23 uses) recovered from cv2.randu(CV_64F): cv2 returns
                 (int64)(u<<32 | carry) * 2^-64 + 0.5, so u = floor((v - 0.5) * 2^32 mod 2^32).
  cv2_gray.json  cv2.cvtColor(BGR2GRAY) on probe pixels (the conversion upstream of the path, Matrix.cpp:518-533)
  cv2_hsv.json   cv2.cvtColor(BGR2HSV) (8-bit, H in [0,180)) on probe pixels incl. grey / saturated / tie cases
  cv2_round.json cvRound semantics probes (numpy rint == round-half-even == SSE cvtsd2si)
  cv2_fuser.json OpenCV's small-matrix arithmetic behind the geometric fuser (GeometryBasedInformationFuser.cpp): cv2.invert
                 (2x2 f32, 3x3 f64), cv2.gemm (f32), cv2.SVDecomp null vector of the stacked ground-plane lines, and a
                 cv2.KalmanFilter(4, 2, 0) run configured as InitializeKalman does (R reset to diag(5) before every update)
  hand.json      hand-computed integral / sumRect / resampler cases from SURVEY.md 8c
"""
import json
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = {}
    for seed in (1, 7, 12345, 99991):
        cv2.setRNGSeed(seed)
        a = np.zeros((1, 64), np.float64)
        cv2.randu(a, 0, 1)
        u = np.floor(np.mod((a[0] - 0.5) * 4294967296.0, 4294967296.0)).astype(np.uint64)
        rng[str(seed)] = [int(x) for x in u]
    json.dump(rng, open(os.path.join(HERE, "cv2_rng.json"), "w"))

    r = np.random.default_rng(3)
    bgr = r.integers(0, 256, (1, 256, 3), dtype=np.uint8)
    g = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    json.dump({"bgr": bgr[0].tolist(), "gray": g[0].tolist()}, open(os.path.join(HERE, "cv2_gray.json"), "w"))

    r2 = np.random.default_rng(4)
    px = r2.integers(0, 256, (4000, 3), dtype=np.uint8)
    edge = np.array([[0, 0, 0], [255, 255, 255], [10, 10, 10], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [0, 255, 255],
                     [255, 0, 255], [1, 0, 0], [0, 1, 0], [0, 0, 1], [200, 200, 199], [128, 127, 128], [5, 250, 5]], np.uint8)
    bgr2 = np.concatenate([edge, px])[None]
    hsv = cv2.cvtColor(bgr2, cv2.COLOR_BGR2HSV)
    json.dump({"bgr": bgr2[0].tolist(), "hsv": hsv[0].tolist()}, open(os.path.join(HERE, "cv2_hsv.json"), "w"))

    probes = [2.5, 3.5, -2.5, -3.5, 0.5, 1.5, 0.49999, 2.500001, -0.5, 1e6 + 0.5]
    json.dump({"x": probes, "r": [int(np.rint(p)) for p in probes]}, open(os.path.join(HERE, "cv2_round.json"), "w"))

    hand = {
        "integral": {"img": [[1, 2, 3], [4, 5, 6]],
                     "ii": [[0, 0, 0, 0], [0, 1, 3, 6], [0, 5, 12, 21]]},
        "sum_rect": {"img": [[1, 2, 3], [4, 5, 6]], "rect": [1, 0, 2, 2], "sum": 16},
        "resample": {"w": [0.1, 0.2, 0.3, 0.4], "u01": 0.5, "cdf": [0, 0.2, 0.5, 0.9],
                     "thr": [0.125, 0.375, 0.625, 0.875], "idx": [1, 2, 3, 3]},
        "rng_float": {"1": [0.9697172069, 0.2704580063, 0.3693985634, 0.0550134850],
                      "7": [0.7880204483, 0.8932060455, 0.5857899440, 0.3850943956],
                      "12345": [0.1589191691, 0.8040905611, 0.2252659721, 0.1414736339]},
    }
    json.dump(hand, open(os.path.join(HERE, "hand.json"), "w"))


def fuser():
    r = np.random.default_rng(11)
    out = {"inv2": [], "inv3": [], "gemm": [], "svd": [], "kalman": {}}
    for _ in range(6):
        a = r.normal(size=(2, 2)).astype(np.float32)
        S = (a @ a.T + np.eye(2, dtype=np.float32) * r.uniform(0.1, 5)).astype(np.float32)
        ok, Si = cv2.invert(S)
        out["inv2"].append({"S": S.tolist(), "Si": Si.tolist(), "det": float(cv2.determinant(S))})
    Hs = []
    for k in range(4):
        # plausible image -> ground homographies: scale, shear, small perspective terms
        H = np.array([[r.uniform(0.5, 2), r.uniform(-0.3, 0.3), r.uniform(-200, 200)],
                      [r.uniform(-0.3, 0.3), r.uniform(0.5, 2), r.uniform(-200, 200)],
                      [r.uniform(-1e-3, 1e-3), r.uniform(-1e-3, 1e-3), 1.0]], np.float64)
        ok, Hi = cv2.invert(H)
        Hs.append(H)
        out["inv3"].append({"H": H.tolist(), "Hi": Hi.tolist()})
    for _ in range(4):
        A = r.normal(size=(1, 2)).astype(np.float32) * 30
        B = r.normal(size=(2, 2)).astype(np.float32)
        out["gemm"].append({"A": A.tolist(), "B": B.tolist(), "AB": cv2.gemm(A, B, 1.0, None, 0.0).tolist()})
    for C in (2, 3, 4):
        foot = r.uniform(50, 600, size=(C, 2)).astype(np.float32)
        L = np.zeros((C, 3))
        for c in range(C):
            L[c] = np.array([1.0, 0.0, -float(foot[c, 0])]) @ cv2.invert(Hs[c])[1]
        w, u, vt = cv2.SVDecomp(L, flags=cv2.SVD_FULL_UV)
        v = vt[2]                                  # cvSVD's V (not transposed) column 2 = V->data.db[2], [5], [8]
        out["svd"].append({"foot": foot.tolist(), "H": [Hs[c].tolist() for c in range(C)], "L": L.tolist(),
                           "g": [float(v[0] / v[2]), float(v[1] / v[2])]})
    kf = cv2.KalmanFilter(4, 2, 0)
    kf.transitionMatrix = np.array([[1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    kf.measurementMatrix = np.array([[1, 0, 0, 0], [0, 1, 0, 0]], np.float32)
    kf.processNoiseCov = np.eye(4, dtype=np.float32)
    kf.measurementNoiseCov = np.array([[10, 0], [0, 10]], np.float32)
    kf.errorCovPost = (np.eye(4) * 100).astype(np.float32)
    kf.statePost = np.array([[123.5], [-40.25], [0], [0]], np.float32)
    zs = (np.array([123.5, -40.25]) + np.cumsum(r.normal(size=(8, 2)) * 3 + np.array([2.0, -1.0]), axis=0)).astype(np.float32)
    states, covs = [], []
    for z in zs:
        kf.measurementNoiseCov = np.array([[5, 0], [0, 5]], np.float32)
        kf.predict()
        kf.correct(z.reshape(2, 1))
        states.append(kf.statePost[:, 0].tolist()); covs.append(kf.errorCovPost.tolist())
    out["kalman"] = {"x0": [123.5, -40.25], "z": zs.tolist(), "state_post": states, "cov_post": covs}
    json.dump(out, open(os.path.join(HERE, "cv2_fuser.json"), "w"))


if __name__ == "__main__":
    main()
    fuser()
