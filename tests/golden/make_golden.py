"""Generates the committed golden fixtures under tests/golden/ (run once, here, in the build container).

  cv2_rng.json   uint32 stream of OpenCV's RNG (the generator behind cvRNG/cvRandInt/cvRandReal that
                 Public.cpp:10-23 uses) recovered from cv2.randu(CV_64F): cv2 returns
                 (int64)(u<<32 | carry) * 2^-64 + 0.5, so u = floor((v - 0.5) * 2^32 mod 2^32).
  cv2_gray.json  cv2.cvtColor(BGR2GRAY) on probe pixels (the conversion upstream of the path, Matrix.cpp:518-533)
  cv2_hsv.json   cv2.cvtColor(BGR2HSV) (8-bit, H in [0,180)) on probe pixels incl. grey / saturated / tie cases
  cv2_round.json cvRound semantics probes (numpy rint == round-half-even == SSE cvtsd2si)
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


if __name__ == "__main__":
    main()
