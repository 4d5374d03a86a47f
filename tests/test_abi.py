"""CPU tests of the drop-in boundary: libmct_b200.so builds for sm_100a, loads, exports every symbol
include/mct_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import math
import os
import re

import numpy as np

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mct_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mct_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from multicameratracking_b200 import capi
    lib = capi.load()
    declared = _header_symbols()
    assert sorted(capi.SYMBOLS) == declared
    for s in declared:
        assert hasattr(lib, s), s


def test_ctypes_mirrors_have_the_layout_of_the_c_structs():
    """every Structure in capi.py is as large as the struct the library was compiled with (no device needed)"""
    import ctypes as C
    from multicameratracking_b200 import capi
    L = capi.load()
    L.mct_abi_struct_sizes.argtypes = [C.POINTER(C.c_int), C.c_int]
    out = (C.c_int * 32)()
    n = L.mct_abi_struct_sizes(out, 32)
    layouts = capi.abi_struct_layouts()
    assert n == len(layouts)
    for k, (name, mirror) in enumerate(layouts):
        assert C.sizeof(mirror) == out[k], (name, C.sizeof(mirror), out[k])


def test_host_object_params_mirror_has_the_layout_of_the_c_struct():
    import ctypes as C
    from multicameratracking_b200 import host as mhost
    L = mhost.load()
    assert C.sizeof(mhost.ObjectParams) == L.mcth_abi_object_params_size()


def test_library_is_sm100a_only():
    import subprocess
    from multicameratracking_b200 import capi
    capi.load()
    out = subprocess.run(["cuobjdump", "--list-elf", capi.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from multicameratracking_b200 import capi
    with pytest.raises(capi.MctError) as e:
        capi.Context(0)
    assert e.value.code == capi.E_CUDA


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multicameratracking_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "mct_oracle" not in txt and "libmct_oracle" not in txt, f


# ---------------------------------------------------------------- host side of the geometric fuser (no GPU needed)
def _fuser_gold():
    import json
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_fuser.json")))


def test_host_fuser_principal_axis_matches_oracle_and_opencv():
    from multicameratracking_b200 import host
    from oracle import fuser_oracle as fo
    for c in _fuser_gold()["svd"]:
        foot = np.array(c["foot"], np.float32)
        g = host.principal_axis_intersection(c["H"], foot)
        np.testing.assert_allclose(g, np.array(c["g"]), rtol=1e-9)                  # cv2.SVDecomp
        np.testing.assert_allclose(g, fo.principal_axis_intersection(foot, c["H"]), rtol=1e-9)


def test_host_fuser_kalman_matches_oracle_and_opencv():
    """cvCreateKalman(4, 2, 0) as InitializeKalman configures it; a fuser whose two cameras' axes meet where the golden
    measurement sequence lies reproduces cv2.KalmanFilter's posterior"""
    from multicameratracking_b200 import host
    from oracle import fuser_oracle as fo
    k = _fuser_gold()["kalman"]
    # camera 0 sees ground x as image x, camera 1 sees ground y as image x: the intersection IS the foot-point pair
    H0 = np.eye(3)
    H1 = np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    f = host.GeometricFuser([H0, H1])
    kf = None
    x0 = np.array(k["x0"], np.float32)
    f.initialize(np.array([[x0[0], 0.0], [x0[1], 0.0]], np.float32))
    kf = fo.KalmanCV(x0[0], x0[1])
    for z, s, P in zip(k["z"], k["state_post"], k["cov_post"]):
        z = np.array(z, np.float32)
        f.fuse(np.array([[z[0], 7.0], [z[1], -3.0]], np.float32))
        kf.update(z, np.array([[5, 0], [0, 5]], np.float32))
        mean, cov, st, Pp = f.posterior()
        np.testing.assert_allclose(st, np.array(s, np.float32), rtol=1e-5, atol=1e-5)           # cv2
        np.testing.assert_allclose(Pp, np.array(P, np.float32), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(st, kf.x_post[:, 0], rtol=1e-6, atol=1e-6)                   # oracle
        np.testing.assert_allclose(cov, kf.cov(), rtol=1e-6, atol=1e-6)
    f.close()
    with pytest.raises(host.HostError):
        host.GeometricFuser([H0])


def test_host_randgaus_stream_matches_restatement(oracle):
    """Public.cpp:37-49 polar Box-Muller over randfloat(), the noise of RearrangeParticlesBasedOnGroundLocation"""
    from multicameratracking_b200 import host
    got = host.randgaus_stream(4242, 10.0, 64)
    r = oracle.rng(4242)
    exp = []
    for _ in range(64):
        while True:
            # `-1 + 2 * randfloat()` is float arithmetic in C++ (int * float), widened to double only on assignment
            x = float(np.float32(-1) + np.float32(2) * np.float32(oracle.randfloat(r)))
            y = float(np.float32(-1) + np.float32(2) * np.float32(oracle.randfloat(r)))
            r2 = x * x + y * y
            if not (r2 > 1.0 or r2 == 0):
                break
        exp.append(np.float32(np.float32(10.0 * y * math.sqrt(-2.0 * math.log(r2) / r2)) + np.float32(0)))   # libm, as the C++ side
    np.testing.assert_array_equal(got, np.array(exp, np.float32))


def test_host_sampler_equals_the_naive_listing(tmp_path):
    """SampleSet::SampleImage of the host mirror enumerates the candidates as per-row column runs and never builds the larger
    set; tests/cpp/host_sampler_check.cpp compares it with the naive enumerate-then-thin listing of SampleSet.cpp on 20 000
    random regions (samples, order, RNG state).  Host code only: runs without a GPU."""
    import subprocess
    from multicameratracking_b200 import build
    from multicameratracking_b200.host import build_host
    lib_dev, lib_host = build.build(), build_host.build()
    hd, dd = os.path.dirname(lib_host), os.path.dirname(lib_dev)
    exe = str(tmp_path / "host_sampler_check")
    cmd = ["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "host_sampler_check.cpp"), "-o", exe,
           "-I", hd, "-I", os.path.join(ROOT, "include"), "-L", hd, "-lmct_host", "-L", dd, "-lmct_b200",
           "-Wl,-rpath," + hd, "-Wl,-rpath," + dd]
    subprocess.run(cmd, check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad 0" in r.stdout and int(r.stdout.split("thinned cases")[1]) > 1000
