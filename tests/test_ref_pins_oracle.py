"""CPU tests that PIN THE ORACLE TO THE REFERENCE'S OWN CODE.

oracle/_ref/libmct_ref.so is the unmodified reference (/root/reference/src/*.cpp) compiled against header shims for IPP /
OpenCV / Boost (oracle/build_ref.py, oracle/ref_shim/) and driven through oracle/ref_harness.cpp.  Every test here runs the
reference and oracle/mct_oracle.c on the same seeded inputs, row by row of SURVEY 8a, and requires

  * bit-exact equality for integer / index work and for everything whose f32 / f64 operation order the oracle restates
    (integral image, rect sums, Haar features, samplers and RNG consumption, selector lists, resampling indices,
    particle states, tracked boxes);
  * 1e-5 relative where the oracle accumulates in another precision than the reference by design.

Stated assumptions that remain (oracle/ref_shim/ipp.h): the integral image and the mean / variance reductions are the
shim's restatement of closed-source IPP; std::sort's order of exactly tied scores is unspecified by the language: the
default _ref build keeps tied candidates in index order, libmct_ref_native.so has libstdc++'s introsort
(test_std_sort_tie_order_is_the_only_difference_of_the_native_build).

Skipped when the library is not built and /root/reference is absent (the GPU box receives the prebuilt library)."""
import ctypes as C

import numpy as np
import pytest

from multicameratracking_b200 import synth
from oracle import oracle_py
from oracle.ref_py import Ref, TrackerParams

pytestmark = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built and /root/reference absent")

ONE = lambda n: np.ones(n, np.float32)


@pytest.fixture(scope="module")
def ref():
    return Ref()


@pytest.fixture(scope="module")
def seq():
    return synth.Sequence(640, 480, 3, n_frames=12)


def _boxes(seq, n, seed=0, smin=0.8, smax=1.25):
    rng = np.random.default_rng(seed)
    sx = rng.uniform(smin, smax, n).astype(np.float32); sy = rng.uniform(smin, smax, n).astype(np.float32)
    col = rng.integers(1, seq.W - int(seq.w0 * smax) - 3, n).astype(np.int32)
    row = rng.integers(1, seq.H - int(seq.h0 * smax) - 3, n).astype(np.int32)
    return col, row, sx, sy


# ---------------------------------------------------------------- a1, a2, RNG, cvRound, frame pre-processing
def test_rng_streams_and_cvround(oracle, ref):
    for seed in (1, 7, 12345, 0, -3):
        ref.seed(seed); r = oracle.rng(seed)
        assert [ref.randfloat() for _ in range(200)] == [oracle.randfloat(r) for _ in range(200)]
        assert [ref.randint(2, 6) for _ in range(50)] == [oracle.randint(r, 2, 6) for _ in range(50)]
        assert [ref.randint(0, 37) for _ in range(50)] == [oracle.randint(r, 0, 37) for _ in range(50)]
    for v in (2.5, 3.5, -2.5, 0.49999, 1e6 + 0.5, -0.5, 17.500001):
        assert ref.lib.ref_cvround(v) == oracle.lib.orc_cvround(v)


@pytest.mark.parametrize("shape", [(640, 480), (97, 33), (1280, 720)])
def test_integral_and_sum_rect(oracle, ref, shape):
    W, H = shape
    rng = np.random.default_rng(W)
    img = rng.integers(0, 256, (H, W)).astype(np.uint8)
    if W == 1280:
        img[:] = 255                                   # table entries beyond 2^24: the single rounding is what is compared
    fr = ref.frame(img)
    ii_r = ref.integral(fr); ii_o = oracle.integral(img)
    np.testing.assert_array_equal(ii_r, ii_o)
    for _ in range(300):
        x = int(rng.integers(0, W - 1)); y = int(rng.integers(0, H - 1))
        w = int(rng.integers(1, W - x + 1)); h = int(rng.integers(1, H - y + 1))
        assert ref.sum_rect(fr, x, y, w, h) == oracle.sum_rect(ii_o, x, y, w, h)
    ref.frame_free(fr)


def test_bgr_unpack_and_gray(oracle, ref):
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, (60, 83, 3)).astype(np.uint8)
    g_r, r_r, gg_r, b_r = ref.bgr_to_planes(bgr)
    g_o, r_o, gg_o, b_o = oracle.bgr_to_planes(bgr, use_hsv=False)
    for a, b in ((g_r, g_o), (r_r, r_o), (gg_r, gg_o), (b_r, b_o)):
        np.testing.assert_array_equal(a, b)


# ---------------------------------------------------------------- a4 samplers
def test_samplers_and_thinning(oracle, ref, seq):
    gray, cr, cg, cb = seq.frame(0)
    fr = ref.frame(gray, cr, cg, cb)
    ref.seed(3); r = oracle.rng(3)
    cases = [(7.0, 0.0, 1000000), (10.0, 0.0, 1000000), (40.0, 10.5, 30), (30.0, 25.0, 30), (40.0, 10.5, 65)]
    for o in range(3):
        x, y, w, h = seq.box(0, o)
        for outer, inner, mx in cases:
            a = ref.sample_ring(fr, x, y, w, h, outer, inner, mx)
            b = oracle.sample_ring(r, seq.W, seq.H, x, y, w, h, outer, inner, mx, cap=1 << 20)
            np.testing.assert_array_equal(a[0], b[0]); np.testing.assert_array_equal(a[1], b[1])
        for (mxx, mxy, mnx, mny, mx) in [(40.0, 40.0, 12.0, 9.0, 30), (25.5, 31.0, 0.0, 0.0, 1000000), (60.0, 20.0, 15.0, 15.0, 50)]:
            a = ref.sample_rect(fr, x, y, w, h, mxx, mxy, mnx, mny, mx)
            b = oracle.sample_rect(r, seq.W, seq.H, x, y, w, h, mxx, mxy, mnx, mny, mx, cap=1 << 20)
            np.testing.assert_array_equal(a[0], b[0]); np.testing.assert_array_equal(a[1], b[1])
    # the whole-image form (Negative_Sampling_Strategy 0: outer radius beyond the frame), once: the reference needs seconds for it
    a = ref.sample_ring(fr, 300, 200, seq.w0, seq.h0, 2000.0, 150.0, 100); b = oracle.sample_ring(r, seq.W, seq.H, 300, 200, seq.w0, seq.h0, 2000.0, 150.0, 100, cap=1 << 20)
    np.testing.assert_array_equal(a[0], b[0]); np.testing.assert_array_equal(a[1], b[1])
    # a box near the border (clipped window) and the RNG consumption
    a = ref.sample_ring(fr, 3, 2, seq.w0, seq.h0, 30.0, 5.0, 40); b = oracle.sample_ring(r, seq.W, seq.H, 3, 2, seq.w0, seq.h0, 30.0, 5.0, 40)
    np.testing.assert_array_equal(a[0], b[0]); np.testing.assert_array_equal(a[1], b[1])
    assert [ref.randfloat() for _ in range(3)] == [oracle.randfloat(r) for _ in range(3)]
    ref.frame_free(fr)


# ---------------------------------------------------------------- a5 - a9 features
def test_haar_table_and_features(oracle, ref, seq):
    gray, cr, cg, cb = seq.frame(0)
    fr = ref.frame(gray, cr, cg, cb); of = oracle.frame(gray, cr, cg, cb)
    for seed, F in ((7, 250), (1, 64)):
        ref.seed(seed); r = oracle.rng(seed)
        t = oracle.haar_generate(r, F, seq.w0, seq.h0)
        oc = oracle.clf_create(0, F, 8, seq.w0, seq.h0, 0, 2, F // 5, 0.85, t)
        rc = ref.clf_create(0, F, 8, seq.w0, seq.h0, 0, 2, F // 5, 0.85)
        for a, b in zip(oracle.haar_arrays(t), ref.clf_haar_table(rc, F)):
            np.testing.assert_array_equal(a, b)
        assert ref.randfloat() == oracle.randfloat(r)                  # Generate consumed exactly the same draws
        col, row, sx, sy = _boxes(seq, 400, seed)
        np.testing.assert_array_equal(ref.clf_features(rc, fr, col, row, sx, sy), oracle.clf_features(oc, of, oracle.boxes(col, row, sx, sy)))
        ref.lib.ref_clf_free(rc); oracle.lib.orc_clf_free(oc)
    ref.frame_free(fr)


@pytest.mark.parametrize("nb", [8, 4])
def test_mdch_features(oracle, ref, seq, nb):
    gray, cr, cg, cb = seq.frame(1)
    fr = ref.frame(gray, cr, cg, cb); of = oracle.frame(gray, cr, cg, cb)
    F = nb ** 3
    oc = oracle.clf_create(2, 0, nb, seq.w0, seq.h0, 0, 4, F // 5, 0.85, None)
    rc = ref.clf_create(2, 0, nb, seq.w0, seq.h0, 0, 4, F // 5, 0.85)
    col, row, sx, sy = _boxes(seq, 120, nb)
    a = ref.clf_features(rc, fr, col, row, sx, sy); b = oracle.clf_features(oc, of, oracle.boxes(col, row, sx, sy))
    np.testing.assert_array_equal(a, b)
    # concatenated Haar + MDCH layout (HaarAndColorHistogramFeatureVector.cpp:12-44): rows [0,Fh) Haar, [Fh, Fh + nb^3) colour
    ref.seed(9); r = oracle.rng(9)
    t = oracle.haar_generate(r, 30, seq.w0, seq.h0)
    oc2 = oracle.clf_create(3, 30, nb, seq.w0, seq.h0, 0, 2, 10, 0.85, t)
    rc2 = ref.clf_create(3, 30, nb, seq.w0, seq.h0, 0, 2, 10, 0.85)
    a = ref.clf_features(rc2, fr, col, row, sx, sy); b = oracle.clf_features(oc2, of, oracle.boxes(col, row, sx, sy))
    assert a.shape == (30 + F, 120)
    np.testing.assert_array_equal(a, b)
    ref.frame_free(fr)


# ---------------------------------------------------------------- a10 - a13 weak classifiers
@pytest.mark.parametrize("weak_type,lr", [(0, 0.85), (0, 0.0), (1, 0.85), (2, 0.85)])
def test_weak_classifiers_on_given_feature_values(oracle, ref, weak_type, lr):
    """OnlineStumps / WeightedStumps (explicit weights: the NULL-weights path crashes in the reference) / Perceptron:
    state after 1 and 3 updates and ClassifyF, against the oracle's strong classifier fed the same single feature row"""
    rng = np.random.default_rng(weak_type * 10 + int(lr * 100))
    for trial in range(4):
        P, Nn = int(rng.integers(5, 300)), int(rng.integers(3, 60))
        xp = (rng.normal(3.0, 2.0, P) * 50).astype(np.float32); xn = (rng.normal(-1.0, 4.0, Nn) * 50).astype(np.float32)
        xt = (rng.normal(0.0, 5.0, 64) * 50).astype(np.float32)
        wp = rng.uniform(0.1, 2.0, P).astype(np.float32) if weak_type == 1 else None
        wn = rng.uniform(0.1, 2.0, Nn).astype(np.float32) if weak_type == 1 else None
        for n_up in (1, 3):
            st_r, out_r, valid = ref.weak_run(weak_type, lr, n_up, xp, xn, wp, wn, xt)
            oc = oracle.clf_create(2, 0, 1, 4, 4, weak_type, 2, 1, lr, None)          # F = 1 (nb = 1): one weak classifier
            for _ in range(n_up):
                oracle.clf_update_from_features(oc, xp.reshape(1, P), xn.reshape(1, Nn), wp, wn)
            st_o = oracle.clf_weak_state(oc)[:, 0]
            k = 2 if weak_type == 2 else 8
            np.testing.assert_allclose(st_o[:k], st_r[:k], rtol=1e-6, atol=0)
            if valid:
                out_o = oracle.clf_classify_from_features(oc, xt.reshape(1, -1))
                np.testing.assert_allclose(out_o, out_r, rtol=1e-5, atol=1e-6)
            oracle.lib.orc_clf_free(oc)


# ---------------------------------------------------------------- a14 - a17 (+ f4) strong classifiers
STRONG_CASES = {
    "haar_stump_milboost": (0, 0, 2, 100, 20, 0.0),
    "haarmdch_stump_milboost": (3, 0, 2, 60, 114, 0.0),
    "mdch_stump_anyboost": (2, 0, 4, 0, 102, 0.0),
    "haar_stump_anyboost": (0, 0, 4, 80, 16, 0.0),
    "haar_perceptron_milboost": (0, 2, 2, 80, 16, 0.0),
    "haar_stump_adaboost": (0, 0, 0, 80, 16, 0.0),
    "haar_perceptron_ensemble": (0, 2, 3, 80, 16, 50.0),
    # MILEnsemble over stumps is not pinnable: the first retain step classifies with the UNTRAINED stump 0, whose m_n0/m_e0 are
    # uninitialised memory in the reference (OnlineStumpsWeakClassifier::Initialize :53-60 sets only mu and sigma)
    "k_clamped_to_half": (0, 0, 2, 40, 0, 0.0),          # StrongClassifierBase.cpp:20-23: K < 1 -> F / 2
    "k_larger_than_f": (0, 0, 2, 40, 41, 0.0),
}


@pytest.mark.parametrize("case", list(STRONG_CASES))
def test_strong_classifier_update_and_classify(oracle, ref, seq, case):
    ft, weak, strong, nh, K, pct = STRONG_CASES[case]
    ref.seed(11); r = oracle.rng(11)
    t = oracle.haar_generate(r, nh, seq.w0, seq.h0) if ft in (0, 3) else None
    oc = oracle.clf_create(ft, nh, 8, seq.w0, seq.h0, weak, strong, K, 0.85, t, pct_retained=pct)
    rc = ref.clf_create(ft, nh, 8, seq.w0, seq.h0, weak, strong, K, 0.85, 0, pct)
    col, row, sx, sy = _boxes(seq, 150, 3, 0.9, 1.1)
    for it in range(4):
        g = seq.frame(it); fr = ref.frame(*g); of = oracle.frame(*g)
        x, y, w, h = seq.box(it, it % 3)
        pc, pr = oracle.sample_ring(r, seq.W, seq.H, x, y, w, h, 5.0, 0.0, 1000000)
        nc, nr = oracle.sample_ring(r, seq.W, seq.H, x, y, w, h, 40.0, 8.0, 30)
        pc2, pr2 = ref.sample_ring(fr, x, y, w, h, 5.0, 0.0, 1000000)
        nc2, nr2 = ref.sample_ring(fr, x, y, w, h, 40.0, 8.0, 30)
        np.testing.assert_array_equal(pc, pc2); np.testing.assert_array_equal(nc, nc2)
        oracle.clf_update(oc, of, oracle.boxes(pc, pr, ONE(len(pc)), ONE(len(pc))), oracle.boxes(nc, nr, ONE(len(nc)), ONE(len(nc))))
        ref.clf_update(rc, fr, (pc, pr, ONE(len(pc)), ONE(len(pc))), (nc, nr, ONE(len(nc)), ONE(len(nc))))
        assert oracle.clf_selectors(oc).tolist() == ref.clf_selectors(rc).tolist(), (case, it)
        k = 2 if weak == 2 else 8
        np.testing.assert_allclose(oracle.clf_weak_state(oc)[:k], ref.clf_weak_state(rc)[:k], rtol=1e-6, atol=0, err_msg="%s frame %d" % (case, it))
        if strong == 0:
            np.testing.assert_allclose(oracle.clf_alphas(oc), ref.clf_alphas(rc), rtol=1e-6)
        for log_ratio in (True, False):
            Ho = oracle.clf_classify(oc, of, oracle.boxes(col, row, sx, sy), log_ratio)
            Hr = ref.clf_classify(rc, fr, col, row, sx, sy, log_ratio)
            np.testing.assert_allclose(Ho, Hr, rtol=1e-5, atol=1e-5 * max(1e-30, float(np.abs(Hr).max())), err_msg="%s frame %d" % (case, it))
        ref.frame_free(fr)
    ref.lib.ref_clf_free(rc); oracle.lib.orc_clf_free(oc)


def test_std_sort_tie_order_is_the_only_difference_of_the_native_build(oracle, seq):
    """libmct_ref_native.so = the same sources with libstdc++'s own std::sort.  On a case whose candidate scores tie exactly
    (perceptron votes are +-1) the selector lists differ from the index-order build; on a case without exact ties they agree."""
    from oracle import build_ref
    path = build_ref.build(native=True)
    if path is None:
        pytest.skip("native variant not built")
    L = C.CDLL(path)
    Ref._proto(L)
    native = Ref.__new__(Ref); native.lib = L; L.ref_set_log(0)
    stable = Ref()
    out = {}
    for name, lib in (("native", native), ("stable", stable)):
        res = []
        for (ft, weak, strong, nh, K) in ((0, 0, 2, 100, 20), (0, 2, 2, 80, 16)):
            lib.seed(11)
            rc = lib.clf_create(ft, nh, 8, seq.w0, seq.h0, weak, strong, K, 0.85)
            g = seq.frame(1); fr = lib.frame(*g)
            x, y, w, h = seq.box(1, 0)
            pc, pr = lib.sample_ring(fr, x, y, w, h, 5.0, 0.0, 1000000); nc, nr = lib.sample_ring(fr, x, y, w, h, 40.0, 8.0, 30)
            lib.clf_update(rc, fr, (pc, pr, ONE(len(pc)), ONE(len(pc))), (nc, nr, ONE(len(nc)), ONE(len(nc))))
            res.append(lib.clf_selectors(rc).tolist())
            lib.frame_free(fr)
        out[name] = res
    # stumps: the leading selections (distinct scores) agree between the builds; the tail, picked among exactly tied scores
    # once the bags saturate, may not
    n_common = 0
    for a, b in zip(out["native"][0], out["stable"][0]):
        if a != b:
            break
        n_common += 1
    assert n_common >= 3
    # perceptron votes (+-1) tie from the first round on: the two builds pick different candidates of equal score
    assert len(out["native"][1]) == len(out["stable"][1])
    assert out["native"] != out["stable"]


# ---------------------------------------------------------------- a18 - a25 particle filter
def _pf_pair(oracle, ref, N, cx, cy, W=640, H=480):
    rp = ref.pf_create(N, W, H, cx, cy)
    op = oracle.pf_create(N, W, H)
    oracle.lib.orc_pf_initialize(op, cx, cy, 1, 1, 0.5, 10, 0.1)
    return rp, op


def _pf_equal(oracle, ref, rp, op, what=""):
    st, rs, idx, fs = ref.pf_get(rp)
    np.testing.assert_array_equal(st, oracle.pf_state(op), err_msg="state " + what)
    np.testing.assert_array_equal(rs, oracle.pf_resampled(op), err_msg="resampled " + what)
    assert fs == op.contents.filter_state, what
    return st, rs, idx


@pytest.mark.parametrize("N", [4, 1000, 2000])
def test_particle_filter_predict_update_resample_readout(oracle, ref, N):
    rng = np.random.default_rng(N)
    w0, h0 = 40.0, 80.0
    rp, op = _pf_pair(oracle, ref, N, 300.0, 200.0)
    _pf_equal(oracle, ref, rp, op, "after Initialize")
    ref.seed(5); r = oracle.rng(5)
    for t in range(6):
        sd = (20.0, 20.0, 0.05, 0.05) if t % 2 else (20.0, 20.0, 1e-7, 1e-7)
        nz = np.stack([(rng.standard_normal(N) * s).astype(np.float32) for s in sd])
        if t == 3:
            nz[2] *= 30                                           # exercise the max-scale-change clamp
        ref.pf_predict(rp, nz, sd, w0, h0)
        oracle.lib.orc_pf_predict(op, nz.ctypes.data_as(C.POINTER(C.c_float)), w0, h0)
        _pf_equal(oracle, ref, rp, op, "after predict %d" % t)
        # weights: peaked likelihoods so that N_eff collapses in some frames and not in others
        st = ref.pf_get(rp)[0]
        d2 = (st[:, 0] - 300.0 - 3 * t) ** 2 + (st[:, 1] - 200.0) ** 2
        prob = np.exp(-d2 / (2 * (6.0 if t % 3 else 60.0) ** 2)).astype(np.float32)
        if t == 4:
            prob[:] = 0                                           # total weight 0 -> all weights 1e-2 (a20)
        n_nonzero = int((st[:, 4] != 0).sum())
        ref.pf_update_all_weights(rp, prob[:n_nonzero])
        resampled = oracle.lib.orc_pf_update_all_weights(op, prob.ctypes.data_as(C.POINTER(C.c_float)), 1, 0, 1, C.byref(r))
        st, rs, idx = _pf_equal(oracle, ref, rp, op, "after update %d" % t)
        if resampled:
            np.testing.assert_array_equal(idx, oracle.pf_resample_idx(op))
        assert ref.randfloat() == oracle.randfloat(r)             # same number of draws consumed
        a = np.zeros(4, np.float32); oracle.lib.orc_pf_get_average(op, a.ctypes.data_as(C.POINTER(C.c_float)))
        np.testing.assert_array_equal(ref.pf_average(rp), a)
        b = np.zeros(5, np.float32); oracle.lib.orc_pf_get_highest_weight(op, b.ctypes.data_as(C.POINTER(C.c_float)))
        np.testing.assert_array_equal(ref.pf_highest(rp), b)
        uq = ref.pf_ordered_unique(rp)
        oracle.lib.orc_pf_find_ordered_unique(op)
        ou = np.ctypeslib.as_array(op.contents.uniq, (op.contents.n_uniq, 5))
        assert len(uq) == op.contents.n_uniq
        np.testing.assert_array_equal(uq, ou)
        c2 = np.array([300.0, 200.0], np.float32); b2 = np.zeros(5, np.float32)
        oracle.lib.orc_pf_get_highest_unique_close(op, c2.ctypes.data_as(C.POINTER(C.c_float)), b2.ctypes.data_as(C.POINTER(C.c_float)))
        np.testing.assert_array_equal(ref.pf_highest_unique_close(rp, c2), b2)
    ref.lib.ref_pf_free(rp[0]); oracle.lib.orc_pf_free(op)


def test_forced_resample_and_index_list(oracle, ref):
    """ResampleParticles(false) as the fusion path calls it every frame (Object.cpp:449): indices bit-exact, state untouched"""
    N = 1500
    rng = np.random.default_rng(2)
    rp, op = _pf_pair(oracle, ref, N, 200.0, 150.0)
    st = np.zeros((N, 5), np.float32)
    st[:, 0] = rng.uniform(50, 500, N); st[:, 1] = rng.uniform(50, 400, N); st[:, 2:4] = 1
    st[:, 4] = rng.uniform(0, 1, N) ** 8
    ref.pf_set(rp, st, None, 2)
    oracle.pf_state(op)[:] = st; op.contents.filter_state = 2
    ref.seed(9); r = oracle.rng(9)
    ref.lib.ref_pf_resample(rp[0], 0)
    oracle.lib.orc_pf_resample.argtypes = [C.POINTER(oracle_py.Pf), C.c_int, C.POINTER(oracle_py.Rng)]
    oracle.lib.orc_pf_resample(op, 0, C.byref(r))
    s2, rs, idx = _pf_equal(oracle, ref, rp, op, "after forced resample")
    np.testing.assert_array_equal(idx, oracle.pf_resample_idx(op))
    ref.lib.ref_pf_free(rp[0]); oracle.lib.orc_pf_free(op)


# ---------------------------------------------------------------- the tracker frame loop
TRACKER_CASES = {
    # name: (n_obj, n_frames, overrides)
    "cfg1_haar_milboost_1000": (1, 8, dict()),
    "haarmdch_stump_2obj": (2, 6, dict(feature_type=3, n_particles=400, n_haar=60)),
    "mdch_anyboost_avg": (1, 8, dict(feature_type=2, strong_type=4, n_particles=400, trajectory_option=1)),
    "greedy_pos_random_neg": (2, 8, dict(n_particles=500, n_haar=80, pos_strategy=1, neg_strategy=1, max_num_pos_examples=60)),
    "semi_greedy_pos": (1, 8, dict(n_particles=500, n_haar=80, pos_strategy=2, max_num_pos_examples=300, sd_x=4.0, sd_y=4.0)),
    "particle_random_pos": (2, 8, dict(n_particles=600, n_haar=80, pos_strategy=3, neg_strategy=1, max_num_pos_examples=50, trajectory_option=1)),
    "haar_adaboost": (1, 6, dict(strong_type=0, n_particles=400, n_haar=80)),
    "haar_ensemble": (1, 6, dict(strong_type=3, weak_type=2, n_particles=400, n_haar=80, percent_retained=50.0)),
}
DEFAULTS = dict(feature_type=0, n_haar=250, n_bins=8, weak_type=0, strong_type=2, percent_selected=20, n_particles=1000,
                sd_x=20.0, sd_y=20.0, sd_sx=1e-7, sd_sy=1e-7, pos_radius_train=10.0, init_pos_radius_train=7.0,
                n_neg_train=30, init_n_neg_train=30, search_win=20, trajectory_option=0, percent_retained=0.0,
                pos_strategy=0, neg_strategy=0, max_num_pos_examples=30)


@pytest.mark.parametrize("case", list(TRACKER_CASES))
def test_particle_filter_tracker_sequence(oracle, ref, case):
    """ParticleFilterTracker::InitializeTrackerWithParameters + TrackObjectAndSaveState of the REFERENCE against the oracle's
    restatement of the frame loop: boxes, particle matrices, selectors and the shared RNG stream equal after every frame.
    All objects of the camera share the reference's one global MWC stream (Public.h:156); objects are processed in order."""
    n_obj, n_frames, over = TRACKER_CASES[case]
    p = dict(DEFAULTS); p.update(over)
    s = synth.Sequence(640, 480, n_obj, n_frames=n_frames)
    F = {0: p["n_haar"], 1: 11, 2: 512, 3: p["n_haar"] + 512}[p["feature_type"]]
    K = int(F * p["percent_selected"] / 100)
    g = s.frame(0)
    fr = ref.frame(*g); of = oracle.frame(*g)
    ref.seed(1); r = oracle.rng(1)
    rt, ot = [], []
    for o in range(n_obj):
        x, y, w, h = s.box(0, o)
        init = np.array([x, y, w, h], np.float32)
        tp_r = TrackerParams(p["n_particles"], p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"], p["pos_radius_train"], p["init_pos_radius_train"],
                             p["n_neg_train"], p["init_n_neg_train"], 100000, p["search_win"], p["trajectory_option"], p["pos_strategy"],
                             p["neg_strategy"], p["max_num_pos_examples"])
        rt.append(ref.tracker_create(True, tp_r, p["feature_type"], p["n_haar"], p["n_bins"], p["weak_type"], p["strong_type"], K, 0.85, fr,
                                     init, n_frames, pct_retained=p["percent_retained"]))
        t = oracle.haar_generate(r, p["n_haar"], s.w0, s.h0) if p["feature_type"] in (0, 3) else None
        clf = oracle.clf_create(p["feature_type"], p["n_haar"], p["n_bins"], s.w0, s.h0, p["weak_type"], p["strong_type"], K, 0.85, t,
                                pct_retained=p["percent_retained"])
        tp_o = oracle_py.TrackerParams(*[getattr(tp_r, f[0]) for f in TrackerParams._fields_])
        trk = oracle.lib.orc_tracker_create(C.byref(tp_o), clf, C.byref(of), init.ctypes.data_as(C.POINTER(C.c_float)), C.byref(r))
        ot.append((trk, clf))
        assert ref.tracker_selectors(rt[o], F).tolist() == oracle.clf_selectors(clf).tolist(), (case, "init", o)
    N = p["n_particles"]
    for t in range(1, n_frames):
        g = s.frame(t)
        ref.frame_free(fr)
        fr = ref.frame(*g); of = oracle.frame(*g)
        for o in range(n_obj):
            nz = synth.noise(N, (p["sd_x"], p["sd_y"], p["sd_sx"], p["sd_sy"]), obj=o, t=t)
            rb = ref.tracker_track(rt[o], fr, nz, 3)
            ob = np.zeros(4, np.int32)
            oracle.lib.orc_tracker_track(ot[o][0], C.byref(of), nz.ctypes.data_as(C.POINTER(C.c_float)), ob.ctypes.data_as(C.POINTER(C.c_int)), 3)
            np.testing.assert_array_equal(rb, ob, err_msg="%s frame %d object %d" % (case, t, o))
            st, rs, idx, fs = ref.pf_get(ref.tracker_pf(rt[o]))
            opf = oracle.lib.orc_tracker_pf(ot[o][0])
            np.testing.assert_array_equal(st, oracle.pf_state(opf), err_msg="%s frame %d object %d state" % (case, t, o))
            np.testing.assert_array_equal(rs, oracle.pf_resampled(opf), err_msg="%s frame %d object %d resampled" % (case, t, o))
            assert ref.tracker_selectors(rt[o], F).tolist() == oracle.clf_selectors(ot[o][1]).tolist(), (case, t, o)
            np.testing.assert_array_equal(ref.tracker_state(rt[o]), np.ctypeslib.as_array(oracle.lib.orc_tracker_state(ot[o][0]), (6,)))
    assert [ref.randfloat() for _ in range(2)] == [oracle.randfloat(r) for _ in range(2)], case
    for o in range(n_obj):
        ref.lib.ref_tracker_free(rt[o][0]); oracle.lib.orc_tracker_free(ot[o][0]); oracle.lib.orc_clf_free(ot[o][1])
    ref.frame_free(fr)
