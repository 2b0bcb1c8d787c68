"""Pins the CPU oracle (oracle/vq_oracle.c) against the reference's own known-answer
tests and against hsdlib compiled verbatim from the reference (oracle/_ref).

Each case cites the reference test it was taken from (paths relative to the vq repo)."""
import math
import os

import numpy as np
import pytest

from oracle import oracle as O

F = np.float32
MINPOS = np.finfo(np.float32).tiny  # f32::MIN_POSITIVE


# ---------------------------------------------------------------- distances
def test_distance_kats_rust(oracle):
    # src/core/distance.rs:131-166
    a, b = [1, 2, 3], [4, 6, 8]
    for sem in ("scalar", "avx512", "avx2"):
        assert oracle.distance("squared_euclidean", a, b, sem) == 50.0
        assert oracle.distance("euclidean", a, b, sem) == F(math.sqrt(50.0))
        assert oracle.distance("manhattan", a, b, sem) == 12.0
        assert abs(oracle.distance("cosine", [1, 0, 0], [0, 1, 0], sem) - 1.0) < 1e-6
        assert abs(oracle.distance("cosine", [1, 2, 3], [1, 2, 3], sem)) < 1e-6
    # src/core/vector.rs:distance2 test, src/core/hsdlib_ffi.rs:169-206
    assert oracle.distance2([1, 2, 3], [4, 5, 6]) == 27.0
    assert oracle.distance("manhattan", [1, 2, 3], [2, 3, 4], "avx512") == 3.0


def test_distance_kats_pyvq(oracle):
    # pyvq/tests/test_distance.py:30-59
    a, b = [1.0, 2.0], [3.0, 4.0]
    for sem in ("scalar", "avx512", "avx2"):
        assert np.isclose(oracle.distance("euclidean", a, b, sem), 2.8284, rtol=1e-4)
        assert np.isclose(oracle.distance("squared_euclidean", a, b, sem), 8.0, rtol=1e-4)
        assert np.isclose(oracle.distance("cosine", a, b, sem), 0.01613, rtol=1e-3)
        assert np.isclose(oracle.distance("manhattan", a, b, sem), 4.0, rtol=1e-4)


def test_cosine_zero_rules(oracle):
    z, v = [0.0, 0.0, 0.0], [1.0, 2.0, 3.0]
    # tests/regression_tests.rs:241-262: zero or near-zero norm -> exactly 1.0 (both builds)
    for sem in ("scalar", "avx512", "avx2"):
        assert oracle.distance("cosine", z, v, sem) == 1.0
        assert oracle.distance("cosine", [1e-20] * 3, v, sem) == 1.0
    # hsdlib tests/test_cosine.c: zero-vs-zero similarity 1 -> distance 0 (simd build only)
    assert oracle.distance("cosine", z, z, "avx512") == 0.0
    assert oracle.distance("cosine", z, z, "scalar") == 1.0  # distance.rs:112-114
    # anti-parallel: simd range is [0,2] (distance.rs:100-104), non-simd clamps to [0,1] (:118)
    assert oracle.distance("cosine", [1, 2, 3], [-1, -2, -3], "avx512") == 2.0
    assert oracle.distance("cosine", [1, 2, 3], [-1, -2, -3], "scalar") == 1.0


HSD_SQ = [  # external/hsdlib/tests/test_euclidean.c:14-75
    ([1, 2, 3, 4, 5, 6, 7, 8, 9], [9, 8, 7, 6, 5, 4, 3, 2, 1], 240.0),
    ([1.1, -2.2, 3.3, -4.4], [1.1, -2.2, 3.3, -4.4], 0.0),
    ([0, 0, 0], [3, 4, 0], 25.0), ([-1, -2], [-4, -6], 25.0), ([], [], 0.0),
    ([5.5], [-2.0], 56.25), ([1, 2, 3], [4, 5, 6], 27.0),
] + [([1] * n, [2] * n, float(n)) for n in (7, 8, 9, 15, 16, 17)]
HSD_L1 = [  # external/hsdlib/tests/test_manhattan.c
    ([1, 2, 3, 4, 5, 6, 7, 8, 9], [9, 8, 7, 6, 5, 4, 3, 2, 1], 40.0),
    ([0, 0, 0], [3, 4, 0], 7.0), ([-1, -2], [-4, -6], 7.0), ([], [], 0.0), ([5.5], [-2.0], 7.5),
    ([1, 2, 3], [4, 5, 6], 9.0),
] + [([1] * n, [2] * n, float(n)) for n in (7, 8, 9, 15, 16, 17)]
HSD_COS = [  # external/hsdlib/tests/test_cosine.c:7-75 (similarity)
    ([1.1, -2.2, 3.3, -4.4], [1.1, -2.2, 3.3, -4.4], 1.0, 1.5e-7),
    ([1, 2, 3], [2, 4, 6], 1.0, 1e-7), ([1, 2, 3], [-1, -2, -3], -1.0, 1e-7),
    ([1, 2, 3], [-2, -4, -6], -1.0, 1e-7), ([1, 0, 0], [0, 1, 0], 0.0, 1e-7),
    ([1, 1], [-1, 1], 0.0, 1e-7), ([0, 0, 0], [3, 4, 0], 0.0, 1e-7), ([3, 4, 0], [0, 0, 0], 0.0, 1e-7),
    ([0, 0, 0], [0, 0, 0], 1.0, 1e-7), ([], [], 1.0, 0), ([5.5], [-2.0], -1.0, 1e-7), ([5.5], [2.0], 1.0, 1e-7),
]


@pytest.mark.parametrize("sem", ["scalar", "avx512", "avx2"])
def test_hsdlib_kats_restated(oracle, sem):
    for a, b, want in HSD_SQ:
        st, got = oracle.hsd_restated("sqeuclidean", a, b, sem)
        assert st == 0 and abs(got - want) <= 1e-5
    for a, b, want in HSD_L1:
        st, got = oracle.hsd_restated("manhattan", a, b, sem)
        assert st == 0 and abs(got - want) <= 1e-6
    for a, b, want, tol in HSD_COS:
        st, got = oracle.hsd_restated("cosine", a, b, sem)
        assert st == 0 and abs(got - want) <= tol
    # test_euclidean.c:77-92: overflow -> HSD_ERR_INVALID_INPUT (-3)
    big = np.finfo(np.float32).max / 1.5
    assert oracle.hsd_restated("sqeuclidean", [big, 0], [-big, 0], sem)[0] == -3
    # NaN / Inf in the scalar tail -> -3
    assert oracle.hsd_restated("sqeuclidean", [np.nan, 1], [0, 1], sem)[0] == -3
    assert oracle.hsd_restated("cosine", [np.inf, 1], [0, 1], sem)[0] == -3


needs_ref = pytest.mark.skipif(not os.path.exists(O.HSD_PATH), reason="oracle/_ref not built")


@needs_ref
def test_real_hsdlib_kats():
    h = O.Hsdlib("auto")
    assert h.backend()
    for a, b, want in HSD_SQ:
        st, got = h.sqeuclidean(a, b)
        assert st == 0 and abs(got - want) <= 1e-5
    for a, b, want in HSD_L1:
        st, got = h.manhattan(a, b)
        assert st == 0 and abs(got - want) <= 1e-6
    for a, b, want, tol in HSD_COS:
        st, got = h.cosine(a, b)
        assert st == 0 and abs(got - want) <= tol


@needs_ref
@pytest.mark.parametrize("backend,sem", [("scalar", "scalar"), ("avx2", "avx2"), ("avx512f", "avx512")])
def test_restatement_bit_exact_vs_real_hsdlib(oracle, backend, sem):
    """The restated kernels must reproduce the verbatim-compiled hsdlib bit for bit, for each
    dispatch target this host can run (lane-boundary sizes of hsdlib's own tests + PQ/TSVQ sizes)."""
    h = O.Hsdlib(backend)
    if backend == "avx512f" and not h.has_avx512():
        pytest.skip("host has no AVX-512F")
    rng = np.random.default_rng(7)
    for n in (1, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 48, 100, 128, 1536):
        for _ in range(20):
            a = (rng.standard_normal(n) * rng.choice([1e-3, 1, 1e3])).astype(F)
            b = (rng.standard_normal(n)).astype(F)
            for which, fn in (("sqeuclidean", h.sqeuclidean), ("manhattan", h.manhattan), ("cosine", h.cosine)):
                st_r, v_r = fn(a, b)
                st_o, v_o = oracle.hsd_restated(which, a, b, sem)
                assert st_r == st_o
                assert np.float32(v_r).tobytes() == np.float32(v_o).tobytes(), (which, n, v_r, v_o)


@needs_ref
def test_distance_through_real_hsdlib_matches_restated(oracle):
    h = O.Hsdlib("auto")
    sem = "avx512" if h.has_avx512() else "avx2"
    rng = np.random.default_rng(3)
    for n in (8, 16, 100):
        a = rng.standard_normal(n).astype(F); b = rng.standard_normal(n).astype(F)
        for metric in O.METRICS:
            assert oracle.distance(metric, a, b, "hsdlib") == oracle.distance(metric, a, b, sem)


def test_simd_consistency(oracle):
    # src/core/distance.rs:176-223: scalar vs simd within 1e-4 (relative here) on len-100 vectors
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, 100).astype(F); b = rng.uniform(-1, 1, 100).astype(F)
    for metric in O.METRICS:
        s = oracle.distance(metric, a, b, "scalar")
        for sem in ("avx512", "avx2"):
            assert abs(s - oracle.distance(metric, a, b, sem)) <= 1e-4 * max(1.0, abs(s))


# ---------------------------------------------------------------- f16
def test_f16_matches_numpy(oracle):
    rng = np.random.default_rng(1)
    xs = np.concatenate([
        rng.standard_normal(4000).astype(F) * 10.0 ** rng.integers(-9, 6, 4000),
        np.array([0.0, -0.0, np.inf, -np.inf, 65504.0, 65520.0, 65519.99, 6e-8, 5.96e-8, 2.98e-8, 2.99e-8,
                  6.1e-5, 6.103515625e-05, 1.0, 1.00048828125, 1.000488281, 1.0014648], F)])
    with np.errstate(over="ignore"):
        want = xs.astype(np.float16).view(np.uint16)
    got = np.array([oracle.f32_to_f16_bits(float(v)) for v in xs], np.uint16)
    assert np.array_equal(got, want)
    # all 65536 half patterns convert back exactly
    allh = np.arange(65536, dtype=np.uint16)
    back = oracle.dequantize_f16(allh)
    ref = allh.view(np.float16).astype(F)
    assert np.array_equal(back.view(np.uint32)[~np.isnan(ref)], ref.view(np.uint32)[~np.isnan(ref)])
    assert np.all(np.isnan(back[np.isnan(ref)]))


# ---------------------------------------------------------------- BQ / SQ
def test_bq_kats(oracle):
    # src/bq.rs:126-144
    assert list(oracle.bq_quantize([-1.0, 0.0, 1.0, -0.5, 0.5], 0.0, 0, 1)) == [0, 1, 1, 0, 1]
    # tests/integration_tests.rs:284-294, :702-711
    assert list(oracle.bq_quantize([0.0, -0.0, MINPOS, -MINPOS], 0.0, 0, 1)) == [1, 1, 1, 0]
    # :477-501
    assert list(oracle.bq_quantize([np.nan, 1.0, -1.0, np.nan], 0.0, 0, 1)) == [0, 1, 0, 0]
    assert list(oracle.bq_quantize([np.inf, -np.inf, 0.0], 0.0, 0, 1)) == [1, 0, 1]
    assert oracle.bq_quantize([], 0.0, 0, 1).size == 0
    # tests/regression_tests.rs:17-32: dequantize maps to low/high, not 0/1
    assert list(oracle.bq_dequantize([10, 20, 10, 20], 10, 20)) == [10.0, 20.0, 10.0, 20.0]
    # subnormal negative is < 0 (no flush-to-zero)
    assert list(oracle.bq_quantize([-MINPOS / 2, MINPOS / 2], 0.0, 0, 1)) == [0, 1]


def test_sq_kats(oracle):
    # src/sq.rs doc-test :13-20
    assert list(oracle.sq_quantize([0.0, 0.5, 1.0], 0.0, 1.0, 11)) == [0, 5, 10]
    # pyvq/tests/test_sq.py:37-54
    x = [-1.2, -1.0, -0.8, -0.3, 0.0, 0.3, 0.6, 1.0, 1.2]
    assert list(oracle.sq_quantize(x, -1.0, 1.0, 5)) == [0, 0, 0, 1, 2, 3, 3, 4, 4]
    # tests/integration_tests.rs:685-699
    b = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]
    assert list(oracle.sq_quantize(b, 0.0, 1.0, 11)) == list(range(11))
    # :516-527 (+Inf -> 255, -Inf -> 0), :505-513 (NaN -> saturating cast -> 0)
    assert list(oracle.sq_quantize([np.inf, -np.inf, np.nan], -1.0, 1.0, 256)) == [255, 0, 0]
    # :530-547
    r = oracle.sq_quantize([MINPOS / 2, -MINPOS / 2, MINPOS, -MINPOS], -1.0, 1.0, 256)
    assert all(126 <= v <= 129 for v in r)
    # :549-565
    fmax = np.finfo(np.float32).max
    r = oracle.sq_quantize([fmax, MINPOS, -fmax, 0.0], -1e10, 1e10, 256)
    assert r[0] == 255 and r[2] == 0 and 126 <= r[1] <= 129 and 126 <= r[3] <= 129
    # dequantize: min + idx*step with two roundings
    step = F(2.0) / F(255)
    want = (F(-1.0) + np.arange(256, dtype=F) * step).astype(F)
    assert np.array_equal(oracle.sq_dequantize(np.arange(256), -1.0, 1.0, 256), want)


def test_sq_matches_numpy_model(oracle):
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(20000) * 0.7).astype(F)
    mn, mx, lv = F(-1.0), F(1.0), 256
    step = (mx - mn) / F(lv - 1)
    c = np.clip(x, mn, mx)
    q = (c - mn) / step
    r = np.where(q >= 0, np.floor(q + F(0.5)), q)  # q >= 0 always after clamp
    # floor(q+0.5) in f32 can differ from roundf for q = k+0.5-ulp; use exact model instead
    r = np.trunc(q) + (q - np.trunc(q) >= F(0.5))
    want = np.minimum(r, lv - 1).astype(np.uint8)
    assert np.array_equal(oracle.sq_quantize(x, mn, mx, lv), want)


# ---------------------------------------------------------------- mean / LBG / PQ
def test_lbg_step_small(oracle):
    # src/core/vector.rs:527-538 (mean [4,5,6]) via a k=1 step
    x = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], F)
    cent, assign, changed, empt = oracle.lbg_step(x, 0, 3, x[:1])
    assert np.array_equal(cent, [[4, 5, 6]]) and list(assign) == [0, 0, 0] and changed and empt.size == 0
    # tie -> lowest index (vector.rs:357 strict '<'); empty cluster reported, not reseeded
    cent, assign, changed, empt = oracle.lbg_step(x, 0, 3, np.array([[4, 5, 6], [4, 5, 6]], F))
    assert list(assign) == [0, 0, 0] and list(empt) == [1] and not changed


def test_lbg_convergence_regression(oracle):
    # tests/regression_tests.rs:208-225
    x = np.array([[1, 1], [1.0001, 1.0001], [10, 10], [10.0001, 10.0001]], F)
    cb, iters = oracle.pq_train(x, 1, 2, 100, [0, 2])
    assert iters[0] < 100
    assert np.allclose(np.sort(cb[0][:, 0]), [1.00005, 10.00005], atol=1e-4)
    # max_iters == 0 returns the sampled rows (vector.rs:415)
    cb0, it0 = oracle.pq_train(x, 1, 2, 0, [3, 1])
    assert it0[0] == 0 and np.array_equal(cb0[0], x[[3, 1]])


def test_pq_matches_numpy_kmeans(oracle):
    """Independent numpy restatement of vector.rs:415-457 (sequential f32 arithmetic)."""
    rng = np.random.default_rng(11)
    n, dim, m, k, iters = 400, 8, 2, 8, 6
    x = rng.standard_normal((n, dim)).astype(F)
    init = np.stack([rng.choice(n, k, replace=False) for _ in range(m)]).astype(np.uint64)
    cb, it = oracle.pq_train(x, m, k, iters, init, reseed=lambda s: 0)
    d = dim // m
    for s in range(m):
        sub = x[:, s * d:(s + 1) * d]
        c = sub[init[s]].copy()
        ran = 0
        for _ in range(iters):
            dist = np.zeros((n, k), F)
            for t in range(d):
                df = (sub[:, t:t + 1] - c[None, :, t]).astype(F)
                dist = (dist + df * df).astype(F)
            a = dist.argmin(1)  # first minimum == strict '<'
            changed = False
            for j in range(k):
                idx = np.nonzero(a == j)[0]
                if idx.size:
                    acc = np.zeros(d, F)
                    for i in idx:
                        acc = (acc + sub[i]).astype(F)
                    new = (acc / F(idx.size)).astype(F)
                    if not np.all(np.abs(new - c[j]) < 1e-6):
                        changed = True
                    c[j] = new
                else:
                    c[j] = sub[0]
            ran += 1
            if not changed:
                break
        assert ran == it[s]
        assert np.array_equal(c, cb[s])


def test_pq_encode_semantics(oracle):
    rng = np.random.default_rng(2)
    cb = rng.standard_normal((3, 16, 4)).astype(F)
    x = rng.standard_normal((50, 12)).astype(F)
    for metric in O.METRICS:
        codes, recon = oracle.pq_encode(cb, metric, x, sem="avx512")
        assert codes.shape == (50, 3) and recon.shape == (50, 12) and recon.dtype == np.float16
        # recon holds the chosen centroids rounded to f16 (pq.rs:193-195)
        want = np.concatenate([cb[s][codes[:, s]] for s in range(3)], axis=1).astype(np.float16)
        assert np.array_equal(recon.view(np.uint16), want.view(np.uint16))
    # duplicate centroids: first index wins
    cb2 = np.repeat(cb[:, :1], 4, axis=1)
    codes, _ = oracle.pq_encode(cb2, "euclidean", x)
    assert not codes.any()


# ---------------------------------------------------------------- TSVQ
def test_tsvq_identical_vectors(oracle):
    # src/tsvq.rs:273-285: ten identical vectors reconstruct within 1e-2; no split possible
    x = np.tile(np.array([[1, 2, 3, 4]], F), (10, 1))
    tree = oracle.tsvq_build(x, 3)
    assert len(tree["left"]) == 1 and tree["left"][0] == -1 and tree["right"][0] == -1
    leaf, recon = oracle.tsvq_encode(tree, "euclidean", x[:1])
    assert np.allclose(recon.astype(F), x[:1], atol=1e-2)


def test_tsvq_two_clusters(oracle):
    # pyvq/tests/test_tsvq.py:68-86
    rng = np.random.default_rng(42)
    c1 = (rng.standard_normal((50, 4)) * 0.1).astype(F)
    c2 = (rng.standard_normal((50, 4)) * 0.1 + 10.0).astype(F)
    x = np.vstack([c1, c2])
    tree = oracle.tsvq_build(x, 2)
    _, r1 = oracle.tsvq_encode(tree, "euclidean", c1[:1])
    _, r2 = oracle.tsvq_encode(tree, "euclidean", c2[:1])
    assert np.linalg.norm(r1.astype(F) - r2.astype(F)) > 5.0


def test_tsvq_structure_small(oracle):
    # hand-checked: values 0..7 in dim 1 dominate the variance; median (3+4)/2 = 3.5 (tsvq.rs:77-81)
    x = np.zeros((8, 2), F); x[:, 1] = np.arange(8); x[:, 0] = [0, 1, 0, 1, 0, 1, 0, 1]
    tree = oracle.tsvq_build(x, 1)
    assert tree["split_dim"][0] == 1 and tree["median"][0] == 3.5
    assert list(tree["count"]) == [8, 4, 4]
    assert np.array_equal(tree["centroids"][1], [0.5, 1.5]) and np.array_equal(tree["centroids"][2], [0.5, 5.5])
    # variance tie -> LAST maximal dim (Iterator::max_by, tsvq.rs:59-66)
    y = np.array([[0, 0], [1, 1]], F)
    assert oracle.tsvq_build(y, 1)["split_dim"][0] == 1
    # NaN row (tests/regression_tests.rs:282-297): builds, NaN goes right
    z = np.array([[1, 2, 3, 4], [5, np.nan, 7, 8], [9, 10, 11, 12]], F)
    t = oracle.tsvq_build(z, 2)
    assert t["count"][0] == 3
    # depth 0 / single vector -> one leaf holding the mean
    assert len(oracle.tsvq_build(x, 0)["left"]) == 1
    assert len(oracle.tsvq_build(x[:1], 5)["left"]) == 1


def test_recon_mse(oracle):
    x = np.array([1.0, 2.0, 3.0, 4.0], F)
    r = np.array([1.0, 2.0, 3.0, 6.0], np.float16)
    assert oracle.recon_mse(x, r) == 1.0


def test_chebyshev_extension_definition(oracle):
    """EXTENSION (VQB_CHEBYSHEV): the reference has no such metric (src/core/distance.rs:8-17), so there is no reference
    value to pin against -- these hand-computed cases pin the restated definition max_i |a_i - b_i| (NaN differences skipped)
    that the GPU path is compared with."""
    f = lambda a, b: oracle.distance("chebyshev", np.asarray(a, np.float32), np.asarray(b, np.float32))
    assert f([1, 2, 3], [4, 0, 3.5]) == 3.0
    assert f([1, 2, 3], [1, 2, 3]) == 0.0
    assert f([], []) == 0.0
    assert f([-5.5], [2.25]) == 7.75
    assert f([np.nan, 1.0], [0.0, 3.0]) == 2.0            # NaN difference skipped
    assert f([np.inf, 1.0], [0.0, 3.0]) == np.inf
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(257).astype(np.float32), rng.standard_normal(257).astype(np.float32)
    assert f(a, b) == np.abs(a - b).max()
