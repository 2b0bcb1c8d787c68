"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded inputs.  Bars (BASELINE.json north_star): BQ/SQ bit-exact; PQ/TSVQ assignments >= 99.9 %
identical with every disagreement a near-tie (<= 1e-5 relative distance); codebooks within 1e-4
relative.  The exact CUDA-core kernels and the ordered update are held to the stricter bar of
bit-equality with the oracle (same operation order)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
F = np.float32
MINPOS = np.finfo(np.float32).tiny
METRICS = ["squared_euclidean", "euclidean", "manhattan", "cosine"]


@pytest.fixture(scope="module")
def vq():
    import vq_b200
    return vq_b200


@pytest.fixture(scope="module")
def eng(vq):
    return vq.default_engine()


def mixture(n, dim, seed, comps=64, sigma=0.25):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((comps, dim)).astype(F)
    x = centers[rng.integers(0, comps, n)] + sigma * rng.standard_normal((n, dim)).astype(F)
    return np.ascontiguousarray(x, dtype=F)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({2: np.uint16, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


# ------------------------------------------------------------------ BQ / SQ / f16
SPECIAL = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, MINPOS, -MINPOS, MINPOS / 2, -MINPOS / 2,
                    np.finfo(F).max, -np.finfo(F).max, 0.5, -0.5, 1.0, -1.0, 0.49999997, 0.50000006], F)


@pytest.mark.parametrize("n", [0, 1, 15, 4095, 4096, 4097, 1_000_003])
def test_bq_sq_bit_exact(vq, oracle, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 0.5).astype(F)
    k = min(n, SPECIAL.size)
    x[:k] = SPECIAL[:k]
    bq = vq.BinaryQuantizer(0.0, 0, 1)
    assert np.array_equal(bq.quantize(x), oracle.bq_quantize(x, 0.0, 0, 1))
    bq2 = vq.BinaryQuantizer(0.25, 10, 20)
    c = bq2.quantize(x)
    assert np.array_equal(c, oracle.bq_quantize(x, 0.25, 10, 20))
    assert np.array_equal(bits(bq2.dequantize(c)), bits(oracle.bq_dequantize(c, 10, 20)))
    for mn, mx, lv in [(-1.0, 1.0, 256), (0.0, 1.0, 11), (-1e10, 1e10, 256), (-1.0, 1.0, 5), (-3.0, 7.0, 2)]:
        sq = vq.ScalarQuantizer(mn, mx, lv)
        assert sq.step == oracle.sq_step(mn, mx, lv)
        q = sq.quantize(x)
        assert np.array_equal(q, oracle.sq_quantize(x, mn, mx, lv)), (mn, mx, lv)
        assert np.array_equal(bits(sq.dequantize(q)), bits(oracle.sq_dequantize(q, mn, mx, lv)))


def test_sq_division_on_rounding_boundaries(vq, oracle):
    """The SQ kernel divides by `step` with a reciprocal + two residual corrections instead of the generic
    IEEE sequence; it must still be the correctly rounded quotient (sq.rs:125).  Stress it where a wrong last
    bit changes the code: numerators a few ulps around (j + 0.5) * step, steps with awkward mantissas
    (all ones, just above a power of two), tiny / huge ranges that take the generic path."""
    rng = np.random.default_rng(99)
    cases = [(-1.0, 1.0, 256), (0.0, 1.0, 255), (0.0, 0.99999994, 2), (0.0, 1.0000001, 2), (-0.33333334, 0.6666667, 4),
             (0.0, 3.0, 256), (-7.0, 9.0, 255), (0.0, 1e-30, 17), (-1e25, 1e25, 256), (1.0, 1.0000038, 33),
             (0.0, 16777215.0, 256), (-123.456, 789.012, 200)]
    for mn, mx, lv in cases:
        mn, mx = F(mn), F(mx)
        step = oracle.sq_step(float(mn), float(mx), lv)
        j = rng.integers(0, lv, 400_000).astype(F)
        base = (mn + (j + F(0.5)) * F(step)).astype(F)
        x = base.copy()
        for k in range(1, 5):   # +-1..4 ulps around each half-way point
            x = np.concatenate([x, np.nextafter(x[-base.size:], F(np.inf)), ])
        lo = base.copy()
        for k in range(4):
            lo = np.nextafter(lo, F(-np.inf)); x = np.concatenate([x, lo])
        x = np.concatenate([x, rng.uniform(float(mn), float(mx), 400_000).astype(F)])
        sq = vq.ScalarQuantizer(float(mn), float(mx), lv)
        assert np.array_equal(sq.quantize(x), oracle.sq_quantize(x, float(mn), float(mx), lv)), (mn, mx, lv)


def test_bq_sq_unaligned_and_kats(vq, oracle):
    x = (np.random.default_rng(1).standard_normal(10_001)).astype(F)[1:]  # 4-byte aligned only
    assert np.array_equal(vq.BinaryQuantizer(0.1).quantize(x), oracle.bq_quantize(x, 0.1, 0, 1))
    assert np.array_equal(vq.ScalarQuantizer(-1, 1).quantize(x), oracle.sq_quantize(x, -1.0, 1.0, 256))
    # src/bq.rs:126-144, tests/integration_tests.rs:284-294,477-501
    assert list(vq.BinaryQuantizer(0.0).quantize(np.array([-1.0, 0.0, 1.0, -0.5, 0.5], F))) == [0, 1, 1, 0, 1]
    assert list(vq.BinaryQuantizer(0.0).quantize(np.array([0.0, -0.0, MINPOS, -MINPOS], F))) == [1, 1, 1, 0]
    assert list(vq.BinaryQuantizer(0.0).quantize(np.array([np.nan, 1.0, -1.0, np.nan], F))) == [0, 1, 0, 0]
    # pyvq/tests/test_sq.py:37-54, tests/integration_tests.rs:516-527,685-699
    xs = np.array([-1.2, -1.0, -0.8, -0.3, 0.0, 0.3, 0.6, 1.0, 1.2], F)
    assert list(vq.ScalarQuantizer(-1.0, 1.0, 5).quantize(xs)) == [0, 0, 0, 1, 2, 3, 3, 4, 4]
    assert list(vq.ScalarQuantizer(-1.0, 1.0, 256).quantize(np.array([np.inf, -np.inf, np.nan], F))) == [255, 0, 0]
    b = np.arange(11, dtype=F) / F(10)
    assert list(vq.ScalarQuantizer(0.0, 1.0, 11).quantize(np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0], F))) == list(range(11))


def test_f16_dequantize_all_patterns(vq, eng, oracle):
    from vq_b200.api import _dequantize_f16
    allh = np.arange(65536, dtype=np.uint16).view(np.float16)
    got = _dequantize_f16(eng, allh)
    want = oracle.dequantize_f16(allh)
    ok = ~np.isnan(want)
    assert np.array_equal(bits(got)[ok], bits(want)[ok]) and np.all(np.isnan(got[~ok]))


def test_elementwise_device_pointers(vq, oracle):
    torch = pytest.importorskip("torch")
    x = torch.randn(300_001, device="cuda")
    q = vq.ScalarQuantizer(-1.0, 1.0, 256).quantize(x)
    assert q.is_cuda and np.array_equal(q.cpu().numpy(), oracle.sq_quantize(x.cpu().numpy(), -1.0, 1.0, 256))
    b = vq.BinaryQuantizer(0.0).quantize(x)
    assert np.array_equal(b.cpu().numpy(), oracle.bq_quantize(x.cpu().numpy(), 0.0, 0, 1))


# ------------------------------------------------------------------ Distance
@pytest.mark.parametrize("metric", METRICS)
def test_distance_batch_bit_exact(vq, oracle, metric):
    rng = np.random.default_rng(5)
    d = vq.Distance(metric)
    for n in (1, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 100, 1536):
        a = (rng.standard_normal((64, n)) * rng.choice([1e-3, 1.0, 1e3], (64, 1))).astype(F)
        b = rng.standard_normal((64, n)).astype(F)
        # zero vectors, NaN / Inf in body and tail, identical rows
        a[0] = 0; b[1] = 0; a[2] = 0; b[2] = 0; a[3] = b[3]
        a[4, 0] = np.nan; b[5, -1] = np.inf; a[6, n // 2] = -np.inf
        a[7] = 1e-25; a[8] = 3e19
        got = d.compute_batch(a, b)
        want = np.array([oracle.distance(metric, a[i], b[i], "avx512") for i in range(64)], F)
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan), (metric, n)
        assert np.array_equal(bits(got)[~nan], bits(want)[~nan]), (metric, n)
    # KATs: src/core/distance.rs:131-166, pyvq/tests/test_distance.py:30-59
    assert vq.Distance("squared_euclidean").compute([1, 2, 3], [4, 6, 8]) == 50.0
    assert vq.Distance.manhattan().compute([1, 2, 3], [4, 6, 8]) == 12.0
    assert np.isclose(vq.Distance.cosine().compute([1.0, 2.0], [3.0, 4.0]), 0.01613, rtol=1e-3)
    with pytest.raises(ValueError, match="Dimension mismatch"):
        vq.Distance.euclidean().compute([1.0, 2.0], [3.0, 4.0, 5.0])


# ------------------------------------------------------------------ PQ: assignment / step / train
def gpu_assign_train(eng, x, cb):
    n, dim = x.shape
    m, k, d = cb.shape
    codes = np.empty((m, n), np.uint32)
    eng.check(eng.lib.vqb_pq_assign_train(eng.h, x.ctypes.data, n, dim, m, k, cb.ctypes.data, 1, codes.ctypes.data))
    return codes


@pytest.mark.parametrize("dim,m,k", [(32, 4, 256), (64, 4, 256), (32, 8, 64), (64, 2, 300), (20, 4, 17), (128, 8, 256)])
def test_assign_train_exact(eng, oracle, dim, m, k):
    n = 5000
    x = mixture(n, dim, 3)
    rng = np.random.default_rng(4)
    d = dim // m
    cb = np.stack([x[rng.choice(n, k, replace=False), s * d:(s + 1) * d] for s in range(m)]).astype(F)
    cb[0, 5] = cb[0, 2]  # duplicate centroid: lowest index must win
    codes = gpu_assign_train(eng, x, np.ascontiguousarray(cb))
    for s in range(m):
        _, want, _, _ = oracle.lbg_step(x, s * d, d, cb[s])
        assert np.array_equal(codes[s], want), f"subspace {s}"


def gpu_train_step(eng, x, cb, update="ordered"):
    from vq_b200 import _lib
    n, dim = x.shape
    m, k, d = cb.shape
    out = np.ascontiguousarray(cb).copy()
    changed = np.zeros(m, np.uint32); counts = np.zeros((m, k), np.uint32)
    opts = _lib.TrainOpts(); opts.struct_size = C.sizeof(_lib.TrainOpts)
    opts.update_mode = 0 if update == "ordered" else 1
    eng.check(eng.lib.vqb_pq_train_step(eng.h, x.ctypes.data, n, dim, m, k, out.ctypes.data, C.byref(opts),
                                        changed.ctypes.data, counts.ctypes.data))
    return out, changed, counts


@pytest.mark.parametrize("dim,m,k,n", [(64, 8, 256, 20000), (48, 3, 40, 3000), (64, 2, 300, 4000)])
def test_train_step_teacher_forced(eng, oracle, dim, m, k, n):
    """One iteration from the oracle's state at every iteration t (SURVEY 8d parity protocol i)."""
    x = mixture(n, dim, 11)
    d = dim // m
    rng = np.random.default_rng(12)
    state = np.stack([x[rng.choice(n, k, replace=False), s * d:(s + 1) * d] for s in range(m)]).astype(F)
    for t in range(4):
        got, changed, counts = gpu_train_step(eng, x, state)
        fast, _, _ = gpu_train_step(eng, x, state, update="fast")
        nxt = np.empty_like(state)
        for s in range(m):
            want, assign, ch, empt = oracle.lbg_step(x, s * d, d, state[s])
            nonempty = np.ones(k, bool); nonempty[empt] = False
            assert np.array_equal(counts[s], np.bincount(assign, minlength=k))
            assert bool(changed[s]) == ch
            # ordered update: the reference's summation order -> bit-identical means
            assert np.array_equal(bits(got[s][nonempty]), bits(want[nonempty])), (t, s)
            assert np.array_equal(bits(got[s][~nonempty]), bits(state[s][~nonempty]))  # empties untouched
            rel = np.linalg.norm(fast[s] - want) / np.linalg.norm(want)
            assert rel <= 1e-4
            nxt[s] = want
        state = nxt


@pytest.mark.parametrize("sub_dim", [4, 8, 12, 16, 20, 24, 32, 40, 64])
@pytest.mark.parametrize("k,n", [(256, 9000), (16, 4100), (256, 5000)])
def test_train_step_every_update_kernel_shape(eng, oracle, sub_dim, k, n):
    """One teacher-forced iteration across the shapes that select different update / assignment kernels (tile kernels
    for sub_dim 8 / 16 / 32 and n >= 4096, radix grouping otherwise; tensor assignment for sub_dim 8 / 16 / 24 / 32):
    ORDERED means bit-identical with the oracle, FAST within 1e-4.  (sub_dim 32 with the ordered tile kernel once read a
    swizzled tile as plain rows -- no test had that shape.)"""
    m = 3 if sub_dim <= 32 else 2
    dim = m * sub_dim
    x = mixture(n, dim, 77 + sub_dim)
    rng = np.random.default_rng(13)
    state = np.stack([x[rng.choice(n, k, replace=False), s * sub_dim:(s + 1) * sub_dim] for s in range(m)]).astype(F)
    got, changed, counts = gpu_train_step(eng, x, state)
    fast, _, _ = gpu_train_step(eng, x, state, update="fast")
    for s in range(m):
        want, assign, ch, empt = oracle.lbg_step(x, s * sub_dim, sub_dim, state[s])
        nonempty = np.ones(k, bool); nonempty[empt] = False
        assert np.array_equal(counts[s], np.bincount(assign, minlength=k))
        assert bool(changed[s]) == ch
        assert np.array_equal(bits(got[s][nonempty]), bits(want[nonempty])), s
        assert np.linalg.norm(fast[s] - want) / np.linalg.norm(want) <= 1e-4


def test_pq_train_end_to_end_bit_exact(vq, oracle):
    """C1-shaped (reduced n): 128-d, m 8, k 256, 10 iterations, explicit index stream."""
    n, dim, m, k, iters = 20000, 128, 8, 256, 10
    x = mixture(n, dim, 20240, comps=1024)
    init, _ = vq.draw_init_indices(n, m, k, 42)
    pq = vq.ProductQuantizer(x, m, k, iters, vq.Distance.euclidean(), init_idx=init, reseed=lambda s: 0)
    want, it = oracle.pq_train(x, m, k, iters, init, reseed=lambda s: 0)
    assert np.array_equal(pq.iters_run, it)
    assert np.array_equal(bits(pq.codebooks), bits(want))
    fast = vq.ProductQuantizer(x, m, k, iters, init_idx=init, reseed=lambda s: 0, update="fast")
    for s in range(m):
        assert np.linalg.norm(fast.codebooks[s] - want[s]) / np.linalg.norm(want[s]) <= 1e-4


def test_pq_train_reseed_and_early_exit(vq, oracle):
    """Duplicated rows force empty clusters (reseed path, vector.rs:448-452) and quick convergence
    (per-subspace early exit, vector.rs:455-457); the same seeded stream drives both sides."""
    rng = np.random.default_rng(9)
    base = rng.standard_normal((12, 8)).astype(F)
    x = np.ascontiguousarray(base[rng.integers(0, 12, 600)])
    x[:, 4:] += (rng.standard_normal((600, 4)) * 0.01).astype(F)  # second subspace has distinct rows
    m, k, iters = 2, 16, 25
    init, _ = vq.draw_init_indices(600, m, k, 7)

    def make_reseed():
        from vq_b200.rand09 import IndexStream
        st = [IndexStream(1234, s) for s in range(m)]
        calls = []
        def f(s):
            calls.append(s)
            return st[s].choose(600)
        return f, calls
    f1, calls1 = make_reseed(); f2, calls2 = make_reseed()
    pq = vq.ProductQuantizer(x, m, k, iters, init_idx=init, reseed=f1)
    want, it = oracle.pq_train(x, m, k, iters, init, reseed=f2)
    assert calls1 == calls2 and len(calls1) > 0
    assert np.array_equal(pq.iters_run, it) and it[0] < iters
    assert np.array_equal(bits(pq.codebooks), bits(want))
    # max_iters == 0 returns the sampled rows (vector.rs:415)
    pq0 = vq.ProductQuantizer(x, m, k, 0, init_idx=init)
    assert np.array_equal(pq0.codebooks[1], x[init[1].astype(int), 4:])


def test_chebyshev_extension_bit_exact(vq, oracle):
    """EXTENSION (north_star names a Chebyshev assignment; the reference has no such metric, src/core/distance.rs:8-17):
    parity is against the oracle's restated definition only."""
    rng = np.random.default_rng(9)
    d = vq.Distance.chebyshev()
    with pytest.raises(ValueError):
        vq.Distance("chebyshev")           # the reference's constructor knows four kinds
    for n in (1, 7, 8, 33, 1536):
        a = (rng.standard_normal((64, n)) * rng.choice([1e-3, 1.0, 1e3], (64, 1))).astype(F)
        b = rng.standard_normal((64, n)).astype(F)
        a[0] = 0; b[1] = 0; a[3] = b[3]; a[4, 0] = np.nan; b[5, -1] = np.inf
        got = d.compute_batch(a, b)
        want = np.array([oracle.distance("chebyshev", a[i], b[i]) for i in range(64)], F)
        assert np.array_equal(bits(got), bits(want)), n
    for dim, m, k in ((64, 8, 256), (96, 3, 50), (40, 8, 256), (128, 8, 300)):
        n = 3000
        x = mixture(n, dim, 41)
        sd = dim // m
        cb = np.stack([x[rng.choice(n, k, replace=False), s * sd:(s + 1) * sd] for s in range(m)]).astype(F)
        cb[0, 3] = cb[0, 1]          # duplicate: lowest index wins (pq.rs:183-191)
        x[1, :sd] = np.nan
        pq = vq.ProductQuantizer.from_codebooks(cb, d)
        codes, recon = pq.encode_with_recon(x)
        want_codes, want_recon = oracle.pq_encode(cb, "chebyshev", x)
        assert np.array_equal(codes.astype(np.uint32), want_codes)
        assert np.array_equal(bits(recon), bits(want_recon))


# ------------------------------------------------------------------ PQ: encode
@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dim,m,k", [(64, 8, 256), (128, 8, 256), (96, 3, 50), (40, 8, 256)])
def test_pq_encode_exact(vq, oracle, metric, dim, m, k):
    n = 4000
    x = mixture(n, dim, 31)
    d = dim // m
    rng = np.random.default_rng(32)
    cb = np.stack([x[rng.choice(n, k, replace=False), s * d:(s + 1) * d] for s in range(m)]).astype(F)
    cb[0, 3] = cb[0, 1]          # duplicate
    cb[m - 1, 0] = 0.0           # zero centroid (cosine zero rules, cosine.c:38-45)
    x[0] = 0.0                   # zero vector
    x[1, :d] = np.nan            # NaN sub-vector: index 0 sticks (pq.rs:183-191)
    x[2, 0] = np.inf
    pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric))
    codes, recon = pq.encode_with_recon(x)
    want_codes, want_recon = oracle.pq_encode(cb, metric, x, sem="avx512")
    assert np.array_equal(codes.astype(np.uint32), want_codes)
    assert np.array_equal(bits(recon), bits(want_recon))
    # single-vector reference API
    q = pq.quantize(x[7])
    assert q.dtype == np.float16 and np.array_equal(bits(q), bits(want_recon[7]))
    assert np.array_equal(pq.dequantize(q), oracle.dequantize_f16(want_recon[7]))
    # decode == f16 round trip of the chosen centroids
    assert np.array_equal(bits(pq.decode(codes)), bits(want_recon.astype(F)))


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("sub_dim", [4, 8, 12, 16, 24, 32, 40])
@pytest.mark.parametrize("k,n", [(7, 1030), (256, 1000), (256, 4100), (300, 4100)])
def test_pq_encode_kernel_selection_sweep(vq, oracle, metric, sub_dim, k, n):
    """Encode in AUTO mode across the shapes that select different assignment kernels (tensor kernel for sub_dim 8 / 16 /
    24 / 32, k <= 256, n >= 1024; tiled Manhattan for sub_dim 8, n >= 4096; the CUDA-core kernel otherwise): codes and f16
    reconstructions identical with the oracle, from host buffers and from device buffers."""
    torch = pytest.importorskip("torch")
    m = 3
    dim = m * sub_dim
    x = mixture(n, dim, 500 + sub_dim + k)
    rng = np.random.default_rng(sub_dim * 1000 + k)
    cb = np.stack([x[rng.choice(n, k, replace=False), s * sub_dim:(s + 1) * sub_dim] for s in range(m)]).astype(F)
    cb[0, min(3, k - 1)] = cb[0, 1]
    x[0] = 0.0
    x[1, :sub_dim] = np.nan
    pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric))
    want_codes, want_recon = oracle.pq_encode(cb, metric, x, sem="avx512")
    codes, recon = pq.encode_with_recon(x)
    assert np.array_equal(codes.astype(np.uint32), want_codes)
    assert np.array_equal(bits(recon), bits(want_recon))
    # decode (tiled kernel for u8 codes, sub_dim 4 / 8 / 16 / 32, n >= 4096; gather kernel otherwise) == f16 round trip
    assert np.array_equal(bits(pq.decode(codes)), bits(want_recon.astype(F)))
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(bits(pq.decode(torch.from_numpy(codes).cuda()).cpu().numpy()), bits(want_recon.astype(F)))
    codes_d, recon_d = pq.encode_with_recon(xd)
    assert np.array_equal(codes_d.cpu().numpy().astype(np.int64) & 0xFFFFFFFF, want_codes.astype(np.int64))
    assert np.array_equal(recon_d.cpu().numpy().view(np.uint16), want_recon.view(np.uint16))


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("sub_dim,k", [(4, 256), (8, 256), (13, 50), (24, 300), (1, 7)])
def test_pq_encode_a_handful_of_rows(vq, oracle, metric, sub_dim, k):
    """The reference's call shape is one vector per call (src/pq.rs:167): up to 64 rows take the warp-per-(row, subspace)
    kernel.  Codes and reconstructions identical with the oracle for n = 1 ... 65, with NaN / zero / duplicate cases."""
    m = 3
    dim = m * sub_dim
    base = mixture(400, dim, 900 + sub_dim)
    rng = np.random.default_rng(901 + sub_dim)
    cb = np.stack([base[rng.choice(400, k, replace=k > 400), s * sub_dim:(s + 1) * sub_dim] for s in range(m)]).astype(F)
    cb[0, min(3, k - 1)] = cb[0, 1]          # duplicate: lowest index wins
    cb[m - 1, 0] = 0.0
    pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric))
    for n in (1, 2, 5, 33, 64, 65):
        x = mixture(n, dim, 950 + n)
        x[0, :sub_dim] = np.nan if n > 1 else x[0, :sub_dim]
        if n > 2: x[2] = 0.0
        if n > 4: x[4, :sub_dim] = cb[0, 1]
        codes, recon = pq.encode_with_recon(x)
        want_codes, want_recon = oracle.pq_encode(cb, metric, x, sem="avx512")
        assert np.array_equal(codes.astype(np.uint32), want_codes), n
        assert np.array_equal(bits(recon), bits(want_recon)), n
    q = pq.quantize(base[7])
    assert np.array_equal(bits(q), bits(oracle.pq_encode(cb, metric, base[7:8], sem="avx512")[1][0]))


def test_pq_encode_vs_real_hsdlib_near_tie_rule(vq, oracle):
    """Against hsdlib compiled verbatim from the reference, whatever this host dispatches to:
    >= 99.9 % agreement and every disagreement within 1e-5 relative distance."""
    if oracle.hsd is None:
        pytest.skip("oracle/_ref not available")
    n, dim, m, k = 20000, 64, 8, 256
    x = mixture(n, dim, 41)
    d = dim // m
    rng = np.random.default_rng(42)
    cb = np.stack([x[rng.choice(n, k, replace=False), s * d:(s + 1) * d] for s in range(m)]).astype(F)
    for metric in METRICS:
        got = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric)).encode(x).astype(np.uint32)
        want, _ = oracle.pq_encode(cb, metric, x, sem="hsdlib", want_recon=False)
        miss = np.argwhere(got != want)
        assert 1.0 - len(miss) / got.size >= 0.999
        for i, s in miss:
            v = x[i, s * d:(s + 1) * d]
            dg = oracle.distance(metric, v, cb[s, got[i, s]], "hsdlib")
            dw = oracle.distance(metric, v, cb[s, want[i, s]], "hsdlib")
            assert abs(dg - dw) <= 1e-5 * max(abs(dw), 1e-30)


def test_pq_encode_round_trip_property(vq):
    """Size-independent property: vectors assembled from centroids encode to their own codes."""
    rng = np.random.default_rng(50)
    m, k, d, n = 96, 256, 8, 200_000
    cb = rng.standard_normal((m, k, d)).astype(F)
    codes = rng.integers(0, k, (n, m)).astype(np.uint8)
    x = np.ascontiguousarray(cb[np.arange(m)[None, :], codes].reshape(n, m * d))
    for metric in ("squared_euclidean", "euclidean", "manhattan"):
        pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric))
        assert np.array_equal(pq.encode(x), codes)


def test_pq_device_pointer_and_chunked_host_paths_agree(vq):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(60)
    m, k, d, n = 96, 256, 8, 150_000  # host path: 128 MiB chunks -> several chunks at 768 dims
    cb = rng.standard_normal((m, k, d)).astype(F)
    x = rng.standard_normal((n, m * d)).astype(F)
    pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance.euclidean())
    c_host, r_host = pq.encode_with_recon(x)
    c_dev, r_dev = pq.encode_with_recon(torch.from_numpy(x).cuda())
    assert np.array_equal(c_host, c_dev.cpu().numpy())
    assert np.array_equal(bits(r_host), bits(r_dev.cpu().numpy()))


# ------------------------------------------------------------------ TSVQ
@pytest.mark.parametrize("n,dim,depth", [(3000, 32, 5), (1000, 33, 4), (500, 1536, 3), (257, 8, 8), (64, 4, 10),
                                         (6000, 256, 8), (2500, 1536, 5)])  # all three ring configurations of k_colsum_w
def test_tsvq_build_bit_exact(vq, oracle, n, dim, depth):
    x = mixture(n, dim, 70 + dim, comps=8, sigma=0.5)
    if dim == 8:
        x[:, 3] = np.round(x[:, 3])  # many ties at the median: extra points go left (tsvq.rs:84)
    want = oracle.tsvq_build(x, depth)
    t = vq.TSVQ(x, depth, vq.Distance.euclidean()).tree()
    assert np.array_equal(t["left"], want["left"]) and np.array_equal(t["right"], want["right"])
    assert np.array_equal(t["count"], want["count"])
    assert np.array_equal(t["split_dim"], want["split_dim"])
    ok = ~np.isnan(want["median"])
    assert np.array_equal(np.isnan(t["median"]), ~ok) and np.array_equal(bits(t["median"])[ok], bits(want["median"])[ok])
    assert np.array_equal(bits(t["centroids"]), bits(want["centroids"]))


def test_tsvq_reference_cases(vq, oracle):
    # src/tsvq.rs:273-285: ten identical vectors -> single leaf, reconstruction within 1e-2
    x = np.tile(np.array([[1, 2, 3, 4]], F), (10, 1))
    t = vq.TSVQ(x, 3)
    assert len(t.tree()["left"]) == 1
    assert np.allclose(t.quantize(x[0]).astype(F), x[0], atol=1e-2)
    # pyvq/tests/test_tsvq.py:68-86
    rng = np.random.default_rng(42)
    c1 = (rng.standard_normal((50, 4)) * 0.1).astype(F); c2 = (rng.standard_normal((50, 4)) * 0.1 + 10).astype(F)
    t = vq.TSVQ(np.vstack([c1, c2]), 2)
    assert np.linalg.norm(t.quantize(c1[0]).astype(F) - t.quantize(c2[0]).astype(F)) > 5.0
    # tests/regression_tests.rs:282-297: NaN in the data must not crash; same tree as the oracle
    z = np.array([[1, 2, 3, 4], [5, np.nan, 7, 8], [9, 10, 11, 12]], F)
    tz = vq.TSVQ(z, 2).tree(); wz = oracle.tsvq_build(z, 2)
    assert np.array_equal(tz["left"], wz["left"]) and np.array_equal(tz["count"], wz["count"])
    # depth 0 and single vector
    assert len(vq.TSVQ(x, 0).tree()["left"]) == 1 and len(vq.TSVQ(x[:1], 5).tree()["left"]) == 1


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dim", [8, 33, 100, 128, 1536])
def test_tsvq_encode_exact(vq, oracle, metric, dim):
    n = 2000 if dim < 1000 else 600
    x = mixture(n, dim, 80, comps=16, sigma=0.5)
    tree = oracle.tsvq_build(x, 6)
    q = mixture(1500, dim, 81, comps=16, sigma=0.5)
    q[0] = 0.0; q[1, dim - 1] = np.nan; q[2, 0] = np.inf
    t = vq.TSVQ.from_tree(tree["centroids"], tree["left"], tree["right"], vq.Distance(metric))
    leaf = t.encode(q); recon = t.quantize_batch(q)
    want_leaf, want_recon = oracle.tsvq_encode(tree, metric, q, sem="avx512")
    assert np.array_equal(leaf.astype(np.uint32), want_leaf)
    assert np.array_equal(bits(recon), bits(want_recon))
    assert np.array_equal(bits(t.quantize(q[5])), bits(want_recon[5]))


# ------------------------------------------------------------------ reference-API behaviour (pyvq tests)
def test_api_shapes_and_errors(vq):
    rng = np.random.default_rng(0)
    training = rng.random((100, 16)).astype(F)
    pq = vq.ProductQuantizer(training_data=training, num_subspaces=4, num_centroids=8, max_iters=10, seed=42)
    assert (pq.dim, pq.num_subspaces, pq.sub_dim) == (16, 4, 4)  # pyvq/tests/test_pq.py:6-18
    codes = pq.quantize(training[0].copy())
    assert isinstance(codes, np.ndarray) and codes.dtype == np.float16 and len(codes) == 16
    rec = pq.dequantize(codes)
    assert rec.dtype == np.float32 and len(rec) == 16
    assert "ProductQuantizer" in repr(pq) and "dim=16" in repr(pq)
    with pytest.raises(ValueError, match="Dimension mismatch"):
        pq.quantize(rng.random(10).astype(F))
    # same input, same output (tests/integration_tests.rs:40-53)
    pq2 = vq.ProductQuantizer(training, 4, 8, 10, seed=42)
    assert np.array_equal(bits(pq.codebooks), bits(pq2.codebooks))
    # pyvq/tests/test_integrations.py:42-76: RMSE < 2.0 on randn
    data = rng.standard_normal((200, 16)).astype(F)
    pq3 = vq.ProductQuantizer(data, 4, 16, 10)
    r = pq3.quantize_batch(data).astype(F)
    assert np.sqrt(np.mean((data - r) ** 2)) < 2.0
    ts = vq.TSVQ(data, 4)
    r = ts.quantize_batch(data).astype(F)
    assert np.sqrt(np.mean((data - r) ** 2)) < 2.0
    assert ts.dim == 16 and "TSVQ" in repr(ts)


@pytest.mark.parametrize("n,dim,m,k", [(5000, 64, 8, 256), (4097, 40, 5, 200), (9000, 96, 12, 17), (6000, 8, 1, 256)])
def test_manhattan_tiled_kernel_equals_exact_and_oracle(vq, oracle, n, dim, m, k):
    """sub_dim 8, k <= 256, n >= 4096: Manhattan encode runs the tiled kernel (k_assign_l1_tiles).  Codes and f16
    reconstructions must equal the generic exact kernel and the oracle bit for bit, with a ragged last tile, a last column
    group of fewer than four subspaces, NaN / Inf / zero rows and duplicate centroids (lowest index wins)."""
    x = mixture(n, dim, 5)
    d = dim // m
    rng = np.random.default_rng(6)
    cb = np.stack([x[rng.choice(n, k, replace=False), s * d:(s + 1) * d] for s in range(m)]).astype(F)
    cb[0, min(3, k - 1)] = cb[0, 1]
    x[0] = 0.0
    x[1, :d] = np.nan
    x[2, 0] = np.inf
    x[n - 1] = -x[n - 2]
    pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance.manhattan())
    c_t, r_t = pq.encode_with_recon(x)                       # auto -> tiled kernel
    c_e, r_e = pq.encode_with_recon(x, assign="exact")       # generic CUDA-core kernel
    assert np.array_equal(c_t, c_e)
    assert np.array_equal(bits(r_t), bits(r_e))
    want_codes, want_recon = oracle.pq_encode(cb, "manhattan", x, sem="avx512")
    assert np.array_equal(c_t.astype(np.uint32), want_codes)
    assert np.array_equal(bits(r_t), bits(want_recon))
