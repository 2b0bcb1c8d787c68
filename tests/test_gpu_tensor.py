"""GPU tests of the tcgen05 GEMM-form assignment kernel (vq_b200/csrc/pq_tc.cu).

The kernel decides every code with the reference's arithmetic (the tensor core only prunes
candidates), so it is held to the strict bar: codes identical to the CUDA-core exact kernel and
to the CPU oracle, including ties, NaN/Inf rows, zero vectors and padded codebooks.  The raw
tensor-core scores are also dumped and compared with float64 to back the margin constant."""
import ctypes as C

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
F = np.float32
KAPPA = 2.0 ** -17  # pq_tc.cu


@pytest.fixture(scope="module")
def vq():
    import vq_b200
    return vq_b200


@pytest.fixture(scope="module")
def eng(vq):
    return vq.default_engine()


def mixture(n, dim, seed, comps=1024, sigma=0.25, scale=1.0):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((comps, dim)).astype(F)
    x = centers[rng.integers(0, comps, n)] + sigma * rng.standard_normal((n, dim)).astype(F)
    return np.ascontiguousarray(x * F(scale), dtype=F)


def sample_codebooks(x, m, k, seed):
    n, dim = x.shape
    d = dim // m
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(
        np.stack([x[rng.choice(n, k, replace=False), s * d:(s + 1) * d] for s in range(m)]).astype(F))


def assign_train(eng, x, cb, mode):
    n, dim = x.shape
    m, k, _ = cb.shape
    codes = np.empty((m, n), np.uint32)
    eng.check(eng.lib.vqb_pq_assign_train(eng.h, x.ctypes.data, n, dim, m, k, cb.ctypes.data, mode, codes.ctypes.data))
    return codes


def debug_scores(eng, x, cb, sub, cosine):
    n, dim = x.shape
    m, k, _ = cb.shape
    scores = np.empty((n, 256), F)
    rescans = np.zeros(1, np.uint64)
    codes = np.empty((m, n), np.uint32)
    eng.check(eng.lib.vqb_debug_tc_scores(eng.h, int(cosine), x.ctypes.data, n, dim, m, k, cb.ctypes.data, sub,
                                          scores.ctypes.data, rescans.ctypes.data, codes.ctypes.data))
    return scores, int(rescans[0]), codes


@pytest.mark.parametrize("sub_dim", [8, 16, 24, 32])
@pytest.mark.parametrize("scale", [1.0, 1e-3, 1e4])
def test_tensor_scores_within_margin(eng, scale, sub_dim):
    """|tcgen05 score - float64 score| must stay far inside the margin M = KAPPA * S used to prune.

    Budget: pruning is sound while 2 * (err_tensor + err_reference) <= KAPPA * S.  The
    reference's own rounding is <= 8e-7 * S (cosine, worst case), the tensor path's rigorous bound is
    1.2e-6 * S (x_lo truncated by the tensor core, dropped x_lo.c_lo, c_lo rounding); KAPPA / 8 = 9.5e-7
    is the measured-error guard that keeps the sum under KAPPA / 2 with room to spare.  The margin of sub_dim D is
    KAPPA * D / 8 (more MMAs and products per score, longer sums in the reference)."""
    KAPPA = 2.0 ** -17 * (sub_dim // 8)
    n, m, k = 20_000, 8, 256
    dim = m * sub_dim
    x = mixture(n, dim, 5, scale=scale)
    cb = sample_codebooks(x, m, k, 6)
    for sub in (0, 5):
        d = dim // m
        xs = x[:, sub * d:(sub + 1) * d].astype(np.float64)
        c = cb[sub].astype(np.float64)
        # training / L2 kinds: ||c||^2 - 2 x.c
        got, rescans, codes = debug_scores(eng, x, cb, sub, cosine=False)
        want = (c * c).sum(1)[None, :] - 2.0 * xs @ c.T
        S = (np.sqrt((xs * xs).sum(1)) + np.sqrt((c * c).sum(1).max())) ** 2
        err = np.abs(got - want).max(1) / S
        print(f"scale={scale} sub={sub} L2: max err/S = {err.max():.3e} (margin {KAPPA:.3e}), "
              f"re-scanned pairs = {rescans} of {n * m}")
        assert err.max() <= KAPPA / 8
        assert rescans <= 0.005 * n * m
        assert np.array_equal(codes, assign_train(eng, x, cb, 1))
        # cosine: -x.c/||c||
        got, rescans, _ = debug_scores(eng, x, cb, sub, cosine=True)
        want = -(xs @ (c / np.sqrt((c * c).sum(1))[:, None]).T)
        S = np.sqrt((xs * xs).sum(1))
        err = np.abs(got - want).max(1) / S
        print(f"scale={scale} sub={sub} cos: max err/S = {err.max():.3e}, re-scanned pairs = {rescans} of {n * m}")
        assert err.max() <= KAPPA / 8
        assert rescans <= 0.005 * n * m


@pytest.mark.parametrize("n,dim,m,k", [(1024, 32, 4, 256), (5000, 64, 8, 256), (4097, 40, 5, 256), (3000, 96, 12, 100),
                                       (2500, 8, 1, 16), (20_000, 768, 96, 256), (1500, 24, 3, 7),
                                       # sub_dim 16 (BASELINE config 1), 24 (the reference's eval default, 384 / 16), 32
                                       (5000, 128, 8, 256), (3000, 384, 16, 256), (2500, 96, 3, 100), (4097, 48, 2, 256),
                                       (2048, 64, 2, 33), (1500, 80, 5, 256), (1200, 24, 1, 9), (20_000, 384, 16, 256)])
def test_tensor_train_assign_equals_exact_and_oracle(eng, oracle, n, dim, m, k):
    x = mixture(n, dim, n)
    cb = sample_codebooks(x, m, k, 1)
    cb[0, min(5, k - 1)] = cb[0, 2]      # duplicate centroid: lowest index must win (vector.rs:354-361)
    x[3] = x[2]                          # duplicate rows
    x[10, :dim // m] = cb[0, 2]          # a row that sits exactly on the duplicated centroid
    x[11] = 0.0
    x[12, 0] = np.nan                    # NaN distance at every centroid: index 0 sticks
    x[13, 3] = np.inf
    x[14] = 1e20                         # squares overflow
    x[15] = 1e-30
    tc = assign_train(eng, x, cb, 2)
    ex = assign_train(eng, x, cb, 1)
    assert np.array_equal(tc, ex)
    d = dim // m
    for s in range(0, m, max(1, m // 4)):
        _, want, _, _ = oracle.lbg_step(x, s * d, d, cb[s])
        assert np.array_equal(tc[s], want), f"subspace {s}"


@pytest.mark.parametrize("metric", ["squared_euclidean", "euclidean", "cosine"])
@pytest.mark.parametrize("n,dim,m,k", [(4000, 64, 8, 256), (3001, 40, 5, 50), (10_000, 768, 96, 256),
                                       (4000, 128, 8, 256), (3001, 384, 16, 256), (2500, 96, 3, 50), (10_000, 1536, 96, 256)])
def test_tensor_encode_equals_exact_and_oracle(vq, oracle, metric, n, dim, m, k):
    x = mixture(n, dim, 31)
    cb = sample_codebooks(x, m, k, 32)
    cb[0, 3] = cb[0, 1]          # duplicate
    cb[m - 1, 0] = 0.0           # zero centroid (cosine zero rules, cosine.c:38-45)
    x[0] = 0.0                   # zero vector
    x[1, :dim // m] = np.nan     # NaN sub-vector: index 0 sticks (pq.rs:183-191)
    x[2, 0] = np.inf
    x[3] = -x[4]                 # anti-correlated pair (cosine range [0, 2])
    x[5] = 3e19
    x[6] = 1e-25
    pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric))
    c_tc, r_tc = pq.encode_with_recon(x, assign="tensor")
    c_ex, r_ex = pq.encode_with_recon(x, assign="exact")
    assert np.array_equal(c_tc, c_ex)
    assert np.array_equal(r_tc.view(np.uint16), r_ex.view(np.uint16))
    if n <= 4000:
        want_codes, want_recon = oracle.pq_encode(cb, metric, x, sem="avx512")
        assert np.array_equal(c_tc.astype(np.uint32), want_codes)
        assert np.array_equal(r_tc.view(np.uint16), want_recon.view(np.uint16))


def test_tensor_near_tie_lattice(vq, eng):
    """Integer lattice data: thousands of exact ties and near-ties inside the margin; the re-scan path
    must reproduce the exact kernel bit for bit."""
    rng = np.random.default_rng(77)
    n, dim, m, k = 30_000, 32, 4, 256
    x = rng.integers(-3, 4, (n, dim)).astype(F)
    cb = rng.integers(-3, 4, (m, k, 8)).astype(F)
    assert np.array_equal(assign_train(eng, x, cb, 2), assign_train(eng, x, cb, 1))
    for metric in ("euclidean", "cosine"):
        pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric))
        assert np.array_equal(pq.encode(x, assign="tensor"), pq.encode(x, assign="exact"))


def test_tensor_unsafe_codebook_falls_back_per_row(vq, eng):
    """A NaN / Inf / huge centroid marks the subspace unsafe: every row is decided by the full scan."""
    n, dim, m, k = 2048, 32, 4, 64
    x = mixture(n, dim, 3)
    cb = sample_codebooks(x, m, k, 4)
    cb[1, 7, 2] = np.nan
    cb[2, 0, 0] = np.inf
    cb[3, 9] = 1e25
    assert np.array_equal(assign_train(eng, x, cb, 2), assign_train(eng, x, cb, 1))
    for metric in ("squared_euclidean", "cosine"):
        pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric))
        assert np.array_equal(pq.encode(x, assign="tensor"), pq.encode(x, assign="exact"))


def test_tensor_mode_rejects_unsupported_shapes(vq):
    x = mixture(2000, 64, 1)
    pq = vq.ProductQuantizer.from_codebooks(sample_codebooks(x, 16, 32, 2), vq.Distance.euclidean())  # sub_dim 4
    with pytest.raises(ValueError):
        pq.encode(x, assign="tensor")
    pq = vq.ProductQuantizer.from_codebooks(sample_codebooks(x, 8, 32, 2), vq.Distance.manhattan())
    with pytest.raises(ValueError):
        pq.encode(x, assign="tensor")
    assert pq.encode(x).shape == (2000, 8)  # auto: CUDA-core kernel


@pytest.mark.parametrize("dim,m", [(64, 8), (128, 8), (96, 4), (64, 2)])   # sub_dim 8, 16, 24, 32
def test_tensor_training_end_to_end_matches_oracle(vq, oracle, dim, m):
    """Ordered update + tensor assignment: the trained codebooks stay bit-identical with the oracle."""
    n, k, iters = 20_000, 256, 8
    x = mixture(n, dim, 20240)
    init, _ = vq.draw_init_indices(n, m, k, 42)
    pq = vq.ProductQuantizer(x, m, k, iters, vq.Distance.cosine(), init_idx=init, reseed=lambda s: 0, assign="tensor")
    want, it = oracle.pq_train(x, m, k, iters, init, reseed=lambda s: 0)
    assert np.array_equal(pq.iters_run, it)
    assert np.array_equal(pq.codebooks.view(np.uint32), want.view(np.uint32))


def test_full_size_properties_1Mx768(vq):
    """BASELINE.json's metric shape (1M x 768, m = 96, k = 256), checked through size-independent properties: the
    tensor-core route and the CUDA-core exact route give identical codes for every metric they share, training is
    deterministic (same indices -> bit-identical codebooks, tests/integration_tests.rs:40-53), the f16 reconstruction
    is exactly the chosen centroids, and quantisation error is far below the data variance."""
    torch = pytest.importorskip("torch")
    n, dim, m, k = 1_000_000, 768, 96, 256
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    centers = torch.randn(1024, dim, device="cuda", generator=g)
    x = torch.empty(n, dim, device="cuda")
    for r0 in range(0, n, 125_000):
        ids = torch.randint(0, 1024, (125_000,), device="cuda", generator=g)
        x[r0:r0 + 125_000] = centers[ids] + 0.25 * torch.randn(125_000, dim, device="cuda", generator=g)
    init, _ = vq.draw_init_indices(n, m, k, 42)
    pq = vq.ProductQuantizer(x, m, k, 3, vq.Distance.cosine(), init_idx=init, reseed=lambda s: 0, update="fast")
    pq2 = vq.ProductQuantizer(x, m, k, 3, vq.Distance.cosine(), init_idx=init, reseed=lambda s: 0, update="fast")
    assert np.array_equal(pq.codebooks.view(np.uint32), pq2.codebooks.view(np.uint32))
    assert int(pq.iters_run.min()) == 3
    for metric in ("cosine", "squared_euclidean", "euclidean"):
        q = vq.ProductQuantizer.from_codebooks(pq.codebooks, vq.Distance(metric))
        c_t = q.encode(x, assign="tensor")
        c_e = q.encode(x, assign="exact")
        assert torch.equal(c_t, c_e), metric
    codes, recon = pq.encode_with_recon(x)
    cb16 = torch.from_numpy(pq.codebooks.astype(np.float16)).cuda()            # [m, k, 8]
    want = cb16[torch.arange(m, device="cuda")[None, :], codes[:200_000].long()].reshape(200_000, dim)
    assert torch.equal(recon[:200_000].view(torch.int16), want.view(torch.int16))
    l2 = vq.ProductQuantizer.from_codebooks(pq.codebooks, vq.Distance.euclidean())
    rec = l2.decode(l2.encode(x))
    mse = float(((rec - x) ** 2).mean())
    assert np.isfinite(mse) and mse < 0.5 * float(x.var())


def test_tensor_adversarial_magnitudes_equals_exact(vq, eng):
    """The margin M = KAPPA * S is argued from S = (||x|| + max||c||)^2 (or ||x|| for cosine): stress exactly the cases
    that argument leans on -- magnitudes that differ by orders INSIDE a sub-vector, a codebook whose largest norm dwarfs
    the typical one, rows far smaller / larger than every centroid, and centroid pairs separated by a few ulps.  The
    pruned (tensor) route must still return the exact route's codes for the training distance and every encode metric."""
    rng = np.random.default_rng(4242)
    n, dim, m, k = 24_000, 32, 4, 256
    x = mixture(n, dim, 99)
    # magnitudes spread over 6 decades inside each sub-vector
    x[:8000] *= np.exp(rng.uniform(np.log(1e-3), np.log(1e3), (8000, dim))).astype(F)
    # rows that are tiny / huge against the codebook
    x[8000:9000] *= F(1e-4)
    x[9000:10000] *= F(3e3)
    cb = sample_codebooks(x[10000:], m, k, 5)
    cb[0, 17] *= F(2e3)                      # one centroid with a norm 2000 x the typical one: S is dominated by it
    cb[1, 3] = cb[1, 2] * F(1.0 + 2 ** -22)  # near-duplicates: distances differ by a few ulps
    cb[1, 5] = np.nextafter(cb[1, 4], F(np.inf))
    cb[2, 40:48] *= np.exp(rng.uniform(np.log(1e-3), np.log(1e3), (8, 8))).astype(F)
    cb[3, 0] = 0.0
    # rows sitting (almost) on the near-duplicate pairs
    x[10:20, 8:16] = cb[1, 2] * F(1.0 + 2 ** -23)
    x[20:30, 8:16] = (cb[1, 4].astype(np.float64) * 0.5 + cb[1, 5].astype(np.float64) * 0.5).astype(F)
    assert np.array_equal(assign_train(eng, x, cb, 2), assign_train(eng, x, cb, 1))
    for metric in ("squared_euclidean", "euclidean", "cosine"):
        pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric))
        c_t, r_t = pq.encode_with_recon(x, assign="tensor")
        c_e, r_e = pq.encode_with_recon(x, assign="exact")
        assert np.array_equal(c_t, c_e), metric
        assert np.array_equal(r_t.view(np.uint16), r_e.view(np.uint16)), metric
    # and the raw tensor scores stay inside the margin on this data too
    d = dim // m
    for sub, cosine in ((0, False), (2, False), (0, True), (2, True)):
        got, rescans, _ = debug_scores(eng, x, cb, sub, cosine=cosine)
        xs = x[:, sub * d:(sub + 1) * d].astype(np.float64)
        c = cb[sub].astype(np.float64)
        nc = np.sqrt((c * c).sum(1))
        if cosine:
            want = -(xs @ (c / np.where(nc > 0, nc, 1.0)[:, None]).T)
            S = np.sqrt((xs * xs).sum(1))
        else:
            want = (c * c).sum(1)[None, :] - 2.0 * xs @ c.T
            S = (np.sqrt((xs * xs).sum(1)) + nc.max()) ** 2
        ok = S > 0
        err = (np.abs(got - want).max(1)[ok] / S[ok]).max()
        print(f"adversarial sub={sub} cosine={cosine}: max err/S = {err:.3e}, re-scanned pairs = {rescans}")
        assert err <= KAPPA / 8
