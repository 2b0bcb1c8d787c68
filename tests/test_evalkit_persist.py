"""the reference's evaluation metrics (src/bin/common.rs) over this engine, and model
persistence.  CPU tests cover the metric arithmetic and the file-format validation; GPU tests the round trips."""
import os

import numpy as np
import pytest

from vq_b200 import evalkit, persist

F = np.float32


def test_reconstruction_error_is_the_reference_arithmetic():
    # common.rs:61-78 restated with explicit f32 loops: per-vector sequential sum, sequential sum over vectors
    rng = np.random.default_rng(3)
    a = rng.random((37, 19), dtype=F)
    b = (a + rng.normal(0, 0.05, a.shape)).astype(F)
    total = F(0.0)
    for o, r in zip(a, b):
        s = F(0.0)
        for x, y in zip(o, r):
            d = F(x - y)
            s = F(s + F(d * d))
        total = F(total + s)
    want = F(total / F(a.size))
    got = evalkit.calculate_reconstruction_error(a, b, block=8)   # block boundary inside the data
    assert got.dtype == np.float32 and got.tobytes() == want.tobytes()
    assert evalkit.calculate_reconstruction_error(a, a) == 0.0
    with pytest.raises(ValueError):
        evalkit.calculate_reconstruction_error(a, b[:, :5])


def test_recall_definition():
    rng = np.random.default_rng(4)
    x = rng.random((300, 8), dtype=F)
    assert evalkit.calculate_recall(x, x, 10) == 1.0                      # identical geometry
    y = x[::-1].copy()                                                     # unrelated geometry: recall ~ k / n
    r = evalkit.calculate_recall(x, y, 10)
    assert 0.0 <= r < 0.3
    # hand-checkable case: 1-D points, k = 1; approx swaps the neighbours of point 0 only
    o = np.array([[0.0], [1.0], [3.0], [7.0]], F)
    a = np.array([[0.0], [4.0], [1.0], [7.0]], F)
    # the window of common.rs:101-104 is [i - n/2, min(i + n/2, n)), not the whole set: query 0 only sees point 1
    # (agrees), queries 1..3: true NN 0, 1, 2 against approximate NN 2, 0, 1 (disagree)  =>  1/4
    assert evalkit.calculate_recall(o, a, 1) == 0.25
    assert evalkit.generate_synthetic_data(5, 3, 66).dtype == np.float32
    g = evalkit.generate_synthetic_data(1000, 4, 66)
    assert g.min() >= 0.0 and g.max() < 1.0 and np.array_equal(g, evalkit.generate_synthetic_data(1000, 4, 66))


def test_persist_format_validation(tmp_path):
    cb = np.zeros((2, 4, 3), F)
    good = {"kind": np.array("pq"), "version": np.array(persist.FORMAT_VERSION, np.int32), "metric": np.array("cosine"),
            "codebooks": cb}
    assert persist.check(good) == "pq"
    for bad in ({**good, "version": np.array(99, np.int32)}, {**good, "codebooks": cb.astype(np.float64)},
                {**good, "codebooks": cb[0]}, {**good, "kind": np.array("nope")}):
        with pytest.raises(ValueError):
            persist.check(bad)
    tree = {"kind": np.array("tsvq"), "version": np.array(persist.FORMAT_VERSION, np.int32), "metric": np.array("euclidean"),
            "centroids": np.zeros((3, 2), F), "left": np.array([1, -1, -1], np.int32), "right": np.array([2, -1, -1], np.int32)}
    assert persist.check(tree) == "tsvq"
    with pytest.raises(ValueError):   # a child that points backwards would make the descent loop for ever
        persist.check({**tree, "left": np.array([1, 0, -1], np.int32)})
    with pytest.raises(ValueError):
        persist.check({**tree, "right": np.array([5, -1, -1], np.int32)})
    p = os.path.join(tmp_path, "m.npz")
    np.savez(p, **good)
    with np.load(p, allow_pickle=False) as z:
        assert persist.check({k: z[k] for k in z.files}) == "pq"


@pytest.mark.gpu
def test_persist_round_trips(tmp_path):
    import vq_b200 as vq
    rng = np.random.default_rng(5)
    x = rng.standard_normal((5000, 32)).astype(F)
    pq = vq.ProductQuantizer(x, 4, 64, 5, vq.Distance.cosine(), 7)
    p = os.path.join(tmp_path, "pq.npz")
    persist.save(pq, p)
    pq2 = persist.load(p)
    assert pq2.distance_metric() == pq.distance_metric()
    assert np.array_equal(pq2.codebooks.view(np.uint32), pq.codebooks.view(np.uint32))
    assert np.array_equal(pq2.encode(x), pq.encode(x))
    t = vq.TSVQ(x, 5, vq.Distance.manhattan())
    p = os.path.join(tmp_path, "tsvq.npz")
    persist.save(t, p)
    t2 = persist.load(p)
    assert t2.distance_metric() == "manhattan" and np.array_equal(t2.encode(x), t.encode(x))
    assert np.array_equal(t2.quantize_batch(x).view(np.uint16), t.quantize_batch(x).view(np.uint16))
    for m in (vq.ScalarQuantizer(-2.0, 3.0, 100), vq.BinaryQuantizer(0.25, 3, 9)):
        p = os.path.join(tmp_path, "s.npz")
        persist.save(m, p)
        assert np.array_equal(persist.load(p).quantize(x), m.quantize(x))


@pytest.mark.gpu
def test_eval_harness_runs_and_is_sane():
    r = evalkit.eval_pq(5000, dim=64, m=8, k=64, max_iters=5, recall_k=10)
    assert r["n_samples"] == 5000 and r["n_dims"] == 64 and r["memory_reduction_ratio"] == 32.0
    assert 0.0 < r["reconstruction_error"] < 1.0 / 12.0          # below the variance of uniform [0, 1)
    assert 0.0 < r["recall"] <= 1.0
    assert evalkit.eval_sq(2000, dim=32)["reconstruction_error"] < 1e-5      # 256 levels over [0, 1]
    assert abs(evalkit.eval_bq(2000, dim=32)["reconstruction_error"] - 1.0 / 12.0) < 0.01   # E[(u - [u >= .5])^2] = 1/12
    t = evalkit.eval_tsvq(3000, dim=32, max_depth=4)
    assert 0.0 < t["reconstruction_error"] < 1.0 / 12.0
