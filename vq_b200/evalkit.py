"""Evaluation harness of the reference, over this engine.

Mirrors `src/bin/common.rs:43-130` (synthetic data, reconstruction error, recall) and the loop of
`src/bin/eval_pq.rs:32-80` / `eval_tsvq.rs` / `eval_sq.rs` / `eval_bq.rs`: train, quantize every sample, report
training time, quantization time and the mean squared reconstruction error.  The metric definitions follow the
reference's arithmetic (f32, sequential sums); the generator cannot follow it bit for bit -- the reference draws
from `rand 0.9`'s `StdRng` + `rand_distr::Uniform`, which is not available here ("parity unpinned") --
so it draws the same distribution (uniform [0, 1), f32) from numpy's seeded generator instead.

Nothing in here is on the hot path: the quantizers it calls are (`vq_b200.api`), the metrics are host numpy.
"""
from __future__ import annotations

import time

import numpy as np

# src/bin/common.rs:9-15
SEED = 66
NUM_SAMPLES = (1_000, 5_000, 10_000, 50_000, 100_000, 1_000_000)
DIM, M, K, MAX_ITERS = 384, 16, 256, 10


def generate_synthetic_data(n_samples: int, n_dims: int, seed: int = SEED) -> np.ndarray:
    """Uniform [0, 1) f32 matrix [n_samples, n_dims] (common.rs:43-54; distribution only, see the module docstring)."""
    return np.random.default_rng(seed).random((n_samples, n_dims), dtype=np.float32)


def _seq_sum_f32(a: np.ndarray, axis: int) -> np.ndarray:
    """Left-to-right f32 sum along `axis` (Rust's `Iterator::sum::<f32>()`); numpy's own sum is pairwise."""
    a = np.asarray(a, dtype=np.float32)
    if a.shape[axis] == 0:
        return np.zeros(np.delete(a.shape, axis), np.float32)
    return np.take(np.cumsum(a, axis=axis, dtype=np.float32), -1, axis=axis)


def calculate_reconstruction_error(original: np.ndarray, reconstructed: np.ndarray, block: int = 65536) -> np.float32:
    """Mean squared error, common.rs:61-78: per vector an f32 sum of (x - y)^2, an f32 sum of those over the
    vectors, divided by `(n * dim) as f32`."""
    original = np.asarray(original, dtype=np.float32)
    reconstructed = np.asarray(reconstructed, dtype=np.float32)
    if original.shape != reconstructed.shape or original.ndim != 2:
        raise ValueError("original and reconstructed must be 2-D arrays of the same shape")
    n, dim = original.shape
    total = np.float32(0.0)
    for r0 in range(0, n, block):
        d = original[r0:r0 + block] - reconstructed[r0:r0 + block]
        rows = _seq_sum_f32(d * d, axis=1)
        # continue the running f32 sum through this block, in row order
        total = np.cumsum(np.concatenate(([total], rows)), dtype=np.float32)[-1]
    return np.float32(total / np.float32(n * dim))


def calculate_recall(original: np.ndarray, approx: np.ndarray, k: int) -> float:
    """recall@k estimate of common.rs:90-130: at most 1000 evenly spaced queries, neighbours searched in a window of
    5000 rows around the query (the whole set when n <= 10000), true neighbours by Euclidean distance in the
    original vectors, approximate ones in the reconstructed vectors, ties kept in index order (stable sort)."""
    original = np.asarray(original, dtype=np.float32)
    approx = np.asarray(approx, dtype=np.float32)
    n = original.shape[0]
    if n == 0 or k == 0:
        raise ValueError("recall needs samples and k > 0")
    eval_samples = min(n, 1000)
    step = max(n // eval_samples, 1)
    window = 5000 if n > 10_000 else n
    total = 0.0
    for i in range(0, n, step):
        lo, hi = max(i - window // 2, 0), min(i + window // 2, n)
        idx = np.array([j for j in range(lo, hi) if j != i], dtype=np.int64)

        def nearest(data):
            d = data[idx] - data[i]
            dist = np.sqrt(_seq_sum_f32(d * d, axis=1))
            return idx[np.argsort(dist, kind="stable")[:k]]

        true_nb, approx_nb = nearest(original), set(nearest(approx).tolist())   # HashSet of common.rs:122
        total += sum(1 for j in true_nb.tolist() if j in approx_nb) / float(k)
    return total / float(n // step)


def _result(n, dim, train_ms, quant_ms, err, ratio, recall=None):
    """BenchmarkResult of common.rs:18-34 as a dict."""
    return {"n_samples": int(n), "n_dims": int(dim), "training_time_ms": float(train_ms),
            "quantization_time_ms": float(quant_ms), "reconstruction_error": float(err),
            "recall": None if recall is None else float(recall), "memory_reduction_ratio": float(ratio)}


def eval_pq(n_samples: int, dim: int = DIM, m: int = M, k: int = K, max_iters: int = MAX_ITERS, seed: int = SEED,
            recall_k: int | None = None, engine=None) -> dict:
    """One iteration of eval_pq.rs's loop: train on the samples, quantize all of them (batch call instead of the
    reference's per-vector loop, same f16 output), reconstruction error through f16."""
    from . import api
    x = generate_synthetic_data(n_samples, dim, seed)
    t0 = time.perf_counter()
    pq = api.ProductQuantizer(x, m, k, max_iters, api.Distance.euclidean(), seed, engine=engine)
    train_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    q = pq.quantize_batch(x)                      # [n, dim] f16 == Vec<f16> per vector
    quant_ms = (time.perf_counter() - t0) * 1e3
    rec = np.asarray(q, dtype=np.float16).astype(np.float32)
    err = calculate_reconstruction_error(x, rec)
    recall = calculate_recall(x, rec, recall_k) if recall_k else None
    return _result(n_samples, dim, train_ms, quant_ms, err, (dim * 4) / float(m * (1 if k <= 256 else 2)), recall)


def eval_tsvq(n_samples: int, dim: int = DIM, max_depth: int = 5, seed: int = SEED, engine=None) -> dict:
    """eval_tsvq.rs:33-60 (default depth 5)."""
    from . import api
    x = generate_synthetic_data(n_samples, dim, seed)
    t0 = time.perf_counter()
    t = api.TSVQ(x, max_depth, api.Distance.euclidean(), engine=engine)
    train_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    q = t.quantize_batch(x)
    quant_ms = (time.perf_counter() - t0) * 1e3
    err = calculate_reconstruction_error(x, np.asarray(q, dtype=np.float16).astype(np.float32))
    return _result(n_samples, dim, train_ms, quant_ms, err, 2.0)   # the reference stores dim f16 values per vector


def eval_sq(n_samples: int, dim: int = DIM, levels: int = 256, seed: int = SEED, engine=None) -> dict:
    """eval_sq.rs:31-55: ScalarQuantizer::new(0.0, 1.0, levels)."""
    from . import api
    x = generate_synthetic_data(n_samples, dim, seed)
    t0 = time.perf_counter()
    sq = api.ScalarQuantizer(0.0, 1.0, levels, engine=engine)
    train_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    codes = sq.quantize(x)
    quant_ms = (time.perf_counter() - t0) * 1e3
    err = calculate_reconstruction_error(x, np.asarray(sq.dequantize(codes), dtype=np.float32).reshape(x.shape))
    return _result(n_samples, dim, train_ms, quant_ms, err, 4.0)


def eval_bq(n_samples: int, dim: int = DIM, threshold: float = 0.5, seed: int = SEED, engine=None) -> dict:
    """eval_bq.rs:28-52: BinaryQuantizer::new(0.5, 0, 1)."""
    from . import api
    x = generate_synthetic_data(n_samples, dim, seed)
    t0 = time.perf_counter()
    bq = api.BinaryQuantizer(threshold, 0, 1, engine=engine)
    train_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    codes = bq.quantize(x)
    quant_ms = (time.perf_counter() - t0) * 1e3
    err = calculate_reconstruction_error(x, np.asarray(bq.dequantize(codes), dtype=np.float32).reshape(x.shape))
    return _result(n_samples, dim, train_ms, quant_ms, err, 4.0)
