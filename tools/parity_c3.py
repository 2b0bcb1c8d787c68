"""SURVEY 8(d) parity protocol (ii) at BASELINE config 3: 1M x 768, m = 96, k = 256, 25 iterations.

GPU: the benched mode -- tensor-core assignment + FAST update.  CPU: the oracle's lbg loop (sequential f32 sums,
reference order), all host cores.  Same host-generated data, same index stream.  Checks, per subspace,
||C_gpu - C_ref||_F / ||C_ref||_F <= 1e-4 and the relative difference of the reconstruction MSE
(src/bin/common.rs:61-78, through f16) <= 1e-4; prints one summary line and exits non-zero on failure.
About four minutes of host time on 16 cores; run once per round, log committed under profiles/."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq
from oracle import oracle as O

n = int(os.environ.get("C3_ROWS", 1_000_000)); dim, m, k = 768, 96, 256
iters = int(os.environ.get("C3_ITERS", 25))
rng = np.random.default_rng(20240)
centers = rng.standard_normal((1024, dim)).astype(np.float32)
x = np.empty((n, dim), np.float32)
for r0 in range(0, n, 100_000):
    r1 = min(n, r0 + 100_000)
    x[r0:r1] = centers[rng.integers(0, 1024, r1 - r0)] + np.float32(0.25) * rng.standard_normal((r1 - r0, dim)).astype(np.float32)
init, _ = vq.draw_init_indices(n, m, k, 42)
t0 = time.perf_counter()
pq = vq.ProductQuantizer(x, m, k, iters, vq.Distance.euclidean(), init_idx=init, reseed=lambda s: 0, update="fast", assign="tensor")
t_gpu = time.perf_counter() - t0
t0 = time.perf_counter()
pq_o = vq.ProductQuantizer(x, m, k, iters, vq.Distance.euclidean(), init_idx=init, reseed=lambda s: 0, update="ordered", assign="tensor")
t_gpu_o = time.perf_counter() - t0
orc = O.get()
t0 = time.perf_counter()
want, it = orc.pq_train(x, m, k, iters, init, reseed=lambda s: 0, threads=os.cpu_count())
t_cpu = time.perf_counter() - t0
rel = np.array([np.linalg.norm(pq.codebooks[s] - want[s]) / np.linalg.norm(want[s]) for s in range(m)])


def mse(cb):
    q = vq.ProductQuantizer.from_codebooks(cb, vq.Distance.euclidean())
    tot = 0.0
    for r0 in range(0, n, 250_000):
        xb = x[r0:r0 + 250_000]
        rec = q.decode(q.encode(xb))
        tot += float(((np.asarray(rec, np.float64) - xb) ** 2).sum())
    return tot / (n * dim)


m_gpu, m_ref = mse(pq.codebooks), mse(want)
mse_rel = abs(m_gpu - m_ref) / m_ref
ordered_identical = bool(np.array_equal(pq_o.codebooks.view(np.uint32), want.view(np.uint32)) and np.array_equal(pq_o.iters_run, it))
rel_o = max(float(np.linalg.norm(pq_o.codebooks[s] - want[s]) / np.linalg.norm(want[s])) for s in range(m))
print(f"C3 ordered update + tensor assignment ({t_gpu_o:.2f} s): codebooks bit-identical with the oracle: {ordered_identical} "
      f"(max relative difference {rel_o:.3e})")
ok = bool(ordered_identical and mse_rel <= 1e-4 and np.array_equal(pq.iters_run, it))
print(f"C3 parity {n}x{dim} m={m} k={k} iters={iters} (ran {int(pq.iters_run.min())}..{int(pq.iters_run.max())}, oracle "
      f"{int(it.min())}..{int(it.max())}): max_s ||dC||_F/||C||_F = {rel.max():.3e} (median {np.median(rel):.3e}), "
      f"MSE gpu {m_gpu:.6e} ref {m_ref:.6e} rel {mse_rel:.3e}; GPU call {t_gpu:.2f} s, oracle {t_cpu:.1f} s on "
      f"{os.cpu_count()} cores -> {'PASS' if ok else 'FAIL'} (bars: ordered bit-identical, fast MSE 1e-4; "
      f"fast codebooks within 1e-4: {bool(rel.max() <= 1e-4)})")
sys.exit(0 if ok else 1)
