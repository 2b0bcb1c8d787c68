"""Latency of the reference's call shape: ProductQuantizer.quantize(one vector) / TSVQ.quantize / BQ / SQ with host buffers."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq
rng = np.random.default_rng(0)
x = rng.standard_normal((4096, 768)).astype(np.float32)
cb = np.stack([x[rng.choice(4096, 256, replace=False), s * 8:(s + 1) * 8] for s in range(96)]).astype(np.float32)
pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance("cosine"))
t = vq.TSVQ(x[:2048, :256].copy(), 6, vq.Distance("euclidean"))
bq, sq = vq.BinaryQuantizer(0.0, 0, 1), vq.ScalarQuantizer(-4.0, 4.0, 256)
def lat(fn, reps=300):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e6
v = x[7]; v2 = x[9, :256].copy()
print(f"pq.quantize(768-d): {lat(lambda: pq.quantize(v)):.1f} us   tsvq.quantize(256-d): {lat(lambda: t.quantize(v2)):.1f} us   "
      f"bq.quantize(768): {lat(lambda: bq.quantize(v)):.1f} us   sq.quantize(768): {lat(lambda: sq.quantize(v)):.1f} us", flush=True)
