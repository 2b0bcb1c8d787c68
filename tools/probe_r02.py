"""Launches the round-2 kernels that bench.py's short profiler run does not reach: ordered tile update, Manhattan tiles."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq
eng = vq.Engine(0)
n, dim, m, k = 1_000_000, 768, 96, 256
g = torch.Generator(device="cuda"); g.manual_seed(3)
centers = torch.randn(1024, dim, device="cuda", generator=g)
x = torch.empty(n, dim, device="cuda")
for r0 in range(0, n, 125_000):
    x[r0:r0 + 125_000] = centers[torch.randint(0, 1024, (125_000,), device="cuda", generator=g)] + 0.25 * torch.randn(125_000, dim, device="cuda", generator=g)
init, _ = vq.draw_init_indices(n, m, k, 42)
for upd in ("fast", "ordered"):
    pq = vq.ProductQuantizer(x, m, k, 2, vq.Distance.euclidean(), init_idx=init, reseed=lambda s: 0, update=upd, engine=eng)
q = vq.ProductQuantizer.from_codebooks(pq.codebooks, vq.Distance.manhattan(), engine=eng)
q.encode(x)
print("probe done")
