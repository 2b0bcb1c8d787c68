"""Cumulative TSVQ build time by depth (1M x 1536): differences give the cost of each level."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq
eng = vq.Engine(0)
n, dim = 1_000_000, 1536
g = torch.Generator(device="cuda"); g.manual_seed(5)
x = torch.randn(n, dim, device="cuda", generator=g)
prev = 0.0
out = []
for depth in range(1, 9):
    ts = []
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        t = vq.TSVQ(x, depth, vq.Distance("squared_euclidean"), engine=eng)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    m = min(ts[1:])
    out.append(f"d{depth}: {m:.1f} (+{m - prev:.1f})")
    prev = m
print("PC_ROWS=" + os.environ.get("VQB_TSVQ_PC_ROWS", "auto") + "  " + "  ".join(out), flush=True)
