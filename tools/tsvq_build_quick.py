"""TSVQ build timing (1M x 1536, depth 8): median / min of 5 calls, and a checksum of the tree (split dims, medians,
centroid bits) to compare builds of different kernels.  VQB_TSVQ_SLICE=32 forces the 32-column slices of round 1."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq

eng = vq.Engine(0)
n, dim, depth = 1_000_000, 1536, 8
g = torch.Generator(device="cuda"); g.manual_seed(5)
x = torch.randn(n, dim, device="cuda", generator=g)
ts = []
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t = vq.TSVQ(x, depth, vq.Distance("squared_euclidean"), engine=eng)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
tree = t.export() if hasattr(t, "export") else t.tree()
h = int(np.asarray(tree["split_dim"], np.int64).sum()) , int(np.asarray(tree["median"], np.float32).view(np.uint32).astype(np.int64).sum()), int(np.asarray(tree["centroids"], np.float32).view(np.uint32).astype(np.int64).sum())
ts = ts[1:]
print(f"build 1M x 1536 depth 8: median {np.median(ts):.2f} ms  min {min(ts):.2f}  max {max(ts):.2f}  checksums {h}", flush=True)
