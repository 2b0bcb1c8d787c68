"""Diagnostics: where does a k-means iteration's wall time go?  (run on a GPU box, VQB_TRACE=1)"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq
from vq_b200 import _lib

rows, DIM, M, K = 1_000_000, 768, 96, 256
eng = vq.Engine(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
centers = torch.randn(1024, DIM, device="cuda", generator=g)
x = torch.empty(rows, DIM, device="cuda")
for r0 in range(0, rows, 131072):
    r1 = min(rows, r0 + 131072)
    ids = torch.randint(0, 1024, (r1 - r0,), device="cuda", generator=g)
    x[r0:r1] = centers[ids] + 0.25 * torch.randn(r1 - r0, DIM, device="cuda", generator=g)
torch.cuda.synchronize()
opts = _lib.TrainOpts(); opts.struct_size = C.sizeof(_lib.TrainOpts); opts.update_mode = _lib.UPDATE_FAST
cb = np.empty((M, K, DIM // M), np.float32); it_run = np.zeros(M, np.uint32)
ginit, _ = vq.draw_init_indices(rows, M, K, 42)
ginit = np.ascontiguousarray(ginit.reshape(-1))
for iters in (1, 1, 2, 6):
    t0 = time.perf_counter()
    eng.check(eng.lib.vqb_pq_train(eng.h, x.data_ptr(), rows, DIM, M, K, iters, ginit.ctypes.data, C.byref(opts),
                                   cb.ctypes.data, it_run.ctypes.data))
    torch.cuda.synchronize()
    print(f"train({iters}): {1e3 * (time.perf_counter() - t0):.2f} ms wall", flush=True)
