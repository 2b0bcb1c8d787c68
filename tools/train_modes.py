"""k-means iteration time of the update modes on 1M x 768 (m = 96, k = 256, tensor assignment): device time per iteration
(vqb_train_opts.iter_ms) and wall time of a 25-iteration call.  Usage: python tools/train_modes.py [rows]"""
import ctypes as C, os, statistics, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq
from vq_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim, m, k, iters = 768, 96, 256, 25
eng = vq.Engine(0)
g = torch.Generator(device="cuda"); g.manual_seed(20240)
centers = torch.randn(1024, dim, device="cuda", generator=g)
x = torch.empty(n, dim, device="cuda")
for r0 in range(0, n, 131072):
    r1 = min(n, r0 + 131072)
    x[r0:r1] = centers[torch.randint(0, 1024, (r1 - r0,), device="cuda", generator=g)] + 0.25 * torch.randn(r1 - r0, dim, device="cuda", generator=g)
init, _ = vq.draw_init_indices(n, m, k, 42)
init = np.ascontiguousarray(init.reshape(-1))
cb = np.empty((m, k, dim // m), np.float32); it_run = np.zeros(m, np.uint32)
for name, mode in (("fast", _lib.UPDATE_FAST), ("ordered", _lib.UPDATE_ORDERED)):
    opts = _lib.TrainOpts(); opts.struct_size = C.sizeof(_lib.TrainOpts); opts.update_mode = mode
    ms = (C.c_float * iters)(); opts.iter_ms = C.cast(ms, C.POINTER(C.c_float))
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.check(eng.lib.vqb_pq_train(eng.h, x.data_ptr(), n, dim, m, k, iters, init.ctypes.data, C.byref(opts), cb.ctypes.data, it_run.ctypes.data))
        wall = time.perf_counter() - t0
    print(f"{name:8s}: median iteration {statistics.median(list(ms)):.3f} ms, 25-iteration call {wall * 1e3:.1f} ms, iterations run {int(it_run.min())}", flush=True)
