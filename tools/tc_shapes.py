"""Tensor-core vs CUDA-core assignment at the sub-vector lengths the tensor kernel covers (8, 16, 24, 32):
ms per 1M rows, encode (codes only), euclidean and cosine.  Shapes: BASELINE config 1 (128-d, m 8), the reference's eval
default (384-d, m 16, src/bin/common.rs:9-15), 256-d m 8, and the metric's 768-d m 96."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq

eng = vq.Engine(0)
lib = eng.lib
ext = torch.cuda.ExternalStream(eng.stream, device=0)
n, k = 1_000_000, 256
g = torch.Generator(device="cuda"); g.manual_seed(3)
for dim, m in ((128, 8), (384, 16), (256, 8), (768, 96)):
    d = dim // m
    centers = torch.randn(1024, dim, device="cuda", generator=g)
    x = centers[torch.randint(0, 1024, (n,), device="cuda", generator=g)] + 0.25 * torch.randn(n, dim, device="cuda", generator=g)
    cb = x[:k * 4:4].reshape(k, m, d).permute(1, 0, 2).contiguous().cpu().numpy()
    codes = {a: torch.empty(n, m, dtype=torch.uint8, device="cuda") for a in (1, 2)}
    for metric in ("euclidean", "cosine"):
        pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric), engine=eng)
        ms = {}
        for mode in (2, 1):   # VQB_ASSIGN_TENSOR, VQB_ASSIGN_EXACT
            reps = 5 if mode == 2 else 2
            for _ in range(2):
                eng.check(lib.vqb_pq_encode(pq._handle, x.data_ptr(), n, mode, codes[mode].data_ptr(), 1, None))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(reps):
                eng.check(lib.vqb_pq_encode(pq._handle, x.data_ptr(), n, mode, codes[mode].data_ptr(), 1, None))
            e1.record(ext)
            torch.cuda.synchronize()
            ms[mode] = e0.elapsed_time(e1) / reps
        same = bool(torch.equal(codes[1], codes[2]))
        tf = 2.0 * n * dim * k / ms[2] / 1e9
        print(f"{dim:4d}-d m={m:2d} sub_dim {d:2d} {metric:10s}: tensor {ms[2]:7.3f} ms ({n / ms[2] / 1e3:7.1f} Mvec/s, {tf:6.1f} TFLOP/s)  "
              f"CUDA-core {ms[1]:7.3f} ms  x{ms[1] / ms[2]:.1f}  codes equal on 1M rows: {same}", flush=True)
