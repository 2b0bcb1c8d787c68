"""CUDA-core exact assignment at sub-vector lengths outside the tensor kernel's set: ms per 1M rows (euclidean, k = 256)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq
eng = vq.Engine(0); lib = eng.lib
ext = torch.cuda.ExternalStream(eng.stream, device=0)
n, k = 1_000_000, 256
g = torch.Generator(device="cuda"); g.manual_seed(3)
for dim, m in ((192, 16), (320, 8), (384, 16), (160, 8), (512, 8), (208, 16)):
    d = dim // m
    x = torch.randn(n, dim, device="cuda", generator=g)
    cb = x[:k * 4:4].reshape(k, m, d).permute(1, 0, 2).contiguous().cpu().numpy()
    pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance("euclidean"), engine=eng)
    codes = torch.empty(n, m, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eng.check(lib.vqb_pq_encode(pq._handle, x.data_ptr(), n, 1, codes.data_ptr(), 1, None))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    eng.check(lib.vqb_pq_encode(pq._handle, x.data_ptr(), n, 1, codes.data_ptr(), 1, None))
    e1.record(ext); torch.cuda.synchronize()
    print(f"{dim}-d m={m} sub_dim {d}: exact kernel {e0.elapsed_time(e1):.2f} ms per 1M rows", flush=True)
