"""Quick timing of TSVQ encode (1M x 1536, depth 8) + bit-comparison with the generic kernel (VQB_TSVQ_OLD_ENCODE=1 in a
second process).  Usage: python tools/tsvq_quick.py [dump.npy]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq

eng = vq.Engine(0)
ext = torch.cuda.ExternalStream(eng.stream, device=0)
n, dim, depth = 1_000_000, 1536, 8
g = torch.Generator(device="cuda"); g.manual_seed(5)
x = torch.randn(n, dim, device="cuda", generator=g)
for metric in ("squared_euclidean", "manhattan"):
    t = vq.TSVQ(x[:200_000], depth, vq.Distance(metric), engine=eng)
    leaf, recon = t._encode(x, True, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(ext)
    for _ in range(reps):
        leaf, recon = t._encode(x, True, True)
    e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = n * dim * 6 / 1e9
    h = int(torch.sum(leaf.to(torch.int64) * (torch.arange(n, device="cuda") % 1000 + 1)).item())
    hr = int(recon.view(torch.int16).to(torch.int64).sum().item())
    print(f"{metric:18s} {ms:.3f} ms per 1M x 1536 depth 8  ({gb / ms:.0f} GB/s algorithmic)  leaf checksum {h} recon checksum {hr}", flush=True)
