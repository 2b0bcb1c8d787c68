#!/usr/bin/env python
"""PCIe probe for the host-pointer encode pipeline: H2D, D2H and both at once from pinned memory."""
import time
import torch

def bw(fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9

for mb in (16, 128, 1024):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def h2d():
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    def both():
        h2d(); d2h()
    print(f"{mb:5d} MB chunks: H2D {bw(h2d, n):6.1f} GB/s   D2H {bw(d2h, n // 2):6.1f} GB/s   "
          f"both (H2D bytes / time) {bw(both, n):6.1f} GB/s", flush=True)
