"""Quick timing + cross-check of the tensor assignment kernel: 1M x 768, m = 96, k = 256, codebook = sampled rows.
Usage: python tools/tc_quick.py [variants...]   (variants = vqb_debug_tc_variant ids, default "0 1")"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq

eng = vq.Engine(0)
lib = eng.lib
ext = torch.cuda.ExternalStream(eng.stream, device=0)
n, dim, m, k = 1_000_000, 768, 96, 256
g = torch.Generator(device="cuda"); g.manual_seed(3)
centers = torch.randn(1024, dim, device="cuda", generator=g)
x = torch.empty(n, dim, device="cuda")
for r0 in range(0, n, 131072):
    r1 = min(n, r0 + 131072)
    x[r0:r1] = centers[torch.randint(0, 1024, (r1 - r0,), device="cuda", generator=g)] + 0.25 * torch.randn(r1 - r0, dim, device="cuda", generator=g)
cb = x[:k * 4:4].reshape(k, m, 8).permute(1, 0, 2).contiguous().cpu().numpy()
codes = torch.empty(n, m, dtype=torch.uint8, device="cuda")
chk = 65536
ref = torch.empty(chk, m, dtype=torch.uint8, device="cuda")
variants = [int(a) for a in sys.argv[1:]] or [0, 1]
for metric in ("cosine", "squared_euclidean", "euclidean"):
    pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance(metric), engine=eng)
    eng.check(lib.vqb_pq_encode(pq._handle, x.data_ptr(), chk, 1, ref.data_ptr(), 1, None))   # exact CUDA-core kernel
    for v in variants:
        lib.vqb_debug_tc_variant(v)
        for _ in range(3):
            eng.check(lib.vqb_pq_encode(pq._handle, x.data_ptr(), n, 2, codes.data_ptr(), 1, None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        reps = 10
        for _ in range(reps):
            eng.check(lib.vqb_pq_encode(pq._handle, x.data_ptr(), n, 2, codes.data_ptr(), 1, None))
        e1.record(ext)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        same = bool(torch.equal(codes[:chk], ref))
        print(f"{metric:18s} variant {v}: {ms:.3f} ms per 1M x 768  ({n / ms / 1e3:.1f} Mvec/s)  codes == exact kernel on {chk} rows: {same}", flush=True)
    # training metric through the one-step entry point is covered by tests/test_gpu_tensor.py
