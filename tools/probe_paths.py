"""Small driver for ncu launch lists: one TSVQ depth-8 build + encode on 1M x 1536 and two k-means iterations
on 1M x 768 (run as: ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/probe_paths.py)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq
from vq_b200 import _lib

eng = vq.Engine(0)
lib = eng.lib
what = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator(device="cuda"); g.manual_seed(5)
if what in ("all", "codec"):
    n, d = 1_000_000, 1536
    e = torch.empty(n, d, device="cuda").normal_(0.0, 0.5, generator=g)
    q = torch.empty(n, d, dtype=torch.uint8, device="cuda")
    ne = n * d
    eng.check(lib.vqb_bq_quantize(eng.h, e.data_ptr(), ne, 0.0, 0, 1, q.data_ptr()))
    eng.check(lib.vqb_sq_quantize(eng.h, e.data_ptr(), ne, -1.0, 1.0, float(np.float32(2.0) / np.float32(255.0)), 256, q.data_ptr()))
    eng.check(lib.vqb_sq_dequantize(eng.h, q.data_ptr(), ne, -1.0, float(np.float32(2.0) / np.float32(255.0)), e.data_ptr()))
    torch.cuda.synchronize()
    del e, q
if what in ("all", "manhattan"):
    rows, DIM, M, K = 1_000_000, 768, 96, 256
    x = torch.randn(rows, DIM, device="cuda", generator=g)
    cb = x[:K * 4:4].reshape(K, M, 8).permute(1, 0, 2).contiguous().cpu().numpy()
    pq = vq.ProductQuantizer.from_codebooks(cb, vq.Distance("manhattan"), engine=eng)
    codes = pq.encode(x)
    rec = pq.decode(codes)
    torch.cuda.synchronize()
    del x, codes, rec, pq
if what in ("all", "tsvq"):
    n, d = 1_000_000, 1536
    x = torch.empty(n, d, device="cuda").normal_(0.0, 0.5, generator=g)
    h = C.c_void_p()
    eng.check(lib.vqb_tsvq_train(eng.h, x.data_ptr(), n, d, 8, 1, C.byref(h)))
    r = torch.empty(n, d, dtype=torch.float16, device="cuda")
    eng.check(lib.vqb_tsvq_encode(h, x.data_ptr(), n, None, r.data_ptr()))
    torch.cuda.synchronize()
    lib.vqb_tsvq_destroy(h)
    del x, r
if what in ("all", "kmeans"):
    rows, DIM, M, K = 1_000_000, 768, 96, 256
    centers = torch.randn(1024, DIM, device="cuda", generator=g)
    x = torch.empty(rows, DIM, device="cuda")
    for r0 in range(0, rows, 131072):
        r1 = min(rows, r0 + 131072)
        ids = torch.randint(0, 1024, (r1 - r0,), device="cuda", generator=g)
        x[r0:r1] = centers[ids] + 0.25 * torch.randn(r1 - r0, DIM, device="cuda", generator=g)
    opts = _lib.TrainOpts(); opts.struct_size = C.sizeof(_lib.TrainOpts); opts.update_mode = _lib.UPDATE_FAST
    cb = np.empty((M, K, DIM // M), np.float32); it_run = np.zeros(M, np.uint32)
    ginit, _ = vq.draw_init_indices(rows, M, K, 42)
    ginit = np.ascontiguousarray(ginit.reshape(-1))
    eng.check(lib.vqb_pq_train(eng.h, x.data_ptr(), rows, DIM, M, K, 3, ginit.ctypes.data, C.byref(opts),
                               cb.ctypes.data, it_run.ctypes.data))
    torch.cuda.synchronize()
print("probe done")
