"""Per-unit hand-off timeline of the tensor assignment kernel (CTA 0): where the MMA -> scan -> MMA cycle spends its time."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq

eng = vq.Engine(0)
n, dim, m, k = 1_000_000, 768, 96, 256
g = torch.Generator(device="cuda"); g.manual_seed(3)
centers = torch.randn(1024, dim, device="cuda", generator=g)
x = centers[torch.randint(0, 1024, (n,), device="cuda", generator=g)] + 0.25 * torch.randn(n, dim, device="cuda", generator=g)
cb = x[:k * 4:4].reshape(k, m, 8).permute(1, 0, 2).contiguous()
units = 400
ts = np.zeros((units, 8), np.uint64)
for _ in range(2):
    eng.check(eng.lib.vqb_debug_tc_timeline(eng.h, x.data_ptr(), n, dim, m, k, cb.data_ptr(), ts.ctypes.data, units))
t = ts.astype(np.int64)
t0 = t[t > 0].min()
names = ["iss_saw_empty", "iss_issued", "scan_saw_full", "scan_released", "scan_published", "res_saw", "res_done", "released_last"]
lo, hi = 200, 216
print("unit " + " ".join(f"{nm:>14s}" for nm in names))
for u in range(lo, hi):
    print(f"{u:4d} " + " ".join(f"{int(v - t0):14d}" for v in t[u]))
d = lambda a, b: np.median((t[100:380, a] - t[100:380, b]))
print("median cycles: unit period (issuer)", np.median(np.diff(t[100:380, 0])),
      "| issue dur", d(1, 0), "| issued -> scan saw full", d(2, 1), "| drain (saw full -> released)", d(3, 2),
      "| post (released -> published)", d(4, 3), "| published -> resolve saw", d(5, 4), "| resolve dur", d(6, 5))
# accumulator a is released by unit u and next observed empty by unit u+2's issue
rel = t[100:378, 3]; nxt = t[102:380, 0]
rl = t[100:378, 7]
print("median released by warp 0 (u) -> issuer saw empty(u+2):", np.median(nxt - rel), "| last warp's release after warp 0's:", d(7, 3),
      "| last release(u) -> issuer saw empty(u+2):", np.median(nxt - rl))
