"""Numerics behind DESIGN.md 6 item 1 (CPU only): error of GEMM-form scores ||c||^2 - 2 x.c when x and -2c are split into
two fp16 terms (x1 c1 + x1 c2 + x2 c1, products exact, accumulation in higher precision) against the 3xTF32 split the
tensor kernel uses today, both relative to S = (||x|| + max ||c||)^2, the quantity the kernel's margin KAPPA * S is built on."""
import numpy as np

rng = np.random.default_rng(0)
n, k, d = 4096, 256, 8
centers = rng.standard_normal((1024, d)).astype(np.float32)
x = (centers[rng.integers(0, 1024, n)] + 0.25 * rng.standard_normal((n, d))).astype(np.float32)
c = x[rng.choice(n, k, replace=False)].copy()
n2 = (c.astype(np.float64) ** 2).sum(1)
exact = (-2 * x.astype(np.float64)) @ c.astype(np.float64).T
S = (np.linalg.norm(x.astype(np.float64), axis=1)[:, None] + np.sqrt(n2.max())) ** 2


def split16(a, scale):
    a = a.astype(np.float32) * np.float32(scale)
    h = a.astype(np.float16)
    lo = (a - h.astype(np.float32)).astype(np.float16)
    return h.astype(np.float64), lo.astype(np.float64)


def tf32(a):
    return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


print(f"KAPPA = 2^-17 = {2.0 ** -17:.3e}")
for s in (1.0, 2.0 ** -6, 2.0 ** 8):
    xh, xl = split16(x, s)
    ch, cl = split16(-2 * c, s)
    dot = (xh @ ch.T + xh @ cl.T + xl @ ch.T) / (s * s)
    print(f"fp16 two-term split, operands scaled by {s:g}: max |err| / S = {(np.abs(dot - exact) / S).max():.3e}")
xh = tf32(x); xl = tf32((x - xh).astype(np.float32))
b = (-2 * c).astype(np.float32); bh = tf32(b); bl = tf32((b - bh).astype(np.float32))
dot = xh.astype(np.float64) @ bh.astype(np.float64).T + xh.astype(np.float64) @ bl.astype(np.float64).T + \
      xl.astype(np.float64) @ bh.astype(np.float64).T
print(f"3xTF32 split (today's kernel, truncating hi part):  max |err| / S = {(np.abs(dot - exact) / S).max():.3e}")
