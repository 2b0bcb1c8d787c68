"""ctypes front-end of the CPU oracle (oracle/vq_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py -- never by vq_b200/.

The oracle restates the reference's Rust loops (file:line cited in vq_oracle.c)
and, when ``oracle/_ref/libhsd_ref.so`` exists (hsdlib compiled verbatim from the
reference checkout by oracle/Makefile), routes the `simd`-build distance calls
through the real hsdlib (``sem="hsdlib"``).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvq_oracle.so")
HSD_PATH = os.path.join(HERE, "_ref", "libhsd_ref.so")

METRICS = {"squared_euclidean": 0, "euclidean": 1, "manhattan": 2, "cosine": 3,
           "chebyshev": 5}   # 5: extension, not in the reference
SEMS = {"scalar": 0, "avx512": 1, "avx2": 2, "hsdlib": 3}
# hsdlib.h:57-68
HSD_BACKENDS = {"auto": 0, "scalar": 1, "avx": 2, "avx2": 3, "avx512f": 4}

RESEED_FN = C.CFUNCTYPE(C.c_uint64, C.c_void_p, C.c_uint32)
HSD_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_float))


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    src = os.path.join(HERE, "vq_oracle.c")
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    need_ref = os.path.isdir("/root/reference/external/hsdlib/src") and not os.path.exists(HSD_PATH)
    if force or stale or need_ref:
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Hsdlib:
    """The reference's own C SIMD library, compiled verbatim (oracle/_ref).

    ``backend`` forces a dispatch target through hsd_set_manual_backend
    (utils.c:128-132); a private copy of the .so is loaded per instance because
    hsdlib caches the resolved kernel on first use (euclidean.c:264-276)."""

    def __init__(self, backend: str = "auto"):
        if not os.path.exists(HSD_PATH):
            raise FileNotFoundError(HSD_PATH)
        self._tmp = tempfile.NamedTemporaryFile(suffix=f"_hsd_{backend}.so", delete=False)
        self._tmp.close()
        shutil.copyfile(HSD_PATH, self._tmp.name)
        self.lib = C.CDLL(self._tmp.name)
        os.unlink(self._tmp.name)
        for name in ("hsd_dist_sqeuclidean_f32", "hsd_dist_manhattan_f32", "hsd_sim_cosine_f32"):
            f = getattr(self.lib, name)
            f.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_float)]
            f.restype = C.c_int
        self.lib.hsd_get_backend.restype = C.c_char_p
        self.lib.hsd_set_manual_backend.argtypes = [C.c_int]
        self.lib.hsd_cpu_has_avx512f.restype = C.c_bool
        if backend != "auto":
            self.lib.hsd_set_manual_backend(HSD_BACKENDS[backend])

    def backend(self) -> str:
        return self.lib.hsd_get_backend().decode()

    def has_avx512(self) -> bool:
        return bool(self.lib.hsd_cpu_has_avx512f())

    def _call(self, fn, a, b):
        a = _f32(a); b = _f32(b)
        out = C.c_float(0.0)
        st = fn(_fp(a), _fp(b), a.size, C.byref(out))
        return st, out.value

    def sqeuclidean(self, a, b):
        return self._call(self.lib.hsd_dist_sqeuclidean_f32, a, b)

    def manhattan(self, a, b):
        return self._call(self.lib.hsd_dist_manhattan_f32, a, b)

    def cosine(self, a, b):
        return self._call(self.lib.hsd_sim_cosine_f32, a, b)


class Oracle:
    def __init__(self, use_hsdlib: bool = True):
        build()
        L = self.lib = C.CDLL(LIB_PATH)
        self.hsd = None
        f32p, u32p, u16p, u8p, u64p, i32p = (C.POINTER(t) for t in
                                             (C.c_float, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint64, C.c_int32))
        sz = C.c_size_t
        L.vqo_distance2.argtypes = [f32p, f32p, sz]; L.vqo_distance2.restype = C.c_float
        L.vqo_distance.argtypes = [C.c_int, C.c_int, f32p, f32p, sz]; L.vqo_distance.restype = C.c_float
        L.vqo_f32_to_f16.argtypes = [C.c_float]; L.vqo_f32_to_f16.restype = C.c_uint16
        L.vqo_f16_to_f32.argtypes = [C.c_uint16]; L.vqo_f16_to_f32.restype = C.c_float
        for n in ("vqo_hsd_sqeuclid", "vqo_hsd_manhattan", "vqo_hsd_cosine"):
            getattr(L, n).argtypes = [C.c_int, f32p, f32p, sz, f32p]; getattr(L, n).restype = C.c_int
        L.vqo_set_hsdlib.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vqo_lbg_step.argtypes = [f32p, sz, sz, sz, sz, f32p, sz, u32p, u32p, u32p, C.c_int]
        L.vqo_lbg_step.restype = C.c_int
        L.vqo_pq_train.argtypes = [f32p, sz, sz, sz, sz, sz, u64p, RESEED_FN, C.c_void_p, f32p, u32p, C.c_int]
        L.vqo_pq_train.restype = C.c_int
        L.vqo_pq_encode.argtypes = [f32p, sz, sz, sz, C.c_int, C.c_int, f32p, sz, u32p, u16p, C.c_int]
        L.vqo_pq_encode.restype = C.c_int
        L.vqo_dequantize_f16.argtypes = [u16p, sz, f32p]
        L.vqo_bq_quantize.argtypes = [f32p, sz, C.c_float, C.c_uint8, C.c_uint8, u8p]
        L.vqo_bq_dequantize.argtypes = [u8p, sz, C.c_uint8, C.c_uint8, f32p]
        L.vqo_sq_step.argtypes = [C.c_float, C.c_float, sz]; L.vqo_sq_step.restype = C.c_float
        L.vqo_sq_quantize.argtypes = [f32p, sz, C.c_float, C.c_float, C.c_float, sz, u8p]
        L.vqo_sq_dequantize.argtypes = [u8p, sz, C.c_float, C.c_float, f32p]
        L.vqo_tsvq_build.argtypes = [f32p, sz, sz, sz, f32p, i32p, i32p, i32p, f32p, u64p, sz]
        L.vqo_tsvq_build.restype = C.c_int64
        L.vqo_tsvq_encode.argtypes = [f32p, i32p, i32p, sz, C.c_int, C.c_int, f32p, sz, u32p, u16p, C.c_int]
        L.vqo_tsvq_encode.restype = C.c_int
        L.vqo_recon_mse.argtypes = [f32p, u16p, sz]; L.vqo_recon_mse.restype = C.c_double
        if use_hsdlib and os.path.exists(HSD_PATH):
            self.hsd = Hsdlib("auto")
            addr = lambda f: C.cast(f, C.c_void_p)
            L.vqo_set_hsdlib(addr(self.hsd.lib.hsd_dist_sqeuclidean_f32),
                             addr(self.hsd.lib.hsd_dist_manhattan_f32),
                             addr(self.hsd.lib.hsd_sim_cosine_f32))

    # ---- which `simd` semantics stands for "the reference on this host" ----
    def default_sem(self) -> str:
        """'hsdlib' when the real library is loadable, else the AVX-512 restatement."""
        return "hsdlib" if self.hsd is not None else "avx512"

    @property
    def threads(self) -> int:
        return os.cpu_count() or 1

    # ---- distances ----
    def distance2(self, a, b):
        a = _f32(a); b = _f32(b)
        return self.lib.vqo_distance2(_fp(a), _fp(b), a.size)

    def distance(self, metric, a, b, sem="avx512"):
        a = _f32(a); b = _f32(b)
        return self.lib.vqo_distance(METRICS[metric], SEMS[sem], _fp(a), _fp(b), a.size)

    def hsd_restated(self, which, a, b, sem="avx512"):
        a = _f32(a); b = _f32(b)
        out = C.c_float(0.0)
        fn = {"sqeuclidean": self.lib.vqo_hsd_sqeuclid, "manhattan": self.lib.vqo_hsd_manhattan,
              "cosine": self.lib.vqo_hsd_cosine}[which]
        st = fn(SEMS[sem], _fp(a), _fp(b), a.size, C.byref(out))
        return st, out.value

    def f32_to_f16_bits(self, x: float) -> int:
        return self.lib.vqo_f32_to_f16(x)

    def f16_bits_to_f32(self, h: int) -> float:
        return self.lib.vqo_f16_to_f32(h)

    # ---- k-means ----
    def lbg_step(self, x, col0, d, centroids, threads=None):
        """One iteration on one subspace. Returns (new_centroids, assign, changed, empties)."""
        x = _f32(x)
        n, ld = x.shape
        cent = _f32(centroids).copy()
        k = cent.shape[0]
        assign = np.empty(n, np.uint32)
        empt = np.empty(k, np.uint32)
        ne = C.c_uint32(0)
        ch = self.lib.vqo_lbg_step(_fp(x), n, ld, col0, d, _fp(cent), k, _p(assign, C.c_uint32),
                                   _p(empt, C.c_uint32), C.byref(ne), threads or self.threads)
        return cent, assign, bool(ch), empt[: ne.value].copy()

    def pq_train(self, x, m, k, max_iters, init_idx, reseed=None, threads=None):
        """Returns (codebooks [m,k,sub_dim], iters_run [m]). ``reseed(subspace)->row``."""
        x = _f32(x)
        n, dim = x.shape
        init_idx = np.ascontiguousarray(init_idx, dtype=np.uint64).reshape(-1)
        assert init_idx.size == m * k
        d = dim // m if m else 0
        cb = np.zeros((m, k, d), np.float32)
        iters = np.zeros(m, np.uint32)
        cbk = RESEED_FN(lambda user, s: int(reseed(int(s))) if reseed else 0)
        rc = self.lib.vqo_pq_train(_fp(x), n, dim, m, k, max_iters, _p(init_idx, C.c_uint64), cbk, None,
                                   _fp(cb), _p(iters, C.c_uint32), threads or self.threads)
        if rc != 0:
            raise ValueError(f"oracle pq_train failed: {rc}")
        return cb, iters

    def pq_encode(self, codebooks, metric, x, sem="avx512", want_recon=True, threads=None):
        cb = _f32(codebooks)
        m, k, d = cb.shape
        x = _f32(x).reshape(-1, m * d)
        n = x.shape[0]
        codes = np.empty((n, m), np.uint32)
        recon = np.empty((n, m * d), np.uint16) if want_recon else None
        self.lib.vqo_pq_encode(_fp(cb), m, k, d, METRICS[metric], SEMS[sem], _fp(x), n,
                               _p(codes, C.c_uint32), _p(recon, C.c_uint16), threads or self.threads)
        return codes, (recon.view(np.float16) if want_recon else None)

    def dequantize_f16(self, q):
        q = np.ascontiguousarray(q).view(np.uint16).reshape(-1)
        out = np.empty(q.size, np.float32)
        self.lib.vqo_dequantize_f16(_p(q, C.c_uint16), q.size, _fp(out))
        return out

    # ---- BQ / SQ ----
    def bq_quantize(self, x, thr, low, high):
        x = _f32(x).reshape(-1)
        out = np.empty(x.size, np.uint8)
        self.lib.vqo_bq_quantize(_fp(x), x.size, thr, low, high, _p(out, C.c_uint8))
        return out

    def bq_dequantize(self, c, low, high):
        c = np.ascontiguousarray(c, dtype=np.uint8).reshape(-1)
        out = np.empty(c.size, np.float32)
        self.lib.vqo_bq_dequantize(_p(c, C.c_uint8), c.size, low, high, _fp(out))
        return out

    def sq_step(self, mn, mx, levels):
        return self.lib.vqo_sq_step(mn, mx, levels)

    def sq_quantize(self, x, mn, mx, levels):
        x = _f32(x).reshape(-1)
        out = np.empty(x.size, np.uint8)
        self.lib.vqo_sq_quantize(_fp(x), x.size, mn, mx, self.sq_step(mn, mx, levels), levels, _p(out, C.c_uint8))
        return out

    def sq_dequantize(self, c, mn, mx, levels):
        c = np.ascontiguousarray(c, dtype=np.uint8).reshape(-1)
        out = np.empty(c.size, np.float32)
        self.lib.vqo_sq_dequantize(_p(c, C.c_uint8), c.size, mn, self.sq_step(mn, mx, levels), _fp(out))
        return out

    # ---- TSVQ ----
    def tsvq_build(self, x, max_depth):
        x = _f32(x)
        n, dim = x.shape
        max_nodes = (1 << (max_depth + 1)) - 1
        max_nodes = min(max_nodes, 2 * n + 1)
        cent = np.zeros((max_nodes, dim), np.float32)
        left = np.full(max_nodes, -1, np.int32)
        right = np.full(max_nodes, -1, np.int32)
        sd = np.full(max_nodes, -1, np.int32)
        med = np.full(max_nodes, np.nan, np.float32)
        cnt = np.zeros(max_nodes, np.uint64)
        nn = self.lib.vqo_tsvq_build(_fp(x), n, dim, max_depth, _fp(cent), _p(left, C.c_int32),
                                     _p(right, C.c_int32), _p(sd, C.c_int32), _fp(med),
                                     _p(cnt, C.c_uint64), max_nodes)
        if nn <= 0:
            raise ValueError(f"oracle tsvq_build failed: {nn}")
        return dict(centroids=cent[:nn].copy(), left=left[:nn].copy(), right=right[:nn].copy(),
                    split_dim=sd[:nn].copy(), median=med[:nn].copy(), count=cnt[:nn].copy())

    def tsvq_encode(self, tree, metric, x, sem="avx512", want_recon=True, threads=None):
        cent = _f32(tree["centroids"])
        dim = cent.shape[1]
        x = _f32(x).reshape(-1, dim)
        n = x.shape[0]
        left = np.ascontiguousarray(tree["left"], np.int32)
        right = np.ascontiguousarray(tree["right"], np.int32)
        leaf = np.empty(n, np.uint32)
        recon = np.empty((n, dim), np.uint16) if want_recon else None
        self.lib.vqo_tsvq_encode(_fp(cent), _p(left, C.c_int32), _p(right, C.c_int32), dim,
                                 METRICS[metric], SEMS[sem], _fp(x), n, _p(leaf, C.c_uint32),
                                 _p(recon, C.c_uint16), threads or self.threads)
        return leaf, (recon.view(np.float16) if want_recon else None)

    def recon_mse(self, x, recon_f16):
        x = _f32(x).reshape(-1)
        r = np.ascontiguousarray(recon_f16).view(np.uint16).reshape(-1)
        return self.lib.vqo_recon_mse(_fp(x), _p(r, C.c_uint16), x.size)


_singleton = None


def get() -> Oracle:
    global _singleton
    if _singleton is None:
        _singleton = Oracle()
    return _singleton
