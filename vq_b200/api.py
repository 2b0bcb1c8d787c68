"""Host-side mirror of the reference's quantizer API over the C ABI (include/vqb200.h).

Same class names, constructor arguments, properties and error messages as the
reference's Python surface (pyvq/pyvq.pyi:91-383, pyvq/src/{bq,sq,pq,tsvq,distance}.rs)
and therefore of the Rust `Quantizer` trait it wraps (src/core/quantizer.rs:29-63):
`quantize` takes ONE vector and returns float16 (PQ/TSVQ: the reconstructed centroid
values, src/pq.rs:193-195) or uint8 (BQ/SQ); `dequantize` returns float32.  Batch
methods (`quantize_batch`, `encode`, `decode`) are the additions that let a caller keep
data on the GPU; they accept numpy arrays or torch CUDA tensors (zero-copy).

Everything computes in libvqb200.so on the GPU.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
import weakref

import numpy as np

from . import _lib
from .rand09 import IndexStream

try:  # torch is optional plumbing: only needed for CUDA-tensor inputs and torch.distributed
    import torch
except Exception:  # pragma: no cover
    torch = None


# --------------------------------------------------------------------------- errors
class VqError(ValueError):
    """src/core/error.rs:5-28; pyvq maps every VqError to ValueError(str(e))."""


class DimensionMismatch(VqError):
    def __init__(self, expected, found):
        super().__init__(f"Dimension mismatch: expected {expected}, found {found}")
        self.expected, self.found = expected, found


class EmptyInput(VqError):
    def __init__(self):
        super().__init__("Empty input: at least one vector is required")


class InvalidParameter(VqError):
    def __init__(self, parameter, reason):
        super().__init__(f"Invalid parameter '{parameter}': {reason}")
        self.parameter, self.reason = parameter, reason


class FfiError(VqError):
    def __init__(self, msg):
        super().__init__(f"FFI error: {msg}")


# --------------------------------------------------------------------------- engine
_live_engines = weakref.WeakSet()   # every open Engine: CUDA-tensor arguments are ordered against their streams


class Engine:
    """One vqb_ctx: a GPU, its streams and scratch memory."""

    def __init__(self, device: int | None = None):
        self.lib = _lib.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        rc = self.lib.vqb_ctx_create(device, C.byref(h))
        if rc == _lib.ERR_UNSUPPORTED_DEVICE:
            raise RuntimeError("vq_b200 needs an sm_100 (B200) GPU: no CUDA device or wrong architecture; "
                               "there is no CPU fallback")
        if rc != 0:
            raise RuntimeError(f"vqb_ctx_create failed ({rc})")
        self.h = h
        self.device = device
        _live_engines.add(self)

    def close(self):
        if getattr(self, "h", None):
            self.lib.vqb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc == 0:
            return
        msg = self.lib.vqb_last_error(self.h).decode(errors="replace")
        if rc == _lib.ERR_EMPTY_INPUT:
            raise EmptyInput()
        if rc in (_lib.ERR_INVALID_INPUT, _lib.ERR_DIM_MISMATCH, _lib.ERR_NULL_PTR):
            raise VqError(msg or f"invalid input ({rc})")
        raise FfiError(msg or f"status {rc}")

    def synchronize(self):
        self.check(self.lib.vqb_ctx_synchronize(self.h))

    @property
    def stream(self) -> int:
        return self.lib.vqb_ctx_stream(self.h) or 0

    def set_stream(self, cuda_stream: int):
        self.check(self.lib.vqb_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.vqb_ctx_launch_count(self.h))

    # pinned host memory as a numpy array (for end-to-end pipelines)
    def pinned_empty(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self.check(self.lib.vqb_host_alloc(self.h, max(nbytes, 1), C.byref(p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        return arr


_engines: dict[int, Engine] = {}
_engines_lock = threading.Lock()


def default_engine(like=None) -> Engine:
    """The process-wide engine of a device: the device of `like` when it is a CUDA tensor, else torch's current device
    (LOCAL_RANK without torch)."""
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    if torch is not None and torch.cuda.is_available():
        try:
            dev = torch.cuda.current_device()
        except Exception:
            pass
    if like is not None and torch is not None and isinstance(like, torch.Tensor) and like.is_cuda:
        dev = like.device.index
    with _engines_lock:
        if dev not in _engines:
            _engines[dev] = Engine(dev)
        return _engines[dev]


def get_simd_backend() -> str:
    """pyvq.get_simd_backend (src/core/hsdlib_ffi.rs:144-155): names the compute backend."""
    return _lib.load().vqb_backend_name().decode()


# --------------------------------------------------------------------------- array plumbing
def _is_torch(a):
    return torch is not None and isinstance(a, torch.Tensor)


def _torch_dtype(dtype, signed_alias=False):
    """numpy dtype -> torch dtype, looked up lazily (torch.uint16 / uint32 exist from torch 2.3 on)."""
    name = {np.float32: "float32", np.uint8: "uint8", np.float16: "float16", np.int32: "int32",
            np.uint16: "int16" if signed_alias else "uint16", np.uint32: "int32" if signed_alias else "uint32"}[dtype]
    return getattr(torch, name)


def _order_after_torch(t):
    """The engine launches on its own stream: make it wait for whatever torch has queued for this tensor (and for the
    caching allocator's previous users of its memory) on torch's current stream.  Every call that takes CUDA tensors
    synchronises the engine before it returns (_sync_if_torch / blocking training calls), so temporaries made here are
    never recycled while an engine kernel still reads them."""
    if not t.is_cuda:
        return
    cur = torch.cuda.current_stream(t.device)
    for e in list(_live_engines):
        if e.h and e.device == t.device.index:
            es = torch.cuda.ExternalStream(e.stream, device=t.device)
            if es.cuda_stream != cur.cuda_stream:
                es.wait_stream(cur)


def _same_device(eng, *arrays):
    """A CUDA tensor must live on the engine's GPU: the library only checks that a pointer is device memory, not whose."""
    for a in arrays:
        if _is_torch(a) and a.is_cuda and a.device.index != eng.device:
            raise ValueError(f"tensor on cuda:{a.device.index} passed to the engine of cuda:{eng.device}")
    return eng


def _in(a, dtype):
    """-> (pointer, keepalive, on_cuda). numpy arrays are made contiguous; torch tensors pass zero-copy."""
    if _is_torch(a):
        tdt = _torch_dtype(dtype)
        t = a.detach()
        if t.dtype != tdt:
            # same-width signed integers (what _out_like hands out for u16 / u32 codes) are reinterpreted, not converted
            if dtype in (np.uint16, np.uint32) and t.dtype == _torch_dtype(dtype, signed_alias=True):
                t = t.contiguous().view(tdt)
            else:
                t = t.to(tdt)
        t = t.contiguous()
        _order_after_torch(t)
        return C.c_void_p(t.data_ptr()), t, t.is_cuda
    arr = np.ascontiguousarray(a, dtype=dtype)
    return C.c_void_p(arr.ctypes.data), arr, False


def _out_like(src, shape, dtype):
    """Output buffer on the same side as `src`."""
    if _is_torch(src) and src.is_cuda:
        t = torch.empty(tuple(shape), dtype=_torch_dtype(dtype, signed_alias=True), device=src.device)
        _order_after_torch(t)
        return C.c_void_p(t.data_ptr()), t
    arr = np.empty(shape, dtype=dtype)
    return C.c_void_p(arr.ctypes.data), arr


def _sync_if_torch(engine, src):
    # device-pointer calls are asynchronous on the engine stream; hand results back completed
    if _is_torch(src) and src.is_cuda:
        engine.synchronize()


# --------------------------------------------------------------------------- Distance
class Distance:
    """pyvq.Distance (pyvq/src/distance.rs:33-97) over src/core/distance.rs:8-65."""

    _NAMES = {"euclidean": "euclidean", "squaredeuclidean": "squared_euclidean",
              "squared_euclidean": "squared_euclidean", "cosine": "cosine", "cosine_distance": "cosine",
              "manhattan": "manhattan"}

    def __init__(self, metric: str):
        key = str(metric).lower()
        if key not in self._NAMES:
            raise ValueError("Invalid distance metric. Choose from: euclidean, squared_euclidean, cosine, manhattan")
        self.metric = self._NAMES[key]

    @staticmethod
    def euclidean():
        return Distance("euclidean")

    @staticmethod
    def squared_euclidean():
        return Distance("squared_euclidean")

    @staticmethod
    def manhattan():
        return Distance("manhattan")

    @staticmethod
    def cosine():
        return Distance("cosine")

    @staticmethod
    def chebyshev():
        """EXTENSION, not in pyvq (the reference's Distance has four kinds, src/core/distance.rs:8-17): max_i |a_i - b_i|.
        Reachable through this method only -- Distance("chebyshev") raises like the reference does.  Usable for
        Distance.compute and ProductQuantizer encoding (CUDA-core assignment kernel); TSVQ rejects it."""
        d = Distance("manhattan")
        d.metric = "chebyshev"
        return d

    @property
    def id(self) -> int:
        return _lib.METRIC_IDS[self.metric]

    def name(self) -> str:
        return self.metric

    def compute(self, a, b, engine: Engine | None = None) -> float:
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
        b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
        if a.size != b.size:
            raise DimensionMismatch(a.size, b.size)  # distance.rs:49-54
        return float(self.compute_batch(a[None, :], b[None, :], engine)[0])

    def compute_batch(self, a, b, engine: Engine | None = None):
        eng = _same_device(engine or default_engine(a), a, b)
        pa, ka, _ = _in(a, np.float32)
        pb, kb, _ = _in(b, np.float32)
        if tuple(ka.shape) != tuple(kb.shape):
            raise DimensionMismatch(ka.shape[-1], kb.shape[-1])
        rows, n = (ka.shape[0], ka.shape[1]) if ka.ndim == 2 else (1, ka.shape[0])
        po, out = _out_like(a, (rows,), np.float32)
        eng.check(eng.lib.vqb_distance_batch(eng.h, self.id, pa, pb, rows, n, po))
        _sync_if_torch(eng, a)
        return out

    def __repr__(self):
        return f"Distance(metric={self.metric})"

    def __eq__(self, other):
        return isinstance(other, Distance) and other.metric == self.metric


# --------------------------------------------------------------------------- BQ / SQ
class BinaryQuantizer:
    """pyvq.BinaryQuantizer (pyvq/src/bq.rs:32-37) over src/bq.rs:55-118."""

    def __init__(self, threshold: float, low: int = 0, high: int = 1, engine: Engine | None = None):
        thr = float(np.float32(threshold))
        if not math.isfinite(thr):
            raise InvalidParameter("threshold", "must be finite (not NaN or infinite)")
        if not (0 <= int(low) <= 255 and 0 <= int(high) <= 255):
            raise OverflowError("low/high must fit in u8")
        if int(low) >= int(high):
            raise InvalidParameter("low/high", "low must be less than high")
        self._thr, self._low, self._high = thr, int(low), int(high)
        self._engine = engine

    threshold = property(lambda self: self._thr)
    low = property(lambda self: self._low)
    high = property(lambda self: self._high)

    def quantize(self, values):
        eng = _same_device(self._engine or default_engine(values), values)
        p, keep, _ = _in(values, np.float32)
        po, out = _out_like(values, tuple(keep.shape), np.uint8)
        n = int(np.prod(keep.shape))
        eng.check(eng.lib.vqb_bq_quantize(eng.h, p, n, self._thr, self._low, self._high, po))
        _sync_if_torch(eng, values)
        return out

    def dequantize(self, codes):
        eng = _same_device(self._engine or default_engine(codes), codes)
        p, keep, _ = _in(codes, np.uint8)
        po, out = _out_like(codes, tuple(keep.shape), np.float32)
        eng.check(eng.lib.vqb_bq_dequantize(eng.h, p, int(np.prod(keep.shape)), self._low, self._high, po))
        _sync_if_torch(eng, codes)
        return out

    def __repr__(self):
        return f"BinaryQuantizer(threshold={self._thr}, low={self._low}, high={self._high})"


class ScalarQuantizer:
    """pyvq.ScalarQuantizer (pyvq/src/sq.rs:32-37) over src/sq.rs:63-151."""

    def __init__(self, min: float, max: float, levels: int = 256, engine: Engine | None = None):
        mn, mx = np.float32(min), np.float32(max)
        if not np.isfinite(mn):
            raise InvalidParameter("min", "must be finite (not NaN or infinite)")
        if not np.isfinite(mx):
            raise InvalidParameter("max", "must be finite (not NaN or infinite)")
        if mx <= mn:
            raise InvalidParameter("max", "must be greater than min")
        if levels < 2:
            raise InvalidParameter("levels", "must be at least 2")
        if levels > 256:
            raise InvalidParameter("levels", "must be no more than 256 to fit in u8")
        self._min, self._max, self._levels = mn, mx, int(levels)
        with np.errstate(over="ignore"):
            self._step = np.float32(mx - mn) / np.float32(levels - 1)  # sq.rs:94, all f32
        self._engine = engine

    min = property(lambda self: float(self._min))
    max = property(lambda self: float(self._max))
    levels = property(lambda self: self._levels)
    step = property(lambda self: float(self._step))

    def quantize(self, values):
        eng = _same_device(self._engine or default_engine(values), values)
        p, keep, _ = _in(values, np.float32)
        po, out = _out_like(values, tuple(keep.shape), np.uint8)
        eng.check(eng.lib.vqb_sq_quantize(eng.h, p, int(np.prod(keep.shape)), float(self._min), float(self._max),
                                          float(self._step), self._levels, po))
        _sync_if_torch(eng, values)
        return out

    def dequantize(self, codes):
        eng = _same_device(self._engine or default_engine(codes), codes)
        p, keep, _ = _in(codes, np.uint8)
        po, out = _out_like(codes, tuple(keep.shape), np.float32)
        eng.check(eng.lib.vqb_sq_dequantize(eng.h, p, int(np.prod(keep.shape)), float(self._min),
                                            float(self._step), po))
        _sync_if_torch(eng, codes)
        return out

    def __repr__(self):
        return f"ScalarQuantizer(min={self.min}, max={self.max}, levels={self._levels}, step={self.step})"


def _dequantize_f16(eng: Engine, q):
    p, keep, _ = _in(q, np.float16)
    po, out = _out_like(q, tuple(keep.shape), np.float32)
    eng.check(eng.lib.vqb_f16_dequantize(eng.h, p, int(np.prod(keep.shape)), po))
    _sync_if_torch(eng, q)
    return out


# --------------------------------------------------------------------------- PQ
def draw_init_indices(n: int, m: int, k: int, seed: int):
    """[m, k] initial rows + the per-subspace streams that later serve re-seeds
    (StdRng::seed_from_u64(seed + i), src/pq.rs:130; choose_multiple, src/core/vector.rs:413)."""
    streams = [IndexStream(seed, s) for s in range(m)]
    init = np.empty((m, k), np.uint64)
    for s in range(m):
        init[s] = streams[s].choose_multiple(n, k)
    return init, streams


class ProductQuantizer:
    """pyvq.ProductQuantizer (pyvq/src/pq.rs:49-87) over src/pq.rs:83-209.

    Extra keyword-only arguments (not in the reference): ``engine``; ``update`` ("ordered" =
    the reference's summation order, "fast" = segmented sums); ``assign`` ("auto", "exact",
    "tensor"); ``init_idx`` / ``reseed`` to supply the index stream explicitly; ``dist`` =
    a vq_b200.dist.RowShard describing this rank's share of a row-sharded training set.
    """

    def __init__(self, training_data, num_subspaces: int, num_centroids: int, max_iters: int = 10,
                 distance: Distance | None = None, seed: int = 42, *, engine: Engine | None = None,
                 update: str = "ordered", assign: str = "auto", init_idx=None, reseed=None, dist=None):
        shape = tuple(training_data.shape)
        if len(shape) != 2:
            raise ValueError("training_data must be a 2D array")
        n_local, dim = shape
        n = dist.n_global if dist is not None else n_local
        if n == 0:
            raise ValueError("Training data cannot be empty")  # pyvq/src/pq.rs:60-62
        m, k = int(num_subspaces), int(num_centroids)
        if m == 0:
            # the reference evaluates `dim % 0` and panics (pq.rs:112); surfaced here as a parameter error
            raise InvalidParameter("m", "must be greater than 0")
        if dim < m:
            raise InvalidParameter("m", f"must be at most the data dimension ({dim})")
        if dim % m != 0:
            raise InvalidParameter("m", f"dimension ({dim}) must be divisible by m")
        if k == 0:
            raise InvalidParameter("k", "must be greater than 0")
        if n < k:
            raise InvalidParameter("k", f"not enough data points ({n}) for {k} clusters")
        self._m, self._k, self._dim, self._sub_dim = m, k, dim, dim // m
        self._distance = distance or Distance.euclidean()  # pyvq/src/pq.rs:73-75
        self._engine = eng = _same_device(engine or default_engine(training_data), training_data)
        self._handle = None

        if init_idx is None:
            init_idx, streams = draw_init_indices(n, m, k, int(seed))
            if reseed is None:
                reseed = lambda s: streams[s].choose(n)
        init_idx = np.ascontiguousarray(init_idx, dtype=np.uint64).reshape(m * k)
        self.init_idx = init_idx.reshape(m, k).copy()

        opts = _lib.TrainOpts()
        opts.struct_size = C.sizeof(_lib.TrainOpts)
        opts.update_mode = {"ordered": _lib.UPDATE_ORDERED, "fast": _lib.UPDATE_FAST}[update]
        opts.assign_mode = {"auto": _lib.ASSIGN_AUTO, "exact": _lib.ASSIGN_EXACT, "tensor": _lib.ASSIGN_TENSOR}[assign]
        self._reseed_cb = _lib.RESEED_FN(lambda user, s: int(reseed(int(s))) if reseed else 0)
        if reseed is not None:
            opts.reseed = self._reseed_cb
        self._allreduce_cb = None
        if dist is not None:
            opts.row_offset, opts.n_global = dist.row_offset, dist.n_global
            if getattr(dist, "use_comm", False):   # the engine's own NCCL communicator (vq_b200.dist.init_comm)
                opts.flags = _lib.TRAIN_USE_COMM
            else:
                self._allreduce_cb = dist.allreduce_callback()
                opts.allreduce = self._allreduce_cb
        px, keep, _ = _in(training_data, np.float32)
        cb = np.empty((m, k, self._sub_dim), np.float32)
        iters = np.zeros(m, np.uint32)
        eng.check(eng.lib.vqb_pq_train(eng.h, px, n_local, dim, m, k, int(max_iters), init_idx.ctypes.data,
                                       C.byref(opts), cb.ctypes.data, iters.ctypes.data))
        self.iters_run = iters
        self._codebooks = cb
        self._make_handle()

    @classmethod
    def from_codebooks(cls, codebooks, distance: Distance | None = None, engine: Engine | None = None):
        """Wraps existing codebooks [m, k, sub_dim] without training."""
        self = cls.__new__(cls)
        cb = np.ascontiguousarray(codebooks, dtype=np.float32)
        self._m, self._k, self._sub_dim = cb.shape
        self._dim = self._m * self._sub_dim
        self._distance = distance or Distance.euclidean()
        self._engine = engine or default_engine()
        self._codebooks = cb
        self.iters_run = np.zeros(self._m, np.uint32)
        self.init_idx = None
        self._handle = None
        self._make_handle()
        return self

    def _make_handle(self):
        eng = self._engine
        h = C.c_void_p()
        eng.check(eng.lib.vqb_pq_create(eng.h, self._codebooks.ctypes.data, self._m, self._k, self._sub_dim,
                                        self._distance.id, C.byref(h)))
        self._handle = h

    def __del__(self):  # pragma: no cover
        try:
            if self._handle and self._engine.h:
                self._engine.lib.vqb_pq_destroy(self._handle)
        except Exception:
            pass

    num_subspaces = property(lambda self: self._m)
    sub_dim = property(lambda self: self._sub_dim)
    dim = property(lambda self: self._dim)
    num_centroids = property(lambda self: self._k)
    codebooks = property(lambda self: self._codebooks)

    def distance_metric(self) -> str:
        return self._distance.name()

    # ---- reference API: one vector -------------------------------------------------------------
    def quantize(self, vector):
        v = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        if v.size != self._dim:
            raise DimensionMismatch(self._dim, v.size)  # pq.rs:168-174
        return self.quantize_batch(v[None, :])[0]

    def dequantize(self, codes):
        q = np.ascontiguousarray(codes, dtype=np.float16).reshape(-1)
        if q.size != self._dim:
            raise DimensionMismatch(self._dim, q.size)  # pq.rs:202-207
        return _dequantize_f16(self._engine, q)

    # ---- batch API ---------------------------------------------------------------------------
    def _encode(self, x, want_codes, want_recon, assign="auto"):
        eng = _same_device(self._engine, x)
        px, keep, _ = _in(x, np.float32)
        if keep.ndim != 2 or keep.shape[1] != self._dim:
            raise DimensionMismatch(self._dim, keep.shape[-1] if keep.ndim else 0)
        n = keep.shape[0]
        cdt = np.uint8 if self._k <= 256 else (np.uint16 if self._k <= 65536 else np.uint32)
        pc, codes = _out_like(x, (n, self._m), cdt) if want_codes else (None, None)
        pr, recon = _out_like(x, (n, self._dim), np.float16) if want_recon else (None, None)
        mode = {"auto": _lib.ASSIGN_AUTO, "exact": _lib.ASSIGN_EXACT, "tensor": _lib.ASSIGN_TENSOR}[assign]
        eng.check(eng.lib.vqb_pq_encode(self._handle, px, n, mode, pc, np.dtype(cdt).itemsize, pr))
        _sync_if_torch(eng, x)
        return codes, recon

    def quantize_batch(self, x, assign="auto"):
        """[n, dim] f32 -> [n, dim] f16: the reference's output format, one row per vector."""
        return self._encode(x, False, True, assign)[1]

    def encode(self, x, assign="auto"):
        """[n, dim] f32 -> [n, m] code indices."""
        return self._encode(x, True, False, assign)[0]

    def encode_with_recon(self, x, assign="auto"):
        return self._encode(x, True, True, assign)

    def decode(self, codes):
        eng = _same_device(self._engine, codes)
        cdt = np.uint8 if self._k <= 256 else (np.uint16 if self._k <= 65536 else np.uint32)
        pc, keep, _ = _in(codes, cdt)
        n = keep.shape[0]
        po, out = _out_like(codes, (n, self._dim), np.float32)
        eng.check(eng.lib.vqb_pq_decode(self._handle, pc, np.dtype(cdt).itemsize, n, po))
        _sync_if_torch(eng, codes)
        return out

    def __repr__(self):
        return (f"ProductQuantizer(dim={self._dim}, num_subspaces={self._m}, sub_dim={self._sub_dim}, "
                f"distance={self._distance.name()})")


# --------------------------------------------------------------------------- TSVQ
class TSVQ:
    """pyvq.TSVQ (pyvq/src/tsvq.rs:42-70) over src/tsvq.rs:195-265."""

    def __init__(self, training_data, max_depth: int, distance: Distance | None = None,
                 engine: Engine | None = None):
        shape = tuple(training_data.shape)
        if len(shape) != 2:
            raise ValueError("training_data must be a 2D array")
        if shape[0] == 0:
            raise ValueError("Training data cannot be empty")
        self._dim = shape[1]
        self._distance = distance or Distance.euclidean()
        self._engine = eng = _same_device(engine or default_engine(training_data), training_data)
        px, keep, _ = _in(training_data, np.float32)
        h = C.c_void_p()
        eng.check(eng.lib.vqb_tsvq_train(eng.h, px, shape[0], shape[1], int(max_depth), self._distance.id, C.byref(h)))
        self._handle = h

    @classmethod
    def from_tree(cls, centroids, left, right, distance: Distance | None = None, engine: Engine | None = None):
        self = cls.__new__(cls)
        cent = np.ascontiguousarray(centroids, dtype=np.float32)
        l = np.ascontiguousarray(left, dtype=np.int32)
        r = np.ascontiguousarray(right, dtype=np.int32)
        self._dim = cent.shape[1]
        self._distance = distance or Distance.euclidean()
        self._engine = eng = engine or default_engine()
        h = C.c_void_p()
        eng.check(eng.lib.vqb_tsvq_create(eng.h, cent.ctypes.data, l.ctypes.data, r.ctypes.data, cent.shape[0],
                                          cent.shape[1], self._distance.id, C.byref(h)))
        self._handle = h
        return self

    def __del__(self):  # pragma: no cover
        try:
            if self._handle and self._engine.h:
                self._engine.lib.vqb_tsvq_destroy(self._handle)
        except Exception:
            pass

    dim = property(lambda self: self._dim)

    def distance_metric(self) -> str:
        return self._distance.name()

    def tree(self) -> dict:
        """Breadth-first node arrays: centroids, left, right, split_dim, median, count."""
        eng = self._engine
        nn, dd = C.c_size_t(), C.c_size_t()
        eng.check(eng.lib.vqb_tsvq_num_nodes(self._handle, C.byref(nn), C.byref(dd)))
        n = nn.value
        cent = np.empty((n, self._dim), np.float32)
        left = np.empty(n, np.int32); right = np.empty(n, np.int32); sd = np.empty(n, np.int32)
        med = np.empty(n, np.float32); cnt = np.empty(n, np.uint64)
        eng.check(eng.lib.vqb_tsvq_export(self._handle, cent.ctypes.data, left.ctypes.data, right.ctypes.data,
                                          sd.ctypes.data, med.ctypes.data, cnt.ctypes.data))
        return dict(centroids=cent, left=left, right=right, split_dim=sd, median=med, count=cnt)

    def _encode(self, x, want_leaf, want_recon):
        eng = _same_device(self._engine, x)
        px, keep, _ = _in(x, np.float32)
        if keep.ndim != 2 or keep.shape[1] != self._dim:
            raise DimensionMismatch(self._dim, keep.shape[-1] if keep.ndim else 0)
        n = keep.shape[0]
        pl, leaf = _out_like(x, (n,), np.uint32) if want_leaf else (None, None)
        pr, recon = _out_like(x, (n, self._dim), np.float16) if want_recon else (None, None)
        eng.check(eng.lib.vqb_tsvq_encode(self._handle, px, n, pl, pr))
        _sync_if_torch(eng, x)
        return leaf, recon

    def quantize(self, vector):
        v = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        if v.size != self._dim:
            raise DimensionMismatch(self._dim, v.size)  # tsvq.rs:240-245
        return self._encode(v[None, :], False, True)[1][0]

    def quantize_batch(self, x):
        return self._encode(x, False, True)[1]

    def encode(self, x):
        """[n, dim] f32 -> [n] leaf node ids (breadth-first numbering)."""
        return self._encode(x, True, False)[0]

    def dequantize(self, codes):
        q = np.ascontiguousarray(codes, dtype=np.float16).reshape(-1)
        if q.size != self._dim:
            raise DimensionMismatch(self._dim, q.size)  # tsvq.rs:258-263
        return _dequantize_f16(self._engine, q)

    def __repr__(self):
        return f"TSVQ(dim={self._dim}, distance={self._distance.name()})"
