"""Row-sharded multi-GPU plumbing: one process per GPU, torch.distributed for the exchange.

PQ training shards the ROWS of the training set over the ranks; codebooks are replicated.
Each k-means iteration exchanges ONE fused buffer [sums | count_lo | count_hi] (f32) with a
sum all-reduce (NCCL over NVLink on GPUs), after which every rank runs the identical
finalize step, so the replicated codebooks stay bit-identical across ranks.  Encoding and
the element-wise codecs shard rows with no collective at all.

The engine calls back into `RowShard.allreduce_callback()` with a raw pointer; this module
wraps that pointer as a tensor (CUDA: __cuda_array_interface__, CPU: ctypes buffer) and calls
torch.distributed.all_reduce on the stream the engine is using.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

try:
    import torch
    import torch.distributed as td
except Exception:  # pragma: no cover
    torch = None
    td = None


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced row range [begin, end) of `rank` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class _CudaView:
    """Minimal __cuda_array_interface__ carrier for a raw device pointer."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False),
                                         "version": 3, "strides": None}


def tensor_from_pointer(ptr: int, count: int, cuda: bool):
    if cuda:
        return torch.as_tensor(_CudaView(ptr, count), device="cuda")
    buf = (C.c_float * count).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=np.float32, count=count))


def pointer_is_cuda(ptr: int) -> bool:
    if torch is None or not torch.cuda.is_available():
        return False
    try:
        from cuda.bindings import runtime as cudart  # cuda-python
        err, attr = cudart.cudaPointerGetAttributes(ptr)
        return int(attr.type) in (2, 3)
    except Exception:
        return True  # engine buffers are device memory whenever a GPU is present


def init_comm(engine, group=None):
    """Gives `engine` the library-owned NCCL communicator (include/vqb200.h, multi-GPU section): rank 0 draws the
    128-byte id, torch.distributed carries it to the other ranks (any transport would do), every rank joins.
    Afterwards RowShard(..., use_comm=True) trains with the all-reduce issued by the library itself -- no Python in
    the loop."""
    rank, world = td.get_rank(group), td.get_world_size(group)
    buf = (C.c_char * _lib.COMM_ID_BYTES)()
    if rank == 0:
        engine.check(engine.lib.vqb_comm_unique_id(buf))
    box = [bytes(buf.raw)]
    td.broadcast_object_list(box, src=td.get_global_rank(group, 0) if group is not None else 0, group=group)
    ident = (C.c_char * _lib.COMM_ID_BYTES).from_buffer_copy(box[0])
    engine.check(engine.lib.vqb_comm_init_rank(engine.h, ident, rank, world))
    return rank, world


@dataclass
class RowShard:
    """This rank's share of a row-sharded training set."""
    row_offset: int
    n_global: int
    group: object = None  # torch.distributed process group (None = default group)
    use_comm: bool = False  # True: the engine's own communicator (init_comm) instead of the torch.distributed callback

    @classmethod
    def for_rank(cls, n_global: int, rank: int | None = None, world: int | None = None, group=None, use_comm: bool = False):
        rank = td.get_rank(group) if rank is None else rank
        world = td.get_world_size(group) if world is None else world
        b, _ = shard_bounds(n_global, rank, world)
        return cls(b, n_global, group, use_comm)

    def all_reduce_pointer(self, ptr: int, count: int, stream: int | None, cuda: bool | None = None) -> int:
        if count == 0:
            return 0
        cuda = pointer_is_cuda(ptr) if cuda is None else cuda
        t = tensor_from_pointer(ptr, count, cuda)
        if cuda and stream:
            with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                td.all_reduce(t, op=td.ReduceOp.SUM, group=self.group)
        else:
            td.all_reduce(t, op=td.ReduceOp.SUM, group=self.group)
        return 0

    def allreduce_callback(self):
        def cb(user, buf, count, stream):
            try:
                return self.all_reduce_pointer(int(buf), int(count), int(stream) if stream else None)
            except Exception as e:  # never unwind through C
                import sys
                print(f"[vq_b200.dist] all-reduce failed: {e!r}", file=sys.stderr)
                return 1
        return _lib.ALLREDUCE_FN(cb)


# ---- host-side statement of the exchange protocol (what the CUDA pack/finalize kernels do) ----
def pack_partial(sums: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """[sums | count_lo | count_hi] as f32: counts travel as two exact 16-bit halves so that a
    float sum over <= 256 ranks stays exact (vq_b200/csrc/pq_train.cu k_pack)."""
    counts = counts.astype(np.uint32).reshape(-1)
    return np.concatenate([sums.astype(np.float32).reshape(-1),
                           (counts & 0xFFFF).astype(np.float32), (counts >> 16).astype(np.float32)])


def unpack_reduced(buf: np.ndarray, n_sums: int) -> tuple[np.ndarray, np.ndarray]:
    n_c = (buf.size - n_sums) // 2
    lo = buf[n_sums:n_sums + n_c].astype(np.uint64)
    hi = buf[n_sums + n_c:].astype(np.uint64)
    return buf[:n_sums], hi * 65536 + lo


# ---- shard by SUBSPACE: the collective-free cross-check of SURVEY.md 8(e) ---------------------------------------
# Subspaces are fully independent (src/pq.rs:121-132: one lbg_quantize per subspace with seed + i), so rank r can train
# subspaces [s0, s1) on ALL rows of its column slice with the ordered update and the result is bit-identical to a
# single-GPU run -- no all-reduce, the reference's exact summation order.  It needs every rank to hold all rows, which is
# why row sharding (above) is the production path and this one the cross-check.
def subspace_bounds(m: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced subspace range [s0, s1) of `rank`."""
    return shard_bounds(m, rank, world)


def gather_codebooks(local: np.ndarray, m: int, group=None) -> np.ndarray:
    """All ranks contribute their [m_local, k, sub_dim] block (rank order = subspace order); returns [m, k, sub_dim].
    Blocks travel as objects, not through a sum, so signed zeros and NaNs arrive untouched."""
    world = td.get_world_size(group)
    parts = [None] * world
    td.all_gather_object(parts, np.ascontiguousarray(local, dtype=np.float32), group=group)
    parts = [p for p in parts if p is not None and p.shape[0] > 0]
    full = np.concatenate(parts, axis=0) if parts else np.empty((0,) + tuple(local.shape[1:]), np.float32)
    if full.shape[0] != m:
        raise RuntimeError(f"gathered {full.shape[0]} subspaces, expected {m}")
    return full


def train_pq_by_subspace(training_data, num_subspaces: int, num_centroids: int, max_iters: int = 10, distance=None,
                         seed: int = 42, *, engine=None, group=None, assign: str = "auto"):
    """ProductQuantizer trained with the subspaces split over the ranks of `group` (every rank passes the SAME full
    training set).  Index streams are drawn exactly as a single process draws them (seed + global subspace id), so
    the returned quantizer equals `ProductQuantizer(training_data, m, k, max_iters, distance, seed)` bit for bit."""
    from . import api
    rank, world = td.get_rank(group), td.get_world_size(group)
    n, dim = tuple(training_data.shape)
    m, k = int(num_subspaces), int(num_centroids)
    if m == 0 or dim % m:
        raise api.InvalidParameter("m", f"dimension ({dim}) must be divisible by m")
    d = dim // m
    s0, s1 = subspace_bounds(m, rank, world)
    init, streams = api.draw_init_indices(n, m, k, int(seed))
    if s1 > s0:
        cols = training_data[:, s0 * d:s1 * d]
        cols = cols.contiguous() if hasattr(cols, "contiguous") else np.ascontiguousarray(cols)
        local = api.ProductQuantizer(cols, s1 - s0, k, max_iters, distance, seed, engine=engine, update="ordered",
                                     assign=assign, init_idx=init[s0:s1],
                                     reseed=lambda s: streams[s0 + s].choose(n))
        block, iters = local.codebooks, local.iters_run
    else:  # more ranks than subspaces
        block, iters = np.empty((0, k, d), np.float32), np.empty(0, np.uint32)
    full = gather_codebooks(block, m, group)
    its = [None] * world
    td.all_gather_object(its, np.asarray(iters, dtype=np.uint32), group=group)
    pq = api.ProductQuantizer.from_codebooks(full, distance, engine=engine)
    pq.iters_run = np.concatenate([np.asarray(i, dtype=np.uint32) for i in its])
    pq.init_idx = init
    return pq
