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


@dataclass
class RowShard:
    """This rank's share of a row-sharded training set."""
    row_offset: int
    n_global: int
    group: object = None  # torch.distributed process group (None = default group)

    @classmethod
    def for_rank(cls, n_global: int, rank: int | None = None, world: int | None = None, group=None):
        rank = td.get_rank(group) if rank is None else rank
        world = td.get_world_size(group) if world is None else world
        b, _ = shard_bounds(n_global, rank, world)
        return cls(b, n_global, group)

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
