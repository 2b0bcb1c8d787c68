"""torchrun --nproc-per-node N tools/check_row_shard.py : row-sharded PQ training on N GPUs against single-GPU
training and the CPU oracle.  Run by tests/test_gpu_multirank.py; prints ROW-SHARD OK on rank 0."""
import os, sys
import numpy as np, torch, torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200 as vq
from vq_b200.dist import RowShard, init_comm, shard_bounds

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = td.get_rank(), td.get_world_size()
eng = vq.Engine(local)
init_comm(eng)


def rel_diff(a, b):
    return max(float(np.linalg.norm(a[s] - b[s]) / max(np.linalg.norm(b[s]), 1e-30)) for s in range(a.shape[0]))


def same_on_all_ranks(cb):
    t = torch.from_numpy(cb.view(np.int32).astype(np.int64)).cuda()
    lo, hi = t.clone(), t.clone()
    td.all_reduce(lo, op=td.ReduceOp.MIN); td.all_reduce(hi, op=td.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


for (n, dim, m, k, iters, update, seed) in [(60_000, 64, 8, 256, 6, "fast", 11), (20_001, 48, 3, 64, 5, "ordered", 12),
                                            (9_000, 32, 4, 300, 4, "fast", 13)]:
    rng = np.random.default_rng(seed)                      # same data on every rank
    centers = rng.standard_normal((1024, dim)).astype(np.float32)   # SURVEY 8d's separated mixture: k-means stays well conditioned
    x = (centers[rng.integers(0, 1024, n)] + 0.25 * rng.standard_normal((n, dim))).astype(np.float32)
    init, _ = vq.draw_init_indices(n, m, k, 42)
    init[0, 1] = init[0, 0]                                # duplicate seed row -> an empty cluster -> re-seeding
    reseed_row = n - 7                                      # owned by the last rank
    b, e = shard_bounds(n, rank, world)
    xs = torch.from_numpy(x[b:e]).cuda()
    dbg = lambda msg: print(f"[rank {rank}] n={n} {update}: {msg}", flush=True) if os.environ.get("ROW_SHARD_VERBOSE") else None
    dbg("start")
    lib = vq.ProductQuantizer(xs, m, k, iters, vq.Distance.euclidean(), engine=eng, init_idx=init, update=update,
                              reseed=lambda s: reseed_row, dist=RowShard.for_rank(n, use_comm=True))
    dbg("library-communicator training done")
    cbk = vq.ProductQuantizer(xs, m, k, iters, vq.Distance.euclidean(), engine=eng, init_idx=init, update=update,
                              reseed=lambda s: reseed_row, dist=RowShard.for_rank(n))
    dbg("callback training done")
    assert np.array_equal(lib.codebooks.view(np.uint32), cbk.codebooks.view(np.uint32)), "library communicator != host callback"
    assert np.array_equal(lib.iters_run, cbk.iters_run)
    assert same_on_all_ranks(lib.codebooks), "replicated codebooks differ between ranks"
    single = vq.ProductQuantizer(torch.from_numpy(x).cuda(), m, k, iters, vq.Distance.euclidean(), engine=eng, init_idx=init,
                                 update=update, reseed=lambda s: reseed_row)
    dbg("single-GPU training done")
    r1 = rel_diff(lib.codebooks, single.codebooks)
    assert r1 <= 1e-4, ("vs single GPU", r1)
    if rank == 0:
        from oracle import oracle as O
        want, it = O.get().pq_train(x, m, k, iters, init, reseed=lambda s: reseed_row)
        r2 = rel_diff(lib.codebooks, want)
        assert r2 <= 1e-4, ("vs oracle", r2)
        assert np.array_equal(lib.iters_run, it), (lib.iters_run, it)
        print(f"n={n} dim={dim} m={m} k={k} {update}: {world} ranks, library comm == callback bit for bit; "
              f"vs single GPU {r1:.2e}, vs oracle {r2:.2e} (bar 1e-4)", flush=True)
td.barrier()
if rank == 0:
    print("ROW-SHARD OK", flush=True)
td.destroy_process_group()
