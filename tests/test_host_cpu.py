"""CPU-only tests: host logic of the package, the C-ABI library's export table, the index
stream, and the multi-rank exchange protocol over gloo (world_size 2).  No compute call is
made on the library here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


@pytest.fixture(scope="module")
def built_lib():
    from vq_b200.build import build_lib
    return build_lib()


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "vqb200.h")).read()
    declared = set(re.findall(r"\b(vqb_[a-z0-9_]+)\s*\(", hdr)) - {"vqb_reseed_fn", "vqb_allreduce_fn"}
    declared = {d for d in declared if not d.endswith("_fn")}
    lib = C.CDLL(built_lib)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from vq_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    assert _lib.load().vqb_backend_name().decode().startswith("vqb200")


def test_library_contains_sm100a_code(built_lib):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import vq_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq_b200.Engine(0)
    h = C.c_void_p()
    from vq_b200 import _lib
    assert _lib.load().vqb_ctx_create(0, C.byref(h)) == _lib.ERR_UNSUPPORTED_DEVICE


def test_product_path_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "vq_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "vq_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


# ---------------------------------------------------------------- constructor validation (no GPU needed)
def test_bq_sq_validation_messages():
    import vq_b200 as vq
    # src/bq.rs:55-75
    with pytest.raises(ValueError, match="Invalid parameter 'threshold': must be finite"):
        vq.BinaryQuantizer(float("nan"))
    with pytest.raises(ValueError, match="low must be less than high"):
        vq.BinaryQuantizer(0.0, 5, 5)
    b = vq.BinaryQuantizer(0.5, 0, 1)
    assert (b.threshold, b.low, b.high) == (0.5, 0, 1) and "BinaryQuantizer" in repr(b)
    # src/sq.rs:63-101
    for args, msg in [((float("inf"), 1.0, 4), "'min'"), ((0.0, float("nan"), 4), "'max'"),
                      ((1.0, 1.0, 4), "must be greater than min"), ((0.0, 1.0, 1), "at least 2"),
                      ((0.0, 1.0, 257), "no more than 256")]:
        with pytest.raises(ValueError, match=msg):
            vq.ScalarQuantizer(*args)
    s = vq.ScalarQuantizer(-1.0, 1.0, 256)
    assert s.levels == 256 and s.step == float(F(2.0) / F(255))


def test_pq_tsvq_validation_order_and_messages():
    import vq_b200 as vq
    x = np.zeros((50, 7), F)
    # pyvq/tests/test_pq.py:78-90, src/pq.rs:91-117, src/core/vector.rs:396-410
    with pytest.raises(ValueError, match="empty"):
        vq.ProductQuantizer(np.zeros((0, 8), F), 2, 4)
    with pytest.raises(ValueError, match="must be at most the data dimension"):
        vq.ProductQuantizer(x, 8, 4)
    with pytest.raises(ValueError, match="must be divisible by m"):
        vq.ProductQuantizer(x, 2, 4)
    with pytest.raises(ValueError, match="'k': must be greater than 0"):
        vq.ProductQuantizer(np.zeros((50, 8), F), 2, 0)
    with pytest.raises(ValueError, match=r"not enough data points \(50\) for 64 clusters"):
        vq.ProductQuantizer(np.zeros((50, 8), F), 2, 64)
    with pytest.raises(ValueError, match="empty"):
        vq.TSVQ(np.zeros((0, 8), F), 3)
    with pytest.raises(ValueError, match="Invalid distance metric"):
        vq.Distance("chebyshev")  # not a variant of the reference's Distance (src/core/distance.rs:8-17)
    assert vq.Distance("SquaredEuclidean").name() == "squared_euclidean"
    assert repr(vq.Distance.cosine()) == "Distance(metric=cosine)"
    # the Chebyshev extension is reachable through its own constructor only, with the header's id
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "vqb200.h")).read()
    assert vq.Distance.chebyshev().id == int(re.search(r"#define VQB_CHEBYSHEV\s+(\d+)", hdr).group(1)) == 5
    assert [vq.Distance(n).id for n in ("squared_euclidean", "euclidean", "manhattan", "cosine")] == [0, 1, 2, 3]


# ---------------------------------------------------------------- index stream
def test_chacha_block_rfc7539_vector():
    from vq_b200.rand09 import _chacha_block
    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    # RFC 7539 section 2.3.2: counter 1, nonce 00:00:00:09 00:00:00:4a 00:00:00:00, 20 rounds
    out = _chacha_block(key, 1 | (0x09000000 << 32), 0x4A000000, rounds=20)
    assert out[:4] == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3]
    assert out[12:] == [0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]


def test_index_stream_properties():
    from vq_b200.rand09 import IndexStream
    import vq_b200 as vq
    for n, k in [(100, 4), (100, 50), (5000, 256), (100_000, 256), (1_000_000, 256), (300, 256)]:
        a = IndexStream(42, 3).choose_multiple(n, k)
        b = IndexStream(42, 3).choose_multiple(n, k)
        assert a == b and len(set(a)) == k and all(0 <= v < n for v in a)
        assert a != IndexStream(43, 3).choose_multiple(n, k)
    st = IndexStream(7, 0)
    draws = [st.choose(10) for _ in range(2000)]
    assert set(draws) == set(range(10))
    init, streams = vq.draw_init_indices(1000, 4, 16, 42)
    assert init.shape == (4, 16) and init.dtype == np.uint64
    # seed + i (src/pq.rs:130): subspace 1 of seed 42 == subspace 0 of seed 43
    init2, _ = vq.draw_init_indices(1000, 4, 16, 43)
    assert np.array_equal(init[1], init2[0])


# ---------------------------------------------------------------- row sharding + exchange protocol
def test_shard_bounds():
    from vq_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 1_000_003):
        for w in (1, 2, 3, 8):
            parts = [shard_bounds(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_counts_exact():
    from vq_b200.dist import pack_partial, unpack_reduced
    sums = np.arange(12, dtype=F)
    counts = np.array([0, 1, 65535, 65536, 70000, 2**31 + 5], np.uint64)
    total = pack_partial(sums, counts) + pack_partial(sums, counts)  # "all-reduce" of two ranks
    s, c = unpack_reduced(total, 12)
    assert np.array_equal(s, 2 * sums) and np.array_equal(c, 2 * counts)


_WORKER = r'''
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as td
from vq_b200.dist import RowShard, shard_bounds, pack_partial, unpack_reduced
from oracle import oracle as O

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
td.init_process_group("gloo", rank=rank, world_size=world)
F = np.float32
orc = O.get()

# 1. the callback the engine would invoke: in-place sum of a raw float buffer across ranks
shard = RowShard.for_rank(1000, group=None)
cb = shard.allreduce_callback()
buf = (C.c_float * 5)(*[float(rank + 1)] * 5)
assert cb(None, C.cast(buf, C.c_void_p), 5, None) == 0
assert list(buf) == [float(sum(range(1, world + 1)))] * 5

# 2. row-sharded k-means iterations, stated with the exchange protocol of pq_train.cu:
#    local ordered sums + counts -> pack -> all-reduce -> identical finalize on every rank.
rng = np.random.default_rng(5)
n, dim, m, k = 3000, 16, 2, 32
centers = rng.standard_normal((64, dim)).astype(F)
x = (centers[rng.integers(0, 64, n)] + 0.25 * rng.standard_normal((n, dim))).astype(F)
d = dim // m
init = np.stack([rng.choice(n, k, replace=False) for _ in range(m)])
b, e = shard_bounds(n, rank, world)
xl = x[b:e]
cb_state = np.stack([x[init[s], s * d:(s + 1) * d] for s in range(m)]).astype(F)
ref_state = cb_state.copy()
for it in range(5):
    sums = np.zeros((m, k, d), F); counts = np.zeros((m, k), np.uint32)
    for s in range(m):
        _, assign, _, _ = orc.lbg_step(xl, s * d, d, cb_state[s])
        for i, j in enumerate(assign):          # ascending local row order, sequential f32 adds
            sums[s, j] = (sums[s, j] + xl[i, s * d:(s + 1) * d]).astype(F)
        counts[s] = np.bincount(assign, minlength=k)
    pack = np.ascontiguousarray(pack_partial(sums, counts))
    assert cb(None, C.c_void_p(pack.ctypes.data), pack.size, None) == 0
    rs, rc = unpack_reduced(pack, m * k * d)
    rs = rs.reshape(m, k, d)
    rc = rc.reshape(m, k)
    nz = rc > 0
    cb_state[nz] = (rs[nz] / rc[nz][:, None].astype(F)).astype(F)
    for s in range(m):                           # single-process oracle, same state
        ref_state[s], _, _, _ = orc.lbg_step(x, s * d, d, ref_state[s])
    # replicated state must be bit-identical on all ranks
    gathered = [torch.empty(cb_state.size, dtype=torch.float32) for _ in range(world)]
    td.all_gather(gathered, torch.from_numpy(cb_state.reshape(-1).copy()))
    assert all(torch.equal(gathered[0], g) for g in gathered)
    # row sharding changes the summation order: tolerance 1e-4 relative (BASELINE north_star)
    for s in range(m):
        rel = np.linalg.norm(cb_state[s] - ref_state[s]) / np.linalg.norm(ref_state[s])
        assert rel <= 1e-4, rel

# 3. shard by SUBSPACE (SURVEY 8e cross-check): every rank trains its subspaces on all rows with the reference's own
#    summation order; gathered blocks must equal the single-process result bit for bit, iteration counts included.
from vq_b200.dist import subspace_bounds, gather_codebooks
m3, k3, it3 = 5, 16, 4                            # 5 subspaces over 2 ranks: uneven split (3 + 2)
x3 = np.ascontiguousarray(x[:, :15])              # dim 15, sub_dim 3
init3 = np.stack([np.random.default_rng(100 + s).choice(n, k3, replace=False) for s in range(m3)]).astype(np.uint64)
want_cb, want_it = orc.pq_train(x3, m3, k3, it3, init3, reseed=lambda s: 0)
s0, s1 = subspace_bounds(m3, rank, world)
assert (s0, s1) == ((0, 3) if rank == 0 else (3, 5))
d3 = 3
blk, blk_it = orc.pq_train(np.ascontiguousarray(x3[:, s0 * d3:s1 * d3]), s1 - s0, k3, it3, init3[s0:s1], reseed=lambda s: 0)
full = gather_codebooks(blk, m3)
assert np.array_equal(full.view(np.uint32), want_cb.view(np.uint32))
its = [None] * world
td.all_gather_object(its, np.asarray(blk_it, dtype=np.uint32))
assert np.array_equal(np.concatenate(its), want_it)
td.destroy_process_group()
print("RANK_OK", rank)
'''


def test_gloo_two_rank_exchange_protocol(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2", OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {r}" in o, o[-3000:]


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` needs no GPU: it times the restated reference loop on the host cores and prints one
    JSON line with the same metric / unit / config keys as the CUDA arm plus impl, cpu_baseline and a zero-copy e2e."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-sample", "20000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "pq_encode_throughput" and line["unit"] == "Mvec/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mvec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["m"] == 96 and line["config"]["k"] == 256 and line["config"]["dim"] == 768
