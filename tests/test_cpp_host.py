"""include/vq.hpp -- the C++ host mirror of the vq crate API -- built against libvqb200.so and driven through
tests/cpp/test_vq_hpp.cpp.  CPU: validation order and error kinds (src/pq.rs:91-117, src/core/vector.rs:396-410,
src/bq.rs:55-75, src/sq.rs:63-101), agreement of its index streams with vq_b200/rand09.py, and the loud failure when
no sm_100 GPU is present.  GPU: the C++ mirror and the Python mirror give identical results on the same data."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    from vq_b200.build import build_lib
    lib = build_lib()
    libdir = os.path.dirname(lib)
    exe = str(tmp_path_factory.mktemp("cpp") / "test_vq_hpp")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_vq_hpp.cpp"),
           "-o", exe, "-L", libdir, "-lvqb200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def run(exe, *args):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=300)
    return r.returncode, r.stdout


def test_cpp_validation_order_and_error_kinds(driver):
    rc, out = run(driver, "host")
    assert rc == 0 and out.strip().endswith("host ok"), out


@pytest.mark.parametrize("seed,n,k", [(42, 1000, 16), (43, 100_000, 256), (7, 300, 256), (1, 1_000_000, 256)])
def test_cpp_index_stream_equals_python_restatement(driver, seed, n, k):
    from vq_b200.rand09 import StdRng
    rc, out = run(driver, "stream", seed, n, k)
    assert rc == 0, out
    lines = out.strip().splitlines()
    rng = StdRng.seed_from_u64(seed)
    assert [int(v) for v in lines[0].split()] == [int(v) for v in rng.sample_indices(n, k)]
    assert [int(v) for v in lines[1].split()] == [rng.random_range_usize(n) for _ in range(5)]


def test_cpp_engine_fails_loudly_without_a_gpu(driver):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc, out = run(driver, "nogpu")
    assert rc == 0 and out.strip(), out      # FfiError with a message; no CPU fallback


def make_rows(n, dim, seed):
    """tests/cpp/test_vq_hpp.cpp::make_rows (32-bit LCG), value for value."""
    out = np.empty((n, dim), F)
    s = np.uint32(seed)
    flat = out.reshape(-1)
    with np.errstate(over="ignore"):
        for i in range(flat.size):
            s = np.uint32(s * np.uint32(1664525) + np.uint32(1013904223))
            hi = np.int32(np.uint32(s >> np.uint32(8)))
            flat[i] = F(F(F(hi) / F(8388608.0)) - F(1.0)) + F(int((s >> np.uint32(3)) & np.uint32(7)))
    return out


@pytest.mark.gpu
def test_cpp_mirror_equals_python_mirror(driver):
    import vq_b200 as vq
    rc, out = run(driver, "gpu")
    assert rc == 0, out
    got = {ln.split()[0]: ln.split()[1:] for ln in out.strip().splitlines()}
    rows = make_rows(600, 16, 7)
    pq = vq.ProductQuantizer(rows, 2, 16, 4, vq.Distance.cosine(), 42)
    assert [int(h, 16) for h in got["cb"]] == pq.codebooks.reshape(-1).view(np.uint32).tolist()
    assert [int(h, 16) for h in got["q"]] == pq.quantize(rows[5]).view(np.uint16).tolist()
    t = vq.TSVQ(rows, 3, vq.Distance.euclidean())
    assert [int(h, 16) for h in got["t"]] == t.quantize(rows[9]).view(np.uint16).tolist()
    assert [int(v) for v in got["b"]] == vq.BinaryQuantizer(0.5, 0, 1).quantize(rows[0]).tolist()
    assert [int(v) for v in got["s"]] == vq.ScalarQuantizer(-1.0, 8.0, 256).quantize(rows[0]).tolist()
    assert F(float(got["d"][0])) == F(vq.Distance.manhattan().compute(rows[0], rows[1]))
    # Engine::comm_init (library NCCL communicator, one rank) + the RowShard constructor: same codebooks
    assert got["comm"] == ["0", "1", "1"]
