#!/usr/bin/env python
"""Generates tests/golden/*.npz -- golden input/output vectors for the vq hot path.

Run in the BUILD container, where the reference checkout exists at /root/reference:

    python tests/golden/make_golden.py

Sources of truth, strongest first:
  hsdlib_distances.npz  outputs of the reference's OWN native code: hsdlib compiled verbatim from
                        /root/reference/external/hsdlib (oracle/Makefile -> oracle/_ref/libhsd_ref.so), called through
                        the same entry points the Rust crate binds (src/core/hsdlib_ffi.rs:38-66), with the backend
                        forced to scalar / AVX2 / AVX512F (hsd_set_manual_backend) -- lane-boundary dims of
                        external/hsdlib/tests/test_*.c (0,1,3,7,8,9,15,16,17) plus 100, 768, 1536.
  kats.npz              the known-answer values written in the reference's own tests (file:line in the arrays' names).
  pq_*.npz, tsvq_*.npz, codec_*.npz
                        outputs of the CPU restatement (oracle/vq_oracle.c) with distances routed through the real
                        hsdlib, on seeded inputs.  They freeze the checker so that the GPU parity tests on the GPU box
                        (where /root/reference does not exist) compare against files made next to the reference.
The fixtures are small (a few hundred KB) and committed; this script is the only thing that writes them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402


def mixture(n, dim, seed, comps=32, sigma=0.25):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((comps, dim)).astype(np.float32)
    x = centers[rng.integers(0, comps, n)] + np.float32(sigma) * rng.standard_normal((n, dim)).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def main():
    if not os.path.isdir("/root/reference/external/hsdlib/src"):
        raise SystemExit("the reference checkout is required to (re)generate the golden vectors")
    O.build(force=True)
    if not os.path.exists(O.HSD_PATH):
        raise SystemExit("oracle/_ref/libhsd_ref.so was not built")
    orc = O.Oracle(use_hsdlib=True)
    assert orc.hsd is not None

    # ---- 1. real hsdlib, every backend ------------------------------------------------------
    rng = np.random.default_rng(1234)
    dims = [0, 1, 3, 7, 8, 9, 15, 16, 17, 100, 768, 1536]
    out = {"dims": np.array(dims)}
    for d in dims:
        a = (rng.standard_normal((8, d)) * 3).astype(np.float32)
        b = (rng.standard_normal((8, d)) * 3).astype(np.float32)
        if d >= 3:
            b[1] = a[1]                      # identical pair
            a[2] = 0.0                       # zero vector vs non-zero (cosine zero rule)
            a[3] = 0.0; b[3] = 0.0           # both zero
            a[4] = -b[4]                     # anti-parallel
        out[f"a_{d}"] = a; out[f"b_{d}"] = b
        for backend in ("scalar", "avx2", "avx512f"):
            try:
                h = O.Hsdlib(backend)
            except Exception:
                continue
            res = np.zeros((3, 8), np.float32); st = np.zeros((3, 8), np.int32)
            for r in range(8):
                for q, fn in enumerate((h.sqeuclidean, h.manhattan, h.cosine)):
                    s, v = fn(a[r], b[r])
                    st[q, r] = s; res[q, r] = v
            out[f"val_{backend}_{d}"] = res; out[f"status_{backend}_{d}"] = st
    np.savez_compressed(os.path.join(HERE, "hsdlib_distances.npz"), **out)

    # ---- 2. KATs copied from the reference's tests (values only; no code) --------------------
    kats = {
        # src/core/distance.rs:131-166
        "distance_rs_131_a": np.array([1, 2, 3], np.float32), "distance_rs_131_b": np.array([4, 6, 8], np.float32),
        "distance_rs_131_sq_l2_l1": np.array([50.0, np.sqrt(np.float32(50.0)), 12.0], np.float32),
        # src/core/hsdlib_ffi.rs:169-206
        "hsdlib_ffi_rs_169_a": np.array([1, 2, 3], np.float32), "hsdlib_ffi_rs_169_b": np.array([4, 5, 6], np.float32),
        "hsdlib_ffi_rs_169_sq_l1": np.array([27.0, 9.0], np.float32),
        # external/hsdlib/tests/test_euclidean.c:14-19
        "test_euclidean_c_14_a": np.arange(1, 10, dtype=np.float32), "test_euclidean_c_14_b": np.arange(9, 0, -1).astype(np.float32),
        "test_euclidean_c_14_sq": np.array([240.0], np.float32),
        # src/sq.rs:13-20  (0,1,11): [0,.5,1] -> [0,5,10]
        "sq_rs_13_in": np.array([0.0, 0.5, 1.0], np.float32), "sq_rs_13_out": np.array([0, 5, 10], np.uint8),
        # pyvq/tests/test_sq.py:37-54  (-1,1,5)
        "test_sq_py_37_in": np.array([-1.2, -1.0, -0.8, -0.3, 0.0, 0.3, 0.6, 1.0, 1.2], np.float32),
        "test_sq_py_37_out": np.array([0, 0, 0, 1, 2, 3, 3, 4, 4], np.uint8),
        # src/bq.rs:126-144  thr 0: [-1,0,1,-.5,.5] -> [0,1,1,0,1]
        "bq_rs_126_in": np.array([-1.0, 0.0, 1.0, -0.5, 0.5], np.float32), "bq_rs_126_out": np.array([0, 1, 1, 0, 1], np.uint8),
        # src/core/vector.rs:527-538  mean of [1,2,3],[4,5,6],[7,8,9]
        "vector_rs_527_in": np.arange(1, 10, dtype=np.float32).reshape(3, 3), "vector_rs_527_mean": np.array([4, 5, 6], np.float32),
    }
    np.savez_compressed(os.path.join(HERE, "kats.npz"), **kats)

    # ---- 3. PQ: train (explicit index stream) + encode with every metric ---------------------
    for name, (n, dim, m, k, iters, seed) in {"pq_d8": (3000, 64, 8, 64, 6, 11), "pq_d16": (2000, 64, 4, 32, 5, 12),
                                               "pq_d5": (1500, 20, 4, 17, 4, 13)}.items():
        x = mixture(n, dim, seed)
        rs = np.random.default_rng(seed + 100)
        init = np.stack([rs.choice(n, k, replace=False) for _ in range(m)]).astype(np.uint64)
        reseed_rows = rs.integers(0, n, 4096).astype(np.uint64)
        pos = [0]

        def reseed(s, rows=reseed_rows, pos=pos):
            v = int(rows[pos[0] % rows.size]); pos[0] += 1
            return v
        cb, it = orc.pq_train(x, m, k, iters, init, reseed=reseed, threads=1)
        fx = dict(x=x, init_idx=init, reseed_rows=reseed_rows, reseeds_used=np.array([pos[0]]), m=np.array([m]), k=np.array([k]),
                  max_iters=np.array([iters]), codebooks=cb, iters_run=it)
        xq = mixture(500, dim, seed + 7)
        xq[0] = 0.0                                # zero query (cosine zero rule)
        xq[1] = cb[:, 3, :].reshape(-1)            # a query that IS a centroid in every subspace
        fx["xq"] = xq
        for metric in ("squared_euclidean", "euclidean", "manhattan", "cosine"):
            codes, recon = orc.pq_encode(cb, metric, xq, sem="hsdlib", want_recon=True, threads=1)
            fx[f"codes_{metric}"] = codes.astype(np.uint16)
            fx[f"recon_{metric}"] = recon.view(np.uint16)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **fx)

    # ---- 4. TSVQ -----------------------------------------------------------------------------
    for name, (n, dim, depth, seed) in {"tsvq_a": (1200, 24, 5, 21), "tsvq_b": (300, 33, 8, 22)}.items():
        x = mixture(n, dim, seed, comps=8)
        x[5] = x[6]                                # duplicates around a median
        tree = orc.tsvq_build(x, depth)
        fx = dict(x=x, depth=np.array([depth]), **{f"tree_{k_}": v for k_, v in tree.items()})
        xq = mixture(400, dim, seed + 3, comps=8)
        fx["xq"] = xq
        for metric in ("squared_euclidean", "euclidean", "manhattan", "cosine"):
            leaf, recon = orc.tsvq_encode(tree, metric, xq, sem="hsdlib", want_recon=True, threads=1)
            fx[f"leaf_{metric}"] = leaf
            fx[f"recon_{metric}"] = recon.view(np.uint16)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **fx)

    # ---- 5. BQ / SQ / f16 codecs incl. the edge values of tests/integration_tests.rs:283-321,476-565
    rng = np.random.default_rng(31)
    v = np.concatenate([
        (rng.standard_normal(4000) * 0.7).astype(np.float32),
        np.array([0.0, -0.0, np.finfo(np.float32).tiny, -np.finfo(np.float32).tiny, np.inf, -np.inf, np.nan,
                  0.5, -0.5, 1.0, -1.0, 1e-45, 3.4e38, -3.4e38], np.float32),
        np.arange(0, 1.0001, 0.1, dtype=np.float32),
    ])
    fx = dict(values=v)
    for tag, (thr, lo, hi) in {"bq0": (0.0, 0, 1), "bq1": (0.5, 10, 200)}.items():
        c = orc.bq_quantize(v, thr, lo, hi)
        fx[f"{tag}_params"] = np.array([thr, lo, hi], np.float32)
        fx[f"{tag}_codes"] = c
        fx[f"{tag}_deq"] = orc.bq_dequantize(c, lo, hi).view(np.uint32)
    for tag, (mn, mx, lv) in {"sq0": (-1.0, 1.0, 256), "sq1": (0.0, 1.0, 11), "sq2": (-3.0, 2.0, 2)}.items():
        c = orc.sq_quantize(v, mn, mx, lv)
        fx[f"{tag}_params"] = np.array([mn, mx, lv], np.float32)
        fx[f"{tag}_codes"] = c
        fx[f"{tag}_deq"] = orc.sq_dequantize(c, mn, mx, lv).view(np.uint32)
    allh = np.arange(65536, dtype=np.uint16)
    fx["f16_all_to_f32_bits"] = orc.dequantize_f16(allh).view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "codec.npz"), **fx)
    print("golden vectors written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(HERE, f)) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
