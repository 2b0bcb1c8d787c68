"""Model persistence (the reference has none, `ROADMAP.md:28`, so the format is free).

One `.npz` file per model: a `kind` tag, a format version, the metric name and the arrays that define the model --
for a ProductQuantizer the codebooks `[m, k, sub_dim]` f32, for a TSVQ the breadth-first node arrays
(`centroids [n_nodes, dim]` f32, `left` / `right` i32, plus the split dim / median / count of every node for inspection).
Loading re-creates the immutable engine handle from the arrays (`ProductQuantizer.from_codebooks`, `TSVQ.from_tree`),
so a loaded model encodes bit-identically to the one that was saved.  BQ / SQ are three scalars each and are stored too
for completeness.
"""
from __future__ import annotations

import numpy as np

FORMAT_VERSION = 1


def _pack(kind: str, metric: str | None, **arrays) -> dict:
    d = {"kind": np.array(kind), "version": np.array(FORMAT_VERSION, np.int32)}
    if metric is not None:
        d["metric"] = np.array(metric)
    d.update(arrays)
    return d


def dumps(model) -> dict:
    """The arrays that define `model` (the dict that `save` writes)."""
    from . import api
    if isinstance(model, api.ProductQuantizer):
        return _pack("pq", model.distance_metric(), codebooks=np.ascontiguousarray(model.codebooks, dtype=np.float32))
    if isinstance(model, api.TSVQ):
        t = model.tree()
        return _pack("tsvq", model.distance_metric(), centroids=t["centroids"], left=t["left"], right=t["right"],
                     split_dim=t["split_dim"], median=t["median"], count=t["count"])
    if isinstance(model, api.ScalarQuantizer):
        return _pack("sq", None, min=np.float32(model.min), max=np.float32(model.max), levels=np.int64(model.levels))
    if isinstance(model, api.BinaryQuantizer):
        return _pack("bq", None, threshold=np.float32(model.threshold), low=np.uint8(model.low), high=np.uint8(model.high))
    raise TypeError(f"cannot persist {type(model).__name__}")


def save(model, path: str) -> None:
    np.savez(path, **dumps(model))


def check(d) -> str:
    """Validates a loaded dict; returns its kind.  Raises ValueError on anything a handle must not be built from."""
    kind = str(d["kind"])
    if int(d["version"]) != FORMAT_VERSION:
        raise ValueError(f"unsupported model format version {int(d['version'])}")
    if kind == "pq":
        cb = d["codebooks"]
        if cb.dtype != np.float32 or cb.ndim != 3 or 0 in cb.shape:
            raise ValueError("codebooks must be a non-empty [m, k, sub_dim] float32 array")
    elif kind == "tsvq":
        c, l, r = d["centroids"], d["left"], d["right"]
        n = c.shape[0]
        if c.dtype != np.float32 or c.ndim != 2 or n == 0 or l.shape != (n,) or r.shape != (n,):
            raise ValueError("tree arrays have inconsistent shapes")
        for ch in (l, r):   # breadth-first numbering: a child's id is larger than its parent's (no cycles)
            idx = np.nonzero(ch >= 0)[0]
            if np.any(ch[idx] <= idx) or np.any(ch[idx] >= n):
                raise ValueError("tree has an invalid child index")
    elif kind not in ("sq", "bq"):
        raise ValueError(f"unknown model kind {kind!r}")
    return kind


def load(path: str, engine=None):
    from . import api
    with np.load(path, allow_pickle=False) as z:
        d = {k: z[k] for k in z.files}
    kind = check(d)
    if kind == "pq":
        return api.ProductQuantizer.from_codebooks(d["codebooks"], api.Distance(str(d["metric"])), engine=engine)
    if kind == "tsvq":
        return api.TSVQ.from_tree(d["centroids"], d["left"], d["right"], api.Distance(str(d["metric"])), engine=engine)
    if kind == "sq":
        return api.ScalarQuantizer(float(d["min"]), float(d["max"]), int(d["levels"]), engine=engine)
    return api.BinaryQuantizer(float(d["threshold"]), int(d["low"]), int(d["high"]), engine=engine)
