"""Oracle: the numeric core of SemanticDedupActor (cosmos_curate/pipelines/video/dedup/dedup_actor.py), in numpy.

  pairwise_max        dedup() :398-470 for one cluster - sort by cosine_dist_to_cent descending, L2-normalise
                      (x / max(|x|, 1e-12)), strict upper-triangular cosine matrix (only earlier rows i < j), clip to
                      [-1, 1], per column the maximum and the FIRST row attaining it (cp.argmax; a later tile only wins
                      with a strictly greater value, :437-440), start values -1.0 / -1, row 0 forced to (0.0, 0),
                      kept = count(max <= 1 - eps)
  assign              nearest centroid as KMeans defines it (argmin squared Euclidean distance) and
                      cosine_dist_to_cent = 1 - clip(x . c/|c|) (:244-249)

Pinned (tests/test_dedup_cpu.py, tests/golden/dedup_ref.npz) against the reference's OWN array code for dedup(): the section
dedup_actor.py:404-466 is executed from its source with numpy standing in for cupy (oracle/ref_import.dedup_core; CuPy mirrors
numpy for every call in it), tiles of 256 and of the default 4096, exact / scaled / near duplicates and sort-key ties.  What stays
unpinned is library behaviour outside that section: cudf's `sort_values` tie order (restated as a stable sort) and KMeansMG
(scalable k-means++ seeding on cuML's RNG is not reproducible outside cuML: only its defining properties are tested - labels are
nearest centroids, centroids are cluster means, multi-rank == single-rank).  The reference's 4096-row tiling does not change
any result (max / first-argmax are tiling-invariant by the `>` rule above).

Test infrastructure only (see oracle/__init__.py).
"""

from __future__ import annotations

import numpy as np


def l2_normalize(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    n = np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    return x / np.maximum(n, np.float32(1e-12))


def pairwise_max(ids, embeddings, cosine_dist_to_cent, eps: float) -> dict:
    ids = np.asarray(ids)
    dist = np.asarray(cosine_dist_to_cent, dtype=np.float32)
    m = len(ids)
    order = np.argsort(-dist, kind="stable")  # descending, ties keep their input order
    e = l2_normalize(np.asarray(embeddings, dtype=np.float32)[order])
    maxv = np.full(m, -1.0, dtype=np.float32)
    argi = np.full(m, -1, dtype=np.int32)
    tile = 512
    for j0 in range(0, m, tile):
        j1 = min(m, j0 + tile)
        s = np.clip(e[:j1] @ e[j0:j1].T, -1.0, 1.0).astype(np.float32)  # rows i < j1, columns j0..j1
        i_idx = np.arange(j1)[:, None]
        j_idx = np.arange(j0, j1)[None, :]
        s = np.where(i_idx < j_idx, s, -np.inf)
        a = np.argmax(s, axis=0)  # first maximum
        v = s[a, np.arange(j1 - j0)]
        better = v > maxv[j0:j1]
        maxv[j0:j1] = np.where(better, v, maxv[j0:j1])
        argi[j0:j1] = np.where(better, a.astype(np.int32), argi[j0:j1])
    if m:
        maxv[0], argi[0] = 0.0, 0
    kept = int(np.count_nonzero(maxv <= np.float32(1 - eps))) if m else 0
    argi = np.where(argi < 0, 0, argi)
    sid = ids[order]
    return {"id": sid, "max_id": sid[argi], "cosine_sim_score": maxv, "kept": kept, "total": m, "sim_matrix_unit": e}


def assign(x_unit: np.ndarray, centroids: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    x = np.asarray(x_unit, dtype=np.float64)
    c = np.asarray(centroids, dtype=np.float64)
    d2 = (x * x).sum(1)[:, None] - 2.0 * x @ c.T + (c * c).sum(1)[None, :]
    labels = np.argmin(d2, axis=1).astype(np.int32)
    cu = c / np.maximum(np.linalg.norm(c, axis=1, keepdims=True), 1e-12)
    sim = (x * cu[labels]).sum(1)
    return labels, (1.0 - np.clip(sim, -1.0, 1.0)).astype(np.float32)
