"""Semantic dedup on the gathered embeddings: the consumer of the embedding path.

Same quantities as SemanticDedupActor (cosmos_curate/pipelines/video/dedup/dedup_actor.py):

    semdedup_cluster     the body of dedup() for one cluster (:398-470): sort by cosine_dist_to_cent descending, L2-normalise,
                         for every row the maximum cosine similarity to any EARLIER row + the first row attaining it,
                         `kept` = number of rows with score <= 1 - eps
    assign_to_centroids  nearest centroid (Euclidean, as KMeans) + cosine distance to it (:244-249)
    spherical_kmeans     the KMeansMG call on unit-norm rows (:222-241) as plain Lloyd iterations; with a process group the
                         per-rank centroid sums / counts are all-reduced over NCCL (the one collective of this step)

Everything numeric runs in libcurate_b200 (cb_rowdot_argmax / cb_rows_l2_normalize / cb_cluster_sums) in fp32; torch is the
container (allocation, sort, all_reduce).  Reading / writing the parquet shards stays with the reference's actor.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check
from .runtime import Context, _stream_ptr, get_context


def _f32(x, device) -> torch.Tensor:
    t = torch.as_tensor(x)
    return t.to(device=device, dtype=torch.float32).contiguous()


def _pad16(x: torch.Tensor) -> torch.Tensor:
    d = x.shape[1]
    if d % 16 == 0:
        return x
    out = torch.zeros((x.shape[0], (d + 15) // 16 * 16), dtype=x.dtype, device=x.device)
    out[:, :d] = x
    return out


def l2_normalize_rows_(x: torch.Tensor, ctx: Context | None = None) -> torch.Tensor:
    """In place x[r] /= max(|x[r]|, 1e-12); returns the norms."""
    ctx = ctx or get_context()
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    norms = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    check(ctx.lib.cb_rows_l2_normalize(ctx.h, x.data_ptr(), x.shape[0], x.shape[1], norms.data_ptr(), _stream_ptr()), "cb_rows_l2_normalize", ctx.h)
    return norms


def rowdot_argmax(a: torch.Tensor, b: torch.Tensor, *, bias: torch.Tensor | None = None, upper: bool = False, clip: bool = False,
                  init_val: float = float("-inf"), ctx: Context | None = None) -> tuple[torch.Tensor, torch.Tensor]:  # fmt: skip
    """For every row j of b: (max_i a_i . b_j + bias_i, first i attaining it)."""
    ctx = ctx or get_context()
    assert a.is_cuda and b.is_cuda and a.dtype == b.dtype == torch.float32 and a.shape[1] == b.shape[1]
    assert a.is_contiguous() and b.is_contiguous() and a.shape[1] % 16 == 0
    val = torch.empty((b.shape[0],), dtype=torch.float32, device=b.device)
    idx = torch.empty((b.shape[0],), dtype=torch.int32, device=b.device)
    flags = (_lib.ROWDOT_UPPER if upper else 0) | (_lib.ROWDOT_CLIP if clip else 0)
    check(ctx.lib.cb_rowdot_argmax(ctx.h, a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], a.shape[1], bias.data_ptr() if bias is not None else None, flags,
                                   C.c_float(init_val), val.data_ptr(), idx.data_ptr(), _stream_ptr()), "cb_rowdot_argmax", ctx.h)  # fmt: skip
    return val, idx


def semdedup_cluster(ids, embeddings, cosine_dist_to_cent, eps: float, ctx: Context | None = None) -> dict:
    """One cluster of dedup() (dedup_actor.py:398-470).  Returns the pruning table columns (in the sorted order the reference
    writes them) and the kept/total counters."""
    ctx = ctx or get_context()
    dev = f"cuda:{ctx.device}"
    ids = np.asarray(ids)
    dist = _f32(cosine_dist_to_cent, dev)
    m = len(ids)
    if m == 0:
        return {"id": ids, "max_id": ids, "cosine_sim_score": np.zeros(0, np.float32), "kept": 0, "total": 0}
    order = torch.sort(dist, descending=True, stable=True).indices  # farthest from the centroid first
    e = _pad16(_f32(embeddings, dev)[order].contiguous())
    l2_normalize_rows_(e, ctx)
    maxv, argi = rowdot_argmax(e, e, upper=True, clip=True, init_val=-1.0, ctx=ctx)
    maxv[0], argi[0] = 0.0, 0  # legacy: the first item has no earlier neighbour -> itself with score 0.0 (:464-466)
    threshold = float(1 - eps)
    kept = int((maxv <= threshold).sum().item())
    argi = torch.where(argi < 0, torch.zeros_like(argi), argi)
    order_h = order.cpu().numpy()
    sorted_ids = ids[order_h]
    return {"id": sorted_ids, "max_id": sorted_ids[argi.cpu().numpy()], "cosine_sim_score": maxv.cpu().numpy(), "kept": kept, "total": m}


def assign_to_centroids(x_unit: torch.Tensor, centroids: torch.Tensor, ctx: Context | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """labels = argmin_c |x - c|^2 (first on ties); cosine_dist = 1 - clip(x . c/|c|, -1, 1) to the assigned centroid."""
    ctx = ctx or get_context()
    bias = -0.5 * (centroids * centroids).sum(dim=1)
    _, labels = rowdot_argmax(centroids, x_unit, bias=bias.contiguous(), ctx=ctx)
    cu = centroids.clone()
    l2_normalize_rows_(cu, ctx)
    sim = (x_unit * cu[labels.long()]).sum(dim=1)
    return labels, 1.0 - sim.clamp(-1.0, 1.0)


def spherical_kmeans(x, n_clusters: int, max_iter: int = 300, seed: int = 0, tol: float = 1e-4, group=None, ctx: Context | None = None) -> dict:
    """Lloyd's k-means on the unit-normalised rows of x (this rank's shard when `group` is given).

    Initial centroids: n_clusters rows drawn with a seeded permutation from rank 0's shard (broadcast).  Every iteration:
    assignment (cb_rowdot_argmax), per-cluster sums (cb_cluster_sums, deterministic), all_reduce of sums and counts over
    `group`, centroid = sum / count (an empty cluster keeps its centroid).  Stops when the squared centroid shift falls
    below tol * mean squared centroid norm, the criterion sklearn/cuML use on inertia-free convergence."""
    import torch.distributed as dist

    ctx = ctx or get_context()
    dev = f"cuda:{ctx.device}"
    x = _f32(x, dev).clone()
    d_in = x.shape[1]
    x = _pad16(x)
    l2_normalize_rows_(x, ctx)
    n, d = x.shape
    multi = group is not None
    rank = dist.get_rank(group) if multi else 0
    bad = torch.tensor([1.0 if (rank == 0 and n < n_clusters) else 0.0], dtype=torch.float32, device=dev)
    if multi:
        dist.all_reduce(bad, group=group)  # every rank raises together instead of hanging in the broadcast below
    if float(bad.item()) > 0:
        msg = f"n_clusters={n_clusters} exceeds the rows of rank 0's shard: the initial centroids are drawn from it"
        raise ValueError(msg)
    if rank == 0:
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed))[:n_clusters]
        cent = x[perm.to(dev)].clone()
    else:
        cent = torch.empty((n_clusters, d), dtype=torch.float32, device=dev)
    if multi:
        dist.broadcast(cent, src=dist.get_global_rank(group, 0), group=group)
    n_iter = 0
    for n_iter in range(1, max_iter + 1):  # noqa: B007
        bias = (-0.5 * (cent * cent).sum(dim=1)).contiguous()
        _, labels = rowdot_argmax(cent, x, bias=bias, ctx=ctx)
        order = torch.sort(labels.long(), stable=True).indices.contiguous()
        counts = torch.bincount(labels.long(), minlength=n_clusters)
        seg = torch.zeros(n_clusters + 1, dtype=torch.int64, device=dev)
        seg[1:] = torch.cumsum(counts, 0)
        sums = torch.zeros((n_clusters, d), dtype=torch.float32, device=dev)
        check(ctx.lib.cb_cluster_sums(ctx.h, x.data_ptr(), order.data_ptr(), seg.data_ptr(), n_clusters, d, sums.data_ptr(), _stream_ptr()), "cb_cluster_sums", ctx.h)
        counts_f = counts.to(torch.float32)
        if multi:
            dist.all_reduce(sums, group=group)
            dist.all_reduce(counts_f, group=group)
        new = torch.where(counts_f[:, None] > 0, sums / counts_f.clamp(min=1.0)[:, None], cent)
        shift = ((new - cent) ** 2).sum().item()
        scale = (new**2).sum(dim=1).mean().item()
        cent = new
        if shift <= tol * scale:
            break
    labels, cos_dist = assign_to_centroids(x, cent, ctx)
    return {"centroids": cent[:, :d_in].contiguous(), "labels": labels, "cosine_dist_to_cent": cos_dist, "n_iter": n_iter, "x_unit": x}
