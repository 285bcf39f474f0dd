"""Multi-GPU partitioning of the path: independent clips, one process per GPU, no data-path collective.

`shard_by_weight` mirrors how the reference balances work by `PipelineTask.weight`
(stage_interface.py:31-39; Video.weight, data_model.py:509-523): longest-processing-time-first greedy assignment so
per-rank decoded-pixel totals match.  `all_gather_embeddings` is the single optional exchange step of the design
(SURVEY.md 8e): per-rank [n_i, D] embedding matrices gathered on every rank before dedup (replaces the reference's
parquet round trip, metadata_writer_stage.py:467-485 -> dedup_actor.py:199-204).  NCCL over NVLink on GPUs; the same
code runs on gloo for the CPU tests.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_by_weight(weights, world_size: int) -> list[list[int]]:
    """Indices per rank; deterministic; heaviest first onto the lightest rank."""
    if world_size <= 0:
        msg = "world_size must be positive"
        raise ValueError(msg)
    order = sorted(range(len(weights)), key=lambda i: (-float(weights[i]), i))
    loads = [0.0] * world_size
    out: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += float(weights[i])
    for lst in out:
        lst.sort()
    return out


def shard_tasks(tasks: list, world_size: int, rank: int) -> list:
    """This rank's share of `tasks`, balanced by `PipelineTask.weight` (a SplitPipeTask's weight is the summed duration of its videos
    normalised to 5 minutes x the fraction of clips it carries, data_model.py:509-523, 779-790).  Every rank computes the same
    partition from the same list: no communication; input order is kept within a rank."""
    if not 0 <= rank < world_size:
        msg = f"rank {rank} outside world of {world_size}"
        raise ValueError(msg)
    return [tasks[i] for i in shard_by_weight([t.weight for t in tasks], world_size)[rank]]


def clip_weight(width: int, height: int, n_frames: int) -> float:
    """Decode cost proxy: pixels that NVDEC has to produce."""
    return float(width) * float(height) * float(n_frames)


def all_gather_embeddings(local: torch.Tensor, ids: torch.Tensor | None = None, group=None):
    """local [n_i, D] (any n_i per rank, same D) -> ([sum n_i, D], ids or None), identical on every rank, rank order.

    One padded all_gather (+ one for the counts): <= 46 MB at 10k x 1152 fp32, latency-bound on NVSwitch, so no custom
    kernel and no overlap machinery."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local, ids
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(counts)
    pad = torch.zeros((cap, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    emb = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    if ids is None:
        return emb, None
    ipad = torch.zeros((cap,), dtype=ids.dtype, device=ids.device)
    ipad[: ids.shape[0]] = ids
    iparts = [torch.empty_like(ipad) for _ in range(world)]
    dist.all_gather(iparts, ipad, group=group)
    return emb, torch.cat([p[:c] for p, c in zip(iparts, counts)], dim=0)


def rank_slice(n_items: int, rank: int, world_size: int) -> np.ndarray:
    """Round-robin shard (uniform synthetic sets: bench.py)."""
    return np.arange(rank, n_items, world_size)
