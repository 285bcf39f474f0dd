"""Oracle: frame-index sampling (integer results, must be bit-exact).

Restates cosmos_curate/pipelines/video/utils/decoder_utils.py:
  find_closest_indices  :281-312
  sample_closest        :315-386
  extract_frames policy :640-652  (sequence / middle timestamp selection)
  lcm + stride rule     cosmos_curate/pipelines/video/clipping/clip_frame_extraction_stages.py:94-137

Test infrastructure only (see oracle/__init__.py).
"""

from __future__ import annotations

import math
from functools import reduce

import numpy as np


def find_closest_indices(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """decoder_utils.py:281-312 - nearest index in sorted `src` for each `dst`; ties go left;
    anything >= src[-1] maps to the last index."""
    src = np.asarray(src)
    dst = np.asarray(dst)
    right = np.searchsorted(src, dst)
    right = np.clip(right, 1, len(src) - 1)
    out = right - 1
    take_right = np.abs(dst - src[right]) < np.abs(dst - src[out])
    out[take_right] = right[take_right]
    out[dst >= src[-1]] = len(src) - 1
    return out.astype(np.int32)


def sample_closest(src, sample_rate, start=None, stop=None, endpoint=True, dedup=True):
    """decoder_utils.py:315-386.  Note the float32 ``np.arange`` and the half-interval
    extension of ``stop`` - both matter for which indices come out."""
    if sample_rate <= 0:
        raise ValueError(f"Sample rate must be greater than 0, got sample_rate={sample_rate}")
    src = np.asarray(src)
    interval = 1.0 / sample_rate
    _start = start if start is not None else src[0]
    _stop = stop if stop is not None else src[-1]
    if endpoint:
        _stop += interval * 0.5
    samples = np.arange(_start, _stop, interval, dtype=np.float32)
    idx = find_closest_indices(src, samples)
    if not endpoint and np.isclose(samples[-1], _stop):
        idx = idx[:-1]
        samples = samples[:-1]
    if dedup:
        idx, counts = np.unique(idx, return_counts=True)
        idx = idx.astype(np.int32)
        counts = counts.astype(np.int32)
    else:
        counts = np.ones_like(idx, dtype=np.int32)
    return idx, counts, samples


def select_policy_timestamps(all_ts: np.ndarray, policy: str) -> np.ndarray:
    """decoder_utils.py:640-652.  `sequence` keeps every timestamp; `middle` keeps ONE timestamp,
    which then makes decode_video_cpu sample index 0 of a 1-element array (reference quirk,
    SURVEY.md a6): the returned frame id is 0, not the middle frame."""
    if len(all_ts) == 0:
        raise ValueError("Can't extract frames from empty video")
    if policy == "sequence" or len(all_ts) == 1:
        return all_ts
    if policy == "middle":
        n = len(all_ts)
        i = n // 2 - 1 if n % 2 == 0 else n // 2
        return all_ts[i : i + 1]
    raise NotImplementedError("Extraction policies apart from Sequence and Middle not available yet")


def frame_ids_for(all_ts: np.ndarray, policy: str, fps: float):
    """extract_frames -> decode_video_cpu -> sample_closest chain (decoder_utils.py:611-672, 505-570)."""
    ts = select_policy_timestamps(np.asarray(all_ts, dtype=np.float32), policy)
    ids, counts, _ = sample_closest(ts, fps, start=ts[0], stop=ts[-1], endpoint=True, dedup=True)
    return ids, counts


def lcm_multiple(fps_list):
    """clip_frame_extraction_stages.py:94-100."""

    def lcm(a, b):
        return abs(a * b) // math.gcd(int(a), int(b))

    return reduce(lcm, fps_list)


def signature(policy: str, fps: float) -> str:
    """decoder_utils.py:110-117: f"{policy!s}-{int(fps*1000)}" with policy an enum member."""
    return f"FrameExtractionPolicy.{policy}-{int(fps * 1000)}"


def frames_per_signature(all_ts, policy: str, target_fps: list):
    """clip_frame_extraction_stages.py:111-152: which decoded-frame list (as positions into the
    expanded id list) every signature receives.  Returns {signature: (ids_expanded)}."""
    use_lcm = len(target_fps) > 1 and all(
        (f.is_integer() if isinstance(f, float) else isinstance(f, int)) for f in target_fps
    )
    out = {}
    if use_lcm:
        lcm = lcm_multiple(target_fps)
        ids, counts = frame_ids_for(all_ts, policy, lcm)
        expanded = np.repeat(ids, counts)
        for f in target_fps:
            out[signature(policy, f)] = expanded[:: int(lcm / f)]
    else:
        for f in target_fps:
            ids, counts = frame_ids_for(all_ts, policy, f)
            out[signature(policy, f)] = np.repeat(ids, counts)
    return out
