"""Shot boundaries from per-frame transition probabilities: the integer frame arithmetic around the shot network.

Same results as the module-level helpers of cosmos_curate/pipelines/video/clipping/transnetv2_extraction_stages.py:

    predictions_from_probabilities   _get_predictions' last line (:263-264): prob > threshold, compared in float32
    scenes_from_predictions          _get_scenes (:267-299)
    filter_scenes                    _get_filtered_scenes (:302-366) with _crop_scenes (:351-366) and _create_spans (:369-392)
    stage_lengths                    TransNetV2ClipExtractionStage._get_min_length / _get_max_length / crop (:150-158, :196)
    clips_from_scenes                uuid5 + span seconds (:198-209)

Host-only and tiny (a few thousand integers per hour of video); the network itself runs in libcurate_b200.
"""

from __future__ import annotations

import math
import uuid

import numpy as np


def predictions_from_probabilities(prob: np.ndarray, threshold: float) -> np.ndarray:
    """fp32 [n] -> uint8 [n, 1].  torch evaluates `fp32_tensor > python_float` in float32, so the threshold is rounded
    to float32 before the comparison (0.4 becomes 0.4000000059604645)."""
    p = np.asarray(prob, dtype=np.float32).reshape(-1)
    return (p > np.float32(threshold)).astype(np.uint8).reshape(-1, 1)


def scenes_from_predictions(predictions: np.ndarray, *, entire_scene_as_clip: bool) -> np.ndarray:
    """0/1 per frame -> int32 [k, 2] rows (first frame of the shot, frame where the next transition starts).

    A shot closes at every 0->1 edge (except at frame 0) and the next one opens at the following 1->0 edge.  When at
    least one shot closed and the track ends on 0, a trailing shot runs to the LAST INDEX (n-1, not n); with no edges
    at all the whole video [0, n) is one shot if `entire_scene_as_clip`."""
    p = np.asarray(predictions).reshape(-1).astype(np.int8)
    n = len(p)
    if n == 0:
        return np.zeros((0, 2), dtype=np.int32)
    prev = np.concatenate([[0], p[:-1]])
    rising = np.flatnonzero((prev == 0) & (p == 1))
    rising = rising[rising != 0]
    falling = np.flatnonzero((prev == 1) & (p == 0))
    # the start in force at a rising edge i is the latest falling edge <= i (0 before any)
    k = np.searchsorted(falling, rising, side="right")
    starts = np.where(k > 0, falling[np.maximum(k - 1, 0)], 0) if len(falling) else np.zeros(len(rising), dtype=np.int64)
    rows = [(int(s), int(e)) for s, e in zip(starts, rising)]
    if rows and p[-1] == 0:
        rows.append((int(falling[-1]) if len(falling) else 0, n - 1))
    if not rows and entire_scene_as_clip:
        rows.append((0, n))
    return np.array(rows, dtype=np.int32).reshape(-1, 2)


def _stride(start: int, end: int, max_length: int, min_length: int | None) -> list[tuple[int, int]]:
    full, rest = divmod(end - start, max_length)
    out = [(start + i * max_length, start + (i + 1) * max_length) for i in range(full)]
    if rest:
        out.append((start + full * max_length, end))
    # only the span that reaches `end` (always the last one) can be dropped for being shorter than min_length
    if min_length and out and out[-1][1] - out[-1][0] < min_length:
        out.pop()
    return out


def filter_scenes(scenes: np.ndarray, min_length: int | None = None, max_length: int | None = None, max_length_mode: str = "truncate",
                  crop_length: int | None = None) -> np.ndarray:  # fmt: skip
    sc = np.array(scenes, dtype=np.int32)
    if sc.ndim != 2:
        error_msg = "Scenes numpy array needs to be a 2D rank matrix!"
        raise ValueError(error_msg)
    if max_length is not None:
        if max_length_mode == "truncate":
            sc[:, 1] = np.minimum(sc[:, 0] + max_length, sc[:, 1])
        elif max_length_mode == "stride":
            rows: list[tuple[int, int]] = []
            for a, b in sc:
                rows.extend(_stride(int(a), int(b), int(max_length), min_length))
            sc = np.array(rows, dtype=np.int32).reshape(-1, 2)
        else:
            error_msg = f"Method `{max_length_mode}` not implemented!"
            raise NotImplementedError(error_msg)
    if crop_length is not None:
        sc = np.stack([sc[:, 0] + crop_length, sc[:, 1] - crop_length], axis=1)
        sc = sc[(sc[:, 1] - sc[:, 0]) > 0]
    if min_length is not None:
        sc = sc[(sc[:, 1] - sc[:, 0]) >= min_length]
    return sc


def stage_lengths(framerate: float, min_length_s: float | None, min_length_frames: int | None, max_length_s: float | None, crop_s: float | None):
    """-> (min_length, max_length, crop_length) in frames."""
    mn = math.ceil(min_length_s * framerate) if min_length_s is not None else None
    if min_length_frames is not None:
        mn = max(mn, min_length_frames) if mn is not None else min_length_frames
    mx = math.ceil(max_length_s * framerate) if max_length_s is not None else None
    crop = int(crop_s * framerate) if crop_s else None
    return mn, mx, crop


def clips_from_scenes(source: str, scenes: np.ndarray, framerate: float, limit_clips: int = 0) -> list[tuple[uuid.UUID, tuple[float, float]]]:
    out = []
    for a, b in scenes:
        out.append((uuid.uuid5(uuid.NAMESPACE_URL, f"{source}_{a}_{b}"), (float(a) / framerate, float(b) / framerate)))
        if limit_clips > 0 and len(out) >= limit_clips:
            break
    return out
