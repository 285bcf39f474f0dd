"""Fixed-stride clip spans: the host arithmetic of the reference's second splitting algorithm
(`--splitting-algorithm fixed-stride`, cosmos_curate/pipelines/video/clipping/clip_extraction_stages.py:444-660): the caller
side of the hot path - it decides which `(start_s, end_s)` windows of a source video become clips.  Same function names minus
the leading underscore, same results (floats compared bit for bit in tests/test_spans_cpu.py against the reference functions
executed from their own source).

    videos_durations              num_frames / framerate per video, -1.0 when unknown (:490-512; not last - first timestamp, as
                                  the reference itself notes)
    videos_timestamps             per-video timestamp arrays, `errors["timestamps"] = "missing"` on every video without (:465-487)
    validate_video_timestamps     (:444-462)
    make_spans_fixed_stride       the window walk (:515-552)
    make_clip_uuids               uuid5(NAMESPACE_URL, f"{session_id}_{start}_{end}") (:555-565)
    populate_clips_fixed_stride   one shared span list for all cameras of a session (:568-660)
    assert_video_clip_alignment   multi-camera time alignment check (pipelines/video/utils/data_model.py:634-690)
    split_by_chunk_size           core/utils/misc/grouping.py:36-66
    slice_video_clips, chunk_tasks   one task per ~num_clips_per_chunk * 8 s of clips, all cameras cut at the same clip indices
                                  (clip_extraction_stages.py:46-163; what ClipTranscodingStage returns, :301)
"""

from __future__ import annotations

import copy
import uuid
from uuid import UUID

import numpy as np

from .data_model import Clip, SplitPipeTask, Video

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


def validate_video_timestamps(video_timestamps: list[np.ndarray]) -> None:
    if len(video_timestamps) == 0:
        msg = "No timestamps found for videos"
        raise ValueError(msg)
    if any(len(ts) == 0 for ts in video_timestamps):
        msg = "Some videos have no timestamps"
        raise ValueError(msg)


def videos_timestamps(videos: list[Video]) -> list[np.ndarray]:
    without = [v for v in videos if v.timestamps is None or len(v.timestamps) == 0]
    for v in without:
        v.errors.setdefault("timestamps", "missing")  # an earlier stage's message wins
    if without:
        msg = f"Videos missing timestamps: {[str(v.input_video) for v in without]}"
        raise ValueError(msg)
    return [v.timestamps for v in videos]


def videos_durations(videos: list[Video]) -> list[float]:
    out = []
    for v in videos:
        n, fps = v.metadata.num_frames, v.metadata.framerate
        out.append(float(n / fps) if n is not None and fps is not None and fps > 0 else -1.0)
    return out


def make_spans_fixed_stride(start_s: float, end_s: float, clip_len_s: float, clip_stride_s: float, min_clip_length_s: float) -> list[tuple[float, float]]:
    """Windows of clip_len_s every clip_stride_s from start_s while the window START is before end_s; the last ones are cut at
    end_s and kept only if still >= min_clip_length_s long.  The start accumulates by repeated addition (float semantics)."""
    out: list[tuple[float, float]] = []
    t = start_s
    while t < end_s:
        stop = end_s if t + clip_len_s > end_s else t + clip_len_s  # == min(t + clip_len_s, end_s)
        if stop - t >= min_clip_length_s:
            out.append((t, stop))
        t += clip_stride_s  # accumulated, never recomputed as start + k * stride: the reference's floats depend on it
    return out


def make_clip_uuids(session_id: str, spans: list[tuple[float, float]]) -> list[UUID]:
    return [uuid.uuid5(uuid.NAMESPACE_URL, f"{session_id}_{span[0]}_{span[1]}") for span in spans]


def populate_clips_fixed_stride(videos: list[Video], session_id: str, clip_len_s: float, clip_stride_s: float, min_clip_length_s: float,
                                *, limit_clips: int = 0) -> None:  # fmt: skip
    """Appends one `Clip(uuid, source_video, span)` per span to EVERY video of the session (multi-camera sessions share spans).
    The window is [0, min over videos of (first timestamp + duration) - max first timestamp), the reference's backwards-compatible
    choice (:582-638)."""
    durations = videos_durations(videos)
    if any(d <= 0 for d in durations):
        msg = "Some videos have invalid (zero or negative) duration"
        raise ValueError(msg)
    stamps = videos_timestamps(videos)
    validate_video_timestamps(stamps)
    first = [float(ts[0]) for ts in stamps]
    latest_start = max(first)
    if latest_start > 0.1:  # noqa: PLR2004
        logger.warning(f"cameras start at {latest_start:.2f} s, but the window below is measured from 0: span boundaries may surprise")
    window_end = min(t0 + d for t0, d in zip(first, durations, strict=True)) - latest_start
    spans = make_spans_fixed_stride(0.0, window_end, clip_len_s, clip_stride_s, min_clip_length_s)
    spans = spans[:limit_clips] if limit_clips > 0 else spans
    ids = make_clip_uuids(session_id, spans)
    for video in videos:
        video.clips.extend(Clip(uuid=u, source_video=str(video.input_video), span=sp) for u, sp in zip(ids, spans, strict=True))


def check_clip_time_alignment(clips_per_video: list[list[Clip]]) -> list[int]:
    """Indices at which the cameras' clips do not share one span (data_model.py:595-631); ValueError when the cameras hold
    different numbers of clips."""
    if not clips_per_video:
        return []
    counts = [len(c) for c in clips_per_video]
    if not all(n == counts[0] for n in counts):
        msg = f"Cannot check time alignment: videos have different clip counts {counts}. All videos must have the same number of clips."
        raise ValueError(msg)
    return [i for i in range(counts[0]) if any(c[i].span != clips_per_video[0][i].span for c in clips_per_video[1:])]


def assert_video_clip_alignment(videos: list[Video]) -> None:
    if not videos:
        return
    processed = [len(v.clips) + len(v.filtered_clips) for v in videos]
    if not all(p == processed[0] for p in processed):
        msg = (f"Multi-cam videos have processed different numbers of clips: {processed}. "
               "All cameras should process clips together to maintain time alignment.")  # fmt: skip
        raise ValueError(msg)
    for name, per_video in (("clips", [v.clips for v in videos]), ("filtered clips", [v.filtered_clips for v in videos])):
        bad = check_clip_time_alignment(per_video)
        if bad:
            spans = [c[bad[0]].span for c in per_video]
            msg = f"Multi-cam {name} at index {bad[0]} have misaligned spans: {spans}. Misaligned indices: {bad}"
            raise ValueError(msg)


def split_by_chunk_size(iterable, chunk_size: int, custom_size_func=lambda x: 1, *, drop_incomplete_chunk: bool = False):  # noqa: ARG005
    """Greedy chunks: a chunk closes as soon as its accumulated size reaches chunk_size (the closing item included)."""
    out, cur = [], 0
    for value in iterable:
        out.append(value)
        cur += custom_size_func(value)
        if cur >= chunk_size:
            yield out
            out, cur = [], 0
    if out and not drop_incomplete_chunk:
        yield out


def slice_video_clips(video: Video, start: int, end: int, chunk_index: int, num_chunks: int) -> Video:
    """A new Video holding clips[start:end]; payloads, metadata, timestamps and clip_stats are shared, errors are copied."""
    if end < start:
        msg = f"End index {end} is less than start index {start}"
        raise ValueError(msg)
    if start < 0 or end > len(video.clips):
        msg = f"Start index {start} or end index {end} is out of range [0, {len(video.clips)})"
        raise ValueError(msg)
    return Video(input_video=video.input_video, relative_path=video.relative_path, encoded_data=video.encoded_data, metadata=video.metadata,
                 frame_array=video.frame_array, timestamps=video.timestamps, clips=video.clips[start:end], num_total_clips=len(video.clips),
                 num_clip_chunks=num_chunks, clip_chunk_index=chunk_index, clip_stats=video.clip_stats, errors=copy.deepcopy(video.errors))  # fmt: skip


def chunk_tasks(tasks: list[SplitPipeTask], num_clips_per_chunk: int, *, verbose: bool = False) -> list[SplitPipeTask]:
    """Each task becomes one subtask per chunk of the PRIMARY video's clips; a chunk closes once its clips' whole-second durations
    reach num_clips_per_chunk * 8; every camera is cut at the same clip indices.  stage_perf is carried by the first subtask and
    reset on the others."""
    out: list[SplitPipeTask] = []
    for task in tasks:
        chunks = list(split_by_chunk_size(task.videos[0].clips, num_clips_per_chunk * 8, lambda c: int(c.span[1] - c.span[0])))
        start = 0
        for idx, chunk in enumerate(chunks):
            end = start + len(chunk)
            sub = SplitPipeTask(session_id=task.session_id, videos=[slice_video_clips(v, start, end, idx, len(chunks)) for v in task.videos],
                                stage_perf=copy.deepcopy(task.stage_perf))  # fmt: skip
            start = end
            if idx > 0:
                for stats in sub.stage_perf.values():
                    stats.reset()
            if verbose:
                logger.info(f"Spawning subtask {idx} with {len(sub.video.clips)} clips and weight={sub.weight:.2f}")
            out.append(sub)
    for task in out:
        assert_video_clip_alignment(task.videos)
    return out
