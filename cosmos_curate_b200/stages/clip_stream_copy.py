"""ClipStreamCopyStage: clip mp4s cut out of the source video WITHOUT a transcode.

Drop-in alternative to ClipTranscodingStage (clip_extraction_stages.py:167-316): same inputs (video.encoded_data +
video.clips[*].span), same outputs (clip.encoded_data holds a standalone mp4 per clip, video.encoded_data dropped,
clip.errors["transcode"] on failure, task.stage_perf).  Where the reference runs one `ffmpeg -ss/-to ... -c:v libopenh264
-b:v 4M` re-encode per clip (B200 has no NVENC, :360-368 falls back to CPU encoders there), this stage copies the clip's
coded pictures (cb_mp4_cut): a clip starts on the sync sample at or before span[0] and ends with the last frame presented
before span[1], so it can be up to one GOP longer at the front than the re-encoded clip - pixels are the source's own, not a
4 Mb/s re-encode's.  Cost: a memcpy of the GOPs (measured in bench.py `clip_cut`), no decode, no encode.

`snap_spans=True` (default) rewrites clip.span to the frames actually contained, so downstream sampling (which is relative to
the clip file) and the metadata agree with the bytes.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import CurateB200Error, check
from ..data_model import StageTimer
from ..interfaces import CuratorStage, CuratorStageResource
from ..runtime import _as_u8, mp4_index
from ..spans import chunk_tasks

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


def mp4_cut(data, first_sample: int, n_samples: int) -> np.ndarray:
    """cb_mp4_cut: samples [first, first + n) of the video track (decode order, first = sync sample) as a standalone mp4."""
    buf = _as_u8(data)
    lib = _lib.load()
    need = C.c_size_t(0)
    check(lib.cb_mp4_cut(None, buf.ctypes.data, buf.size, int(first_sample), int(n_samples), None, 0, C.byref(need)), "cb_mp4_cut")
    out = np.empty(need.value, dtype=np.uint8)
    check(lib.cb_mp4_cut(None, buf.ctypes.data, buf.size, int(first_sample), int(n_samples), out.ctypes.data, out.size, C.byref(need)), "cb_mp4_cut")
    return out


def span_sample_range(idx: dict, span: tuple[float, float]) -> tuple[int, int, float, float]:
    """(first sample, sample count, start_s, end_s) of the stream-copied clip for `span` seconds of the indexed video.

    Frames with span[0] <= pts < span[1] are wanted (ffmpeg -ss/-to semantics); the cut starts at the sync sample at or before
    the first of them in decode order and runs to the last sample (decode order) any wanted frame needs."""
    ts = idx["pts"].astype(np.float64) / float(idx["timescale"])
    start, end = float(span[0]), float(span[1])
    wanted = np.flatnonzero((ts >= start - 1e-9) & (ts < end - 1e-9))
    if wanted.size == 0:
        msg = f"span {span} holds no frame of the video ({ts.min():.3f}..{ts.max():.3f} s)"
        raise ValueError(msg)
    first = int(wanted.min())
    sync = np.flatnonzero(idx["sync"][: first + 1])
    if sync.size == 0:
        msg = "no sync sample at or before the span"
        raise ValueError(msg)
    first = int(sync[-1])
    last = int(wanted.max())
    contained = ts[first : last + 1]
    frame = float(np.median(np.diff(np.sort(ts)))) if len(ts) > 1 else 0.0
    return first, last - first + 1, float(contained.min()), float(contained.max() + frame)


class ClipStreamCopyStage(CuratorStage):
    def __init__(self, *, snap_spans: bool = True, num_cpus_per_worker: float = 1.0, num_clips_per_chunk: int | None = None, verbose: bool = False,
                 log_stats: bool = False) -> None:  # fmt: skip
        self._timer = StageTimer(self)
        self._snap, self._cpus, self._verbose, self._log_stats = snap_spans, num_cpus_per_worker, verbose, log_stats
        # ClipTranscodingStage re-chunks its output into tasks of ~num_clips_per_chunk * 8 s of clips (default 32, :184, :301); None keeps
        # the incoming tasks (one chunk each)
        self._num_clips_per_chunk = num_clips_per_chunk

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(cpus=self._cpus)

    def _process_video(self, video) -> None:
        if not video.encoded_data:
            error_msg = "Please load video!"
            raise ValueError(error_msg)
        if not video.clips:
            logger.warning(f"No clips to cut for {video.input_video}. Skipping...")
            video.encoded_data.drop()
            return
        data = video.encoded_data.resolve()
        idx = mp4_index(data)
        for clip in video.clips:
            try:
                first, count, start_s, end_s = span_sample_range(idx, clip.span)
                clip.encoded_data = mp4_cut(data, first, count)
                if self._snap:
                    clip.span = (start_s, end_s)
            except (CurateB200Error, ValueError) as e:
                logger.error(f"stream copy failed for clip {clip.uuid} of {video.input_video}: {e}")
                clip.errors["transcode"] = str(e)

    def process_data(self, tasks):
        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            for video in task.videos:
                with self._timer.time_process(len(video.clips), video.metadata.duration if video.metadata.duration else 0):
                    try:
                        self._process_video(video)
                    except Exception as e:  # noqa: BLE001 - same convention as the reference stage (:283-287)
                        logger.exception(f"Error processing video {video.input_video}")
                        video.errors[self.__class__.__name__] = str(e)
                video.encoded_data.drop()
                video.num_total_clips, video.num_clip_chunks, video.clip_chunk_index = len(video.clips), 1, 0
            if self._log_stats:
                stage_name, stats = self._timer.log_stats()
                task.stage_perf[stage_name] = stats
        if self._num_clips_per_chunk is not None:
            return chunk_tasks(tasks, self._num_clips_per_chunk, verbose=self._verbose)
        return tasks
