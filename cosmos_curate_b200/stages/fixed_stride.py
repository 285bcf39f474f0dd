"""FixedStrideExtractorStage: drop-in for the reference's fixed-interval splitter
(cosmos_curate/pipelines/video/clipping/clip_extraction_stages.py:664-760) - same class name, constructor and task mutations.
Host code only; the span arithmetic lives in `cosmos_curate_b200/spans.py`.  Its output (`clip.span` on every camera of the
session) is what `NvdecClipAestheticStage(source="video_span")` and `ClipStreamCopyStage` consume."""

from __future__ import annotations

from ..data_model import StageTimer
from ..interfaces import CuratorStage
from ..spans import assert_video_clip_alignment, populate_clips_fixed_stride

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


class FixedStrideExtractorStage(CuratorStage):
    """Cuts every session into windows of `clip_len_s` seconds every `clip_stride_s` seconds."""

    def __init__(self, clip_len_s: float = 10, clip_stride_s: float = 10, min_clip_length_s: float = 10, limit_clips: int = 0, *,
                 verbose: bool = False, log_stats: bool = False) -> None:  # fmt: skip
        assert clip_stride_s, "a zero stride would never advance"
        self._timer = StageTimer(self)
        self.clip_len_s, self.clip_stride_s, self.min_clip_length_s = clip_len_s, clip_stride_s, min_clip_length_s
        self._limit_clips, self._verbose, self._log_stats = limit_clips, verbose, log_stats

    def _split(self, task) -> None:
        first_incomplete = next((v for v in task.videos if not v.has_metadata()), None)
        if first_incomplete is not None:  # the reference stops at the first camera without metadata and marks only that one (:729-737)
            first_incomplete.errors["metadata"] = "incomplete"
            msg = f"Incomplete metadata for {first_incomplete.input_video}. Skipping"
            raise ValueError(msg)
        populate_clips_fixed_stride(task.videos, task.session_id, self.clip_len_s, self.clip_stride_s, self.min_clip_length_s, limit_clips=self._limit_clips)

    def process_data(self, tasks):
        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            with self._timer.time_process():
                try:
                    self._split(task)
                except Exception as exc:  # noqa: BLE001 - a session that cannot be split is reported on the task, the batch goes on (:744-752)
                    task.errors["FixedStrideExtractorStage"] = f"failed to populate clips: {exc}"
                    logger.error(f"session {task.session_id}: {exc}")
            if self._log_stats:
                name, stats = self._timer.log_stats()
                task.stage_perf[name] = stats
            assert_video_clip_alignment(task.videos)
        return tasks
