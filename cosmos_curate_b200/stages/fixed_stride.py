"""FixedStrideExtractorStage - same name, constructor and task mutations as the reference stage
(cosmos_curate/pipelines/video/clipping/clip_extraction_stages.py:664-760): host-only, no GPU.  It writes the clip spans the
decode stages consume (`NvdecClipAestheticStage(source="video_span")` / `ClipStreamCopyStage` read `clip.span`)."""

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
    """Stage that extracts video clips using fixed-length intervals."""

    def __init__(self, clip_len_s: float = 10, clip_stride_s: float = 10, min_clip_length_s: float = 10, limit_clips: int = 0, *,
                 verbose: bool = False, log_stats: bool = False) -> None:  # fmt: skip
        self._timer = StageTimer(self)
        self.clip_stride_s = clip_stride_s
        assert clip_stride_s
        self.clip_len_s = clip_len_s
        self.min_clip_length_s = min_clip_length_s
        self._limit_clips = limit_clips
        self._verbose = verbose
        self._log_stats = log_stats

    def process_data(self, tasks):
        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            with self._timer.time_process():
                try:
                    for video in task.videos:
                        if not video.has_metadata():
                            video.errors["metadata"] = "incomplete"
                            error_msg = f"Incomplete metadata for {video.input_video}. Skipping"
                            raise ValueError(error_msg)  # noqa: TRY301
                    populate_clips_fixed_stride(task.videos, task.session_id, self.clip_len_s, self.clip_stride_s, self.min_clip_length_s,
                                                limit_clips=self._limit_clips)  # fmt: skip
                except Exception as e:  # noqa: BLE001
                    logger.error(f"Failed to populate clips for {task.session_id}: {e}")
                    task.errors["FixedStrideExtractorStage"] = f"failed to populate clips: {e}"
            if self._log_stats:
                stage_name, stage_perf_stats = self._timer.log_stats()
                task.stage_perf[stage_name] = stage_perf_stats
        for task in tasks:
            assert_video_clip_alignment(task.videos)
        return tasks
