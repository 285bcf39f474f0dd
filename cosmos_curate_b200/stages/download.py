"""VideoDownloader for local files - same name, constructor and task mutations as the reference stage
(cosmos_curate/pipelines/video/read_write/download_stages.py:40-228): bytes into `video.encoded_data`, `populate_metadata()`,
`populate_timestamps()`, the reference's error keys (`download`, `remux`, `timestamps`).  Object-store clients and the mpegts -> mp4
remux (an ffmpeg subprocess) are outside this path: a non-local input or a non-MP4 container is recorded as the error the
reference would record when its client / remux fails, never raised."""

from __future__ import annotations

import pathlib

import numpy as np

from ..data_model import StageTimer
from ..interfaces import CuratorStage, CuratorStageResource

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


class VideoDownloader(CuratorStage):
    """Stage that reads the source video(s) of a task into memory and indexes them."""

    def __init__(self, input_path: str = "", input_s3_profile_name: str = "default", *, verbose: bool = False, log_stats: bool = False) -> None:
        self._timer = StageTimer(self)
        self._input_path = input_path
        self._input_s3_profile_name = input_s3_profile_name
        self._verbose = verbose
        self._log_stats = log_stats

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(cpus=1.0)

    def _download_video_bytes(self, video) -> bool:
        try:
            src = video.input_video
            if isinstance(src, str) and "://" in src:
                msg = f"no storage client for {src.split('://', 1)[0]}:// inputs in this build (local paths only)"
                raise ValueError(msg)  # noqa: TRY301
            video.encoded_data = np.fromfile(pathlib.Path(src), dtype=np.uint8)
        except Exception as e:  # noqa: BLE001
            logger.error(f"Got an exception {e!s} when trying to read {video.input_video}")
            video.errors["download"] = str(e)
            return False
        if not video.encoded_data:
            logger.error(f"video.encoded_data is None for {video.input_video} without exceptions ???")
            video.encoded_data = np.array([], dtype=np.uint8)
        return True

    def process_data(self, tasks):
        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            for video in task.videos:
                with self._timer.time_process():
                    if not self._download_video_bytes(video):
                        continue
                    try:
                        video.populate_metadata()
                    except Exception as e:  # noqa: BLE001
                        logger.warning(f"Failed to extract metadata for {video.input_video}: {e}")
                        continue
                    if video.metadata.video_codec is None:
                        logger.warning(f"Codec could not be extracted for {video.input_video}!")
                    if video.metadata.pixel_format is None:
                        logger.warning(f"Pixel format could not be extracted for {video.input_video}!")
                    try:
                        video.populate_timestamps()
                    except Exception as e:  # noqa: BLE001
                        video.errors["timestamps"] = str(e)
                        logger.error(f"Failed to populate timestamps for {video.input_video}: {e}")
                    if self._verbose:
                        m = video.metadata
                        logger.info(f"Downloaded {video.input_video} size={video.encoded_data.nbytes:,}B res={m.width}x{m.height} fps={m.framerate} "
                                    f"duration={m.duration} weight={video.weight:.2f} bit_rate={m.bit_rate_k}K.")  # fmt: skip
            if self._log_stats:
                stage_name, stage_perf_stats = self._timer.log_stats()
                task.stage_perf[stage_name] = stage_perf_stats
        return tasks
