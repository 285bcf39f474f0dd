"""VideoDownloader for local files: drop-in for the first stage of the reference's split pipeline
(cosmos_curate/pipelines/video/read_write/download_stages.py:40-228) - same class name and constructor, same task mutations
(`video.encoded_data`, `video.metadata`, `video.timestamps`) and the same `video.errors` keys (`download`, `timestamps`).

What differs underneath: metadata and per-frame timestamps both come from ONE pass over the MP4 sample tables (`cb_mp4_index`)
instead of an ffprobe subprocess plus a PyAV demux of every packet.  Object-store inputs and the mpegts -> mp4 remux (an ffmpeg
subprocess in the reference) are outside this path: such inputs end up with the error a failed client / remux leaves behind,
nothing is raised."""

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


def read_local_video(source) -> np.ndarray:
    """uint8 view of the file behind `source`; URL-style sources are refused (no storage clients in this build)."""
    text = str(source)
    if "://" in text:
        scheme = text.split("://", 1)[0]
        msg = f"no storage client for {scheme}:// inputs in this build (local paths only)"
        raise ValueError(msg)
    return np.fromfile(pathlib.Path(text), dtype=np.uint8)


class VideoDownloader(CuratorStage):
    """Reads every source video of a task into memory and indexes it."""

    def __init__(self, input_path: str = "", input_s3_profile_name: str = "default", *, verbose: bool = False, log_stats: bool = False) -> None:
        self._timer = StageTimer(self)
        self._input_path, self._input_s3_profile_name = input_path, input_s3_profile_name
        self._verbose, self._log_stats = verbose, log_stats

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(cpus=1.0)

    def _ingest(self, video) -> None:
        """bytes -> metadata -> timestamps; each step records its own failure and stops the chain for this video."""
        try:
            video.encoded_data = read_local_video(video.input_video)
        except Exception as exc:  # noqa: BLE001
            video.errors["download"] = str(exc)
            logger.error(f"cannot read {video.input_video}: {exc}")
            return
        try:
            video.populate_metadata()
        except Exception as exc:  # noqa: BLE001 - the reference logs and moves on without an error key here (:127-141)
            logger.warning(f"Failed to extract metadata for {video.input_video}: {exc}")
            return
        for field in ("video_codec", "pixel_format"):
            if getattr(video.metadata, field) is None:
                logger.warning(f"{field} could not be extracted for {video.input_video}!")
        try:
            video.populate_timestamps()
        except Exception as exc:  # noqa: BLE001
            video.errors["timestamps"] = str(exc)
            logger.error(f"Failed to populate timestamps for {video.input_video}: {exc}")
        if self._verbose:
            m = video.metadata
            logger.info(f"{video.input_video}: {video.encoded_data.nbytes:,} B, {m.width}x{m.height} @ {m.framerate} fps, {m.duration} s, "
                        f"weight {video.weight:.2f}, {m.bit_rate_k} kb/s")  # fmt: skip

    def process_data(self, tasks):
        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            for video in task.videos:
                with self._timer.time_process():
                    self._ingest(video)
            if self._log_stats:
                name, stats = self._timer.log_stats()
                task.stage_perf[name] = stats
        return tasks
