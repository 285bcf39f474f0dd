"""TransNetV2ClipExtractionStage on the B200 path.

Drop-in for cosmos_curate/pipelines/video/clipping/transnetv2_extraction_stages.py:39-212: same constructor, same
task mutations (`video.clips` gets one `Clip(uuid5, source_video, span seconds)` per kept shot, `video.frame_array`
is dropped, `task.stage_perf`), same skip/raise behaviour.  Differences underneath: one library call per video instead
of one model call + H2D copy per 100-frame window, and `video.frame_array` may already be a CUDA tensor (left there by
VideoFrameExtractionStage(keep_on_device=True)), in which case the frames never visit the host.
"""

from __future__ import annotations

import logging

import numpy as np
import torch

from .. import shots
from ..data_model import Clip, StageTimer
from ..interfaces import CuratorStage, CuratorStageResource, ModelInterface
from ..models.transnetv2 import TransNetV2

logger = logging.getLogger(__name__)


class TransNetV2ClipExtractionStage(CuratorStage):
    def __init__(self, threshold: float = 0.4, min_length_s: float | None = 2.0, min_length_frames: int | None = 48, max_length_s: float | None = 60.0,
                 max_length_mode: str = "stride", crop_s: float | None = 0.5, *, entire_scene_as_clip: bool = True, num_gpus_per_worker: float = 0.25,
                 limit_clips: int = 0, verbose: bool = False, log_stats: bool = False, model: TransNetV2 | None = None) -> None:  # fmt: skip
        super().__init__()
        self._timer = StageTimer(self)
        self.threshold = threshold
        self.min_length_s, self.min_length_frames, self.max_length_s = min_length_s, min_length_frames, max_length_s
        if self.min_length_s and self.max_length_s and self.max_length_s < self.min_length_s:
            error_msg = "Max length is smaller than min length!"
            raise ValueError(error_msg)
        self.max_length_mode = max_length_mode
        self.crop_s = crop_s
        self.entire_scene_as_clip = entire_scene_as_clip
        self._num_gpus_per_worker, self._limit_clips, self._verbose, self._log_stats = num_gpus_per_worker, limit_clips, verbose, log_stats
        self._model = model if model is not None else TransNetV2()

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(gpus=self._num_gpus_per_worker)

    @property
    def model(self) -> ModelInterface:
        return self._model

    def stage_setup(self) -> None:
        self._model.setup()

    def _assign_clips(self, video, frames) -> None:
        """frames [n,27,48,3] uint8 (numpy or cuda tensor) -> video.clips (transnetv2_extraction_stages.py:185-209)."""
        fps = video.metadata.framerate
        prob = self._model.predict_video(frames).cpu().numpy()
        predictions = shots.predictions_from_probabilities(prob, self.threshold)
        scenes = shots.scenes_from_predictions(predictions, entire_scene_as_clip=self.entire_scene_as_clip)
        if self._verbose:
            logger.info(f"{video.input_video} returned {scenes.shape[0]} scenes")
        mn, mx, crop = shots.stage_lengths(fps, self.min_length_s, self.min_length_frames, self.max_length_s, self.crop_s)
        filtered = shots.filter_scenes(scenes, min_length=mn, max_length=mx, max_length_mode=self.max_length_mode, crop_length=crop)
        if self._verbose:
            logger.info(f"{video.input_video} returned {filtered.shape[0]} filtered scenes")
        for uid, span in shots.clips_from_scenes(str(video.input_video), filtered, fps, self._limit_clips):
            video.clips.append(Clip(uuid=uid, source_video=str(video.input_video), span=span))

    def process_data(self, tasks):
        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            video = task.video
            if not video.has_metadata():
                logger.warning(f"Incomplete metadata for {video.input_video}. Skipping...")
                continue
            if not video.frame_array:
                logger.warning(f"No frame array for {video.input_video}. Skipping...")
                continue
            with self._timer.time_process():
                frames = video.frame_array.resolve()
                if frames is None:
                    msg = f"frame_array resolved to None for {video.input_video}"
                    raise ValueError(msg)
                if tuple(frames.shape[1:4]) != (27, 48, 3):
                    error_msg = f"Expected frames of shape 27x48x3, got {frames.shape[1:4]}."
                    raise ValueError(error_msg)
                self._assign_clips(video, frames)
                video.frame_array.drop()
                if not video.clips:
                    logger.warning(f"No scene cut predicted for {video.input_video}.")
            if self._log_stats:
                stage_name, stats = self._timer.log_stats()
                task.stage_perf[stage_name] = stats
        return tasks


class NvdecShotDetectionStage(TransNetV2ClipExtractionStage):
    """VideoFrameExtractionStage + TransNetV2ClipExtractionStage in one GPU pass (recommended on B200).

    Decodes every frame of `video.encoded_data` on NVDEC straight to 27x48 RGB thumbnails in HBM
    (cb_decoder_decode_thumbnails, the pynvc path of frame_extraction_stages.py:71-204), feeds them to the shot network
    without a host round trip and writes `video.clips`.  `video.frame_array` is never materialised.  Metadata is filled
    from the moov index when the task arrives without it (the reference's downloader stage normally does that)."""

    def __init__(self, *args, num_gpus_per_worker: float = 1.0, **kwargs) -> None:
        super().__init__(*args, num_gpus_per_worker=num_gpus_per_worker, **kwargs)

    def stage_setup(self) -> None:
        from ..runtime import SessionTable, get_context

        super().stage_setup()
        self._sessions = SessionTable(get_context())  # one NVDEC session per source resolution

    def destroy(self) -> None:
        if getattr(self, "_sessions", None):
            self._sessions.close()

    def process_data(self, tasks):
        from .._lib import CurateB200Error
        from ..runtime import decode_thumbnails, mp4_index

        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            video = task.video
            data = video.encoded_data.resolve()
            if data is None:
                error_msg = "Please load video bytes!"
                raise ValueError(error_msg)
            with self._timer.time_process():
                try:
                    if not video.has_metadata():
                        video.populate_metadata()
                    idx = mp4_index(data)
                    frames = decode_thumbnails(self._sessions.get((idx["width"], idx["height"])), data, 48, 27, idx["n_samples"])
                except (CurateB200Error, KeyError) as e:
                    logger.error(f"Video frame extraction failed on {video.input_video}: {e}")
                    video.errors["frame_extraction"] = "null"
                    continue
                self._assign_clips(video, frames)
                if not video.clips:
                    logger.warning(f"No scene cut predicted for {video.input_video}.")
            if self._log_stats:
                stage_name, stats = self._timer.log_stats()
                task.stage_perf[stage_name] = stats
        return tasks
