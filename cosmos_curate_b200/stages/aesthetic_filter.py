"""AestheticFilterStage on the B200 path - same name, constructor and task mutations as the reference stage
(cosmos_curate/pipelines/video/filtering/aesthetics/aesthetic_filter_stages.py:41-221).

Differences that do not change results: frames of ALL clips of ALL tasks in one `process_data` call are scored in
shared batches (the reference calls the model once per clip, :181-183, batch ~11) and there is one device->host
copy per batch instead of one `.cpu()` per clip.
"""

from __future__ import annotations

from typing import Literal

import numpy as np

from ..data_model import StageTimer
from ..interfaces import CuratorStage, CuratorStageResource, ModelInterface
from ..models.clip_aesthetics import CLIPAestheticScorer
from ..sampling import FrameExtractionPolicy, FrameExtractionSignature

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


def score_frame_groups(model: CLIPAestheticScorer, groups: list[np.ndarray], max_batch: int) -> list[np.ndarray]:
    """Per-frame scores for each uint8 [n_i,H,W,3] group, batching across groups of equal frame size."""
    out: list[np.ndarray | None] = [None] * len(groups)
    by_shape: dict[tuple, list[int]] = {}
    for i, g in enumerate(groups):
        by_shape.setdefault(tuple(g.shape[1:]), []).append(i)
    for idxs in by_shape.values():
        start = 0
        while start < len(idxs):  # greedy fill up to max_batch frames
            end, n = start, 0
            while end < len(idxs) and (n == 0 or n + len(groups[idxs[end]]) <= max_batch):
                n += len(groups[idxs[end]])
                end += 1
            batch = np.concatenate([groups[i] for i in idxs[start:end]], axis=0)
            scores = model(batch).cpu().numpy()
            off = 0
            for i in idxs[start:end]:
                out[i] = scores[off : off + len(groups[i])]
                off += len(groups[i])
            start = end
    return out  # type: ignore[return-value]


class AestheticFilterStage(CuratorStage):
    """Stage for filtering video clips based on aesthetic score."""

    def __init__(  # noqa: PLR0913
        self,
        score_threshold: float,
        reduction: Literal["mean", "min"] = "min",
        target_fps: float = 1.0,
        num_gpus_per_worker: float = 0.25,
        *,
        verbose: bool = False,
        log_stats: bool = False,
        max_batch: int = 256,
        stage_batch_size: int = 1,
        model: CLIPAestheticScorer | None = None,
    ) -> None:
        self._timer = StageTimer(self)
        self._score_threshold = score_threshold
        self._reduction = reduction
        self._reduce_fn = np.min
        self._frame_extraction_signature = FrameExtractionSignature(extraction_policy=FrameExtractionPolicy.sequence, target_fps=target_fps).to_str()
        self._num_gpus_per_worker = num_gpus_per_worker
        self._verbose = verbose
        self._log_stats = log_stats
        self._max_batch = max_batch
        self._stage_batch_size = stage_batch_size
        self._model = model if model is not None else CLIPAestheticScorer(max_batch=max_batch)
        self._process_count = 0

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(gpus=self._num_gpus_per_worker)

    @property
    def model(self) -> ModelInterface:
        return self._model

    @property
    def stage_batch_size(self) -> int:
        return self._stage_batch_size

    def stage_setup(self) -> None:
        self._model.setup()
        if self._reduction == "mean":
            self._reduce_fn = np.mean
        elif self._reduction == "min":
            self._reduce_fn = np.min
        else:
            error_msg = f"Reduction `{self._reduction}` not implemented."
            raise NotImplementedError(error_msg)

    def destroy(self) -> None:
        return

    def process_data(self, tasks):
        # StageTimer call pattern of the reference (aesthetic_filter_stages.py:161-203): reinit BEFORE the work so that
        # process_time covers the model calls.  Clips of all tasks of the call share batches, so the window is per call.
        self._timer.reinit(self, sum(task.get_major_size() for task in tasks))
        work: list[tuple[object, np.ndarray]] = []  # (clip, frames) in task order
        n_clips = sum(len(task.video.clips) for task in tasks)
        with self._timer.time_process(num_samples=max(1, n_clips)):
            for task in tasks:
                for clip in task.video.clips:
                    if not clip.encoded_data:
                        logger.warning(f"Clip {clip.uuid} has no encoded_data.")
                        clip.errors["encoded_data"] = "empty"
                        clip.aesthetic_score = -1.0
                        continue
                    ef = clip.extracted_frames.resolve()
                    if ef is None or self._frame_extraction_signature not in ef:
                        clip.errors[f"frames-{self._frame_extraction_signature}"] = "missing"
                        logger.error(f"Clip {clip.uuid} has buffer but no extracted frames for {self._frame_extraction_signature}")
                        clip.aesthetic_score = -1.0
                        continue
                    frames = ef.pop(self._frame_extraction_signature)  # pop: other consumers own the other keys
                    if not ef:
                        clip.extracted_frames.drop()
                    work.append((clip, frames))
            if work:
                per_clip = score_frame_groups(self._model, [f for _, f in work], self._max_batch)
                for (clip, _), scores in zip(work, per_clip):
                    clip.aesthetic_score = float(self._reduce_fn(scores))

            for task in tasks:
                video = task.video
                passed = []
                for clip in video.clips:
                    if clip.aesthetic_score < self._score_threshold:
                        video.filtered_clips.append(clip)
                        video.clip_stats.num_filtered_by_aesthetic += 1
                        if self._verbose:
                            logger.info(f"Clip {clip.uuid} has aesthetic score {clip.aesthetic_score:.3f} below threshold {self._score_threshold}, skipped.")
                    else:
                        passed.append(clip)
                        if self._verbose:
                            logger.info(f"Clip {clip.uuid} has aesthetic score {clip.aesthetic_score:.3f} above threshold {self._score_threshold}, kept.")
                video.clips = passed
        if self._log_stats:
            stage_name, stats = self._timer.log_stats()
            for task in tasks:
                task.stage_perf[stage_name] = stats
        self._process_count += 1
        return tasks
