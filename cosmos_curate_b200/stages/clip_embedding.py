"""ClipFrameEmbeddingStage: the LOCAL producer of `clip.openai_embedding` - the generic "frames -> one vector per clip" slot the
reference fills through a remote OpenAI-compatible endpoint (OpenAIEmbeddingStage,
cosmos_curate/pipelines/video/embedding/openai_embedding_stage.py:47-190) and ClipWriterStage persists with
`--embedding-algorithm openai` (metadata_writer_stage.py:745-752).  Same input slot (`clip.extracted_frames[signature]` at
`target_fps`, default 2.0), same error key and message (`clip.errors["openai_embedding"] = "extracted frames missing"`), same
`extracted_frames.drop()` after a successful embedding; the vectors come from the image tower of this repo instead of a network
call.  The reference has no rule for pooling per-frame embeddings into a clip vector: mean over frames, then L2 re-normalisation
(the same choice as the fused NvdecClipAestheticStage).

Frames of all clips of all tasks of a call share tower batches (one device -> host copy per batch)."""

from __future__ import annotations

import numpy as np

from ..data_model import StageTimer
from ..interfaces import CuratorStage, CuratorStageResource, ModelInterface
from ..sampling import FrameExtractionPolicy, FrameExtractionSignature
from .aesthetic_filter import score_frame_groups

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


def pool_clip_embedding(per_frame: np.ndarray) -> np.ndarray:
    """[n, D] unit-norm frame embeddings -> [D] float32: mean, then L2 norm."""
    m = np.asarray(per_frame, dtype=np.float32).mean(axis=0)
    return (m / np.linalg.norm(m)).astype(np.float32)


class ClipFrameEmbeddingStage(CuratorStage):
    """Generate one embedding per clip from its pre-extracted frames on the local image tower."""

    def __init__(self, *, model_name: str = "openai/clip-vit-large-patch14", target_fps: float = 2.0, num_gpus_per_worker: float = 0.25,
                 max_batch: int = 256, stage_batch_size: int = 1, verbose: bool = False, log_stats: bool = False, model: ModelInterface | None = None) -> None:  # fmt: skip
        self._timer = StageTimer(self)
        self._model_name = model_name
        self._num_gpus, self._max_batch, self._stage_batch_size = num_gpus_per_worker, max_batch, stage_batch_size
        self._verbose, self._log_stats = verbose, log_stats
        self._frame_extraction_signature = FrameExtractionSignature(extraction_policy=FrameExtractionPolicy.sequence, target_fps=target_fps).to_str()
        if model is None:
            from ..models.clip import CLIPImageEmbeddings
            from ..models.siglip import SigLIPImageEmbeddings

            model = SigLIPImageEmbeddings(max_batch=max_batch) if "siglip" in model_name.lower() else CLIPImageEmbeddings(max_batch=max_batch)
        self._model = model

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(gpus=self._num_gpus)

    @property
    def model(self) -> ModelInterface:
        return self._model

    @property
    def stage_batch_size(self) -> int:
        return self._stage_batch_size

    def process_data(self, tasks):
        self._timer.reinit(self, sum(task.get_major_size() for task in tasks))
        n_clips = sum(len(task.video.clips) for task in tasks)
        with self._timer.time_process(max(1, n_clips)):
            work = []
            for task in tasks:
                for clip in task.video.clips:
                    ef = clip.extracted_frames.resolve()
                    if ef is None or self._frame_extraction_signature not in ef:
                        clip.errors["openai_embedding"] = "extracted frames missing"
                        logger.error(f"Clip {clip.uuid} has no extracted frames for {self._frame_extraction_signature}")
                        continue
                    work.append((clip, ef[self._frame_extraction_signature]))
            try:
                per_clip = score_frame_groups(self._model, [f for _, f in work], self._max_batch) if work else []
            except Exception as exc:  # noqa: BLE001 - per-item convention of the reference stage (:160-167): record, do not raise
                for clip, _ in work:
                    clip.errors["openai_embedding"] = str(exc)
                logger.warning(f"local embedding failed for {len(work)} clips: {exc}")
                per_clip, work = [], []
            for (clip, frames), emb in zip(work, per_clip):
                clip.openai_embedding = pool_clip_embedding(emb)
                if self._verbose:
                    logger.info(f"embedded clip {clip.uuid}: {len(frames)} frames, shape={clip.openai_embedding.shape}")
                clip.extracted_frames.drop()
        if self._log_stats:
            stage_name, stats = self._timer.log_stats()
            for task in tasks:
                task.stage_perf[stage_name] = stats
        return tasks
