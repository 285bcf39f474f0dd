"""ImageCLIPEmbeddingStage on the B200 path (image_embedding_stages.py:219-283): image.embeddings["clip"] per task;
images of equal size in one `process_data` call share a batch (the reference embeds one image per model call)."""

from __future__ import annotations

import numpy as np

from ..data_model import StageTimer
from ..interfaces import CuratorStage, CuratorStageResource, ModelInterface
from ..models.clip import CLIPImageEmbeddings


class ImageCLIPEmbeddingStage(CuratorStage):
    def __init__(self, num_gpus_per_worker: float = 0.25, *, verbose: bool = False, log_stats: bool = False, stage_batch_size: int = 1,
                 model: CLIPImageEmbeddings | None = None) -> None:  # fmt: skip
        self._timer = StageTimer(self)
        self._num_gpus_per_worker, self._verbose, self._log_stats = num_gpus_per_worker, verbose, log_stats
        self._stage_batch_size = stage_batch_size
        self._model = model if model is not None else CLIPImageEmbeddings()

    @property
    def model(self) -> ModelInterface:
        return self._model

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(gpus=self._num_gpus_per_worker)

    @property
    def stage_batch_size(self) -> int:
        return self._stage_batch_size

    def stage_setup(self) -> None:
        self._model.setup()

    def process_data(self, tasks):
        self._timer.reinit(self, sum(task.get_major_size() for task in tasks))  # before the work (image_embedding_stages.py:262)
        with self._timer.time_process(num_samples=max(1, len(tasks))):
            by_shape: dict[tuple, list] = {}
            for task in tasks:
                image = task.image
                if image.image_data is None or len(image.image_data.frames) == 0:
                    image.errors["clip_embedding"] = "no image_data"
                    continue
                frame = image.image_data.frames[0]
                by_shape.setdefault(tuple(frame.shape), []).append((image, frame))
            for items in by_shape.values():
                emb = self._model(np.stack([f for _, f in items])).cpu().numpy()
                for (image, _), e in zip(items, emb):
                    image.embeddings["clip"] = e
        if self._log_stats:
            stage_name, stats = self._timer.log_stats()
            for task in tasks:
                task.stage_perf[stage_name] = stats
        return tasks
