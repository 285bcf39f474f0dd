"""CLIP embeddings chained with the aesthetic head.  Drop-in for cosmos_curate/models/clip_aesthetics.py:27-76.

One fused forward: the tower's tail kernel emits the score next to the embedding, so `__call__(images)` is a single
pass (the reference runs two modules and five cuBLAS GEMMs for the head).
"""

from __future__ import annotations

import numpy as np
import torch

from ..interfaces import ModelInterface
from .aesthetics import _AESTHETICS_MODEL_ID, AestheticScorer
from .clip import _CLIP_MODEL_ID, CLIPImageEmbeddings


class CLIPAestheticScorer(ModelInterface):
    """A model that chains CLIPImageEmbeddings and AestheticScorer models."""

    def __init__(self, *, clip_weights_dir=None, aesthetic_weights_dir=None, seed: int | None = None, max_batch: int = 256, config=None) -> None:
        super().__init__()
        self._kw = {"clip_weights_dir": clip_weights_dir, "aesthetic_weights_dir": aesthetic_weights_dir, "seed": seed, "max_batch": max_batch,
                    "config": config}  # fmt: skip
        self._clip_model: CLIPImageEmbeddings | None = None
        self._aesthetic_model: AestheticScorer | None = None

    @property
    def conda_env_name(self) -> str:
        return "unified"

    @property
    def model_id_names(self) -> list[str]:
        return [_AESTHETICS_MODEL_ID, _CLIP_MODEL_ID]

    def setup(self) -> None:
        if self._clip_model is not None:  # idempotent: several stages of one actor process may share the model object
            return
        kw = self._kw
        dim = (kw["config"].proj_dim or kw["config"].hidden) if kw["config"] is not None else 768
        self._aesthetic_model = AestheticScorer(weights_dir=kw["aesthetic_weights_dir"], seed=kw["seed"], dim=dim)
        w, b = self._aesthetic_model.load()
        self._clip_model = CLIPImageEmbeddings(weights_dir=kw["clip_weights_dir"], seed=kw["seed"], max_batch=kw["max_batch"], config=kw["config"],
                                               aesthetic=(w, b))  # fmt: skip
        self._clip_model.setup()

    @property
    def tower(self):
        assert self._clip_model is not None
        return self._clip_model._tower

    def embed_and_score(self, images) -> tuple[torch.Tensor, torch.Tensor]:
        assert self._clip_model is not None
        return self._clip_model._embed(images)

    def __call__(self, images: torch.Tensor | np.ndarray) -> torch.Tensor:
        return self.embed_and_score(images)[1]
