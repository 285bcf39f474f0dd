"""Aesthetic scorer on the B200 path.  Drop-in for cosmos_curate/models/aesthetics.py:109-155.

The reference MLP (768 -> 1024 -> 128 -> 64 -> 16 -> 1, Dropout only, aesthetics.py:44-53) has no activation, so it
is folded once in float64 to score = w . e + b; the dot products run in libcurate_b200 (cb_affine_score).
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from ..interfaces import ModelInterface
from ..runtime import affine_score, get_context
from . import _weights_source as src
from . import weights as W

_AESTHETICS_MODEL_ID = "ttj/sac-logos-ava1-l14-linearMSE"


class AestheticScorer(ModelInterface):
    """Public interface for aesthetic scoring of video embeddings."""

    def __init__(self, *, weights_dir: str | Path | None = None, seed: int | None = None, dim: int = 768) -> None:
        super().__init__()
        self._weights_dir, self._seed, self._dim = weights_dir, seed, dim
        self.w: np.ndarray | None = None
        self.b: float = 0.0

    @property
    def conda_env_name(self) -> str:
        return "unified"

    @property
    def model_id_names(self) -> list[str]:
        return [_AESTHETICS_MODEL_ID]

    def load(self) -> tuple[np.ndarray, float]:
        """Host part of setup(): the folded (w, b)."""
        d = src.resolve_dir(self.model_id_names[0], self._weights_dir)
        if d is not None:
            self.w, self.b = W.load_aesthetic_mlp(d / "model.safetensors")
        else:
            seed = src.synthetic_seed(self._seed)
            if seed is None:
                msg = f"weights for {self.model_id_names[0]} not found and synthetic weights were not requested"
                raise FileNotFoundError(msg)
            self.w, self.b = W.seeded_aesthetic(self._dim, seed)
        return self.w, self.b

    def setup(self) -> None:
        self.load()
        self._ctx = get_context()
        self._w_dev = torch.from_numpy(self.w).to(f"cuda:{self._ctx.device}")

    def __call__(self, embeddings: torch.Tensor | np.ndarray) -> torch.Tensor:
        if isinstance(embeddings, np.ndarray):
            embeddings = torch.from_numpy(embeddings.copy())
        e = embeddings.to(f"cuda:{self._ctx.device}", dtype=torch.float32).contiguous()
        return affine_score(self._ctx, e, self._w_dev, self.b)
