"""CLIP image embeddings on the B200 path.

Drop-in for cosmos_curate/models/clip.py: `CLIPImageEmbeddings(ModelInterface)` with the same
`conda_env_name`, `model_id_names`, `setup()` and `__call__(images) -> Tensor[N, 768]` (unit-norm fp32 on the
GPU, clip.py:64-74).  Underneath: H2D of the uint8 frames, the fused resize/crop/normalise kernel and the
hand-written tower in libcurate_b200 - no torchvision, no transformers, no torch compute.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from ..interfaces import ModelInterface
from ..runtime import CLIP_MEAN, CLIP_STD, VitTower, get_context
from . import _weights_source as src
from . import weights as W

_CLIP_MODEL_ID = "openai/clip-vit-large-patch14"


class CLIPImageEmbeddings(ModelInterface):
    """Interface for generating CLIP image embeddings from input images."""

    def __init__(self, *, weights_dir: str | Path | None = None, seed: int | None = None, max_batch: int = 256,
                 config: W.VitConfig | None = None, aesthetic: tuple | None = None) -> None:  # fmt: skip
        super().__init__()
        self._weights_dir, self._seed, self._max_batch = weights_dir, seed, max_batch
        self._config = config
        self._aesthetic = aesthetic
        self._tower: VitTower | None = None

    @property
    def conda_env_name(self) -> str:
        return "unified"

    @property
    def model_id_names(self) -> list[str]:
        return [_CLIP_MODEL_ID]

    def setup(self) -> None:
        ctx = get_context()
        d = src.resolve_dir(self.model_id_names[0], self._weights_dir)
        if d is not None:
            cfg, weights = W.load_hf_clip_dir(d)
        else:
            seed = src.synthetic_seed(self._seed)
            if seed is None:
                msg = (f"weights for {self.model_id_names[0]} not found (reference weight cache, CURATE_B200_WEIGHTS_DIR) and "
                       "synthetic weights were not requested (seed= / CURATE_B200_SYNTHETIC_WEIGHTS)")  # fmt: skip
                raise FileNotFoundError(msg)
            cfg = self._config or W.CLIP_VIT_L14
            weights = W.seeded_weights(cfg, seed)
        self._cfg = cfg
        self._tower = VitTower(ctx, cfg.to_dict(), weights, max_batch=self._max_batch, aesthetic=self._aesthetic)

    @property
    def embedding_dim(self) -> int:
        assert self._tower is not None
        return self._tower.out_dim

    def _embed(self, images):
        assert self._tower is not None, "setup() was not called"
        ctx = self._tower.ctx
        if isinstance(images, np.ndarray):
            if images.ndim != 4 or images.shape[-1] != 3 or images.dtype != np.uint8:
                msg = f"expected uint8 [N,H,W,3] frames, got {images.dtype} {images.shape}"
                raise ValueError(msg)
            dev = torch.from_numpy(np.ascontiguousarray(images)).to(f"cuda:{ctx.device}", non_blocking=True)
        else:
            t = images
            if t.dim() != 4 or t.dtype != torch.uint8:
                msg = f"expected uint8 [N,C,H,W] or [N,H,W,C] tensor, got {t.dtype} {tuple(t.shape)}"
                raise ValueError(msg)
            if t.shape[1] == 3 and t.shape[-1] != 3:  # the reference accepts NCHW tensors (clip.py:64-70)
                t = t.permute(0, 2, 3, 1)
            dev = t.to(f"cuda:{ctx.device}").contiguous()
        if dev.shape[0] == 0:
            z = torch.empty((0, self._tower.out_dim), dtype=torch.float32, device=dev.device)
            return z, torch.empty((0,), dtype=torch.float32, device=dev.device)
        pool = ctx.rgb_pool(dev)
        emb, _, score = self._tower.embed_pool(pool, mean=CLIP_MEAN, std=CLIP_STD)
        return emb, score

    def __call__(self, images: torch.Tensor | np.ndarray) -> torch.Tensor:
        return self._embed(images)[0]
