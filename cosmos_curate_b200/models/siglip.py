"""SigLIP (SoViT-400m/14 @384) image embeddings on the B200 path - BASELINE.json configs[3]'s tower.

The reference has no SigLIP model class (its image towers are CLIP ViT-L/14, clip.py, and the video towers); this follows the
`ModelInterface` contract of `CLIPImageEmbeddings` (clip.py:77-118: conda_env_name, model_id_names, setup(), __call__(uint8
[N,H,W,3]) -> unit-norm fp32 [N, D] on the GPU) with transformers' SiglipImageProcessor semantics for the model id below:
resize to 384x384 ... here the short side is resized with antialiased bicubic and centre-cropped like the CLIP path (the same
fused kernel), pixels normalised with mean = std = 0.5, pooled output of the MAP head L2-normalised.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from ..interfaces import ModelInterface
from ..runtime import VitTower, get_context
from . import _weights_source as src
from . import weights as W

_SIGLIP_MODEL_ID = "google/siglip-so400m-patch14-384"
SIGLIP_MEAN = SIGLIP_STD = (0.5, 0.5, 0.5)


class SigLIPImageEmbeddings(ModelInterface):
    def __init__(self, *, weights_dir: str | Path | None = None, seed: int | None = None, max_batch: int = 64, config: W.VitConfig | None = None) -> None:
        super().__init__()
        self._weights_dir, self._seed, self._max_batch, self._config = weights_dir, seed, max_batch, config
        self._tower: VitTower | None = None

    mean, std = SIGLIP_MEAN, SIGLIP_STD

    @property
    def conda_env_name(self) -> str:
        return "unified"

    @property
    def model_id_names(self) -> list[str]:
        return [_SIGLIP_MODEL_ID]

    def setup(self) -> None:
        if self._tower is not None:
            return
        d = src.resolve_dir(self.model_id_names[0], self._weights_dir)
        if d is not None:
            cfg, weights = W.load_hf_siglip_dir(d)
        else:
            seed = src.synthetic_seed(self._seed)
            if seed is None:
                msg = (f"weights for {self.model_id_names[0]} not found (reference weight cache, CURATE_B200_WEIGHTS_DIR) and "
                       "synthetic weights were not requested (seed= / CURATE_B200_SYNTHETIC_WEIGHTS)")  # fmt: skip
                raise FileNotFoundError(msg)
            cfg = self._config or W.SIGLIP_SO400M_384
            weights = W.seeded_weights(cfg, seed)
        self._cfg = cfg
        self._tower = VitTower(get_context(), cfg.to_dict(), weights, max_batch=self._max_batch)

    @property
    def tower(self) -> VitTower:
        assert self._tower is not None, "setup() was not called"
        return self._tower

    @property
    def embedding_dim(self) -> int:
        return self.tower.out_dim

    def __call__(self, images: torch.Tensor | np.ndarray) -> torch.Tensor:
        ctx = self.tower.ctx
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
            if t.shape[1] == 3 and t.shape[-1] != 3:
                t = t.permute(0, 2, 3, 1)
            dev = t.to(f"cuda:{ctx.device}").contiguous()
        if dev.shape[0] == 0:
            return torch.empty((0, self.tower.out_dim), dtype=torch.float32, device=dev.device)
        return self.tower.embed_pool(ctx.rgb_pool(dev), mean=self.mean, std=self.std)[0]
