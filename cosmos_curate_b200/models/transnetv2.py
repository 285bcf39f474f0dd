"""TransNetV2 shot-transition model on the B200 path.  Drop-in for cosmos_curate/models/transnetv2.py:530-580.

The network (rf=16, rl=3, rs=2, rd=1024 with frame-similarity and colour-histogram branches: the only configuration
the reference builds, :563) runs in fp32 inside libcurate_b200 (cb_transnet_*); this class only finds the weights and
hands the reference's state_dict over under its own key names.
"""

from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import torch

from ..interfaces import ModelInterface
from ..runtime import ShotNet, get_context
from . import _weights_source as src

_TRANSNETV2_MODEL_ID = "Sn4kehead/TransNetV2"
_TRANSNETV2_MODEL_WEIGHTS = "transnetv2-pytorch-weights.pth"
_DILATIONS = (1, 2, 4, 8)


def seeded_state_dict(seed: int = 0) -> dict[str, np.ndarray]:
    """Synthetic weights with the reference's key names and shapes (benchmarks / tests only; never a silent substitute)."""
    rng = np.random.default_rng(seed)
    sd: dict[str, np.ndarray] = {}

    def rn(shape, std):
        return (rng.standard_normal(shape) * std).astype(np.float32)

    for s in range(3):
        f = 16 << s
        stack_in = 3 if s == 0 else (16 << (s - 1)) * 4
        for b in range(2):
            cin = stack_in if b == 0 else 4 * f
            p = f"SDDCNN.{s}.DDCNN.{b}"
            for d in _DILATIONS:
                sd[f"{p}.Conv3D_{d}.layers.0.weight"] = rn((2 * f, cin, 1, 3, 3), math.sqrt(2.0 / (cin * 9)))
                sd[f"{p}.Conv3D_{d}.layers.1.weight"] = rn((f, 2 * f, 3, 1, 1), math.sqrt(1.0 / (2 * f * 3)))
            c = 4 * f
            sd[f"{p}.bn.weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
            sd[f"{p}.bn.bias"] = rn((c,), 0.1)
            sd[f"{p}.bn.running_mean"] = rn((c,), 0.1)
            sd[f"{p}.bn.running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
            sd[f"{p}.bn.num_batches_tracked"] = np.array(0, dtype=np.int64)
    sd["frame_sim_layer.projection.weight"] = rn((128, 448), 1.0 / math.sqrt(448))
    sd["frame_sim_layer.projection.bias"] = rn((128,), 0.05)
    sd["frame_sim_layer.fc.weight"] = rn((128, 101), 1.0 / math.sqrt(101))
    sd["frame_sim_layer.fc.bias"] = rn((128,), 0.05)
    sd["color_hist_layer.fc.weight"] = rn((128, 101), 1.0 / math.sqrt(101))
    sd["color_hist_layer.fc.bias"] = rn((128,), 0.05)
    sd["fc1.weight"] = rn((1024, 4864), 1.0 / math.sqrt(4864))
    sd["fc1.bias"] = rn((1024,), 0.05)
    sd["cls_layer1.weight"] = rn((1, 1024), 2.0 / math.sqrt(1024))
    sd["cls_layer1.bias"] = rn((1,), 0.05) - np.float32(4.75)  # centres the seeded logits around the decision threshold
    sd["cls_layer2.weight"] = rn((1, 1024), 2.0 / math.sqrt(1024))
    sd["cls_layer2.bias"] = rn((1,), 0.05)
    return sd


class TransNetV2(ModelInterface):
    """Interface for TransNetV2 shot transition detection model."""

    def __init__(self, *, weights_dir: str | Path | None = None, seed: int | None = None, max_windows: int = 16) -> None:
        super().__init__()
        self._weights_dir, self._seed, self._max_windows = weights_dir, seed, max_windows
        self._net: ShotNet | None = None

    @property
    def conda_env_name(self) -> str:
        return "unified"

    @property
    def model_id_names(self) -> list[str]:
        return [_TRANSNETV2_MODEL_ID]

    def load(self) -> dict:
        """Host part of setup(): the state_dict (transnetv2.py:560-566)."""
        d = src.resolve_dir(_TRANSNETV2_MODEL_ID, self._weights_dir)
        if d is not None:
            model_file = d / _TRANSNETV2_MODEL_WEIGHTS
            if not model_file.exists():
                error_msg = f"{model_file} not found!"
                raise FileNotFoundError(error_msg)
            sd = torch.load(model_file.as_posix(), map_location="cpu", weights_only=True)
            return {k: v.detach().to(torch.float32).numpy() if v.is_floating_point() else v.numpy() for k, v in sd.items()}
        seed = src.synthetic_seed(self._seed)
        if seed is None:
            msg = f"weights for {_TRANSNETV2_MODEL_ID} not found and synthetic weights were not requested"
            raise FileNotFoundError(msg)
        return seeded_state_dict(seed)

    def setup(self) -> None:
        self._ctx = get_context()
        self._net = ShotNet(self._ctx, self.load(), max_windows=self._max_windows)

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        """[# batch, # frames, 27, 48, RGB] uint8 -> [# batch, # frames, 1] transition probabilities."""
        assert isinstance(inputs, torch.Tensor), "inputs must be a torch.Tensor"
        assert list(inputs.shape[2:]) == [27, 48, 3], f"incorrect shape: expected [*, *, 27, 48, 3], got {inputs.shape}"
        assert inputs.dtype == torch.uint8, f"incorrect dtype: expected torch.uint8, got {inputs.dtype}"
        return self._net.forward(inputs.to(f"cuda:{self._ctx.device}"))

    def predict_video(self, frames: torch.Tensor | np.ndarray) -> torch.Tensor:
        """[n, 27, 48, 3] uint8 (host or device) -> fp32 cuda [n]: every window of the video in one library call."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        return self._net.predict(frames.to(f"cuda:{self._ctx.device}", non_blocking=True))
