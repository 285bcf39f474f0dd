"""InternVideo2 input formulation on the B200 path.

Mirror of `InternVideo2MultiModality(utils_only=True)` (cosmos_curate/models/internvideo2_mm.py:335-443) - the object the
reference's InternVideo2FrameCreationStage holds (internvideo2_stages.py:73) - restricted to what that stage calls:
`get_target_num_frames()` (:417-424), `formulate_input_frames(frames)` (:426-438 -> `_construct_frames` :390-405) and
`formulate_input_image(frame)` (:440-443 -> `_construct_image` :407-415).  The tower itself is not part of this path.

Underneath: only the `fnum` frames the stride keeps are uploaded; cv2.resize's fixed-point bilinear and the float32
ImageNet normalisation run in one kernel of libcurate_b200 (cb_video_tube), bit-identical to the reference's numpy result.
"""

from __future__ import annotations

import numpy as np
import torch

from ..interfaces import ModelInterface
from ..runtime import IMAGENET_MEAN, IMAGENET_STD, Pool, get_context

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


def select_frame_ids(n_frames: int, fnum: int) -> list[int]:
    """`vid_list[::step][:fnum]` with step = len // fnum (internvideo2_mm.py:399-400); [] when n_frames < fnum."""
    if n_frames < fnum:
        return []
    step = n_frames // fnum
    return list(range(0, n_frames, step))[:fnum]


class InternVideo2FrameFormulator(ModelInterface):
    """utils-only InternVideo2 interface: frame selection, resize, normalisation -> float32 [1, T, 3, size, size]."""

    def __init__(self, *, num_frames: int = 8, size_t: int = 224) -> None:
        super().__init__()
        self._num_frames, self._size_t = num_frames, size_t  # config "num_frames" / "size_t" defaults (:424, :437)
        self._ctx = None

    @property
    def conda_env_name(self) -> str:
        return "unified"

    @property
    def model_id_names(self) -> list[str]:
        return []  # no weights: the formulation has no parameters (the reference lists the tower's ids even with utils_only)

    def setup(self) -> None:
        self._ctx = get_context()

    def get_target_num_frames(self) -> int:
        return self._num_frames

    def _device(self) -> str:
        if self._ctx is None:
            self.setup()
        return f"cuda:{self._ctx.device}"

    def formulate_pool(self, pool: Pool, slots) -> torch.Tensor:
        """Frames already in HBM (NV12 decode surfaces or RGB): -> float32 cuda [len(slots), 3, size, size]."""
        if self._ctx is None:
            self.setup()
        return self._ctx.video_tube(pool, self._size_t, self._size_t, slots=slots, mean=IMAGENET_MEAN, std=IMAGENET_STD)

    def _formulate_host(self, picked: list[np.ndarray]) -> np.ndarray:
        dev = self._device()
        by_shape: dict[tuple, list[int]] = {}
        for i, f in enumerate(picked):
            if f.ndim != 3 or f.shape[-1] != 3 or f.dtype != np.uint8:
                msg = f"expected uint8 [H,W,3] frames, got {f.dtype} {f.shape}"
                raise ValueError(msg)
            by_shape.setdefault(f.shape, []).append(i)
        res: list[torch.Tensor | None] = [None] * len(picked)
        for idxs in by_shape.values():  # the reference resizes frame by frame, so mixed sizes are legal
            batch = torch.from_numpy(np.ascontiguousarray(np.stack([picked[i] for i in idxs]))).to(dev, non_blocking=True)
            t = self.formulate_pool(self._ctx.rgb_pool(batch), None)
            for k, i in enumerate(idxs):
                res[i] = t[k]
        out = torch.stack(res)  # [T, 3, size, size]
        return out.unsqueeze(0).cpu().numpy()

    def formulate_input_frames(self, frames) -> np.ndarray:
        """list / array of uint8 [H,W,3] frames -> float32 [1, fnum, 3, size, size]; empty float32 array if too few frames."""
        fn = self.get_target_num_frames()
        if len(frames) < fn:
            logger.error(f"Frame count {len(frames)} is smaller than minimal requirement {fn}")
            return np.empty(0, dtype=np.float32)
        return self._formulate_host([np.asarray(frames[i]) for i in select_frame_ids(len(frames), fn)])

    def formulate_input_image(self, frame: np.ndarray) -> np.ndarray:
        """one uint8 [H,W,3] frame -> float32 [1, 1, 3, size, size]."""
        return self._formulate_host([np.asarray(frame)])
