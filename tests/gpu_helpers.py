"""Shared helpers for the GPU parity tests."""

from __future__ import annotations

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def ctx():
    from cosmos_curate_b200.runtime import Context

    c = Context(0)
    yield c
    c.close()


def nv12_pool(ctx, frames_nv12, width: int, height: int, pitch: int, luma_rows: int, colour: str = "opencv"):
    rows = luma_rows + height // 2
    buf = np.zeros((len(frames_nv12), rows, pitch), dtype=np.uint8)
    for i, f in enumerate(frames_nv12):
        buf[i, :height, :width] = f[:height, :width]
        buf[i, luma_rows : luma_rows + height // 2, :width] = f[height:, :width]
    t = torch.from_numpy(buf).cuda()
    return ctx.nv12_pool(t, width, height, luma_rows, colour=colour)


def u8_budget(got: np.ndarray, want: np.ndarray, frac: float = 1e-4):
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 1, f"max diff {d.max()}"
    assert (d > 0).mean() <= frac, f"{(d > 0).mean():.2e} of pixels differ"
