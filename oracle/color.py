"""Oracle: NV12 <-> RGB (byte arithmetic, bit-exact).

nv12_to_rgb restates what the reference's GPU decode path does with every decoded surface,
``cvcuda.cvtcolor_into(..., cvcuda.ColorConversion.YUV2RGB_NV12)``
(cosmos_curate/pipelines/video/utils/nvcodec_utils.py:35-38,146,178): the conversion code is
chosen from the pixel format alone, the stream's colour metadata is ignored.  CV-CUDA
(cvcuda-cu13 0.16.0, pixi.lock:11231) is a closed wheel that cannot be installed here; it
documents its colour codes as OpenCV-compatible, so this follows OpenCV's
``COLOR_YUV2RGB_NV12``: ITU-R BT.601 limited range, 20-bit fixed point, chroma replicated to
the 2x2 luma block (no interpolation), result clamped to u8.  tests/test_oracle_cpu.py pins
this bit-exactly against cv2.cvtColor on random surfaces; parity against CV-CUDA itself is
UNPINNED (stated in DESIGN.md).

nv12_to_rgb_swscale restates what the reference's CPU decode path - the one that feeds CLIP - does with every decoded
frame: ``frame.to_ndarray(format="rgb24")`` (decoder_utils.py:439-451) = libswscale's unscaled yuv420p -> rgb24 converter
(third-party: FFmpeg libswscale, conda-forge build pinned by pixi.lock; cv2 bundles libswscale 9.1 here).  Restated from
libswscale's published code (yuv2rgb.c ff_yuv2rgb_c_init_tables for the coefficients, x86/yuv_2_rgb.asm for the arithmetic):
ITU-R BT.601 limited range whatever the stream says (neither PyAV nor cv2 calls sws_setColorspaceDetails), chroma replicated
to the 2x2 block, 16-bit fixed point with every product TRUNCATED (pmulhw): y' = ((8Y - 128) * 9539) >> 16, r = y' +
((8V - 1024) * 13075 >> 16), ... clamp.  Pinned bit-exactly against cv2.VideoCapture's BGR output over the whole u8 range
(tests/test_oracle_cpu.py decodes lossless I_PCM pictures of known YUV).  aarch64 builds use a NEON converter with other
rounding: this restates the x86 path the reference's deployment runs.

rgb_to_nv12 is only a generator of plausible synthetic surfaces for tests/bench (not on the
reference path).

Test infrastructure only (see oracle/__init__.py).
"""

from __future__ import annotations

import numpy as np

# OpenCV imgproc color_yuv: ITUR_BT_601_* constants, ITUR_BT_601_SHIFT = 20
CY, CUB, CUG, CVG, CVR, SHIFT = 1220542, 2116026, -409993, -852492, 1673527, 20


def nv12_to_rgb(nv12: np.ndarray, height: int, width: int) -> np.ndarray:
    """nv12: uint8 [height*3/2, pitch>=width] (Y plane rows then interleaved UV rows).
    Returns uint8 [height, width, 3] RGB."""
    assert nv12.dtype == np.uint8 and nv12.shape[0] >= height * 3 // 2
    y = nv12[:height, :width].astype(np.int64)
    uv = nv12[height : height + height // 2, :width]
    u = uv[:, 0::2].astype(np.int64) - 128
    v = uv[:, 1::2].astype(np.int64) - 128
    u = np.repeat(np.repeat(u, 2, axis=0), 2, axis=1)[:height, :width]
    v = np.repeat(np.repeat(v, 2, axis=0), 2, axis=1)[:height, :width]
    yy = np.maximum(y - 16, 0) * CY
    half = 1 << (SHIFT - 1)
    r = (yy + half + CVR * v) >> SHIFT
    g = (yy + half + CVG * v + CUG * u) >> SHIFT
    b = (yy + half + CUB * u) >> SHIFT
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


# libswscale ff_yuv2rgb_c_init_tables, ITU601 (SWS_CS_DEFAULT), limited range, brightness 0 / contrast 1 / saturation 1:
# roundToInt16(c * 2^13): cy = 65536 * 255 / 219, crv = 104597, cbu = 132201, cgu = -25675, cgv = -53279
SWS_Y, SWS_VR, SWS_UB, SWS_VG, SWS_UG, SWS_YOFF, SWS_COFF = 9539, 13075, 16525, -6660, -3209, 128, 1024


def nv12_to_rgb_swscale(nv12: np.ndarray, height: int, width: int) -> np.ndarray:
    """Same layout contract as nv12_to_rgb; libswscale (x86) yuv420p -> rgb24 arithmetic."""
    assert nv12.dtype == np.uint8 and nv12.shape[0] >= height * 3 // 2
    y = nv12[:height, :width].astype(np.int64)
    uv = nv12[height : height + height // 2, :width]
    u = np.repeat(np.repeat(uv[:, 0::2].astype(np.int64), 2, axis=0), 2, axis=1)[:height, :width]
    v = np.repeat(np.repeat(uv[:, 1::2].astype(np.int64), 2, axis=0), 2, axis=1)[:height, :width]
    yy = (((y << 3) - SWS_YOFF) * SWS_Y) >> 16  # numpy >> on negative ints floors, like pmulhw
    uu, vv = (u << 3) - SWS_COFF, (v << 3) - SWS_COFF
    r = yy + ((vv * SWS_VR) >> 16)
    g = yy + ((uu * SWS_UG) >> 16) + ((vv * SWS_VG) >> 16)
    b = yy + ((uu * SWS_UB) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def rgb_to_nv12(rgb: np.ndarray, pitch: int | None = None) -> np.ndarray:
    """Synthetic-surface generator (BT.601 limited, float, 2x2 box chroma).  rgb: uint8 [H,W,3],
    H and W even.  Returns uint8 [H*3/2, pitch]."""
    h, w, _ = rgb.shape
    assert h % 2 == 0 and w % 2 == 0
    pitch = w if pitch is None else pitch
    f = rgb.astype(np.float64)
    r, g, b = f[..., 0], f[..., 1], f[..., 2]
    y = 16 + (65.481 * r + 128.553 * g + 24.966 * b) / 255.0
    cb = 128 + (-37.797 * r - 74.203 * g + 112.0 * b) / 255.0
    cr = 128 + (112.0 * r - 93.786 * g - 18.214 * b) / 255.0
    box = lambda a: a.reshape(h // 2, 2, w // 2, 2).mean(axis=(1, 3))  # noqa: E731
    out = np.zeros((h * 3 // 2, pitch), dtype=np.uint8)
    out[:h, :w] = np.clip(np.rint(y), 0, 255).astype(np.uint8)
    out[h:, 0:w:2] = np.clip(np.rint(box(cb)), 0, 255).astype(np.uint8)
    out[h:, 1:w:2] = np.clip(np.rint(box(cr)), 0, 255).astype(np.uint8)
    return out


def synthetic_nv12(height: int, width: int, seed: int, pitch: int | None = None) -> np.ndarray:
    """SURVEY.md 8d kernel microbench surfaces: moving gradient + per-pixel noise, legal video
    range (Y 16..235, UV 16..240).  Deterministic in `seed`."""
    rng = np.random.default_rng(1234 + seed)
    pitch = width if pitch is None else pitch
    out = np.zeros((height * 3 // 2, pitch), dtype=np.uint8)
    yy, xx = np.mgrid[0:height, 0:width]
    base = 16 + ((xx * 3 + yy * 2 + seed * 37) % 220)
    noise = rng.integers(-24, 25, size=(height, width))
    out[:height, :width] = np.clip(base + noise, 16, 235).astype(np.uint8)
    out[height:, :width] = rng.integers(16, 241, size=(height // 2, width), dtype=np.uint8)
    return out
