"""Oracle: frame preprocess (resize / crop / normalise), numpy float32.

Restates the transform chain of the reference's CLIP wrapper,
cosmos_curate/models/clip.py:48-62:

    Resize(224, BICUBIC, antialias=True) -> CenterCrop(224) -> ConvertImageDtype(float32)
    -> Normalize(mean, std)

applied to a uint8 NCHW tensor (clip.py:66-70).  The arithmetic lives in third-party code
that is not under /root/reference: torchvision 0.25 (pixi.toml:77) ``transforms.v1`` ->
``torchvision.transforms._functional_tensor.resize`` -> ATen ``_upsample_bicubic2d_aa``.
Published algorithm restated here:

  * output size: short side -> `size`, long side -> int(size * long / short)
    (torchvision.transforms.functional._compute_resized_output_size);
  * u8 -> float32; separable antialiased bicubic (Keys a = -0.5), per output index i:
        scale   = in / out                       (float32, align_corners=False)
        support = 2 * scale  if scale >= 1 else 2
        center  = scale * (i + 0.5)
        xmin    = max(int(center - support + 0.5), 0)
        xsize   = min(int(center + support + 0.5), in) - xmin
        w_j     = cubic((j + xmin - center + 0.5) * (1/scale if scale >= 1 else 1)),  normalised by sum
    horizontal pass over every source row first, then the vertical pass (ATen
    UpSampleKernel.cpp separable_upsample_generic_Nd_kernel_impl / UpSampleBilinear2d.cu
    upsample_gen2d_aa_out_frame), fp32 accumulation in tap order;
  * clamp to [0,255], round half-to-even, cast to u8 (torchvision _cast_squeeze_out);
  * centre crop with top/left = int(round((in - crop) / 2.0)) (python banker's rounding);
  * x / 255 (float32 division), then (x - mean) / std (float32 sub then division).

Also restates the bilinear path of the reference's NVDEC extractor (nvcodec_utils.py:189-194,
``cvcuda.resize_into(..., Interp.LINEAR)``): half-pixel centres, no antialias, 4 taps, float
arithmetic, round-to-nearest-even.  CV-CUDA rounding is UNPINNED (no reference test runs it);
tests compare against cv2.resize(INTER_LINEAR) with a 1 LSB budget.

Test infrastructure only (see oracle/__init__.py).
"""

from __future__ import annotations

import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
F32 = np.float32


def resized_output_size(h: int, w: int, size: int) -> tuple[int, int]:
    """torchvision _compute_resized_output_size for an int `size` (no max_size)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)  # (new_h, new_w)


def center_crop_offsets(h: int, w: int, crop: int) -> tuple[int, int]:
    """torchvision center_crop (no padding case): top, left."""
    return int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))


def _cubic(x: np.ndarray) -> np.ndarray:
    a = F32(-0.5)
    x = np.abs(x).astype(F32)
    near = ((a + F32(2)) * x - (a + F32(3))) * x * x + F32(1)
    far = (((x - F32(5)) * x + F32(8)) * x - F32(4)) * a
    return np.where(x < 1, near, np.where(x < 2, far, F32(0))).astype(F32)


def aa_bicubic_taps(in_size: int, out_size: int):
    """Per-output tap tables: xmin int32[out], xsize int32[out], w float32[out, maxtaps]."""
    scale = F32(in_size) / F32(out_size)
    support = F32(2.0) * scale if scale >= 1 else F32(2.0)
    invscale = F32(1.0) / scale if scale >= 1 else F32(1.0)
    i = np.arange(out_size, dtype=F32)
    center = scale * (i + F32(0.5))
    xmin = np.maximum((center - support + F32(0.5)).astype(np.int32), 0)
    xsize = np.minimum((center + support + F32(0.5)).astype(np.int32), in_size) - xmin
    maxt = int(xsize.max())
    j = np.arange(maxt, dtype=F32)[None, :]
    arg = (j + (xmin.astype(F32) - center)[:, None] + F32(0.5)) * invscale
    w = _cubic(arg)
    w[np.arange(maxt)[None, :] >= xsize[:, None]] = 0
    total = np.zeros(out_size, dtype=F32)
    for t in range(maxt):  # sequential fp32 sum, tap order
        total = (total + w[:, t]).astype(F32)
    w = np.where(total[:, None] != 0, w / total[:, None], w).astype(F32)
    return xmin.astype(np.int32), xsize.astype(np.int32), w


def _apply_taps_last_axis(x: np.ndarray, xmin, xsize, w) -> np.ndarray:
    """x float32 [..., in] -> [..., out]; t = x0*w0; t += xj*wj  (tap order, fp32)."""
    maxt = w.shape[1]
    idx = np.minimum(xmin[:, None] + np.arange(maxt)[None, :], x.shape[-1] - 1)  # [out, taps]
    acc = x[..., idx[:, 0]] * w[:, 0]
    for t in range(1, maxt):
        acc = (acc + x[..., idx[:, t]] * w[:, t]).astype(F32)
    return acc.astype(F32)


def resize_bicubic_aa_u8(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """img uint8 [..., H, W] (channel-planar) -> uint8 [..., new_h, new_w]."""
    h, w = img.shape[-2:]
    if (h, w) == (new_h, new_w):
        return img.copy()  # torchvision F.resize early return
    x = img.astype(F32)
    xm, xs, ww = aa_bicubic_taps(w, new_w)
    x = _apply_taps_last_axis(x, xm, xs, ww)  # horizontal
    ym, ys, wh = aa_bicubic_taps(h, new_h)
    x = np.swapaxes(_apply_taps_last_axis(np.swapaxes(x, -1, -2), ym, ys, wh), -1, -2)  # vertical
    return np.rint(np.clip(x, 0, 255)).astype(np.uint8)


def normalize_lut(mean=CLIP_MEAN, std=CLIP_STD) -> np.ndarray:
    """float32 [3,256]: ((v / 255) - mean_c) / std_c exactly as torchvision computes it."""
    v = np.arange(256, dtype=F32) / F32(255)
    m = np.asarray(mean, dtype=F32)[:, None]
    s = np.asarray(std, dtype=F32)[:, None]
    return ((v[None, :] - m) / s).astype(F32)


def clip_resize_crop_u8(frames_nhwc: np.ndarray, res: int = 224) -> np.ndarray:
    """uint8 [N,H,W,3] -> uint8 [N,3,res,res] (the u8 stage before /255 + Normalize)."""
    n, h, w, _ = frames_nhwc.shape
    x = np.ascontiguousarray(frames_nhwc.transpose(0, 3, 1, 2))
    nh, nw = resized_output_size(h, w, res)
    x = resize_bicubic_aa_u8(x, nh, nw)
    top, left = center_crop_offsets(nh, nw, res)
    return np.ascontiguousarray(x[:, :, top : top + res, left : left + res])


def clip_preprocess(frames_nhwc: np.ndarray, res: int = 224, mean=CLIP_MEAN, std=CLIP_STD) -> np.ndarray:
    """Full clip.py:48-62 chain: uint8 [N,H,W,3] -> float32 [N,3,res,res]."""
    u8 = clip_resize_crop_u8(frames_nhwc, res)
    lut = normalize_lut(mean, std)
    out = np.empty(u8.shape, dtype=F32)
    for c in range(3):
        out[:, c] = lut[c][u8[:, c]]
    return out


def to_patches(x_nchw: np.ndarray, patch: int, k_pad: int | None = None) -> np.ndarray:
    """[N,3,R,R] -> [N, (R/p)^2, 3*p*p (zero-padded to k_pad)], row order (c, py, px) - the
    im2col of Conv2d(3, D, kernel=p, stride=p) (HF CLIPVisionEmbeddings.patch_embedding)."""
    n, c, r, _ = x_nchw.shape
    g = r // patch  # a stride-p conv drops the remainder rows / columns (SigLIP 384 / 14 = 27)
    x_nchw = x_nchw[:, :, : g * patch, : g * patch]
    p = x_nchw.reshape(n, c, g, patch, g, patch).transpose(0, 2, 4, 1, 3, 5).reshape(n, g * g, c * patch * patch)
    if k_pad is not None and k_pad > p.shape[-1]:
        p = np.concatenate([p, np.zeros((n, g * g, k_pad - p.shape[-1]), dtype=p.dtype)], axis=-1)
    return np.ascontiguousarray(p)


def resize_bilinear_u8(img_hwc: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """nvcodec_utils.py:189-194 stand-in: half-pixel bilinear, float32, round-half-even.
    img uint8 [H,W,C] -> uint8 [new_h,new_w,C]."""
    h, w = img_hwc.shape[:2]
    sy, sx = F32(h) / F32(new_h), F32(w) / F32(new_w)
    fy = (np.arange(new_h, dtype=F32) + F32(0.5)) * sy - F32(0.5)
    fx = (np.arange(new_w, dtype=F32) + F32(0.5)) * sx - F32(0.5)
    y0 = np.floor(fy).astype(np.int32)
    x0 = np.floor(fx).astype(np.int32)
    wy = (fy - y0.astype(F32)).astype(F32)
    wx = (fx - x0.astype(F32)).astype(F32)
    y0c, y1c = np.clip(y0, 0, h - 1), np.clip(y0 + 1, 0, h - 1)
    x0c, x1c = np.clip(x0, 0, w - 1), np.clip(x0 + 1, 0, w - 1)
    f = img_hwc.astype(F32)
    wx_ = wx[None, :, None]
    wy_ = wy[:, None, None]
    top = f[y0c][:, x0c] * (F32(1) - wx_) + f[y0c][:, x1c] * wx_
    bot = f[y1c][:, x0c] * (F32(1) - wx_) + f[y1c][:, x1c] * wx_
    out = top * (F32(1) - wy_) + bot * wy_
    return np.rint(np.clip(out, 0, 255)).astype(np.uint8)


def pynvc_target_size(width: int, height: int, target_w: int = -1, target_h: int = -1) -> tuple[int, int]:
    """nvcodec_utils.py:128-135: if either target is -1 both are recomputed from the decoded
    size: downscale = width // 256 (1 below 256), python round()."""
    if target_w != -1 and target_h != -1:
        return target_w, target_h
    ds = 1 if width < 256 else width // 256
    return round(width / ds), round(height / ds)
