"""GPU parity tests (B200): every call goes through the C ABI of libcurate_b200.so.

Integer / byte stages are compared bit-exactly with the oracle where the arithmetic is pinned
(colour conversion, frame indices), within the stated fp32-summation budget where it is not
(u8 stage of the antialiased resize: <= 1 LSB on <= 1e-4 of the pixels - the same budget the
oracle itself needs against ATen, tests/test_oracle_cpu.py).  Floating-point stages: tolerance in
each test.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import golden_json, load_golden
from gpu_helpers import ctx, nv12_pool as _nv12_pool, u8_budget as _u8_budget  # noqa: F401
from oracle import color, preprocess, vit

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------ colour / bilinear
@pytest.mark.parametrize(("h", "w", "pitch", "luma_rows"), [(64, 96, 128, 64), (480, 854, 1024, 480), (1080, 1920, 2048, 1088)])
def test_nv12_to_rgb_bit_exact(ctx, h, w, pitch, luma_rows):
    if w % 2:
        pytest.skip("odd width")
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, size=(h * 3 // 2, w), dtype=np.uint8) for _ in range(2)]
    pool = _nv12_pool(ctx, frames, w, h, pitch, luma_rows)
    got = ctx.nv12_to_rgb(pool).cpu().numpy()
    for i, f in enumerate(frames):
        np.testing.assert_array_equal(got[i], color.nv12_to_rgb(f, h, w))


def test_bilinear_27x48_matches_oracle(ctx):
    frames = [color.synthetic_nv12(1080, 1920, seed=s) for s in range(3)]
    pool = _nv12_pool(ctx, frames, 1920, 1080, 2048, 1088)
    got = ctx.preprocess_bilinear_u8(pool, 48, 27).cpu().numpy()
    assert got.shape == (3, 27, 48, 3)
    for i, f in enumerate(frames):
        want = preprocess.resize_bilinear_u8(color.nv12_to_rgb(f, 1080, 1920), 27, 48)
        d = np.abs(got[i].astype(int) - want.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3  # fp32 contraction differences only


# ------------------------------------------------------------------------------------ CLIP preprocess
@pytest.mark.parametrize(("h", "w", "pitch", "luma_rows"), [(1080, 1920, 2048, 1088), (720, 1280, 1280, 720), (480, 854, 1024, 480), (1920, 1080, 1280, 1920)])
def test_clip_preprocess_nv12_u8_stage(ctx, h, w, pitch, luma_rows):
    frames = [color.synthetic_nv12(h, w, seed=s) for s in range(2)]
    pool = _nv12_pool(ctx, frames, w, h, pitch, luma_rows)
    got = ctx.preprocess_clip_u8(pool, res=224).cpu().numpy()
    rgb = np.stack([color.nv12_to_rgb(f, h, w) for f in frames])
    _u8_budget(got, preprocess.clip_resize_crop_u8(rgb, 224))


def test_clip_preprocess_typed_and_patch_layout(ctx):
    h, w = 1080, 1920
    frames = [color.synthetic_nv12(h, w, seed=10 + s) for s in range(3)]
    pool = _nv12_pool(ctx, frames, w, h, 2048, 1088)
    u8 = ctx.preprocess_clip_u8(pool).cpu().numpy()
    lut = preprocess.normalize_lut()
    want32 = np.stack([lut[c][u8[:, c]] for c in range(3)], axis=1)
    # given the u8 stage, normalise + cast must be exact
    got32 = ctx.preprocess_clip(pool, dtype=torch.float32).cpu().numpy()
    np.testing.assert_array_equal(got32, want32)
    got16 = ctx.preprocess_clip(pool, dtype=torch.float16).cpu().numpy()
    np.testing.assert_array_equal(got16, want32.astype(np.float16))
    gotbf = ctx.preprocess_clip(pool, dtype=torch.bfloat16).float().cpu().numpy()
    np.testing.assert_array_equal(gotbf, torch.from_numpy(want32).to(torch.bfloat16).float().numpy())
    for patch, k_pad in ((14, 640), (32, 3072), (16, 768)):
        gp = ctx.preprocess_clip(pool, dtype=torch.float16, layout="patch", patch=patch, k_pad=k_pad).cpu().numpy()
        np.testing.assert_array_equal(gp, preprocess.to_patches(want32.astype(np.float16), patch, k_pad))


def test_clip_preprocess_rgb_frames_vs_reference_golden(ctx):
    """Host RGB frames (the ModelInterface input of clip.py:64-70) against the REFERENCE's own transform output."""
    g = load_golden("clip_preprocess_ref.npz")
    lut = preprocess.normalize_lut()
    for name in sorted(k[:-3] for k in g.files if k.endswith("_in")):
        fr, want = g[name + "_in"], g[name + "_out"]
        pool = ctx.rgb_pool(torch.from_numpy(fr).cuda())
        got = ctx.preprocess_clip(pool, dtype=torch.float32).cpu().numpy()
        assert got.shape == want.shape
        bad = 0
        for c in range(3):
            ref_u8 = np.abs(want[:, c, :, :, None] - lut[c]).argmin(-1)
            got_u8 = np.abs(got[:, c, :, :, None] - lut[c]).argmin(-1)
            assert np.abs(lut[c][got_u8] - got[:, c]).max() == 0.0
            d = np.abs(ref_u8 - got_u8)
            assert d.max() <= 1, name
            bad += int((d > 0).sum())
        assert bad <= 1e-4 * want.size, (name, bad)


def test_clip_preprocess_vs_torchvision_cuda(ctx):
    """The reference's actual CUDA path: torchvision transforms on a CUDA uint8 tensor (clip.py:48-70)."""
    tv = pytest.importorskip("torchvision.transforms")
    t = tv.Compose([tv.Resize(224, interpolation=tv.InterpolationMode.BICUBIC, antialias=True), tv.CenterCrop(224)])
    rng = np.random.default_rng(5)
    fr = rng.integers(0, 256, size=(2, 1080, 1920, 3), dtype=np.uint8)
    x = torch.from_numpy(fr).cuda()
    want = t(x.permute(0, 3, 1, 2)).cpu().numpy()
    got = ctx.preprocess_clip_u8(ctx.rgb_pool(x)).cpu().numpy()
    _u8_budget(got, want)


def test_preprocess_edge_cases(ctx):
    from cosmos_curate_b200._lib import CurateB200Error

    frames = [color.synthetic_nv12(64, 96, seed=1)]
    pool = _nv12_pool(ctx, frames, 96, 64, 128, 64)
    assert ctx.preprocess_clip_u8(pool, slots=np.zeros(0, np.int32)).shape[0] == 0  # empty batch
    out = ctx.preprocess_clip_u8(pool, slots=[0, 0, 0]).cpu().numpy()  # repeated slot (supersampled frame ids)
    np.testing.assert_array_equal(out[0], out[2])
    with pytest.raises(CurateB200Error):
        ctx.preprocess_clip_u8(pool, slots=[-1])
    # upscale (scale < 1) path
    rgb = color.nv12_to_rgb(frames[0], 64, 96)[None]
    _u8_budget(out[:1], preprocess.clip_resize_crop_u8(rgb, 224))


