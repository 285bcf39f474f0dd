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
from oracle import color, preprocess, vit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from cosmos_curate_b200.runtime import Context

    c = Context(0)
    yield c
    c.close()


def _nv12_pool(ctx, frames_nv12: list[np.ndarray], width: int, height: int, pitch: int, luma_rows: int):
    rows = luma_rows + height // 2
    buf = np.zeros((len(frames_nv12), rows, pitch), dtype=np.uint8)
    for i, f in enumerate(frames_nv12):
        buf[i, :height, :width] = f[:height, :width]
        buf[i, luma_rows : luma_rows + height // 2, :width] = f[height:, :width]
    t = torch.from_numpy(buf).cuda()
    return ctx.nv12_pool(t, width, height, luma_rows)


def _u8_budget(got: np.ndarray, want: np.ndarray, frac: float = 1e-4):
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 1, f"max diff {d.max()}"
    assert (d > 0).mean() <= frac, f"{(d > 0).mean():.2e} of pixels differ"


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


# ------------------------------------------------------------------------------------ GEMM / LN / attention
@pytest.mark.parametrize(("m", "n", "k"), [(128, 128, 64), (300, 256, 192), (1000, 1024, 1024), (2570, 3072, 1024), (20000, 1024, 4096), (257, 136, 72), (40000, 4096, 1024)])
def test_gemm_plain(ctx, m, n, k):
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = (torch.randn(m, k, device="cuda", generator=g) * 0.5).half()
    w = (torch.randn(n, k, device="cuda", generator=g) * 0.5).half()
    bias = torch.randn(n, device="cuda", generator=g)
    got = ctx.gemm(a, w, bias=bias).float()
    want = a.float() @ w.float().T + bias
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= 2e-3 * scale + 1e-2, (err, scale)  # fp16 output rounding of values ~ sqrt(k)/4


def test_gemm_epilogues(ctx):
    g = torch.Generator(device="cuda").manual_seed(3)
    m, n, k = 3000, 512, 256
    a = (torch.randn(m, k, device="cuda", generator=g) * 0.3).half()
    w = (torch.randn(n, k, device="cuda", generator=g) * 0.2).half()
    bias = torch.randn(n, device="cuda", generator=g)
    z = a.float() @ w.float().T + bias
    from cosmos_curate_b200 import _lib

    got = ctx.gemm(a, w, bias=bias, epilogue=_lib.EPI_QUICK_GELU).float()
    torch.testing.assert_close(got, z * torch.sigmoid(1.702 * z), rtol=2e-3, atol=2e-3)
    got = ctx.gemm(a, w, bias=bias, epilogue=_lib.EPI_GELU_TANH).float()
    torch.testing.assert_close(got, torch.nn.functional.gelu(z, approximate="tanh"), rtol=2e-3, atol=2e-3)
    res = torch.randn(m, n, device="cuda", generator=g)
    want = res + z
    got = ctx.gemm(a, w, bias=bias, residual=res.clone(), out_f32=True)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-3)
    got = ctx.gemm(a, w, out_f32=True)  # no bias, no residual
    torch.testing.assert_close(got, a.float() @ w.float().T, rtol=1e-4, atol=1e-3)


def test_layernorm(ctx):
    g = torch.Generator(device="cuda").manual_seed(4)
    for rows, d in ((1000, 1024), (77, 768), (513, 1152), (9, 256)):
        x = torch.randn(rows, d, device="cuda", generator=g) * 3 + 1
        gamma = torch.randn(d, device="cuda", generator=g)
        beta = torch.randn(d, device="cuda", generator=g)
        got = ctx.layernorm(x, gamma, beta, 1e-5).float()
        want = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5)
        torch.testing.assert_close(got, want, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize(("n", "t", "heads", "hd"), [(3, 257, 16, 64), (2, 50, 12, 64), (2, 64, 4, 64), (1, 256, 16, 72), (2, 17, 2, 32)])
def test_attention(ctx, n, t, heads, hd):
    g = torch.Generator(device="cuda").manual_seed(t)
    d = heads * hd
    qkv = (torch.randn(n, t, 3 * d, device="cuda", generator=g) * 1.5).half()
    got = ctx.attention(qkv, heads).float()
    q, k, v = qkv.float().view(n, t, 3, heads, hd).permute(2, 0, 3, 1, 4)
    p = torch.softmax(q @ k.transpose(-1, -2) * hd**-0.5, dim=-1)
    want = (p @ v).permute(0, 2, 1, 3).reshape(n, t, d)
    torch.testing.assert_close(got, want, rtol=1e-2, atol=4e-3)  # P and O rounded to fp16


# ------------------------------------------------------------------------------------ tower
def _rel(got: np.ndarray, want: np.ndarray) -> float:
    return float((np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)).max())


def _tower(ctx, cfg, w, max_batch, aesthetic=None):
    from cosmos_curate_b200.runtime import VitTower

    return VitTower(ctx, cfg.to_dict(), w, max_batch=max_batch, aesthetic=aesthetic)


def test_tower_tiny_vs_reference_wrapper_golden(ctx):
    """RGB frames -> embeddings against the REFERENCE's own _CLIPImageEmbeddings.__call__ outputs."""
    g = load_golden("clip_tiny_ref.npz")
    cfg = vit.VitConfig(**golden_json(g, "cfg"))
    w = {k[2:]: g[k] for k in g.files if k.startswith("w_")}
    if cfg.hidden % 128:
        pytest.skip("golden tiny config hidden not a multiple of 128")
    tower = _tower(ctx, cfg, w, 4)
    for name in sorted(k[:-3] for k in g.files if k.endswith("_in")):
        pool = ctx.rgb_pool(torch.from_numpy(g[name + "_in"]).cuda())
        emb, _, _ = tower.embed_pool(pool)
        assert _rel(emb.cpu().numpy(), g[name + "_emb"]) < 1e-3, name


@pytest.mark.parametrize("cfg_name", ["CLIP_TINY", "CLIP_VIT_B32", "CLIP_VIT_L14"])
def test_tower_vs_oracle(ctx, cfg_name):
    cfg = getattr(vit, cfg_name)
    w = vit.random_weights(cfg, seed=1)
    sd = vit.random_aesthetic_mlp(seed=2, in_dim=cfg.proj_dim)
    aw, ab = vit.collapse_aesthetic_mlp(sd)
    n = 5 if cfg_name != "CLIP_VIT_L14" else 3
    frames = [color.synthetic_nv12(1080, 1920, seed=40 + s) for s in range(n)]
    pool = _nv12_pool(ctx, frames, 1920, 1080, 2048, 1088)
    tower = _tower(ctx, cfg, w, max_batch=4, aesthetic=(aw, ab))  # max_batch < n: exercises chunking
    emb, feat, score = tower.embed_pool(pool, want_features=True)
    # oracle on the SAME u8 stage (isolates the tower from the <=1 LSB resize budget)
    u8 = ctx.preprocess_clip_u8(pool).cpu().numpy()
    lut = preprocess.normalize_lut()
    x = np.stack([lut[c][u8[:, c]] for c in range(3)], axis=1)
    ref = vit.forward(cfg, w, x)
    assert _rel(emb.cpu().numpy(), ref["embedding"]) < 1e-3  # BASELINE.json: fp embeddings within 1e-3 relative
    assert _rel(feat.cpu().numpy(), ref["features"]) < 1e-3
    want_score = vit.aesthetic_mlp_forward(sd, ref["embedding"])
    np.testing.assert_allclose(score.cpu().numpy(), want_score, rtol=0, atol=2e-3)  # reference test tolerance 0.002
