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


def test_siglip_tower_vs_hf_golden(ctx):
    """SigLIP-style tower (no CLS, patch bias, gelu_tanh, MAP pooling head) against transformers' SiglipVisionModel."""
    g = load_golden("siglip_tiny_hf.npz")
    cfg = vit.VitConfig(**golden_json(g, "cfg"))
    w = {k[2:]: g[k] for k in g.files if k.startswith("w_")}
    tower = _tower(ctx, cfg, w, 4)
    x16 = g["x"].astype(np.float16)
    patches = torch.from_numpy(preprocess.to_patches(x16, cfg.patch, tower.k_pad)).cuda()
    _, feat, _ = tower.forward_patches(patches, want_features=True)
    want = vit.forward(cfg, w, x16.astype(np.float32))["features"]  # oracle on the same fp16-rounded pixels
    np.testing.assert_allclose(want, g["pooled"], rtol=2e-2, atol=2e-3)  # ... which itself tracks HF on the fp32 pixels
    assert _rel(feat.cpu().numpy(), want) < 2e-3


def test_siglip_so400m_shape_two_layers(ctx):
    """SoViT-400m/14 @384 geometry (729 tokens, 16 heads x 72, MLP 4304, 384 = 27*14 + 6) with 2 layers: streamed attention,
    N=1152 / K=4304 GEMM tails, patch rows from a non-divisible image size; embeddings vs the oracle."""
    cfg = vit.SIGLIP_SO400M_2L
    w = vit.random_weights(cfg, seed=4)
    frames = [color.synthetic_nv12(1080, 1920, seed=70 + s) for s in range(2)]
    pool = _nv12_pool(ctx, frames, 1920, 1080, 2048, 1088)
    tower = _tower(ctx, cfg, w, max_batch=2)
    mean = std = (0.5, 0.5, 0.5)  # SigLIP normalisation
    emb, feat, _ = tower.embed_pool(pool, mean=mean, std=std, want_features=True)
    u8 = ctx.preprocess_clip_u8(pool, res=384).cpu().numpy()
    lut = preprocess.normalize_lut(mean, std)
    x = np.stack([lut[c][u8[:, c]] for c in range(3)], axis=1)
    ref = vit.forward(cfg, w, x)
    assert _rel(emb.cpu().numpy(), ref["embedding"]) < 2e-3


def test_siglip_so400m_full_depth_vs_oracle(ctx):
    """The full SoViT-400m/14 @384 tower (27 layers, 1152 hidden, 16 x 72 heads, MLP 4304, 729 tokens, MAP head) - BASELINE.json
    configs[3]'s network - against the fp32 oracle at n = 2, from NV12 surfaces (384 = 27 * 14 + 6 geometry included)."""
    from cosmos_curate_b200.models import weights as W

    cfg = vit.SIGLIP_SO400M_384
    assert (cfg.layers, cfg.hidden, cfg.heads, cfg.mlp, cfg.tokens) == (27, 1152, 16, 4304, 729)
    w = vit.random_weights(cfg, seed=8)
    frames = [color.synthetic_nv12(1080, 1920, seed=90 + s) for s in range(2)]
    pool = _nv12_pool(ctx, frames, 1920, 1080, 2048, 1088)
    tower = _tower(ctx, cfg, w, max_batch=2)
    mean = std = (0.5, 0.5, 0.5)
    emb, feat, _ = tower.embed_pool(pool, mean=mean, std=std, want_features=True)
    u8 = ctx.preprocess_clip_u8(pool, res=384).cpu().numpy()
    lut = preprocess.normalize_lut(mean, std)
    x = np.stack([lut[c][u8[:, c]] for c in range(3)], axis=1)
    ref = vit.forward(cfg, w, x)
    rel = _rel(emb.cpu().numpy(), ref["embedding"])
    print(f"\n[SoViT-400m/14@384, 27 layers] embedding rel err {rel:.2e}")
    assert rel < 1e-3
    assert abs(W.SIGLIP_SO400M_384.flops_per_image() / 1e9 - 666.5) < 15  # ~666.5 GFLOP per image
    # the product's own seeded weights carry the MAP head too (bench.py's secondary line uses them)
    w2 = W.seeded_weights(W.SIGLIP_SO400M_384.__class__(**{**W.SIGLIP_SO400M_384.to_dict(), "layers": 1}), seed=1)
    assert {"map_probe", "map_in_w", "map_fc2_b", "patch_b"} <= set(w2)
