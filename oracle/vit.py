"""Oracle: CLIP / SigLIP image tower + aesthetic head, torch fp32 on CPU.

This is the floating-point part of the path, so the oracle is a plain fp32 torch restatement
(tolerance for the CUDA path: 1e-3 relative on the unit-norm embedding, BASELINE.json).

What it restates:
  * cosmos_curate/models/clip.py:71-74 - ``CLIPModel.get_image_features(pixel_values)`` then
    ``embed / ||embed||_2``.  The tower itself is third-party (transformers, pinned ``<5`` in
    pixi.toml:156; 5.5.0 installed here - same math, different return type):
    patch-embed Conv2d(3,D,p,p,bias=False) -> [CLS] concat -> + learned position embedding ->
    pre_layrnorm -> L x [LN1, MHA(q,k,v,out with bias; softmax in fp32; scale d^-1/2), +res,
    LN2, fc1, quick_gelu (x*sigmoid(1.702x)), fc2, +res] -> post_layernorm(CLS) ->
    visual_projection (no bias).
  * cosmos_curate/models/aesthetics.py:44-53,94-106 - the 5-Linear MLP has no activation
    (Dropout is identity in eval), so it is one affine map 768 -> 1; ``collapse_aesthetic_mlp``
    folds it in float64 and ``aesthetic_mlp_forward`` keeps the layer-by-layer form for pinning.
  * SigLIP vision tower (BASELINE.json config 4; not in the reference tree - transformers'
    SiglipVisionModel is the stated stand-in): conv patch-embed with bias, no
    CLS, gelu_pytorch_tanh, post_layernorm on all tokens, MAP attention-pooling head.

Pinned by tests/golden/vit_*.npz which oracle/make_golden.py generates with transformers'
own CLIPModel / SiglipVisionModel (seeded random weights).

Test infrastructure only (see oracle/__init__.py).
"""

from __future__ import annotations

from dataclasses import asdict, dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class VitConfig:
    image_size: int = 224
    patch: int = 14
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    proj_dim: int = 768  # 0 = no projection (SigLIP pooled output)
    act: str = "quick_gelu"  # or "gelu_tanh"
    ln_eps: float = 1e-5
    arch: str = "clip"  # or "siglip"

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + (1 if self.arch == "clip" else 0)

    def to_dict(self):
        return asdict(self)


CLIP_VIT_L14 = VitConfig()
CLIP_VIT_B32 = VitConfig(patch=32, hidden=768, layers=12, heads=12, mlp=3072, proj_dim=512)
SIGLIP_SO400M_384 = VitConfig(
    image_size=384, patch=14, hidden=1152, layers=27, heads=16, mlp=4304, proj_dim=0, act="gelu_tanh", ln_eps=1e-6, arch="siglip"
)
SIGLIP_SO400M_2L = VitConfig(image_size=384, patch=14, hidden=1152, layers=2, heads=16, mlp=4304, proj_dim=0, act="gelu_tanh", ln_eps=1e-6, arch="siglip")
# tiny configs for fast parity tests (same code paths, small sizes)
CLIP_TINY = VitConfig(image_size=224, patch=32, hidden=256, layers=2, heads=4, mlp=512, proj_dim=128)


def random_weights(cfg: VitConfig, seed: int = 0, scale: float = 1.0) -> dict[str, np.ndarray]:
    """Seeded weights with realistic magnitudes (independent of transformers' init)."""
    g = torch.Generator().manual_seed(seed)
    d, m = cfg.hidden, cfg.mlp
    kp = 3 * cfg.patch * cfg.patch

    def rn(*shape, std):
        return (torch.randn(*shape, generator=g) * std * scale).numpy().astype(np.float32)

    w = {"patch_w": rn(d, kp, std=kp**-0.5), "pos": rn(cfg.tokens, d, std=0.02)}
    if cfg.arch == "clip":
        w["cls"] = rn(d, std=0.02)
        w["pre_ln_w"] = 1 + rn(d, std=0.05)
        w["pre_ln_b"] = rn(d, std=0.02)
    else:
        w["patch_b"] = rn(d, std=0.02)
    for i in range(cfg.layers):
        p = f"L{i}."
        w[p + "ln1_w"] = 1 + rn(d, std=0.05)
        w[p + "ln1_b"] = rn(d, std=0.02)
        w[p + "qkv_w"] = rn(3 * d, d, std=d**-0.5)
        w[p + "qkv_b"] = rn(3 * d, std=0.02)
        w[p + "out_w"] = rn(d, d, std=d**-0.5 / (2 * cfg.layers) ** 0.5)
        w[p + "out_b"] = rn(d, std=0.02)
        w[p + "ln2_w"] = 1 + rn(d, std=0.05)
        w[p + "ln2_b"] = rn(d, std=0.02)
        w[p + "fc1_w"] = rn(m, d, std=d**-0.5)
        w[p + "fc1_b"] = rn(m, std=0.02)
        w[p + "fc2_w"] = rn(d, m, std=m**-0.5 / (2 * cfg.layers) ** 0.5)
        w[p + "fc2_b"] = rn(d, std=0.02)
    w["post_ln_w"] = 1 + rn(d, std=0.05)
    w["post_ln_b"] = rn(d, std=0.02)
    if cfg.proj_dim:
        w["proj_w"] = rn(cfg.proj_dim, d, std=d**-0.5)
    if cfg.arch == "siglip":
        w["map_probe"] = rn(d, std=0.02)
        w["map_in_w"] = rn(3 * d, d, std=d**-0.5)
        w["map_in_b"] = rn(3 * d, std=0.02)
        w["map_out_w"] = rn(d, d, std=d**-0.5)
        w["map_out_b"] = rn(d, std=0.02)
        w["map_ln_w"] = 1 + rn(d, std=0.05)
        w["map_ln_b"] = rn(d, std=0.02)
        w["map_fc1_w"] = rn(m, d, std=d**-0.5)
        w["map_fc1_b"] = rn(m, std=0.02)
        w["map_fc2_w"] = rn(d, m, std=m**-0.5)
        w["map_fc2_b"] = rn(d, std=0.02)
    return w


def weights_from_hf_clip(model) -> tuple[VitConfig, dict[str, np.ndarray]]:
    """Flatten a transformers CLIPModel (vision side) into this module's weight names."""
    vc = model.config.vision_config
    cfg = VitConfig(
        image_size=vc.image_size, patch=vc.patch_size, hidden=vc.hidden_size, layers=vc.num_hidden_layers,
        heads=vc.num_attention_heads, mlp=vc.intermediate_size, proj_dim=model.config.projection_dim,
        act=vc.hidden_act, ln_eps=vc.layer_norm_eps, arch="clip",
    )  # fmt: skip
    sd = {k: v.detach().float().numpy() for k, v in model.state_dict().items()}
    v = "vision_model."
    w = {
        "patch_w": sd[v + "embeddings.patch_embedding.weight"].reshape(cfg.hidden, -1),
        "cls": sd[v + "embeddings.class_embedding"],
        "pos": sd[v + "embeddings.position_embedding.weight"],
        "pre_ln_w": sd[v + "pre_layrnorm.weight"], "pre_ln_b": sd[v + "pre_layrnorm.bias"],
        "post_ln_w": sd[v + "post_layernorm.weight"], "post_ln_b": sd[v + "post_layernorm.bias"],
        "proj_w": sd["visual_projection.weight"],
    }  # fmt: skip
    for i in range(cfg.layers):
        s, p = f"{v}encoder.layers.{i}.", f"L{i}."
        w[p + "ln1_w"], w[p + "ln1_b"] = sd[s + "layer_norm1.weight"], sd[s + "layer_norm1.bias"]
        w[p + "ln2_w"], w[p + "ln2_b"] = sd[s + "layer_norm2.weight"], sd[s + "layer_norm2.bias"]
        w[p + "qkv_w"] = np.concatenate([sd[s + f"self_attn.{n}_proj.weight"] for n in "qkv"], axis=0)
        w[p + "qkv_b"] = np.concatenate([sd[s + f"self_attn.{n}_proj.bias"] for n in "qkv"], axis=0)
        w[p + "out_w"], w[p + "out_b"] = sd[s + "self_attn.out_proj.weight"], sd[s + "self_attn.out_proj.bias"]
        w[p + "fc1_w"], w[p + "fc1_b"] = sd[s + "mlp.fc1.weight"], sd[s + "mlp.fc1.bias"]
        w[p + "fc2_w"], w[p + "fc2_b"] = sd[s + "mlp.fc2.weight"], sd[s + "mlp.fc2.bias"]
    return cfg, w


def weights_from_hf_siglip(model) -> tuple[VitConfig, dict[str, np.ndarray]]:
    """Flatten a transformers SiglipVisionModel."""
    vc = model.config
    cfg = VitConfig(
        image_size=vc.image_size, patch=vc.patch_size, hidden=vc.hidden_size, layers=vc.num_hidden_layers,
        heads=vc.num_attention_heads, mlp=vc.intermediate_size, proj_dim=0,
        act="gelu_tanh", ln_eps=vc.layer_norm_eps, arch="siglip",
    )  # fmt: skip
    sd = {k: v.detach().float().numpy() for k, v in model.state_dict().items()}
    v = "vision_model."
    w = {
        "patch_w": sd[v + "embeddings.patch_embedding.weight"].reshape(cfg.hidden, -1),
        "patch_b": sd[v + "embeddings.patch_embedding.bias"],
        "pos": sd[v + "embeddings.position_embedding.weight"],
        "post_ln_w": sd[v + "post_layernorm.weight"], "post_ln_b": sd[v + "post_layernorm.bias"],
        "map_probe": sd[v + "head.probe"].reshape(-1),
        "map_in_w": sd[v + "head.attention.in_proj_weight"], "map_in_b": sd[v + "head.attention.in_proj_bias"],
        "map_out_w": sd[v + "head.attention.out_proj.weight"], "map_out_b": sd[v + "head.attention.out_proj.bias"],
        "map_ln_w": sd[v + "head.layernorm.weight"], "map_ln_b": sd[v + "head.layernorm.bias"],
        "map_fc1_w": sd[v + "head.mlp.fc1.weight"], "map_fc1_b": sd[v + "head.mlp.fc1.bias"],
        "map_fc2_w": sd[v + "head.mlp.fc2.weight"], "map_fc2_b": sd[v + "head.mlp.fc2.bias"],
    }  # fmt: skip
    for i in range(cfg.layers):
        s, p = f"{v}encoder.layers.{i}.", f"L{i}."
        w[p + "ln1_w"], w[p + "ln1_b"] = sd[s + "layer_norm1.weight"], sd[s + "layer_norm1.bias"]
        w[p + "ln2_w"], w[p + "ln2_b"] = sd[s + "layer_norm2.weight"], sd[s + "layer_norm2.bias"]
        w[p + "qkv_w"] = np.concatenate([sd[s + f"self_attn.{n}_proj.weight"] for n in "qkv"], axis=0)
        w[p + "qkv_b"] = np.concatenate([sd[s + f"self_attn.{n}_proj.bias"] for n in "qkv"], axis=0)
        w[p + "out_w"], w[p + "out_b"] = sd[s + "self_attn.out_proj.weight"], sd[s + "self_attn.out_proj.bias"]
        w[p + "fc1_w"], w[p + "fc1_b"] = sd[s + "mlp.fc1.weight"], sd[s + "mlp.fc1.bias"]
        w[p + "fc2_w"], w[p + "fc2_b"] = sd[s + "mlp.fc2.weight"], sd[s + "mlp.fc2.bias"]
    return cfg, w


def _act(x: torch.Tensor, kind: str) -> torch.Tensor:
    if kind == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    if kind in ("gelu_tanh", "gelu_pytorch_tanh"):
        return F.gelu(x, approximate="tanh")
    raise ValueError(kind)


def _mha(x, w_in, b_in, heads):
    n, t, d = x.shape
    qkv = F.linear(x, w_in, b_in).view(n, t, 3, heads, d // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = torch.matmul(q, k.transpose(-1, -2)) * (d // heads) ** -0.5
    p = torch.softmax(s.float(), dim=-1)
    return torch.matmul(p, v).permute(0, 2, 1, 3).reshape(n, t, d)


@torch.no_grad()
def forward(cfg: VitConfig, w: dict, pixels: np.ndarray | torch.Tensor, return_hidden: bool = False):
    """pixels float32 [N,3,R,R] (already normalised).  Returns dict(features, embedding[, hidden])."""
    W = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in w.items()}
    x = torch.as_tensor(pixels, dtype=torch.float32)
    n, d, p, g = x.shape[0], cfg.hidden, cfg.patch, cfg.grid
    x = x[:, :, : g * p, : g * p]  # Conv2d(stride = kernel = p) ignores the remainder (SigLIP: 384 = 27 * 14 + 6)
    patches = x.reshape(n, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(n, g * g, 3 * p * p)
    tok = patches @ W["patch_w"].T
    if cfg.arch == "clip":
        tok = torch.cat([W["cls"].expand(n, 1, d), tok], dim=1) + W["pos"]
        h = F.layer_norm(tok, (d,), W["pre_ln_w"], W["pre_ln_b"], cfg.ln_eps)
    else:
        h = tok + W["patch_b"] + W["pos"]
    hidden = [h]
    for i in range(cfg.layers):
        q = f"L{i}."
        a = F.layer_norm(h, (d,), W[q + "ln1_w"], W[q + "ln1_b"], cfg.ln_eps)
        a = _mha(a, W[q + "qkv_w"], W[q + "qkv_b"], cfg.heads)
        h = h + F.linear(a, W[q + "out_w"], W[q + "out_b"])
        m = F.layer_norm(h, (d,), W[q + "ln2_w"], W[q + "ln2_b"], cfg.ln_eps)
        m = _act(F.linear(m, W[q + "fc1_w"], W[q + "fc1_b"]), cfg.act)
        h = h + F.linear(m, W[q + "fc2_w"], W[q + "fc2_b"])
        hidden.append(h)
    if cfg.arch == "clip":
        pooled = F.layer_norm(h[:, 0], (d,), W["post_ln_w"], W["post_ln_b"], cfg.ln_eps)
        feat = pooled @ W["proj_w"].T if cfg.proj_dim else pooled
    else:
        hs = F.layer_norm(h, (d,), W["post_ln_w"], W["post_ln_b"], cfg.ln_eps)
        # MAP head: single learned query attends over all tokens (nn.MultiheadAttention)
        hd = d // cfg.heads
        wq, wk, wv = W["map_in_w"][:d], W["map_in_w"][d : 2 * d], W["map_in_w"][2 * d :]
        bq, bk, bv = W["map_in_b"][:d], W["map_in_b"][d : 2 * d], W["map_in_b"][2 * d :]
        qh = (F.linear(W["map_probe"].expand(n, 1, d), wq, bq)).view(n, 1, cfg.heads, hd).transpose(1, 2)
        kh = F.linear(hs, wk, bk).view(n, -1, cfg.heads, hd).transpose(1, 2)
        vh = F.linear(hs, wv, bv).view(n, -1, cfg.heads, hd).transpose(1, 2)
        pr = torch.softmax((qh @ kh.transpose(-1, -2)) * hd**-0.5, dim=-1)
        o = (pr @ vh).transpose(1, 2).reshape(n, 1, d)
        o = F.linear(o, W["map_out_w"], W["map_out_b"])
        r = o
        o = F.layer_norm(o, (d,), W["map_ln_w"], W["map_ln_b"], cfg.ln_eps)
        o = r + F.linear(_act(F.linear(o, W["map_fc1_w"], W["map_fc1_b"]), cfg.act), W["map_fc2_w"], W["map_fc2_b"])
        feat = o[:, 0]
    emb = feat / torch.linalg.vector_norm(feat, dim=-1, keepdim=True)
    out = {"features": feat.numpy(), "embedding": emb.numpy()}
    if return_hidden:
        out["hidden"] = [t.numpy() for t in hidden]
    return out


# ---- aesthetic head (aesthetics.py:30-66) ---------------------------------------------------
AES_DIMS = (768, 1024, 128, 64, 16, 1)
AES_KEYS = ("layers.0", "layers.2", "layers.4", "layers.6", "layers.7")  # nn.Sequential indices with Linear


def random_aesthetic_mlp(seed: int = 0, in_dim: int = 768) -> dict[str, np.ndarray]:
    g = torch.Generator().manual_seed(1000 + seed)
    dims = (in_dim,) + AES_DIMS[1:]
    sd = {}
    for k, (i, o) in zip(AES_KEYS, zip(dims[:-1], dims[1:])):
        sd[k + ".weight"] = (torch.randn(o, i, generator=g) * i**-0.5).numpy().astype(np.float32)
        sd[k + ".bias"] = (torch.randn(o, generator=g) * 0.1).numpy().astype(np.float32)
    return sd


def aesthetic_mlp_forward(sd: dict, emb: np.ndarray) -> np.ndarray:
    """Layer-by-layer fp32 (the reference's form, aesthetics.py:44-53; Dropout = identity)."""
    x = torch.as_tensor(emb, dtype=torch.float32)
    for k in AES_KEYS:
        x = F.linear(x, torch.as_tensor(sd[k + ".weight"]), torch.as_tensor(sd[k + ".bias"]))
    return x.squeeze(1).numpy()


def collapse_aesthetic_mlp(sd: dict) -> tuple[np.ndarray, float]:
    """Fold the activation-free MLP into score = w . e + b (float64 fold, float32 result)."""
    a = np.eye(sd[AES_KEYS[0] + ".weight"].shape[1], dtype=np.float64)
    b = np.zeros(a.shape[0], dtype=np.float64)
    for k in AES_KEYS:
        wk, bk = sd[k + ".weight"].astype(np.float64), sd[k + ".bias"].astype(np.float64)
        a = wk @ a
        b = wk @ b + bk
    return a.reshape(-1).astype(np.float32), float(b.reshape(-1)[0])
