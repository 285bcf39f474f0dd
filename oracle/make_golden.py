"""Generate tests/golden/* by executing the REFERENCE's own code (build container only).

    python -m oracle.make_golden            # writes tests/golden/*.npz

The reference is Python, so its hot-path modules are imported from /root/reference
(oracle/ref_import.py installs the two stubs SURVEY.md 8c documents).  The vectors are what
travels to the GPU box; /root/reference itself does not.

Files written (all small):
  sampling_ref.npz        reference decoder_utils.sample_closest / find_closest_indices on a
                          spread of PTS tracks (CFR 24/30/29.97, VFR, dropped frames, offsets)
  clip_preprocess_ref.npz reference _CLIPImageEmbeddings.transforms (clip.py:48-62) on u8 frames
  clip_tiny_ref.npz       reference _CLIPImageEmbeddings.__call__ (clip.py:64-74), unmodified class,
                          on a tiny seeded CLIPModel saved with save_pretrained (weights included)
  aesthetic_ref.npz       reference aesthetics.MLP (aesthetics.py:30-66) seeded state_dict + outputs
  siglip_tiny_hf.npz      transformers SiglipVisionModel tiny seeded (stand-in; not in the reference)

Test infrastructure only (see oracle/__init__.py).
"""

from __future__ import annotations

import json
from fractions import Fraction
import tempfile
from pathlib import Path

import numpy as np
import torch

from oracle import ref_import, vit

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def pts_tracks() -> dict[str, np.ndarray]:
    """PTS tracks as the reference builds them: float(pts) * float(time_base) -> float32, sorted
    (decoder_utils.py:268-278)."""
    rng = np.random.default_rng(7)
    tracks = {}

    def from_ticks(ticks, timescale):
        tb = float(Fraction(1, timescale))  # PyAV time_base is a Fraction
        return np.sort(np.array([float(t) * tb for t in ticks], dtype=np.float32))

    tracks["cfr24_240_ts12288"] = from_ticks(np.arange(240) * 512, 12288)  # the Sintel fixture's timing
    tracks["cfr24_720_ts12288"] = from_ticks(np.arange(720) * 512, 12288)
    tracks["cfr30_300_ts15360"] = from_ticks(np.arange(300) * 512, 15360)
    tracks["cfr30_300_ts90000"] = from_ticks(np.arange(300) * 3000, 90000)
    tracks["ntsc_2997_300"] = from_ticks(np.arange(300) * 1001, 30000)
    tracks["cfr60_600_ts60"] = from_ticks(np.arange(600), 60)
    tracks["cfr24_120_5s"] = from_ticks(np.arange(120) * 512, 12288)  # config C1: 5 s
    tracks["offset_start_30"] = from_ticks(3003 + np.arange(150) * 3000, 90000)  # non-zero first PTS
    jitter = np.cumsum(rng.integers(2000, 4000, size=200))
    tracks["vfr_200"] = from_ticks(jitter, 90000)
    keep = np.ones(300, dtype=bool)
    keep[rng.choice(300, size=40, replace=False)] = False
    keep[0] = keep[-1] = True
    tracks["dropped_30"] = from_ticks((np.arange(300) * 3000)[keep], 90000)
    tracks["single_frame"] = np.array([0.0], dtype=np.float32)
    tracks["two_frames"] = np.array([0.0, 1.0 / 30], dtype=np.float32)
    tracks["lowfps_5_50"] = from_ticks(np.arange(50) * 18000, 90000)  # 5 fps source: supersampling
    return tracks


def gen_sampling() -> None:
    du = ref_import.decoder_utils()
    out, meta = {}, []
    case = 0
    for name, ts in pts_tracks().items():
        for rate in (0.5, 1.0, 2.0, 4.0, 8.0, 30.0):
            for endpoint in (True, False):
                if len(ts) < 2 and not endpoint:
                    continue
                try:
                    ids, counts, samples = du.sample_closest(ts, rate, start=ts[0], stop=ts[-1], endpoint=endpoint, dedup=True)
                except Exception as e:  # noqa: BLE001 - record reference failures too
                    meta.append({"case": case, "track": name, "rate": rate, "endpoint": endpoint, "raises": type(e).__name__})
                    case += 1
                    continue
                out[f"c{case}_ts"], out[f"c{case}_ids"], out[f"c{case}_counts"] = ts, ids, counts
                meta.append({"case": case, "track": name, "rate": rate, "endpoint": endpoint, "n": int(len(ids))})
                case += 1
    # find_closest_indices on random monotone arrays incl. exact ties and out-of-range queries
    rng = np.random.default_rng(11)
    for k in range(8):
        src = np.sort(rng.uniform(0, 10, size=rng.integers(2, 60))).astype(np.float32)
        dst = np.sort(np.concatenate([rng.uniform(-1, 12, size=40), (src[:-1] + src[1:]) / 2, src])).astype(np.float32)
        out[f"f{k}_src"], out[f"f{k}_dst"] = src, dst
        out[f"f{k}_idx"] = du.find_closest_indices(src, dst)
    # extract_frames policy path: `middle` quirk (one timestamp -> id 0)
    for name in ("cfr24_240_ts12288", "cfr30_300_ts90000", "single_frame"):
        ts = pts_tracks()[name]
        n = len(ts)
        mid = ts[(n // 2 - 1 if n % 2 == 0 else n // 2) :][:1]
        ids, counts, _ = du.sample_closest(mid, 1.0, start=mid[0], stop=mid[-1], endpoint=True, dedup=True)
        out[f"m_{name}_ts"], out[f"m_{name}_ids"], out[f"m_{name}_counts"] = ts, ids, counts
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT / "sampling_ref.npz", **out)
    print("sampling_ref.npz:", case, "cases")


def test_frames() -> dict[str, np.ndarray]:
    rng = np.random.default_rng(3)
    frames = {}

    def scene(h, w, seed):
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        base = np.stack(
            [128 + 100 * np.sin(xx / (7 + seed) + yy / 13), 128 + 90 * np.cos(xx / 11 - yy / (5 + seed)), (xx * 255 / w + yy * 3) % 256],
            axis=-1,
        )
        noise = rng.integers(-40, 41, size=(h, w, 3))
        img = np.clip(base + noise, 0, 255).astype(np.uint8)
        img[h // 4 : h // 4 + 9, :, :] = 255  # hard edges: overshoot/clamp exercise
        img[:, w // 3 : w // 3 + 5, :] = 0
        return img

    frames["land_270x480"] = np.stack([scene(270, 480, s) for s in range(2)])
    frames["port_480x270"] = scene(480, 270, 5)[None]
    frames["square_224"] = scene(224, 224, 6)[None]  # resize is the identity
    frames["small_100x160"] = scene(100, 160, 7)[None]  # upscale (scale < 1)
    frames["odd_301x225"] = scene(301, 225, 8)[None]  # odd crop offset (banker's rounding)
    frames["rand_360x640"] = rng.integers(0, 256, size=(1, 360, 640, 3), dtype=np.uint8)
    return frames


def _tiny_clip_dir(tmp: Path):
    from transformers import CLIPConfig, CLIPModel

    torch.manual_seed(0)
    cfg = CLIPConfig(
        text_config={"hidden_size": 32, "intermediate_size": 64, "num_hidden_layers": 1, "num_attention_heads": 2,
                     "vocab_size": 64, "max_position_embeddings": 8, "projection_dim": 64},
        vision_config={"hidden_size": 128, "intermediate_size": 256, "num_hidden_layers": 2, "num_attention_heads": 2,
                       "image_size": 224, "patch_size": 32, "projection_dim": 64},
        projection_dim=64,
    )  # fmt: skip
    model = CLIPModel(cfg).eval()
    with torch.no_grad():  # HF init leaves LN at (1,0) and biases at 0: perturb so every term is exercised
        g = torch.Generator().manual_seed(1)
        for n, p in model.named_parameters():
            if n.endswith("bias") or "layer_norm" in n or "layrnorm" in n or "layernorm" in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    model.save_pretrained(tmp)
    return model


def gen_clip() -> None:
    cm = ref_import.clip_module()
    with tempfile.TemporaryDirectory() as td:
        hf = _tiny_clip_dir(Path(td))
        ref_import.register_weights_dir("tiny-clip", Path(td))
        ref = cm._CLIPImageEmbeddings("tiny-clip")  # the reference class, unmodified
        # transformers >= 5 returns an output object from get_image_features;
        # adapt the return type only, the reference's __call__ body runs as written.
        orig = ref.clip.get_image_features

        def _tensor_get_image_features(*a, **k):
            o = orig(*a, **k)
            return o if isinstance(o, torch.Tensor) else o.pooler_output

        ref.clip.get_image_features = _tensor_get_image_features
        pre, emb = {}, {}
        for name, fr in test_frames().items():
            x = torch.from_numpy(fr).permute(0, 3, 1, 2)
            pre[f"{name}_in"] = fr
            pre[f"{name}_out"] = ref.transforms(x).numpy()
            emb[f"{name}_in"] = fr
            emb[f"{name}_emb"] = ref(fr).cpu().numpy()
        np.savez_compressed(OUT / "clip_preprocess_ref.npz", **pre)
        cfg, w = vit.weights_from_hf_clip(hf)
        emb["cfg"] = np.frombuffer(json.dumps(cfg.to_dict()).encode(), dtype=np.uint8)
        for k, v in w.items():
            emb["w_" + k] = v
        np.savez_compressed(OUT / "clip_tiny_ref.npz", **emb)
    print("clip_preprocess_ref.npz, clip_tiny_ref.npz")


def gen_aesthetic() -> None:
    am = ref_import.aesthetics_module()
    torch.manual_seed(5)
    mlp = am.MLP().eval()
    sd = {k: v.detach().numpy() for k, v in mlp.state_dict().items()}
    rng = np.random.default_rng(9)
    e = rng.standard_normal((16, 768)).astype(np.float32)
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    with torch.no_grad():
        y = mlp(torch.from_numpy(e)).squeeze(1).numpy()
    np.savez_compressed(OUT / "aesthetic_ref.npz", emb=e, score=y, **{"sd_" + k: v for k, v in sd.items()})
    print("aesthetic_ref.npz")


def gen_siglip() -> None:
    from transformers import SiglipVisionConfig, SiglipVisionModel

    torch.manual_seed(2)
    cfg = SiglipVisionConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, image_size=64, patch_size=16)
    m = SiglipVisionModel(cfg).eval()
    with torch.no_grad():
        g = torch.Generator().manual_seed(3)
        for n, p in m.named_parameters():
            if n.endswith("bias") or "norm" in n or "probe" in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    x = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        y = m(pixel_values=x).pooler_output.numpy()
    c, w = vit.weights_from_hf_siglip(m)
    out = {"x": x.numpy(), "pooled": y, "cfg": np.frombuffer(json.dumps(c.to_dict()).encode(), dtype=np.uint8)}
    out.update({"w_" + k: v for k, v in w.items()})
    np.savez_compressed(OUT / "siglip_tiny_hf.npz", **out)
    print("siglip_tiny_hf.npz")


def main() -> None:
    assert ref_import.available(), "needs /root/reference (build container)"
    OUT.mkdir(parents=True, exist_ok=True)
    gen_sampling()
    gen_clip()
    gen_aesthetic()
    gen_siglip()


if __name__ == "__main__":
    main()
