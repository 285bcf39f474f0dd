"""Oracle: TransNetV2 shot-transition model + the shot logic around it.

CPU restatement (torch fp32 functional ops + numpy integer arithmetic) of

  _TransNetV2.forward                cosmos_curate/models/transnetv2.py:103-148
    StackedDDCNNV2.forward           :190-221   (blocks -> relu -> + first block's output -> avg-pool (1,2,2))
    DilatedDCNNV2.forward            :252-276   (4 dilated branches -> concat -> BatchNorm3d(eps=1e-3, eval) -> relu?)
    Conv3DConfigurable               :279-343   ((1,3,3) conv in->2F no bias, then (3,1,1) conv 2F->F dilation d)
    FrameSimilarity.forward          :377-418   (spatial means -> Linear -> L2 norm -> T x T cosine -> 101-window -> Linear -> relu)
    ColorHistograms                  :440-527   (512-bin RGB histogram per frame -> L2 norm -> T x T -> window -> Linear -> relu)
  _get_batches / _get_predictions    cosmos_curate/pipelines/video/clipping/transnetv2_extraction_stages.py:215-264
  _get_scenes                        :267-299
  _get_filtered_scenes / _crop_scenes / _create_spans   :302-392
  TransNetV2ClipExtractionStage._get_min_length/_get_max_length/process_data   :150-212

The reference's weights (Sn4kehead/TransNetV2) are not in this image, so parity is pinned on seeded weights:
`random_state_dict(seed)` builds a full state_dict with the reference's key names, oracle/make_golden.py loads it into
the reference's own `_TransNetV2` and records its outputs (tests/golden/transnetv2_ref.npz); this file must reproduce
them, and the CUDA path is then checked against this file.

Quirk kept on purpose: `_get_batches` never pads the END of a video (its `end_idx > total_frames` branch is dead,
:232-235), so the last one or two windows reach the model with fewer than 100 frames.

Test infrastructure only (see oracle/__init__.py).
"""

from __future__ import annotations

import math
import uuid

import numpy as np
import torch
import torch.nn.functional as F

RF, RL, RS, RD = 16, 3, 2, 1024  # transnetv2.py:43-46 defaults
LOOKUP = 101
DILATIONS = (1, 2, 4, 8)
FRAME_H, FRAME_W = 27, 48


def layer_plan() -> list[dict]:
    """[(stack, block, in_filters, filters)] in execution order (transnetv2.py:66-72, 176-186)."""
    plan = []
    for s in range(RL):
        filters = RF * 2**s
        stack_in = 3 if s == 0 else (RF * 2 ** (s - 1)) * 4
        for b in range(RS):
            plan.append({"stack": s, "block": b, "in": stack_in if b == 0 else filters * 4, "filters": filters, "relu": b != RS - 1})
    return plan


def random_state_dict(seed: int = 0) -> dict[str, np.ndarray]:
    """A full state_dict (reference key names/shapes) with He-scaled convs and non-trivial BatchNorm statistics."""
    rng = np.random.default_rng(seed)
    sd: dict[str, np.ndarray] = {}

    def rn(shape, std):
        return (rng.standard_normal(shape) * std).astype(np.float32)

    for lp in layer_plan():
        p = f"SDDCNN.{lp['stack']}.DDCNN.{lp['block']}"
        cin, f = lp["in"], lp["filters"]
        for d in DILATIONS:
            sd[f"{p}.Conv3D_{d}.layers.0.weight"] = rn((2 * f, cin, 1, 3, 3), math.sqrt(2.0 / (cin * 9)))
            sd[f"{p}.Conv3D_{d}.layers.1.weight"] = rn((f, 2 * f, 3, 1, 1), math.sqrt(1.0 / (2 * f * 3)))
        c = 4 * f
        sd[f"{p}.bn.weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        sd[f"{p}.bn.bias"] = rn((c,), 0.1)
        sd[f"{p}.bn.running_mean"] = rn((c,), 0.1)
        sd[f"{p}.bn.running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        sd[f"{p}.bn.num_batches_tracked"] = np.array(0, dtype=np.int64)
    sim_in = sum((RF * 2**i) * 4 for i in range(RL))
    sd["frame_sim_layer.projection.weight"] = rn((128, sim_in), 1.0 / math.sqrt(sim_in))
    sd["frame_sim_layer.projection.bias"] = rn((128,), 0.05)
    sd["frame_sim_layer.fc.weight"] = rn((128, LOOKUP), 1.0 / math.sqrt(LOOKUP))
    sd["frame_sim_layer.fc.bias"] = rn((128,), 0.05)
    sd["color_hist_layer.fc.weight"] = rn((128, LOOKUP), 1.0 / math.sqrt(LOOKUP))
    sd["color_hist_layer.fc.bias"] = rn((128,), 0.05)
    fc_in = ((RF * 2 ** (RL - 1)) * 4) * 3 * 6 + 256
    sd["fc1.weight"] = rn((RD, fc_in), 1.0 / math.sqrt(fc_in))
    sd["fc1.bias"] = rn((RD,), 0.05)
    sd["cls_layer1.weight"] = rn((1, RD), 2.0 / math.sqrt(RD))
    # the offset centres the seeded logits (mean ~ +4.8 on synthetic_frames) so probabilities straddle the thresholds tested
    sd["cls_layer1.bias"] = rn((1,), 0.05) - np.float32(4.75)
    sd["cls_layer2.weight"] = rn((1, RD), 2.0 / math.sqrt(RD))
    sd["cls_layer2.bias"] = rn((1,), 0.05)
    return sd


def synthetic_frames(n: int, seed: int = 0, cuts: tuple[int, ...] = ()) -> np.ndarray:
    """uint8 [n, 27, 48, 3]: smooth drifting scenes with hard cuts at the given frame numbers."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:FRAME_H, 0:FRAME_W].astype(np.float32)
    out = np.empty((n, FRAME_H, FRAME_W, 3), dtype=np.uint8)
    bounds = [0, *sorted(cuts), n]
    for a, b in zip(bounds[:-1], bounds[1:]):
        base = rng.uniform(20, 235, 3)
        fx, fy, ph = rng.uniform(0.05, 0.4, 3), rng.uniform(0.05, 0.4, 3), rng.uniform(0, 6.28, 3)
        amp = rng.uniform(10, 60, 3)
        for t in range(a, b):
            for c in range(3):
                v = base[c] + amp[c] * np.sin(fx[c] * xx + fy[c] * yy + ph[c] + 0.07 * (t - a))
                out[t, :, :, c] = np.clip(v + rng.integers(-3, 4, (FRAME_H, FRAME_W)), 0, 255).astype(np.uint8)
    return out


# ---- model ----------------------------------------------------------------------------------------------------------
def _t(sd, k):
    return torch.as_tensor(sd[k])


def _windowed(sim: torch.Tensor) -> torch.Tensor:
    """[B,T,T] -> [B,T,101]: row t keeps columns t-50 .. t+50, zero outside (transnetv2.py:393-416 / :503-522)."""
    b, t, _ = sim.shape
    half = (LOOKUP - 1) // 2
    padded = F.pad(sim, [half, half])
    idx = torch.arange(t).view(t, 1) + torch.arange(LOOKUP).view(1, LOOKUP)  # padded column = t + j
    return padded[:, torch.arange(t).view(t, 1), idx]


def color_histograms(frames: torch.Tensor) -> torch.Tensor:
    """uint8 [B,T,H,W,3] -> float32 [B,T,512] unit-norm (transnetv2.py:440-486)."""
    b, t = frames.shape[:2]
    f = frames.to(torch.int32).view(b * t, -1, 3)
    bins = ((f[..., 0] >> 5) << 6) + ((f[..., 1] >> 5) << 3) + (f[..., 2] >> 5)
    hist = torch.zeros(b * t, 512, dtype=torch.int32)
    hist.scatter_add_(1, bins.long(), torch.ones_like(bins, dtype=torch.int32))
    return F.normalize(hist.view(b, t, 512).float(), p=2, dim=2)


@torch.no_grad()
def forward(sd: dict, frames: np.ndarray | torch.Tensor, return_parts: bool = False):
    """uint8 [B,T,27,48,3] -> float32 [B,T,1] transition probabilities (one_hot head only, transnetv2.py:142-148)."""
    inputs = torch.as_tensor(frames)
    assert inputs.dtype == torch.uint8 and list(inputs.shape[2:]) == [FRAME_H, FRAME_W, 3]
    x = inputs.permute(0, 4, 1, 2, 3).float() / 255.0
    feats = []
    parts = {}
    plan = layer_plan()
    for s in range(RL):
        shortcut = None
        for lp in (p for p in plan if p["stack"] == s):
            p = f"SDDCNN.{s}.DDCNN.{lp['block']}"
            branches = []
            for d in DILATIONS:
                y = F.conv3d(x, _t(sd, f"{p}.Conv3D_{d}.layers.0.weight"), padding=(0, 1, 1))
                y = F.conv3d(y, _t(sd, f"{p}.Conv3D_{d}.layers.1.weight"), padding=(d, 0, 0), dilation=(d, 1, 1))
                branches.append(y)
            x = torch.cat(branches, dim=1)
            x = F.batch_norm(x, _t(sd, f"{p}.bn.running_mean"), _t(sd, f"{p}.bn.running_var"), _t(sd, f"{p}.bn.weight"), _t(sd, f"{p}.bn.bias"),
                             training=False, eps=1e-3)  # fmt: skip
            if lp["relu"]:
                x = F.relu(x)
            if shortcut is None:
                shortcut = x
        x = F.relu(x) + shortcut
        x = F.avg_pool3d(x, kernel_size=(1, 2, 2))
        feats.append(x)
        parts[f"stack{s}"] = x
    x = x.permute(0, 2, 3, 4, 1)
    x = x.reshape(x.shape[0], x.shape[1], -1)

    f = torch.cat([torch.mean(v, dim=[3, 4]) for v in feats], dim=1).transpose(1, 2)  # [B,T,448]
    f = F.linear(f, _t(sd, "frame_sim_layer.projection.weight"), _t(sd, "frame_sim_layer.projection.bias"))
    f = F.normalize(f, p=2, dim=2)
    sim = _windowed(torch.bmm(f, f.transpose(1, 2)))
    fs = F.relu(F.linear(sim, _t(sd, "frame_sim_layer.fc.weight"), _t(sd, "frame_sim_layer.fc.bias")))

    h = color_histograms(inputs)
    hsim = _windowed(torch.bmm(h, h.transpose(1, 2)))
    ch = F.relu(F.linear(hsim, _t(sd, "color_hist_layer.fc.weight"), _t(sd, "color_hist_layer.fc.bias")))

    x = torch.cat([ch, torch.cat([fs, x], 2)], 2)  # [colour(128), similarity(128), trunk(4608)]
    parts["concat"] = x
    x = F.relu(F.linear(x, _t(sd, "fc1.weight"), _t(sd, "fc1.bias")))
    logit = F.linear(x, _t(sd, "cls_layer1.weight"), _t(sd, "cls_layer1.bias"))
    prob = torch.sigmoid(logit)
    if return_parts:
        parts["logit"] = logit
        return prob, parts
    return prob


# ---- windowing + shot logic -------------------------------------------------------------------------------------------
def window_plan(total: int) -> list[tuple[int, int, int]]:
    """[(first_frame, n_real_frames, n_front_pad)] per window (transnetv2_extraction_stages.py:215-236)."""
    out = []
    rem = -total % 50
    for i in range(0, total + rem, 50):
        a, b = max(i - 25, 0), min(i + 75, total)
        out.append((a, b - a, max(25 - i, 0)))
    return out


def windows(frames: np.ndarray) -> list[np.ndarray]:
    out = []
    for a, n, pad in window_plan(len(frames)):
        w = frames[a : a + n]
        if pad:
            w = np.concatenate([np.repeat(frames[:1], pad, axis=0), w], axis=0)
        out.append(w)
    return out


def probabilities(sd: dict, frames: np.ndarray) -> np.ndarray:
    """float32 [n]: per-frame transition probability, windows stitched as _get_predictions does (:253-263)."""
    parts = [forward(sd, w[None])[0, 25:75, 0] for w in windows(frames)]
    return torch.cat(parts)[: len(frames)].numpy()


def predictions(sd: dict, frames: np.ndarray, threshold: float) -> np.ndarray:
    """uint8 [n,1] (:263-264).  torch compares the fp32 tensor with the Python scalar in the tensor's dtype, i.e.
    prob > float32(threshold)."""
    p = torch.from_numpy(probabilities(sd, frames))
    return (p > threshold).to(torch.uint8).numpy().reshape(-1, 1)


def scenes_from_predictions(pred: np.ndarray, entire_scene_as_clip: bool) -> np.ndarray:
    """0/1 per frame -> int32 [k,2] (start, end) pairs (:267-299)."""
    flat = [int(v) for v in np.asarray(pred).reshape(-1)]
    scenes = []
    prev, start, cur, i = 0, 0, -1, 0
    for i, cur in enumerate(flat):
        if prev == 1 and cur == 0:
            start = i
        if prev == 0 and cur == 1 and i != 0:
            scenes.append((start, i))
        prev = cur
    if scenes and cur == 0:
        scenes.append((start, i))
    if not scenes and entire_scene_as_clip:
        scenes.append((0, len(flat)))
    return np.array(scenes, dtype=np.int32).reshape(-1, 2)


def stride_spans(start: int, end: int, max_length: int, min_length: int | None) -> list[list[int]]:
    """(:369-392)"""
    spans, cur = [], start
    while cur < end:
        stop = min(cur + max_length, end)
        if min_length and stop - cur < min_length and stop == end:
            break
        spans.append([cur, stop])
        cur = stop
    return spans


def filter_scenes(scenes: np.ndarray, min_length=None, max_length=None, max_length_mode="truncate", crop_length=None) -> np.ndarray:
    """(:302-366)"""
    scenes = np.array(scenes, dtype=np.int32).reshape(-1, 2)
    if max_length is not None:
        if max_length_mode == "truncate":
            scenes[:, 1] = np.minimum(scenes[:, 0] + max_length, scenes[:, 1])
        elif max_length_mode == "stride":
            new = []
            for a, b in scenes:
                new.extend(stride_spans(int(a), int(b), max_length, min_length))
            scenes = np.array(new, dtype=np.int32).reshape(-1, 2)
        else:
            raise NotImplementedError(max_length_mode)
    if crop_length is not None:
        cropped = np.stack([scenes[:, 0] + crop_length, scenes[:, 1] - crop_length]).T
        scenes = cropped[(cropped[:, 1] - cropped[:, 0]) > 0]
    if min_length is not None:
        scenes = scenes[(scenes[:, 1] - scenes[:, 0]) >= min_length]
    return scenes


def stage_lengths(framerate: float, min_length_s=2.0, min_length_frames=48, max_length_s=60.0, crop_s=0.5):
    """(min_length, max_length, crop_length) in frames, as the stage derives them (:150-158, :196)."""
    mn = math.ceil(min_length_s * framerate) if min_length_s is not None else None
    if min_length_frames is not None:
        mn = max(mn, min_length_frames) if mn is not None else min_length_frames
    mx = math.ceil(max_length_s * framerate) if max_length_s is not None else None
    crop = int(crop_s * framerate) if crop_s else None
    return mn, mx, crop


def clips_for_video(name: str, scenes: np.ndarray, framerate: float, limit_clips: int = 0) -> list[tuple[uuid.UUID, tuple[float, float]]]:
    """(uuid5, span seconds) per scene (:198-209)."""
    out = []
    for a, b in scenes:
        out.append((uuid.uuid5(uuid.NAMESPACE_URL, f"{name}_{a}_{b}"), (float(a) / framerate, float(b) / framerate)))
        if limit_clips > 0 and len(out) >= limit_clips:
            break
    return out
