"""Host-side frame-index math of the path (integer results; numpy, like the reference).

Mirrors, with the same names and argument meaning, the reference's
cosmos_curate/pipelines/video/utils/decoder_utils.py:
    FrameExtractionPolicy / FrameExtractionSignature   :86-117
    get_video_timestamps                               :230-278  (container demux -> cb_mp4_index)
    find_closest_indices / sample_closest              :281-386
    extract_frames' policy selection                   :640-652
and the LCM + stride rule of ClipFrameExtractionStage (clip_frame_extraction_stages.py:94-137).
This is O(#frames) scalar host work in the reference too; the decode it drives is NVDEC (runtime.Decoder).
"""

from __future__ import annotations

import enum
import math
from fractions import Fraction
from functools import reduce

import attrs
import numpy as np


class FrameExtractionPolicy(enum.Enum):
    first = 0
    middle = 1
    last = 2
    sequence = 3


@attrs.define
class FrameExtractionSignature:
    extraction_policy: FrameExtractionPolicy
    target_fps: float

    def to_str(self) -> str:
        return f"{self.extraction_policy!s}-{int(self.target_fps * 1000)}"


def timestamps_from_index(pts_ticks: np.ndarray, timescale: int) -> np.ndarray:
    """get_video_timestamps: float(packet.pts) * float(time_base), then sort as float32."""
    tb = float(Fraction(1, int(timescale)))
    ts = [float(int(p)) * tb for p in pts_ticks]
    return np.sort(np.array(ts, dtype=np.float32))


def find_closest_indices(src, dst):
    src, dst = np.asarray(src), np.asarray(dst)
    hi = np.clip(np.searchsorted(src, dst), 1, len(src) - 1)
    lo = hi - 1
    pick_hi = np.abs(dst - src[hi]) < np.abs(dst - src[lo])  # ties stay left
    idx = np.where(pick_hi, hi, lo)
    idx[dst >= src[-1]] = len(src) - 1
    return idx.astype(np.int32)


def sample_closest(src, sample_rate, start=None, stop=None, endpoint=True, dedup=True):
    if sample_rate <= 0:
        msg = f"Sample rate must be greater than 0, got {sample_rate=}"
        raise ValueError(msg)
    src = np.asarray(src)
    step = 1.0 / sample_rate
    first = src[0] if start is None else start
    last = src[-1] if stop is None else stop
    if endpoint:
        last += step * 0.5
    wanted = np.arange(first, last, step, dtype=np.float32)
    idx = find_closest_indices(src, wanted)
    if not endpoint and np.isclose(wanted[-1], last):
        idx, wanted = idx[:-1], wanted[:-1]
    if dedup:
        idx, counts = np.unique(idx, return_counts=True)
        return idx.astype(np.int32), counts.astype(np.int32), wanted
    return idx, np.ones_like(idx, dtype=np.int32), wanted


def policy_timestamps(all_ts: np.ndarray, policy: FrameExtractionPolicy) -> np.ndarray:
    if len(all_ts) == 0:
        msg = "Can't extract frames from empty video"
        raise ValueError(msg)
    if policy == FrameExtractionPolicy.sequence or len(all_ts) == 1:
        return all_ts
    if policy == FrameExtractionPolicy.middle:
        n = len(all_ts)
        i = n // 2 - 1 if n % 2 == 0 else n // 2
        return all_ts[i : i + 1]
    msg = "Extraction policies apart from Sequence and Middle not available yet"
    raise NotImplementedError(msg)


def frame_ids(all_ts: np.ndarray, policy: FrameExtractionPolicy, fps: float):
    """extract_frames -> decode_video_cpu -> sample_closest: (ids, counts) in display order."""
    ts = policy_timestamps(all_ts, policy)
    ids, counts, _ = sample_closest(ts, fps, start=ts[0], stop=ts[-1], endpoint=True, dedup=True)
    return ids, counts


def lcm_multiple(fps: list) -> float | int:
    def lcm(a, b):
        return abs(a * b) // math.gcd(int(a), int(b))

    return reduce(lcm, fps)


def plan_extraction(all_ts: np.ndarray, policies, target_fps: list):
    """Which decoded frames every signature receives (clip_frame_extraction_stages.py:111-152).

    Returns {signature: int32 array of display-order frame ids, repeats expanded}."""
    use_lcm = len(target_fps) > 1 and all((f.is_integer() if isinstance(f, float) else isinstance(f, int)) for f in target_fps)
    plan = {}
    for policy in policies:
        if use_lcm:
            lcm = lcm_multiple(target_fps)
            ids, counts = frame_ids(all_ts, policy, lcm)
            expanded = np.repeat(ids, counts)
            for f in target_fps:
                plan[FrameExtractionSignature(policy, f).to_str()] = expanded[:: int(lcm / f)]
        else:
            for f in target_fps:
                ids, counts = frame_ids(all_ts, policy, f)
                plan[FrameExtractionSignature(policy, f).to_str()] = np.repeat(ids, counts)
    return plan


def pynvc_target_size(width: int, height: int, target_w: int = -1, target_h: int = -1) -> tuple[int, int]:
    """(target_w, target_h) of the pynvc thumbnail path when either side is -1 (VideoBatchDecoder.__call__,
    nvcodec_utils.py:129-136): downscale by width // 256 (no downscale below 256 px), Python round() (half to even)."""
    if target_w != -1 and target_h != -1:
        return int(target_w), int(target_h)
    factor = 1 if width < 256 else width // 256
    return round(width / factor), round(height / factor)


def span_frame_ids(ts: np.ndarray, span: tuple[float, float], fps: float) -> np.ndarray:
    """Source-video frame indices that a clip cut at `span` = (start_s, end_s) and then sampled at `fps` would show.

    The cut keeps the source frames whose PTS lies in [start, end) (half a frame period of slack absorbs the float32 PTS vs
    float64 span rounding: TransNetV2 spans are frame_index / fps, transnetv2_extraction_stages.py:201-206), re-times them from
    the first kept frame - what ClipTranscodingStage's `-ss start -t duration` produces - and samples that clip with the same
    rule as a stand-alone clip (sample_closest with endpoint, decoder_utils.py:315-386).  Returns int32 ids, repeats expanded."""
    ts = np.asarray(ts, dtype=np.float32)
    if len(ts) == 0:
        msg = "video has no frames"
        raise ValueError(msg)
    dt = float(np.median(np.diff(ts))) if len(ts) > 1 else 0.0
    start, end = float(span[0]), float(span[1])
    keep = np.flatnonzero((ts >= start - 0.5 * dt) & (ts < end - 0.5 * dt))
    if len(keep) == 0:
        msg = f"span {span} selects no frame of the source video"
        raise ValueError(msg)
    clip_ts = (ts[keep] - ts[keep[0]]).astype(np.float32)
    ids, counts = frame_ids(clip_ts, FrameExtractionPolicy.sequence, fps)
    return np.repeat(keep[ids], counts).astype(np.int32)


@attrs.define
class VideoMetadata:
    """Same fields as the reference's VideoMetadata (decoder_utils.py:57-84)."""

    height: int
    width: int
    fps: float
    num_frames: int
    video_codec: str
    pixel_format: str
    video_duration: float
    bit_rate_k: int
    format_name: str = "unknown"
    audio_codec: str | None = None

    @property
    def length_s(self) -> float:
        return self.video_duration


def video_metadata_from_index(idx: dict) -> VideoMetadata:
    """extract_video_metadata (decoder_utils.py:120-197) without the ffprobe subprocess + temp file: the same
    quantities ffprobe prints for an MP4 video stream, derived from the moov index (cb_mp4_index).
        avg_frame_rate = nb_frames / stream duration, fps = num / den; num_frames = int(duration * fps);
        bit_rate_k = int(bit_rate / 1024) with bit_rate = 8 * sample bytes / duration."""
    ts, dur, n = int(idx["timescale"]), int(idx["duration"]), int(idx["n_samples"])
    if dur <= 0 or ts <= 0:
        msg = "Could not find `duration` in video metadata."
        raise KeyError(msg)
    rate = Fraction(n * ts, dur).limit_denominator(60000)
    fps = rate.numerator / rate.denominator
    duration_s = float(f"{dur / ts:.6f}")  # ffprobe prints seconds with 6 decimals
    bit_rate = int(8 * int(idx["sample_bytes"]) / (dur / ts))
    return VideoMetadata(
        height=int(idx["height"]), width=int(idx["width"]), fps=fps, num_frames=int(duration_s * fps),
        video_codec={4: "h264", 8: "hevc"}.get(int(idx["codec"]), "unknown"), pixel_format="yuv420p", video_duration=duration_s,
        bit_rate_k=int(bit_rate / 1024), format_name="mov,mp4,m4a,3gp,3g2,mj2",
    )  # fmt: skip
