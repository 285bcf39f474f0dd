"""CPU: transcode-free clip cutting (cb_mp4_cut + ClipStreamCopyStage).  The cut clips are standalone MP4s
whose frames, decoded by libavcodec (cv2), are bit-identical to the corresponding frames of the source video."""

from __future__ import annotations

import uuid

import numpy as np
import pytest

from conftest import GOLDEN
from cosmos_curate_b200.data_model import Clip, SplitPipeTask, Video
from cosmos_curate_b200.runtime import mp4_index
from cosmos_curate_b200.stages import ClipStreamCopyStage
from cosmos_curate_b200.stages.clip_stream_copy import mp4_cut, span_sample_range
from tools import synth_h264


def _luma_frames(data: bytes, tmp_path, name="x.mp4"):
    import cv2

    p = tmp_path / name
    p.write_bytes(bytes(data))
    cap = cv2.VideoCapture(str(p))
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    out = []
    while True:
        ok, y = cap.read()
        if not ok:
            break
        out.append(y.copy())
    return out


def test_cut_is_a_standalone_mp4_with_identical_pictures(tmp_path):
    src = synth_h264.make_coded_clip(640, 368, 30, 4.0, seed=3, bitrate=1.0e6)
    idx = mp4_index(src)
    assert idx["n_samples"] == 120 and idx["n_sync"] == 4
    full = _luma_frames(src, tmp_path, "src.mp4")
    cut = mp4_cut(src, 30, 45)  # GOP 1 and half of GOP 2
    cidx = mp4_index(cut)
    assert cidx["n_samples"] == 45 and cidx["n_sync"] == 2 and (cidx["width"], cidx["height"]) == (640, 368)
    assert cidx["timescale"] == idx["timescale"] and cidx["pts"][0] == 0 and np.array_equal(np.diff(cidx["pts"]), np.diff(idx["pts"][30:75]))
    assert len(cut) < len(src) * 0.5
    got = _luma_frames(cut, tmp_path, "cut.mp4")
    assert len(got) == 45
    for k, y in enumerate(got):
        assert np.array_equal(y, full[30 + k]), k  # stream copy: the coded pictures are the source's own
    from cosmos_curate_b200._lib import CurateB200Error

    with pytest.raises(CurateB200Error, match="sync sample"):
        mp4_cut(src, 31, 10)
    with pytest.raises(CurateB200Error, match="outside the track"):
        mp4_cut(src, 90, 31)
    again = mp4_cut(cut, 0, 45)  # idempotent on its own output (stsd copied verbatim, same tables)
    assert np.array_equal(again, cut)


def test_span_rule_and_real_fixture(tmp_path):
    """ffmpeg -ss/-to semantics snapped to keyframes; the Sintel fixture (single GOP, High profile) cut to its first 5 seconds -
    BASELINE.json configs[0]'s "480p 5 s" clips."""
    src = synth_h264.make_coded_clip(640, 368, 30, 4.0, seed=4, bitrate=1.0e6)
    idx = mp4_index(src)
    assert span_sample_range(idx, (0.0, 4.0))[:2] == (0, 120)
    first, count, s0, s1 = span_sample_range(idx, (1.5, 2.5))  # frames 45..74 wanted; GOP starts at 30
    assert (first, count) == (30, 45) and s0 == pytest.approx(1.0) and s1 == pytest.approx(2.5)
    assert span_sample_range(idx, (2.0, 2.0 + 1 / 30))[:2] == (60, 1)
    with pytest.raises(ValueError):
        span_sample_range(idx, (10.0, 11.0))
    sintel = (GOLDEN / "sintel_clip_10s.mp4").read_bytes()
    sidx = mp4_index(sintel)
    first, count, _, _ = span_sample_range(sidx, (0.0, 5.0))
    assert (first, count) == (0, 120)
    cut = mp4_cut(sintel, first, count)
    full, got = _luma_frames(sintel, tmp_path, "s.mp4"), _luma_frames(cut, tmp_path, "c.mp4")
    assert len(got) == 120 and all(np.array_equal(a, b) for a, b in zip(got, full))
    assert mp4_index(cut)["duration"] == 120 * 512  # mdhd duration in media ticks (12288 / 24 fps)


def test_stage_contract_matches_clip_transcoding_stage(tmp_path):
    src = synth_h264.make_coded_clip(640, 368, 30, 4.0, seed=5, bitrate=1.0e6)
    clips = [Clip(uuid=uuid.uuid4(), source_video="v.mp4", span=s) for s in ((0.0, 2.0), (1.5, 3.0), (7.0, 9.0))]
    video = Video(input_video="v.mp4", encoded_data=src, clips=clips)
    video.metadata.duration = 4.0
    task = SplitPipeTask(session_id="s", video=video)
    stage = ClipStreamCopyStage(log_stats=True)
    stage.stage_setup()
    out = stage.process_data([task])
    assert out == [task] and "ClipStreamCopyStage" in task.stage_perf
    assert not video.encoded_data  # dropped like ClipTranscodingStage does (:289)
    a, b, c = video.clips
    assert mp4_index(a.encoded_data.resolve())["n_samples"] == 60 and a.span == pytest.approx((0.0, 2.0))
    assert mp4_index(b.encoded_data.resolve())["n_samples"] == 60 and b.span == pytest.approx((1.0, 3.0))  # snapped to the GOP at 1.0 s
    assert not c.encoded_data and "transcode" in c.errors  # span beyond the video: the reference's error key
    assert (video.num_total_clips, video.num_clip_chunks, video.clip_chunk_index) == (3, 1, 0)
    empty = SplitPipeTask(session_id="e", video=Video(input_video="e.mp4", clips=[clips[0]]))
    stage.process_data([empty])
    assert "ClipStreamCopyStage" in empty.video.errors  # "Please load video!" recorded, not raised (:283-287)


def test_stream_copy_stage_rechunks_like_the_transcoding_stage():
    """num_clips_per_chunk: ClipStreamCopyStage returns what ClipTranscodingStage returns (clip_extraction_stages.py:301) - one
    task per ~num_clips_per_chunk * 8 s of clips, bookkeeping fields set for the downstream weight / fraction arithmetic."""
    data = synth_h264.make_clip(320, 192, 30, 4.0, seed=2, gop=30)
    clips = [Clip(uuid=uuid.uuid4(), source_video="v.mp4", span=(float(s), float(s + 1))) for s in range(4)]
    video = Video(input_video="v.mp4", encoded_data=data, clips=clips)
    video.metadata.duration, video.metadata.size = 4.0, len(data)
    out = ClipStreamCopyStage(num_clips_per_chunk=1).process_data([SplitPipeTask(session_id="s", video=video)])  # 8 s of clips per chunk -> one chunk
    assert len(out) == 1 and len(out[0].video.clips) == 4 and all(c.encoded_data for c in out[0].video.clips)
    video2 = Video(input_video="v.mp4", encoded_data=data, clips=[Clip(uuid=uuid.uuid4(), source_video="v.mp4", span=(0.0, 4.0)), Clip(uuid=uuid.uuid4(), source_video="v.mp4", span=(0.0, 4.0)),
                                                                   Clip(uuid=uuid.uuid4(), source_video="v.mp4", span=(1.0, 4.0))])
    video2.metadata.duration, video2.metadata.size = 4.0, len(data)
    out = ClipStreamCopyStage(num_clips_per_chunk=1, snap_spans=False).process_data([SplitPipeTask(session_id="s", video=video2)])
    assert [len(t.video.clips) for t in out] == [2, 1]  # 4 s + 4 s reaches 8: the chunk closes
    assert [(t.video.num_total_clips, t.video.num_clip_chunks, t.video.clip_chunk_index) for t in out] == [(3, 2, 0), (3, 2, 1)]
    assert out[0].weight == pytest.approx(4.0 / 300 * 2 / 3) and not out[0].video.encoded_data
