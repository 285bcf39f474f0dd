"""Fixed-stride clip spans (host logic, no GPU): cosmos_curate_b200/spans.py + FixedStrideExtractorStage against
(a) vectors produced by the reference's own functions executed from source (tests/golden/fixed_stride_ref.json, floats compared
bit for bit) and (b) the known answers of the reference's test-suite
(tests/cosmos_curate/pipelines/video/clipping/test_fixed_stride_extraction.py)."""

from __future__ import annotations

import json
import uuid

import numpy as np
import pytest

from conftest import GOLDEN
from cosmos_curate_b200 import spans as S
from cosmos_curate_b200.data_model import Clip, SplitPipeTask, Video
from cosmos_curate_b200.stages import FixedStrideExtractorStage


def _video(name="v.mp4", n_frames=900, fps=30.0, t0=0.0, with_ts=True):
    v = Video(input_video=name)
    m = v.metadata
    m.num_frames, m.framerate, m.height, m.width, m.duration, m.video_codec = n_frames, fps, 1080, 1920, n_frames / fps if fps else 0.0, "h264"
    if with_ts:
        v.timestamps = (t0 + np.arange(n_frames) / (fps or 1.0)).astype(np.float32)
    return v


def test_spans_and_uuids_match_the_reference_functions_bit_for_bit():
    g = json.loads((GOLDEN / "fixed_stride_ref.json").read_text())
    assert len(g["spans"]) == 36
    for case in g["spans"]:
        args = [float.fromhex(a) for a in case["args"]]
        spans = S.make_spans_fixed_stride(*args)
        assert [[float.hex(a), float.hex(b)] for a, b in spans] == case["spans"], case["args"]
        assert [str(u) for u in S.make_clip_uuids("s3://bucket/session-7", spans)] == case["uuids"]
    for case in g["populate"]:
        vids = [_video(n, nf, float.fromhex(fps), float.fromhex(t0)) for n, nf, fps, t0 in case["videos"]]
        clip_len, stride, min_len = (float.fromhex(a) for a in case["args"][:3])
        S.populate_clips_fixed_stride(vids, "session/" + case["name"], clip_len, stride, min_len, limit_clips=case["args"][3])
        got = [[[str(c.uuid), c.source_video, float.hex(c.span[0]), float.hex(c.span[1])] for c in v.clips] for v in vids]
        assert got == case["clips"], case["name"]
        S.assert_video_clip_alignment(vids)


def test_reference_known_answers():
    """test_fixed_stride_extraction.py:662-805, 959-990 (spans) and :605-660 (durations)."""
    assert S.make_spans_fixed_stride(0.0, 30.0, 10.0, 10.0, 5.0) == [(0.0, 10.0), (10.0, 20.0), (20.0, 30.0)]
    assert S.make_spans_fixed_stride(0.0, 30.0, 10.0, 5.0, 5.0) == [(0.0, 10.0), (5.0, 15.0), (10.0, 20.0), (15.0, 25.0), (20.0, 30.0), (25.0, 30.0)]
    assert S.make_spans_fixed_stride(0.0, 12.0, 10.0, 10.0, 5.0) == [(0.0, 10.0)]  # the 2 s tail is below min_clip_length_s
    assert S.make_spans_fixed_stride(0.0, 12.0, 10.0, 10.0, 2.0) == [(0.0, 10.0), (10.0, 12.0)]
    assert S.make_spans_fixed_stride(0.0, 5.0, 10.0, 10.0, 10.0) == []
    a = S.make_clip_uuids("session", [(0.0, 10.0), (10.0, 20.0)])
    assert a == S.make_clip_uuids("session", [(0.0, 10.0), (10.0, 20.0)]) and a[0] != a[1]
    assert a[0] == uuid.uuid5(uuid.NAMESPACE_URL, "session_0.0_10.0")
    assert S.make_clip_uuids("other", [(0.0, 10.0)]) != a[:1]
    assert S.videos_durations([]) == []
    assert S.videos_durations([_video(n_frames=90, fps=30.0), _video(n_frames=100, fps=0.0)]) == [3.0, -1.0]
    with pytest.raises(ValueError, match="No timestamps"):
        S.validate_video_timestamps([])
    with pytest.raises(ValueError, match="no timestamps"):
        S.validate_video_timestamps([np.zeros(3, np.float32), np.zeros(0, np.float32)])
    S.validate_video_timestamps([np.zeros(3, np.float32)])
    cams = [_video("a.mp4"), _video("b.mp4", with_ts=False), _video("c.mp4", with_ts=False)]
    cams[2].errors["timestamps"] = "demux failed"  # an earlier stage's message is preserved (:1075-1084)
    with pytest.raises(ValueError, match="missing timestamps"):
        S.videos_timestamps(cams)
    assert cams[0].errors == {} and cams[1].errors == {"timestamps": "missing"} and cams[2].errors == {"timestamps": "demux failed"}


def test_stage_default_parameters_and_task_mutations():
    """FixedStrideExtractorStage on SplitPipeTasks (test_fixed_stride_extraction.py:49-145, 245-404, 933-957)."""
    st = FixedStrideExtractorStage()
    assert (st.clip_len_s, st.clip_stride_s, st.min_clip_length_s, st._limit_clips) == (10, 10, 10, 0)
    task = SplitPipeTask(session_id="sess-1", video=_video("s3://b/v.mp4"))
    assert st.process_data([task]) == [task] and not task.errors
    clips = task.video.clips
    assert [c.span for c in clips] == [(0.0, 10.0), (10.0, 20.0), (20.0, 30.0)]
    assert all(c.source_video == "s3://b/v.mp4" and not c.encoded_data for c in clips)
    assert clips[1].uuid == uuid.uuid5(uuid.NAMESPACE_URL, "sess-1_10.0_20.0")
    # limit, overlap, stats
    t2 = SplitPipeTask(session_id="sess-2", video=_video())
    FixedStrideExtractorStage(clip_len_s=10, clip_stride_s=5, min_clip_length_s=5, limit_clips=2, log_stats=True).process_data([t2])
    assert [c.span for c in t2.video.clips] == [(0.0, 10.0), (5.0, 15.0)] and "FixedStrideExtractorStage" in t2.stage_perf
    # too short for one clip: no clips, no error
    t3 = SplitPipeTask(session_id="sess-3", video=_video(n_frames=150))
    FixedStrideExtractorStage().process_data([t3])
    assert t3.video.clips == [] and not t3.errors
    # missing timestamps / incomplete metadata: recorded on the task (and the video), never raised
    t4 = SplitPipeTask(session_id="sess-4", video=_video(with_ts=False))
    FixedStrideExtractorStage().process_data([t4])
    assert "failed to populate clips" in t4.errors["FixedStrideExtractorStage"] and t4.video.errors == {"timestamps": "missing"} and not t4.video.clips
    t5 = SplitPipeTask(session_id="sess-5", video=_video())
    t5.video.metadata.video_codec = None
    FixedStrideExtractorStage().process_data([t5])
    assert "Incomplete metadata" in t5.errors["FixedStrideExtractorStage"] and t5.video.errors == {"metadata": "incomplete"}
    # multi-camera session: every camera receives the same spans and uuids
    t6 = SplitPipeTask(session_id="rig", videos=[_video("cam0.mp4"), _video("cam1.mp4", n_frames=600, fps=24.0)])
    FixedStrideExtractorStage(clip_len_s=10, clip_stride_s=10, min_clip_length_s=5).process_data([t6])
    a, b = t6.videos
    assert [c.span for c in a.clips] == [c.span for c in b.clips] == [(0.0, 10.0), (10.0, 20.0), (20.0, 25.0)]  # the shorter camera bounds the window
    assert [c.uuid for c in a.clips] == [c.uuid for c in b.clips] and {c.source_video for c in b.clips} == {"cam1.mp4"}
    b.clips[1] = Clip(uuid=b.clips[1].uuid, source_video="cam1.mp4", span=(10.0, 19.0))
    with pytest.raises(ValueError, match="misaligned"):
        S.assert_video_clip_alignment(t6.videos)
    b.clips.pop()
    with pytest.raises(ValueError, match="different numbers"):
        S.assert_video_clip_alignment(t6.videos)


def test_video_weight_fraction_and_decoder_heuristics():
    """Video / SplitPipeTask helpers of the reference data model the path's callers read (data_model.py:496-583, 744-790;
    reference test: tests/cosmos_curate/pipelines/video/utils/test_data_model.py:340-380)."""
    v = _video()
    assert v.weight == 0  # size unknown: not downloaded yet
    v.metadata.size = 1 << 20
    assert v.fraction == 1.0 and v.weight == pytest.approx(30.0 / 300)
    v.num_total_clips = 4
    v.clips = [Clip(uuid=uuid.uuid4(), source_video="v.mp4", span=(0.0, 1.0))]
    assert v.fraction == 0.25 and v.weight == pytest.approx(0.1 * 0.25)
    w = _video("w.mp4", n_frames=1800)
    w.metadata.size = 5
    task = SplitPipeTask(session_id="s", videos=[v, w])
    assert task.weight == pytest.approx(v.weight + w.weight)  # multi-camera: the sum over its videos
    assert task.fraction == pytest.approx(1 / 4)
    for codec, pix, want in (("h264", "yuv420p", True), ("h264", "yuv420p10le", True), ("h264", "yuv444p", False), ("hevc", "yuv420p10le", True),
                             ("hevc", "yuv444p", True), ("hevc", "yuv422p", False), ("av1", "yuv420p", False), ("h264", None, False), (None, "yuv420p", False)):
        v.metadata.video_codec, v.metadata.pixel_format = codec, pix
        assert v.nvdec_support() is want, (codec, pix)
    v.metadata.pixel_format = "yuv420p10le"
    assert v.is_10_bit_color() is True
    v.metadata.pixel_format = "yuv420p"
    assert v.is_10_bit_color() is False
    v.metadata.pixel_format = None
    assert v.is_10_bit_color() is None


def test_chunk_tasks_matches_the_reference_chunking():
    """chunk_tasks / slice_video_clips (clip_extraction_stages.py:46-163): chunk sizes equal to the reference's
    split_by_chunk_size on the same spans; every camera cut at the same indices; chunk bookkeeping fields; stats reset."""
    from cosmos_curate_b200.data_model import StagePerfStats

    g = json.loads((GOLDEN / "fixed_stride_ref.json").read_text())
    assert list(S.split_by_chunk_size(range(10), 4)) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    assert list(S.split_by_chunk_size(range(10), 4, drop_incomplete_chunk=True)) == [[0, 1, 2, 3], [4, 5, 6, 7]]
    for case in g["chunks"]:
        spans = [(float.fromhex(a), float.fromhex(b)) for a, b in case["spans"]]
        cams = []
        for name in ("cam0.mp4", "cam1.mp4"):
            v = _video(name)
            v.metadata.size = 1
            v.clips = [Clip(uuid=uuid.uuid5(uuid.NAMESPACE_URL, f"{i}"), source_video=name, span=sp) for i, sp in enumerate(spans)]
            v.errors["note"] = "kept"
            cams.append(v)
        task = SplitPipeTask(session_id="rig", videos=cams, stage_perf={"Up": StagePerfStats(process_time=2.0)})
        subs = S.chunk_tasks([task], case["num_clips_per_chunk"])
        assert [len(t.video.clips) for t in subs] == case["chunk_sizes"]
        pos = 0
        for i, t in enumerate(subs):
            assert t.session_id == "rig" and len(t.videos) == 2
            for v, src in zip(t.videos, cams):
                assert v.clips == src.clips[pos : pos + len(v.clips)]
                assert (v.num_total_clips, v.num_clip_chunks, v.clip_chunk_index) == (len(spans), len(subs), i)
                assert v.metadata is src.metadata and v.clip_stats is src.clip_stats and v.errors == {"note": "kept"} and v.errors is not src.errors
            assert t.stage_perf["Up"].process_time == (2.0 if i == 0 else 0.0)  # carried by the first subtask only
            assert t.fraction == pytest.approx(len(t.video.clips) / len(spans))
            pos += len(t.video.clips)
        assert pos == len(spans)
    v = _video()
    v.clips = [Clip(uuid=uuid.uuid4(), source_video="v.mp4", span=(0.0, 1.0))]
    with pytest.raises(ValueError, match="less than start"):
        S.slice_video_clips(v, 1, 0, 0, 1)
    with pytest.raises(ValueError, match="out of range"):
        S.slice_video_clips(v, 0, 2, 0, 1)


def test_host_side_chain_download_split_cut_on_the_reference_fixture(tmp_path):
    """VideoDownloader -> FixedStrideExtractorStage -> ClipStreamCopyStage(num_clips_per_chunk) on the reference's own media
    fixture, all host code: what reaches the GPU stages is a list of SplitPipeTasks whose clips carry standalone MP4 bytes."""
    import shutil

    from cosmos_curate_b200.interfaces import run_pipeline
    from cosmos_curate_b200.runtime import mp4_index
    from cosmos_curate_b200.stages import ClipStreamCopyStage, VideoDownloader

    src = tmp_path / "sintel.mp4"
    shutil.copy(GOLDEN / "sintel_clip_10s.mp4", src)
    task = SplitPipeTask(session_id=str(src), video=Video(input_video=src))
    missing = SplitPipeTask(session_id="gone", video=Video(input_video=tmp_path / "gone.mp4"))
    remote = SplitPipeTask(session_id="s3", video=Video(input_video="s3://bucket/v.mp4"))
    out = run_pipeline([task, missing, remote], [VideoDownloader(log_stats=True), FixedStrideExtractorStage(clip_len_s=4, clip_stride_s=4, min_clip_length_s=2)])
    assert out is not None and len(out) == 3
    v = task.video
    assert v.metadata.num_frames == 240 and v.metadata.framerate == 24.0 and v.metadata.video_codec == "h264" and v.has_metadata()
    assert v.timestamps.dtype == np.float32 and len(v.timestamps) == 240 and v.timestamps[1] == np.float32(1 / 24)
    assert v.nvdec_support() and v.weight == pytest.approx(10.0 / 300) and "VideoDownloader" in task.stage_perf
    assert [c.span for c in v.clips] == [(0.0, 4.0), (4.0, 8.0), (8.0, 10.0)]
    assert "download" in missing.video.errors and "FixedStrideExtractorStage" in missing.errors and not missing.video.clips
    assert "no storage client" in remote.video.errors["download"]
    cut = ClipStreamCopyStage(num_clips_per_chunk=1).process_data([task])
    assert [len(t.video.clips) for t in cut] == [2, 1]  # 4 s + 4 s of clips close the first chunk
    first = cut[0].video.clips[0]
    idx = mp4_index(first.encoded_data.resolve())
    assert idx["width"] == 854 and idx["n_samples"] >= 96 and first.span[0] == 0.0  # the fixture is one GOP: every cut starts at its only sync sample
    assert cut[1].video.clip_chunk_index == 1 and cut[1].fraction == pytest.approx(1 / 3)
    # ... and out again: the writer stage's layout under a local directory (no GPU stage in between: no scores, no embeddings)
    from cosmos_curate_b200.stages import ClipWriterStage

    ClipWriterStage(str(tmp_path / "out"), str(tmp_path), generate_embeddings=False).process_data(cut)
    assert sorted(p.name for p in (tmp_path / "out" / "processed_clip_chunks").iterdir()) == ["sintel.mp4_0.json", "sintel.mp4_1.json"]
    assert (tmp_path / "out" / "processed_videos" / "sintel.mp4.json").exists() and len(list((tmp_path / "out" / "clips").glob("*.mp4"))) == 3
    back = (tmp_path / "out" / "clips" / f"{first.uuid}.mp4").read_bytes()
    assert mp4_index(np.frombuffer(back, dtype=np.uint8))["n_samples"] == idx["n_samples"]
