"""CPU: the on-disk formats either side of the embedding path - layout, schema and loader semantics of
ClipWriterStage (metadata_writer_stage.py:226-265, 467-485, 752-781, 796-905) and SemanticDedupActor's reader
(dedup_actor.py:199-222)."""

from __future__ import annotations

import json
import pickle
import uuid

import numpy as np
import pandas as pd
import pytest

from cosmos_curate_b200 import embedding_io as E
from cosmos_curate_b200.data_model import Clip, SplitPipeTask, Video


def _task(n_pass=3, n_filtered=2, d=768, name="s3://bucket/videos/a.mp4", chunk=2):
    rng = np.random.default_rng(1)
    mk = lambda i: Clip(uuid=uuid.uuid5(uuid.NAMESPACE_URL, f"{name}_{i}"), source_video=name, span=(float(i), float(i + 5)), encoded_data=b"x")  # noqa: E731
    clips = [mk(i) for i in range(n_pass + n_filtered)]
    for i, c in enumerate(clips):
        c.aesthetic_score = 4.0 + i
        if i < n_pass:
            e = rng.standard_normal(d).astype(np.float32)
            c.openai_embedding = e / np.linalg.norm(e)
    v = Video(input_video=name, clips=clips[:n_pass], filtered_clips=clips[n_pass:], clip_chunk_index=chunk)
    v.metadata.width, v.metadata.height, v.metadata.framerate = 1920, 1080, 30.0
    return SplitPipeTask(session_id="s", video=v)


def test_stems_and_paths_follow_the_reference_layout():
    assert E.embd_stem("internvideo2") == "iv2_embd" and E.embd_stem("openai") == "openai_embd"
    assert E.embd_stem("cosmos-embed1-224p") == "ce1_embd_224p" and E.embd_stem("foo") == "foo_embd"
    assert E.get_output_path_embds("/out/", "openai") == "/out/openai_embd"
    assert E.get_output_path_embd_parquets("/out", "internvideo2") == "/out/iv2_embd_parquet"
    assert E.get_output_path_metas("/out") == "/out/metas/v0"
    vid = E.get_video_uuid("s3://bucket/videos/a.mp4")
    assert vid == uuid.uuid5(uuid.NAMESPACE_URL, "s3://bucket/videos/a.mp4")
    assert str(E.get_grouped_clips_uri(vid, 3, "/out/openai_embd_parquet", "parquet")) == f"/out/openai_embd_parquet/{vid}_3.parquet"


def test_writer_outputs_and_dedup_reader_round_trip(tmp_path):
    task = _task()
    stats = E.write_task_outputs(task, str(tmp_path), "openai")
    assert stats == {"num_with_embeddings": 3, "parquets": 1, "metas": 5}
    v = task.video
    # per-clip pickle: the numpy array itself
    for c in v.clips:
        with open(tmp_path / "openai_embd" / f"{c.uuid}.pickle", "rb") as f:
            back = pickle.load(f)  # noqa: S301
        assert isinstance(back, np.ndarray) and back.dtype == np.float32 and np.array_equal(back, c.openai_embedding)
    # grouped parquet: id + embedding, index not stored, one file per (video, chunk)
    pq_path = tmp_path / "openai_embd_parquet" / f"{E.get_video_uuid(str(v.input_video))}_2.parquet"
    df = pd.read_parquet(pq_path)
    assert list(df.columns) == ["id", "embedding"] and len(df) == 3
    assert df["id"].tolist() == [str(c.uuid) for c in v.clips]
    import pyarrow.parquet as pq

    schema = pq.read_schema(pq_path)
    assert str(schema.field("id").type) in ("string", "large_string") and str(schema.field("embedding").type).startswith("list<")
    # the dedup loader: float32 (n, d), ids aligned, values exactly the float32 embeddings (float -> double -> float is lossless)
    ids, x = E.read_embedding_parquets([pq_path])
    assert x.dtype == np.float32 and x.shape == (3, 768) and x.flags["C_CONTIGUOUS"]
    assert np.array_equal(x, np.stack([c.openai_embedding for c in v.clips])) and ids.tolist() == df["id"].tolist()
    # per-clip json: aesthetic_score present, vector absent, filtered clips point at filtered_clips/
    for c, filtered in [(v.clips[0], False), (v.filtered_clips[0], True)]:
        meta = json.loads((tmp_path / "metas" / "v0" / f"{c.uuid}.json").read_text())
        assert meta["span_uuid"] == str(c.uuid) and meta["aesthetic_score"] == c.aesthetic_score and "embedding" not in meta
        assert meta["duration_span"] == list(c.span) and meta["width_source"] == 1920 and meta["framerate_source"] == 30.0
        assert ("filtered_clips" in meta["clip_location"]) == filtered
        assert "embedding_model_name" not in meta and "embedding_model_version" not in meta  # only the grouped rows carry them (:896-905)
        assert "extract_metadata" in meta["errors"]  # the fake clip bytes are not an MP4: recorded like the reference does (:811-815)


def test_reader_error_convention(tmp_path):
    with pytest.raises(ValueError, match="no rows to cluster"):
        E.read_embedding_parquets([])
    p = tmp_path / "ragged.parquet"
    pd.DataFrame([{"id": "a", "embedding": [1.0, 2.0]}, {"id": "b", "embedding": [1.0]}]).to_parquet(p, index=False)
    with pytest.raises(ValueError, match=r"ragged embeddings \(min=1, max=2\); SemDeDup requires fixed-length vectors"):
        E.read_embedding_parquets([p])
    t = _task(n_pass=0, n_filtered=1)
    assert E.write_grouped_embeddings_parquet(t.video, str(tmp_path)) is None  # empty buffer: nothing written (:468)


def test_clip_writer_stage_layout_stats_and_cleanup(tmp_path):
    """ClipWriterStage on a local directory: the reference's file layout and json contents (metadata_writer_stage.py:365-452,
    :658-681, :769-789, :893-1019) for a two-camera, two-chunk session; stats, cleanup, error routing, jsonl mode, dry run."""
    from conftest import GOLDEN
    from cosmos_curate_b200.stages import ClipWriterStage
    from cosmos_curate_b200.stages.clip_stream_copy import mp4_cut

    src = np.fromfile(GOLDEN / "sintel_clip_10s.mp4", dtype=np.uint8)
    clip_bytes = mp4_cut(src, 0, 48)  # 2 s of the fixture
    in_dir, out_dir = "/data/in/", str(tmp_path / "out")

    def video(name, rel, chunk, n_pass=2, n_filt=1):
        rng = np.random.default_rng(chunk)
        clips = []
        for i in range(n_pass + n_filt):
            c = Clip(uuid=uuid.uuid5(uuid.NAMESPACE_URL, f"sess_{chunk}_{i}"), source_video=name, span=(float(i), float(i + 2 + i)), encoded_data=clip_bytes)
            c.aesthetic_score = 5.0 - i
            if i < n_pass:
                e = rng.standard_normal(16).astype(np.float32)
                c.openai_embedding = e / np.linalg.norm(e)
            clips.append(c)
        import pathlib

        v = Video(input_video=pathlib.Path(name), relative_path=rel, clips=clips[:n_pass], filtered_clips=clips[n_pass:], clip_chunk_index=chunk,
                  num_total_clips=6, num_clip_chunks=2)
        m = v.metadata
        m.width, m.height, m.framerate, m.num_frames, m.duration, m.video_codec, m.pixel_format, m.audio_codec = 854, 480, 24.0, 240, 10.0, "h264", "yuv420p", None
        v.clip_stats.num_filtered_by_aesthetic = n_filt
        return v

    stage = ClipWriterStage(out_dir, in_dir, "default", upload_clips=True, upload_clip_info_in_chunks=False, upload_clip_info_in_lance=False,
                            upload_cds_parquet=False, dry_run=False, generate_embeddings=True, embedding_algorithm="openai", embedding_model_version="r2",
                            generate_previews=False, log_stats=True)
    tasks = [SplitPipeTask(session_id="sess", videos=[video("/data/in/rig/cam0.mp4", "cam0", k), video("/data/in/rig/cam1.mp4", "cam1", k)]) for k in (0, 1)]
    embeddings = {str(c.uuid): c.openai_embedding.copy() for t in tasks for c in t.videos[0].clips}
    assert stage.process_data(tasks) == tasks and "ClipWriterStage" in tasks[0].stage_perf
    out = tmp_path / "out"
    for t in tasks:
        for cam, v in enumerate(t.videos):
            for c in v.clips:
                assert (out / "clips" / str(c.uuid) / f"cam{cam}.mp4").read_bytes() == bytes(clip_bytes)  # every camera writes its own mp4s
            for c in v.filtered_clips:
                assert (out / "filtered_clips" / str(c.uuid) / f"cam{cam}.mp4").exists()
            assert all(not c.encoded_data and c.openai_embedding is None for c in v.clips + v.filtered_clips)  # cleaned up
            s = v.clip_stats
            assert (s.num_passed, s.num_transcoded, s.num_filtered_by_aesthetic) == (2, 3, 1)
            assert s.num_with_embeddings == (2 if cam == 0 else 0) and s.max_clip_duration == (4.0 if cam == 0 else 0.0)
    # uuid-keyed files come from the primary camera only
    assert len(list((out / "openai_embd").glob("*.pickle"))) == 4 and len(list((out / "metas" / "v0").glob("*.json"))) == 6
    some = tasks[0].videos[0].clips[1]
    with open(out / "openai_embd" / f"{some.uuid}.pickle", "rb") as f:
        assert np.array_equal(pickle.load(f), embeddings[str(some.uuid)])  # noqa: S301
    meta = json.loads((out / "metas" / "v0" / f"{some.uuid}.json").read_text())
    assert meta["width"] == 854 and meta["num_frames"] == 48 and meta["video_codec"] == "h264" and meta["num_bytes"] == len(clip_bytes)  # the clip's own facts
    assert meta["source_video"] == "/data/in/rig/cam0.mp4" and meta["clip_location"].endswith(f"out/clips/{some.uuid}.mp4") and "embedding" not in meta
    vid = E.get_video_uuid("/data/in/rig/cam0.mp4")
    ids, x = E.read_embedding_parquets(sorted((out / "openai_embd_parquet").glob(f"{vid}_*.parquet")))
    assert len(ids) == 4 and all(np.array_equal(x[i], embeddings[ids[i]]) for i in range(4))
    pv = json.loads((out / "processed_videos" / "rig" / "cam1.mp4.json").read_text())  # first chunk only
    assert pv["num_total_clips"] == 6 and pv["num_clip_chunks"] == 2 and pv["video_uuid"] == str(E.get_video_uuid("/data/in/rig/cam1.mp4")) and pv["audio_format"] is None
    ck = json.loads((out / "processed_clip_chunks" / "rig" / "cam0.mp4_1.json").read_text())
    assert ck["clip_chunk_index"] == 1 and ck["num_clips_passed"] == 2 and ck["num_clips_with_embeddings"] == 2 and ck["num_clips_filtered_by_aesthetic"] == 1
    assert ck["clips"] == [str(c.uuid) for c in tasks[1].videos[0].clips] and set(ck["all_windows"]) == set(ck["clips"])
    assert not (out / "processed_videos" / "rig" / "cam0.mp4_1.json").exists() and not (out / "video_errors").exists()
    # a video with errors writes video_errors/ instead of the chunk summary; jsonl mode groups the clip metadata; dry run writes nothing
    out2 = tmp_path / "out2"
    st2 = ClipWriterStage(str(out2), in_dir, upload_clip_info_in_chunks=True, embedding_model_version="r2")
    bad = video("/data/in/bad.mp4", "", 0)
    bad.errors["download"] = "boom"
    good = video("/data/in/good.mp4", "", 0)
    st2.process_data([SplitPipeTask(session_id="a", video=bad), SplitPipeTask(session_id="b", video=good)])
    assert json.loads((out2 / "video_errors" / "bad.mp4_0.json").read_text())["errors"] == {"download": "boom"}
    assert not (out2 / "processed_clip_chunks" / "bad.mp4_0.json").exists() and not (out2 / "metas").exists() and not (out2 / "openai_embd").exists()
    rows = [json.loads(ln) for ln in (out2 / "metas_jsonl" / "v0" / f"{E.get_video_uuid('/data/in/good.mp4')}_0.jsonl").read_text().splitlines()]
    assert len(rows) == 2 and rows[0]["embedding_model_name"] == "openai" and rows[0]["embedding_model_version"] == "r2" and "embedding" not in rows[0]
    assert (out2 / "clips" / f"{good.clips[0].uuid}.mp4").exists()  # no relative_path: flat layout
    out3 = tmp_path / "out3"
    dry = video("/data/in/dry.mp4", "", 0)
    ClipWriterStage(str(out3), in_dir, dry_run=True).process_data([SplitPipeTask(session_id="d", video=dry)])
    assert not out3.exists() and dry.clip_stats.num_transcoded == 3
    for kw in ({"upload_clip_info_in_lance": True}, {"upload_cds_parquet": True}, {"generate_previews": True}, {"generate_cosmos_predict_dataset": "predict2"}):
        with pytest.raises(NotImplementedError):
            ClipWriterStage(str(out3), in_dir, **kw)
    with pytest.raises(NotImplementedError, match="storage client"):
        ClipWriterStage("s3://bucket/out", in_dir)
