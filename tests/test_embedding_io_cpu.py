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
        assert ("embedding_model_name" in meta) == (not filtered)


def test_reader_error_convention(tmp_path):
    with pytest.raises(ValueError, match="no rows to cluster"):
        E.read_embedding_parquets([])
    p = tmp_path / "ragged.parquet"
    pd.DataFrame([{"id": "a", "embedding": [1.0, 2.0]}, {"id": "b", "embedding": [1.0]}]).to_parquet(p, index=False)
    with pytest.raises(ValueError, match=r"ragged embeddings \(min=1, max=2\); SemDeDup requires fixed-length vectors"):
        E.read_embedding_parquets([p])
    t = _task(n_pass=0, n_filtered=1)
    assert E.write_grouped_embeddings_parquet(t.video, str(tmp_path)) is None  # empty buffer: nothing written (:468)
