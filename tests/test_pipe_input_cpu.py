"""Input listing and the resume rule for local directories (cosmos_curate_b200/pipe_input.py; reference:
pipelines/video/utils/video_pipe_input.py:40-300), exercised against the files ClipWriterStage really writes."""

from __future__ import annotations

import json
import uuid

import pytest

from cosmos_curate_b200 import pipe_input as P
from cosmos_curate_b200.data_model import Clip, SplitPipeTask, Video
from cosmos_curate_b200.stages import ClipWriterStage


def _touch(p, text="x"):
    p.parent.mkdir(parents=True, exist_ok=True)
    p.write_text(text)


def test_resume_skips_only_fully_written_videos(tmp_path):
    inp, out = tmp_path / "in", tmp_path / "out"
    for name in ("a.mp4", "b.mp4", "sub/c.mp4", "d.mp4", "e.mp4"):
        _touch(inp / name)
    assert P.files_relative(inp) == ["a.mp4", "b.mp4", "d.mp4", "e.mp4", "sub/c.mp4"] and P.files_relative(inp, 2) == ["a.mp4", "b.mp4"]
    pv, pc = ClipWriterStage.get_output_path_processed_videos(str(out)), ClipWriterStage.get_output_path_processed_clip_chunks(str(out))
    videos, all_videos, n_done = P.extract_single_cam_split_tasks(str(inp), None, str(out), pv, pc)
    assert [str(v.input_video) for v in videos] == [str(inp / n) for n in all_videos] and n_done == 0 and all(v.relative_path == "" for v in videos)

    # the writer stage produces the summaries: a.mp4 complete (2 chunks), sub/c.mp4 killed after its first chunk of 2, d.mp4 errored
    def chunk(name, k, n, errors=None):
        v = Video(input_video=inp / name, clips=[Clip(uuid=uuid.uuid4(), source_video=str(inp / name), span=(0.0, 1.0))], clip_chunk_index=k, num_clip_chunks=n, num_total_clips=n)
        v.errors.update(errors or {})
        return SplitPipeTask(session_id=name, video=v)

    st = ClipWriterStage(str(out), str(inp), generate_embeddings=False)
    st.process_data([chunk("a.mp4", 0, 2), chunk("a.mp4", 1, 2), chunk("sub/c.mp4", 0, 2), chunk("d.mp4", 0, 1, {"download": "boom"})])
    assert P.find_fully_processed_videos(pv, pc) == {"a.mp4"}
    videos, all_videos, n_done = P.extract_single_cam_split_tasks(str(inp), None, str(out), pv, pc, verbose=True)
    assert n_done == 1 and [v.input_video.name for v in videos] == ["b.mp4", "d.mp4", "e.mp4", "c.mp4"] and len(all_videos) == 5
    # the limit counts NEW videos: listing is widened by the number already processed (video_pipe_input.py:155-157)
    videos, all_videos, _ = P.extract_single_cam_split_tasks(str(inp), None, str(out), pv, pc, limit=2)
    assert [v.input_video.name for v in videos] == ["b.mp4", "d.mp4"] and all_videos == ["a.mp4", "b.mp4", "d.mp4"]
    st.process_data([chunk("sub/c.mp4", 1, 2)])  # the missing chunk arrives: now complete
    assert P.find_fully_processed_videos(pv, pc) == {"a.mp4", "sub/c.mp4"}
    _touch(out / "processed_videos" / "broken.mp4.json", "{not json")
    assert P.find_fully_processed_videos(pv, pc) == {"a.mp4", "sub/c.mp4"}  # unreadable summary: redo that video
    # explicit list
    lst = tmp_path / "list.json"
    lst.write_text(json.dumps([str(inp / "e.mp4"), str(inp / "a.mp4")]))
    videos, all_videos, _ = P.extract_single_cam_split_tasks(str(inp), str(lst), str(out), pv, pc)
    assert all_videos == ["e.mp4", "a.mp4"] and [v.input_video.name for v in videos] == ["e.mp4"]
    lst.write_text(json.dumps(["/elsewhere/x.mp4"]))
    with pytest.raises(ValueError, match="is not in"):
        P.extract_single_cam_split_tasks(str(inp), str(lst), str(out), pv, pc)


def test_multi_camera_sessions(tmp_path):
    s1, s2 = str(uuid.UUID(int=1)), str(uuid.UUID(int=2))
    for p in (f"{s1}/rec0/front_wide.mp4", f"{s1}/rec0/front_tele.mp4", f"{s1}/rec1/rear.mp4", f"{s1}/notes.txt", f"{s2}/front_wide.mp4", "not-a-uuid/front_wide.mp4", f"{uuid.UUID(int=3)}/readme.txt"):
        _touch(tmp_path / p)
    tasks = P.extract_multi_cam_split_tasks(str(tmp_path), "front_wide", {".mp4"})
    assert [t.session_id for t in tasks] == [s1, s2]
    assert [v.relative_path for v in tasks[0].videos] == ["rec0/front_wide", "rec0/front_tele", "rec1/rear"]  # primary first, the rest sorted
    assert str(tasks[0].videos[0].input_video) == str(tmp_path / s1 / "rec0" / "front_wide.mp4")
    assert len(P.extract_multi_cam_split_tasks(str(tmp_path), "front_wide", {".mp4"}, limit=1)) == 1
    assert P.order_video_paths(["a.txt"], {".mp4"}, "front") == []
    with pytest.raises(ValueError, match="Multiple primary"):
        P.order_video_paths(["front_a.mp4", "front_b.mp4"], {".mp4"}, "front")
    with pytest.raises(ValueError, match="No primary camera"):
        P.order_video_paths(["rear.mp4"], {".mp4"}, "front")


def test_split_summary_aggregates_the_writer_stage_outputs(tmp_path):
    """summary.json (summary_writers.py:127-262) from the files ClipWriterStage wrote: totals, per-video records, unprocessed inputs."""
    inp, out = tmp_path / "in", tmp_path / "out"

    def chunk(name, k, n, n_pass, n_filt, dur):
        mk = lambda i: Clip(uuid=uuid.uuid5(uuid.NAMESPACE_URL, f"{name}{k}{i}"), source_video=name, span=(0.0, dur))  # noqa: E731
        v = Video(input_video=inp / name, clips=[mk(i) for i in range(n_pass)], filtered_clips=[mk(9 + i) for i in range(n_filt)], clip_chunk_index=k, num_clip_chunks=n,
                  num_total_clips=7)
        v.metadata.duration = 100.0
        v.clip_stats.num_filtered_by_aesthetic = n_filt
        return SplitPipeTask(session_id=name, video=v)

    ClipWriterStage(str(out), str(inp), generate_embeddings=False).process_data([chunk("a.mp4", 0, 2, 3, 1, 4.0), chunk("a.mp4", 1, 2, 2, 1, 6.0), chunk("b.mp4", 0, 1, 1, 0, 2.5)])
    s = P.write_split_summary(str(inp), ["a.mp4", "b.mp4", "c.mp4"], 3, str(out), embedding_algorithm="openai", limit=5, pipeline_run_time=1.5)
    assert json.loads((out / "summary.json").read_text()) == s
    assert (s["num_input_videos"], s["num_processed_videos"], s["total_num_clips_passed"], s["total_num_clips_filtered_by_aesthetic"]) == (3, 2, 6, 2)
    assert s["total_video_duration"] == 200.0 and s["max_clip_duration"] == 6.0 and s["total_clip_duration"] == pytest.approx(4 * 4.0 + 3 * 6.0 + 2.5)
    assert s["a.mp4"]["num_clip_chunks"] == 2 and len(s["a.mp4"]["clips"]) == 5 and len(s["a.mp4"]["filtered_clips"]) == 2 and s["a.mp4"]["num_total_clips"] == 7
    assert s["a.mp4"]["video_uuid"] == str(uuid.uuid5(uuid.NAMESPACE_URL, str(inp / "a.mp4"))) and s["c.mp4"] == {"source_video": str(inp / "c.mp4"), "processed": False}
    assert s["embedding_algorithm"] == "openai" and s["pipeline_run_time"] == 1.5
