"""Which videos become tasks, and which are already done: the input side of the split pipeline for local directories, with the
resume rule of the reference (cosmos_curate/pipelines/video/utils/video_pipe_input.py:40-300).  A video counts as processed when
`processed_videos/<rel>.json` exists AND every `processed_clip_chunks/<rel>_<k>.json` for k < its `num_clip_chunks` exists - the
files ClipWriterStage writes last - so a run killed between two clip chunks of a video redoes that video and nothing else.

    files_relative                  sorted recursive listing, paths relative to the root (storage_utils.py:407-421)
    find_fully_processed_videos     the resume check (:45-83, :122-150)
    extract_single_cam_split_tasks  (videos to process, all input videos, number already processed) (:111-190); the limit
                                    applies to the NEW videos, like the reference's
    order_video_paths               primary camera first, the rest sorted (:193-211)
    extract_multi_cam_split_tasks   one task per UUID-named session directory (:214-281)
    write_split_summary             `summary.json`: totals + one record per input video aggregated over its clip-chunk summaries
                                    (pipelines/video/read_write/summary_writers.py:127-262; captions / perf stats are other
                                    stages' business)

Object stores are out of this path's scope: every path is a local directory.
"""

from __future__ import annotations

import json
import pathlib
import uuid

from .data_model import SplitPipeTask, Video

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


def files_relative(path: str | pathlib.Path, limit: int = 0) -> list[str]:
    root = pathlib.Path(path)
    files = sorted(str(x.relative_to(root)) for x in root.rglob("*") if x.is_file())
    return files[:limit] if limit > 0 else files


def find_fully_processed_videos(output_video_path: str, output_clip_chunk_path: str) -> set[str]:
    """Relative input paths whose video summary and all clip-chunk summaries exist (unreadable summaries count as not processed)."""
    done = set()
    for name in files_relative(output_video_path):
        try:
            n_chunks = int(json.loads((pathlib.Path(output_video_path) / name).read_text())["num_clip_chunks"])
            stem = name.removesuffix(".json")
            if all((pathlib.Path(output_clip_chunk_path) / f"{stem}_{k}.json").exists() for k in range(n_chunks)):
                done.add(stem)
            else:
                logger.debug(f"Semi-processed video {name} is missing a clip chunk")
        except Exception as e:  # noqa: BLE001
            logger.error(f"Failed to read processed video json {name}: {e}")
    return done


def _read_video_list_json(input_path: str, list_path: str) -> list[str]:
    prefix = input_path.rstrip("/") + "/"
    out = []
    for video_path in [str(x) for x in json.loads(pathlib.Path(list_path).read_text())]:
        if not video_path.startswith(prefix):
            error_msg = f"Input video {video_path} is not in {prefix}"
            raise ValueError(error_msg)
        out.append(video_path[len(prefix) :])
    return out


def extract_single_cam_split_tasks(input_path: str, input_video_list_json_path: str | None, output_path: str, output_video_path: str,
                                   output_clip_chunk_path: str, limit: int = 0, *, verbose: bool = False) -> tuple[list[Video], list[str], int]:  # fmt: skip
    if (pathlib.Path(output_path) / "summary.json").exists():
        logger.warning(f"Output path {output_path} already concluded with a summary.json file")
    processed = find_fully_processed_videos(output_video_path, output_clip_chunk_path)
    logger.info(f"Found {len(processed)} fully processed videos in {output_video_path}")
    if input_video_list_json_path is not None:
        all_videos = _read_video_list_json(input_path, input_video_list_json_path)
    else:
        all_videos = files_relative(input_path, 0 if limit == 0 else len(processed) + limit)
    if verbose:
        logger.info(f"Skipping {sum(x in processed for x in all_videos)} already-processed video(s)")
    raw = [x for x in all_videos if x not in processed]
    if limit > 0:
        raw = raw[:limit]
    root = pathlib.Path(input_path)
    return [Video(root / x, relative_path="") for x in raw], all_videos, len(processed)  # single camera: clips go to clips/<uuid>.mp4


def order_video_paths(paths: list[str], video_extensions: set[str], primary_camera_keyword: str) -> list[str]:
    video_paths = sorted(p for p in paths if any(p.endswith(ext) for ext in video_extensions))
    if not video_paths:
        return []
    primary = [p for p in video_paths if primary_camera_keyword in p]
    if len(primary) > 1:
        msg = f"Multiple primary cameras found: {primary=}, need distinct primary camera to run multicam pipeline"
        raise ValueError(msg)
    if len(primary) == 0:
        msg = f"No primary camera found with keyword {primary_camera_keyword}, need distinct primary camera to run multicam pipeline"
        raise ValueError(msg)
    return primary + sorted(p for p in video_paths if p not in primary)


def _is_uuid(value: str) -> bool:
    try:
        uuid.UUID(value)
    except (ValueError, AttributeError, TypeError):
        return False
    return True


def extract_multi_cam_split_tasks(sessions_prefix: str, primary_camera_keyword: str, video_extensions: set[str], limit: int = 0) -> list[SplitPipeTask]:
    root = pathlib.Path(sessions_prefix)
    tasks: list[SplitPipeTask] = []
    for session_id in sorted({f.split("/")[0] for f in files_relative(root)}):
        if not _is_uuid(session_id):
            continue
        paths = order_video_paths(files_relative(root / session_id), video_extensions, primary_camera_keyword)
        if paths:
            tasks.append(SplitPipeTask(session_id=session_id, videos=[Video(root / session_id / p, relative_path=str(pathlib.Path(p).with_suffix(""))) for p in paths]))
        if limit > 0 and len(tasks) >= limit:
            break
    logger.info(f"Extracted {len(tasks)} session tasks from {sessions_prefix}")
    return tasks


CLIP_STATS_KEYS = ("num_clips_filtered_by_motion", "num_clips_filtered_by_aesthetic", "num_clips_filtered_by_qwen_classifier",
                   "num_clips_filtered_by_qwen_semantic", "num_clips_filtered_by_artificial_text", "num_clips_passed", "num_clips_transcoded",
                   "num_clips_with_embeddings", "num_clips_with_caption", "num_clips_with_webp")  # fmt: skip


def write_split_summary(input_path: str, input_videos_relative: list[str], num_input_videos_selected: int, output_path: str, *,
                        embedding_algorithm: str, limit: int = 0, pipeline_run_time: float = 0.0, video_bytes: int = 0, num_remuxed_videos: int = 0,
                        multi_cam: bool = False) -> dict:  # fmt: skip
    """Aggregates processed_videos/<rel>.json and processed_clip_chunks/<rel>_<k>.json into <output_path>/summary.json; returns it."""
    out = pathlib.Path(output_path)
    pv, pc = out / "processed_videos", out / "processed_clip_chunks"
    listed = files_relative(pv)
    sessions = sorted({f.split("/")[0] for f in listed}) if multi_cam else listed
    summary: dict = {"num_input_videos": len(input_videos_relative), "num_input_videos_selected": num_input_videos_selected,
                     "num_processed_videos": len(sessions), "embedding_algorithm": embedding_algorithm, "total_video_duration": 0, "total_clip_duration": 0,
                     "max_clip_duration": 0, "pipeline_run_time": pipeline_run_time, "total_video_bytes": video_bytes, "num_remuxed_videos": num_remuxed_videos,
                     "total_prompt_tokens": 0, "total_output_tokens": 0}  # fmt: skip
    for key in CLIP_STATS_KEYS:
        summary[f"total_{key}"] = 0
    for video in input_videos_relative:
        rec: dict = {"source_video": str(pathlib.Path(input_path) / video)}
        summary[video] = rec
        meta_path = pv / f"{video}.json"
        if not meta_path.exists():
            if limit == 0:
                logger.error(f"video process-record {video} not found ???")
            rec["processed"] = False
            continue
        meta = json.loads(meta_path.read_text())
        chunks = []
        for k in range(meta.get("num_clip_chunks", 0)):
            cp = pc / f"{video}_{k}.json"
            if cp.exists():
                chunks.append(json.loads(cp.read_text()))
            else:
                logger.error(f"clip chunk record {cp} not found ???")
        rec.update({"video_uuid": meta.get("video_uuid", "N/A"), "num_clip_chunks": len(chunks), "num_total_clips": meta.get("num_total_clips", "N/A"),
                    "clips": [], "filtered_clips": []})  # fmt: skip
        for key in CLIP_STATS_KEYS:
            rec[key] = 0
        for ch in chunks:
            for key in CLIP_STATS_KEYS:
                rec[key] += ch.get(key, 0)
                summary[f"total_{key}"] += ch.get(key, 0)
            rec["clips"].extend(ch.get("clips", []))
            rec["filtered_clips"].extend(ch.get("filtered_clips", []))
            summary["total_clip_duration"] += ch.get("total_clip_duration", 0)
            summary["max_clip_duration"] = max(summary["max_clip_duration"], ch.get("max_clip_duration", 0))
            summary["total_prompt_tokens"] += ch.get("total_prompt_tokens", 0)
            summary["total_output_tokens"] += ch.get("total_output_tokens", 0)
        summary["total_video_duration"] += meta.get("duration", 0) or 0
    out.mkdir(parents=True, exist_ok=True)
    (out / "summary.json").write_text(json.dumps(summary, indent=4))
    return summary
