"""ClipWriterStage for a local output directory - same name, keyword arguments, directory layout, file contents and task
mutations as the reference stage (cosmos_curate/pipelines/video/read_write/metadata_writer_stage.py:66-1020) for what this path
produces:

    clips/<uuid>[/<relative_path>].mp4, filtered_clips/...      clip bytes                                   (:658-681)
    <stem>_embd/<uuid>.pickle                                   per-clip embedding                            (:769-789)
    <stem>_embd_parquet/<video uuid>_<chunk>.parquet            id + embedding rows of one clip chunk         (:467-485)
    metas/v0/<uuid>.json  or  metas_jsonl/v0/<video uuid>_<chunk>.jsonl   clip metadata                       (:487-507, :893-910)
    processed_videos/<rel>.json, processed_clip_chunks/<rel>_<chunk>.json, video_errors/<rel>_<chunk>.json     (:939-1019)

The primary camera of a multi-camera task writes the uuid-keyed files; every camera writes its own mp4s (:365-372).  Object
stores, Lance datasets, CDS parquet, previews and the cosmos-predict dataset belong to other stages' outputs or to the storage
layer and are refused at construction instead of being silently skipped.
"""

from __future__ import annotations

import json
import pathlib
from typing import Any

import numpy as np

from .. import embedding_io as E
from ..data_model import ClipStats, StageTimer
from ..interfaces import CuratorStage, CuratorStageResource

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


def _out(output_path: str, extra: str) -> str:
    return output_path.rstrip("/") + "/" + extra.strip("/")


class ClipWriterStage(CuratorStage):
    """Stage that writes clips, embeddings and metadata for the split pipeline."""

    def __init__(  # noqa: PLR0913
        self,
        output_path: str,
        input_path: str,
        output_s3_profile_name: str = "default",
        *,
        upload_clips: bool = True,
        upload_clip_info_in_chunks: bool = False,
        upload_clip_info_in_lance: bool = False,
        upload_cds_parquet: bool = False,
        dry_run: bool = False,
        generate_embeddings: bool = True,
        embedding_algorithm: str = "openai",
        embedding_model_version: str = "",
        generate_previews: bool = False,
        caption_models: list[str] | None = None,
        enhanced_caption_models: list[str] | None = None,
        generate_cosmos_predict_dataset: str | None = None,
        verbose: bool = False,
        log_stats: bool = False,
    ) -> None:
        for flag, name in ((upload_clip_info_in_lance, "upload_clip_info_in_lance"), (upload_cds_parquet, "upload_cds_parquet"),
                           (generate_previews, "generate_previews"), (generate_cosmos_predict_dataset not in (None, "disable"), "generate_cosmos_predict_dataset")):  # fmt: skip
            if flag:
                msg = f"{name} is not part of this path's outputs"
                raise NotImplementedError(msg)
        if "://" in output_path:
            msg = f"no storage client for {output_path.split('://', 1)[0]}:// outputs in this build (local paths only)"
            raise NotImplementedError(msg)
        self._timer = StageTimer(self)
        self._input_path = input_path.rstrip("/") + "/"
        self._output_path = output_path.rstrip("/") + "/"
        self._upload_clips, self._dry_run = upload_clips, dry_run
        self._emit_per_clip_metadata = not upload_clip_info_in_chunks
        self._emit_jsonl_metadata = upload_clip_info_in_chunks
        self._generate_embeddings = generate_embeddings
        self._embedding_algorithm, self._embedding_model_version = embedding_algorithm, embedding_model_version
        self._verbose, self._log_stats = verbose, log_stats

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(cpus=0.25)

    # ---- paths (static helpers of the reference class, :166-336) ------------------------------------
    get_output_path_clips = staticmethod(E.get_output_path_clips)
    get_output_path_metas = staticmethod(E.get_output_path_metas)
    get_output_path_embds = staticmethod(E.get_output_path_embds)
    get_output_path_embd_parquets = staticmethod(E.get_output_path_embd_parquets)
    get_video_uuid = staticmethod(E.get_video_uuid)
    get_grouped_clips_uri = staticmethod(E.get_grouped_clips_uri)

    @staticmethod
    def get_output_path_processed_videos(output_path: str) -> str:
        return _out(output_path, "processed_videos")

    @staticmethod
    def get_output_path_processed_clip_chunks(output_path: str) -> str:
        return _out(output_path, "processed_clip_chunks")

    @staticmethod
    def get_output_path_video_errors(output_path: str) -> str:
        return _out(output_path, "video_errors")

    @staticmethod
    def get_output_path_meta_jsonls(output_path: str, version: str) -> str:
        return _out(output_path, f"metas_jsonl/{version}")

    def _rel(self, input_video_path: str) -> str:
        assert input_video_path.startswith(self._input_path), f"{input_video_path} is not under {self._input_path}"
        return input_video_path[len(self._input_path) :]

    def _write(self, dest: pathlib.Path, data: bytes) -> None:
        if self._dry_run:
            return
        dest.parent.mkdir(parents=True, exist_ok=True)
        dest.write_bytes(data)

    def _write_json(self, dest: pathlib.Path, data: dict) -> None:
        self._write(dest, json.dumps(data, indent=4, default=str).encode())

    # ---- per clip -----------------------------------------------------------------------------------
    def _write_clip_mp4(self, clip, relative_path: str, stats: ClipStats, *, filtered: bool = False) -> None:
        data = clip.encoded_data.resolve() if clip.encoded_data else None
        if data is not None:
            name = f"{clip.uuid}/{relative_path}.mp4" if relative_path else f"{clip.uuid}.mp4"
            if self._upload_clips:
                self._write(pathlib.Path(self.get_output_path_clips(self._output_path, filtered=filtered)) / name, np.asarray(data).tobytes())
            stats.num_transcoded += 1
        else:
            logger.warning(f"Clip {clip.uuid} from {clip.source_video} has no buffer, skip uploading to s3")
        if not filtered:
            stats.num_passed += 1

    def _write_clip_embedding(self, clip, stats: ClipStats) -> None:
        emb = E.get_clip_embedding(clip, self._embedding_algorithm)
        if emb is not None:
            if not self._dry_run and self._emit_per_clip_metadata:
                E.write_clip_embedding_pickle(clip, self._output_path, self._embedding_algorithm)
            stats.num_with_embeddings += 1
        elif self._generate_embeddings:
            logger.error(f"Clip {clip.uuid} from {clip.source_video} has no {self._embedding_algorithm} embedding, skip uploading")

    def _clip_metadata(self, clip, video, stats: ClipStats, *, filtered: bool = False) -> dict[str, Any]:
        data = E.make_clip_metadata(clip, video, self._output_path, self._embedding_algorithm, filtered=filtered, embedding_model_version=self._embedding_model_version)
        if self._emit_per_clip_metadata:
            keep = {k: v for k, v in data.items() if k not in ("embedding", "embedding_model_name", "embedding_model_version")}
            self._write_json(E.get_clip_uri(clip.uuid, self.get_output_path_metas(self._output_path, "v0"), "json"), keep)
        duration = clip.span[1] - clip.span[0]
        stats.total_clip_duration += duration
        stats.max_clip_duration = max(stats.max_clip_duration, duration)
        return data

    # ---- per video ----------------------------------------------------------------------------------
    def _write_video_metadata(self, video) -> None:
        path = str(video.input_video)
        rel = self._rel(path)
        if video.errors:
            self._write_json(pathlib.Path(self.get_output_path_video_errors(self._output_path)) / f"{rel}_{video.clip_chunk_index}.json",
                             {"video": path, "clip_chunk_index": video.clip_chunk_index, "errors": video.errors})  # fmt: skip
            return
        m, s = video.metadata, video.clip_stats
        if video.clip_chunk_index == 0:
            self._write_json(pathlib.Path(self.get_output_path_processed_videos(self._output_path)) / f"{rel}.json", {
                "video": path, "height": m.height, "width": m.width, "framerate": m.framerate, "num_frames": m.num_frames, "duration": m.duration,
                "video_codec": m.video_codec, "pixel_format": m.pixel_format, "audio_format": m.audio_codec, "num_total_clips": video.num_total_clips,
                "num_clip_chunks": video.num_clip_chunks, "video_uuid": self.get_video_uuid(path)})  # fmt: skip
        self._write_json(pathlib.Path(self.get_output_path_processed_clip_chunks(self._output_path)) / f"{rel}_{video.clip_chunk_index}.json", {
            "video": path, "clip_chunk_index": video.clip_chunk_index, "num_clips_filtered_by_motion": s.num_filtered_by_motion,
            "num_clips_filtered_by_aesthetic": s.num_filtered_by_aesthetic, "num_clips_filtered_by_qwen_classifier": s.num_filtered_by_qwen_classifier,
            "num_clips_filtered_by_qwen_semantic": s.num_filtered_by_qwen_semantic, "num_clips_filtered_by_artificial_text": s.num_filtered_by_artificial_text,
            "num_clips_passed": s.num_passed, "num_clips_transcoded": s.num_transcoded, "num_clips_with_embeddings": s.num_with_embeddings,
            "num_clips_with_caption": s.num_with_caption, "num_clips_with_webp": s.num_with_webp, "total_clip_duration": s.total_clip_duration,
            "max_clip_duration": s.max_clip_duration, "total_prompt_tokens": s.total_prompt_tokens, "total_output_tokens": s.total_output_tokens,
            "clips": [str(c.uuid) for c in video.clips], "filtered_clips": [str(c.uuid) for c in video.filtered_clips],
            "all_windows": {str(c.uuid): {} for c in video.clips}, "all_windows_enhanced_caption": {str(c.uuid): {} for c in video.clips}})  # fmt: skip

    def _process_video(self, video, *, is_primary: bool) -> None:
        stats = ClipStats()
        rows: list[dict[str, Any]] = []
        with self._timer.time_process(len(video.clips)):
            for clip in video.clips:
                self._write_clip_mp4(clip, video.relative_path, stats)
                if is_primary:
                    self._write_clip_embedding(clip, stats)
                    rows.append(self._clip_metadata(clip, video, stats))
            for clip in video.filtered_clips:
                self._write_clip_mp4(clip, video.relative_path, stats, filtered=True)
                if is_primary:
                    self._clip_metadata(clip, video, stats, filtered=True)
            video.clip_stats.combine(stats)
            self._write_video_metadata(video)
            if is_primary and not self._dry_run:
                E.write_grouped_embeddings_parquet(video, self._output_path, self._embedding_algorithm)
                if self._emit_jsonl_metadata and rows:
                    lines = "\n".join(json.dumps({k: v for k, v in r.items() if k != "embedding"}, default=str) for r in rows) + "\n"
                    self._write(pathlib.Path(self.get_grouped_clips_uri(self.get_video_uuid(str(video.input_video)), video.clip_chunk_index,
                                                                        self.get_output_path_meta_jsonls(self._output_path, "v0"), "jsonl")), lines.encode())  # fmt: skip
        for clip in video.clips + video.filtered_clips:  # clean up intermediate data (:429-451)
            clip.encoded_data.drop()
            clip.intern_video_2_embedding = clip.cosmos_embed1_embedding = clip.openai_embedding = None

    def process_data(self, tasks):
        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            for video_index, video in enumerate(task.videos):
                self._process_video(video, is_primary=video_index == 0)
            if self._log_stats:
                stage_name, stage_perf_stats = self._timer.log_stats()
                task.stage_perf[stage_name] = stage_perf_stats
        return tasks
