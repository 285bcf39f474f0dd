"""On-disk contract either side of the embedding path: what `ClipWriterStage` persists of the fields this
path fills, and what the dedup actors read back - restated so a run of the B200 stages can hand its results to the
reference's own dedup pipeline (or to `cosmos_curate_b200.dedup`) through files, layout and schema unchanged.

    <out>/<stem>/<clip-uuid>.pickle                    pickle of the float32 embedding (metadata_writer_stage.py:769-781)
    <out>/<stem>_parquet/<video-uuid>_<chunk>.parquet   columns id (str), embedding (list of floats), one row per clip
                                                       (:752-765 buffer, :467-485 write; pandas -> pyarrow, index=False)
    <out>/metas/v0/<clip-uuid>.json                    per-clip metadata incl. aesthetic_score (:796-886, :896-905)

with <stem> = iv2_embd | ce1_embd_<variant> | openai_embd (:226-236), video uuid = uuid5(NAMESPACE_URL, input path) (:294-296).
`read_embedding_parquets` is the loader of SemanticDedupActor.kmeans / dedup (dedup_actor.py:199-222: columns id + embedding,
fixed-length check with the same error text, float32 (n, d) matrix) on pyarrow instead of cudf.

Host-side file I/O only: no GPU work, nothing here is on the timed path.
"""

from __future__ import annotations

import io
import json
import pathlib
import pickle
import uuid
from typing import Any

import numpy as np

EMBEDDING_FIELDS = {"internvideo2": "intern_video_2_embedding", "openai": "openai_embedding"}


def embd_stem(embedding_algorithm: str) -> str:
    if embedding_algorithm == "internvideo2":
        return "iv2_embd"
    if embedding_algorithm.startswith("cosmos-embed1-"):
        return f"ce1_embd_{embedding_algorithm.removeprefix('cosmos-embed1-')}"
    if embedding_algorithm == "openai":
        return "openai_embd"
    return f"{embedding_algorithm}_embd"


def _join(output_path: str, extra: str) -> str:
    return output_path.rstrip("/") + "/" + extra.strip("/")


def get_output_path_embds(output_path: str, embedding_algorithm: str) -> str:
    return _join(output_path, embd_stem(embedding_algorithm))


def get_output_path_embd_parquets(output_path: str, embedding_algorithm: str) -> str:
    return _join(output_path, f"{embd_stem(embedding_algorithm)}_parquet")


def get_output_path_metas(output_path: str, version: str = "v0") -> str:
    return _join(output_path, f"metas/{version}")


def get_output_path_clips(output_path: str, *, filtered: bool = False) -> str:
    return _join(output_path, "filtered_clips" if filtered else "clips")


def get_video_uuid(input_video_path: str) -> uuid.UUID:
    return uuid.uuid5(uuid.NAMESPACE_URL, f"{input_video_path}")


def get_clip_uri(clip_uuid, path_prefix: str, file_type: str) -> pathlib.Path:
    return pathlib.Path(path_prefix) / f"{clip_uuid}.{file_type}"


def get_grouped_clips_uri(video_uuid, chunk_index: int, path_prefix: str, file_type: str) -> pathlib.Path:
    return pathlib.Path(path_prefix) / f"{video_uuid}_{chunk_index}.{file_type}"


def get_clip_embedding(clip, embedding_algorithm: str):
    if embedding_algorithm.startswith("cosmos-embed1-"):
        return clip.cosmos_embed1_embedding
    field = EMBEDDING_FIELDS.get(embedding_algorithm)
    return getattr(clip, field) if field else None


def _write(dest: pathlib.Path, data: bytes) -> pathlib.Path:
    dest.parent.mkdir(parents=True, exist_ok=True)
    dest.write_bytes(data)
    return dest


def write_clip_embedding_pickle(clip, output_path: str, embedding_algorithm: str = "openai") -> pathlib.Path | None:
    emb = get_clip_embedding(clip, embedding_algorithm)
    if emb is None:
        return None
    buf = io.BytesIO()
    pickle.dump(emb, buf)
    return _write(get_clip_uri(clip.uuid, get_output_path_embds(output_path, embedding_algorithm), "pickle"), buf.getvalue())


def embedding_rows(clips, embedding_algorithm: str = "openai") -> list[dict[str, Any]]:
    rows = []
    for clip in clips:
        emb = get_clip_embedding(clip, embedding_algorithm)
        if emb is not None:
            rows.append({"id": str(clip.uuid), "embedding": np.asarray(emb).reshape(-1).tolist()})
    return rows


def write_grouped_embeddings_parquet(video, output_path: str, embedding_algorithm: str = "openai", clips=None) -> pathlib.Path | None:
    """One parquet per (video, clip chunk): the file set the dedup pipeline lists as its input."""
    import pandas as pd

    rows = embedding_rows(video.clips if clips is None else clips, embedding_algorithm)
    if not rows:
        return None
    buf = io.BytesIO()
    pd.DataFrame(rows).to_parquet(buf, index=False)
    dest = get_grouped_clips_uri(get_video_uuid(str(video.input_video)), video.clip_chunk_index, get_output_path_embd_parquets(output_path, embedding_algorithm), "parquet")
    return _write(dest, buf.getvalue())


def extract_clip_metadata(clip) -> dict[str, Any] | None:
    """width / height / framerate / num_frames / video_codec / num_bytes of the clip's encoded_data, None without bytes."""
    data = clip.encoded_data.resolve() if clip.encoded_data else None
    if data is None:
        return None
    from .runtime import mp4_index
    from .sampling import video_metadata_from_index

    m = video_metadata_from_index(mp4_index(data))
    return {"width": m.width, "height": m.height, "framerate": m.fps, "num_frames": m.num_frames, "video_codec": m.video_codec,
            "num_bytes": int(np.asarray(data).nbytes)}  # fmt: skip


def make_clip_metadata(clip, video, output_path: str, embedding_algorithm: str = "openai", *, filtered: bool = False, embedding_model_version: str = "") -> dict[str, Any]:
    """The fields of _make_clip_metadata (:796-886) this path owns; windows / captions are other stages' business."""
    m = video.metadata
    data: dict[str, Any] = {
        "span_uuid": str(clip.uuid), "source_video": str(clip.source_video), "duration_span": list(clip.span),
        "width_source": m.width, "height_source": m.height, "framerate_source": m.framerate,
        "clip_location": str(get_clip_uri(clip.uuid, get_output_path_clips(output_path, filtered=filtered), "mp4")),
    }  # fmt: skip
    clip_metadata = None
    try:  # Clip.extract_metadata (data_model.py:283-308): the clip file's own facts - from its MP4 index here, ffprobe there
        clip_metadata = extract_clip_metadata(clip)
    except Exception as e:  # noqa: BLE001
        clip.errors["extract_metadata"] = str(e)
    if clip_metadata:
        data.update(clip_metadata)
    if clip.aesthetic_score is not None:
        data["aesthetic_score"] = clip.aesthetic_score
    if len(clip.errors) > 0:
        data["errors"] = list(clip.errors)
    data["windows"], data["filtered_windows"] = [], []
    data["valid"] = False  # bool(clip.encoded_data and len(clip.windows) > 0): this path creates no caption windows
    data["has_caption"], data["total_prompt_tokens"], data["total_output_tokens"] = False, 0, 0
    emb = get_clip_embedding(clip, embedding_algorithm)
    if emb is not None:
        data["embedding"] = np.asarray(emb).reshape(-1).tolist()
        data["embedding_model_name"] = embedding_algorithm
        data["embedding_model_version"] = embedding_model_version
    return data


def write_clip_metadata(clip, video, output_path: str, embedding_algorithm: str = "openai", *, filtered: bool = False) -> pathlib.Path:
    data = make_clip_metadata(clip, video, output_path, embedding_algorithm, filtered=filtered)
    data = {k: v for k, v in data.items() if k not in ("embedding", "embedding_model_name", "embedding_model_version")}  # :896-905: the pickle holds the vector
    return _write(get_clip_uri(clip.uuid, get_output_path_metas(output_path, "v0"), "json"), json.dumps(data, indent=4).encode())


def write_task_outputs(task, output_path: str, embedding_algorithm: str = "openai") -> dict[str, int]:
    """What ClipWriterStage._process_video does for the outputs of this path: pickle + parquet row per passing clip, json per clip."""
    stats = {"num_with_embeddings": 0, "parquets": 0, "metas": 0}
    for video in task.videos:
        for clip in video.clips:
            stats["num_with_embeddings"] += write_clip_embedding_pickle(clip, output_path, embedding_algorithm) is not None
            write_clip_metadata(clip, video, output_path, embedding_algorithm)
            stats["metas"] += 1
        for clip in video.filtered_clips:
            write_clip_metadata(clip, video, output_path, embedding_algorithm, filtered=True)
            stats["metas"] += 1
        stats["parquets"] += write_grouped_embeddings_parquet(video, output_path, embedding_algorithm) is not None
    return stats


def read_embedding_parquets(paths, display_name: str = "SemanticDedup") -> tuple[np.ndarray, np.ndarray]:
    """-> (ids [n] str, X [n, d] float32 C-contiguous), dedup_actor.py:199-222 on pyarrow."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    tables = [pq.read_table(str(p), columns=["id", "embedding"]) for p in paths]
    t = pa.concat_tables(tables) if tables else None
    n = 0 if t is None else t.num_rows
    if n == 0:
        error_message = f"{display_name}: no rows to cluster"
        raise ValueError(error_message)
    emb = t.column("embedding").combine_chunks()
    lens = np.diff(np.asarray(emb.offsets))
    dmin, dmax = int(lens.min()), int(lens.max())
    if dmin != dmax:
        error_message = f"{display_name}: ragged embeddings (min={dmin}, max={dmax}); SemDeDup requires fixed-length vectors"
        raise ValueError(error_message)
    x = np.ascontiguousarray(np.asarray(emb.flatten()).reshape(n, dmin).astype("float32"))
    return np.asarray(t.column("id").to_pylist()), x
