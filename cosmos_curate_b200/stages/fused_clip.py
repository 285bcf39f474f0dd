"""NvdecClipAestheticStage: clip mp4 bytes -> NVDEC -> sampled NV12 surfaces -> fused preprocess -> tower -> scores.

Replaces the pair ClipFrameExtractionStage (CPU PyAV decode of every frame + RGB frames pickled to the next actor,
clip_frame_extraction_stages.py:102-165) -> AestheticFilterStage (aesthetic_filter_stages.py:120-221) with one GPU
stage that never materialises RGB frames.  Task mutations and the error convention are the reference's:

    clip.aesthetic_score, video.filtered_clips, video.clip_stats.num_filtered_by_aesthetic, task.stage_perf,
    no encoded_data            -> clip.errors["encoded_data"] = "empty", score -1.0
    demux / decode failure     -> clip.errors["frame_extraction"] = "video_decode_failed", encoded_data dropped, score -1.0
    (optional) clip.openai_embedding = L2-normalised mean of the per-frame embeddings (the reference has
    no pooling rule for frame embeddings; this documented choice fills the existing generic clip-embedding slot).
"""

from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Literal

import numpy as np
import torch

from .. import sampling
from .._lib import CurateB200Error
from ..data_model import StageTimer
from ..interfaces import CuratorStage, CuratorStageResource, ModelInterface
from ..models.clip_aesthetics import CLIPAestheticScorer
from ..runtime import Decoder, alloc_nv12_pool, get_context, mp4_index

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)


class NvdecClipAestheticStage(CuratorStage):
    def __init__(  # noqa: PLR0913
        self,
        score_threshold: float,
        reduction: Literal["mean", "min"] = "min",
        target_fps: float = 1.0,
        num_gpus_per_worker: float = 1.0,
        *,
        write_embedding: bool = False,
        max_batch: int = 256,
        num_decoders: int = 8,
        stage_batch_size: int = 8,
        seek_keyframes: bool = True,
        source: Literal["clip", "video_span"] = "clip",
        verbose: bool = False,
        log_stats: bool = False,
        model: CLIPAestheticScorer | None = None,
    ) -> None:
        self._timer = StageTimer(self)
        self._score_threshold, self._reduction, self._target_fps = score_threshold, reduction, target_fps
        self._num_gpus_per_worker = num_gpus_per_worker
        self._write_embedding, self._max_batch, self._num_decoders = write_embedding, max_batch, num_decoders
        self._stage_batch_size, self._verbose, self._log_stats = stage_batch_size, verbose, log_stats
        self._seek = seek_keyframes  # decode only the GOPs that contain sampled frames (identical frames, fewer decoded)
        if source not in ("clip", "video_span"):
            error_msg = f"source={source!r} not in ('clip', 'video_span')"
            raise ValueError(error_msg)
        # "video_span": analysis without the transcode - the clip's frames are decoded straight out of the SOURCE
        # video (video.encoded_data) at clip.span, so ClipTranscodingStage's re-encode + this stage's re-decode disappear for runs
        # that only need scores / embeddings.  Pixels are the source's, not the 4 Mb/s re-encode's: not bit-comparable with "clip".
        self._source = source
        self._video_index: dict[int, tuple] = {}
        self._model = model if model is not None else CLIPAestheticScorer(max_batch=max_batch)
        self._reduce_fn = np.min
        self._pools: dict[tuple[int, int], object] = {}

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(gpus=self._num_gpus_per_worker)

    @property
    def model(self) -> ModelInterface:
        return self._model

    @property
    def stage_batch_size(self) -> int:
        return self._stage_batch_size

    def stage_setup(self) -> None:
        if self._reduction not in ("mean", "min"):
            error_msg = f"Reduction `{self._reduction}` not implemented."
            raise NotImplementedError(error_msg)
        self._reduce_fn = np.mean if self._reduction == "mean" else np.min
        self._model.setup()
        self._ctx = get_context()
        self._decoders = [Decoder(self._ctx) for _ in range(self._num_decoders)]
        self._threads = ThreadPoolExecutor(max_workers=self._num_decoders)

    def destroy(self) -> None:
        for d in getattr(self, "_decoders", []):
            d.close()
        self._decoders = []
        if getattr(self, "_threads", None):
            self._threads.shutdown(wait=True)
        self._pools.clear()

    # ---- helpers ---------------------------------------------------------------------------------
    def _plan_span(self, clip, video):
        """Frames of the source video inside clip.span, re-timed from the clip start and sampled like a clip of its own."""
        data = video.encoded_data.resolve() if video.encoded_data else None
        if data is None:
            logger.warning(f"Clip {clip.uuid}: source video has no encoded_data.")
            clip.errors["encoded_data"] = "empty"
            clip.aesthetic_score = -1.0
            return None
        try:
            cached = self._video_index.get(id(video))
            if cached is None or cached[0] is not data:
                idx = mp4_index(data)
                cached = self._video_index[id(video)] = (data, idx, sampling.timestamps_from_index(idx["pts"], idx["timescale"]))
            _, idx, ts = cached
            ids = sampling.span_frame_ids(ts, clip.span, self._target_fps)
            return data, ids, ((idx["width"] + 1) & ~1, (idx["height"] + 1) & ~1)
        except (CurateB200Error, ValueError) as e:
            logger.error(f"Error extracting frames from clip {clip.uuid}: {e}")
            clip.errors["frame_extraction"] = "video_decode_failed"
            clip.aesthetic_score = -1.0
            return None

    def _plan(self, clip, video=None):
        """-> (data u8 array, frame ids expanded, (w, h)) or None (errors recorded on the clip)."""
        if self._source == "video_span":
            return self._plan_span(clip, video)
        data = clip.encoded_data.resolve() if clip.encoded_data else None
        if data is None:
            logger.warning(f"Clip {clip.uuid} has no encoded_data.")
            clip.errors["encoded_data"] = "empty"
            clip.aesthetic_score = -1.0
            return None
        try:
            idx = mp4_index(data)
            ts = sampling.timestamps_from_index(idx["pts"], idx["timescale"])
            ids, counts = sampling.frame_ids(ts, sampling.FrameExtractionPolicy.sequence, self._target_fps)
            return data, np.repeat(ids, counts).astype(np.int32), ((idx["width"] + 1) & ~1, (idx["height"] + 1) & ~1)
        except (CurateB200Error, ValueError) as e:
            self._decode_failed(clip, e)
            return None

    def _decode_failed(self, clip, e) -> None:
        logger.error(f"Error extracting frames from clip {clip.uuid}: {e}")
        clip.errors["frame_extraction"] = "video_decode_failed"
        if self._source == "clip":
            clip.encoded_data.drop()
        clip.aesthetic_score = -1.0

    def _pool(self, size):
        p = self._pools.get(size)
        if p is None:
            p = self._pools[size] = alloc_nv12_pool(self._ctx, self._max_batch, size[0], size[1])
        return p

    def _run_batch(self, pool, items) -> None:
        """items: [(clip, data, ids, first_slot)] whose frames fit the pool.  Decode in parallel, embed once."""

        def work(arg):
            j, (clip, data, ids, first) = arg
            try:
                self._decoders[j % self._num_decoders].decode(data, ids, pool, np.arange(first, first + len(ids), dtype=np.int32), seek_keyframes=self._seek)
                return None
            except CurateB200Error as e:
                return e

        errs = []
        for wave in range(0, len(items), self._num_decoders):  # a decoder is owned by one thread per wave
            errs.extend(self._threads.map(work, list(enumerate(items))[wave : wave + self._num_decoders]))
        n = sum(len(ids) for _, _, ids, _ in items)
        tower = self._model.tower
        emb, _, score = tower.embed_pool(pool, slots=np.arange(n, dtype=np.int32))
        score_h = score.cpu().numpy()
        emb_h = emb.cpu().numpy() if self._write_embedding else None
        for (clip, _, ids, first), err in zip(items, errs):
            if err is not None:
                self._decode_failed(clip, err)
                continue
            clip.aesthetic_score = float(self._reduce_fn(score_h[first : first + len(ids)]))
            if emb_h is not None:
                m = emb_h[first : first + len(ids)].mean(axis=0)
                clip.openai_embedding = (m / np.linalg.norm(m)).astype(np.float32)

    # ---- stage entry -----------------------------------------------------------------------------
    def process_data(self, tasks):
        by_size: dict[tuple[int, int], list] = {}
        for task in tasks:
            for video in task.videos:
                for clip in video.clips:
                    plan = self._plan(clip, video)
                    if plan is not None:
                        data, ids, size = plan
                        by_size.setdefault(size, []).append((clip, data, ids))
        for size, clips in by_size.items():
            pool = self._pool(size)
            batch, used = [], 0
            for clip, data, ids in clips:
                if len(ids) > self._max_batch:
                    self._decode_failed(clip, ValueError(f"{len(ids)} sampled frames exceed max_batch={self._max_batch}"))
                    continue
                if used + len(ids) > self._max_batch:
                    self._run_batch(pool, batch)
                    batch, used = [], 0
                batch.append((clip, data, ids, used))
                used += len(ids)
            if batch:
                self._run_batch(pool, batch)

        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            for video in task.videos:
                passed = []
                for clip in video.clips:
                    if clip.aesthetic_score is None:
                        clip.aesthetic_score = -1.0
                    if clip.aesthetic_score < self._score_threshold:
                        video.filtered_clips.append(clip)
                        video.clip_stats.num_filtered_by_aesthetic += 1
                    else:
                        passed.append(clip)
                video.clips = passed
            if self._log_stats:
                stage_name, stats = self._timer.log_stats()
                task.stage_perf[stage_name] = stats
        torch.cuda.current_stream().synchronize()
        self._video_index.clear()
        return tasks
