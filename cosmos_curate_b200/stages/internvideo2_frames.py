"""InternVideo2FrameCreationStage on the B200 path - same name, constructor and task mutations as the reference stage
(cosmos_curate/pipelines/video/embedding/internvideo2_stages.py:43-184): `clip.intern_video_2_frames` <- float32
[1, 8, 3, 224, 224], the tube the video tower's encode_video_frames consumes (:277-296).

source="frames" (default, drop-in): reads the host RGB frames a ClipFrameExtractionStage left in `clip.extracted_frames`
    under this stage's signature; only the 8 frames the stride keeps cross PCIe.  A clip with fewer sampled frames than the
    model needs is re-extracted at 2x the rate (up to 20 fps) from `encoded_data` (:157-176) - on NVDEC here.
source="nvdec": no upstream frame extraction; the sampled frame ids come from the MP4 index, the re-extraction rule is
    applied to the id lists (no decode needed to know how many frames a rate yields), the 8 kept frames are decoded
    straight into NV12 surfaces and resized / normalised from there.  Host frames never exist; 4.8 MB per clip come back.
"""

from __future__ import annotations

import numpy as np

from .. import sampling
from .._lib import CurateB200Error
from ..data_model import StageTimer
from ..interfaces import CuratorStage, CuratorStageResource, ModelInterface
from ..models.internvideo2_frames import InternVideo2FrameFormulator, select_frame_ids
from ..runtime import DecoderPool, alloc_nv12_pool, get_context, mp4_index
from ..sampling import FrameExtractionPolicy, FrameExtractionSignature

try:
    from loguru import logger
except Exception:  # noqa: BLE001
    import logging

    logger = logging.getLogger(__name__)

MAX_FPS = 20  # internvideo2_stages.py:138


def sampled_ids_with_regen(ts: np.ndarray, target_fps: float, target_num_frames: int, max_fps: int = MAX_FPS):
    """Frame ids of the `sequence` policy at target_fps, the rate doubled while it yields fewer than target_num_frames
    frames and stays <= max_fps (internvideo2_stages.py:157-176).  Returns (ids, rate actually used)."""
    def expanded(rate):
        ids, counts = sampling.frame_ids(ts, FrameExtractionPolicy.sequence, rate)
        return np.repeat(ids, counts)

    fps = used = target_fps
    ids = expanded(fps)
    while len(ids) < target_num_frames:
        fps *= 2
        if fps > max_fps:
            break
        ids, used = expanded(fps), fps
    return ids, used


class InternVideo2FrameCreationStage(CuratorStage):
    """Stage for creating InternVideo2 input frames from video clips."""

    GROUP = 32  # clips per decode group of the nvdec source (32 x 8 surfaces: 0.8 GB of 1080p NV12 per pool, two pools)

    def __init__(self, target_fps: float = 2.0, *, verbose: bool = False, log_stats: bool = False, source: str = "frames",
                 num_gpus_per_worker: float = 0.1, num_decoders: int = 8, stage_batch_size: int = 1, colour: str = "swscale",
                 model: InternVideo2FrameFormulator | None = None) -> None:  # fmt: skip
        if source not in ("frames", "nvdec"):
            msg = f"source={source!r} not in ('frames', 'nvdec')"
            raise ValueError(msg)
        self._timer = StageTimer(self)
        self._target_fps = target_fps
        self._extraction_policy = FrameExtractionPolicy.sequence
        self._frame_extraction_signature = FrameExtractionSignature(extraction_policy=FrameExtractionPolicy.sequence, target_fps=target_fps).to_str()
        self._model = model if model is not None else InternVideo2FrameFormulator()
        self._verbose, self._log_stats = verbose, log_stats
        self._source, self._colour, self._num_gpus = source, colour, num_gpus_per_worker
        self._num_decoders, self._stage_batch_size = num_decoders, stage_batch_size
        self._decode_pool = None
        self._pools: dict[tuple, list] = {}

    @property
    def model(self) -> ModelInterface:
        return self._model

    @property
    def resources(self) -> CuratorStageResource:
        return CuratorStageResource(cpus=1.0, gpus=self._num_gpus)  # the reference stage is CPU-only (cpus=1.0, :93)

    @property
    def stage_batch_size(self) -> int:
        return self._stage_batch_size

    def stage_setup(self) -> None:
        self._model.setup()
        self._ctx = get_context()

    def destroy(self) -> None:
        if self._decode_pool is not None:
            self._decode_pool.close()
            self._decode_pool = None
        self._pools.clear()

    # ---- NVDEC side --------------------------------------------------------------------------------
    def _decoders(self) -> DecoderPool:
        if self._decode_pool is None:
            self._decode_pool = DecoderPool(self._ctx, self._num_decoders)
        return self._decode_pool

    def _pool(self, size: tuple[int, int], r: int, n: int):
        ring = self._pools.get(size)
        if ring is None:
            if len(self._pools) >= 4:
                self._pools.pop(next(iter(self._pools)))
            ring = self._pools[size] = [None, None]
        if ring[r] is None or ring[r].buf.shape[0] < n:
            ring[r] = None
            ring[r] = alloc_nv12_pool(self._ctx, n, size[0], size[1], self._colour)
        return ring[r]

    def _plan(self, clip, data):
        """-> (size, distinct frame ids to decode, slot of every kept frame relative to the clip's first slot), None when the
        clip is too short (tube = the reference's empty array), or raises for an unreadable container."""
        idx = mp4_index(data)
        ts = sampling.timestamps_from_index(idx["pts"], idx["timescale"])
        fn = self._model.get_target_num_frames()
        ids, fps = sampled_ids_with_regen(ts, self._target_fps, fn)
        if len(ids) < fn:
            logger.error(f"Clip {clip.uuid} is too short to extract enough frames.")
            logger.error(f"Frame count {len(ids)} is smaller than minimal requirement {fn}")
            return None
        if self._verbose and fps != self._target_fps:
            logger.warning(f"Clip {clip.uuid} has <{fn} frames at target_fps={self._target_fps}; sampled at {fps}.")
        keep = np.asarray(ids)[select_frame_ids(len(ids), fn)]
        uniq, inverse = np.unique(keep, return_inverse=True)  # a frame kept twice (supersampled clip) is decoded once
        return ((idx["width"] + 1) & ~1, (idx["height"] + 1) & ~1), uniq.astype(np.int32), inverse.astype(np.int32)

    def _tubes_from_streams(self, items) -> None:
        """items: [(clip, data)].  Decode groups of GROUP clips on the session pool (group k+1 decodes while group k is
        resized, normalised and copied out), one tube kernel launch per group."""
        fn = self._model.get_target_num_frames()
        by_size: dict[tuple, list] = {}
        for clip, data in items:
            try:
                plan = self._plan(clip, data)
            except (CurateB200Error, ValueError) as e:
                self._decode_failed(clip, e)
                continue
            if plan is None:
                clip.intern_video_2_frames = np.empty(0, dtype=np.float32)
                continue
            by_size.setdefault(plan[0], []).append((clip, data, plan[1], plan[2]))
        groups = [(size, clips[i : i + self.GROUP]) for size, clips in by_size.items() for i in range(0, len(clips), self.GROUP)]
        ring_pos: dict[tuple, int] = {}

        def decode_one(dec, data, ids, pool, first):
            dec.decode(data, ids, pool, np.arange(first, first + len(ids), dtype=np.int32))

        def submit(k):
            size, clips = groups[k]
            r = ring_pos.get(size, 0)
            ring_pos[size] = r ^ 1
            cap, need = fn, sum(len(ids) for _, _, ids, _ in clips)
            while cap < need:
                cap *= 2
            pool = self._pool(size, r, cap)
            futs, first = [], 0
            for _, data, ids, _ in clips:
                futs.append((first, self._decoders().submit(decode_one, data, ids, pool, first, shape=size)))
                first += len(ids)
            return pool, futs

        pending = submit(0) if groups else None
        for k, (_, clips) in enumerate(groups):
            pool, futs = pending
            ok, slots = [], []
            for (clip, _, _, inverse), (first, fut) in zip(clips, futs):
                try:
                    fut.result()
                except CurateB200Error as e:
                    self._decode_failed(clip, e)
                    continue
                ok.append(clip)
                slots.append(first + inverse)
            pending = submit(k + 1) if k + 1 < len(groups) else None  # next group's decode overlaps this group's kernel + D2H
            if ok:
                tubes = self._model.formulate_pool(pool, np.concatenate(slots)).cpu().numpy()  # [len(ok) * fn, 3, s, s]
                for i, clip in enumerate(ok):
                    clip.intern_video_2_frames = tubes[i * fn : (i + 1) * fn][None].copy()

    @staticmethod
    def _decode_failed(clip, e) -> None:
        logger.error(f"Error extracting frames from clip {clip.uuid}: {e}")
        clip.errors["frame_extraction"] = "video_decode_failed"

    # ---- the reference's process_data ----------------------------------------------------------
    def process_data(self, tasks):
        if self._source == "nvdec":
            return self._process_streams(tasks)
        for task in tasks:
            self._timer.reinit(self, task.get_major_size())
            video = task.video
            for clip in video.clips:
                data = clip.encoded_data.resolve() if clip.encoded_data else None
                if data is None:
                    clip.errors["encoded_data"] = "empty"
                    continue
                ef = clip.extracted_frames.resolve()
                if ef is None or self._frame_extraction_signature not in ef:
                    clip.errors[f"frames-{self._frame_extraction_signature}"] = "missing"
                    logger.error(f"Clip {clip.uuid} has buffer but no extracted frames for {self._frame_extraction_signature}")
                    continue
                with self._timer.time_process():
                    frames = ef[self._frame_extraction_signature]
                    if frames.shape[0] < self._model.get_target_num_frames():
                        # re-extract at a higher rate from the stream (internvideo2_stages.py:157-176): the id-list rule lands on
                        # the rate the reference's decode-and-count loop stops at; a clip still too short gets the empty
                        # float32 array `_construct_frames` returns (internvideo2_mm.py:396-398)
                        self._tubes_from_streams([(clip, data)])
                    else:
                        clip.intern_video_2_frames = self._model.formulate_input_frames(list(frames))
                clip.extracted_frames.drop()

            if self._log_stats:
                stage_name, stage_perf_stats = self._timer.log_stats()
                task.stage_perf[stage_name] = stage_perf_stats
        return tasks

    def _process_streams(self, tasks):
        """source="nvdec": clips of all tasks of the call share the decode groups, so the timer window is per call."""
        self._timer.reinit(self, sum(task.get_major_size() for task in tasks))
        items = []
        for task in tasks:
            for clip in task.video.clips:
                data = clip.encoded_data.resolve() if clip.encoded_data else None
                if data is None:
                    clip.errors["encoded_data"] = "empty"
                    continue
                items.append((clip, data))
        with self._timer.time_process(num_samples=max(1, len(items))):
            self._tubes_from_streams(items)
        if self._log_stats:
            stage_name, stats = self._timer.log_stats()
            for task in tasks:
                task.stage_perf[stage_name] = stats
        return tasks
