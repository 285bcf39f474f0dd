"""Oracle: the reference's CPU path for one clip, end to end (timed by bench.py as the CPU baseline).

Restates the call chain SURVEY.md 3.3 documents:

  ClipFrameExtractionStage._process_video   clip_frame_extraction_stages.py:102-165
    extract_frames                          decoder_utils.py:611-672
      get_video_timestamps                  :230-278   (container open #1)
      sample_closest                        :315-386
      decode_video_cpu_frame_ids            :389-461   (decode EVERY frame up to the last id, keep the wanted ones as RGB)
  AestheticFilterStage.process_data         aesthetic_filter_stages.py:181-183  (one model call per clip)
    _CLIPImageEmbeddings.__call__           clip.py:64-74  (torchvision transforms -> CLIP image tower fp32 -> L2 norm)
    MLP                                     aesthetics.py:44-53

Stand-ins forced by this image (no PyAV, no ffmpeg CLI): decode goes through cv2.VideoCapture - the same
libavcodec H.264 decoder and swscale colour conversion PyAV would drive - from a temp file in /dev/shm, with
per-frame PTS taken as index / fps for the constant-frame-rate synthetic clips.  The transforms are the real
torchvision ones on torch-CPU; the tower is oracle.vit.forward (torch fp32 on all host threads).

Test infrastructure only (see oracle/__init__.py).
"""

from __future__ import annotations

import os
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

os.environ.setdefault("OPENCV_LOG_LEVEL", "ERROR")

import numpy as np
import torch

from oracle import sampling, vit


def decode_sampled_frames(mp4: bytes, fps: float, decode_threads: int = 4) -> tuple[np.ndarray, np.ndarray]:
    """-> (uint8 [n, H, W, 3] RGB frames, frame ids).  Mirrors extract_frames(policy=sequence, sample_rate_fps=fps)."""
    import cv2

    with tempfile.NamedTemporaryFile(suffix=".mp4", dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as f:
        f.write(mp4)
        f.flush()
        cap = cv2.VideoCapture(f.name, cv2.CAP_FFMPEG, [cv2.CAP_PROP_N_THREADS, int(decode_threads)])
        n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        rate = float(cap.get(cv2.CAP_PROP_FPS))
        ts = (np.arange(n, dtype=np.float64) / rate).astype(np.float32)
        ids, counts = sampling.frame_ids_for(ts, "sequence", fps)
        want = {int(i): int(c) for i, c in zip(ids, counts)}
        last = int(ids[-1])
        out = []
        for i in range(last + 1):
            ok, bgr = cap.read()  # decodes every frame, like the PyAV loop of decode_video_cpu_frame_ids
            if not ok:
                raise RuntimeError(f"decode failed at frame {i}")
            if i in want:
                out.extend([bgr[..., ::-1]] * want[i])
        cap.release()
    return np.ascontiguousarray(np.stack(out)), np.repeat(ids, counts)


def reference_transforms():
    """The transform chain of clip.py:48-62, built from torchvision itself."""
    from torchvision import transforms

    return transforms.Compose(
        [
            transforms.Resize(224, interpolation=transforms.InterpolationMode.BICUBIC, antialias=True),
            transforms.CenterCrop(224),
            transforms.ConvertImageDtype(torch.float32),
            transforms.Normalize(mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711)),
        ]
    )


class CpuReferencePath:
    """decode -> sample -> preprocess -> embed -> score for a list of clips, timing each phase."""

    def __init__(self, cfg: vit.VitConfig, weights: dict, aesthetic_sd: dict | None, threads: int | None = None):
        self.cfg, self.w, self.sd = cfg, {k: torch.as_tensor(v) for k, v in weights.items()}, aesthetic_sd
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.tf = reference_transforms()

    @torch.no_grad()
    def run(self, clips: list[bytes], fps: float = 1.0, decode_workers: int | None = None, decode_threads: int = 4) -> dict:
        t0 = time.perf_counter()
        # the reference runs several ClipFrameExtractionStage actors (3 CPUs each, 4 decode threads) side by side
        workers = decode_workers or max(1, self.threads // 4)
        with ThreadPoolExecutor(max_workers=min(workers, len(clips))) as ex:
            decoded = list(ex.map(lambda c: decode_sampled_frames(c, fps, decode_threads), clips))
        t1 = time.perf_counter()
        embs, scores, n_frames = [], [], 0
        t_pre = t_model = 0.0
        for frames, _ in decoded:  # one model call per clip (aesthetic_filter_stages.py:181-183)
            ta = time.perf_counter()
            x = self.tf(torch.from_numpy(frames).permute(0, 3, 1, 2))
            tb = time.perf_counter()
            e = vit.forward(self.cfg, self.w, x)["embedding"]
            s = vit.aesthetic_mlp_forward(self.sd, e) if self.sd is not None else None
            tc = time.perf_counter()
            t_pre += tb - ta
            t_model += tc - tb
            embs.append(e)
            scores.append(s)
            n_frames += len(frames)
        t2 = time.perf_counter()
        return {"seconds": t2 - t0, "decode_s": t1 - t0, "preprocess_s": t_pre, "model_s": t_model, "clips": len(clips), "frames": n_frames,
                "embeddings": embs, "scores": scores}  # fmt: skip


# ---- several reference actors side by side (one process each, like the reference's per-stage Ray actors) ------------
_W = {}


def _worker_init(cfg_dict: dict, seed: int, threads: int) -> None:
    torch.set_num_threads(threads)
    cfg = vit.VitConfig(**cfg_dict)
    _W["path"] = CpuReferencePath(cfg, vit.random_weights(cfg, seed=seed), vit.random_aesthetic_mlp(seed=seed, in_dim=cfg.proj_dim or cfg.hidden),
                                  threads=threads)  # fmt: skip


def _worker_run(args):
    clip, fps = args
    r = _W["path"].run([clip], fps, decode_workers=1)
    return {k: r[k] for k in ("seconds", "decode_s", "preprocess_s", "model_s", "clips", "frames")}


class CpuReferencePool:
    """`procs` worker processes x `threads` torch threads; each worker runs whole clips through CpuReferencePath."""

    def __init__(self, cfg: vit.VitConfig, seed: int, procs: int, threads: int):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor

        self.procs, self.threads = procs, threads
        self.ex = ProcessPoolExecutor(max_workers=procs, mp_context=mp.get_context("spawn"), initializer=_worker_init, initargs=(cfg.to_dict(), seed, threads))
        list(self.ex.map(_noop, range(procs * 2)))  # start the workers (imports + weights) before anything is timed

    def run(self, clips: list[bytes], fps: float = 1.0) -> dict:
        t0 = time.perf_counter()
        parts = list(self.ex.map(_worker_run, [(c, fps) for c in clips]))
        sec = time.perf_counter() - t0
        out = {"seconds": sec, "clips": len(clips), "frames": sum(p["frames"] for p in parts)}
        for k in ("decode_s", "preprocess_s", "model_s"):
            out[k] = sum(p[k] for p in parts)  # summed over workers (CPU-seconds of wall inside workers)
        return out

    def close(self):
        self.ex.shutdown(wait=True, cancel_futures=True)


def _noop(i):
    time.sleep(0.05)
    return i
