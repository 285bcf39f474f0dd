#!/usr/bin/env python
"""bench.py - throughput of the decode -> sample -> preprocess -> embed/classify hot path on B200.

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps K --warmup W # the reference's CPU path (oracle) on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic 1920x1080 30 fps 10 s H.264 clips (tools/synth_h264.py: no encoder
exists in this image), sampled at 1 fps like AestheticFilterStage (11 frames per clip), CLIP ViT-L/14 image tower
(seeded random weights) + aesthetic affine head.  One STEP = `--clips-per-step` clips (default 24 -> 264 frames,
the closest whole-clip count to the 256-frame batch BASELINE.json names).

Printed JSON line (rank 0):
  value   clips/s, whole job, decoded NV12 surfaces of the step already resident in HBM (preprocess + tower + head),
          timed with CUDA events, max over ranks.
  e2e     clips/s through the public stage API from HOST mp4 buffers: MP4 index + NVDEC decode + fused preprocess +
          tower + D2H of embeddings/scores, wall clock between device synchronisations, max over ranks.
  roofline      the dominant kernel (tcgen05 GEMM): algorithmic FLOPs per launch / CUDA-event time per launch vs the
                measured sustained bf16 peak (MEASURED_PEAKS.json); roofline_other has preprocess / LayerNorm (HBM).
  cpu_baseline  the oracle's CPU restatement of the reference path timed on the host cores (N=1, rank 0), bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("OPENCV_LOG_LEVEL", "ERROR")

import numpy as np  # noqa: E402

FRAME_W, FRAME_H, FPS, SECONDS = 1920, 1080, 30, 10.0
SAMPLE_FPS = 1.0
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2 :] if sm else []  # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_clips(n_distinct: int, rank: int) -> list[bytes]:
    from tools import synth_h264

    return [synth_h264.make_clip(FRAME_W, FRAME_H, FPS, SECONDS, seed=1000 * rank + i, gop=FPS, pan=(2, 0)) for i in range(n_distinct)]


# ================================================================================================ reference arm
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 measures the CPU path
    from oracle import cpu_path, vit

    cores = os.cpu_count() or 1
    procs, threads = cpu_layout(cores)
    pool = cpu_path.CpuReferencePool(vit.CLIP_VIT_L14, seed=0, procs=procs, threads=threads)
    clips = make_clips(2, 0)
    sample = max(procs, args.ref_clips)  # at least one clip per worker so every host core is busy
    batch = [clips[i % len(clips)] for i in range(sample)]
    pool.run(batch[:procs], SAMPLE_FPS)  # one warm-up pass (a CPU step takes ~25 s; more warm-up would only burn minutes)
    t, frames, phases = 0.0, 0, {"decode_s": 0.0, "preprocess_s": 0.0, "model_s": 0.0}
    for _ in range(args.steps):
        r = pool.run(batch, SAMPLE_FPS)
        t += r["seconds"]
        frames += r["frames"]
        for k in phases:
            phases[k] += r[k]
    pool.close()
    value = sample * args.steps / t
    line = {
        "impl": "reference", "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_sec": frames / t,
        "config": {"workload": WORKLOAD, "implementation": "reference CPU path (oracle port): libavcodec decode, torchvision transforms, torch-fp32 tower", "clips_per_step": sample,
                   "frames_per_clip": frames // (sample * args.steps), "network": "clip-vit-large-patch14 (seeded random weights), fp32", "sharding": "host worker processes"},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": procs * threads, "kind": "port",
                         "sample": f"{sample} clip(s) per step x {args.steps} steps over {procs} worker processes x {threads} threads; cv2/libavcodec decode "
                                   "(PyAV stand-in) + torchvision transforms + oracle torch-fp32 tower, one model call per clip",
                         "worker_seconds": phases},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": cores,
    }  # fmt: skip
    print(json.dumps(line))


# ================================================================================================ this repo's arm
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # its banner goes to stdout: keep stdout to the one JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from concurrent.futures import ThreadPoolExecutor

    from cosmos_curate_b200 import sampling
    from cosmos_curate_b200.models import weights as W
    from cosmos_curate_b200.runtime import Context, Decoder, VitTower, alloc_nv12_pool, mp4_index

    ctx = Context(local)
    cfg = W.CLIP_VIT_L14
    cps = args.clips_per_step
    clips = make_clips(args.distinct_clips, rank)
    plans = []
    for c in clips:
        idx = mp4_index(c)
        ts = sampling.timestamps_from_index(idx["pts"], idx["timescale"])
        ids, counts = sampling.frame_ids(ts, sampling.FrameExtractionPolicy.sequence, SAMPLE_FPS)
        plans.append(np.repeat(ids, counts).astype(np.int32))
    fpc = len(plans[0])
    frames_per_step = cps * fpc
    tower = VitTower(ctx, cfg.to_dict(), W.seeded_weights(cfg, 0), max_batch=frames_per_step, aesthetic=W.seeded_aesthetic(cfg.proj_dim, 0))
    pools = [alloc_nv12_pool(ctx, frames_per_step, FRAME_W, FRAME_H) for _ in range(3)]
    n_dec = args.decoders
    tp = ThreadPoolExecutor(max_workers=n_dec)       # one NVDEC session per worker thread (thread-local, reused across clips)
    tls = threading.local()
    all_decoders = []
    host_emb = torch.empty((frames_per_step, cfg.proj_dim), dtype=torch.float32).pin_memory()
    host_score = torch.empty((frames_per_step,), dtype=torch.float32).pin_memory()
    step_clips = [i % len(clips) for i in range(cps)]

    def my_decoder():
        d = getattr(tls, "dec", None)
        if d is None:
            d = tls.dec = Decoder(ctx)
            all_decoders.append(d)
        return d

    def decode_clip(j, pool, seek):
        k = step_clips[j]
        return my_decoder().decode(clips[k], plans[k], pool, np.arange(j * fpc, (j + 1) * fpc, dtype=np.int32), seek_keyframes=seek)["frames_decoded"]

    def decode_step(pool, seek=False):
        return sum(tp.map(lambda j: decode_clip(j, pool, seek), range(cps)))

    def decode_ceiling(reps: int) -> float:
        """All NVDEC sessions decoding whole clips, surfaces discarded: frames/s."""
        from cosmos_curate_b200.runtime import decode_discard

        list(tp.map(lambda j: decode_discard(my_decoder(), clips[step_clips[j]]), range(n_dec)))  # warm-up: sessions created
        t0 = time.perf_counter()
        n = sum(tp.map(lambda j: decode_discard(my_decoder(), clips[step_clips[j % cps]]), range(reps)))
        return n / (time.perf_counter() - t0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident-input measurement (value): decoded surfaces of one step already in HBM
    decode_step(pools[0])
    barrier()
    for _ in range(max(args.warmup, 3)):
        tower.embed_pool(pools[0])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.profile_begin()
    ev0.record()
    for _ in range(args.steps):
        emb, _, score = tower.embed_pool(pools[0])
    ev1.record()
    prof = ctx.profile_end()
    barrier()
    dev_s = max_over_ranks(ev0.elapsed_time(ev1) / 1e3)
    launches = ctx.launch_count() - l0
    value = world * cps * args.steps / dev_s

    # ---- end-to-end measurement (e2e): host mp4 bytes -> NVDEC -> preprocess -> tower -> host results.
    # Decode of steps i+1, i+2 (NVDEC engines + host parsing threads) overlaps the tower of step i (SMs); three surface pools.
    def e2e_run(n_steps: int, seek: bool):
        # three surface pools: the clips of steps i+1 and i+2 are queued on the decode threads while the tower works on step i,
        # so no NVDEC session idles at a step boundary waiting for the slowest clip of the step
        futs = {}

        def submit(i):
            futs[i] = [tp.submit(decode_clip, j, pools[i % 3], seek) for j in range(cps)]

        for i in range(min(2, n_steps)):
            submit(i)
        decoded = 0
        for i in range(n_steps):
            decoded += sum(f.result() for f in futs.pop(i))
            if i + 2 < n_steps:
                submit(i + 2)  # pool (i+2)%3 was last read by the tower of step i-1, synchronised below
            emb, _, score = tower.embed_pool(pools[i % 3])
            host_emb.copy_(emb, non_blocking=True)
            host_score.copy_(score, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return decoded

    def e2e_measure(seek: bool):
        e2e_run(max(1, min(args.warmup, 2)), seek)
        barrier()
        t0 = time.perf_counter()
        decoded = e2e_run(args.e2e_steps, seek)
        barrier()
        sec = max_over_ranks(time.perf_counter() - t0)
        h2d = sum(len(clips[k]) for k in step_clips)
        return {"value": world * cps * args.e2e_steps / sec, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": frames_per_step * (cfg.proj_dim + 1) * 4,
                "steps": args.e2e_steps, "ms_per_step": 1e3 * sec / args.e2e_steps, "frames_per_sec": world * frames_per_step * args.e2e_steps / sec,
                "decoded_frames_per_sec": world * decoded / sec, "decoded_frames_per_step": decoded // args.e2e_steps, "nvdec_sessions": n_dec}  # fmt: skip

    e2e = e2e_sparse = ceiling = None
    e2e_error = None
    if not args.no_e2e:
        try:
            e2e = e2e_measure(seek=False)
            e2e["note"] = ("every frame up to the last sampled one is decoded (reference semantics, decoder_utils.py:439-455); synthetic I_PCM + "
                           "motion-only P pictures; decode of the next two steps overlaps the tower of step i")
            e2e_sparse = e2e_measure(seek=True)
            e2e_sparse["note"] = "CB_DECODE_SEEK_SYNC: only GOPs holding sampled frames are decoded (identical frames); closed GOP = 30, 1 fps sampling"
            ceil_fps = decode_ceiling(2 * n_dec)
            ceiling = {"decode_only_frames_per_sec_per_gpu": ceil_fps, "sessions": n_dec,
                       "e2e_fraction_of_decode_ceiling": (e2e["decoded_frames_per_sec"] / world) / ceil_fps if ceil_fps > 0 else None}
        except Exception as exc:  # noqa: BLE001 - e.g. libnvcuvid missing on the box: report it, keep the device-resident numbers
            e2e_error = f"{type(exc).__name__}: {exc}"
    shot = None
    if rank == 0 and not args.no_shots:
        try:
            shot = shot_detection_measure(ctx, torch)
        except Exception as exc:  # noqa: BLE001
            shot = {"error": f"{type(exc).__name__}: {exc}"}
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    traffic = ncu_traffic()
    gemm_flops = cfg.gemm_flops_per_image() * frames_per_step * args.steps
    gemm_ms, gemm_n = prof["gemm"]["ms"], max(1, prof["gemm"]["launches"])
    ach_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    pre_bytes = (1.5 * min(FRAME_W, FRAME_H) ** 2 + 3 * 224 * 224 * 2) * frames_per_step * args.steps
    ln_bytes = 6.0 * cfg.tokens * cfg.hidden * frames_per_step * (2 * cfg.layers) * args.steps  # fp32 in + fp16 out per LayerNorm
    other = {}
    for name, nbytes, key in (("preprocess", pre_bytes, "preprocess"), ("layernorm", ln_bytes, "layernorm")):
        ms = prof[key]["ms"]
        gbs = nbytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        other[name] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                       "traffic": traffic.get({"preprocess": "clip_preprocess", "layernorm": "layernorm_kernel"}[name]),
                       "ms_per_step": ms / args.steps, "launches_per_step": prof[key]["launches"] / args.steps}  # fmt: skip
    other["attention"] = {"ms_per_step": prof["attention"]["ms"] / args.steps, "launches_per_step": prof["attention"]["launches"] / args.steps}
    step_ms = 1e3 * dev_s / args.steps
    line = {
        "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "frames_per_sec": world * frames_per_step * args.steps / dev_s,
        "config": {"workload": WORKLOAD, "implementation": "one process per B200: NVDEC + fused preprocess kernel + tcgen05 tower",
                   "clips_per_step": cps, "frames_per_clip": fpc, "frames_per_step": frames_per_step, "sample_fps": SAMPLE_FPS, "distinct_clips": args.distinct_clips,
                   "network": "clip-vit-large-patch14, seeded random weights, fp16 operands / fp32 accumulate+residual", "sharding": f"{world} rank(s), clips sharded per rank, no data-path collective",
                   "l2": "inputs (NV12 pool 0.88 GB + activations > 1 GB) exceed the 126 MB L2", "value_inputs": "decoded NV12 surfaces resident in HBM"},
        "clocks": clocks, "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_2cta_kernel", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                     "traffic": traffic.get("gemm_tcgen05_2cta"), "traffic_source": traffic.get("_source"),
                     "peak_source": f"{pk_src} bf16_tflops_sustained (kernel timed inside a long step)", "launches_per_step": gemm_n / args.steps,
                     "ms_per_step": gemm_ms / args.steps, "share_of_step": gemm_ms / args.steps / step_ms},
        "roofline_other": other,
        "e2e": e2e, "e2e_keyframe_seek": e2e_sparse, "decode_roofline": ceiling,
    }  # fmt: skip
    if shot is not None:
        if "tflops_fp32" in shot and clocks:
            peak = ctx.device_info()["sm_count"] * 128 * 2 * clocks["sm_max_mhz"] * 1e6 / 1e12  # FFMA lanes x 2 flop x clock
            shot["fp32_peak_tflops"] = peak
            shot["frac_of_fp32_peak"] = shot["tflops_fp32"] / peak
        line["shot_detection"] = shot
    if e2e_error:
        line["e2e_error"] = e2e_error
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(clips[:1])
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def shot_flops_per_window(frames: int = 100) -> float:
    """2*M*N*K over the convolutions and Linear layers of the shot network for one window (transnetv2.py:66-100)."""
    total = 0.0
    hw = [(27, 48), (13, 24), (6, 12)]
    for s in range(3):
        f = 16 << s
        m = frames * hw[s][0] * hw[s][1]
        for cin in ((3 if s == 0 else 2 * f), 4 * f):  # first block reads the previous stack (4 * f/2 channels), second block 4f
            total += 2.0 * m * (8 * f) * (9 * cin) + 2.0 * m * (4 * f) * (6 * f)
    return total + 2.0 * frames * (4864 * 1024 + 448 * 128 + 1024)


def shot_detection_measure(ctx, torch, n_frames: int = 9000, reps: int = 3) -> dict:
    """Secondary row: the fp32 shot-transition network over one 5-minute video of resident 27x48 thumbnails."""
    from cosmos_curate_b200.models.transnetv2 import seeded_state_dict
    from cosmos_curate_b200.runtime import ShotNet

    net = ShotNet(ctx, seeded_state_dict(0), max_windows=16)
    frames = torch.randint(0, 256, (n_frames, 27, 48, 3), dtype=torch.uint8, device=f"cuda:{ctx.device}")
    for _ in range(2):
        net.predict(frames)
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        net.predict(frames)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    windows = -(-n_frames // 50)
    net.close()
    return {"workload": f"{n_frames} frames 27x48 RGB (5 min @ 30 fps) resident in HBM, {windows} windows of 100 frames / stride 50, fp32",
            "frames_per_sec": n_frames / ms * 1e3, "ms_per_video": ms, "ms_per_window": ms / windows, "gflop_per_window": shot_flops_per_window() / 1e9,
            "tflops_fp32": windows * shot_flops_per_window() / ms / 1e9, "gpu_launches": (ctx.launch_count() - l0) // reps}  # fmt: skip


# the same workload name on both arms (the driver pairs the lines by metric + config)
WORKLOAD = "1080p30 10 s H.264 clips -> 1 fps frame sampling -> preprocess -> CLIP-ViT-L/14 embed + aesthetic score (BASELINE.json configs[1])"


def ncu_traffic() -> dict:
    """DRAM bytes per launch from the committed `ncu --set full` capture (profiles/rNN_traffic.json, newest round); {} if none."""
    import glob

    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_traffic.json")))
    if not files:
        return {}
    with open(files[-1]) as f:
        d = json.load(f)
    out = {k: v["traffic_bytes_per_launch"] for k, v in d["kernels"].items()}
    out["_source"] = os.path.join("profiles", os.path.basename(files[-1]))
    return out


def cpu_layout(cores: int) -> tuple[int, int]:
    """(worker processes, torch threads each): the reference scales this path by replicating actors, not by threading one model."""
    if os.environ.get("CB_REF_PROCS") and os.environ.get("CB_REF_THREADS"):  # tuning override
        return int(os.environ["CB_REF_PROCS"]), int(os.environ["CB_REF_THREADS"])
    # measured on the 128-thread (64-core) B200 host: 16 workers x 4 threads = 0.58 clips/s, 8 x 8 = 0.51, 8 x 16 = 0.27-0.32,
    # 16 x 8 = 0.36, 4 x 32 = 0.19 -> one torch thread per PHYSICAL core, four per worker
    if cores >= 32:
        return cores // 8, 4
    return max(1, cores // 4), min(4, cores) if cores >= 4 else 1


def cpu_baseline(clips: list[bytes]) -> dict:
    """Bounded CPU sample of the same workload: the oracle's restatement of the reference path (kind 'port')."""
    from oracle import cpu_path, vit

    cores = os.cpu_count() or 1
    procs, threads = cpu_layout(cores)
    pool = cpu_path.CpuReferencePool(vit.CLIP_VIT_L14, seed=0, procs=procs, threads=threads)
    try:
        pool.run([clips[i % len(clips)] for i in range(procs)], SAMPLE_FPS)  # warm-up: one clip per worker
        n = 2 * procs
        r = pool.run([clips[i % len(clips)] for i in range(n)], SAMPLE_FPS)
    finally:
        pool.close()
    return {"value": r["clips"] / r["seconds"], "unit": "clips/s", "cores": procs * threads, "kind": "port",
            "sample": f"{n} clips (1080p30 10 s, {r['frames']} sampled frames) over {procs} worker processes x {threads} threads: cv2/libavcodec decode + "
                      "torchvision transforms + oracle torch-fp32 ViT-L/14, one model call per clip",
            "frames_per_sec": r["frames"] / r["seconds"], "worker_seconds": {k: r[k] for k in ("decode_s", "preprocess_s", "model_s")}}  # fmt: skip


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips-per-step", type=int, default=24)
    ap.add_argument("--distinct-clips", type=int, default=4)
    ap.add_argument("--decoders", type=int, default=20, help="concurrent NVDEC sessions per GPU (7 engines on B200)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-clips", type=int, default=2, help="clips per step of the reference arm (bounded sample)")
    ap.add_argument("--no-shots", action="store_true", help="skip the shot-detection secondary measurement")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
