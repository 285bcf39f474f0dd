"""Throughput of the shot-transition network on resident thumbnails (run on the B200 box).

    python tools/prof_shots.py [n_frames] [max_windows]
"""

import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cosmos_curate_b200.models.transnetv2 import seeded_state_dict
from cosmos_curate_b200.runtime import Context, ShotNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 9000  # 5 min @ 30 fps
mw = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = Context(0)
net = ShotNet(ctx, seeded_state_dict(3), max_windows=mw)
frames = torch.randint(0, 256, (n, 27, 48, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    net.predict(frames)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
ev0.record()
for _ in range(reps):
    p = net.predict(frames)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / reps
windows = -(-n // 50)
flop = windows * 82.4e9
print(f"{n} frames, {windows} windows, max_windows={mw}: {ms:.2f} ms/video, {n / ms * 1e3:.0f} frames/s, {flop / ms / 1e9:.1f} TFLOP/s fp32 (82.4 GFLOP/window)")
