"""Standalone preprocess launch for ncu: 64 synthetic 1080p NV12 surfaces -> patch-major fp16."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from cosmos_curate_b200.runtime import Context
from oracle import color
ctx = Context(0)
n, h, w, pitch, lr = 64, 1080, 1920, 2048, 1080
f = color.synthetic_nv12(h, w, seed=1, pitch=pitch)
buf = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(f[None], (n, h * 3 // 2, pitch)))).cuda()
pool = ctx.nv12_pool(buf, w, h, lr)
for _ in range(3):
    out = ctx.preprocess_clip(pool, dtype=torch.float16, layout="patch", patch=14, k_pad=640)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = ctx.preprocess_clip(pool, dtype=torch.float16, layout="patch", patch=14, k_pad=640)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"preprocess {n} frames: {ms:.3f} ms, {ms*1e3/n:.2f} us/frame, {2050656*n/ms/1e6:.1f} GB/s algorithmic")
