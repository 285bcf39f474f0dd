"""Minimal preprocess launch for compute-sanitizer."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from cosmos_curate_b200.runtime import Context
from oracle import color
ctx = Context(0)
h, w, pitch, lr = 64, 96, 128, 64
f = color.synthetic_nv12(h, w, seed=1)
buf = np.zeros((1, lr + h // 2, pitch), dtype=np.uint8)
buf[0, :h, :w] = f[:h]; buf[0, lr:lr + h // 2, :w] = f[h:]
pool = ctx.nv12_pool(torch.from_numpy(buf).cuda(), w, h, lr)
out = ctx.preprocess_clip_u8(pool)
torch.cuda.synchronize()
print("ok", out.float().mean().item())
