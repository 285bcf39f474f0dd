"""Standalone attention launch for ncu (ViT-L/14 shape: 257 tokens, 16 heads x 64)."""
import sys
sys.path.insert(0, ".")
import torch
from cosmos_curate_b200.runtime import Context
ctx = Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 264
qkv = (torch.randn(n, 257, 3072, device="cuda") * 1.5).half()
for _ in range(3):
    out = ctx.attention(qkv, 16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = ctx.attention(qkv, 16)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"attention n={n}: {ms:.3f} ms, {4*257*257*64*16*n/ms/1e9:.1f} TFLOP/s")
