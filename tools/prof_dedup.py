"""Throughput of the fused cosine arg-max kernel (run on the B200 box):  python tools/prof_dedup.py [m] [d]"""
import sys

import torch

sys.path.insert(0, ".")
from cosmos_curate_b200 import dedup
from cosmos_curate_b200.runtime import Context

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
ctx = Context(0)
e = torch.randn(m, d, device="cuda")
dedup.l2_normalize_rows_(e, ctx)
dedup.rowdot_argmax(e, e, upper=True, clip=True, init_val=-1.0, ctx=ctx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
dedup.rowdot_argmax(e, e, upper=True, clip=True, init_val=-1.0, ctx=ctx)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
tiles = (m + 127) // 128
flop = 2.0 * 128 * 128 * d * tiles * (tiles + 1) / 2
print(f"semdedup pairwise m={m} d={d}: {ms:.1f} ms, {flop / ms / 1e9:.1f} TFLOP/s fp32 (upper triangle incl. diagonal tiles), S never written to HBM")
