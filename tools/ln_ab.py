"""A/B of the LayerNorm instantiations (CB_LN_VARIANT is read once per process): rows = 264 x 257 tokens, d = 1024 (the bench
step's shape; 417 MB per launch, larger than L2).  Prints ms per launch, algorithmic GB/s and a hash of the output."""
import hashlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmos_curate_b200.runtime import get_context  # noqa: E402

ctx = get_context()
rows, d = 264 * 257, int(os.environ.get("LN_D", "1024"))
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((rows, d), device="cuda", generator=g)
gamma = torch.randn((d,), device="cuda", generator=g)
beta = torch.randn((d,), device="cuda", generator=g)
y = ctx.layernorm(x, gamma, beta, 1e-5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = []
for _ in range(3):
    e0.record()
    for _ in range(50):
        ctx.layernorm(x, gamma, beta, 1e-5)
    e1.record()
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / 50)
ms = min(best)
print(json.dumps({"variant": os.environ.get("CB_LN_VARIANT", "default"), "d": d, "ms": ms, "gbs": rows * d * 6 / ms / 1e6,
                  "sha": hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:16]}))
