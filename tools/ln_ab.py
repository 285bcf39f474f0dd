"""A/B of the LayerNorm instantiations (CB_LN_VARIANT is read per call): rows = 264 x 257 tokens (the bench step's shape; > L2 per
launch).  Prints ms per launch, algorithmic GB/s and a hash of the output per (d, variant)."""
import hashlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmos_curate_b200.runtime import get_context  # noqa: E402

ctx = get_context()
rows = 264 * 257
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for d in (1024, 1152, 768):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((rows, d), device="cuda", generator=g)
    gamma = torch.randn((d,), device="cuda", generator=g)
    beta = torch.randn((d,), device="cuda", generator=g)
    for v in ("0", "1", "2", "0", "1", "2"):
        os.environ["CB_LN_VARIANT"] = v
        y = ctx.layernorm(x, gamma, beta, 1e-5)
        torch.cuda.synchronize()
        best = []
        for _ in range(3):
            e0.record()
            for _ in range(50):
                ctx.layernorm(x, gamma, beta, 1e-5)
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1) / 50)
        ms = min(best)
        print(json.dumps({"variant": v, "d": d, "ms": ms, "gbs": rows * d * 6 / ms / 1e6, "sha": hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:16]}), flush=True)
