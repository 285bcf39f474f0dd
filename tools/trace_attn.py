import sys, os
sys.path.insert(0, ".")
import torch
from cosmos_curate_b200.runtime import Context
ctx = Context(0)
qkv = (torch.randn(264, 257, 3072, device="cuda") * 1.5).half()
for _ in range(2): ctx.attention(qkv, 16)
torch.cuda.synchronize()
os.environ["CB_ATTN_DEBUG_TRACE"] = "1"
ctx.attention(qkv, 16)
torch.cuda.synchronize()
