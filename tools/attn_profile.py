"""One decoder-shape attention launch per exp2-emulation variant given on the command line (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast3r_b200 import ops, lib as L  # noqa: E402

variants = [int(v) for v in sys.argv[1:]] or [0]
b, h, sq, skv, scale = 1, 16, 23552, 23552, 0.16019
D = h * 64
g = torch.Generator().manual_seed(0)
q = torch.randn(b * sq, D, generator=g).to(torch.bfloat16).cuda()
kv = torch.randn(b * skv, 2 * D, generator=g).to(torch.bfloat16).cuda()
out = torch.empty(b * sq, D, dtype=torch.bfloat16, device="cuda")
for v in variants:
    L.set_option("attn_emu", v)
    ops.attention(q, kv, out, batch=b, heads=h, sq=sq, skv=skv, scale=scale)
    torch.cuda.synchronize()
