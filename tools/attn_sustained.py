"""Sustained (power-capped) attention rate per kernel variant: each variant runs back to back for a few seconds on the
N=32 decoder shape; the steady-state time per launch is taken from the last two thirds.  (Short isolated timings run at
boost clocks and rank the variants differently from a long step under the 1 kW cap.)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast3r_b200 import ops, lib as L  # noqa: E402

b, h, sq, skv, scale = 1, 16, 23552, 23552, 0.16019
D = h * 64
g = torch.Generator().manual_seed(0)
q = torch.randn(b * sq, D, generator=g).to(torch.bfloat16).cuda()
kv = torch.randn(b * skv, 2 * D, generator=g).to(torch.bfloat16).cuda()
out = torch.empty(b * sq, D, dtype=torch.bfloat16, device="cuda")
flops = 4.0 * b * sq * skv * D
secs = float(os.environ.get("SECS", "3"))
res = {}
variants = [(1, 0), (1, 2), (2, 0), (2, 1), (1, 3), (2, 0), (2, 1)]
for split, emu in variants:
    L.set_option("attn_emu", emu)
    L.set_option("attn_split", split)
    ops.attention(q, kv, out, batch=b, heads=h, sq=sq, skv=skv, scale=scale, kv_split=1)
    torch.cuda.synchronize()
    n = int(secs / 2.6e-3)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for i in range(n):
        if i == n // 3:
            evs[0].record()
        ops.attention(q, kv, out, batch=b, heads=h, sq=sq, skv=skv, scale=scale, kv_split=1)
    evs[1].record()
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[1]) / (n - n // 3)
    key = f"split{split}/emu{emu}"
    res.setdefault(key, []).append(ms)
    print(key, f"{ms:.3f} ms sustained  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
json.dump(res, open("gpurun_out/attn_sustained.json", "w"), indent=1)
