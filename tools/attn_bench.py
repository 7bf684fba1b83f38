"""Attention kernel A/B bench (GPU box): times f3r_attention at the bench shapes for every exp2-emulation variant
(f3r_set_option attn_emu = 0..5) and checks each variant against variant 0.  Writes gpurun_out/attn_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast3r_b200 import ops, lib as L  # noqa: E402

SHAPES = {  # name: (batch, heads, sq, skv, scale)
    "dec_N32_1gpu": (1, 16, 23552, 23552, 0.16019),
    "dec_N32_8gpu_shard": (1, 16, 2944, 23552, 0.16019),
    "dec_N32_4gpu_shard": (1, 16, 5888, 23552, 0.16019),
    "dec_N4_1gpu": (1, 16, 2944, 2944, 0.16019),
    "enc_32views": (32, 16, 736, 736, 0.125),
    "enc_4views": (4, 16, 736, 736, 0.125),
}
if "--long" in sys.argv:
    SHAPES["dec_N320_8gpu_shard"] = (1, 16, 29440, 235520, 0.16019)


def main():
    dev = torch.device("cuda")
    res = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, (b, h, sq, skv, scale) in SHAPES.items():
        D = h * 64
        g = torch.Generator().manual_seed(0)
        q = (torch.randn(b * sq, D, generator=g)).to(torch.bfloat16).to(dev)
        kv = (torch.randn(b * skv, 2 * D, generator=g)).to(torch.bfloat16).to(dev)
        out = torch.empty(b * sq, D, dtype=torch.bfloat16, device=dev)
        ref = None
        flops = 4.0 * b * sq * skv * D
        for split, emu in [(1, 0), (1, 2), (1, 3), (2, 0), (2, 1), (2, 2)]:
            L.set_option("attn_emu", emu)
            L.set_option("attn_split", split)
            ops.attention(q, kv, out, batch=b, heads=h, sq=sq, skv=skv, scale=scale)
            torch.cuda.synchronize()
            ts = []
            for it in range(5):
                flush.fill_(it)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.attention(q, kv, out, batch=b, heads=h, sq=sq, skv=skv, scale=scale)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            o = out.float()
            if emu == 0 and split == 1:
                ref = o.clone()
                err = 0.0
            else:
                err = float((o - ref).norm() / ref.norm())
            res[f"{name}/split{split}/emu{emu}"] = dict(ms=ms, tflops=flops / ms / 1e9, rel_vs_emu0=err, nan=bool(torch.isnan(o).any()))
            print(name, "split", split, "emu", emu, f"{ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s  rel vs emu0 {err:.2e}", flush=True)
        L.set_option("attn_emu", -1)
        L.set_option("attn_split", -1)
        for label, kw in (("default_direct", dict(kv_split=1)), ("default_auto_kvsplit", dict())):
            ts = []
            for it in range(6):
                flush.fill_(it)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.attention(q, kv, out, batch=b, heads=h, sq=sq, skv=skv, scale=scale, **kw)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts[1:])[len(ts[1:]) // 2]
            err = float((out.float() - ref).norm() / ref.norm())
            ns = ops.pick_kv_split(b * h * ((sq + 255) // 256), (skv + 127) // 128)
            res[f"{name}/{label}"] = dict(ms=ms, tflops=flops / ms / 1e9, rel_vs_emu0=err, kv_split=ns)
            print(name, label, f"(auto split {ns})", f"{ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s  rel {err:.2e}", flush=True)
        del q, kv, out, ref
    L.set_option("attn_emu", -1)
    L.set_option("attn_split", -1)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/attn_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
