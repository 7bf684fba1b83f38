import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import torch
    from fast3r_b200 import ops
    from tests.kernel_checks import _rand, attention_ref, rel
    use_lse, batch, heads, sq, skv = [int(x) for x in sys.argv[2:7]]
    D = heads * 64
    q = _rand((batch * sq, D), 1); kv = _rand((batch * skv, 2 * D), 2)
    out = torch.zeros(batch * sq, D, dtype=torch.bfloat16, device="cuda")
    lse = torch.zeros(batch, heads, sq, dtype=torch.float32, device="cuda") if use_lse else None
    ops.attention(q, kv, out, batch=batch, heads=heads, sq=sq, skv=skv, scale=0.125, lse=lse)
    torch.cuda.synchronize()
    qh = q.reshape(batch, sq, heads, 64).transpose(1, 2)
    kh = kv[:, :D].reshape(batch, skv, heads, 64).transpose(1, 2)
    vh = kv[:, D:].reshape(batch, skv, heads, 64).transpose(1, 2)
    ref = attention_ref(qh, kh, vh, 0.125).transpose(1, 2).reshape(batch * sq, D)
    print("RESULT", rel(out, ref))
    sys.exit(0)
cases = [(0, 1, 1, 128, 128), (1, 1, 1, 128, 128), (0, 2, 2, 736, 736), (1, 2, 2, 736, 736), (0, 1, 16, 2944, 2944), (0, 1, 2, 512, 3072)]
for split in ("2", "1"):
    for c in cases:
        env = dict(os.environ, F3R_ATTN_SPLIT=split)
        try:
            p = subprocess.run([sys.executable, __file__, "one"] + [str(x) for x in c], capture_output=True, text=True, timeout=45, env=env)
            tail = [l for l in p.stdout.splitlines() if l.startswith("RESULT")] or [(p.stdout + p.stderr)[-300:]]
            print("split", split, c, tail[-1], flush=True)
        except subprocess.TimeoutExpired:
            print("split", split, c, "TIMEOUT", flush=True)
