"""Small instances of every kernel family for compute-sanitizer (memcheck / racecheck are 10-100x slower than native)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import kernel_checks as KC  # noqa: E402

RUN = [("layernorm_128", {}), ("im2col_patch", {}), ("upsample_crop", {}), ("x3_upsample_f32", {}), ("x3_split3", {}),
       ("x3_add_f32", {}), ("cast", {}), ("im2col3x3s2", {}), ("linear_small_tails", {}), ("linear_split", {}),
       ("rope_epilogue", {}), ("idxemb_epilogue", {}), ("conv1x1_tinymap", {}), ("conv3x3_w6_c96_res", {}),
       ("convT_k2", {}), ("final_fused", {}), ("attn_256x384", {}), ("attn_24", {}), ("attn_ranges_merge_rank0", {}),
       ("x3_attn_128", {}), ("x3_linear_gelu_tails", {}), ("linear_resid_splitk", {})]
table = {n: (f, kw) for n, f, kw in KC.ALL}
bad = []
for name, _ in RUN:
    f, kw = table[name]
    err, tol, info = f(**kw)
    torch.cuda.synchronize()
    ok = err <= tol
    print(("OK   " if ok else "FAIL ") + name, f"{err:.3e}", flush=True)
    if not ok:
        bad.append(name)
from fast3r_b200.ingest import ingest_rgb8  # noqa: E402
from oracle import ingest_oracle as O  # noqa: E402
img = np.random.default_rng(0).integers(0, 256, (301, 517, 3), dtype=np.uint8)
out, _ = ingest_rgb8(torch.from_numpy(img).cuda(), 224)
ref, _ = O.ingest(img, 224)
print("OK   ingest" if np.array_equal(out.cpu().numpy(), ref) else "FAIL ingest", flush=True)
print("FAILED:", bad)
