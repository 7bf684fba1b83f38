"""One (or a few) forward passes of the bench workload, for use under ncu.  Never a bench number."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_views  # noqa: E402
from fast3r_b200 import Fast3R, vit_large_args  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=32)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--enc-depth", type=int, default=24)
ap.add_argument("--dec-depth", type=int, default=24)
a = ap.parse_args()
enc, dec, head = vit_large_args()
enc["depth"], dec["depth"] = a.enc_depth, a.dec_depth
torch.manual_seed(0)
with torch.device("cuda"):
    model = Fast3R(enc, dec, head).eval()
views = make_views(a.views, device="cuda")
torch.manual_seed(7)
model(views)  # warm-up (weights packed, kernels loaded) outside the profiled range
torch.cuda.synchronize()
torch.cuda.profiler.start()  # use with: ncu --profile-from-start off
for _ in range(a.steps):
    torch.manual_seed(7)
    model(views)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
