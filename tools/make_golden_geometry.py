"""Generates tests/golden/geometry_tail.pt by running the REFERENCE (in the build container only):

  * estimate_focal                       fast3r/models/multiview_dust3r_module.py:1081-1109   (real, pure torch)
  * estimate_focal_knowing_depth(weiszfeld)  fast3r/dust3r/post_process.py:19-79             (real, pure torch)
  * MultiViewDUSt3RLitModule.align_local_pts3d_to_global  multiview_dust3r_module.py:427-549  (real method; the
    absent third-party ``roma.rigid_points_registration`` it calls is supplied by oracle/geometry_oracle.umeyama,
    so those entries pin the masking / fallback / application logic, not roma's SVD)

Inputs are seeded synthetic pointmaps with the statistics of the model's outputs (a pin-hole camera looking at a smooth
depth surface, a global frame that is a similarity of the local one plus noise, conf = 1 + exp(.)).
Run: python tools/make_golden_geometry.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import geometry_oracle as go  # noqa: E402
from oracle import ref_harness  # noqa: E402


def roma_stub(x, y, compute_scaling=True):
    assert compute_scaling
    r, t, s = go.umeyama(x.double().numpy(), y.double().numpy())
    return torch.from_numpy(r).to(x.dtype), torch.from_numpy(t).to(x.dtype), torch.tensor(s, dtype=x.dtype)


def rand_rotation(g):
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def synth_pred(g, b, h, w, focal):
    v, u = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    yy, xx = v / h, u / w
    out = {k: [] for k in ("pts3d_local", "conf_local", "pts3d_in_other_view", "conf")}
    for _ in range(b):
        a = torch.rand(4, generator=g)
        z = 1.5 + a[0] + 0.6 * torch.sin(3 * xx + 6 * a[1]) * torch.cos(2 * yy + 6 * a[2]) + 0.02 * torch.randn(h, w, generator=g)
        loc = torch.stack([(u - w / 2) * z / focal, (v - h / 2) * z / focal, z], -1)
        loc = loc + 0.01 * torch.randn(h, w, 3, generator=g)
        r = rand_rotation(g).float()
        s = 0.5 + torch.rand(1, generator=g).item()
        t = torch.randn(3, generator=g)
        glob = s * (loc @ r.T) + t + 0.02 * torch.randn(h, w, 3, generator=g)
        out["pts3d_local"].append(loc)
        out["pts3d_in_other_view"].append(glob)
        out["conf_local"].append(1 + torch.exp(torch.randn(h, w, generator=g)))
        out["conf"].append(1 + torch.exp(torch.randn(h, w, generator=g)))
    return {k: torch.stack(v) for k, v in out.items()}


def main():
    lit_mod = ref_harness.import_reference_lit_module(roma_registration=roma_stub)
    from fast3r.dust3r.post_process import estimate_focal_knowing_depth

    g = torch.Generator().manual_seed(20260923)
    b, h, w = 2, 48, 64
    preds = [synth_pred(g, b, h, w, focal=70.0 + 10 * i) for i in range(3)]
    # view 1 carries a valid_mask, view 2 one with fewer than 3 valid pixels in batch item 1 (identity fallback)
    views = [{} for _ in preds]
    vm = torch.rand(b, h, w, generator=g) > 0.3
    views[1]["valid_mask"] = vm
    vm2 = torch.rand(b, h, w, generator=g) > 0.5
    vm2[1] = False
    vm2[1, 0, :2] = True
    views[2]["valid_mask"] = vm2

    cases = []
    for pct in (0, 30):
        ps = [{k: v.clone() for k, v in p.items()} for p in preds]
        lit_mod.MultiViewDUSt3RLitModule.align_local_pts3d_to_global(None, ps, views, min_conf_thr_percentile=pct)
        cases.append({"percentile": pct, "aligned": [p["pts3d_local_aligned_to_global"] for p in ps]})

    focal_masked = [[lit_mod.estimate_focal(p["pts3d_local"][i:i + 1], p["conf_local"][i:i + 1]) for i in range(b)] for p in preds]
    pp = torch.tensor([[w / 2, h / 2]]).expand(b, 2)
    focal_all = [estimate_focal_knowing_depth(p["pts3d_local"], pp, focal_mode="weiszfeld") for p in preds]
    quant = [[float(torch.quantile(p["conf"][i].reshape(-1), q)) for q in (0.0, 0.1, 0.3, 0.85, 1.0)] for p in preds for i in range(b)]

    out = {"preds": preds, "valid_masks": [v.get("valid_mask") for v in views], "align": cases,
           "focal_masked_p10_100it": focal_masked, "focal_all_10it": focal_all,
           "quantiles": {"q": [0.0, 0.1, 0.3, 0.85, 1.0], "values": quant},
           "what": "reference outputs, see tools/make_golden_geometry.py"}
    path = os.path.join(ROOT, "tests", "golden", "geometry_tail.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes;", "focals", focal_masked)


if __name__ == "__main__":
    main()
