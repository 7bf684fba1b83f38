"""The geometry tail every caller runs on the forward's outputs (first slice), on the GPU.

Same names, arguments and results as the reference:

* ``align_local_pts3d_to_global(preds, views, min_conf_thr_percentile=0)`` -
  MultiViewDUSt3RLitModule.align_local_pts3d_to_global (fast3r/models/multiview_dust3r_module.py:427-549): adds
  ``pts3d_local_aligned_to_global`` (B, H, W, 3) to every pred.  The reference loops over (view, batch item) pairs in a
  CPU thread pool (torch.quantile + boolean gathers + roma SVD per pair); here all pairs go through three kernels
  (exact radix-select quantile, masked moments + Umeyama solve, streaming apply).
* ``estimate_focal(pts3d_i, conf_i, pp=None, min_conf_thr_percentile=10)`` - multiview_dust3r_module.py:1081-1109
  (returns a python float), and ``estimate_focal_knowing_depth(pts3d, pp, focal_mode="weiszfeld")`` -
  fast3r/dust3r/post_process.py:19-79 (returns a (B,) tensor).

NOT here (documented in DESIGN.md §1): fast_pnp / cv2.solvePnPRansac (cloud_opt/init_im_poses.py:300-350) and the
"median" focal mode.  Tensors may live on the CPU (what ``inference()`` returns) or on a CUDA device; CPU inputs are
copied to ``device`` (default cuda:0) and the results copied back, so the function is a drop-in either way.  There is
no CPU implementation: without the CUDA library this raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops

_GROUP = 64  # (view, batch) pairs stacked per kernel call: bounds the staging copy to ~64 x 5.3 MB at 512x368


def _device_of(t: torch.Tensor, device) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("fast3r_b200.postprocess needs a CUDA device (there is no CPU path)")
    return torch.device(device if device is not None else "cuda:0")


def _f32(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()


def align_local_pts3d_to_global(preds: List[Dict], views: List[Dict], min_conf_thr_percentile: float = 0, device=None) -> None:
    for pred in preds:
        for key, what in (("pts3d_local", "Key 'pts3d_local' not found in preds."),
                          ("conf_local", "Key 'conf_local' not found in preds."),
                          ("pts3d_in_other_view", "Key 'pts3d_in_other_view' not found in preds."),
                          ("conf", "Key 'conf' (global head confidence) not found in preds.")):
            if key not in pred:
                raise ValueError(what)
    if not preds:
        return
    dev = _device_of(preds[0]["pts3d_local"], device)
    q = float(min_conf_thr_percentile) / 100.0
    for g0 in range(0, len(preds), _GROUP):
        group = preds[g0:g0 + _GROUP]
        gviews = views[g0:g0 + _GROUP] if views is not None else [{}] * len(group)
        shapes = {tuple(p["pts3d_local"].shape) for p in group}
        if len(shapes) != 1:  # mixed resolutions: one call per pred
            for p, v in zip(group, gviews):
                _align_group([p], [v], q, dev)
        else:
            _align_group(group, gviews, q, dev)


def _align_group(group: List[Dict], gviews: List[Dict], q: float, dev: torch.device) -> None:
    b, h, w, _ = group[0]["pts3d_local"].shape
    n = h * w
    x = torch.cat([_f32(p["pts3d_local"], dev).reshape(b, n, 3) for p in group])
    y = torch.cat([_f32(p["pts3d_in_other_view"], dev).reshape(b, n, 3) for p in group])
    conf = torch.cat([_f32(p["conf"], dev).reshape(b, n) for p in group])
    valid = None
    if any("valid_mask" in v for v in gviews):
        valid = torch.cat([
            (v["valid_mask"].to(dev).reshape(b, n) if "valid_mask" in v else torch.ones(b, n, dtype=torch.bool, device=dev))
            .to(torch.uint8) for v in gviews]).contiguous()
    thr = ops.conf_quantile(conf, q)
    rts = ops.similarity_fit(x, y, conf, thr, valid)
    out = ops.similarity_apply(x, rts)
    for i, p in enumerate(group):
        src = p["pts3d_local"]
        aligned = out[i * b:(i + 1) * b].reshape(b, h, w, 3)
        p["pts3d_local_aligned_to_global"] = aligned.to(device=src.device, dtype=src.dtype)  # no copy if already there


def estimate_focal(pts3d_i: torch.Tensor, conf_i: torch.Tensor, pp: Optional[torch.Tensor] = None,
                   min_conf_thr_percentile: float = 10, device=None) -> float:
    b, h, w, three = pts3d_i.shape
    assert three == 3
    assert b == 1  # the reference processes one sample at a time
    dev = _device_of(pts3d_i, device)
    pts = _f32(pts3d_i, dev)
    conf = _f32(conf_i, dev).reshape(b, h, w)
    thr = ops.conf_quantile(conf.reshape(b, h * w), float(min_conf_thr_percentile) / 100.0)
    ppt = None if pp is None else _f32(torch.as_tensor(pp), dev).reshape(b, 2)
    return float(ops.focal_weiszfeld(pts, conf, thr, ppt, iters=100)[0])


def estimate_focal_knowing_depth(pts3d: torch.Tensor, pp: torch.Tensor, focal_mode: str = "weiszfeld", device=None) -> torch.Tensor:
    if focal_mode != "weiszfeld":
        raise ValueError(f"bad {focal_mode=} (only 'weiszfeld' is implemented on the GPU)")
    b, h, w, three = pts3d.shape
    assert three == 3
    dev = _device_of(pts3d, device)
    ppt = _f32(torch.as_tensor(pp), dev).reshape(-1, 2).expand(b, 2).contiguous()
    return ops.focal_weiszfeld(_f32(pts3d, dev), None, None, ppt, iters=10).to(pts3d.device)
