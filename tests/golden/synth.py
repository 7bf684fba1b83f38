"""Deterministic synthetic weights / inputs shared by the golden generator and the tests.

A 647 M-parameter (or even the tiny model's 41 M-parameter DPT heads) state dict is too big to
commit, so fixtures hold only inputs' seeds and the reference's OUTPUTS; weights are re-created
from (key, shape, seed) here, bit-identically on any box with the same torch build.
"""
import math
import torch


def synth_state_dict(shapes: dict, seed: int = 0, gain: float = 1.0) -> dict:
    """shapes: {key: tuple}.  Linear/conv weights ~ N(0, gain/sqrt(fan_in)); norm weights
    1 + 0.1 N(0,1); biases 0.1 N(0,1).  Keys are visited in sorted order."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        if len(shp) >= 2:
            if "act_postprocess" in k and k.endswith(".1.weight") and ("0.1" in k or "1.1" in k) and len(shp) == 4 \
                    and (".act_postprocess.0.1." in k or ".act_postprocess.1.1." in k):
                fan_in = shp[0]  # ConvTranspose2d weight is (in, out, kh, kw), k = s -> one tap per output
            else:
                fan_in = 1
                for d in shp[1:]:
                    fan_in *= d
            # the DPT stack is ~10 convs deep with residual sums: damp it so the head output stays O(1)
            # (pts3d = expm1(|xyz|), conf = 1+exp(c) amplify anything larger)
            g_k = gain * (0.8 if "downstream_head" in k else 1.0)
            sd[k] = r * (g_k / math.sqrt(fan_in))
        elif "norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * r
        else:
            sd[k] = 0.1 * r
    # scratch.layer_rn.{i} aliases scratch.layer{i+1}_rn (same tensor in the reference module)
    for k in list(sd):
        if ".scratch.layer_rn." in k:
            i = int(k.split(".scratch.layer_rn.")[1].split(".")[0])
            sd[k] = sd[k.replace(f".scratch.layer_rn.{i}.", f".scratch.layer{i + 1}_rn.")]
    return sd


def synth_images(n_views: int, batch: int, H: int, W: int, seed0: int = 1234):
    """img_i = rand(B,3,H,W, generator=manual_seed(1234+i))*2-1."""
    return [torch.rand(batch, 3, H, W, generator=torch.Generator().manual_seed(seed0 + i)) * 2 - 1
            for i in range(n_views)]
