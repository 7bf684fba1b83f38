"""Pins oracle/fast3r_oracle.py against outputs of the UNMODIFIED reference (fixtures made by
tests/golden/make_golden.py in the build container).  CPU only."""
import os

import pytest
import torch

from oracle import fast3r_oracle as O
from tests.conftest import rel_l2
from tests.golden.synth import synth_state_dict, synth_images

TOL = 2e-5  # fp32 CPU, different summation order (SDPA-flash vs naive, conv algorithms)


ALL_TAGS = ["tiny_b1_n3", "tiny_b2_n2", "tiny_noattnbias", "tiny_fixedidx", "tiny_nolocal_n2", "tiny_single_view",
            "tiny_trainmode"]


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_tiny_end_to_end(golden_dir, tag):
    g = torch.load(os.path.join(golden_dir, f"{tag}.pt"))
    enc, dec, head = O.tiny_args()
    dec.update(g.get("dec_over", {}))
    head.update(g.get("head_over", {}))
    sd = synth_state_dict(g["shapes"], seed=g["weight_seed"])
    imgs = synth_images(g["N"], g["B"], g["H"], g["W"])
    taps = {}
    torch.manual_seed(g["rng_seed"])  # image ids must come out of the same RNG stream
    preds = O.forward(sd, enc, dec, head, imgs, training=g.get("train_mode", False), taps=taps)
    for i, (p, q) in enumerate(zip(preds, g["preds"])):
        assert sorted(p) == sorted(q)
        for k in q:
            assert p[k].shape == q[k].shape
            assert rel_l2(p[k], q[k]) < TOL, (i, k, rel_l2(p[k], q[k]))
    # per-stage taps (only stored for the B=1 fixture)
    name_map = {"dec_block0": "dec_block0", "dec_block5": "dec_block5", "dec_block11": "dec_block11",
                "patch_embed": "patch_embed", "enc_block0": "enc_block0", "enc_block1": "enc_block1",
                "enc_out": "enc_out", "layer_rn0": "layer_rn0", "layer_rn1": "layer_rn1",
                "layer_rn2": "layer_rn2", "layer_rn3": "layer_rn3"}
    for gk, ok in name_map.items():
        if gk in g["taps"]:
            assert rel_l2(taps[ok], g["taps"][gk]) < TOL, gk
    if "path3_uncropped" in g["taps"]:
        assert rel_l2(taps["path3"], g["taps"]["path3_uncropped"]) < TOL


def test_mixed_resolution_views(golden_dir):
    g = torch.load(os.path.join(golden_dir, "tiny_mixed_res.pt"))
    enc, dec, head = O.tiny_args()
    sd = synth_state_dict(g["shapes"], seed=g["weight_seed"])
    imgs = [synth_images(1, g["B"], h, w, seed0=1234 + i)[0] for i, (h, w) in enumerate(g["sizes"])]
    torch.manual_seed(g["rng_seed"])
    preds = O.forward(sd, enc, dec, head, imgs)
    for i, (p, q) in enumerate(zip(preds, g["preds"])):
        for k in q:
            assert p[k].shape == q[k].shape
            assert rel_l2(p[k], q[k]) < TOL, (i, k, rel_l2(p[k], q[k]))


def test_image_id_rng_stream(golden_dir):
    g = torch.load(os.path.join(golden_dir, "tiny_b1_n3.pt"))
    torch.manual_seed(g["rng_seed"])
    assert torch.equal(O.draw_image_ids(g["B"], g["N"]), g["image_ids"])


def test_vitl_width_blocks(golden_dir):
    g = torch.load(os.path.join(golden_dir, "vitl_blocks.pt"))
    sd = synth_state_dict(g["shapes"], seed=g["weight_seed"])
    x, pos = g["x"], g["pos"]
    y = O.block(x, sd, "", 16, 1e-6, 64 ** -0.5, pos)
    assert rel_l2(y, g["enc_block"]) < TOL
    xd = x.reshape(1, -1, 1024)
    assert rel_l2(O.block(xd, sd, "", 16, 1e-5, O.attn_bias_scale(64), None), g["dec_block_eval"]) < TOL
    assert rel_l2(O.block(xd, sd, "", 16, 1e-5, 64 ** -0.5, None), g["dec_block_train"]) < TOL
    assert rel_l2(O.rope2d(g["rope_q"], pos), g["rope_out"]) < 1e-6
    assert abs(O.attn_bias_scale(64) - 0.16019) < 1e-5


def vitl_n4_model_inputs(g):
    """ViT-L state dict + views of the vitl_n4_368x512 fixture (BASELINE.json configs[0])."""
    from fast3r_b200 import Fast3R, vit_large_args
    enc, dec, head = vit_large_args()
    with torch.device("meta"):
        shapes = {k: tuple(v.shape) for k, v in Fast3R(enc, dec, head).state_dict().items()}
    sd = synth_state_dict(shapes, seed=g["weight_seed"])
    imgs = synth_images(g["N"], g["B"], g["H"], g["W"])
    return (enc, dec, head), sd, imgs


def test_vitl_n4_368x512_config0(golden_dir):
    """Full ViT-L/512, N=4 views 512x368 (BASELINE configs[0]): oracle vs the reference's own inference(dtype="32")
    output (every 4th pixel is stored; per-view full-resolution moments are checked too)."""
    g = torch.load(os.path.join(golden_dir, "vitl_n4_368x512.pt"))
    cfg, sd, imgs = vitl_n4_model_inputs(g)
    torch.manual_seed(g["rng_seed"])
    preds = O.forward(sd, *cfg, imgs)
    st = g["stride"]
    for i, (p, q) in enumerate(zip(preds, g["preds_sub"])):
        for k in q:
            assert rel_l2(p[k][:, ::st, ::st], q[k]) < TOL, (i, k, rel_l2(p[k][:, ::st, ::st], q[k]))
            mean, std, amax = g["moments"][i][k]
            assert abs(float(p[k].double().mean()) - mean) < 1e-4 * max(abs(mean), std)
            assert abs(float(p[k].double().std()) - std) < 1e-4 * std
