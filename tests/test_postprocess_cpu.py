"""Host logic of fast3r_b200.postprocess (grouping / stacking / mask plumbing / result placement) on the CPU emulator of
the C ABI, against the reference's own outputs (tests/golden/geometry_tail.pt, tools/make_golden_geometry.py)."""
import os

import pytest
import torch

from fast3r_b200 import postprocess
from tests import abi_emulator

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "geometry_tail.pt")


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(postprocess, "ops", abi_emulator)
    monkeypatch.setattr(postprocess, "_device_of", lambda t, device: torch.device("cpu"))
    return postprocess


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN, weights_only=False)


def _views(gold):
    return [({} if vm is None else {"valid_mask": vm}) for vm in gold["valid_masks"]]


@pytest.mark.parametrize("group", [64, 2, 1])
def test_align_matches_reference(emulated, gold, monkeypatch, group):
    monkeypatch.setattr(postprocess, "_GROUP", group)
    for case in gold["align"]:
        preds = [{k: v.clone() for k, v in p.items()} for p in gold["preds"]]
        emulated.align_local_pts3d_to_global(preds, _views(gold), min_conf_thr_percentile=case["percentile"])
        for p, want in zip(preds, case["aligned"]):
            got = p["pts3d_local_aligned_to_global"]
            assert got.shape == want.shape and got.dtype == want.dtype
            assert (got - want).abs().max() <= 1e-5 * want.abs().max()


def test_align_requires_the_reference_keys(emulated, gold):
    preds = [{k: v for k, v in gold["preds"][0].items() if k != "conf"}]
    with pytest.raises(ValueError, match="global head confidence"):
        emulated.align_local_pts3d_to_global(preds, [{}])


def test_align_mixed_resolutions(emulated, gold):
    p0 = {k: v.clone() for k, v in gold["preds"][0].items()}
    p1 = {k: v[:, :32, :40].clone() for k, v in gold["preds"][1].items()}
    emulated.align_local_pts3d_to_global([p0, p1], [{}, {}])
    assert p0["pts3d_local_aligned_to_global"].shape == p0["pts3d_local"].shape
    assert p1["pts3d_local_aligned_to_global"].shape == p1["pts3d_local"].shape
    assert (p0["pts3d_local_aligned_to_global"] - gold["align"][0]["aligned"][0]).abs().max() < 1e-4


def test_focal_matches_reference(emulated, gold):
    for p, want_m, want_a in zip(gold["preds"], gold["focal_masked_p10_100it"], gold["focal_all_10it"]):
        b, h, w = p["conf"].shape
        for i in range(b):
            got = emulated.estimate_focal(p["pts3d_local"][i:i + 1], p["conf_local"][i:i + 1])
            assert isinstance(got, float) and abs(got - want_m[i]) <= 2e-5 * want_m[i]
        pp = torch.tensor([[w / 2, h / 2]]).expand(b, 2)
        got = emulated.estimate_focal_knowing_depth(p["pts3d_local"], pp, focal_mode="weiszfeld")
        assert torch.allclose(got, want_a, rtol=2e-5)
    with pytest.raises(ValueError):
        emulated.estimate_focal_knowing_depth(p["pts3d_local"], pp, focal_mode="median")


def test_no_cpu_path_without_cuda(gold):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    preds = [{k: v.clone() for k, v in gold["preds"][0].items()}]
    with pytest.raises(RuntimeError, match="no CPU path"):
        postprocess.align_local_pts3d_to_global(preds, [{}])
