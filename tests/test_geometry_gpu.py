"""Geometry tail on the GPU (csrc/geometry.cu through the C ABI) against the oracle, the reference-generated golden
fixture, and - at the full 512x368 size - size-independent properties.

Tolerances: the confidence quantile is bit-exact (integer radix select + ATen's lerp); the similarity and the focal are
floating point with fp64 accumulators, compared at 1e-5 relative (of the pointmap's magnitude) / 2e-5 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import geometry_oracle as go

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "geometry_tail.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN, weights_only=False)


@pytest.fixture(scope="module")
def ops():
    from fast3r_b200 import ops as o
    return o


def test_quantile_is_exact(ops):
    g = torch.Generator().manual_seed(0)
    for views, n in ((1, 1), (3, 2), (2, 1000), (4, 3072), (5, 4099), (2, 368 * 512)):
        conf = 1 + torch.exp(torch.randn(views, n, generator=g))
        conf[0, : n // 3] = conf[0, 0]  # ties
        if n > 10:
            conf[-1, 5] = float("inf")
            conf[-1, 7] = -3.5
        for q in (0.0, 0.1, 0.3, 0.5, 0.85, 0.999, 1.0):
            got = ops.conf_quantile(conf.cuda(), q).cpu()
            want = torch.stack([torch.quantile(c, q) for c in conf])
            same = (got == want) | (got.isnan() & want.isnan())  # an infinite order statistic interpolates to NaN in ATen
            assert bool(same.all()), (views, n, q, got, want)
            with np.errstate(invalid="ignore"):
                for v in range(views):
                    o = float(go.conf_quantile(conf[v].numpy(), q))
                    assert float(got[v]) == o or (np.isnan(o) and bool(got[v].isnan()))


def test_quantile_mask_matches_torch_at_full_size(ops):
    g = torch.Generator().manual_seed(1)
    conf = (1 + torch.exp(2 * torch.randn(32, 368 * 512, generator=g))).cuda()
    for q in (0.1, 0.85):
        thr = ops.conf_quantile(conf, q)
        want = torch.quantile(conf, q, dim=1)
        assert torch.equal(thr, want)


def test_align_matches_golden_and_oracle(ops, gold):
    from fast3r_b200 import postprocess
    views = [({} if vm is None else {"valid_mask": vm}) for vm in gold["valid_masks"]]
    for case in gold["align"]:
        for on_gpu in (False, True):
            preds = [{k: (v.cuda() if on_gpu else v.clone()) for k, v in p.items()} for p in gold["preds"]]
            postprocess.align_local_pts3d_to_global(preds, views, min_conf_thr_percentile=case["percentile"])
            for p, want in zip(preds, case["aligned"]):
                got = p["pts3d_local_aligned_to_global"]
                assert got.is_cuda == on_gpu and got.shape == want.shape and got.dtype == want.dtype
                assert (got.cpu() - want).abs().max() <= 1e-5 * want.abs().max()


def test_similarity_fit_against_oracle_edge_cases(ops, gold):
    g = torch.Generator().manual_seed(2)
    for n in (3, 7, 130, 4099):  # n % 4 != 0 exercises the scalar apply path
        x = torch.randn(3, n, 3, generator=g)
        y = 1.3 * x.flip(-1) + torch.randn(3, n, 3, generator=g) * 0.05 + 2
        conf = 1 + torch.exp(torch.randn(3, n, generator=g))
        valid = (torch.rand(3, n, generator=g) > 0.2)
        valid[:, :3] = True
        valid[2] = False  # view 2: no valid pixel -> identity
        thr = ops.conf_quantile(conf.cuda(), 0.3)
        rts = ops.similarity_fit(x.cuda(), y.cuda(), conf.cuda(), thr, valid.to(torch.uint8).cuda()).cpu().double()
        out = ops.similarity_apply(x.cuda(), rts.float().cuda()).cpu()
        for v in range(3):
            want, r, t, s = go.align_local_to_global(x[v].reshape(1, n, 3).numpy(), conf[v].reshape(1, n).numpy(),
                                                     y[v].reshape(1, n, 3).numpy(), valid[v].reshape(1, n).numpy(), 30.0)
            assert np.allclose(rts[v, :9].reshape(3, 3).numpy(), r, atol=5e-6), (n, v)
            assert np.allclose(rts[v, 9:12].numpy(), t, atol=2e-5), (n, v)
            assert abs(float(rts[v, 12]) - s) <= 5e-6 * abs(s), (n, v)
            assert np.abs(out[v].numpy() - want.reshape(n, 3)).max() <= 2e-5 * max(1.0, np.abs(want).max())
    # no masks at all
    rts = ops.similarity_fit(x.cuda(), y.cuda()).cpu()
    r, t, s = go.umeyama(x[0].numpy(), y[0].numpy())
    assert np.allclose(rts[0, :9].reshape(3, 3).numpy(), r, atol=5e-6) and abs(float(rts[0, 12]) - s) < 5e-6 * s


def test_full_size_similarity_round_trip(ops):
    """N = 32 views at 512x368: a known similarity of the local pointmap is recovered and applied exactly; aligning twice
    is idempotent (the second fit is the identity)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    views, n = 32, 368 * 512
    x = torch.randn(views, n, 3, device="cuda", generator=g) * 2 + 1
    q, _ = torch.linalg.qr(torch.randn(views, 3, 3, generator=torch.Generator().manual_seed(4), dtype=torch.float64))
    q[:, :, 0] *= torch.sign(torch.det(q)).reshape(-1, 1)
    q = q.cuda()
    s = 0.5 + torch.rand(views, device="cuda", generator=g, dtype=torch.float64)
    t = torch.randn(views, 3, device="cuda", generator=g, dtype=torch.float64)
    y = (s.reshape(-1, 1, 1) * (x.double() @ q.transpose(1, 2)) + t.reshape(-1, 1, 3)).float()
    conf = 1 + torch.exp(torch.randn(views, n, device="cuda", generator=g))
    thr = ops.conf_quantile(conf, 0.5)
    rts = ops.similarity_fit(x, y, conf, thr)
    assert (rts[:, :9].reshape(-1, 3, 3).double() - q).abs().max() < 1e-6
    assert ((rts[:, 12].double() - s).abs() / s).max() < 1e-6
    assert (rts[:, 9:12].double() - t).abs().max() < 1e-5
    out = ops.similarity_apply(x, rts)
    assert (out - y).abs().max() <= 1e-5 * y.abs().max()
    rts2 = ops.similarity_fit(out, y, conf, thr)
    eye = torch.eye(3, device="cuda").reshape(1, 9)
    assert (rts2[:, :9] - eye).abs().max() < 1e-6 and (rts2[:, 12] - 1).abs().max() < 1e-6
    assert rts2[:, 9:12].abs().max() < 1e-5


def test_focal_matches_golden_and_oracle(ops, gold):
    from fast3r_b200 import postprocess
    for p, want_m, want_a in zip(gold["preds"], gold["focal_masked_p10_100it"], gold["focal_all_10it"]):
        b, h, w = p["conf"].shape
        for i in range(b):
            got = postprocess.estimate_focal(p["pts3d_local"][i:i + 1], p["conf_local"][i:i + 1])
            assert abs(got - want_m[i]) <= 2e-5 * want_m[i], (got, want_m[i])
            got = postprocess.estimate_focal(p["pts3d_local"][i:i + 1].cuda(), p["conf_local"][i:i + 1].cuda(),
                                             pp=torch.tensor([w / 2 + 1.5, h / 2 - 2.0]))
            want = go.focal_weiszfeld(p["pts3d_local"][i].numpy(), (w / 2 + 1.5, h / 2 - 2.0),
                                      (p["conf_local"][i] >= torch.quantile(p["conf_local"][i].reshape(-1), 0.1)).numpy(), 100)
            assert abs(got - want) <= 2e-5 * want
        pp = torch.tensor([[w / 2, h / 2]]).expand(b, 2)
        got = postprocess.estimate_focal_knowing_depth(p["pts3d_local"], pp, focal_mode="weiszfeld")
        assert torch.allclose(got, want_a, rtol=2e-5)


def test_focal_edge_cases_and_full_size(ops):
    # z = 0 / NaN pixels contribute nothing (nan_to_num), an empty selection returns the 60-degree default
    h, w, f = 368, 512, 410.0
    v, u = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    z = 2 + torch.sin(u / 50) * torch.cos(v / 40)
    pts = torch.stack([(u - w / 2) * z / f, (v - h / 2) * z / f, z], -1).reshape(1, h, w, 3).repeat(3, 1, 1, 1)
    pts[1, :50] = 0.0            # x/z = 0/0 -> NaN -> 0
    pts[1, 60:70, :, 2] = 0.0    # x/0 -> inf -> 0
    conf = torch.ones(3, h, w)
    thr = torch.tensor([1.0, 1.0, 2.0])  # view 2: nothing selected
    got = ops.focal_weiszfeld(pts.cuda(), conf.cuda(), thr.cuda(), None, iters=100).cpu()
    assert abs(float(got[0]) - f) < 1e-2 and abs(float(got[1]) - f) < 1e-2
    assert abs(float(got[2]) - max(h, w) / (2 * np.tan(np.deg2rad(30)))) < 1e-3
    want = go.focal_weiszfeld(pts[1].numpy(), (w / 2, h / 2), np.ones((h, w), bool), 100)
    assert abs(float(got[1]) - want) <= 2e-5 * want


def test_bad_arguments_fail_loudly(ops):
    conf = torch.ones(2, 16, device="cuda")
    with pytest.raises(RuntimeError, match="q must be in"):
        ops.conf_quantile(conf, 1.5)
    x = torch.zeros(2, 16, 3, device="cuda")
    with pytest.raises(RuntimeError, match="there is no CPU path"):
        ops.similarity_apply(x.cpu(), torch.zeros(2, 13))
