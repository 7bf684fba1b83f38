"""The geometry-tail oracle (oracle/geometry_oracle.py) against the reference's own outputs (tests/golden/geometry_tail.pt,
made by tools/make_golden_geometry.py) and against the defining properties of the similarity fit (the part of the
reference that lives in the absent third-party package roma)."""
import os

import numpy as np
import pytest
import torch

from oracle import geometry_oracle as go

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "geometry_tail.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN, weights_only=False)


def test_quantile_matches_torch(gold):
    qs = gold["quantiles"]["q"]
    rows = iter(gold["quantiles"]["values"])
    for p in gold["preds"]:
        for i in range(p["conf"].shape[0]):
            want = next(rows)
            for q, w in zip(qs, want):
                assert float(go.conf_quantile(p["conf"][i].numpy(), q)) == w, (q, w)
    # short vectors with wide gaps make the interpolation's rounding visible (ATen's lerp is one fused multiply-add)
    rng = np.random.default_rng(5)
    for _ in range(2000):
        v = (1 + np.exp(rng.standard_normal(int(rng.integers(1, 6))))).astype(np.float32)
        q = float(rng.uniform(0, 1))
        assert float(go.conf_quantile(v, q)) == float(torch.quantile(torch.from_numpy(v), q))
    # ties and tiny vectors
    v = np.array([2.0, 2.0, 2.0, 5.0, 1.0], np.float32)
    for q in (0.0, 0.25, 0.5, 0.77, 1.0):
        assert float(go.conf_quantile(v, q)) == float(torch.quantile(torch.from_numpy(v), q))


def test_focal_matches_reference(gold):
    for p, want_m, want_a in zip(gold["preds"], gold["focal_masked_p10_100it"], gold["focal_all_10it"]):
        h, w = p["conf"].shape[1:]
        for i in range(p["conf"].shape[0]):
            got = go.estimate_focal(p["pts3d_local"][i].numpy(), p["conf_local"][i].numpy())
            assert abs(got - want_m[i]) <= 2e-5 * want_m[i], (got, want_m[i])
            got = go.focal_weiszfeld(p["pts3d_local"][i].numpy(), (w / 2, h / 2), None, iters=10)
            assert abs(got - float(want_a[i])) <= 2e-5 * float(want_a[i])


def test_align_matches_reference_call_site(gold):
    for case in gold["align"]:
        for p, vm, want in zip(gold["preds"], gold["valid_masks"], case["aligned"]):
            for i in range(p["conf"].shape[0]):
                got, r, t, s = go.align_local_to_global(p["pts3d_local"][i].numpy(), p["conf"][i].numpy(),
                                                        p["pts3d_in_other_view"][i].numpy(),
                                                        None if vm is None else vm[i].numpy(), case["percentile"])
                assert np.abs(got - want[i].numpy()).max() <= 1e-5 * np.abs(want[i].numpy()).max()
    # the <3 valid pixels unit falls back to the identity
    p, vm = gold["preds"][2], gold["valid_masks"][2]
    _, r, t, s = go.align_local_to_global(p["pts3d_local"][1].numpy(), p["conf"][1].numpy(),
                                          p["pts3d_in_other_view"][1].numpy(), vm[1].numpy(), 0)
    assert np.array_equal(r, np.eye(3)) and s == 1.0 and not t.any()


def test_umeyama_recovers_a_known_similarity_and_is_optimal():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((500, 3))
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    y = 1.7 * x @ q.T + np.array([0.3, -2.0, 5.0])
    r, t, s = go.umeyama(x, y)
    assert np.allclose(r, q, atol=1e-12) and abs(s - 1.7) < 1e-12 and np.allclose(t, [0.3, -2.0, 5.0], atol=1e-12)
    # noisy: no nearby similarity does better; reflections are never returned
    y2 = y + 0.1 * rng.standard_normal(y.shape)
    r, t, s = go.umeyama(x, y2)
    assert abs(np.linalg.det(r) - 1) < 1e-12
    best = ((s * x @ r.T + t - y2) ** 2).sum()
    for _ in range(50):
        w = 1e-3 * rng.standard_normal(3)
        k = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        r2 = r @ (np.eye(3) + k + k @ k / 2)
        u, _, vt = np.linalg.svd(r2)
        r2 = u @ vt
        s2, t2 = s * (1 + 1e-3 * rng.standard_normal()), t + 1e-3 * rng.standard_normal(3)
        assert ((s2 * x @ r2.T + t2 - y2) ** 2).sum() >= best
    # a point set whose best orthogonal map is a reflection still yields a rotation
    xm = x * np.array([1.0, 1.0, -1.0])
    r, t, s = go.umeyama(x, xm)
    assert abs(np.linalg.det(r) - 1) < 1e-12
