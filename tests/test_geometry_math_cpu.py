"""fast3r_b200/csrc/geometry_math.h (the per-view solve of the similarity-fit kernel) compiled for the host and
checked against oracle/geometry_oracle.umeyama (numpy SVD) - including the reflection and degenerate cases."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import geometry_oracle as go

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "fast3r_b200", "csrc")


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("geom") / "libgeom_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC,
                           os.path.join(HERE, "geometry_math_host.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.f3r_test_umeyama_from_moments.argtypes = [C.c_void_p, C.c_void_p]
    lib.f3r_test_eig3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def moments(x, y):
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    return np.concatenate([[len(x)], x.sum(0), y.sum(0), [(x * x).sum()], (y.T @ x).reshape(-1)]).astype(np.float64)


def solve(lib, x, y):
    m = np.ascontiguousarray(moments(x, y))
    out = np.zeros(13, np.float32)
    lib.f3r_test_umeyama_from_moments(m.ctypes.data, out.ctypes.data)
    return out[:9].reshape(3, 3).astype(np.float64), out[9:12].astype(np.float64), float(out[12])


def rot(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def test_eig3(hostlib):
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.standard_normal((3, 3)) * 10 ** rng.uniform(-3, 3)
        s = np.ascontiguousarray(a.T @ a)
        vec, lam = np.zeros((3, 3)), np.zeros(3)
        hostlib.f3r_test_eig3(s.ctypes.data, vec.ctypes.data, lam.ctypes.data)
        want = np.sort(np.linalg.eigvalsh(s))[::-1]
        assert np.allclose(lam, want, rtol=1e-10, atol=1e-12 * want[0])
        assert np.allclose(vec.T @ vec, np.eye(3), atol=1e-12)
        assert np.allclose(s @ vec, vec * lam, atol=1e-10 * want[0])


def test_matches_oracle_on_random_similarities(hostlib):
    rng = np.random.default_rng(2)
    for trial in range(100):
        n = int(rng.integers(3, 400))
        x = rng.standard_normal((n, 3)) * rng.uniform(0.1, 5) + rng.standard_normal(3) * 3
        y = rng.uniform(0.2, 3) * x @ rot(rng).T + rng.standard_normal(3) * 4 + rng.uniform(0, 0.3) * rng.standard_normal((n, 3))
        r, t, s = solve(hostlib, x, y)
        r0, t0, s0 = go.umeyama(x, y)
        assert np.allclose(r, r0, atol=2e-6), trial
        assert np.allclose(t, t0, atol=2e-5 * (1 + np.abs(t0).max())), trial
        assert abs(s - s0) <= 2e-6 * s0, trial


def test_reflection_planar_and_collinear_inputs(hostlib):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((300, 3))
    # best orthogonal map is a reflection: both must return the same proper rotation and scale
    y = x * np.array([1.0, 1.0, -1.0]) + 0.01 * rng.standard_normal(x.shape)
    r, t, s = solve(hostlib, x, y)
    r0, t0, s0 = go.umeyama(x, y)
    assert abs(np.linalg.det(r) - 1) < 1e-5 and np.allclose(r, r0, atol=1e-5) and abs(s - s0) < 1e-5
    # planar point set (third singular value 0): rotation still determined
    xp = x.copy()
    xp[:, 2] = 0
    q = rot(rng)
    yp = 2.0 * xp @ q.T + 1.0
    r, t, s = solve(hostlib, xp, yp)
    assert np.allclose(r, q, atol=1e-5) and abs(s - 2.0) < 1e-5 and np.allclose(t, 1.0, atol=1e-5)
    # collinear: some rotation that maps the line correctly, finite output
    xl = np.outer(np.linspace(-1, 1, 50), [1.0, 2.0, -0.5])
    yl = 1.5 * xl @ q.T
    r, t, s = solve(hostlib, xl, yl)
    assert np.isfinite(r).all() and abs(np.linalg.det(r) - 1) < 1e-5 and abs(s - 1.5) < 1e-5
    assert np.allclose(s * xl @ r.T + t, yl, atol=1e-5)


def test_scale_extremes_and_near_isotropic_clouds(hostlib):
    """Pointmaps in millimetres or kilometres, and clouds whose covariance is nearly a multiple of the identity (all three
    singular values of M close together - the eigenvectors are then arbitrary, the rotation is not)."""
    rng = np.random.default_rng(4)
    for unit in (1e-4, 1.0, 1e4):
        for _ in range(20):
            x = rng.standard_normal((2000, 3)) * unit + rng.standard_normal(3) * 10 * unit
            q = rot(rng)
            s_true = rng.uniform(0.3, 3)
            y = s_true * x @ q.T + rng.standard_normal(3) * unit + 1e-3 * unit * rng.standard_normal(x.shape)
            r, t, s = solve(hostlib, x, y)
            r0, t0, s0 = go.umeyama(x, y)
            assert np.allclose(r, r0, atol=2e-6), unit
            assert abs(s - s0) <= 2e-6 * s0, unit
            assert np.allclose(t, t0, atol=3e-5 * unit * 30), unit
    # exactly isotropic second moments: the 8 corners of a cube
    cube = np.array([[i, j, k] for i in (-1.0, 1.0) for j in (-1.0, 1.0) for k in (-1.0, 1.0)])
    q = rot(rng)
    r, t, s = solve(hostlib, cube, 2.5 * cube @ q.T + np.array([1.0, 2.0, 3.0]))
    assert np.allclose(r, q, atol=1e-6) and abs(s - 2.5) < 1e-6 and np.allclose(t, [1, 2, 3], atol=1e-5)
