"""Pins the oracle's restated third-party small-matrix routines against what IS in the container
(cv2 4.13 runs OpenCV's own JacobiImpl_/QR32f/LU32f; numpy f64 for the plane fit).  SURVEY.md Appendix A."""
import numpy as np
import pytest

import oracle

cv2 = pytest.importorskip("cv2")


def _cov3(rng):
    pts = rng.normal(size=(5, 3)).astype(np.float32) * rng.uniform(0.01, 1.0)
    pts[:, 0] *= rng.uniform(0.1, 20.0)
    c = pts.mean(0)
    d = pts - c
    A = (d.T @ d / 5).astype(np.float32)
    return ((A + A.T) * np.float32(0.5)).astype(np.float32)


def _jtj6(rng, n=400):
    J = rng.normal(size=(n, 6)).astype(np.float32)
    J[:, :3] *= rng.uniform(1.0, 30.0)
    A = (J.astype(np.float64).T @ J.astype(np.float64)).astype(np.float32)
    return np.ascontiguousarray((A + A.T) * np.float32(0.5))


def test_eigen3_bit_exact_vs_cv2():
    rng = np.random.default_rng(1)
    for _ in range(2000):
        A = _cov3(rng)
        ok, w_cv, v_cv = cv2.eigen(A)
        W, V = oracle.eigen_sym(A)
        assert np.array_equal(W, w_cv.reshape(-1))
        assert np.array_equal(V, v_cv)


def test_eigen6_bit_exact_vs_cv2():
    rng = np.random.default_rng(2)
    for _ in range(300):
        A = _jtj6(rng)
        ok, w_cv, v_cv = cv2.eigen(A)
        W, V = oracle.eigen_sym(A)
        assert np.array_equal(W, w_cv.reshape(-1))
        assert np.array_equal(V, v_cv)


def test_qr_solve6_bit_exact_vs_cv2():
    rng = np.random.default_rng(3)
    for _ in range(500):
        A = _jtj6(rng)
        b = rng.normal(size=(6, 1)).astype(np.float32) * 50
        ok_cv, x_cv = cv2.solve(A, b, flags=cv2.DECOMP_QR)
        ok, x = oracle.qr_solve(A, b.reshape(-1))
        assert bool(ok) == bool(ok_cv)
        assert np.array_equal(x, x_cv.reshape(-1))


def test_qr_solve6_singular_matches_cv2():
    A = np.zeros((6, 6), np.float32)
    A[0, 0] = 4.0
    b = np.ones((6, 1), np.float32)
    ok_cv, x_cv = cv2.solve(A, b, flags=cv2.DECOMP_QR)
    ok, x = oracle.qr_solve(A, b.reshape(-1))
    assert bool(ok) == bool(ok_cv)
    if not ok_cv:
        assert np.array_equal(x, np.zeros(6, np.float32))


def test_lu_invert6_bit_exact_vs_cv2():
    rng = np.random.default_rng(4)
    for _ in range(300):
        A = _jtj6(rng)
        _, _, V = cv2.eigen(A)
        ret, Vi_cv = cv2.invert(V, flags=cv2.DECOMP_LU)
        ok, Vi = oracle.lu_invert(V)
        assert np.array_equal(Vi, Vi_cv)


def test_small_matmul_is_f64_accumulate():
    rng = np.random.default_rng(5)
    for _ in range(200):
        A = rng.normal(size=(6, 6)).astype(np.float32)
        B = rng.normal(size=(6, 6)).astype(np.float32)
        want = (A.astype(np.float64) @ B.astype(np.float64)).astype(np.float32)
        assert np.array_equal(oracle.matmul_f64acc(A, B), want)
        # and OpenCV's small cv::Mat product agrees (gemm with f64 accumulation)
        got_cv = cv2.gemm(A, B, 1.0, None, 0.0)
        assert np.array_equal(got_cv, want)


def test_colpiv_householder_plane_fit_close_to_f64_lstsq():
    """Eigen is not in the container ("parity unpinned"): check the restatement solves the same problem."""
    rng = np.random.default_rng(6)
    worst = 0.0
    for _ in range(2000):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        c = rng.uniform(-15, 15, size=3)
        # 5 points spread +-0.4 m on the plane through c
        u = np.cross(n, [1, 0, 0]); u /= np.linalg.norm(u); v = np.cross(n, u)
        pts = c + rng.uniform(-0.4, 0.4, size=(5, 1)) * u + rng.uniform(-0.4, 0.4, size=(5, 1)) * v
        A = pts.astype(np.float32)
        b = -np.ones(5, np.float32)
        x = oracle.colpiv_solve_5x3(A, b)
        x64 = np.linalg.lstsq(A.astype(np.float64), b.astype(np.float64), rcond=None)[0]
        # compare the plane through a query 5 cm off-plane
        q = c + 0.05 * n
        r32 = (x.astype(np.float64) @ q + 1) / np.linalg.norm(x.astype(np.float64))
        r64 = (x64 @ q + 1) / np.linalg.norm(x64)
        worst = max(worst, abs(r32 - r64))
    assert worst < 5e-3     # f32 conditioning at 15 m; the definition is the oracle itself


def test_colpiv_householder_exact_small_integers():
    A = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [0, 1, 1]], np.float32)
    x_true = np.array([2.0, -3.0, 0.5], np.float32)
    b = A @ x_true
    x = oracle.colpiv_solve_5x3(A, b)
    assert np.allclose(x, x_true, atol=1e-5)


def test_get_transformation_roundtrip_and_convention():
    rng = np.random.default_rng(7)
    for _ in range(200):
        pose = np.concatenate([rng.uniform(-1.2, 1.2, 3), rng.uniform(-50, 50, 3)]).astype(np.float32)
        T = oracle.get_transformation(pose)
        roll, pitch, yaw = pose[:3].astype(np.float64)
        Rx = np.array([[1, 0, 0], [0, np.cos(roll), -np.sin(roll)], [0, np.sin(roll), np.cos(roll)]])
        Ry = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
        Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        assert np.allclose(T[:, :3], Rz @ Ry @ Rx, atol=1e-6)
        assert np.array_equal(T[:, 3], pose[3:])
        back = oracle.get_translation_and_euler(T)
        assert np.allclose(back, pose, atol=2e-6 * 50)
