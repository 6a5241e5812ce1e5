"""First-principles pins of the pose-graph checker (oracle/posegraph_oracle.py; "parity unpinned": the
reference has no tests and g2o / Sophus are absent).  Reference: src/loopclosing.cpp:537-646,
include/myslam/g2o_types.h:157-190."""
import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.optimize import least_squares

from oracle import posegraph_oracle as PG


def _mat4(R, t):
    M = np.eye(4)
    M[:3, :3], M[:3, 3] = R, t
    return M


def test_exp_and_log_match_matrix_functions():
    rng = np.random.default_rng(0)
    for scale in (1e-11, 1e-5, 0.1, 1.0, 2.5):
        d = rng.normal(0, scale, 6)
        R, t = PG.se3_exp(d)
        X = np.zeros((4, 4))
        X[:3, :3] = PG.hat(d[3:])
        X[:3, 3] = d[:3]
        assert np.allclose(_mat4(R, t), expm(X), atol=1e-12)
        back = PG.se3_log((R, t))
        assert np.allclose(back, d, atol=1e-9 * max(1.0, scale))
        if scale >= 1e-5:
            L = np.real(logm(_mat4(R, t)))
            assert np.allclose(back[:3], L[:3, 3], atol=1e-8) and np.allclose(PG.hat(back[3:]), L[:3, :3], atol=1e-8)


def test_quaternion_round_trip_all_branches():
    rng = np.random.default_rng(1)
    for _ in range(200):
        w = rng.normal(0, 2.0, 3)
        R, _ = PG.se3_exp(np.concatenate([np.zeros(3), w]))
        assert np.allclose(PG.quat_to_R(PG.R_to_quat(R)), R, atol=1e-12)


def test_numeric_jacobian_is_the_derivative_through_oplus(synth):
    g = synth.pose_graph(2, n=60, n_loops=3)
    R, t = PG.se3_from7(g["poses0"])
    Zinv = PG.se3_inv(PG.se3_from7(g["meas"]))
    v0, v1 = g["v0"].astype(int), g["v1"].astype(int)
    Ji, Jj = PG.numeric_jacobians(R, t, v0, v1, Zinv)
    Ji6, Jj6 = PG.numeric_jacobians(R, t, v0, v1, Zinv, delta=1e-6)
    # step 1e-9 in double with translations of a few hundred metres: ~eps * |t| / 1e-9 = 1e-4 absolute noise
    # (the reference inherits exactly this from g2o's numeric linearizeOplus)
    assert np.abs(Ji - Ji6).max() < 1e-3 and np.abs(Jj - Jj6).max() < 1e-3
    # for a zero error the derivative wrt the two vertices is +-Ad-like and cancels on a common motion
    assert np.abs(Ji).max() > 0.5


def test_fixed_vertices_stay_and_chi2_drops(synth):
    g = synth.pose_graph(3, n=200, n_loops=6)
    p, info = PG.solve(g["poses0"], g["fixed"], g["v0"], g["v1"], g["meas"])
    fx = g["fixed"] == 1
    assert np.allclose(p[fx], g["poses0"][fx], atol=1e-12)
    assert info["chi2"] < 1e-3 * info["chi2_start"]
    assert info["lm_iters"] == 20


def test_reaches_the_least_squares_minimum(synth):
    g = synth.pose_graph(4, n=40, n_loops=2, n_active=3)
    p, info = PG.solve(g["poses0"], g["fixed"], g["v0"], g["v1"], g["meas"], iters=60)
    free = np.nonzero(g["fixed"] == 0)[0]
    R, t = PG.se3_from7(p)
    Zinv = PG.se3_inv(PG.se3_from7(g["meas"]))
    v0, v1 = g["v0"].astype(int), g["v1"].astype(int)

    def resid(z):
        Rz, tz = R.copy(), t.copy()
        E = PG.se3_exp(z.reshape(-1, 6))
        Rz[free], tz[free] = PG.se3_mul(E, (R[free], t[free]))
        return PG.edge_errors(Rz, tz, v0, v1, Zinv).ravel()

    sol = least_squares(resid, np.zeros(6 * len(free)), method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=50)
    c = float(sol.fun @ sol.fun)
    assert c <= info["chi2"] * (1 + 1e-9) + 1e-15
    assert info["chi2"] - c < 1e-6 * max(info["chi2"], 1e-12), (info["chi2"], c)


def test_loop_closure_pulls_the_drifted_trajectory_back(synth):
    g = synth.pose_graph(5, n=300, n_loops=8)
    p, info = PG.solve(g["poses0"], g["fixed"], g["v0"], g["v1"], g["meas"])

    def centres(q):
        R, t = PG.se3_from7(q)
        return -(np.swapaxes(R, 1, 2) @ t[..., None])[..., 0]

    e0 = np.linalg.norm(centres(g["poses0"]) - centres(g["poses_gt"]), axis=1)
    e1 = np.linalg.norm(centres(p) - centres(g["poses_gt"]), axis=1)
    assert e1.mean() < 0.6 * e0.mean()
