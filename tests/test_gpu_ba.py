"""Parity of the CUDA local bundle adjustment with the CPU oracle (g2o-faithful LM + Schur restatement of
Backend::OptimizeActiveMap, src/backend.cpp:126-269).  Tolerance from BASELINE.json north_star: poses and
landmark positions within 1e-4 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def rel_close(got, want):
    return np.all(np.abs(got - want) <= RTOL * np.maximum(1.0, np.abs(want)))


@pytest.fixture(scope="module")
def ba(pkg):
    b = pkg.LocalBA(max_windows=8, max_poses=7, max_points=512, max_obs=4096)
    yield b
    b.close()


def check_window(oracle, synth, w, got, **kw):
    p, x, chi2, outl, info = oracle.ba_solve(w["poses0"], w["points0"], w["fixed"], w["obs_pose"], w["obs_point"], w["uv"],
                                             synth.KITTI_K, **kw)
    gp, gx, gchi2, goutl, ginfo = got
    # outer rounds and inlier / outlier counts must agree; the LM iteration count too, except on windows that
    # converge to round-off level, where "rho == 0 -> terminate" (g2o) is decided by the last bits of chi2
    assert ginfo[0] == info[0] and np.array_equal(ginfo[2:], info[2:]), (ginfo, info)
    assert ginfo[1] == info[1] or np.median(chi2) < 1e-8, (ginfo, info)
    assert rel_close(gp, p), np.abs(gp - p).max()
    assert rel_close(gx, x), np.abs(gx - x).max()
    assert rel_close(gchi2, chi2), np.abs(gchi2 - chi2).max()
    # outlier flags may only differ where chi2 sits within the tolerance of the threshold
    diff = goutl != outl
    assert np.all(np.abs(chi2[diff] - 5.991) < 1e-3)
    return np.abs(gp - p).max(), np.abs(gx - x).max()


def test_batch_of_windows_matches_oracle(ba, oracle, synth):
    windows = [synth.ba_window(s) for s in range(8)]
    res = ba.solve(windows, synth.KITTI_K)
    worst = [check_window(oracle, synth, w, r) for w, r in zip(windows, res)]
    # fp64 on both sides: the agreement is far inside the 1e-4 bound
    assert max(a for a, _ in worst) < 1e-7 and max(b for _, b in worst) < 1e-6
    for w, (p, x, chi2, outl, info) in zip(windows, res):
        assert info[0] == 1 and info[1] == 10
        assert np.abs(p - w["poses_gt"])[:, 4:].max() < np.abs(w["poses0"] - w["poses_gt"])[:, 4:].max()


def test_ragged_windows_and_edge_cases(ba, oracle, synth):
    w_small = synth.ba_window(20, n_poses=3, n_points=40)
    w_clean = synth.ba_window(21, pix_noise=0.0, outlier_frac=0.0)
    w_allfixed = synth.ba_window(22, n_points=100)
    w_allfixed["fixed"][:] = 1                                   # pose-only: every landmark is a constraint
    w_nofixed = synth.ba_window(23, n_points=150, fixed_frac=0.0)  # pure gauge freedom, held by the damping only
    windows = [w_small, w_clean, w_allfixed, w_nofixed]
    res = ba.solve(windows, synth.KITTI_K)
    for w, r in zip(windows, res):
        check_window(oracle, synth, w, r)
    assert np.array_equal(res[2][1], w_allfixed["points0"])     # fixed landmarks never move


def test_many_outliers_trigger_more_outer_rounds(ba, oracle, synth):
    """Inlier ratio <= 0.5 after a round => the reference runs optimize(10) again, up to 5 times (:212-232)."""
    w = synth.ba_window(30, n_points=120, outlier_frac=0.7)
    res = ba.solve([w], synth.KITTI_K)
    check_window(oracle, synth, w, res[0])
    assert res[0][4][0] > 1


def test_other_solver_parameters(ba, oracle, synth):
    w = synth.ba_window(31, n_points=80)
    kw = dict(huber_delta=1.0, chi2_th=3.0, outer_max=2, inner_iters=4)
    res = ba.solve([w], synth.KITTI_K, **kw)
    check_window(oracle, synth, w, res[0], **kw)


def test_bad_input_is_reported(pkg, ba, synth):
    w = synth.ba_window(40, n_points=30)
    w["obs_point"] = w["obs_point"].copy()
    w["obs_point"][3] = 999
    with pytest.raises(pkg.SlamB200Error):
        ba.solve([w], synth.KITTI_K)


def test_a_keyframe_may_observe_a_landmark_more_than_once(ba, oracle, synth):
    """After LoopLocalFusion (src/loopclosing.cpp:478-505) two features of the current keyframe can end up on the same map
    point: g2o then simply has two edges between the same vertices.  Same result as the oracle, which treats every
    observation as its own edge; the window next to it in the batch (no duplicates) is unaffected."""
    rng = np.random.default_rng(5)
    w = synth.ba_window(41, n_points=120)
    pick = rng.choice(len(w["obs_pose"]), 25, replace=False)
    pick = np.concatenate([pick, pick[:5]])                                  # five (keyframe, landmark) pairs even three times
    w["obs_pose"] = np.concatenate([w["obs_pose"], w["obs_pose"][pick]])
    w["obs_point"] = np.concatenate([w["obs_point"], w["obs_point"][pick]])
    w["uv"] = np.concatenate([w["uv"], w["uv"][pick] + rng.normal(0, 0.7, (len(pick), 2))])
    perm = rng.permutation(len(w["obs_pose"]))                               # duplicates anywhere in the edge list
    for k in ("obs_pose", "obs_point", "uv"):
        w[k] = w[k][perm]
    plain = synth.ba_window(42, n_points=120)
    res = ba.solve([w, plain], synth.KITTI_K)
    check_window(oracle, synth, w, res[0])
    check_window(oracle, synth, plain, res[1])


def test_submit_wait_equals_solve(pkg, ba, synth):
    """sb_ba_submit / sb_ba_wait (the asynchronous back end) give bit-identical results to sb_ba_solve; a second submit
    before the wait is refused."""
    windows = [synth.ba_window(s, n_points=120) for s in (50, 51)]
    want = ba.solve(windows, synth.KITTI_K)
    ba.submit(windows, synth.KITTI_K)
    with pytest.raises(pkg.SlamB200Error):
        ba.submit(windows, synth.KITTI_K)
    got = ba.wait()
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert np.array_equal(a, b)
