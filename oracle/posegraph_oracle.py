"""CPU restatement of LoopClosing::PoseGraphOptimization (reference src/loopclosing.cpp:537-646).

TEST INFRASTRUCTURE ONLY — only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.

What is restated (paths relative to /root/reference):
  * the graph: one VertexPose per keyframe, fixed = active keyframes + the loop keyframe + keyframe 0
    (:559-562); one EdgePoseGraph per (keyframe, previous keyframe) and per (keyframe, loop keyframe) with the
    stored relative pose as measurement and information I6 (:571-599); LM for 20 iterations (:605-606);
  * EdgePoseGraph::computeError (include/myslam/g2o_types.h:161-167): log(Z^-1 * T0 * T1^-1), and
    VertexPose::oplusImpl (:32-37): T <- exp(d) * T;
  * g2o (third party, absent, un-pinned master ~2020): the edge has no analytic linearizeOplus (the block at
    g2o_types.h:169-185 is commented out), so g2o's BaseBinaryEdge falls back to central differences with
    step 1e-9 through oplus/computeError; BlockSolver<6,6> without marginalisation; LinearSolverEigen (a sparse
    Cholesky — here scipy's sparse LU on the same SPD system); OptimizationAlgorithmLevenberg as in
    oracle/ba_oracle.c; Sophus SE3d::exp / log (also absent) restated from its published formulas.

PARITY UNPINNED: no reference test or golden vector exists for this path and g2o/Sophus cannot be run here.
tests/test_oracle_posegraph.py pins exp/log against scipy's matrix exponential/logarithm, the numeric
Jacobian against the closed form the reference left commented out, and checks that a drifted loop is pulled
back onto the ground truth.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def hat(w):
    W = np.zeros(w.shape[:-1] + (3, 3))
    W[..., 0, 1], W[..., 0, 2] = -w[..., 2], w[..., 1]
    W[..., 1, 0], W[..., 1, 2] = w[..., 2], -w[..., 0]
    W[..., 2, 0], W[..., 2, 1] = -w[..., 1], w[..., 0]
    return W


def quat_to_R(q):
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def R_to_quat(R):
    """-> (x, y, z, w) with w >= 0, batched (Shepperd's method)."""
    R = np.asarray(R)
    out = np.empty(R.shape[:-2] + (4,))
    flat_R, flat_q = R.reshape(-1, 3, 3), out.reshape(-1, 4)
    for k, M in enumerate(flat_R):
        tr = M[0, 0] + M[1, 1] + M[2, 2]
        if tr > 0:
            s = np.sqrt(tr + 1.0) * 2
            q = [(M[2, 1] - M[1, 2]) / s, (M[0, 2] - M[2, 0]) / s, (M[1, 0] - M[0, 1]) / s, 0.25 * s]
        elif M[0, 0] > M[1, 1] and M[0, 0] > M[2, 2]:
            s = np.sqrt(1.0 + M[0, 0] - M[1, 1] - M[2, 2]) * 2
            q = [0.25 * s, (M[0, 1] + M[1, 0]) / s, (M[0, 2] + M[2, 0]) / s, (M[2, 1] - M[1, 2]) / s]
        elif M[1, 1] > M[2, 2]:
            s = np.sqrt(1.0 + M[1, 1] - M[0, 0] - M[2, 2]) * 2
            q = [(M[0, 1] + M[1, 0]) / s, 0.25 * s, (M[1, 2] + M[2, 1]) / s, (M[0, 2] - M[2, 0]) / s]
        else:
            s = np.sqrt(1.0 + M[2, 2] - M[0, 0] - M[1, 1]) * 2
            q = [(M[0, 2] + M[2, 0]) / s, (M[1, 2] + M[2, 1]) / s, 0.25 * s, (M[1, 0] - M[0, 1]) / s]
        q = np.array(q)
        if q[3] < 0:
            q = -q
        flat_q[k] = q / np.linalg.norm(q)
    return out


def se3_from7(p):
    p = np.asarray(p, np.float64)
    return quat_to_R(p[..., :4]), p[..., 4:7].copy()


def se3_to7(R, t):
    return np.concatenate([R_to_quat(R), t], axis=-1)


def se3_mul(A, B):
    return A[0] @ B[0], (A[0] @ B[1][..., None])[..., 0] + A[1]


def se3_inv(A):
    Rt = np.swapaxes(A[0], -1, -2)
    return Rt, -(Rt @ A[1][..., None])[..., 0]


def se3_exp(d):
    """Sophus SE3d::exp([upsilon, omega]), batched."""
    d = np.asarray(d, np.float64)
    w = d[..., 3:]
    th2 = np.sum(w * w, -1)
    th = np.sqrt(th2)
    small = th < 1e-10
    ths = np.where(small, 1.0, th)
    a = np.where(small, 1 - th2 / 6, np.sin(ths) / ths)
    b = np.where(small, 0.5 - th2 / 24, (1 - np.cos(ths)) / ths ** 2)
    c = np.where(small, 1.0 / 6 - th2 / 120, (ths - np.sin(ths)) / ths ** 3)
    W = hat(w)
    W2 = W @ W
    I = np.eye(3)
    R = I + a[..., None, None] * W + b[..., None, None] * W2
    V = I + b[..., None, None] * W + c[..., None, None] * W2
    return R, (V @ d[..., :3, None])[..., 0]


def so3_log(R):
    """Sophus SO3d::log via the unit quaternion (two_atan_nbyw_by_n), batched."""
    q = R_to_quat(R)
    v, w = q[..., :3], q[..., 3]
    n2 = np.sum(v * v, -1)
    n = np.sqrt(n2)
    small = n2 < 1e-20
    ns = np.where(small, 1.0, n)
    ws = np.where(np.abs(w) < 1e-10, 1.0, w)
    two = np.where(small, 2.0 / ws - (2.0 / 3.0) * n2 / ws ** 3,
                   np.where(np.abs(w) < 1e-10, np.pi / ns, 2 * np.arctan(ns / ws) / ns))
    return two[..., None] * v


def se3_log(A):
    """Sophus SE3d::log -> [upsilon, omega], batched."""
    w = so3_log(A[0])
    th2 = np.sum(w * w, -1)
    th = np.sqrt(th2)
    W = hat(w)
    W2 = W @ W
    small = th < 1e-10
    ths = np.where(small, 1.0, th)
    half = 0.5 * ths
    coef = np.where(small, 1.0 / 12, (1 - ths * np.cos(half) / (2 * np.sin(half))) / ths ** 2)
    Vinv = np.eye(3) - 0.5 * W + coef[..., None, None] * W2
    return np.concatenate([(Vinv @ A[1][..., None])[..., 0], w], -1)


def edge_errors(R, t, v0, v1, Zinv):
    """EdgePoseGraph::computeError for all edges: log(Z^-1 * T[v0] * T[v1]^-1) -> [m, 6]."""
    T0 = (R[v0], t[v0])
    T1i = se3_inv((R[v1], t[v1]))
    return se3_log(se3_mul(Zinv, se3_mul(T0, T1i)))


def numeric_jacobians(R, t, v0, v1, Zinv, delta=1e-9):
    """g2o BaseBinaryEdge::linearizeOplus (numeric): central differences through oplus. -> Ji, Jj [m, 6, 6]."""
    m = len(v0)
    Ji, Jj = np.zeros((m, 6, 6)), np.zeros((m, 6, 6))
    for side, J in ((0, Ji), (1, Jj)):
        for d in range(6):
            e = []
            for sgn in (1.0, -1.0):
                step = np.zeros(6)
                step[d] = sgn * delta
                E = se3_exp(step)
                R0, t0, R1, t1 = R[v0], t[v0], R[v1], t[v1]
                if side == 0:
                    R0, t0 = se3_mul((E[0][None], E[1][None]), (R0, t0))
                else:
                    R1, t1 = se3_mul((E[0][None], E[1][None]), (R1, t1))
                e.append(se3_log(se3_mul(Zinv, se3_mul((R0, t0), se3_inv((R1, t1))))))
            J[:, :, d] = (e[0] - e[1]) / (2 * delta)
    return Ji, Jj


def solve(poses, fixed, v0, v1, meas, iters=20, jacobian="numeric"):
    """LoopClosing::PoseGraphOptimization's solver part.
    poses [n,7] (qx qy qz qw tx ty tz), fixed [n] bool, edges v0 -> v1 with measurement meas [m,7].
    Returns (poses [n,7], info dict)."""
    R, t = se3_from7(poses)
    fixed = np.asarray(fixed, bool)
    v0, v1 = np.asarray(v0, np.int64), np.asarray(v1, np.int64)
    Zinv = se3_inv(se3_from7(meas))
    n = len(R)
    fidx = -np.ones(n, np.int64)
    fidx[~fixed] = np.arange((~fixed).sum())
    nf = int((~fixed).sum())
    a, b = fidx[v0], fidx[v1]
    lam, ni = 0.0, 2.0
    lm_iters, trials = 0, 0
    chi_hist = []
    for it in range(iters):
        e = edge_errors(R, t, v0, v1, Zinv)
        cur_chi = float(np.sum(e * e))
        chi_hist.append(cur_chi)
        if nf == 0:
            break
        Ji, Jj = numeric_jacobians(R, t, v0, v1, Zinv)
        rows, cols, vals = [], [], []
        bvec = np.zeros(6 * nf)

        def add(pa, pb, M):
            rr, cc = np.meshgrid(np.arange(6), np.arange(6), indexing="ij")
            rows.append((6 * pa[:, None, None] + rr).ravel())
            cols.append((6 * pb[:, None, None] + cc).ravel())
            vals.append(M.ravel())

        ma, mb = a >= 0, b >= 0
        add(a[ma], a[ma], np.swapaxes(Ji[ma], 1, 2) @ Ji[ma])
        add(b[mb], b[mb], np.swapaxes(Jj[mb], 1, 2) @ Jj[mb])
        both = ma & mb
        add(a[both], b[both], np.swapaxes(Ji[both], 1, 2) @ Jj[both])
        add(b[both], a[both], np.swapaxes(Jj[both], 1, 2) @ Ji[both])
        np.add.at(bvec.reshape(nf, 6), a[ma], -(np.swapaxes(Ji[ma], 1, 2) @ e[ma][..., None])[..., 0])
        np.add.at(bvec.reshape(nf, 6), b[mb], -(np.swapaxes(Jj[mb], 1, 2) @ e[mb][..., None])[..., 0])
        H = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * nf, 6 * nf))
        if it == 0:
            lam = 1e-5 * float(np.abs(H.diagonal()).max())
            ni = 2.0
        rho, qmax = 0.0, 0
        while True:
            Rb, tb = R.copy(), t.copy()
            try:
                x = spla.splu(H + lam * sp.identity(6 * nf, format="csc")).solve(bvec)
                ok = bool(np.all(np.isfinite(x)))
            except RuntimeError:
                ok, x = False, np.zeros(6 * nf)
            if ok:
                E = se3_exp(x.reshape(nf, 6))
                free = np.nonzero(~fixed)[0]
                R[free], t[free] = se3_mul(E, (R[free], t[free]))
            e_t = edge_errors(R, t, v0, v1, Zinv)
            tmp_chi = float(np.sum(e_t * e_t)) if ok else np.finfo(float).max
            scale = float(x @ (lam * x + bvec)) + 1e-3 if ok else 1e-3
            rho = (cur_chi - tmp_chi) / scale
            trials += 1
            if rho > 0 and np.isfinite(tmp_chi):
                alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                lam *= max(1.0 / 3.0, alpha)
                ni = 2.0
                cur_chi = tmp_chi
            else:
                lam *= ni
                ni *= 2
                R, t = Rb, tb
            qmax += 1
            if not (rho < 0 and qmax < 10):
                break
        lm_iters += 1
        if qmax == 10 or rho == 0:
            break
    e = edge_errors(R, t, v0, v1, Zinv)
    return se3_to7(R, t), {"lm_iters": lm_iters, "trials": trials, "chi2": float(np.sum(e * e)), "chi2_start": chi_hist[0] if chi_hist else 0.0}
