"""CPU-only checks of the ORACLE (test infrastructure) against independent implementations.

The reference (jyakaranda/A-LeGO-LOAM) has no tests, fixtures or golden vectors and cannot be built here (ROS / PCL /
Ceres / Eigen absent) — PARITY IS UNPINNED BY THE REFERENCE.  What can be pinned is that every third-party piece the
oracle restates (FLANN exact k-NN, PCL VoxelGrid, Eigen 3x3 eigen-solve / 5x3 least squares, Ceres LM + Huber) agrees
with an independent implementation available in this image (scipy / numpy / finite differences), and that the four
cost functions are the reference's — including its documented Jacobian quirks.
"""
import numpy as np
import pytest


def _rot(x):
    sr, cr, sp, cp, sy, cy = np.sin(x[3]), np.cos(x[3]), np.sin(x[4]), np.cos(x[4]), np.sin(x[5]), np.cos(x[5])
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _residual_np(f, x):
    """Independent numpy restatement of the four residuals (utility.h:122-349), no Jacobians."""
    kind = int(f[0])
    cp, a, b, c, d = f[1:4], f[4:7], f[7:10], f[10:13], f[13]
    lp = _rot(x) @ cp + x[:3]
    if kind in (0, 2):
        return np.linalg.norm(np.cross(lp - a, lp - b)) / np.linalg.norm(a - b)
    if kind == 1:
        n = np.cross(a - b, a - c) ** 2  # component-wise squares (utility.h:191-193)
        return np.sqrt(np.sum((lp - a) ** 2 * n)) / np.sqrt(np.sum(n))
    return float(a @ lp + d)


def _fd_jac(f, x, h=1e-6):
    J = np.zeros(6)
    for q in range(6):
        e = np.zeros(6)
        e[q] = h
        J[q] = (_residual_np(f, x + e) - _residual_np(f, x - e)) / (2 * h)
    return J


def test_knn_matches_scipy_ckdtree(ob):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    pts = np.zeros((20000, 4), np.float32)
    pts[:, :3] = rng.uniform(-30, 30, (20000, 3)) * np.array([1, 1, 0.1])
    q = np.zeros((500, 4), np.float32)
    q[:, :3] = rng.uniform(-30, 30, (500, 3)) * np.array([1, 1, 0.1])
    idx, d = ob.knn(pts, q, 5)
    idx_b, d_b = ob.knn(pts, q, 5, brute=True)
    assert np.array_equal(idx, idx_b) and np.array_equal(d, d_b)  # kd-tree == O(n) scan, bit for bit
    dd, ii = cKDTree(pts[:, :3].astype(np.float64)).query(q[:, :3].astype(np.float64), k=5)
    # float32 vs float64 accumulation can swap near-ties: compare the distances, and the index sets where well separated
    assert np.allclose(np.sqrt(d.astype(np.float64)), dd, rtol=1e-5, atol=1e-6)
    gap = np.min(np.diff(dd, axis=1), axis=1) > 1e-4
    assert gap.sum() > 400 and np.array_equal(idx[gap], ii[gap])
    # k = 1 (LaserOdometry call sites)
    i1, d1 = ob.knn(pts, q, 1)
    assert np.array_equal(i1[:, 0], idx[:, 0]) and np.array_equal(d1[:, 0], d[:, 0])


def test_knn_distance_is_float_l2_simple(ob):
    """FLANN's L2_Simple<float>: diff = a - b; result += diff*diff, all in float32, x then y then z."""
    rng = np.random.default_rng(1)
    pts = np.zeros((64, 4), np.float32)
    pts[:, :3] = rng.uniform(-50, 50, (64, 3))
    q = np.zeros((8, 4), np.float32)
    q[:, :3] = rng.uniform(-50, 50, (8, 3))
    idx, d = ob.knn(pts, q, 3)
    for a in range(8):
        for t in range(3):
            p = pts[idx[a, t]]
            r = np.float32(0)
            for c in range(3):
                diff = np.float32(q[a, c] - p[c])
                r = np.float32(r + np.float32(diff * diff))
            assert r == d[a, t]


def test_eig3_matches_numpy_eigh(ob):
    rng = np.random.default_rng(2)
    for _ in range(200):
        pts = rng.normal(size=(5, 3)) * rng.uniform(0.01, 2.0, 3)
        zm = pts - pts.mean(0)
        A = zm.T @ zm
        w, V = ob.eig3(A)
        wn, Vn = np.linalg.eigh(A)
        assert np.allclose(w, wn, rtol=1e-10, atol=1e-13)  # ascending, like Eigen::SelfAdjointEigenSolver
        for k in range(3):
            if k == 0 or abs(wn[k] - wn[k - 1]) > 1e-6 * abs(wn[2]):
                assert abs(abs(V[:, k] @ Vn[:, k]) - 1.0) < 1e-8
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-12)
    w, V = ob.eig3(np.diag([3.0, 1.0, 2.0]))
    assert np.array_equal(w, [1.0, 2.0, 3.0])


def test_lstsq5x3_matches_numpy(ob):
    rng = np.random.default_rng(3)
    for _ in range(200):
        n0 = rng.normal(size=3)
        n0 /= np.linalg.norm(n0)
        A = rng.uniform(-20, 20, (5, 3))
        A -= np.outer(A @ n0 + rng.uniform(2, 30), n0) * 0  # keep generic; plane-like case below
        b = -np.ones(5)
        n = ob.lstsq5x3(A, b)
        nn = np.linalg.lstsq(A, b, rcond=None)[0]
        assert np.allclose(n, nn, rtol=1e-8, atol=1e-10)
    # five points close to a plane n.p + d = 0 (the LaserMapping use, laserMapping.cpp:427-452)
    for _ in range(100):
        n0 = rng.normal(size=3)
        n0 /= np.linalg.norm(n0)
        d0 = rng.uniform(1, 40)
        P = rng.uniform(-1, 1, (5, 3))
        P = P - np.outer(P @ n0 + d0, n0) + rng.normal(scale=0.01, size=(5, 3))
        n = ob.lstsq5x3(P, -np.ones(5))
        assert np.allclose(n, np.linalg.lstsq(P, -np.ones(5), rcond=None)[0], rtol=1e-7, atol=1e-9)
        assert abs(abs(n @ n0) / np.linalg.norm(n) - 1) < 1e-2


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_cost_functions_and_jacobian_quirks(ob, kind):
    """Residual == numpy restatement; Jacobian == finite differences EXCEPT for the reference's documented quirks."""
    rng = np.random.default_rng(10 + kind)
    for trial in range(50):
        x = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-0.4, 0.4, 3)])
        f = np.zeros(14)
        f[0] = kind
        f[1:4] = rng.uniform(-20, 20, 3)
        f[4:7] = rng.uniform(-20, 20, 3)
        f[7:10] = f[4:7] + rng.uniform(-2, 2, 3)
        f[10:13] = f[4:7] + rng.uniform(-2, 2, 3)
        if kind == 3:
            n = rng.normal(size=3)
            f[4:7] = n / np.linalg.norm(n)
            f[13] = rng.uniform(-5, 5)
        r, J = ob.eval_residual(f, x)
        assert abs(r - _residual_np(f, x)) < 1e-9 * max(1, abs(r))
        fd = _fd_jac(f, x)
        if kind == 0:   # CornerCostFunction: only x, y, yaw are filled (utility.h:162-167)
            assert np.array_equal(J[[2, 3, 4]], [0, 0, 0])
            assert np.allclose(J[[0, 1, 5]], fd[[0, 1, 5]], rtol=1e-5, atol=1e-6)
        elif kind == 1:  # SurfCostFunction: only z, with an extra 1/k (utility.h:199-203,228)
            n = np.cross(f[4:7] - f[7:10], f[4:7] - f[10:13]) ** 2
            k = np.sqrt(n.sum())
            assert np.array_equal(J[[0, 1, 3, 4, 5]], [0, 0, 0, 0, 0])
            assert np.isclose(J[2] * k, fd[2], rtol=1e-5, atol=1e-7)
        else:           # LidarEdge / LidarPlane: full Jacobian; pitch column carries the cr*sr*cp typo (utility.h:273,325)
            assert np.allclose(J[[0, 1, 2, 3, 5]], fd[[0, 1, 2, 3, 5]], rtol=1e-5, atol=1e-6)
            # the typo replaces sy*cp*cr*z by cr*sr*cp*z in d(lp.y)/d(pitch): undo it and the column matches
            sr, cr, cp_, sy = np.sin(x[3]), np.cos(x[3]), np.cos(x[4]), np.sin(x[5])
            if kind == 3:
                gy = f[5]
            else:
                lp = _rot(x) @ f[1:4] + x[:3]
                a, b = f[4:7], f[7:10]
                cr3 = np.cross(lp - a, lp - b)
                m, kk = np.linalg.norm(cr3), np.linalg.norm(a - b)
                # d r / d lp = ((b - a) x (cross)) / (m k)  → y component
                gy = np.cross(b - a, cr3)[1] / (m * kk)
                gy = -gy if not np.isclose(np.cross(b - a, cr3)[0] / (m * kk), fd[0], atol=1e-5) else gy
            fixed = J[4] + gy * (sy * cp_ * cr - cr * sr * cp_) * f[3]
            assert np.isclose(fixed, fd[4], rtol=1e-5, atol=1e-6)
    # with roll == yaw the typo vanishes and the reference Jacobian is the true one
    x = np.array([0.1, -0.2, 0.3, 0.25, 0.1, 0.25])
    if kind in (2, 3):
        r, J = ob.eval_residual(f, x)
        assert np.allclose(J, _fd_jac(f, x), rtol=1e-5, atol=1e-6)


def _synthetic_lm_problem(rng, n_edge=60, n_plane=300, x_true=(0.3, -0.2, 0.1, 0.01, -0.015, 0.03), noise=0.0):
    x_true = np.array(x_true)
    R, t = _rot(x_true), x_true[:3]
    F = []
    for _ in range(n_edge):
        c = rng.uniform(-30, 30, 3) * np.array([1, 1, 0.1])
        u = np.array([0, 0, 1.0]) + rng.normal(scale=0.05, size=3)
        u /= np.linalg.norm(u)
        w = c + u * rng.uniform(-1, 1) + rng.normal(scale=noise, size=3)     # world point on the line
        cp = R.T @ (w - t)
        F.append(np.concatenate([[2], cp, c + 0.1 * u, c - 0.1 * u, [0, 0, 0], [0]]))
    for _ in range(n_plane):
        n = rng.normal(size=3) * np.array([0.3, 0.3, 1.0]) if rng.uniform() < 0.6 else rng.normal(size=3) * np.array([1, 1, 0.05])
        n /= np.linalg.norm(n)
        d = rng.uniform(-20, 20)
        w = rng.uniform(-30, 30, 3)
        w = w - (n @ w + d) * n + rng.normal(scale=noise, size=3)
        cp = R.T @ (w - t)
        F.append(np.concatenate([[3], cp, n, [0, 0, 0], [0, 0, 0], [d]]))
    return np.array(F), x_true


def test_lm_solver_converges_like_scipy_huber(ob):
    """ceres-like LM + HuberLoss(0.1) vs scipy.optimize.least_squares(loss='huber', f_scale=0.1): same minimiser.
    (Not iteration-exact — different trust-region schedules; and the reference's pitch-column typo is O(roll-yaw).)"""
    from scipy.optimize import least_squares
    rng = np.random.default_rng(5)
    F, x_true = _synthetic_lm_problem(rng, noise=0.02)
    # 5 % gross outliers so the Huber branch is exercised
    out = rng.choice(len(F), len(F) // 20, replace=False)
    F[out, 1:4] += rng.normal(scale=1.0, size=(len(out), 3))
    x, info = ob.solve(F, np.zeros(6), 50)
    assert info["final_cost"] < info["initial_cost"] * 0.2 and info["successful"] >= 3
    res = least_squares(lambda p: np.array([_residual_np(f, p) for f in F]), np.zeros(6), loss="huber", f_scale=0.1, xtol=1e-12,
                        ftol=1e-12, gtol=1e-12)
    assert np.abs(x - res.x).max() < 2e-3, (x, res.x)
    assert np.abs(x - x_true).max() < 2e-2
    # scipy's cost uses the same Huber: 0.5 * sum rho
    assert abs(info["final_cost"] - res.cost) < 1e-3 * res.cost + 1e-6


def test_lm_solver_exact_problem_and_iteration_cap(ob):
    rng = np.random.default_rng(6)
    F, x_true = _synthetic_lm_problem(rng, noise=0.0)
    x, info = ob.solve(F, np.zeros(6), 50)
    assert np.abs(x - x_true).max() < 1e-6 and info["final_cost"] < 1e-12
    # max_num_iterations counts successful + unsuccessful steps (ceres)
    x1, info1 = ob.solve(F, np.zeros(6), 1)
    assert info1["iterations"] == 1 and info1["termination"] == 0
    x0, info0 = ob.solve(F, np.zeros(6), 0)
    assert info0["iterations"] == 0 and np.array_equal(x0, np.zeros(6))
    # zero Jacobian columns (LaserOdometry: roll/pitch never move) are held by the min_lm_diagonal clamp
    surf_only = np.zeros((40, 14))
    surf_only[:, 0] = 1
    surf_only[:, 1:4] = rng.uniform(-10, 10, (40, 3))
    surf_only[:, 4:7] = surf_only[:, 1:4] + np.array([0.3, 0.1, 0.2])
    surf_only[:, 7:10] = surf_only[:, 4:7] + rng.uniform(-1, 1, (40, 3)) * np.array([1, 1, 0.02])
    surf_only[:, 10:13] = surf_only[:, 4:7] + rng.uniform(-1, 1, (40, 3)) * np.array([1, 1, 0.02])
    xs, infos = ob.solve(surf_only, np.zeros(6), 5)
    assert np.abs(xs[[0, 1, 3, 4, 5]]).max() < 1e-9 and abs(xs[2]) > 1e-3 and infos["final_cost"] < infos["initial_cost"]


def _voxel_np(pts, leaf):
    """Independent restatement of pcl::VoxelGrid (PCL 1.8 voxel_grid.hpp) with (key, input order) summation."""
    inv = np.float32(1.0) / np.float32(leaf)
    xyz = pts[:, :3]
    mn, mx = xyz.min(0), xyz.max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = (np.floor(xyz * inv) - min_b.astype(np.float32)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.lexsort((np.arange(len(pts)), key))
    out, keys = [], []
    a = 0
    while a < len(order):
        b = a
        s = np.zeros(4, np.float32)
        while b < len(order) and key[order[b]] == key[order[a]]:
            s = (s + pts[order[b]]).astype(np.float32)
            b += 1
        out.append(s / np.float32(b - a))
        keys.append(key[order[a]])
        a = b
    return np.array(out, np.float32), np.array(keys)


@pytest.mark.parametrize("n,leaf", [(1, 0.4), (300, 0.4), (4000, 0.8), (4000, 1.0)])
def test_voxel_grid_matches_numpy_restatement(ob, n, leaf):
    rng = np.random.default_rng(n)
    pts = np.zeros((n, 4), np.float32)
    pts[:, :2] = rng.uniform(-25, 25, (n, 2))
    pts[:, 2] = rng.uniform(-2, 3, n)
    pts[:, 3] = rng.uniform(0, 16, n)
    out, keys = ob.voxel_grid(pts, leaf, stable=True)
    ref, rkeys = _voxel_np(pts, leaf)
    assert np.array_equal(keys, rkeys) and np.array_equal(out, ref)
    assert np.all(np.diff(keys.astype(np.int64)) > 0)   # one output per voxel, ascending voxel index
    pcl_order, _ = ob.voxel_grid(pts, leaf, stable=False)   # std::sort order inside a voxel: same up to float rounding
    assert pcl_order.shape == out.shape and np.allclose(pcl_order, out, rtol=0, atol=2e-5)


def test_voxel_grid_overflow_returns_input(ob):
    rng = np.random.default_rng(0)
    pts = np.zeros((100, 4), np.float32)
    pts[:, :3] = rng.uniform(-5000, 5000, (100, 3))
    out, _ = ob.voxel_grid(pts, 0.01)  # 1e6^3 cells > INT32_MAX, < INT64 overflow
    assert np.array_equal(out, pts)


def test_segment_bounds_and_label_numbering_properties(alego, ob):
    """Structural invariants of the reference's IP output that the GPU design relies on."""
    P = alego.default_params(alego.PRESET_VLP16_1800)
    w = alego.SynthWorld(seed=0)
    scan = w.render(P, (0, 0, 0, 0), noise_seed=0)
    o = ob.Oracle(P)
    assert o.ip(scan) == 0
    lab = o.get("label_mat").reshape(P.n_scan, P.horizon_scan)
    feas = np.unique(lab[(lab > 0) & (lab < 999999)])
    assert np.array_equal(feas, np.arange(1, len(feas) + 1))            # dense 1..K
    firsts = [np.argmax(lab.reshape(-1) == k) for k in feas]
    assert np.all(np.diff(firsts) > 0)                                  # numbered in raster-seed order
    s, e = o.get("startRingIndex"), o.get("endRingIndex")
    M = len(o.get("segmentedCloudColInd"))
    assert s[0] == 5 and e[-1] == M - 6 and np.all(s[1:] == e[:-1] + 11)  # ring r+1 starts where ring r ends (+5 / -5 trims)
    col = o.get("segmentedCloudColInd")
    for r in range(P.n_scan):
        a, b = s[r] - 5, e[r] + 5
        assert np.all(np.diff(col[a:b + 1]) > 0)                        # ring-major, columns ascending within a ring
    o.lo_features()
    labf = o.get("cloud_label")
    assert set(np.unique(labf)) <= {-1, 0, 1, 2}
    assert np.array_equal(np.sort(o.get("sharp_idx")), np.sort(np.nonzero(labf == 2)[0]))
    assert np.array_equal(np.sort(o.get("flat_idx")), np.sort(np.nonzero(labf == -1)[0]))


def test_local_map_assembly_restatement(ob):
    """N1 (laserMapping.cpp:194-323): keyframe matrix = Rz*Ry*Rx (float, via quaternions like Eigen), transformed keyframes are
    concatenated corner | surf+outlier per keyframe and voxel-filtered; identity poses reduce to VoxelGrid of the concatenation."""
    rng = np.random.default_rng(3)
    K = 4
    c = [rng.uniform(-20, 20, (120, 4)).astype(np.float32) for _ in range(K)]
    s = [rng.uniform(-20, 20, (700, 4)).astype(np.float32) for _ in range(K)]
    o = [rng.uniform(-20, 20, (60, 4)).astype(np.float32) for _ in range(K)]
    poses = np.zeros((K, 6), np.float32)
    cm, sm, M = ob.lm_assemble_map(c, s, o, poses)
    ref_c, _ = ob.voxel_grid(np.concatenate(c), 0.4, stable=True)
    ref_s, _ = ob.voxel_grid(np.concatenate([x for k in range(K) for x in (s[k], o[k])]), 0.8, stable=True)
    assert np.array_equal(cm, ref_c) and np.array_equal(sm, ref_s)
    assert np.array_equal(M.reshape(K, 3, 4)[:, :, :3], np.broadcast_to(np.eye(3, dtype=np.float32), (K, 3, 3)))
    # rotation against a double-precision Rz*Ry*Rx, translation passed through
    poses = rng.uniform(-1, 1, (K, 6)).astype(np.float32) * np.array([30, 30, 2, 0.05, 0.05, 3.0], np.float32)
    _, _, M = ob.lm_assemble_map(c, s, o, poses)
    for k in range(K):
        r, p, y = (float(v) for v in poses[k, 3:])
        Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
        Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
        Mk = M[k].reshape(3, 4)
        assert np.abs(Mk[:, :3] - Rz @ Ry @ Rx).max() < 5e-7
        assert np.array_equal(Mk[:, 3], poses[k, :3])
    # one keyframe, one point: the transform itself
    pt = np.array([[1.5, -2.0, 0.25, 7.0]], np.float32)
    cm, sm, M = ob.lm_assemble_map([pt], [np.zeros((0, 4), np.float32)], [np.zeros((0, 4), np.float32)], poses[:1])
    want = M[0].reshape(3, 4)[:, :3].astype(np.float64) @ pt[0, :3].astype(np.float64) + poses[0, :3]
    assert len(cm) == 1 and len(sm) == 0 and np.allclose(cm[0, :3], want, atol=1e-5) and cm[0, 3] == 7.0
