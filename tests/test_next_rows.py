"""SURVEY §8f "next" rows beyond N1: N2 LaserOdometry::adjustDistortion (laserOdometry.cpp:557-657) and N4 the loop-closure ICP of
LaserMapping::performLoopClosure (laserMapping.cpp:652-711); the N4 part is at the end of the file.

CPU part: the oracle's literal restatement against an independent float64 numpy evaluation of the same formulas and against
properties of the function (no IMU data -> untouched, unsynchronised stamps -> partial update, constant IMU state -> identity).
GPU part: alego_lo_adjust_distortion (prefix-max + search formulation on the device) against the oracle's sequential walk.

Tolerance: the outputs are float32 coordinates produced by two float 3x3 products from interpolated IMU angles; the device
evaluates sinf / cosf in double and rounds (glibc's float routines differ from that in rare 1-ulp cases) and Eigen's evaluation
order is unpinned -> 5e-5 m absolute on points within 100 m (a few float ulps); integer outputs (points visited,
imu_ptr_last_iter_) and every untouched point bit-exact.
"""
import numpy as np
import pytest

DIST_TOL = 5e-5  # metres


def make_queue(rng, scan_time, length=200, n_msgs=120, rate=100.0, lead=0.35, wrap_at=0, motion=1.0):
    """IMU ring buffers as imuHandler leaves them (laserOdometry.cpp:761-804): n_msgs messages at `rate` Hz, the first one
    `lead` seconds before the sweep, written from ring position wrap_at.  Returns ((10, length) array, ptr_last, first index)."""
    q = np.zeros((10, length))
    t = scan_time - lead + np.arange(n_msgs) / rate + rng.uniform(0, 1e-4, n_msgs).cumsum()
    ph = rng.uniform(0, 6.28, 9)
    vals = np.stack([
        t,
        motion * 0.03 * np.sin(2.1 * t + ph[0]), motion * 0.02 * np.sin(1.7 * t + ph[1]), motion * (0.4 * (t - t[0]) + 0.05 * np.sin(3 * t + ph[2])),
        motion * 2.0 * (t - t[0]) + 0.1 * motion * np.sin(t + ph[3]), motion * 0.3 * np.sin(0.9 * t + ph[4]), motion * 0.02 * np.sin(5 * t + ph[5]),
        motion * (2.0 + 0.1 * np.cos(t + ph[6])), motion * 0.27 * np.cos(0.9 * t + ph[7]), motion * 0.1 * np.cos(5 * t + ph[8]),
    ])
    idx = (wrap_at + np.arange(n_msgs)) % length
    q[:, idx] = vals
    return q, int(idx[-1]), int(idx[0])


def numpy_adjust(cloud, col, horizon_scan, scan_time, q, ptr_last, ptr_last_iter, scan_period=0.2):
    """Independent float64 evaluation of laserOdometry.cpp:557-657 (start_ori = end_ori = 0 as for any orientation in
    (-pi, pi]); plain rotation matrices instead of Eigen's quaternion route."""
    out = np.array(cloud, np.float64).copy()
    length = q.shape[1]

    def rot(r, p, y):
        cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
        return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                         [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                         [-sp, cp * sr, cp * cr]])
    it, visited = ptr_last_iter, 0
    start = None
    for i in range(len(out)):
        rel = col[i] * scan_period / horizon_scan
        cur = scan_time + rel
        if ptr_last <= 0:
            continue
        f = it
        while f != ptr_last:
            if cur < q[0, f]:
                break
            f = (f + 1) % length
        if abs(cur - q[0, f]) > scan_period:
            break
        if cur > q[0, f]:
            s = q[1:, f]
        else:
            b = (f - 1 + length) % length
            rf = (cur - q[0, b]) / (q[0, f] - q[0, b])
            s = q[1:, f] * rf + q[1:, b] * (1 - rf)
        Rc = rot(*s[0:3])
        if i == 0:
            start = (np.linalg.inv(Rc), s[3:6].copy(), s[6:9].copy())
        else:
            sh = s[3:6] - start[1] - start[2] * rel
            out[i, :3] = start[0] @ (Rc @ out[i, :3] + sh)
        it = f
        visited += 1
    return out.astype(np.float32), visited, it


def random_cloud(rng, n_rings=8, horizon_scan=1800, keep=0.6, max_col_frac=0.85):
    """Ring-major cloud with ascending columns inside each ring, like ImageProjection's compaction (imageProjection.cpp:158-191).
    max_col_frac < 0.95: see test_oracle_distortion_ring_restart."""
    pts, cols = [], []
    for r in range(n_rings):
        c = np.nonzero(rng.uniform(size=horizon_scan) < keep)[0]
        c = c[c < max_col_frac * horizon_scan]
        az = -(c + 0.5) * (2 * np.pi / horizon_scan)
        rg = rng.uniform(2.0, 90.0, len(c))
        el = np.deg2rad(-15 + 2.0 * r)
        pts.append(np.stack([rg * np.cos(el) * np.cos(az), rg * np.cos(el) * np.sin(az), rg * np.sin(el), r + c / 10000.0], 1))
        cols.append(c)
    return np.concatenate(pts).astype(np.float32), np.concatenate(cols).astype(np.int32)


# ------------------------------------------------------------------------------------------------------------------------
# CPU: oracle restatement
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wrap_at", [0, 150])
def test_oracle_distortion_matches_numpy(ob, wrap_at):
    rng = np.random.default_rng(5 + wrap_at)
    cloud, col = random_cloud(rng, n_rings=3, horizon_scan=600, keep=0.5)
    q, last, first = make_queue(rng, 100.0, wrap_at=wrap_at)
    got, n, it = ob.adjust_distortion(cloud, col, 0.3, 0.3 + 2 * np.pi, 600, 100.0, q, last, first)
    want, n2, it2 = numpy_adjust(cloud, col, 600, 100.0, q, last, first)
    assert (n, it) == (n2, it2) and n == len(cloud)
    assert np.allclose(got[:, :3], want[:, :3], rtol=0, atol=DIST_TOL)
    assert np.array_equal(got[:, 3], cloud[:, 3])  # intensity (ring + col/10000) untouched
    assert np.array_equal(got[0], cloud[0])        # point 0 only fixes the start pose (:633-639)
    assert np.abs(got[:, :3] - cloud[:, :3]).max() > 0.05  # the motion is really applied


def test_oracle_distortion_edge_cases(ob):
    rng = np.random.default_rng(9)
    cloud, col = random_cloud(rng, n_rings=2, horizon_scan=600)
    q, last, first = make_queue(rng, 50.0)
    # fewer than two IMU messages: imu_ptr_last_ <= 0, nothing happens (:583)
    for pl in (-1, 0):
        got, n, it = ob.adjust_distortion(cloud, col, 0.0, 0.0, 600, 50.0, q, pl, 0)
        assert n == 0 and it == 0 and np.array_equal(got, cloud)
    # the queue ends early: points later than last stamp + scan_period are "unsync" -> return with a partial update (:596-600)
    qs, last_s, first_s = make_queue(rng, 50.0, n_msgs=20, lead=0.29)  # last stamp ~ scan_time - 0.1: columns past C/2 are unsync
    got, n, it = ob.adjust_distortion(cloud, col, 0.0, 0.0, 600, 50.0, qs, last_s, first_s)
    want, n2, it2 = numpy_adjust(cloud, col, 600, 50.0, qs, last_s, first_s)
    assert 0 < n < len(cloud) and (n, it) == (n2, it2)
    assert np.array_equal(got[n:], cloud[n:]) and np.allclose(got[:n, :3], want[:n, :3], rtol=0, atol=DIST_TOL)
    # a motionless IMU leaves the cloud where it is (up to float rounding of R^-1 R p)
    q0, last0, first0 = make_queue(rng, 50.0, motion=0.0)
    got, n, _ = ob.adjust_distortion(cloud, col, 0.0, 0.0, 600, 50.0, q0, last0, first0)
    assert n == len(cloud) and np.allclose(got, cloud, rtol=0, atol=2e-5)
    # the pointer never moves backwards: the second ring restarts at column 0 but keeps the far pointer (:587-595, :656)
    got, n, it = ob.adjust_distortion(cloud, col, 0.0, 0.0, 600, 50.0, q, last, first)
    t_max = 50.0 + col.max() * 0.2 / 600
    assert n == len(cloud) and (q[0, it] > t_max or it == last)


def test_oracle_distortion_ring_restart(ob):
    """What the reference's function does on a full 360-degree ring-major cloud (one reason the call is commented out, :115):
    after ring 0 the forward-only pointer sits past the last column's stamp, the first point of ring 1 is a whole scan_period
    earlier, |cur_time - imu_time_[front]| exceeds scan_period and the function returns (:596-600) — only ring 0 is adjusted."""
    rng = np.random.default_rng(21)
    cloud, col = random_cloud(rng, n_rings=4, horizon_scan=600, keep=0.9, max_col_frac=1.0)
    q, last, first = make_queue(rng, 75.0)
    got, n, it = ob.adjust_distortion(cloud, col, 0.0, 0.0, 600, 75.0, q, last, first)
    ring0 = int(np.argmax(np.diff(col) < 0)) + 1
    assert n == ring0 and np.array_equal(got[n:], cloud[n:])
    want, n2, it2 = numpy_adjust(cloud, col, 600, 75.0, q, last, first)
    assert (n, it) == (n2, it2) and np.allclose(got[:, :3], want[:, :3], rtol=0, atol=DIST_TOL)


def test_pointer_walk_equals_prefix_max_search(ob):
    """The device replaces the sequential pointer walk (:587-595) by: pointer after point i = first live queue entry whose stamp
    exceeds cur_time(max column seen so far).  Check that formulation (numpy) against the oracle's literal walk on random ring-
    major clouds, wrapped queues and repeated stamps: same stop index and same final pointer."""
    rng = np.random.default_rng(33)
    for trial in range(40):
        C_ = 600
        cloud, col = random_cloud(rng, n_rings=int(rng.integers(1, 5)), horizon_scan=C_, keep=rng.uniform(0.2, 0.9),
                                  max_col_frac=rng.choice([0.5, 0.85, 1.0]))
        wrap_at = int(rng.integers(0, 200))
        q, last, first = make_queue(rng, 10.0 + trial, n_msgs=int(rng.integers(5, 150)), lead=rng.uniform(0.0, 0.5), wrap_at=wrap_at,
                                    rate=rng.choice([50.0, 100.0, 400.0]))
        if trial % 4 == 0:  # repeated stamps
            live = (first + np.arange((last - first) % 200 + 1)) % 200
            q[0, live[1::2]] = q[0, live[0:-1:2]][:len(live[1::2])]
        it0 = (first + int(rng.integers(0, 3))) % 200 if (last - first) % 200 > 3 else first
        sp = 0.2
        _, n, it = ob.adjust_distortion(cloud, col, 0.0, 0.0, C_, 10.0 + trial, q, last, it0, scan_period=sp)
        if last <= 0:  # imu_ptr_last_ <= 0 (here: the ring wrapped exactly onto slot 0): nothing is visited (:583)
            assert n == 0 and it == it0
            continue
        # parallel formulation
        live_n = (last - it0) % 200
        live_t = q[0, (it0 + np.arange(live_n)) % 200]
        pm = np.maximum.accumulate(col)
        t_max = (10.0 + trial) + pm * sp / C_
        k = np.array([np.argmax(t < live_t) if (t < live_t).any() else live_n for t in t_max])
        front = (it0 + k) % 200
        cur = (10.0 + trial) + col * sp / C_
        bad = np.abs(cur - q[0, front]) > sp
        stop = int(np.argmax(bad)) if bad.any() else len(col)
        assert stop == n, (trial, stop, n)
        assert (front[stop - 1] if stop else it0) == it, trial


def test_host_imu_queue_matches_imu_handler(alego):
    """alego::ImuQueue (host shell) against a numpy restatement of imuHandler (laserOdometry.cpp:761-804): roll / pitch / yaw of the
    orientation (scipy), gravity removal, rotation into the world frame, dead-reckoned velocity and shift, ring-buffer pointers
    (wrap-around, the front pointer pushed ahead of the writer, no integration across gaps >= 1 s)."""
    import ctypes as C
    from scipy.spatial.transform import Rotation
    L = C.CDLL(alego.HOST_PATH)
    L.alego_host_imu_create.restype = C.c_void_p
    L.alego_host_imu_create.argtypes = [C.c_int]
    L.alego_host_imu_destroy.argtypes = [C.c_void_p]
    L.alego_host_imu_push.argtypes = [C.c_void_p, C.c_void_p]
    L.alego_host_imu_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(2)
    length, n = 50, 130  # wraps twice
    q = L.alego_host_imu_create(length)
    ref = np.zeros((10, length))
    last, front = -1, 0
    t = 5.0
    for k in range(n):
        t += 2.5 if k == 60 else 0.01  # one gap: integration restarts from the stale neighbour values without a step
        rpy = np.array([0.05 * np.sin(t), 0.04 * np.cos(1.3 * t), 0.3 * t])
        quat = Rotation.from_euler("ZYX", rpy[::-1]).as_quat()  # x, y, z, w
        a = rng.normal(0, 0.5, 3) + Rotation.from_quat(quat).inv().apply([0, 0, 9.81])
        msg = np.array([t, *quat, *a])
        L.alego_host_imu_push(q, msg.ctypes.data)
        # numpy restatement
        r_, p_, y_ = rpy
        acc = np.array([a[0] + 9.81 * np.sin(p_), a[1] - 9.81 * np.cos(p_) * np.sin(r_), a[2] - 9.81 * np.cos(p_) * np.cos(r_)])
        last = (last + 1) % length
        if (last + 1) % length == front:
            front = (front + 1) % length
        ref[0:4, last] = [t, r_, p_, y_]
        accw = Rotation.from_quat(quat).apply(acc)
        back = (last - 1 + length) % length
        dt = ref[0, last] - ref[0, back]
        if dt < 1.0:
            ref[4:7, last] = ref[4:7, back] + ref[7:10, back] * dt + accw * dt * dt * 0.5
            ref[7:10, last] = ref[7:10, back] + accw * dt
    got = np.zeros((10, length))
    ptrs = np.zeros(3, np.int32)
    L.alego_host_imu_get(q, got.ctypes.data, ptrs.ctypes.data)
    L.alego_host_imu_destroy(q)
    assert list(ptrs) == [front, last, 0]
    assert np.array_equal(got[0], ref[0])
    assert np.allclose(got[1:4], ref[1:4], rtol=0, atol=1e-9)      # rpy in double
    assert np.allclose(got[4:10], ref[4:10], rtol=0, atol=2e-5)    # the world-frame acceleration is a float product (:784-785)


# ------------------------------------------------------------------------------------------------------------------------
# GPU: alego_lo_adjust_distortion against the oracle
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("preset", [0, 1])
def test_gpu_adjust_distortion(alego, ob, preset):
    P = alego.default_params(preset)
    seeds = [0, 1, 2, 3]
    B = len(seeds)
    scans = []
    for s in seeds:
        w = alego.SynthWorld(seed=s)
        sc = w.render(P, alego.trajectory_pose(0, seed=s), noise_seed=7000 + s)
        if s in (0, 1):  # a 300-degree sweep: every ring is visited (test_oracle_distortion_ring_restart explains the full one)
            h = np.degrees(-np.arctan2(sc[:, 1], sc[:, 0]) + 2 * np.pi) % 360.0
            sc = sc[h < 300.0]
        scans.append(sc)
    g = alego.Alego(P, n_seq=B)
    buf, n = g.pack_scans(scans)
    g.ip_process(buf, n)
    before = [g.ip_get(b) for b in range(B)]
    rng = np.random.default_rng(11)
    t0 = np.array([10.0, 20.5, 31.25, 47.0])
    # seq 0: plain; seq 1: ring buffer wrapped; seq 2: full sweep (returns at the first ring restart); seq 3: no IMU data yet
    qa = [make_queue(rng, t0[0]), make_queue(rng, t0[1], wrap_at=130), make_queue(rng, t0[2]), make_queue(rng, t0[3])]
    queues = [x[0] for x in qa]
    ptr_last = [qa[0][1], qa[1][1], qa[2][1], 0]
    ptr_iter = [qa[0][2], qa[1][2], qa[2][2], 0]
    n_adj, it_new = g.lo_adjust_distortion(t0, queues, ptr_last, ptr_iter, scan_period=0.2)
    bit_equal = total = 0
    for b in range(B):
        info = before[b]
        want, nv, it = ob.adjust_distortion(info["segmented_cloud"], info["segmentedCloudColInd"], info["startOrientation"],
                                            info["endOrientation"], P.horizon_scan, t0[b], queues[b], ptr_last[b], ptr_iter[b],
                                            scan_period=0.2)
        got = g.ip_get(b)["segmented_cloud"]
        M = len(want)
        assert M > 1000
        assert n_adj[b] == nv and it_new[b] == it, (b, n_adj[b], nv, it_new[b], it)
        assert np.array_equal(got[nv:], info["segmented_cloud"][nv:])          # untouched tail, bit for bit
        assert np.array_equal(got[:, 3], info["segmented_cloud"][:, 3])
        assert np.allclose(got[:, :3], want[:, :3], rtol=0, atol=DIST_TOL), np.abs(got[:, :3] - want[:, :3]).max()
        bit_equal += int((got[:nv] == want[:nv]).all(axis=1).sum())
        total += nv
    assert n_adj[0] == len(before[0]["segmented_cloud"]) and n_adj[1] == len(before[1]["segmented_cloud"])
    assert 0 < n_adj[2] < len(before[2]["segmented_cloud"]) and n_adj[3] == 0
    assert bit_equal >= 0.98 * total, (bit_equal, total)  # identical formulas: differences only from float sin / cos
    # the feature stage runs on the adjusted cloud (curvature uses the ranges of /seg_info, laserOdometry.cpp:122-129)
    g.lo_extract()
    f = g.lo_get_features(0)
    assert len(f["sharp_idx"]) > 0
    # stamps running backwards over the live entries are rejected loudly, not silently mis-searched
    bad = queues[0].copy()
    bad[0, ptr_iter[0] + 5] -= 1.0
    g.ip_process(buf, n)
    with pytest.raises(alego.AlegoError):
        g.lo_adjust_distortion(t0, [bad] + queues[1:], ptr_last, ptr_iter)
    g.close()


# ========================================================================================================================
# N4: loop-closure ICP (pcl::IterativeClosestPoint as configured at laserMapping.cpp:667-671)
#
# Tolerances: ICP stops on thresholds (|t|^2 <= 1e-6, relative MSE change < 1e-6), so two evaluations that differ in the last
# float bits may stop one iteration apart; the last steps are <= 1e-3 m by construction.  Hence: same iteration count ->
# transforms within 2e-5 (float rounding of a 4x4 chain); one iteration apart -> within 2e-3 m / 2e-4.  The oracle's literal
# float reductions (exact_sums=False, Eigen::umeyama's float sums over ~10^4 points of magnitude 10..50 m) carry ~1e-4 m of
# summation noise themselves, which bounds what any comparison against the real PCL could show.
# ========================================================================================================================
ICP_T_TIGHT, ICP_T_LOOSE = 2e-5, 2e-3
ICP_T_FLOAT_SUMS = 3e-4  # against the literal float reductions


def icp_scene(rng, n, noise=0.01, ext=30.0):
    """ground + two walls + a few poles within +-ext metres"""
    a = rng.uniform(-ext, ext, (n, 2))
    g = np.c_[a, rng.normal(-1.7, noise, n)]
    w1 = np.c_[rng.uniform(-ext, ext, n), 0.4 * ext + rng.normal(0, noise, n), rng.uniform(-1.7, 4, n)]
    w2 = np.c_[-0.5 * ext + rng.normal(0, noise, n), rng.uniform(-ext, ext, n), rng.uniform(-1.7, 4, n)]
    poles = []
    for _ in range(6):
        c = rng.uniform(-0.8 * ext, 0.8 * ext, 2)
        poles.append(np.c_[c[0] + rng.normal(0, 0.03, n // 20), c[1] + rng.normal(0, 0.03, n // 20), rng.uniform(-1.7, 3, n // 20)])
    p = np.concatenate([g, w1, w2] + poles)
    return np.c_[p, np.zeros(len(p))].astype(np.float32)


def misalign(cloud, ypr, t):
    """source such that the aligning transform is (R(ypr), t): src = R^T (p - t)"""
    from scipy.spatial.transform import Rotation
    R = Rotation.from_euler("ZYX", ypr).as_matrix()
    out = cloud.copy()
    out[:, :3] = ((cloud[:, :3].astype(np.float64) - t) @ R).astype(np.float32)
    return out, R


def scipy_icp(src, tgt, max_corr_dist=100.0, max_iterations=100, eps_t=1e-6, eps_f=1e-6):
    """Independent float64 ICP with PCL's control flow: cKDTree 1-NN, Kabsch through numpy's SVD, the same stop rules."""
    from scipy.spatial import cKDTree
    tree = cKDTree(tgt[:, :3].astype(np.float64))
    cur = src[:, :3].astype(np.float64).copy()
    Tf = np.eye(4)
    prev, it, state = np.finfo(np.float64).max, 0, 0
    while state == 0:
        d, j = tree.query(cur)
        keep = d * d <= max_corr_dist ** 2
        if keep.sum() < 3:
            state = 5
            break
        s, t = cur[keep], tgt[j[keep], :3].astype(np.float64)
        sm, tm = s.mean(0), t.mean(0)
        U, _, Vt = np.linalg.svd((t - tm).T @ (s - sm) / len(s))
        S = np.diag([1, 1, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
        R = U @ S @ Vt
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, tm - R @ sm
        cur = cur @ R.T + T[:3, 3]
        Tf = T @ Tf
        mse = float((d[keep] ** 2).mean())
        it += 1
        if it >= max_iterations:
            state = 1
        elif 0.5 * (np.trace(R) - 1) >= 1 - eps_t and (T[:3, 3] ** 2).sum() <= eps_t:
            state = 2
        elif abs(mse - prev) < 1e-12:
            state = 3
        elif abs(mse - prev) / prev < eps_f:
            state = 4
        else:
            prev = mse
    al = src[:, :3].astype(np.float64) @ Tf[:3, :3].T + Tf[:3, 3]
    d, _ = tree.query(al)
    return {"T": Tf, "iterations": it, "state": state, "fitness": float((d * d).mean())}


def assert_icp_close(a, b, tag="", max_iter_diff=1, base_tol=None):
    """two ICP runs of the same problem: equal up to the stop-threshold effect described above.  max_iter_diff > 1 is for the
    comparison with the float-summation variant: while point-to-point ICP creeps along a surface the relative MSE change hovers
    around the 1e-6 threshold for several iterations, and 1e-4 m of summation noise decides which of them stops the loop."""
    d_it = abs(a["iterations"] - b["iterations"])
    assert d_it <= max_iter_diff, (tag, a["iterations"], b["iterations"])
    tol = (ICP_T_TIGHT if base_tol is None else base_tol) + ICP_T_LOOSE * d_it
    assert np.allclose(a["T"][:3, 3], b["T"][:3, 3], rtol=0, atol=tol), (tag, a["T"], b["T"])
    assert np.allclose(a["T"][:3, :3], b["T"][:3, :3], rtol=0, atol=tol / 10 + 1e-6), (tag, a["T"], b["T"])
    assert abs(a["fitness"] - b["fitness"]) <= 1e-3 * max(b["fitness"], 1e-6) + tol, (tag, a["fitness"], b["fitness"])


def test_oracle_icp_matches_scipy_and_recovers_motion(ob):
    rng = np.random.default_rng(4)
    tgt = icp_scene(rng, 3000)
    src, R = misalign(tgt[::3], [0.03, 0.01, -0.008], np.array([0.4, -0.25, 0.08]))  # the same surface samples, moved
    want = scipy_icp(src, tgt)
    for exact in (True, False):
        got = ob.icp(src, tgt, exact_sums=exact)
        assert got["converged"] and got["state"] in (2, 4) and got["iterations"] == len(got["trace"])
        assert abs(got["iterations"] - want["iterations"]) <= 1
        assert np.allclose(got["T"], want["T"], rtol=0, atol=ICP_T_LOOSE), (got["T"], want["T"])
        assert abs(got["fitness"] - want["fitness"]) < 1e-3 * want["fitness"] + 1e-4
        # the planted motion is recovered (point-to-point ICP on identical samples has the exact answer as its fixed point)
        assert np.allclose(got["T"][:3, 3], [0.4, -0.25, 0.08], atol=0.02) and np.allclose(got["T"][:3, :3], R, atol=2e-3)
    a, b = ob.icp(src, tgt, exact_sums=True), ob.icp(src, tgt, exact_sums=False)
    assert_icp_close(a, b, "exact vs float sums", max_iter_diff=10, base_tol=ICP_T_FLOAT_SUMS)


def test_umeyama_rotation_matches_numpy_svd(ob):
    """The rotation step of the transformation estimation (the same routine text runs on the device, lc_icp_kernels.cu): against
    numpy's SVD on random covariances, on covariances whose best orthogonal fit is a reflection (det(U) det(V) < 0 -> the smallest
    singular direction is flipped, Umeyama eq. 39-43) and on rank-2 covariances (planar correspondences)."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(12)

    def ref(sig):
        U, s, Vt = np.linalg.svd(sig.astype(np.float64))
        S = np.diag([1.0, 1.0, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
        return U @ S @ Vt

    for k in range(200):
        Rtrue = Rotation.from_rotvec(rng.normal(0, 1.0, 3)).as_matrix()
        pts = rng.normal(0, 1, (50, 3)) * rng.uniform(0.2, 5.0, 3)
        if k % 4 == 1:
            pts[:, 2] = 0.0                      # planar: rank-2 covariance
        sig = (Rtrue @ pts.T) @ pts / len(pts)   # sum t s^T with t = R s
        if k % 4 == 2:
            sig = sig @ np.diag([1.0, 1.0, -1.0])  # mirrored source: the unconstrained optimum is a reflection
        if k % 4 == 3:
            sig = sig + rng.normal(0, 0.05, (3, 3))  # noisy correspondences
        sig32 = sig.astype(np.float32)
        got = ob.umeyama_rotation(sig32).astype(np.float64)
        assert np.abs(got @ got.T - np.eye(3)).max() < 1e-5 and abs(np.linalg.det(got) - 1) < 1e-5, k
        if k % 4 != 1:
            assert np.abs(got - ref(sig32)).max() < 2e-5, (k, got, ref(sig32))
        else:  # rank 2: unique as long as two singular values are non-zero; compare through the action on the plane
            assert np.abs(got @ pts.T - Rtrue @ pts.T).max() < 1e-4, k
        if k % 4 == 0:
            assert np.abs(got - Rtrue).max() < 2e-5, k


def test_oracle_icp_stop_rules(ob):
    rng = np.random.default_rng(6)
    tgt = icp_scene(rng, 1500)
    # identical clouds: every correspondence has distance 0 -> identity, stops on the transformation threshold at once
    r = ob.icp(tgt, tgt)
    assert r["converged"] and r["state"] == 2 and r["iterations"] == 1 and r["fitness"] == 0.0
    assert np.allclose(r["T"], np.eye(4), atol=1e-6)
    src, _ = misalign(tgt, [0.02, 0.0, 0.0], np.array([0.3, 0.1, 0.0]))
    # maximum iterations reached counts as converged (failure_after_max_iter_ = false)
    r = ob.icp(src, tgt, max_iterations=3)
    assert r["converged"] and r["state"] == 1 and r["iterations"] == 3
    # no correspondence inside the distance gate: "Not enough correspondences found", not converged, identity
    far = tgt.copy()
    far[:, 0] += 500.0
    r = ob.icp(far, tgt, max_corr_dist=1.0)
    assert not r["converged"] and r["state"] == 5 and r["iterations"] == 0 and np.array_equal(r["T"], np.eye(4, dtype=np.float32))
    # the gate is on the squared distance, inclusive (distance > max_dist_sqr is dropped, correspondence_estimation.hpp)
    r2 = ob.icp(src, tgt, max_corr_dist=0.5)
    assert r2["trace"][0, 0] < len(src)


@pytest.mark.gpu
def test_gpu_icp_matches_oracle(alego, ob):
    P = alego.default_params(0)
    g = alego.Alego(P, n_seq=2)  # the ICP buffers do not depend on the batch size of the handle
    rng = np.random.default_rng(8)
    # 1) synthetic planes + poles, moderate misalignment; a dozen source points floating 8-15 m above everything (exhaustive pass)
    tgt = icp_scene(rng, 6000)
    src, _ = misalign(icp_scene(rng, 2500), [0.025, -0.01, 0.006], np.array([0.35, 0.3, -0.05]))
    src[:12, 2] += rng.uniform(8, 15, 12).astype(np.float32)  # neighbours farther than two 2 m cell rings
    want = ob.icp(src, tgt, exact_sums=True)
    got = g.lc_icp(src, tgt)
    assert got["converged"] == want["converged"] and got["state"] == want["state"]
    assert_icp_close(got, want, "planes")
    k = min(got["iterations"], want["iterations"]) - 1
    # iteration by iteration: correspondence counts equal, mean squared distance and increments equal to float rounding
    assert np.array_equal(got["trace"][:k, 0], want["trace"][:k, 0])
    assert np.allclose(got["trace"][:k, 1], want["trace"][:k, 1], rtol=1e-5, atol=1e-9)
    assert np.allclose(got["trace"][:k, 2:], want["trace"][:k, 2:], rtol=0, atol=ICP_T_TIGHT)
    assert_icp_close(got, ob.icp(src, tgt, exact_sums=False), "planes vs float sums", max_iter_diff=10, base_tol=ICP_T_FLOAT_SUMS)
    # 2) the world of the benchmark: local map (1 m voxels) as history cloud, a rendered sweep's features as latest keyframe
    world = alego.SynthWorld(seed=5)
    corner, surf = world.make_map(4000, 40000, seed=5, radius=50.0)
    hist, _ = ob.voxel_grid(np.concatenate([corner, surf]), 1.0)
    key = np.concatenate([corner[::3], surf[::7]]).copy()
    key, _ = misalign(key, [-0.015, 0.004, 0.0], np.array([-0.2, 0.15, 0.03]))
    want = ob.icp(key, hist, exact_sums=True)
    got = g.lc_icp(key, hist)
    assert got["state"] == want["state"] and want["converged"]
    assert_icp_close(got, want, "world")
    # 3) stop rules and gates through the C ABI
    r = g.lc_icp(tgt, tgt)
    assert r["converged"] and r["state"] == 2 and r["iterations"] == 1 and r["fitness"] < 1e-10 and np.allclose(r["T"], np.eye(4), atol=1e-6)
    r = g.lc_icp(src, tgt, max_iterations=3)
    assert r["converged"] and r["state"] == 1 and r["iterations"] == 3
    far = tgt.copy()
    far[:, 0] += 500.0
    r = g.lc_icp(far, tgt, max_corr_dist=1.0)
    assert not r["converged"] and r["state"] == 5 and r["iterations"] == 0 and np.array_equal(r["T"], np.eye(4, dtype=np.float32))
    w2 = ob.icp(src, tgt, max_corr_dist=0.5)
    r2 = g.lc_icp(src, tgt, max_corr_dist=0.5)
    assert r2["trace"][0, 0] == w2["trace"][0, 0] and r2["state"] == w2["state"]
    assert_icp_close(r2, w2, "gated")
    # a larger cloud after a smaller one (buffers grow), then the small one again (capacity kept)
    big_t = icp_scene(rng, 30000)
    big_s, _ = misalign(icp_scene(rng, 9000), [0.01, 0.0, 0.0], np.array([0.1, -0.1, 0.0]))
    assert_icp_close(g.lc_icp(big_s, big_t), ob.icp(big_s, big_t, exact_sums=True), "big")
    assert_icp_close(g.lc_icp(src, tgt), ob.icp(src, tgt, exact_sums=True), "small again")
    g.close()


# ========================================================================================================================
# bootstrap: the first mapped frames of a run (what the nodelet shell does when no map was loaded)
# ========================================================================================================================
@pytest.mark.gpu
def test_gpu_bootstrap_from_empty_local_map(alego, ob):
    """No keyframe yet: extractSurroundingKeyFrames leaves the map clouds empty (laserMapping.cpp:196-199), the guard of
    scan2MapOptimization skips the solve (:350-354), the first keyframe is saved with the unchanged estimate (:491-545); from then
    on the local map is assembled from the stored keyframes (:206-243, alego_lm_assemble_map) and the solves run.  Device against
    the oracle driven through the same steps."""
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 6
    w = alego.SynthWorld(seed=seed)
    g = alego.Alego(P, n_seq=1)
    o = ob.Oracle(P, lm_every=1, stable_voxel=False)
    empty = np.zeros((0, 4), np.float32)
    g.lm_set_map(0, empty, empty)
    o.lm_set_map(empty, empty)
    g.pipeline_config(lm_every=1)
    kfs, poses6 = [], []
    for t in range(5):
        scan = w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=300 + t)
        buf, n = g.pack_scans([scan])
        poses = g.pipeline_step(buf, n)
        o.pipeline_step(scan)
        rep, orep = g.solve_report("lm", 0), o.report("lm")
        assert rep["status"] == orep["status"] and rep["iterations"] == orep["iterations"], (t, rep, orep)
        assert rep["n_corner"] == orep["n_corner"] and rep["n_surf"] == orep["n_surf"], (t, rep, orep)
        if t == 0:
            assert rep["status"] == alego.FEW_FEATURES and rep["iterations"] == 0 and np.all(poses[0, 3:9] == 0.0)
        else:
            assert rep["status"] == alego.OK and rep["iterations"] > 0
        assert np.abs(poses[0, 3:9] - o.get("lm_params")).max() < 1e-4, t   # north_star pose tolerance
        ds = g.lm_get_downsampled(0)
        for a, name in zip(ds, ("lm_corner_ds", "lm_surf_ds", "lm_outlier_ds")):
            assert np.array_equal(a, o.get(name)), (t, name)
        kfs.append(ds)
        poses6.append(np.asarray(o.get("lm_params"), np.float32))  # the same keyframe poses on both sides
        ck, sk, ok_ = [k[0] for k in kfs], [k[1] for k in kfs], [k[2] for k in kfs]
        g.lm_assemble_map(0, ck, sk, ok_, np.stack(poses6))
        cm, sm, _ = ob.lm_assemble_map(ck, sk, ok_, np.stack(poses6), P.lm_corner_leaf, P.lm_surf_leaf, stable=False)
        o.lm_set_map(cm, sm)
        gc, gs = g.lm_get_map(0)
        assert np.array_equal(gc, cm) and np.array_equal(gs, sm), t
    assert abs(poses[0, 3] - alego.trajectory_pose(4, seed=seed)[0]) < 0.3  # sanity: it tracks the trajectory
    g.close()
