"""Oracle cross-check for the stages the north star wants BIT-EXACT: ImageProjection (range image, ground mask, segmentation labels,
compaction) and LaserOdometry's feature selection (curvature, occlusion marks, sharp / less-sharp / flat indices).

The checker here is a second restatement of the reference (src/imageProjection.cpp:49-316, src/laserOdometry.cpp:118-297) written in
numpy with a different structure from oracle/alego_oracle.cpp — segmentation by connected components of the "join" graph
(scipy.sparse.csgraph) instead of the reference's BFS, compaction by boolean masks and cumulative sums instead of the running
counter, vectorised float32 stencils — so that a transcription slip in either shows up as a difference.  PARITY STAYS UNPINNED by
the reference itself (it has no tests or vectors); this narrows the gap to "two independent readings agree".

Further down, in the same spirit: LaserMapping's association (cKDTree + eigh + lstsq), LaserOdometry's surf and corner associations
(brute force), Ceres' trust-region LM (numpy.linalg.lstsq on the damped system), the pose bookkeeping around the solves (scipy
Rotation) and the composition of downsampleCurrentScan.
"""
import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

DBL_MAX = np.finfo(np.float64).max


def numpy_image_projection(scan, P):
    R, C = P.n_scan, P.horizon_scan
    pts = np.asarray(scan, np.float32).reshape(-1, 4)
    pts = pts[np.isfinite(pts[:, :3]).all(axis=1)]  # removeNaNFromPointCloud (:59)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    # float atan2 / hypot / sqrt, widened to double, scaled (:79-88, :99)
    vert = np.arctan2(z, np.hypot(x, y)).astype(np.float64) * 180.0 / np.pi
    row = np.trunc((vert + P.ang_bottom) / P.ang_res_y + 0.5).astype(np.int64)
    hor = (-np.arctan2(y, x).astype(np.float64) + 2 * np.pi) * 180.0 / np.pi
    col = np.trunc(hor / P.ang_res_x).astype(np.int64)
    col = np.where(col >= C, col - C, col)
    ok = (row >= 0) & (row < R) & (col >= 0) & (col < C)
    rng_f = np.sqrt(x * x + y * y + z * z)  # float32, left to right
    cell = (col + row * C)[ok]
    src = np.nonzero(ok)[0]
    # the last point written to a cell stays (:103)
    last = {}
    for c, s in zip(cell.tolist(), src.tolist()):
        last[c] = s
    cells = np.fromiter(last.keys(), np.int64, len(last))
    winners = np.fromiter(last.values(), np.int64, len(last))
    range_mat = np.full(R * C, DBL_MAX)
    range_mat[cells] = rng_f[winners].astype(np.float64)
    cloud = np.zeros((R * C, 4), np.float32)
    cloud[:, 3] = -1.0  # "no point" marker (:24-33)
    cloud[cells, :3] = pts[winners, :3]
    cloud[cells, 3] = (row[winners] + col[winners] / 10000.0).astype(np.float32)
    valid = (range_mat != DBL_MAX).reshape(R, C)
    range_mat = range_mat.reshape(R, C)
    xyz = cloud[:, :3].reshape(R, C, 3)
    # ground (:106-132): rows below ground_scan_id against the row above, both cells present
    g = P.ground_scan_id
    d = (xyz[1:g + 1] - xyz[0:g]).astype(np.float64)  # float differences widened to double
    ang = np.degrees(np.arctan2(d[..., 2], np.hypot(d[..., 0], d[..., 1])))
    hit = valid[0:g] & valid[1:g + 1] & (np.abs(ang - P.sensor_mount_ang) < 10.0)
    ground = np.zeros((R, C), bool)
    ground[0:g] |= hit
    ground[1:g + 1] |= hit
    # segmentation (:134-156, :210-316) as connected components of the join graph over the candidate cells
    cand = valid & ~ground
    idx = np.arange(R * C).reshape(R, C)

    def join(r_a, r_b, alpha):
        d1, d2 = np.maximum(r_a, r_b), np.minimum(r_a, r_b)
        return np.arctan2(d2 * np.sin(alpha), d1 - d2 * np.cos(alpha)) > P.seg_theta

    ax, ay = np.radians(P.ang_res_x), np.radians(P.ang_res_y)
    with np.errstate(over="ignore", invalid="ignore"):
        right = np.roll(range_mat, -1, axis=1)  # column index wraps (:241-248)
        e_h = cand & np.roll(cand, -1, axis=1) & join(range_mat, right, ax)
        e_v = cand[:-1] & cand[1:] & join(range_mat[:-1], range_mat[1:], ay)  # rows do not wrap (:237)
    a = np.concatenate([idx[e_h], idx[:-1][e_v]])
    b = np.concatenate([np.roll(idx, -1, axis=1)[e_h], idx[1:][e_v]])
    n_comp, comp = connected_components(coo_matrix((np.ones(len(a)), (a, b)), shape=(R * C, R * C)), directed=False)
    comp = comp.reshape(R, C)
    label = np.where(cand, 0, -1).astype(np.int32)
    cc = comp[cand]
    size = np.bincount(cc, minlength=n_comp)
    rows_of = np.zeros((n_comp, R), bool)
    rows_of[cc, np.nonzero(cand)[0]] = True
    feasible = (size >= P.seg_min_cluster) | ((size >= P.seg_valid_point_num) & (rows_of.sum(1) >= P.seg_valid_line_num))
    first = np.full(n_comp, R * C, np.int64)
    np.minimum.at(first, cc, idx[cand])
    # labels count the feasible components in the order the raster scan meets them (:303-306)
    order = np.argsort(first[feasible], kind="stable")
    number = np.zeros(n_comp, np.int32)
    number[np.nonzero(feasible)[0][order]] = np.arange(1, feasible.sum() + 1)
    label[cand] = np.where(feasible[cc], number[cc], 999999)
    # compaction (:158-191)
    cols = np.tile(np.arange(C), (R, 1))
    rows = np.repeat(np.arange(R), C).reshape(R, C)
    keep = ((label > 0) | ground) & (label != 999999)
    keep &= ~(ground & (cols % 5 != 0) & (cols > 4) & (cols < C - 5))
    outl = ((label > 0) | ground) & (label == 999999) & (rows > P.ground_scan_id) & (cols % 5 == 0)
    through = np.cumsum(keep.sum(1))
    return {
        "range_mat": range_mat.reshape(-1), "ground_mat": ground.reshape(-1).astype(np.uint8), "label_mat": label.reshape(-1),
        "startRingIndex": (through - keep.sum(1) + 5).astype(np.int32), "endRingIndex": (through - 1 - 5).astype(np.int32),
        "segmentedCloudGroundFlag": ground[keep].astype(np.uint8), "segmentedCloudColInd": cols[keep].astype(np.int32),
        "segmentedCloudRange": range_mat[keep].astype(np.float32), "segmented_cloud": cloud.reshape(R, C, 4)[keep],
        "outlier_cloud": cloud.reshape(R, C, 4)[outl],
    }


def numpy_features(ip, R):
    r = ip["segmentedCloudRange"]
    col = ip["segmentedCloudColInd"].astype(np.int64)
    gflag = ip["segmentedCloudGroundFlag"].astype(bool)
    M = len(r)
    i = np.arange(5, M - 5)
    # float sum, strictly left to right (:124)
    d = r[i - 5]
    for k in (-4, -3, -2, -1):
        d = d + r[i + k]
    d = d - r[i] * np.float32(10)
    for k in (1, 2, 3, 4, 5):
        d = d + r[i + k]
    curv = np.zeros(M)
    curv[i] = d.astype(np.float64) ** 2
    # occlusion marks (:131-159): every effect is "set", so the loop order does not matter
    picked = np.zeros(M, bool)
    r64 = r.astype(np.float64)
    near = np.abs(col[i] - col[i + 1]) < 10
    far_first = near & (r64[i] - r64[i + 1] > 0.5)
    far_second = near & ~far_first & (r64[i + 1] - r64[i] > 0.5)
    for k in range(-5, 1):
        picked[i[far_first] + k] = True
    for k in range(1, 6):
        picked[i[far_second] + k] = True
    lone = ~far_first & (np.abs(r64[i - 1] - r64[i]) > 0.02 * r64[i]) & (np.abs(r64[i + 1] - r64[i]) > 0.02 * r64[i])
    picked[i[lone]] = True
    occluded = picked.copy()
    label = np.zeros(M, np.int32)
    sharp, less_sharp, flat = [], [], []

    def suppress(idx):
        for sgn in (1, -1):
            for l in range(1, 6):
                if abs(col[idx + sgn * l] - col[idx + sgn * (l - 1)]) > 10:
                    break
                picked[idx + sgn * l] = True

    for ring in range(R):
        s, e = int(ip["startRingIndex"][ring]), int(ip["endRingIndex"][ring])
        for j in range(6):
            # C++ integer division truncates toward zero; s, e >= 0 whenever sp < ep can hold
            sp = int((s * (6 - j) + e * j) / 6)
            ep = int((s * (5 - j) + e * (j + 1)) / 6) - 1
            if sp >= ep:
                continue
            order = sorted(range(sp, ep + 1), key=lambda t: curv[t])
            n_pick = 0
            for idx in reversed(order):
                if not picked[idx] and curv[idx] > 0.1 and not gflag[idx]:
                    n_pick += 1
                    picked[idx] = True
                    if n_pick <= 2:
                        label[idx] = 2
                        sharp.append(idx)
                        less_sharp.append(idx)
                    elif n_pick <= 20:
                        label[idx] = 1
                        less_sharp.append(idx)
                    else:
                        break
                    suppress(idx)
            n_pick = 0
            for idx in order:
                if not picked[idx] and curv[idx] < 0.1 and gflag[idx]:
                    label[idx] = -1
                    flat.append(idx)
                    n_pick += 1
                    picked[idx] = True
                    if n_pick >= 4:
                        break
                    suppress(idx)
    return {"cloud_curvature": curv, "occluded": occluded, "cloud_neighbor_picked": picked, "cloud_label": label,
            "sharp_idx": np.array(sharp, np.int32), "less_sharp_idx": np.array(less_sharp, np.int32), "flat_idx": np.array(flat, np.int32)}


@pytest.mark.parametrize("preset,seed", [(0, 0), (0, 3), (1, 1), (2, 5), (3, 2)])
def test_oracle_ip_and_features_match_numpy_restatement(alego, ob, preset, seed):
    P = alego.default_params(preset)
    w = alego.SynthWorld(seed=seed)
    scan = w.render(P, alego.trajectory_pose(1, seed=seed), noise_seed=40 + seed)
    if seed == 3:  # NaN points, duplicated cells (the later point wins) and points outside the vertical field of view
        scan = np.concatenate([scan, scan[100:300] * np.float32(1.0001), [[np.nan, 1, 1, 0], [1, 1, 50, 0], [2, -1, -40, 0]]]).astype(np.float32)
    o = ob.Oracle(P, stable_voxel=True)
    assert o.ip(scan) == 0
    ip = numpy_image_projection(scan, P)
    for k in ("range_mat", "ground_mat", "label_mat", "startRingIndex", "endRingIndex", "segmentedCloudGroundFlag",
              "segmentedCloudColInd", "segmentedCloudRange", "segmented_cloud", "outlier_cloud"):
        got = np.asarray(o.get(k))
        assert got.shape == ip[k].shape and np.array_equal(got, ip[k]), (k, got.shape, ip[k].shape)
    lab = ip["label_mat"]
    assert (lab == 999999).sum() > 100 and lab[(lab > 0) & (lab < 999999)].max() > 50  # rejected and numbered components both occur
    if seed == 3:
        assert int(o.get("n_dup_cells")) > 50  # the duplicated points really shared cells
    o.lo_features()
    f = numpy_features(ip, P.n_scan)
    M = len(ip["segmentedCloudRange"])
    assert np.array_equal(o.get("cloud_curvature")[5:M - 5], f["cloud_curvature"][5:M - 5])
    if int(o.get("tie_sensitive")) == 0:  # std::sort's tie order is not reproducible by a stable sort; the oracle reports when it matters
        assert np.array_equal(o.get("cloud_neighbor_picked")[5:M - 5].astype(bool), f["cloud_neighbor_picked"][5:M - 5])
        assert np.array_equal(o.get("cloud_label")[5:M - 5], f["cloud_label"][5:M - 5])
        for k in ("sharp_idx", "less_sharp_idx", "flat_idx"):
            assert np.array_equal(o.get(k), f[k]), k
        assert len(f["sharp_idx"]) > 0 and len(f["flat_idx"]) > 0


def test_oracle_scan_to_map_association_matches_numpy_restatement(alego, ob):
    """LaserMapping's data association (laserMapping.cpp:371-462) restated with scipy's cKDTree, numpy's eigh and lstsq: the same
    queries produce residual blocks, with the same line points (up to the sign of the eigenvector) and plane parameters."""
    from scipy.spatial import cKDTree
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 7
    w = alego.SynthWorld(seed=seed)
    corner_map, surf_map = w.make_map(5000, 30000, seed=seed, radius=60.0)
    o = ob.Oracle(P, lm_every=1, stable_voxel=True)
    o.lm_set_map(corner_map, surf_map)
    o.pipeline_step(w.render(P, alego.trajectory_pose(0, seed=seed), noise_seed=seed))  # first sweep: map2laser is the identity
    # (the association of that first mapped sweep ran with map2odom = odom2laser = identity; the solve then moved the pose)
    resids = o.get("lm_resids")
    rep = o.report("lm")
    edge, plane = resids[resids[:, 0] == 2], resids[resids[:, 0] == 3]
    assert len(edge) == rep["n_corner"] > 20 and len(plane) == rep["n_surf"] > 200

    def five_nn(tree, pts32, q):
        _, j = tree.query(q.astype(np.float64), k=5)
        # squared L2 accumulated in float like FLANN's L2_Simple: the gate compares that value (:376, :426)
        d = np.zeros(len(q), np.float32)
        far = pts32[j[:, 4]]
        for c in range(3):
            t = q[:, c] - far[:, c]
            d = d + t * t
        return j, d

    # corner: PCA of the five neighbours, line iff the largest eigenvalue exceeds 3x the middle one (:397-403)
    cq = o.get("lm_corner_ds")[:, :3]
    j, d4 = five_nn(cKDTree(corner_map[:, :3].astype(np.float64)), corner_map[:, :3], cq)
    sel, want = [], []
    for i in np.nonzero(d4 < 1.0)[0]:
        nb = corner_map[j[i], :3].astype(np.float64)
        c = nb.sum(0) / 5.0
        lam, vec = np.linalg.eigh((nb - c).T @ (nb - c))
        if lam[2] > 3 * lam[1]:
            sel.append(i)
            want.append(np.concatenate([cq[i], c + 0.1 * vec[:, 2], c - 0.1 * vec[:, 2]]))
    assert np.array_equal(o.get("lm_corner_sel"), np.array(sel, np.int32))
    want = np.array(want)
    got = edge[:, 1:10]
    assert np.array_equal(got[:, :3], want[:, :3])
    same = np.abs(got[:, 3:9] - want[:, 3:9]).max(1)
    swapped = np.abs(got[:, 3:9] - want[:, [6, 7, 8, 3, 4, 5]]).max(1)
    assert np.minimum(same, swapped).max() < 1e-6  # eigenvector sign is free: the residual does not depend on it

    # surf: plane through the five neighbours by least squares of A n = -1, accepted iff all five lie within 0.2 m (:435-452)
    sq = o.get("lm_surf_total_ds")[:, :3]
    j, d4 = five_nn(cKDTree(surf_map[:, :3].astype(np.float64)), surf_map[:, :3], sq)
    sel, want = [], []
    for i in np.nonzero(d4 < 1.0)[0]:
        nb = surf_map[j[i], :3].astype(np.float64)
        n = np.linalg.lstsq(nb, -np.ones(5), rcond=None)[0]
        dd = 1.0 / np.linalg.norm(n)
        n = n / np.linalg.norm(n)
        if (np.abs(nb @ n + dd) > 0.2).any():
            continue
        sel.append(i)
        want.append(np.concatenate([sq[i], n, [dd]]))
    want = np.array(want)
    osel = o.get("lm_surf_sel")
    # a plane test that lands within rounding of 0.2 m may differ between two least-squares routines: allow a handful
    common = np.intersect1d(osel, np.array(sel, np.int32))
    assert len(common) >= len(osel) - 2 and len(common) >= len(sel) - 2
    got = plane[np.isin(osel, common)]
    want = want[np.isin(np.array(sel), common)]
    assert np.array_equal(got[:, 1:4], want[:, :3])
    assert np.abs(got[:, 4:7] - want[:, 3:6]).max() < 1e-7 and np.abs(got[:, 13] - want[:, 6]).max() < 1e-6


def test_oracle_scan_to_scan_surf_association_matches_numpy_restatement(alego, ob):
    """LaserOdometry's surf association (laserOdometry.cpp:337-407): 1-NN in surf_last_ (float squared distance, gate 25), then the
    walk over the neighbouring rings in storage order — best same-ring and best other-ring point, strict '<' so the first minimum
    in walk order stays.  Restated with brute-force numpy on the second sweep (params_ = 0: transformToStart is the identity)."""
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 8
    w = alego.SynthWorld(seed=seed)
    o = ob.Oracle(P, lm_every=0)
    for t in range(2):
        o.ip(w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=60 + t))
        o.lo_features()
        if t == 1:
            flat = o.get("flat")[:, :3].copy()
            target = o.get("surf_last").copy()  # the first sweep's less_flat cloud
        o.lo_scan2scan()
    corr = o.get("lo_surf_corr")
    assert len(corr) > 50
    tx = target[:, :3]
    ring = target[:, 3].astype(np.int64)  # int(intensity) (:347)
    assert (np.diff(ring) >= 0).all()     # ring-major storage, which is what makes the walk's early `break` a range
    want = []
    for j, q in enumerate(flat):
        d = np.zeros(len(tx), np.float32)
        for c in range(3):
            t_ = q[c] - tx[:, c]
            d = d + t_ * t_
        closest = int(np.argmin(d))
        if not d[closest] < 25.0:
            continue
        rc = ring[closest]
        diff = (tx - q).astype(np.float64)  # float subtraction, then pow(double, 2) (:357)
        pd = diff[:, 0] ** 2 + diff[:, 1] ** 2 + diff[:, 2] ** 2
        fwd = [k for k in range(closest + 1, len(tx)) if ring[k] <= rc + 2]
        bwd = [k for k in range(closest - 1, -1, -1) if ring[k] >= rc - 2]
        walk = np.array(fwd + bwd, np.int64)
        best = []
        for same in (True, False):
            cand = walk[(ring[walk] == rc) == same]
            cand = cand[pd[cand] < 25.0]
            best.append(int(cand[np.argmin(pd[cand])]) if len(cand) else -1)
        if best[0] >= 0 and best[1] >= 0:
            want.append([j, closest, best[0], best[1]])
    assert np.array_equal(corr, np.array(want, np.int32))


def numpy_ceres_lm(f14, x0, max_iters, a, ob):
    """Ceres' trust-region Levenberg-Marquardt (1.13/1.14 trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, corrector.cc,
    loss_function.cc; defaults of Solver::Options) restated with numpy: HuberLoss(a) corrector, Jacobi scaling, damped least squares
    through numpy.linalg.lstsq on the augmented system instead of a hand-written QR.  Residuals / Jacobians come from the oracle's
    per-residual evaluator (cross-checked separately against finite differences)."""
    def evaluate(x, want_jac=True):
        rs, Js = [], []
        for f in f14:
            r, J = ob.eval_residual(f, x)
            rs.append(r)
            Js.append(J)
        r, J = np.array(rs), np.array(Js)
        s = r * r
        out = s > a * a
        rho = np.where(out, 2 * a * np.sqrt(np.where(out, s, 1.0)) - a * a, s)
        w = np.sqrt(np.where(out, a / np.sqrt(np.where(out, s, 1.0)), 1.0))  # rho'' <= 0: residual and Jacobian scale by sqrt(rho')
        return 0.5 * rho.sum(), r * w, (J * w[:, None] if want_jac else None)

    x = np.array(x0, np.float64)
    cost, r, J = evaluate(x)
    initial = cost
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    J = J * scale
    radius, decrease, reuse, it, ok_steps, invalid = 1e4, 2.0, False, 0, 0, 0
    diag = None
    while it < max_iters and radius > 1e-32:
        it += 1
        if not reuse:
            diag = np.clip((J * J).sum(0), 1e-6, 1e32)
        D = np.sqrt(diag / radius)
        y = np.linalg.lstsq(np.vstack([J, np.diag(D)]), np.concatenate([r, np.zeros(6)]), rcond=None)[0]
        step = -y
        reuse = True
        Jd = J @ step
        model_change = -(Jd * (r + Jd / 2)).sum()
        if not model_change > 0:
            invalid += 1
            if invalid >= 5:
                break
            radius *= 0.5
            continue
        invalid = 0
        xc = x + step * scale
        cand = evaluate(xc, False)[0]
        if np.linalg.norm(x - xc) <= 1e-8 * (np.linalg.norm(x) + 1e-8):
            break
        if abs(cost - cand) <= 1e-6 * cost:
            break
        rho = (cost - cand) / model_change
        if rho > 1e-3:
            x = xc
            cost, r, Ju = evaluate(x)
            grad = Ju.T @ r
            J = Ju * scale
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease, reuse = 2.0, False
            ok_steps += 1
            if np.abs(grad).max() <= 1e-10:
                break
        else:
            radius /= decrease
            decrease *= 2.0
    return x, {"iterations": it, "initial_cost": initial, "final_cost": cost, "successful": ok_steps}


def test_oracle_lm_solver_matches_numpy_ceres_restatement(alego, ob):
    """The oracle's Ceres-like solver against the numpy restatement above on real scan-to-map problems (edge + plane blocks of a
    mapped sweep, HuberLoss(0.1), 20 iterations) from several start poses: same iteration count, same accepted steps, same
    minimiser and costs."""
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 7
    w = alego.SynthWorld(seed=seed)
    cm, sm = w.make_map(5000, 30000, seed=seed, radius=60.0)
    o = ob.Oracle(P, lm_every=1, stable_voxel=True)
    o.lm_set_map(cm, sm)
    o.pipeline_step(w.render(P, alego.trajectory_pose(0, seed=seed), noise_seed=seed))
    f14 = o.get("lm_resids")[::6].copy()  # every 6th block keeps the python loop short; still ~400 residuals of both kinds
    assert (f14[:, 0] == 2).sum() > 5 and (f14[:, 0] == 3).sum() > 100
    rng = np.random.default_rng(1)
    for trial in range(3):
        x0 = np.zeros(6) if trial == 0 else rng.normal(0, 1, 6) * np.array([0.3, 0.3, 0.1, 0.01, 0.01, 0.03])
        for iters in (20, 4):
            xo, so = ob.solve(f14, x0, iters, 0.1)
            xn, sn = numpy_ceres_lm(f14, x0, iters, 0.1, ob)
            assert so["iterations"] == sn["iterations"] and so["successful"] == sn["successful"], (trial, iters, so, sn)
            assert np.abs(xo - xn).max() < 1e-9, (trial, iters, xo, xn)
            assert abs(so["initial_cost"] - sn["initial_cost"]) <= 1e-12 * sn["initial_cost"]
            assert abs(so["final_cost"] - sn["final_cost"]) <= 1e-9 * sn["final_cost"]
            assert so["final_cost"] < so["initial_cost"]


def test_oracle_lo_solver_with_rank_deficient_jacobian_matches_numpy(alego, ob):
    """LaserOdometry's problems (laserOdometry.cpp:403-492): SurfCostFunction touches only z, CornerCostFunction only x, y and yaw, so
    the Jacobian has all-zero columns (roll, pitch — and x, y, yaw in the surf-only solve) that only Ceres' min_lm_diagonal clamp
    keeps solvable.  Same comparison as above on the residual blocks of a second sweep."""
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 8
    w = alego.SynthWorld(seed=seed)
    o = ob.Oracle(P, lm_every=0)
    for t in range(2):
        o.ip(w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=60 + t))
        o.lo_features()
        o.lo_scan2scan()
    f14 = o.get("lo_resids")
    surf, corner = f14[f14[:, 0] == 1], f14[f14[:, 0] == 0]
    assert len(surf) > 50 and len(corner) > 10
    for blocks in (surf[::2], np.concatenate([surf[::2], corner])):  # the first solve, then the joint one (:418, :492)
        xo, so = ob.solve(blocks, np.zeros(6), 5, 0.1)
        xn, sn = numpy_ceres_lm(blocks, np.zeros(6), 5, 0.1, ob)
        assert so["iterations"] == sn["iterations"] and so["successful"] == sn["successful"], (so, sn)
        assert np.abs(xo - xn).max() < 1e-9, (xo, xn)
        assert abs(xo[3]) < 1e-9 and abs(xo[4]) < 1e-9  # roll and pitch never move (up to the rounding of the damped solve)
    assert xo[2] != 0.0 and xo[0] != 0.0


def test_oracle_pose_bookkeeping_matches_scipy_rotations(alego, ob):
    """Pose integration around the solves: LaserOdometry accumulates translation and yaw only (laserOdometry.cpp:504-508);
    LaserMapping's transformUpdate turns params_ into map2laser = Rz Ry Rx, t and map2odom = map2laser o odom2laser^-1
    (laserMapping.cpp:481-489).  Recomputed from the oracle's per-sweep params with scipy's Rotation."""
    from scipy.spatial.transform import Rotation
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 3
    w = alego.SynthWorld(seed=seed)
    cm, sm = w.make_map(5000, 30000, seed=seed, radius=60.0)
    o = ob.Oracle(P, lm_every=1, stable_voxel=True)
    o.lm_set_map(cm, sm)
    t_w, r_w = np.zeros(3), np.eye(3)
    for t in range(4):
        o.pipeline_step(w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=20 + t))
        lo = o.get("lo_params")
        if t > 0:  # the first sweep only initialises the targets (:316-324)
            t_w = t_w + r_w @ lo[:3]
            r_w = r_w @ Rotation.from_euler("z", lo[5]).as_matrix()
        assert np.abs(o.get("t_w_cur") - t_w).max() < 1e-12 and np.abs(o.get("r_w_cur").reshape(3, 3) - r_w).max() < 1e-12, t
        lm = o.get("lm_params")
        r_m2l = Rotation.from_euler("ZYX", lm[5:2:-1]).as_matrix()  # Rz(yaw) Ry(pitch) Rx(roll)
        assert np.abs(o.get("r_map2laser").reshape(3, 3) - r_m2l).max() < 1e-12 and np.abs(o.get("t_map2laser") - lm[:3]).max() < 1e-12
        r_m2o = r_m2l @ r_w.T
        assert np.abs(o.get("r_map2odom").reshape(3, 3) - r_m2o).max() < 1e-12
        assert np.abs(o.get("t_map2odom") - (lm[:3] - r_m2o @ t_w)).max() < 1e-12
    assert np.linalg.norm(t_w) > 0.2  # the sequence really moved


def test_oracle_scan_to_scan_corner_association_matches_numpy_restatement(alego, ob):
    """LaserOdometry's corner association (laserOdometry.cpp:427-481) after the surf solve has moved params_: transformToStart
    (:728-740, double rotation + translation, float result), 1-NN in corner_last_, then the best point on a HIGHER ring walking
    forward (at most two rings up) or on a LOWER ring walking backward, one running minimum over both walks."""
    from scipy.spatial.transform import Rotation
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 8
    w = alego.SynthWorld(seed=seed)
    o = ob.Oracle(P, lm_every=0)
    for t in range(2):
        o.ip(w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=60 + t))
        o.lo_features()
        if t == 1:
            sharp = o.get("sharp")[:, :3].copy()
            target = o.get("corner_last").copy()
        o.lo_scan2scan()
    resid = o.get("lo_resids")
    surf = resid[resid[:, 0] == 1]
    assert len(surf) >= 10
    x, _ = ob.solve(surf, np.zeros(6), P.lo_surf_iters, P.huber_delta)  # params_ after the first Solve (:418)
    Rm = Rotation.from_euler("ZYX", x[5:2:-1]).as_matrix()
    sel = (sharp.astype(np.float64) @ Rm.T + x[:3]).astype(np.float32)
    tx = target[:, :3]
    ring = target[:, 3].astype(np.int64)
    want = []
    for j, q in enumerate(sel):
        d = np.zeros(len(tx), np.float32)
        for c in range(3):
            t_ = q[c] - tx[:, c]
            d = d + t_ * t_
        closest = int(np.argmin(d))
        if not d[closest] < 25.0:
            continue
        rc = ring[closest]
        diff = (tx - q).astype(np.float64)
        pd = diff[:, 0] ** 2 + diff[:, 1] ** 2 + diff[:, 2] ** 2
        fwd = [k for k in range(closest + 1, len(tx)) if rc < ring[k] <= rc + 2]
        bwd = [k for k in range(closest - 1, -1, -1) if rc - 2 <= ring[k] < rc]
        walk = np.array(fwd + bwd, np.int64)
        walk = walk[pd[walk] < 25.0] if len(walk) else walk
        if len(walk):
            want.append([j, closest, int(walk[np.argmin(pd[walk])])])
    got = o.get("lo_corner_corr")
    assert len(got) > 10 and np.array_equal(got, np.array(want, np.int32))


def test_oracle_downsample_current_scan_composition(alego, ob):
    """downsampleCurrentScan (laserMapping.cpp:325-346): corner 0.4, surf 0.8, outlier 1.0, then surf_ds + outlier_ds through the
    0.8 filter again — recomposed from the stand-alone VoxelGrid (itself cross-checked against a numpy restatement) on the clouds
    LaserOdometry and ImageProjection hand over (/corner_last = less_sharp, /surf_last = less_flat, /outlier)."""
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 2
    w = alego.SynthWorld(seed=seed)
    cm, sm = w.make_map(3000, 10000, seed=seed, radius=50.0)
    for stable in (False, True):
        o = ob.Oracle(P, lm_every=1, stable_voxel=stable)
        o.lm_set_map(cm, sm)
        o.pipeline_step(w.render(P, alego.trajectory_pose(0, seed=seed), noise_seed=5))
        corner, surf, outlier = o.get("less_sharp"), o.get("less_flat_stable" if stable else "less_flat"), o.get("outlier_cloud")
        c_ds, _ = ob.voxel_grid(corner, P.lm_corner_leaf, stable)
        s_ds, _ = ob.voxel_grid(surf, P.lm_surf_leaf, stable)
        o_ds, _ = ob.voxel_grid(outlier, P.lm_outlier_leaf, stable)
        tot_ds, _ = ob.voxel_grid(np.concatenate([s_ds, o_ds]), P.lm_surf_leaf, stable)
        assert np.array_equal(o.get("lm_corner_ds"), c_ds) and np.array_equal(o.get("lm_surf_ds"), s_ds)
        assert np.array_equal(o.get("lm_outlier_ds"), o_ds) and np.array_equal(o.get("lm_surf_total_ds"), tot_ds)
        assert len(c_ds) > 50 and len(tot_ds) > 500 and len(o_ds) > 10
