"""Pins the port oracle (oracle/alego_oracle.cpp) to the REFERENCE'S OWN CODE (oracle/_ref).

oracle/_ref/libalego_ref_{ip,lo,lm}*.so are /root/reference's src/imageProjection.cpp, src/laserOdometry.cpp and src/laserMapping.cpp
(with include/alego/*.h) compiled UNMODIFIED against stand-in headers for ROS / PCL / Eigen / Ceres / GTSAM
(oracle/refbuild/, see its Makefile).  What runs in these tests on the `_ref` side is therefore the reference's own pcCB /
labelComponents, its own mainLoop (smoothness, occlusion, feature selection, association loops, pose integration), its own four cost
functions and its own scan2MapOptimization; the third-party pieces underneath (KdTreeFLANN, VoxelGrid, Ceres' trust-region LM,
Eigen's eigen-solver / QR, GTSAM) are restatements inside the stand-ins and stay "restated".

Bar: integer / index / byte / float32 arrays bit-exact; double solver state to 1e-9 (observed ~1e-13), identical iteration counts.

CPU tests run wherever the libraries exist (built here from /root/reference by __graft_entry__.build(); shipped prebuilt to the GPU
box).  The `-m gpu` tests at the bottom compare the CUDA path, through the C ABI, DIRECTLY with the reference build.
"""
import os

import numpy as np
import pytest

from conftest import first_diff

PRESET_VARIANT = {0: "vlp16_1800", 1: "hdl64_1800", 2: "hdl64_2048", 3: "stock"}
IP_KEYS = ["range_mat", "full_cloud", "ground_mat", "label_mat", "startRingIndex", "endRingIndex", "segmentedCloudGroundFlag",
           "segmentedCloudColInd", "segmentedCloudRange", "segmented_cloud", "outlier_cloud", "startOrientation", "endOrientation",
           "orientationDiff"]
TOL = 1e-9


@pytest.fixture(scope="session")
def rb():
    from oracle import ref_binding
    if not all(ref_binding.available(v) for v in ref_binding.VARIANTS):
        if os.path.exists("/root/reference/src/imageProjection.cpp"):
            ref_binding.build()
        else:
            pytest.skip("oracle/_ref libraries absent and /root/reference not present to build them")
    return ref_binding


def eq(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and np.array_equal(a.ravel(), b.ravel()), what + ": " + first_diff(a.ravel(), b.ravel())


def check_ip(r, o, tag):
    for k in IP_KEYS:
        eq(np.asarray(r.get(k)).ravel(), np.asarray(o.get(k)).ravel(), "%s %s" % (tag, k))


def check_features(rlo, o, tag):
    M = len(o.get("segmentedCloudColInd"))
    for k in ("cloud_curvature", "cloud_neighbor_picked", "cloud_label", "cloud_sort_idx"):
        # the reference (re)initialises only [5, M-5) per sweep (laserOdometry.cpp:122-129)
        eq(rlo.get(k)[5:max(M - 5, 5)], o.get(k)[5:max(M - 5, 5)], "%s %s" % (tag, k))
    seg = o.get("segmented_cloud")
    for cloud, idx in (("sharp", "sharp_idx"), ("less_sharp", "less_sharp_idx"), ("flat", "flat_idx")):
        eq(rlo.get(cloud), seg[o.get(idx)].reshape(-1, 4), "%s %s (points gathered by the port's index list)" % (tag, cloud))
    eq(rlo.get("less_flat"), o.get("less_flat"), tag + " less_flat (per-ring VoxelGrid, PCL order)")


def scans_for(alego, P, seed, n):
    w = alego.SynthWorld(seed=seed)
    return [w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=100 * seed + t) for t in range(n)]


def test_ref_build_constants(alego, rb):
    """each library carries the constants its name says (stock = include/alego/utility.h:50-57 as they are)"""
    for preset, variant in PRESET_VARIANT.items():
        P = alego.default_params(preset)
        for cls in (rb.RefImageProjection, rb.RefLaserOdometry, rb.RefLaserMapping):
            k = cls(variant).constants()
            assert [int(k[0]), int(k[1]), int(k[2])] == [P.n_scan, P.horizon_scan, P.ground_scan_id], (variant, k)
            assert k[3] == P.ang_res_x and k[4] == P.ang_res_y and k[5] == P.ang_bottom, (variant, k)
    k = rb.RefImageProjection("stock").constants()
    assert list(k[:6]) == [16, 4000, 10, 0.09, 2.0, 15.0] and k[7] == 1.047


def test_cost_functions_match_reference(ob, rb):
    """utility.h:122-349 — CornerCostFunction, SurfCostFunction, LidarEdgeCostFunction, LidarPlaneCostFunction: residual and the
    6 Jacobian entries of the port against the reference's own Evaluate() on 10 000 random blocks (rows a12, a13, a18-a20)."""
    import ctypes as C
    L = rb.RefLaserMapping("stock").L
    rng = np.random.default_rng(0)
    worst = 0.0
    for k in range(10000):
        kind = k % 4
        f = np.zeros(14)
        f[0] = kind
        f[1:4] = rng.uniform(-30, 30, 3)
        base = f[1:4] + rng.normal(0, 0.5, 3)
        f[4:7] = base + rng.normal(0, 1.0, 3)
        f[7:10] = base + rng.normal(0, 1.0, 3)
        f[10:13] = base + rng.normal(0, 1.0, 3)
        if kind == 3:
            nrm = rng.normal(0, 1, 3)
            f[4:7] = nrm / np.linalg.norm(nrm)
            f[13] = rng.uniform(-20, 20)
        x = rng.normal(0, 1, 6) * np.array([0.5, 0.5, 0.5, 0.3, 0.3, 1.0])
        r_ref, J_ref = np.zeros(1), np.zeros(6)
        assert L.ref_lm_eval_cost(f.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), r_ref.ctypes.data_as(C.c_void_p),
                                  J_ref.ctypes.data_as(C.c_void_p)) == 0
        r, J = ob.eval_residual(f, x)
        scale = max(1.0, abs(r_ref[0]), np.abs(J_ref).max())
        worst = max(worst, abs(r - r_ref[0]) / scale, np.abs(J - J_ref).max() / scale)
        # structural zeros are exact (partial Jacobians, utility.h:162-167, 226-231)
        assert np.array_equal(J == 0, J_ref == 0), (kind, J, J_ref)
    assert worst < 1e-12, worst


@pytest.mark.parametrize("preset", [3, 0, 1, 2])
def test_image_projection_matches_reference(alego, ob, rb, preset):
    """imageProjection.cpp:49-316 (rows a1-a4): every output of pcCB bit-exact, three seeded sweeps per preset."""
    P = alego.default_params(preset)
    r = rb.RefImageProjection(PRESET_VARIANT[preset])
    for seed in (0, 1, 2):
        scan = scans_for(alego, P, seed, 1)[0]
        o = ob.Oracle(P)
        assert o.ip(scan) == 0 and r.process(scan) == 0
        check_ip(r, o, "preset%d seed%d" % (preset, seed))
        lab = o.get("label_mat")
        assert lab.min() == -1 and (lab == 999999).any() and 3 < lab[(lab > 0) & (lab < 999999)].max() < 3000


@pytest.mark.parametrize("preset", [3, 0])
def test_image_projection_edge_cases_match_reference(alego, ob, rb, preset):
    """NaN / inf points, second returns in a cell (later point wins), rays outside the vertical field of view, half and 3-point
    sweeps, rays anywhere inside their cell — and state carried from one sweep to the next in the same node (the reset at the end
    of pcCB, imageProjection.cpp:193-205)."""
    P = alego.default_params(preset)
    w = alego.SynthWorld(seed=7)
    full = w.render(P, (0, 0, 0, 0), noise_seed=1)
    rng = np.random.default_rng(0)
    with_nan = full.copy()
    with_nan[rng.choice(len(full), 500, replace=False), rng.integers(0, 3, 500)] = np.nan
    with_nan[0, 0] = np.nan
    with_nan[-1, 2] = np.inf
    dup = np.concatenate([full, full[::7] * np.float32(1.01)])
    out_of_fov = full.copy()
    out_of_fov[::11, 2] += 40.0
    jitter = w.render(P, (0.3, -0.2, 0.0, 0.4), noise_seed=2, jitter_cells=0.45)
    r = rb.RefImageProjection(PRESET_VARIANT[preset])  # ONE node for all cases: exercises the end-of-callback reset
    rlo = rb.RefLaserOdometry(PRESET_VARIANT[preset])
    for i, s in enumerate([full, with_nan, dup, out_of_fov, full[: len(full) // 2], full[:3], jitter, full]):
        o = ob.Oracle(P)
        assert o.ip(s) == 0 and r.process(s) == 0
        check_ip(r, o, "case%d" % i)
        o.lo_features()
        assert rlo.process(r) == 0
        check_features(rlo, o, "case%d" % i)


@pytest.mark.parametrize("preset,seed", [(3, 3), (0, 0), (1, 2), (2, 4)])
def test_laser_odometry_matches_reference(alego, ob, rb, preset, seed):
    """laserOdometry.cpp:118-297 (rows a5-a8: curvature, occlusion, std::sort + feature selection, per-ring VoxelGrid) bit-exact and
    :316-535, 728-740 (rows a9-a14: association walks, the two solves over the reference's own cost functions, yaw-only pose
    integration) over a 5-sweep sequence: same correspondences => same iteration counts, cost / pose traces and poses."""
    P = alego.default_params(preset)
    o = ob.Oracle(P, lm_every=0, stable_voxel=False)
    rip, rlo = rb.RefImageProjection(PRESET_VARIANT[preset]), rb.RefLaserOdometry(PRESET_VARIANT[preset])
    moved = 0.0
    for t, scan in enumerate(scans_for(alego, P, seed, 5)):
        assert o.ip(scan) == 0 and rip.process(scan) == 0
        o.lo_features()
        o.lo_scan2scan()
        assert rlo.process(rip) == 0
        tag = "preset%d sweep%d" % (preset, t)
        check_features(rlo, o, tag)
        eq(rlo.get("surf_last"), o.get("surf_last"), tag + " surf_last")
        eq(rlo.get("corner_last"), o.get("corner_last"), tag + " corner_last")
        if t == 0:
            assert len(rlo.get("lo_solve_iterations")) == 0  # first sweep only initialises (:316-324)
            continue
        rep = o.report("lo")
        assert rep["n_surf"] >= 10 and rep["n_corner"] >= 10
        assert int(rlo.get("lo_solve_iterations").sum()) == rep["iterations"], (rlo.get("lo_solve_iterations"), rep)
        tr, to = rlo.get("lo_trace"), o.get("lo_trace")
        assert tr.shape == to.shape and np.abs(tr - to).max() < TOL, tag + " per-attempt cost / pose trace"
        assert np.abs(rlo.get("lo_params") - o.get("lo_params")).max() < TOL
        assert np.abs(rlo.get("t_w_cur") - o.get("t_w_cur")).max() < TOL
        assert np.abs(rlo.get("r_w_cur") - np.asarray(o.get("r_w_cur")).ravel()).max() < TOL
        moved = max(moved, float(np.abs(tr[-1, 1:] - tr[0, 1:]).max()))
    assert moved > 1e-3  # the solves actually move the pose (the comparison is not between two no-ops)


@pytest.mark.parametrize("n_corner,n_surf", [(6000, 30000), (50000, 200000)])
def test_scan_to_map_matches_reference(alego, ob, rb, n_corner, n_surf):
    """laserMapping.cpp:325-489 (rows a8, a15-a21) on BASELINE config 3 (50k corner + 200k surf local map) and a small map: the four
    VoxelGrid outputs bit-exact, same number of LM attempts in both outer iterations, trace / pose / map->odom to 1e-9."""
    P = alego.default_params(alego.PRESET_HDL64_1800)
    w = alego.SynthWorld(seed=5)
    cm, sm = w.make_map(n_corner, n_surf, seed=5, radius=80.0)
    scan = w.render(P, (0.0, 0.0, 0.0, 0.0), noise_seed=5)
    o = ob.Oracle(P, stable_voxel=False)
    o.ip(scan)
    o.lo_features()
    corner, surf, outl = o.get("less_sharp"), o.get("less_flat"), o.get("outlier_cloud")
    yaw = np.deg2rad(1.0)
    R0 = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
    t0 = np.array([0.15, -0.12, 0.05])
    x0 = np.array([0.15, -0.12, 0.05, 0.0, 0.0, yaw])
    o.lm_set_map(cm, sm)
    o.lm_set_scan(corner, surf, outl)
    o.lm_set_odom(t0, R0)
    o.lm_set_params(x0)
    o.lm_scan2map()
    r = rb.RefLaserMapping("hdl64_1800")
    assert r.scan2map(cm, sm, corner, surf, outl, t0, [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)], x0) == 0
    for k in ("lm_corner_ds", "lm_surf_ds", "lm_outlier_ds", "lm_surf_total_ds"):
        eq(r.get(k), o.get(k), k)
    rep = o.report("lm")
    assert rep["n_surf"] > 100 and rep["n_corner"] >= 10
    assert int(r.get("lm_solve_iterations").sum()) == rep["iterations"], (r.get("lm_solve_iterations"), rep)
    tr, to = r.get("lm_trace"), o.get("lm_trace")
    assert tr.shape == to.shape and np.abs(tr - to).max() < TOL
    assert np.abs(r.get("lm_params") - o.get("lm_params")).max() < TOL
    assert np.abs(r.get("t_map2odom") - o.get("t_map2odom")).max() < TOL
    # q_map2odom (w, x, y, z) against the port's rotation matrix
    qw, qx, qy, qz = r.get("q_map2odom")
    Rq = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                   [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                   [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
    assert np.abs(Rq.ravel() - np.asarray(o.get("r_map2odom")).ravel()).max() < TOL
    assert np.linalg.norm(o.get("lm_params")[:3]) < 0.08  # pulled towards the truth (identity)


def test_whole_chain_matches_reference(alego, ob, rb):
    """IP -> LO -> LM chained exactly like the three nodelets (topics replaced by function arguments), 6 sweeps, LaserMapping on
    every sweep against a fixed local map: the port's pipeline_step against the reference build, sweep by sweep."""
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 3
    w = alego.SynthWorld(seed=seed)
    cm, sm = w.make_map(6000, 30000, seed=seed, radius=60.0)
    o = ob.Oracle(P, lm_every=1, stable_voxel=False)
    o.lm_set_map(cm, sm)
    rip, rlo, rlm = rb.RefImageProjection("vlp16_1800"), rb.RefLaserOdometry("vlp16_1800"), rb.RefLaserMapping("vlp16_1800")
    for t in range(6):
        scan = w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=100 + t)
        o.pipeline_step(scan)
        assert rip.process(scan) == 0 and rlo.process(rip) == 0
        odom = rlo.get("odom_lidar") if t > 0 else np.array([0, 0, 0, 1.0, 0, 0, 0])
        # /corner_last, /surf_last, /outlier, /odom/lidar (laserOdometry.cpp:520-546; laserMapping.cpp:88-91)
        assert rlm.scan2map(cm, sm, rlo.get("corner_last"), rlo.get("surf_last"), rip.get("outlier_cloud"), odom[:3], odom[3:]) == 0
        assert np.abs(rlo.get("lo_params") - o.get("lo_params")).max() < TOL
        assert np.abs(rlm.get("lm_params") - o.get("lm_params")).max() < 1e-8, (t, rlm.get("lm_params"), o.get("lm_params"))
        assert np.abs(rlm.get("t_map2odom") - o.get("t_map2odom")).max() < 1e-8
        assert int(rlm.get("lm_solve_iterations").sum()) == o.report("lm")["iterations"]


def test_local_map_assembly_matches_reference(alego, ob, rb):
    """SURVEY §8f row N1 — extractSurroundingKeyFrames + saveKeyFramesAndFactor run closed-loop inside the reference's own mainLoop
    (laserMapping.cpp:102-131, 194-323, 491-545): its keyframe clouds and key poses fed to the port's lm_assemble_map reproduce the
    reference's corner_from_map_ds_ / surf_from_map_ds_ bit for bit."""
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 4
    w = alego.SynthWorld(seed=seed)
    rip, rlo, rlm = rb.RefImageProjection("vlp16_1800"), rb.RefLaserOdometry("vlp16_1800"), rb.RefLaserMapping("vlp16_1800")
    kf_corner, kf_surf, kf_outl = [], [], []
    checked = 0
    for t in range(12):
        # 0.6 m between sweeps: a new keyframe (> 1 m, laserMapping.cpp:501-508) on every mapped sweep (every 2nd, :112)
        scan = w.render(P, (0.6 * t, 0.05 * t, 0.0, 0.01 * t), noise_seed=50 + t)
        assert rip.process(scan) == 0 and rlo.process(rip) == 0
        odom = rlo.get("odom_lidar") if t > 0 else np.array([0, 0, 0, 1.0, 0, 0, 0])
        n_before = int(rlm.get("n_keyframes")[0]) if t > 0 else 0
        assert rlm.frame(rlo.get("corner_last"), rlo.get("surf_last"), rip.get("outlier_cloud"), odom[:3], odom[3:], 10.0 + 0.1 * t) == 0
        if t % 2:
            assert int(rlm.get("n_keyframes")[0]) == n_before  # odd calls are skipped by mainLoop's frame counter
            continue
        poses = rlm.get("keyposes_6d")
        if n_before > 0:
            # the local map this sweep was matched against = the keyframes that existed before it
            co, so, _ = ob.lm_assemble_map(kf_corner, kf_surf, kf_outl, poses[:n_before, :6], stable=False)
            eq(rlm.get("corner_from_map_ds"), co, "sweep %d corner_from_map_ds" % t)
            eq(rlm.get("surf_from_map_ds"), so, "sweep %d surf_from_map_ds" % t)
            checked += 1
        if int(rlm.get("n_keyframes")[0]) > n_before:
            kf_corner.append(rlm.get("lm_corner_ds").copy())
            kf_surf.append(rlm.get("lm_surf_ds").copy())
            kf_outl.append(rlm.get("lm_outlier_ds").copy())
    assert checked >= 4 and len(kf_corner) >= 5


def test_adjust_distortion_matches_reference(alego, ob, rb):
    """SURVEY §8f row N2 — imuHandler + adjustDistortion (laserOdometry.cpp:557-657, 761-804), dead code in the reference's mainLoop
    (:115) but restated by the port: integer outputs and the corrected points bit-exact against the reference's own evaluation."""
    P = alego.default_params(alego.PRESET_REFERENCE)
    rlo = rb.RefLaserOdometry("stock")
    rng = np.random.default_rng(1)
    t0 = 50.0
    for k in range(120):  # 200 Hz IMU around the sweep
        ang = np.array([0.02 * np.sin(0.1 * k), 0.015 * np.cos(0.07 * k), 0.3 * k / 120.0])
        cr, sr, cp, sp, cy, sy = np.cos(ang[0] / 2), np.sin(ang[0] / 2), np.cos(ang[1] / 2), np.sin(ang[1] / 2), np.cos(ang[2] / 2), np.sin(ang[2] / 2)
        q = [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]
        rlo.imu(t0 - 0.1 + 0.005 * k, q, rng.normal(0, 0.01, 3), np.array([0.3, -0.2, 9.81]) + rng.normal(0, 0.05, 3))
    queue, ptrs = rlo.imu_state()
    w = alego.SynthWorld(seed=2)
    o = ob.Oracle(P)
    o.ip(w.render(P, (0, 0, 0, 0), noise_seed=3))
    cloud, col = o.get("segmented_cloud"), o.get("segmentedCloudColInd")
    # one ring only: on a full ring-major cloud the forward-only IMU pointer makes the function return at the first ring restart
    n = int(o.get("endRingIndex")[0]) + 6
    cloud, col = cloud[:n], col[:n]
    so, eo = float(o.get("startOrientation")), float(o.get("endOrientation"))
    ref_out, ref_it = rlo.adjust_distortion(cloud, col, so, eo, t0 + 0.2)
    port_out, visited, port_it = ob.adjust_distortion(cloud, col, so, eo, P.horizon_scan, t0 + 0.2, queue, int(ptrs[1]), int(ptrs[2]))
    assert port_it == ref_it
    eq(port_out, ref_out, "adjusted cloud")
    assert np.abs(ref_out[:, :3] - cloud[:, :3]).max() > 1e-3  # the correction is not a no-op


def test_reference_nodes_are_independent_across_threads(alego, rb):
    """bench.py's parity check drives several reference chains from a thread pool: the stand-in middleware keeps its state per
    thread, so concurrent chains give exactly what they give one after the other."""
    from concurrent.futures import ThreadPoolExecutor
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seeds = [1, 2, 3, 4]
    worlds = {s: alego.SynthWorld(seed=s) for s in seeds}
    maps = {s: worlds[s].make_map(3000, 15000, seed=s, radius=60.0) for s in seeds}
    sweeps = {s: [worlds[s].render(P, alego.trajectory_pose(t, seed=s), noise_seed=10 * s + t) for t in range(3)] for s in seeds}

    def chain(s):
        rip, rlo, rlm = rb.RefImageProjection("vlp16_1800"), rb.RefLaserOdometry("vlp16_1800"), rb.RefLaserMapping("vlp16_1800")
        out = []
        for t, scan in enumerate(sweeps[s]):
            assert rip.process(scan) == 0 and rlo.process(rip) == 0
            odom = rlo.get("odom_lidar") if t > 0 else np.array([0, 0, 0, 1.0, 0, 0, 0])
            assert rlm.scan2map(maps[s][0], maps[s][1], rlo.get("corner_last"), rlo.get("surf_last"), rip.get("outlier_cloud"), odom[:3], odom[3:]) == 0
            out.append((rip.get("label_mat").copy(), rlo.get("less_flat").copy(), rlo.get("lo_params").copy(), rlm.get("lm_params").copy(),
                        rlm.get("lm_trace").copy()))
        return out

    serial = [chain(s) for s in seeds]
    with ThreadPoolExecutor(max_workers=4) as ex:
        threaded = list(ex.map(chain, seeds))
    for a, b in zip(serial, threaded):
        for ra, rb_ in zip(a, b):
            for x, y in zip(ra, rb_):
                assert np.array_equal(x, y)


# ------------------------------------------------------------------------------------------------ GPU: CUDA vs the reference build
def _gpu_vs_ref_ip_features(alego, ob, rb, preset, seeds):
    P = alego.default_params(preset)
    scans = [scans_for(alego, P, s, 1)[0] for s in seeds]
    g = alego.Alego(P, n_seq=len(scans))
    buf, n = g.pack_scans(scans)
    g.ip_process(buf, n)
    g.lo_extract()
    for b, s in enumerate(scans):
        rip, rlo = rb.RefImageProjection(PRESET_VARIANT[preset]), rb.RefLaserOdometry(PRESET_VARIANT[preset])
        assert rip.process(s) == 0 and rlo.process(rip) == 0
        eq(g.debug("label_mat", b).ravel(), rip.get("label_mat"), "label_mat")
        eq(g.debug("ground_mat", b).ravel(), rip.get("ground_mat"), "ground_mat")
        for k in ("startRingIndex", "endRingIndex", "segmentedCloudGroundFlag", "segmentedCloudColInd", "segmentedCloudRange"):
            eq(g.debug(k, b), rip.get(k), k)
        eq(g.debug("segmented_cloud", b), rip.get("segmented_cloud"), "segmented_cloud")
        eq(g.debug("outlier_cloud", b), rip.get("outlier_cloud"), "outlier_cloud")
        M = len(rip.get("segmentedCloudColInd"))
        eq(g.debug("cloud_label", b)[5:M - 5], rlo.get("cloud_label")[5:M - 5], "cloud_label")
        eq(g.debug("cloud_sort_idx", b)[5:M - 5], rlo.get("cloud_sort_idx")[5:M - 5], "cloud_sort_idx")
        seg = rip.get("segmented_cloud")
        for cloud, idx in (("sharp", "sharp_idx"), ("less_sharp", "less_sharp_idx"), ("flat", "flat_idx")):
            eq(seg[g.debug(idx, b)].reshape(-1, 4), rlo.get(cloud), cloud + " (feature indices)")
        lf = g.debug("less_flat", b)
        assert lf.shape == rlo.get("less_flat").shape and np.abs(lf - rlo.get("less_flat")).max() < 2e-5  # a8: summation order
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("preset", [3, 0, 1, 2])
def test_gpu_ip_and_features_match_reference(alego, ob, rb, preset):
    """CUDA (C ABI) vs the reference's own ImageProjection + feature extraction: labels, compaction arrays, feature indices
    bit-exact on every preset, the 64 x 2048 sweep of BASELINE config 4 included."""
    _gpu_vs_ref_ip_features(alego, ob, rb, preset, [0, 1])


@pytest.mark.gpu
@pytest.mark.parametrize("preset,seed", [(0, 3), (1, 2), (2, 4)])
def test_gpu_pipeline_matches_reference(alego, ob, rb, preset, seed):
    """CUDA pipeline (IP -> LO -> LM, host buffers in, poses out) vs the reference build chained the same way: final poses within
    north_star's 1e-4 m / 1e-4 rad after the same number of solver attempts, feature indices bit-exact on every sweep.  The
    reference side sums VoxelGrid centroids in PCL's introsort order (row a8), the device in input order — the pose bar holds
    across that difference."""
    P = alego.default_params(preset)
    w = alego.SynthWorld(seed=seed)
    cm, sm = w.make_map(6000, 30000, seed=seed, radius=60.0)
    rip, rlo, rlm = (c(PRESET_VARIANT[preset]) for c in (rb.RefImageProjection, rb.RefLaserOdometry, rb.RefLaserMapping))
    g = alego.Alego(P, n_seq=1)
    g.lm_set_map(0, cm, sm)
    g.pipeline_config(lm_every=1)
    for t in range(5):
        scan = w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=100 + t)
        buf, n = g.pack_scans([scan])
        poses = g.pipeline_step(buf, n)
        assert rip.process(scan) == 0 and rlo.process(rip) == 0
        odom = rlo.get("odom_lidar") if t > 0 else np.array([0, 0, 0, 1.0, 0, 0, 0])
        assert rlm.scan2map(cm, sm, rlo.get("corner_last"), rlo.get("surf_last"), rip.get("outlier_cloud"), odom[:3], odom[3:]) == 0
        seg = rip.get("segmented_cloud")
        eq(g.debug("label_mat").ravel(), rip.get("label_mat"), "label_mat sweep %d" % t)
        for cloud, idx in (("sharp", "sharp_idx"), ("less_sharp", "less_sharp_idx"), ("flat", "flat_idx")):
            eq(seg[g.debug(idx)].reshape(-1, 4), rlo.get(cloud), "%s sweep %d" % (cloud, t))
        assert np.abs(g.debug("lo_params") - rlo.get("lo_params")).max() < 1e-4
        assert np.abs(poses[0, 3:9] - rlm.get("lm_params")).max() < 1e-4, (t, poses[0, 3:9], rlm.get("lm_params"))
        if t > 0:
            assert g.solve_report("lo")["iterations"] == int(rlo.get("lo_solve_iterations").sum())
    g.close()
