"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py on seeded inputs: outputs of the REFERENCE BUILD
oracle/_ref wherever the reference materialises them — cross-asserted against the port oracle when the files are written — and the
port's index lists where it does not; PCL's literal VoxelGrid record order throughout).

CPU (`-m "not gpu"`): the oracle still reproduces its committed golden outputs — integer/index/byte arrays bit-exact,
float32 pass-through bit-exact, double solver state to 1e-12.
GPU (`-m gpu`): the CUDA path, through the C ABI, reproduces the same files — labels / compaction / feature indices /
correspondences bit-exact, poses within 1e-4 (north_star tolerance).
"""
import os

import numpy as np
import pytest

from conftest import first_diff

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POSE_TOL = 1e-4
IP_KEYS = ["ground_mat", "label_mat", "startRingIndex", "endRingIndex", "segmentedCloudGroundFlag", "segmentedCloudColInd",
           "segmentedCloudRange", "outlier_cloud"]


def load(name):
    return dict(np.load(os.path.join(G, name)))


def vlp16(alego):
    return alego.default_params(alego.PRESET_VLP16_1800)


# ------------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_reproduces_cfg1(alego, ob):
    g = load("cfg1_vlp16_ip_features.npz")
    o = ob.Oracle(vlp16(alego), stable_voxel=False)
    assert o.ip(g["scan"]) == 0
    o.lo_features()
    M = len(g["segmentedCloudColInd"])
    for k in IP_KEYS + ["sharp_idx", "less_sharp_idx", "flat_idx", "less_flat"]:
        assert np.array_equal(np.asarray(o.get(k)), g[k]), k + ": " + first_diff(o.get(k), g[k])
    for k in ("cloud_curvature", "cloud_neighbor_picked", "cloud_label", "cloud_sort_idx"):
        assert np.array_equal(o.get(k)[5:M - 5], g[k]), k
    # known-answer sanity of the fixture itself
    lab = g["label_mat"]
    assert lab.min() == -1 and 5 < lab[(lab > 0) & (lab < 999999)].max() < 500
    assert 10 < len(g["sharp_idx"]) <= 16 * 12 and 100 < len(g["flat_idx"]) <= 16 * 24


@pytest.mark.parametrize("tag,corner_iters", [("5_5", 5), ("5_10", 10)])
def test_oracle_reproduces_cfg2(alego, ob, tag, corner_iters):
    g1, g = load("cfg1_vlp16_ip_features.npz"), load("cfg2_vlp16_scan2scan.npz")
    P = vlp16(alego)
    P.lo_corner_iters = corner_iters
    o = ob.Oracle(P, lm_every=0, stable_voxel=False)
    for t, s in enumerate([g1["scan"], g["scan1"], g["scan2"]]):
        o.ip(s)
        o.lo_features()
        o.lo_scan2scan()
        if t == 0:
            continue
        rep = o.report("lo")
        assert [rep["n_corner"], rep["n_surf"], rep["iterations"]] == list(g["%s_t%d_report" % (tag, t)])
        for k in ("lo_surf_corr", "lo_corner_corr"):
            assert np.array_equal(o.get(k), g["%s_t%d_%s" % (tag, t, k)]), k
        for k in ("lo_params", "t_w_cur", "r_w_cur", "lo_trace"):
            assert np.allclose(o.get(k), g["%s_t%d_%s" % (tag, t, k)], rtol=0, atol=1e-12), k
    # the odometry follows the synthetic motion (speed 0.25 m / sweep)
    assert 0.15 < np.linalg.norm(g[tag + "_t2_lo_params"][:2]) < 0.4


@pytest.mark.parametrize("tag,iters", [("2x20", (2, 20)), ("1x10", (1, 10))])
def test_oracle_reproduces_cfg3_small(alego, ob, tag, iters):
    g = load("cfg3_small_scan2map.npz")
    P = vlp16(alego)
    P.lm_outer_iters, P.lm_max_iters = iters
    o = ob.Oracle(P, stable_voxel=False)
    o.lm_set_map(g["corner_map"], g["surf_map"])
    o.lm_set_scan(g["corner"], g["surf"], g["outlier"])
    o.lm_set_odom(g["t_odom"], g["r_odom"])
    o.lm_set_params(g["x0"])
    o.lm_scan2map()
    rep = o.report("lm")
    assert [rep["n_corner"], rep["n_surf"], rep["iterations"]] == list(g[tag + "_report"])
    for k in ("lm_corner_sel", "lm_surf_sel", "lm_corner_ds", "lm_surf_total_ds"):
        assert np.array_equal(o.get(k), g[tag + "_" + k]), k
    for k in ("lm_params", "t_map2laser", "t_map2odom", "r_map2odom", "lm_trace"):
        assert np.allclose(o.get(k), g[tag + "_" + k], rtol=0, atol=1e-12), k
    assert np.linalg.norm(g[tag + "_lm_params"][:3]) < 0.08   # pulled from the (0.15,-0.12,0.05) prediction towards the truth


# ------------------------------------------------------------------------------------------------ GPU: CUDA path
@pytest.mark.gpu
def test_cuda_reproduces_cfg1(alego):
    g = load("cfg1_vlp16_ip_features.npz")
    a = alego.Alego(vlp16(alego), n_seq=1)
    buf, n = a.pack_scans([g["scan"]])
    a.ip_process(buf, n)
    a.lo_extract()
    M = len(g["segmentedCloudColInd"])
    assert np.array_equal(a.debug("range_mat"), g["range_mat"])
    out = a.ip_get(0)
    for k in IP_KEYS:
        got = out[k] if k in out else a.debug(k)
        assert np.array_equal(np.asarray(got).reshape(-1), np.asarray(g[k]).reshape(-1)), k + ": " + first_diff(np.asarray(got).reshape(-1), g[k].reshape(-1))
    ca = a.debug("cloud_curvature_abs").astype(np.float64)
    assert np.array_equal((ca * ca)[5:M - 5], g["cloud_curvature"])
    assert np.array_equal(a.debug("cloud_neighbor_picked")[5:M - 5], g["cloud_neighbor_picked"])
    assert np.array_equal(a.debug("cloud_label")[5:M - 5], g["cloud_label"])
    assert np.array_equal(a.debug("cloud_sort_idx")[5:M - 5], g["cloud_sort_idx"]), "std::sort permutation"
    for k in ("sharp_idx", "less_sharp_idx", "flat_idx"):
        assert np.array_equal(a.debug(k), g[k]), k
    assert np.array_equal(a.debug("less_flat"), g["less_flat"]), "per-ring VoxelGrid in PCL's record order"
    a.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,corner_iters", [("5_5", 5), ("5_10", 10)])
def test_cuda_reproduces_cfg2(alego, tag, corner_iters):
    g1, g = load("cfg1_vlp16_ip_features.npz"), load("cfg2_vlp16_scan2scan.npz")
    P = vlp16(alego)
    P.lo_corner_iters = corner_iters
    a = alego.Alego(P, n_seq=1)
    for t, s in enumerate([g1["scan"], g["scan1"], g["scan2"]]):
        buf, n = a.pack_scans([s])
        a.ip_process(buf, n)
        a.lo_extract()
        rc, rep = a.lo_scan2scan()
        if t == 0:
            continue
        assert [rep[0]["n_corner"], rep[0]["n_surf"], rep[0]["iterations"]] == list(g["%s_t%d_report" % (tag, t)])
        sc = a.debug("lo_surf_corr")
        cc = a.debug("lo_corner_corr")
        assert np.array_equal(sc[sc[:, 1] >= 0], g["%s_t%d_lo_surf_corr" % (tag, t)])
        assert np.array_equal(cc[cc[:, 1] >= 0], g["%s_t%d_lo_corner_corr" % (tag, t)])
        p, tw, rw = a.lo_get_state(0)
        assert np.abs(p - g["%s_t%d_lo_params" % (tag, t)]).max() < POSE_TOL
        assert np.abs(tw - g["%s_t%d_t_w_cur" % (tag, t)]).max() < POSE_TOL
        assert np.abs(rw.reshape(-1) - g["%s_t%d_r_w_cur" % (tag, t)]).max() < POSE_TOL
        tr = a.debug("lo_trace")
        assert tr.shape == g["%s_t%d_lo_trace" % (tag, t)].shape and np.allclose(tr, g["%s_t%d_lo_trace" % (tag, t)], rtol=1e-7, atol=1e-9)
    a.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,iters", [("2x20", (2, 20)), ("1x10", (1, 10))])
def test_cuda_reproduces_cfg3_small(alego, tag, iters):
    g = load("cfg3_small_scan2map.npz")
    P = vlp16(alego)
    P.lm_outer_iters, P.lm_max_iters = iters
    a = alego.Alego(P, n_seq=1)
    a.lm_set_map(0, g["corner_map"], g["surf_map"])
    a.lm_set_scan(0, g["corner"], g["surf"], g["outlier"])
    a.lm_set_odom(0, g["t_odom"], g["r_odom"])
    a.lm_set_params(0, g["x0"])
    rc, rep = a.lm_scan2map()
    assert [rep[0]["n_corner"], rep[0]["n_surf"], rep[0]["iterations"]] == list(g[tag + "_report"])
    assert np.array_equal(np.nonzero(a.debug("lm_edge")[:, 0])[0], g[tag + "_lm_corner_sel"])
    assert np.array_equal(np.nonzero(a.debug("lm_plane")[:, 0])[0], g[tag + "_lm_surf_sel"])
    assert np.array_equal(a.debug("lm_corner_ds"), g[tag + "_lm_corner_ds"])
    assert np.array_equal(a.debug("lm_surf_total_ds"), g[tag + "_lm_surf_total_ds"])
    st = a.lm_get_state(0)
    assert np.abs(st["params"] - g[tag + "_lm_params"]).max() < POSE_TOL
    assert np.abs(st["t_map2odom"] - g[tag + "_t_map2odom"]).max() < POSE_TOL
    assert np.abs(st["r_map2odom"].reshape(-1) - g[tag + "_r_map2odom"]).max() < POSE_TOL
    tr = a.debug("lm_trace")
    assert tr.shape == g[tag + "_lm_trace"].shape and np.allclose(tr, g[tag + "_lm_trace"], rtol=1e-6, atol=1e-8)
    a.close()


# ------------------------------------------------------------------------------------------------ N1: local-map assembly
def _n1_inputs(g):
    K = len(g["poses6"])
    return ([g["corner%d" % k] for k in range(K)], [g["surf%d" % k] for k in range(K)], [g["outlier%d" % k] for k in range(K)],
            g["poses6"])


def test_oracle_reproduces_n1_local_map(ob):
    g = load("n1_local_map.npz")
    ck, sk, okf, poses = _n1_inputs(g)
    cm, sm, M = ob.lm_assemble_map(ck, sk, okf, poses, 0.4, 0.8, stable=False)
    assert np.array_equal(M, g["matrices"]) and np.array_equal(cm, g["corner_from_map_ds"]) and np.array_equal(sm, g["surf_from_map_ds"])
    # known-answer sanity of the fixture: rotations are orthonormal, fewer map points than inputs, PCL order agrees to 2e-5
    for Mk in M.reshape(-1, 3, 4):
        assert np.abs(Mk[:, :3] @ Mk[:, :3].T - np.eye(3)).max() < 1e-6
    assert 0 < len(sm) < sum(len(x) for x in sk) + sum(len(x) for x in okf)
    cm2, sm2, _ = ob.lm_assemble_map(ck, sk, okf, poses, 0.4, 0.8, stable=True)
    assert cm2.shape == cm.shape and np.allclose(cm2, cm, rtol=0, atol=2e-5) and np.allclose(sm2, sm, rtol=0, atol=2e-5)


@pytest.mark.gpu
def test_gpu_reproduces_n1_local_map(alego):
    g = load("n1_local_map.npz")
    ck, sk, okf, poses = _n1_inputs(g)
    a = alego.Alego(vlp16(alego), n_seq=1)
    a.lm_assemble_map(0, ck, sk, okf, poses)
    cm, sm = a.lm_get_map(0)
    assert np.array_equal(cm, g["corner_from_map_ds"]), first_diff(cm, g["corner_from_map_ds"])
    assert np.array_equal(sm, g["surf_from_map_ds"]), first_diff(sm, g["surf_from_map_ds"])
    a.close()
