"""Golden vectors of the SURVEY §8f rows N2 (adjustDistortion) and N4 (loop-closure ICP): tests/golden/n2_adjust_distortion.npz and
n4_loop_closure_icp.npz, made by tests/golden/make_golden_next.py from the oracle on seeded inputs.

CPU (`-m "not gpu"`): the oracle still reproduces its committed outputs bit for bit.
GPU (`-m gpu`): the CUDA path, through the C ABI, reproduces the same files within the tolerances stated at each assert (the live
device-vs-oracle comparisons of the same code paths are in tests/test_next_rows.py).
"""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name)))


def vlp16(alego):
    return alego.default_params(alego.PRESET_VLP16_1800)


# ------------------------------------------------------------------------------------------------ N2: adjustDistortion
N2_TOL = 5e-5  # metres: float sin / cos of the device vs glibc (tests/test_next_rows.py states the reasoning)


def test_oracle_reproduces_n2_adjust_distortion(alego, ob):
    g = load("n2_adjust_distortion.npz")
    o = ob.Oracle(vlp16(alego))
    assert o.ip(g["scan"]) == 0
    assert np.array_equal(o.get("segmented_cloud"), g["segmented_cloud"]) and np.array_equal(o.get("segmentedCloudColInd"), g["col"])
    out, visited, it = ob.adjust_distortion(g["segmented_cloud"], g["col"], float(o.get("startOrientation")), float(o.get("endOrientation")),
                                            1800, float(g["scan_time"]), g["queue"], int(g["ptr_last"]), int(g["ptr_last_iter"]))
    assert visited == int(g["visited"]) == len(g["segmented_cloud"]) and it == int(g["ptr_last_iter_out"])
    assert np.array_equal(out, g["adjusted"])
    # known-answer sanity of the fixture: point 0 and the intensities untouched, the motion is metres at range
    assert np.array_equal(out[0], g["segmented_cloud"][0]) and np.array_equal(out[:, 3], g["segmented_cloud"][:, 3])
    assert 1.0 < np.abs(out[:, :3] - g["segmented_cloud"][:, :3]).max() < 20.0


@pytest.mark.gpu
def test_gpu_reproduces_n2_adjust_distortion(alego):
    g = load("n2_adjust_distortion.npz")
    a = alego.Alego(vlp16(alego), n_seq=1)
    buf, n = a.pack_scans([g["scan"]])
    a.ip_process(buf, n)
    assert np.array_equal(a.ip_get(0)["segmented_cloud"], g["segmented_cloud"])
    n_adj, it = a.lo_adjust_distortion([float(g["scan_time"])], [g["queue"]], [int(g["ptr_last"])], [int(g["ptr_last_iter"])])
    assert n_adj[0] == int(g["visited"]) and it[0] == int(g["ptr_last_iter_out"])
    got = a.ip_get(0)["segmented_cloud"]
    assert np.array_equal(got[:, 3], g["adjusted"][:, 3])
    assert np.allclose(got[:, :3], g["adjusted"][:, :3], rtol=0, atol=N2_TOL), np.abs(got[:, :3] - g["adjusted"][:, :3]).max()
    a.close()


# ------------------------------------------------------------------------------------------------ N4: loop-closure ICP
def test_oracle_reproduces_n4_icp(ob):
    g = load("n4_loop_closure_icp.npz")
    r = ob.icp(g["source"], g["target"], exact_sums=True)
    assert r["iterations"] == int(g["iterations"]) and r["state"] == int(g["state"]) and r["converged"]
    assert np.array_equal(r["T"], g["T"]) and r["fitness"] == float(g["fitness"]) and np.array_equal(r["trace"], g["trace"])
    rf = ob.icp(g["source"], g["target"], exact_sums=False)
    assert rf["iterations"] == int(g["iterations_float_sums"]) and np.array_equal(rf["T"], g["T_float_sums"])
    # known-answer sanity: a rigid transform near the planted motion (yaw 0.02, t = 0.25, 0.2, -0.04) — point-to-point ICP between
    # two different sparse samplings of the same surfaces settles a few decimetres off (scipy's ICP agrees, test_next_rows.py)
    R = g["T"][:3, :3].astype(np.float64)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(R) - 1) < 1e-6
    assert np.allclose(g["T"][:3, 3], [0.25, 0.2, -0.04], atol=0.5) and abs(np.arctan2(R[1, 0], R[0, 0]) - 0.02) < 0.01


@pytest.mark.gpu
def test_gpu_reproduces_n4_icp(alego):
    g = load("n4_loop_closure_icp.npz")
    a = alego.Alego(vlp16(alego), n_seq=1)
    r = a.lc_icp(g["source"], g["target"])
    assert r["converged"] and r["state"] == int(g["state"]) and abs(r["iterations"] - int(g["iterations"])) <= 1
    # same iteration count -> float rounding of the 4x4 chain; one apart -> the last step, <= 1e-3 m by the stop rule
    tol = 2e-5 if r["iterations"] == int(g["iterations"]) else 2e-3
    assert np.allclose(r["T"], g["T"], rtol=0, atol=tol), (r["T"], g["T"])
    assert abs(r["fitness"] - float(g["fitness"])) < 1e-3 * float(g["fitness"]) + tol
    k = min(r["iterations"], int(g["iterations"])) - 1
    assert np.array_equal(r["trace"][:k, 0], g["trace"][:k, 0])  # correspondence counts, iteration by iteration
    a.close()
