"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see the header of oracle/alego_oracle.cpp).

PARITY UNPINNED: the reference has no tests or golden vectors and cannot be built here; this oracle is a
restatement ("port") of its hot path, cross-checked in tests/test_oracle_*.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(verbose=False):
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-3000:], r.stderr[-3000:])
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_void_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_get.restype = C.c_int64
        L.oracle_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        for f in ("oracle_ip", "oracle_pipeline_step"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        for f in ("oracle_lo_features", "oracle_lo_scan2scan", "oracle_lm_scan2map"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.oracle_config.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_lm_set_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_lm_set_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_lm_set_odom.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_lm_set_params.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_lo_set_params.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_report.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_voxel_grid.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
        L.oracle_knn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_eval_residual.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.oracle_eig3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_lstsq5x3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_std_sort_by_key.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_default_params.argtypes = [C.c_void_p, C.c_int]
        L.oracle_lm_assemble_map.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_float, C.c_float, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_void_p,
                                             C.POINTER(C.c_int), C.c_void_p]
        L.oracle_adjust_distortion.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_double, C.c_double,
                                               C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.oracle_umeyama_rotation.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_icp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p,
                                 C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_DT = {
    "range_mat": np.float64, "full_cloud": np.float32, "ground_mat": np.uint8, "label_mat": np.int32, "startRingIndex": np.int32,
    "endRingIndex": np.int32, "segmentedCloudGroundFlag": np.uint8, "segmentedCloudColInd": np.int32,
    "segmentedCloudRange": np.float32, "segmented_cloud": np.float32, "outlier_cloud": np.float32,
    "startOrientation": np.float32, "endOrientation": np.float32, "orientationDiff": np.float32, "min_margin_row": np.float64,
    "min_margin_col": np.float64, "n_dup_cells": np.int32, "cloud_curvature": np.float64, "cloud_neighbor_picked": np.uint8,
    "cloud_label": np.int32, "cloud_sort_idx": np.int32, "sharp_idx": np.int32, "less_sharp_idx": np.int32, "flat_idx": np.int32,
    "less_flat_scan_idx": np.int32, "sharp": np.float32, "less_sharp": np.float32, "flat": np.float32, "less_flat": np.float32,
    "less_flat_stable": np.float32, "n_tie_segments": np.int32, "tie_sensitive": np.int32, "surf_last": np.float32,
    "corner_last": np.float32, "lo_params": np.float64, "t_w_cur": np.float64, "r_w_cur": np.float64, "lo_surf_corr": np.int32,
    "lo_corner_corr": np.int32, "lo_trace": np.float64, "lm_trace": np.float64, "lm_params": np.float64,
    "t_map2laser": np.float64, "r_map2laser": np.float64, "t_map2odom": np.float64, "r_map2odom": np.float64,
    "lm_corner_ds": np.float32, "lm_surf_ds": np.float32, "lm_outlier_ds": np.float32, "lm_surf_total_ds": np.float32,
    "lm_corner_sel": np.int32, "lm_surf_sel": np.int32, "lm_resids": np.float64, "lo_resids": np.float64, "timings_ms": np.float64,
}
_COLS = {"full_cloud": 4, "segmented_cloud": 4, "outlier_cloud": 4, "sharp": 4, "less_sharp": 4, "flat": 4, "less_flat": 4,
         "less_flat_stable": 4, "surf_last": 4, "corner_last": 4, "lo_surf_corr": 4, "lo_corner_corr": 3, "lo_trace": 7,
         "lm_trace": 7, "lm_corner_ds": 4, "lm_surf_ds": 4, "lm_outlier_ds": 4, "lm_surf_total_ds": 4, "lm_resids": 14, "lo_resids": 14}


class SolveReport(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_corner", C.c_int32), ("n_surf", C.c_int32), ("iterations", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double)]


class Oracle:
    """One sequence of the reference's CPU path.  `params` is any ctypes struct laid out like AlegoParams."""

    def __init__(self, params, lm_every=1, stable_voxel=False):
        self.L = lib()
        self.h = C.c_void_p(self.L.oracle_create(C.byref(params)))
        if not self.h:
            raise RuntimeError("oracle_create failed")
        self.L.oracle_config(self.h, lm_every, int(stable_voxel))

    def __del__(self):
        try:
            if self.h:
                self.L.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def ip(self, scan):
        scan = np.ascontiguousarray(scan, np.float32).reshape(-1, 4)
        return self.L.oracle_ip(self.h, _p(scan), len(scan))

    def lo_features(self):
        return self.L.oracle_lo_features(self.h)

    def lo_scan2scan(self):
        return self.L.oracle_lo_scan2scan(self.h)

    def lm_set_map(self, corner, surf):
        corner = np.ascontiguousarray(corner, np.float32).reshape(-1, 4)
        surf = np.ascontiguousarray(surf, np.float32).reshape(-1, 4)
        return self.L.oracle_lm_set_map(self.h, _p(corner), len(corner), _p(surf), len(surf))

    def lm_set_scan(self, corner, surf, outlier):
        a = [np.ascontiguousarray(x, np.float32).reshape(-1, 4) for x in (corner, surf, outlier)]
        return self.L.oracle_lm_set_scan(self.h, _p(a[0]), len(a[0]), _p(a[1]), len(a[1]), _p(a[2]), len(a[2]))

    def lm_set_odom(self, t, r):
        t = np.ascontiguousarray(t, np.float64)
        r = np.ascontiguousarray(r, np.float64).reshape(9)
        return self.L.oracle_lm_set_odom(self.h, _p(t), _p(r))

    def lm_set_params(self, p):
        p = np.ascontiguousarray(p, np.float64)
        return self.L.oracle_lm_set_params(self.h, _p(p))

    def lo_set_params(self, p):
        p = np.ascontiguousarray(p, np.float64)
        return self.L.oracle_lo_set_params(self.h, _p(p))

    def lm_scan2map(self):
        return self.L.oracle_lm_scan2map(self.h)

    def pipeline_step(self, scan):
        scan = np.ascontiguousarray(scan, np.float32).reshape(-1, 4)
        return self.L.oracle_pipeline_step(self.h, _p(scan), len(scan))

    def report(self, which):
        r = SolveReport()
        self.L.oracle_get_report(self.h, 0 if which == "lo" else 1, C.byref(r))
        return {k: getattr(r, k) for k, _ in r._fields_}

    def get(self, name):
        n = self.L.oracle_get(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        dt = np.dtype(_DT[name])
        a = np.zeros(n // dt.itemsize, dt)
        if n:
            self.L.oracle_get(self.h, name.encode(), _p(a), n)
        c = _COLS.get(name)
        if c:
            return a.reshape(-1, c)
        return a if a.size != 1 or name.endswith("_idx") or name.startswith("cloud_") or name.startswith("segmentedCloud") else a[0]


def voxel_grid(pts, leaf, stable=False):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
    out = np.zeros_like(pts)
    keys = np.zeros(len(pts), np.uint32)
    n = C.c_int(0)
    lib().oracle_voxel_grid(_p(pts), len(pts), leaf, int(stable), _p(out), C.byref(n), _p(keys))
    return out[:n.value], keys[:n.value]


def _cloud_ptrs(clouds):
    clouds = [np.ascontiguousarray(c, np.float32).reshape(-1, 4) for c in clouds]
    ptrs = (C.c_void_p * max(len(clouds), 1))(*[c.ctypes.data for c in clouds])
    counts = np.array([len(c) for c in clouds], np.int32)
    return clouds, ptrs, counts


def lm_assemble_map(corner_kfs, surf_kfs, outlier_kfs, poses6, leaf_c=0.4, leaf_s=0.8, stable=True):
    """extractSurroundingKeyFrames' cloud side (laserMapping.cpp:194-323): returns (corner_from_map_ds, surf_from_map_ds,
    the K row-major 3x4 keyframe matrices)."""
    K = len(corner_kfs)
    ck, cp, cn = _cloud_ptrs(corner_kfs)
    sk, sp, sn = _cloud_ptrs(surf_kfs)
    ok, op, on = _cloud_ptrs(outlier_kfs)
    poses6 = np.ascontiguousarray(poses6, np.float32).reshape(-1, 6)
    co = np.zeros((max(int(cn.sum()), 1), 4), np.float32)
    so = np.zeros((max(int(sn.sum() + on.sum()), 1), 4), np.float32)
    M = np.zeros((max(K, 1), 12), np.float32)
    nco, nso = C.c_int(0), C.c_int(0)
    lib().oracle_lm_assemble_map(K, cp, _p(cn), sp, _p(sn), op, _p(on), _p(poses6), leaf_c, leaf_s, int(stable), _p(co), C.byref(nco),
                                 _p(so), C.byref(nso), _p(M))
    return co[:nco.value], so[:nso.value], M[:K]


def adjust_distortion(cloud, col, start_orientation, end_orientation, horizon_scan, scan_time, queue, ptr_last, ptr_last_iter,
                      scan_period=0.2):
    """LaserOdometry::adjustDistortion, IMU branch (laserOdometry.cpp:557-657).  queue: (10, len) float64 — time, roll, pitch, yaw,
    shift xyz, velocity xyz.  Returns (adjusted cloud, points visited, imu_ptr_last_iter_ afterwards)."""
    out = np.array(cloud, np.float32).reshape(-1, 4).copy()
    col = np.ascontiguousarray(col, np.int32)
    queue = np.ascontiguousarray(queue, np.float64).reshape(10, -1)
    it = C.c_int(int(ptr_last_iter))
    n = lib().oracle_adjust_distortion(_p(out), len(out), _p(col), float(start_orientation), float(end_orientation), int(horizon_scan),
                                       float(scan_period), float(scan_time), _p(queue), queue.shape[1], int(ptr_last), C.byref(it))
    return out, n, it.value


def umeyama_rotation(sigma):
    """R = U diag(1, 1, det(U) det(V)) V^T of the 3x3 covariance sigma = U S V^T (the rotation Eigen::umeyama returns)."""
    sigma = np.ascontiguousarray(sigma, np.float32).reshape(9)
    R = np.zeros(9, np.float32)
    lib().oracle_umeyama_rotation(_p(sigma), _p(R))
    return R.reshape(3, 3)


ICP_STATES = ("not converged", "iterations", "transform", "abs mse", "rel mse", "no correspondences")


def icp(src, tgt, max_corr_dist=100.0, max_iterations=100, transformation_epsilon=1e-6, fitness_epsilon=1e-6, exact_sums=True):
    """pcl::IterativeClosestPoint as LaserMapping::performLoopClosure configures it (laserMapping.cpp:667-688).  Returns a dict:
    T (4x4 float32 final_transformation_), fitness (getFitnessScore), converged, state, iterations, trace [iterations][14]."""
    src = np.ascontiguousarray(src, np.float32).reshape(-1, 4)
    tgt = np.ascontiguousarray(tgt, np.float32).reshape(-1, 4)
    T = np.zeros(16, np.float32)
    trace = np.zeros((max(max_iterations, 1), 14))
    fit, conv, st = C.c_double(0), C.c_int(0), C.c_int(0)
    it = lib().oracle_icp(_p(src), len(src), _p(tgt), len(tgt), max_corr_dist, max_iterations, transformation_epsilon, fitness_epsilon,
                          int(exact_sums), _p(T), C.byref(fit), C.byref(conv), C.byref(st), _p(trace))
    return {"T": T.reshape(4, 4), "fitness": fit.value, "converged": bool(conv.value), "state": st.value, "iterations": it,
            "trace": trace[:it]}


def knn(pts, q, k, brute=False):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
    q = np.ascontiguousarray(q, np.float32).reshape(-1, 4)
    idx = np.zeros((len(q), k), np.int32)
    d = np.zeros((len(q), k), np.float32)
    lib().oracle_knn(_p(pts), len(pts), _p(q), len(q), k, int(brute), _p(idx), _p(d))
    return idx, d


def eval_residual(f14, x):
    f14 = np.ascontiguousarray(f14, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    r = np.zeros(1)
    J = np.zeros(6)
    lib().oracle_eval_residual(_p(f14), _p(x), _p(r), _p(J))
    return r[0], J


def solve(f14, x0, max_iters, huber=0.1):
    f14 = np.ascontiguousarray(f14, np.float64).reshape(-1, 14)
    x = np.array(x0, np.float64)
    s = np.zeros(4)
    it = lib().oracle_solve(_p(f14), len(f14), _p(x), max_iters, huber, _p(s))
    return x, {"iterations": it, "initial_cost": s[0], "final_cost": s[1], "successful": int(s[2]), "termination": int(s[3])}


def eig3(A):
    A = np.ascontiguousarray(A, np.float64).reshape(9)
    w = np.zeros(3)
    V = np.zeros(9)
    lib().oracle_eig3(_p(A), _p(w), _p(V))
    return w, V.reshape(3, 3)


def lstsq5x3(A, b):
    A = np.ascontiguousarray(A, np.float64).reshape(15)
    b = np.ascontiguousarray(b, np.float64)
    n = np.zeros(3)
    lib().oracle_lstsq5x3(_p(A), _p(b), _p(n))
    return n


def std_sort_by_key(key):
    key = np.ascontiguousarray(key, np.float64)
    idx = np.arange(len(key), dtype=np.int32)
    lib().oracle_std_sort_by_key(_p(key), _p(idx), len(key))
    return idx
