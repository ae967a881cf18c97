"""ctypes binding of oracle/_ref/libalego_ref_*.so — TEST INFRASTRUCTURE.

oracle/_ref holds the REFERENCE'S OWN translation units (src/imageProjection.cpp, src/laserOdometry.cpp, src/laserMapping.cpp with
include/alego/*.h) compiled UNMODIFIED from /root/reference by oracle/refbuild/Makefile against stand-in headers for ROS, PCL,
Eigen, Ceres and GTSAM (oracle/refbuild/shims/).  It exists to pin oracle/alego_oracle.cpp (the port) to the reference:
tests/test_ref_pin.py asserts port == _ref.  The third-party pieces (KdTreeFLANN, VoxelGrid, Ceres' trust-region solver, Eigen's
eigen-solver / QR, GTSAM) are restated inside the stand-ins — those semantics stay "restated", everything else that runs is the
reference's code.

The stock build carries the reference's compile-time sensor constants (16 x 4000, include/alego/utility.h:50-57).  The per-preset
variants (suffix _vlp16_1800 / _hdl64_1800 / _hdl64_2048) are built from a temporary copy of the three headers in which ONLY those
constants are rewritten by sed (see the Makefile); the .cpp files are always compiled from where they lie.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
VARIANTS = ("stock", "vlp16_1800", "hdl64_1800", "hdl64_2048")
_libs = {}


def lib_path(stage, variant="stock"):
    suffix = "" if variant == "stock" else "_" + variant
    return os.path.join(REF_DIR, "libalego_ref_%s%s.so" % (stage, suffix))


def available(variant="stock"):
    return all(os.path.exists(lib_path(s, variant)) for s in ("ip", "lo", "lm"))


def build(verbose=False):
    """Needs /root/reference (this container only).  Elsewhere the prebuilt libraries under oracle/_ref are used as they are."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "refbuild")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-3000:], r.stderr[-3000:])
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed")


def _lib(stage, variant):
    key = (stage, variant)
    if key not in _libs:
        L = C.CDLL(lib_path(stage, variant))
        g = getattr(L, "ref_%s_get" % stage)
        g.restype = C.c_int64
        g.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        getattr(L, "ref_%s_create" % stage).restype = C.c_void_p
        getattr(L, "ref_%s_destroy" % stage).argtypes = [C.c_void_p]
        _libs[key] = L
    return _libs[key]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f4(a):
    return np.ascontiguousarray(a, np.float32).reshape(-1, 4)


_DT = {
    "range_mat": np.float64, "label_mat": np.int32, "ground_mat": np.uint8, "full_cloud": np.float32, "segmented_cloud": np.float32,
    "outlier_cloud": np.float32, "startRingIndex": np.int32, "endRingIndex": np.int32, "segmentedCloudGroundFlag": np.uint8,
    "segmentedCloudColInd": np.int32, "segmentedCloudRange": np.float32, "startOrientation": np.float32, "endOrientation": np.float32,
    "orientationDiff": np.float32, "label_cnt": np.int32,
    "cloud_curvature": np.float64, "cloud_neighbor_picked": np.uint8, "cloud_label": np.int32, "cloud_sort_idx": np.int32,
    "sharp": np.float32, "less_sharp": np.float32, "flat": np.float32, "less_flat": np.float32, "surf_last": np.float32,
    "corner_last": np.float32, "lo_params": np.float64, "t_w_cur": np.float64, "r_w_cur": np.float64, "odom_lidar": np.float64,
    "lo_trace": np.float64, "lo_solve_iterations": np.int32,
    "lm_params": np.float64, "lm_trace": np.float64, "lm_solve_iterations": np.int32, "lm_solve_blocks": np.int32,
    "lm_corner_ds": np.float32, "lm_surf_ds": np.float32, "lm_outlier_ds": np.float32, "lm_surf_total_ds": np.float32,
    "t_map2laser": np.float64, "q_map2laser": np.float64, "t_map2odom": np.float64, "q_map2odom": np.float64,
    "corner_from_map_ds": np.float32, "surf_from_map_ds": np.float32, "keyposes_6d": np.float64, "n_keyframes": np.int32,
    "lm_constants": np.float64, "lm_timing_ms": np.float64,
}
_COLS = {"full_cloud": 4, "segmented_cloud": 4, "outlier_cloud": 4, "sharp": 4, "less_sharp": 4, "flat": 4, "less_flat": 4,
         "surf_last": 4, "corner_last": 4, "lo_trace": 7, "lm_trace": 7, "lm_corner_ds": 4, "lm_surf_ds": 4, "lm_outlier_ds": 4,
         "lm_surf_total_ds": 4, "corner_from_map_ds": 4, "surf_from_map_ds": 4, "keyposes_6d": 7}


class _Stage:
    stage = None

    def __init__(self, variant="stock"):
        self.L = _lib(self.stage, variant)
        self.h = C.c_void_p(getattr(self.L, "ref_%s_create" % self.stage)())
        self.variant = variant

    def get(self, name):
        g = getattr(self.L, "ref_%s_get" % self.stage)
        n = g(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        dt = np.dtype(_DT[name])
        a = np.zeros(n // dt.itemsize, dt)
        if n:
            g(self.h, name.encode(), _p(a), n)
        c = _COLS.get(name)
        return a.reshape(-1, c) if c else a

    def constants(self):
        k = np.zeros(9)
        getattr(self.L, "ref_%s_constants" % self.stage)(_p(k))
        return k


class RefImageProjection(_Stage):
    """loam::ImageProjection (src/imageProjection.cpp), one pcCB per process() call."""
    stage = "ip"

    def process(self, scan):
        scan = _f4(scan)
        return self.L.ref_ip_process(self.h, _p(scan), len(scan))


class RefLaserOdometry(_Stage):
    """loam::LaserOdometry (src/laserOdometry.cpp), one mainLoop iteration per process() call."""
    stage = "lo"

    def process(self, ip):
        """ip: anything with .get(name) returning ImageProjection's outputs (a RefImageProjection or the port Oracle)."""
        seg = _f4(ip.get("segmented_cloud"))
        outl = _f4(ip.get("outlier_cloud"))
        sr = np.ascontiguousarray(ip.get("startRingIndex"), np.int32)
        er = np.ascontiguousarray(ip.get("endRingIndex"), np.int32)
        gf = np.ascontiguousarray(ip.get("segmentedCloudGroundFlag"), np.uint8)
        col = np.ascontiguousarray(ip.get("segmentedCloudColInd"), np.int32)
        rng = np.ascontiguousarray(ip.get("segmentedCloudRange"), np.float32)
        so, eo, od = (float(np.asarray(ip.get(k)).ravel()[0]) for k in ("startOrientation", "endOrientation", "orientationDiff"))
        return self.L.ref_lo_process(self.h, _p(seg), len(seg), _p(sr), _p(er), _p(gf), _p(col), _p(rng), C.c_float(so), C.c_float(eo),
                                     C.c_float(od), _p(outl), len(outl))

    def set_params(self, p):
        p = np.ascontiguousarray(p, np.float64)
        self.L.ref_lo_set_params(self.h, _p(p))

    def imu(self, stamp, quat_xyzw, ang_vel, lin_acc):
        q, w, a = (np.ascontiguousarray(v, np.float64) for v in (quat_xyzw, ang_vel, lin_acc))
        self.L.ref_lo_imu(self.h, C.c_double(stamp), _p(q), _p(w), _p(a))

    def imu_state(self):
        n = int(self.constants()[8])
        q = np.zeros((10, n))
        ptrs = np.zeros(3, np.int32)
        self.L.ref_lo_imu_state(self.h, _p(q), _p(ptrs))
        return q, ptrs

    def adjust_distortion(self, cloud, col, start_ori, end_ori, scan_time):
        out = np.array(cloud, np.float32).reshape(-1, 4).copy()
        col = np.ascontiguousarray(col, np.int32)
        it = self.L.ref_lo_adjust_distortion(self.h, _p(out), len(out), _p(col), C.c_float(start_ori), C.c_float(end_ori),
                                             C.c_double(scan_time))
        return out, it


class RefLaserMapping(_Stage):
    """loam::LaserMapping (src/laserMapping.cpp)."""
    stage = "lm"

    def scan2map(self, map_corner, map_surf, corner, surf, outlier, t_odom, q_odom_wxyz, params=None):
        """downsampleCurrentScan + scan2MapOptimization + transformUpdate against a GIVEN local map (corner_from_map_ds_ /
        surf_from_map_ds_ set directly), the odometry pose entering through laserOdomHandler."""
        a = [_f4(x) for x in (map_corner, map_surf, corner, surf, outlier)]
        t = np.ascontiguousarray(t_odom, np.float64)
        q = np.ascontiguousarray(q_odom_wxyz, np.float64)
        pp = None if params is None else np.ascontiguousarray(params, np.float64)
        return self.L.ref_lm_scan2map(self.h, _p(a[0]), len(a[0]), _p(a[1]), len(a[1]), _p(a[2]), len(a[2]), _p(a[3]), len(a[3]),
                                      _p(a[4]), len(a[4]), _p(t), _p(q), None if pp is None else _p(pp))

    def frame(self, corner, surf, outlier, t_odom, q_odom_wxyz, stamp):
        """One synchronised input set through the reference's handlers and mainLoop (maps every 2nd call, laserMapping.cpp:111-126;
        local map from its own keyframes)."""
        a = [_f4(x) for x in (corner, surf, outlier)]
        t = np.ascontiguousarray(t_odom, np.float64)
        q = np.ascontiguousarray(q_odom_wxyz, np.float64)
        return self.L.ref_lm_frame(self.h, _p(a[0]), len(a[0]), _p(a[1]), len(a[1]), _p(a[2]), len(a[2]), _p(t), _p(q), C.c_double(stamp))
