"""alego_b200 — ctypes binding of libalego_b200.so (include/alego_b200.h), the B200-native implementation of
A-LeGO-LOAM's per-scan hot path (ImageProjection -> LaserOdometry -> LaserMapping), plus the synthetic sweep
generator.  There is NO CPU fallback: every stage call goes to the CUDA library and raises if it is missing.

The directory name contains a hyphen; load it with ``alego_pkg.load()`` (repo root) which registers it as
module ``alego_b200``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
SYNTH_DIR = os.path.join(_HERE, "synth")
HOST_DIR = os.path.join(_HERE, "host")
LIB_PATH = os.path.join(CSRC_DIR, "libalego_b200.so")
SYNTH_PATH = os.path.join(SYNTH_DIR, "libalego_synth.so")
HOST_PATH = os.path.join(HOST_DIR, "libalego_host.so")

OK, FEW_FEATURES, BAD_ARG, CUDA_ERROR, NOT_READY = 0, 1, -1, -2, -3
PRESET_VLP16_1800, PRESET_HDL64_1800, PRESET_HDL64_2048, PRESET_REFERENCE = 0, 1, 2, 3
LABEL_INVALID = 999999


class AlegoParams(C.Structure):
    """Mirror of struct AlegoParams (include/alego_b200.h) = runtime form of utility.h:50-73."""
    _fields_ = [
        ("n_scan", C.c_int32), ("horizon_scan", C.c_int32), ("ground_scan_id", C.c_int32),
        ("seg_valid_point_num", C.c_int32), ("seg_valid_line_num", C.c_int32), ("seg_min_cluster", C.c_int32),
        ("lo_surf_iters", C.c_int32), ("lo_corner_iters", C.c_int32), ("lm_outer_iters", C.c_int32),
        ("lm_max_iters", C.c_int32), ("reserved_i", C.c_int32 * 6),
        ("ang_res_x", C.c_double), ("ang_res_y", C.c_double), ("ang_bottom", C.c_double),
        ("sensor_mount_ang", C.c_double), ("seg_theta", C.c_double), ("nearest_feature_dist", C.c_double),
        ("huber_delta", C.c_double), ("less_flat_leaf", C.c_double), ("lm_corner_leaf", C.c_double),
        ("lm_surf_leaf", C.c_double), ("lm_outlier_leaf", C.c_double), ("reserved_d", C.c_double * 5),
    ]

    def copy(self):
        p = AlegoParams()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(AlegoParams))
        return p


class AlegoSolveReport(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_corner", C.c_int32), ("n_surf", C.c_int32), ("iterations", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class AlegoImuQueue(C.Structure):
    """The IMU ring buffers of laserOdometry.h:36-46 (caller-owned arrays of `length` doubles)."""
    _fields_ = [("length", C.c_int32), ("ptr_last", C.c_int32), ("ptr_last_iter", C.c_int32), ("reserved", C.c_int32)] + [
        (k, C.POINTER(C.c_double)) for k in ("time", "roll", "pitch", "yaw", "shift_x", "shift_y", "shift_z", "velo_x", "velo_y", "velo_z")]


class AlegoIcpResult(C.Structure):
    """What performLoopClosure reads from the ICP object (laserMapping.cpp:686-688)."""
    _fields_ = [("final_transformation", C.c_float * 16), ("fitness_score", C.c_double), ("has_converged", C.c_int32),
                ("iterations", C.c_int32), ("convergence_state", C.c_int32), ("n_correspondences", C.c_int32)]


class AlegoCloudInfo(C.Structure):
    """Mirror of msg/cloud_info.msg:1-12 (Header omitted)."""
    _fields_ = [("startRingIndex", C.POINTER(C.c_int32)), ("endRingIndex", C.POINTER(C.c_int32)),
                ("startOrientation", C.c_float), ("endOrientation", C.c_float), ("orientationDiff", C.c_float),
                ("size", C.c_int32), ("segmentedCloudGroundFlag", C.POINTER(C.c_uint8)),
                ("segmentedCloudColInd", C.POINTER(C.c_int32)), ("segmentedCloudRange", C.POINTER(C.c_float))]


def default_params(preset=PRESET_HDL64_1800):
    """Host-side copy of alego_default_params (no GPU / library needed)."""
    p = AlegoParams()
    p.seg_valid_point_num, p.seg_valid_line_num, p.seg_min_cluster = 5, 3, 30
    p.lo_surf_iters, p.lo_corner_iters, p.lm_outer_iters, p.lm_max_iters = 5, 5, 2, 20
    p.sensor_mount_ang, p.seg_theta, p.nearest_feature_dist, p.huber_delta = 0.0, 1.047, 25.0, 0.1
    p.less_flat_leaf, p.lm_corner_leaf, p.lm_surf_leaf, p.lm_outlier_leaf = 0.4, 0.4, 0.8, 1.0
    if preset == PRESET_VLP16_1800:
        p.n_scan, p.ang_res_x, p.ang_res_y, p.ang_bottom, p.ground_scan_id = 16, 0.2, 2.0, 15.0, 7
    elif preset == PRESET_HDL64_1800:
        p.n_scan, p.ang_res_x, p.ang_res_y, p.ang_bottom, p.ground_scan_id = 64, 0.2, 0.427, 24.9, 50
    elif preset == PRESET_HDL64_2048:
        p.n_scan, p.ang_res_x, p.ang_res_y, p.ang_bottom, p.ground_scan_id = 64, 360.0 / 2048.0, 0.427, 24.9, 50
    elif preset == PRESET_REFERENCE:
        p.n_scan, p.ang_res_x, p.ang_res_y, p.ang_bottom, p.ground_scan_id = 16, 0.09, 2.0, 15.0, 10
    else:
        raise ValueError("unknown preset")
    p.horizon_scan = int(360.0 / p.ang_res_x + 0.5)
    return p


# ------------------------------------------------------------------------------------------------------
# build helpers (used by __graft_entry__.build)
# ------------------------------------------------------------------------------------------------------
def build(verbose=False):
    """Compile libalego_b200.so (nvcc, sm_100a), the synthetic generator and the host shells, in-tree."""
    for d in (CSRC_DIR, SYNTH_DIR, HOST_DIR):
        if not os.path.exists(os.path.join(d, "Makefile")):
            continue
        r = subprocess.run(["make", "-C", d, "-j8"], capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout[-4000:], r.stderr[-4000:])
        if r.returncode != 0:
            raise RuntimeError("build failed in %s" % d)


_lib = None


def lib():
    """The CUDA library.  Fails loudly when it has not been built — there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libalego_b200.so is missing (%s): run __graft_entry__.build(); no CPU fallback exists" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    PF, PI, PD = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double)
    sig = {
        "alego_default_params": (C.c_int, [C.POINTER(AlegoParams), C.c_int]),
        "alego_create": (C.c_int, [C.POINTER(AlegoParams), C.c_int, C.c_int, C.c_int, C.POINTER(H)]),
        "alego_destroy": (None, [H]),
        "alego_last_error": (C.c_char_p, [H]),
        "alego_synchronize": (C.c_int, [H]),
        "alego_get_params": (C.c_int, [H, C.POINTER(AlegoParams)]),
        "alego_n_seq": (C.c_int, [H]),
        "alego_set_point_stride": (C.c_int, [H, C.c_int]),
        "alego_lm_assemble_map": (C.c_int, [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
        "alego_lm_get_map": (C.c_int, [H, C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)]),
        "alego_host_alloc": (C.c_void_p, [C.c_size_t]),
        "alego_host_free": (None, [C.c_void_p]),
        "alego_ip_process": (C.c_int, [H, C.c_void_p, C.c_void_p]),
        "alego_ip_upload": (C.c_int, [H, C.c_void_p, C.c_void_p]),
        "alego_ip_run": (C.c_int, [H]),
        "alego_stage_upload": (C.c_int, [H, C.c_int, C.c_void_p, C.c_void_p]),
        "alego_stage_select": (C.c_int, [H, C.c_int]),
        "alego_ip_get": (C.c_int, [H, C.c_int, C.POINTER(AlegoCloudInfo), C.c_void_p, C.c_void_p, PI, C.c_void_p]),
        "alego_lo_extract": (C.c_int, [H]),
        "alego_lc_icp": (C.c_int, [H, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_double,
                                   C.POINTER(AlegoIcpResult), C.c_void_p]),
        "alego_lo_adjust_distortion": (C.c_int, [H, PD, C.POINTER(AlegoImuQueue), C.c_double, PI]),
        "alego_lo_get_features": (C.c_int, [H, C.c_int, C.c_void_p, PI, C.c_void_p, PI, C.c_void_p, PI, C.c_void_p, PI, C.c_void_p]),
        "alego_lo_scan2scan": (C.c_int, [H, C.POINTER(AlegoSolveReport)]),
        "alego_lo_get_state": (C.c_int, [H, C.c_int, PD, PD, PD]),
        "alego_lo_set_params": (C.c_int, [H, C.c_int, PD]),
        "alego_lm_set_map": (C.c_int, [H, C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
        "alego_lm_set_scan": (C.c_int, [H, C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
        "alego_lm_set_odom": (C.c_int, [H, C.c_int, PD, PD]),
        "alego_lm_scan2map": (C.c_int, [H, C.POINTER(AlegoSolveReport)]),
        "alego_lm_get_state": (C.c_int, [H, C.c_int, PD, PD, PD, PD, PD]),
        "alego_lm_set_params": (C.c_int, [H, C.c_int, PD]),
        "alego_lm_get_downsampled": (C.c_int, [H, C.c_int, C.c_void_p, PI, C.c_void_p, PI, C.c_void_p, PI, C.c_void_p, PI]),
        "alego_pipeline_step": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_void_p]),
        "alego_pipeline_config": (C.c_int, [H, C.c_int, C.c_int, C.c_int]),
        "alego_pipeline_submit": (C.c_int, [H, C.c_void_p, C.c_void_p]),
        "alego_pipeline_collect": (C.c_int, [H, C.c_void_p]),
        "alego_pipeline_timeline": (C.c_int, [H, C.c_int, C.c_void_p]),
        "alego_voxel_grid": (C.c_int, [H, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, PI]),
        "alego_timer_mark": (C.c_int, [H, C.c_int]),
        "alego_timer_elapsed_ms": (C.c_int, [H, C.c_int, C.c_int, PF]),
        "alego_profile_enable": (C.c_int, [H, C.c_int]),
        "alego_profile_reset": (C.c_int, [H]),
        "alego_profile_count": (C.c_int, [H]),
        "alego_profile_get": (C.c_int, [H, C.c_int, C.c_char_p, C.POINTER(C.c_int64), PD]),
        "alego_launch_count": (C.c_int64, [H]),
        "alego_debug_get": (C.c_int64, [H, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "alego_default_params", "alego_create", "alego_destroy", "alego_last_error", "alego_synchronize", "alego_get_params",
    "alego_n_seq", "alego_set_point_stride", "alego_lm_assemble_map", "alego_lm_get_map", "alego_host_alloc", "alego_host_free", "alego_stage_upload", "alego_stage_select", "alego_ip_process", "alego_ip_upload", "alego_ip_run", "alego_ip_get", "alego_lo_extract",
    "alego_lo_get_features", "alego_lo_scan2scan", "alego_lo_get_state", "alego_lo_set_params", "alego_lm_set_map",
    "alego_lm_set_scan", "alego_lm_set_odom", "alego_lm_scan2map", "alego_lm_get_state", "alego_lm_set_params",
    "alego_lm_get_downsampled", "alego_pipeline_step", "alego_pipeline_config", "alego_pipeline_submit", "alego_pipeline_collect", "alego_pipeline_timeline", "alego_voxel_grid", "alego_timer_mark",
    "alego_timer_elapsed_ms", "alego_profile_enable", "alego_profile_reset", "alego_profile_count", "alego_profile_get",
    "alego_launch_count", "alego_debug_get", "alego_lo_adjust_distortion", "alego_lc_icp",
]


class AlegoError(RuntimeError):
    pass


_DEBUG_DTYPES = {
    "range_mat": np.float32, "full_cloud": np.float32, "ground_mat": np.uint8, "label_mat": np.int32,
    "startRingIndex": np.int32, "endRingIndex": np.int32, "segmentedCloudGroundFlag": np.uint8,
    "segmentedCloudColInd": np.int32, "segmentedCloudRange": np.float32, "segmented_cloud": np.float32,
    "outlier_cloud": np.float32, "orientation": np.float32, "cloud_curvature_abs": np.float32,
    "cloud_neighbor_picked": np.uint8, "cloud_neighbor_picked_occl": np.uint8, "cloud_label": np.int32,
    "cloud_sort_idx": np.int32, "sharp_idx": np.int32, "less_sharp_idx": np.int32, "flat_idx": np.int32,
    "sharp": np.float32, "flat": np.float32, "less_sharp": np.float32, "less_flat": np.float32, "corner_last": np.float32,
    "surf_last": np.float32, "lo_params": np.float64, "t_w_cur": np.float64, "r_w_cur": np.float64,
    "lo_surf_corr": np.int32, "lo_corner_corr": np.int32, "lo_trace": np.float64, "lm_trace": np.float64,
    "lm_params": np.float64, "lm_corner_ds": np.float32, "lm_surf_ds": np.float32, "lm_outlier_ds": np.float32,
    "lm_surf_total_ds": np.float32, "lm_edge": np.float64, "lm_plane": np.float64, "t_map2laser": np.float64,
    "r_map2laser": np.float64, "t_map2odom": np.float64, "r_map2odom": np.float64,
    "lm_report": np.uint8, "lo_report": np.uint8, "map_index_kind": np.int32,
}
_DEBUG_COLS = {"full_cloud": 4, "segmented_cloud": 4, "outlier_cloud": 4, "sharp": 4, "flat": 4, "less_sharp": 4, "less_flat": 4,
               "corner_last": 4, "surf_last": 4, "lo_surf_corr": 4, "lo_corner_corr": 3, "lo_trace": 7, "lm_trace": 7,
               "lm_corner_ds": 4, "lm_surf_ds": 4, "lm_outlier_ds": 4, "lm_surf_total_ds": 4, "lm_edge": 10, "lm_plane": 8}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Alego:
    """One handle = n_seq independent sequences advancing in lock-step on one GPU (see include/alego_b200.h)."""

    def __init__(self, params, n_seq=1, max_points=None, device=0):
        self.L = lib()
        self.params = params.copy()
        self.n_seq = n_seq
        self.R, self.Cc = params.n_scan, params.horizon_scan
        self.max_points = max_points or self.R * self.Cc
        self.point_stride = 4
        self.h = C.c_void_p()
        rc = self.L.alego_create(C.byref(self.params), device, n_seq, self.max_points, C.byref(self.h))
        if rc != OK:
            msg = self.L.alego_last_error(self.h).decode() if self.h else "alego_create failed"
            raise AlegoError("alego_create rc=%d: %s" % (rc, msg))

    def close(self):
        if getattr(self, "h", None):
            self.L.alego_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, allow_few=True):
        if rc == OK or (allow_few and rc == FEW_FEATURES):
            return rc
        raise AlegoError("rc=%d: %s" % (rc, self.L.alego_last_error(self.h).decode()))

    # ---- ImageProjection
    def set_point_stride(self, floats_per_point):
        """4 = (x, y, z, intensity) per input point (default), 3 = packed (x, y, z): see alego_set_point_stride."""
        self._chk(self.L.alego_set_point_stride(self.h, floats_per_point))
        self.point_stride = floats_per_point

    def pack_scans(self, scans):
        """list of (n_i,4) float32 arrays -> ([n_seq, max_points, point_stride] float32, [n_seq] int32)"""
        assert len(scans) == self.n_seq
        st = self.point_stride
        buf = np.zeros((self.n_seq, self.max_points, st), np.float32)
        n = np.zeros(self.n_seq, np.int32)
        for b, s in enumerate(scans):
            s = np.ascontiguousarray(s, np.float32).reshape(-1, 4)
            n[b] = len(s)
            buf[b, :len(s)] = s[:, :st]
        return buf, n

    def ip_process(self, buf, n):
        return self._chk(self.L.alego_ip_process(self.h, _ptr(buf), _ptr(n)))

    def ip_upload(self, buf, n):
        return self._chk(self.L.alego_ip_upload(self.h, _ptr(buf), _ptr(n)))

    def ip_run(self):
        return self._chk(self.L.alego_ip_run(self.h))

    def stage_upload(self, slot, buf, n):
        return self._chk(self.L.alego_stage_upload(self.h, slot, _ptr(buf), _ptr(n)))

    def stage_select(self, slot):
        return self._chk(self.L.alego_stage_select(self.h, slot))

    def ip_get(self, seq=0, labels=True):
        R, RC = self.R, self.R * self.Cc
        out = {"startRingIndex": np.zeros(R, np.int32), "endRingIndex": np.zeros(R, np.int32),
               "segmentedCloudGroundFlag": np.zeros(RC, np.uint8), "segmentedCloudColInd": np.zeros(RC, np.int32),
               "segmentedCloudRange": np.zeros(RC, np.float32)}
        info = AlegoCloudInfo()
        info.startRingIndex = out["startRingIndex"].ctypes.data_as(C.POINTER(C.c_int32))
        info.endRingIndex = out["endRingIndex"].ctypes.data_as(C.POINTER(C.c_int32))
        info.segmentedCloudGroundFlag = out["segmentedCloudGroundFlag"].ctypes.data_as(C.POINTER(C.c_uint8))
        info.segmentedCloudColInd = out["segmentedCloudColInd"].ctypes.data_as(C.POINTER(C.c_int32))
        info.segmentedCloudRange = out["segmentedCloudRange"].ctypes.data_as(C.POINTER(C.c_float))
        seg = np.zeros((RC, 4), np.float32)
        outl = np.zeros((RC, 4), np.float32)
        n_out = C.c_int32(0)
        lab = np.zeros(RC, np.int32) if labels else None
        self._chk(self.L.alego_ip_get(self.h, seq, C.byref(info), _ptr(seg), _ptr(outl), C.byref(n_out), _ptr(lab)))
        M = info.size
        for k in ("segmentedCloudGroundFlag", "segmentedCloudColInd", "segmentedCloudRange"):
            out[k] = out[k][:M]
        out.update(size=M, segmented_cloud=seg[:M], outlier_cloud=outl[:n_out.value], startOrientation=info.startOrientation,
                   endOrientation=info.endOrientation, orientationDiff=info.orientationDiff)
        if labels:
            out["label_mat"] = lab.reshape(R, self.Cc)
        return out

    # ---- LaserOdometry
    def lo_extract(self):
        return self._chk(self.L.alego_lo_extract(self.h))

    def lo_adjust_distortion(self, scan_times, queues, ptr_last, ptr_last_iter, scan_period=0.2):
        """adjustDistortion (laserOdometry.cpp:557-657) on the segmented cloud of every sequence, in place on the device.
        queues: per sequence a (10, len) float64 array (time, roll, pitch, yaw, shift xyz, velocity xyz); ptr_last /
        ptr_last_iter: per sequence imu_ptr_last_ / imu_ptr_last_iter_.  Returns (points visited [n_seq], new ptr_last_iter)."""
        B = self.n_seq
        t = np.ascontiguousarray(scan_times, np.float64).reshape(B)
        qs = [np.ascontiguousarray(q, np.float64).reshape(10, -1) for q in queues]
        arr = (AlegoImuQueue * B)()
        PD = C.POINTER(C.c_double)
        for b in range(B):
            arr[b].length, arr[b].ptr_last, arr[b].ptr_last_iter = qs[b].shape[1], int(ptr_last[b]), int(ptr_last_iter[b])
            for k, name in enumerate(("time", "roll", "pitch", "yaw", "shift_x", "shift_y", "shift_z", "velo_x", "velo_y", "velo_z")):
                setattr(arr[b], name, qs[b][k].ctypes.data_as(PD))
        n = np.zeros(B, np.int32)
        self._chk(self.L.alego_lo_adjust_distortion(self.h, t.ctypes.data_as(PD), arr, float(scan_period), n.ctypes.data_as(C.POINTER(C.c_int32))))
        return n, np.array([arr[b].ptr_last_iter for b in range(B)], np.int32)

    def lo_get_features(self, seq=0):
        R, RC = self.R, self.R * self.Cc
        s, ls, f = np.zeros(R * 12, np.int32), np.zeros(R * 120, np.int32), np.zeros(R * 24, np.int32)
        lf, lab = np.zeros((RC, 4), np.float32), np.zeros(RC, np.int32)
        ns, nls, nf, nlf = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._chk(self.L.alego_lo_get_features(self.h, seq, _ptr(s), C.byref(ns), _ptr(ls), C.byref(nls), _ptr(f), C.byref(nf), _ptr(lf),
                                               C.byref(nlf), _ptr(lab)))
        M = int(self.debug("segmentedCloudColInd", seq).shape[0])
        return {"sharp_idx": s[:ns.value], "less_sharp_idx": ls[:nls.value], "flat_idx": f[:nf.value], "less_flat": lf[:nlf.value],
                "cloud_label": lab[:M]}

    def lo_scan2scan(self):
        rep = (AlegoSolveReport * self.n_seq)()
        rc = self._chk(self.L.alego_lo_scan2scan(self.h, rep))
        return rc, [r.as_dict() for r in rep]

    def lo_get_state(self, seq=0):
        p, t, r = np.zeros(6), np.zeros(3), np.zeros(9)
        PD = C.POINTER(C.c_double)
        self._chk(self.L.alego_lo_get_state(self.h, seq, p.ctypes.data_as(PD), t.ctypes.data_as(PD), r.ctypes.data_as(PD)))
        return p, t, r.reshape(3, 3)

    def lo_set_params(self, seq, params):
        p = np.ascontiguousarray(params, np.float64)
        return self._chk(self.L.alego_lo_set_params(self.h, seq, p.ctypes.data_as(C.POINTER(C.c_double))))

    # ---- LaserMapping
    def lm_set_map(self, seq, corner, surf):
        corner = np.ascontiguousarray(corner, np.float32).reshape(-1, 4)
        surf = np.ascontiguousarray(surf, np.float32).reshape(-1, 4)
        return self._chk(self.L.alego_lm_set_map(self.h, seq, _ptr(corner), len(corner), _ptr(surf), len(surf)))

    def lm_set_scan(self, seq, corner, surf, outlier):
        a = [np.ascontiguousarray(x, np.float32).reshape(-1, 4) for x in (corner, surf, outlier)]
        return self._chk(self.L.alego_lm_set_scan(self.h, seq, _ptr(a[0]), len(a[0]), _ptr(a[1]), len(a[1]), _ptr(a[2]), len(a[2])))

    def lm_set_odom(self, seq, t, r):
        t = np.ascontiguousarray(t, np.float64)
        r = np.ascontiguousarray(r, np.float64).reshape(9)
        PD = C.POINTER(C.c_double)
        return self._chk(self.L.alego_lm_set_odom(self.h, seq, t.ctypes.data_as(PD), r.ctypes.data_as(PD)))

    def lm_set_params(self, seq, params):
        p = np.ascontiguousarray(params, np.float64)
        return self._chk(self.L.alego_lm_set_params(self.h, seq, p.ctypes.data_as(C.POINTER(C.c_double))))

    def lm_scan2map(self):
        rep = (AlegoSolveReport * self.n_seq)()
        rc = self._chk(self.L.alego_lm_scan2map(self.h, rep))
        return rc, [r.as_dict() for r in rep]

    def lm_get_state(self, seq=0):
        PD = C.POINTER(C.c_double)
        p, t1, r1, t2, r2 = np.zeros(6), np.zeros(3), np.zeros(9), np.zeros(3), np.zeros(9)
        self._chk(self.L.alego_lm_get_state(self.h, seq, *[a.ctypes.data_as(PD) for a in (p, t1, r1, t2, r2)]))
        return {"params": p, "t_map2laser": t1, "r_map2laser": r1.reshape(3, 3), "t_map2odom": t2, "r_map2odom": r2.reshape(3, 3)}

    # ---- whole path
    def pipeline_config(self, lm_every=1, rebuild_map_index_every_step=True, overlap_map_build=True, graphs=False):
        opt = (2 if graphs else 0) if overlap_map_build else -1
        return self._chk(self.L.alego_pipeline_config(self.h, lm_every, int(rebuild_map_index_every_step), opt))

    def pipeline_submit(self, buf, n):
        """Asynchronous pipeline_step: buf must be pinned (pinned_empty) and untouched until collected; <= 2 in flight."""
        return self._chk(self.L.alego_pipeline_submit(self.h, _ptr(buf), _ptr(n)))

    def pipeline_collect(self, want_poses=True):
        poses = np.zeros((self.n_seq, 12), np.float64) if want_poses else None
        self._chk(self.L.alego_pipeline_collect(self.h, _ptr(poses)))
        return poses

    def pipeline_timeline(self, on=True):
        """Switch the per-step timing events on/off; returns the timeline of the step collected last (ms since the first timed
        submit): [H2D starts, H2D done, front end starts, front end done, poses on the host]."""
        t = np.zeros(5, np.float32)
        self._chk(self.L.alego_pipeline_timeline(self.h, int(on), _ptr(t)))
        return t

    def pipeline_step(self, buf=None, n=None, want_poses=True):
        poses = np.zeros((self.n_seq, 12), np.float64) if want_poses else None
        self._chk(self.L.alego_pipeline_step(self.h, _ptr(buf), _ptr(n), _ptr(poses)))
        return poses

    def voxel_grid(self, pts, leaf):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
        out = np.zeros_like(pts)
        n = C.c_int32(0)
        self._chk(self.L.alego_voxel_grid(self.h, _ptr(pts), len(pts), leaf, _ptr(out), C.byref(n)))
        return out[:n.value]

    # ---- measurement
    def synchronize(self):
        return self._chk(self.L.alego_synchronize(self.h))

    def timer_mark(self, slot):
        return self._chk(self.L.alego_timer_mark(self.h, slot))

    def timer_elapsed_ms(self, a, b):
        ms = C.c_float(0)
        self._chk(self.L.alego_timer_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        return self._chk(self.L.alego_profile_enable(self.h, int(on)))

    def profile_reset(self):
        return self._chk(self.L.alego_profile_reset(self.h))

    def profile(self):
        out = {}
        for i in range(self.L.alego_profile_count(self.h)):
            name = C.create_string_buffer(64)
            n, ms = C.c_int64(0), C.c_double(0)
            self._chk(self.L.alego_profile_get(self.h, i, name, C.byref(n), C.byref(ms)))
            out[name.value.decode()] = (n.value, ms.value)
        return out

    def launch_count(self):
        return int(self.L.alego_launch_count(self.h))

    def debug(self, name, seq=0):
        nbytes = self.L.alego_debug_get(self.h, name.encode(), seq, None, 0)
        if nbytes < 0:
            raise AlegoError("debug_get(%s) rc=%d: %s" % (name, nbytes, self.L.alego_last_error(self.h).decode()))
        dt = np.dtype(_DEBUG_DTYPES[name])
        a = np.zeros(max(nbytes // dt.itemsize, 0), dt)
        if nbytes:
            got = self.L.alego_debug_get(self.h, name.encode(), seq, _ptr(a), nbytes)
            if got != nbytes:
                raise AlegoError("debug_get(%s) rc=%d: %s" % (name, got, self.L.alego_last_error(self.h).decode()))
        cols = _DEBUG_COLS.get(name)
        return a.reshape(-1, cols) if cols else a

    def lm_assemble_map(self, seq, corner_kfs, surf_kfs, outlier_kfs, poses6):
        """Local map of sequence seq from keyframe clouds (lists of (n,4) arrays) and their poses [K][6] (x,y,z,roll,pitch,yaw):
        alego_lm_assemble_map (laserMapping.cpp:194-323)."""
        def pack(clouds):
            clouds = [np.ascontiguousarray(c, np.float32).reshape(-1, 4) for c in clouds]
            ptrs = (C.c_void_p * max(len(clouds), 1))(*[c.ctypes.data for c in clouds])
            return clouds, ptrs, np.array([len(c) for c in clouds], np.int32)
        ck, cp, cn = pack(corner_kfs)
        sk, sp, sn = pack(surf_kfs)
        ok, op, on = pack(outlier_kfs)
        poses6 = np.ascontiguousarray(poses6, np.float32).reshape(-1, 6)
        self._cap_map = (int(cn.sum()), int(sn.sum() + on.sum()))
        return self._chk(self.L.alego_lm_assemble_map(self.h, seq, len(ck), cp, _ptr(cn), sp, _ptr(sn), op, _ptr(on), _ptr(poses6)))

    def lc_icp(self, source, target, max_corr_dist=100.0, max_iterations=100, transformation_epsilon=1e-6, fitness_epsilon=1e-6):
        """The ICP of performLoopClosure (laserMapping.cpp:667-688) on two (n,4) clouds: alego_lc_icp.  Returns a dict like
        oracle.binding.icp."""
        src = np.ascontiguousarray(source, np.float32).reshape(-1, 4)
        tgt = np.ascontiguousarray(target, np.float32).reshape(-1, 4)
        res = AlegoIcpResult()
        trace = np.zeros((max_iterations, 14))
        self._chk(self.L.alego_lc_icp(self.h, _ptr(src), len(src), _ptr(tgt), len(tgt), max_corr_dist, max_iterations,
                                      transformation_epsilon, fitness_epsilon, C.byref(res), _ptr(trace)))
        return {"T": np.array(res.final_transformation, np.float32).reshape(4, 4), "fitness": res.fitness_score,
                "converged": bool(res.has_converged), "state": res.convergence_state, "iterations": res.iterations,
                "n_correspondences": res.n_correspondences, "trace": trace[:res.iterations]}

    def lm_get_downsampled(self, seq=0):
        """(laser_corner_ds_, laser_surf_ds_, laser_outlier_ds_) of the last mapped sweep (laserMapping.cpp:325-346)."""
        return self.debug("lm_corner_ds", seq), self.debug("lm_surf_ds", seq), self.debug("lm_outlier_ds", seq)

    def lm_get_map(self, seq=0):
        """(corner_from_map_ds_, surf_from_map_ds_) of sequence seq as (n,4) arrays."""
        nc, ns = C.c_int32(0), C.c_int32(0)
        self._chk(self.L.alego_lm_get_map(self.h, seq, None, C.byref(nc), None, C.byref(ns)))
        c = np.zeros((max(nc.value, 1), 4), np.float32)
        s = np.zeros((max(ns.value, 1), 4), np.float32)
        self._chk(self.L.alego_lm_get_map(self.h, seq, _ptr(c), C.byref(nc), _ptr(s), C.byref(ns)))
        return c[:nc.value], s[:ns.value]

    def solve_report(self, stage, seq=0):
        """AlegoSolveReport of the last LaserOdometry ("lo") / LaserMapping ("lm") solve of sequence seq, as a dict."""
        raw = self.debug(stage + "_report", seq)
        return AlegoSolveReport.from_buffer_copy(raw.tobytes()).as_dict()


def pinned_empty(shape, dtype):
    """numpy array over cudaMallocHost memory (true async H2D).  The buffer is never freed (process lifetime)."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = lib().alego_host_alloc(max(n, 1))
    if not p:
        raise AlegoError("cudaMallocHost failed")
    arr = np.ctypeslib.as_array((C.c_uint8 * n).from_address(p)).view(dt).reshape(shape)
    return arr


# ------------------------------------------------------------------------------------------------------
# synthetic data (host only)
# ------------------------------------------------------------------------------------------------------
_synth = None


def synth_lib():
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_PATH):
            raise RuntimeError("libalego_synth.so is missing: run __graft_entry__.build()")
        S = C.CDLL(SYNTH_PATH)
        S.synth_world_create.restype = C.c_void_p
        S.synth_world_create.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double]
        S.synth_world_destroy.argtypes = [C.c_void_p]
        S.synth_render.restype = C.c_int
        S.synth_render.argtypes = [C.c_void_p, C.POINTER(AlegoParams), C.POINTER(C.c_double), C.c_uint64, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_void_p]
        S.synth_make_map.restype = C.c_int
        S.synth_make_map.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        _synth = S
    return _synth


class SynthWorld:
    """Seeded static world: ground plane + boxes + poles."""

    def __init__(self, seed=0, n_boxes=30, n_poles=20, extent=55.0, sensor_height=1.7):
        self.S = synth_lib()
        self.seed = seed
        self.w = C.c_void_p(self.S.synth_world_create(seed, n_boxes, n_poles, extent, sensor_height))

    def __del__(self):
        try:
            if self.w:
                self.S.synth_world_destroy(self.w)
                self.w = None
        except Exception:
            pass

    def render(self, params, pose=(0.0, 0.0, 0.0, 0.0), noise_seed=0, range_sigma=0.01, dropout=0.03, max_range=100.0, jitter_cells=0.0):
        """One sweep in the sensor frame, (n,4) float32 xyzi, column-major emission order."""
        R, Cc = params.n_scan, params.horizon_scan
        out = np.zeros((R * Cc, 4), np.float32)
        pose4 = (C.c_double * 4)(*pose)
        n = self.S.synth_render(self.w, C.byref(params), pose4, noise_seed, range_sigma, dropout, max_range, jitter_cells, _ptr(out))
        return out[:n].copy()

    def make_map(self, n_corner, n_surf, seed=0, radius=100.0, corner_step=0.4, surf_step=0.8, sigma=0.01):
        corner = np.zeros((n_corner, 4), np.float32)
        surf = np.zeros((n_surf, 4), np.float32)
        n = (C.c_int * 2)()
        self.S.synth_make_map(self.w, seed, n_corner, n_surf, radius, corner_step, surf_step, sigma, _ptr(corner), _ptr(surf), n)
        return corner[:n[0]].copy(), surf[:n[1]].copy()


def trajectory_pose(t, speed=0.15, yaw_rate=0.01, seed=0):
    """Smooth planar sensor trajectory: pose (x,y,z,yaw) of sweep t (t=0 is the origin = map/odom frame)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi)
    yaw = 0.0
    x = y = 0.0
    for k in range(int(t)):
        yaw_k = yaw_rate * k + 0.05 * np.sin(0.1 * k + ph) - 0.05 * np.sin(ph)
        x += speed * np.cos(yaw_k)
        y += speed * np.sin(yaw_k)
    yaw = yaw_rate * t + 0.05 * np.sin(0.1 * t + ph) - 0.05 * np.sin(ph)
    return (x, y, 0.0, yaw)
