"""The C-ABI library loads and exports every symbol include/alego_b200.h declares (no GPU work)."""
import ctypes
import os
import re


def test_library_exports_every_declared_symbol(alego):
    hdr = open(os.path.join(os.path.dirname(alego.CSRC_DIR), "..", "include", "alego_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(alego_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 34
    L = ctypes.CDLL(alego.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    for s in alego.EXPORTED_SYMBOLS:
        assert s in declared


def test_struct_layout_matches_header(alego):
    # 16 int32 + 16 doubles
    assert ctypes.sizeof(alego.AlegoParams) == 16 * 4 + 16 * 8
    assert ctypes.sizeof(alego.AlegoSolveReport) == 4 * 4 + 2 * 8
    L = ctypes.CDLL(alego.LIB_PATH)
    L.alego_default_params.argtypes = [ctypes.c_void_p, ctypes.c_int]
    for preset in range(4):
        p = alego.AlegoParams()
        assert L.alego_default_params(ctypes.byref(p), preset) == 0
        q = alego.default_params(preset)
        assert bytes(p) == bytes(q)
    assert L.alego_default_params(ctypes.byref(p), 99) == alego.BAD_ARG


def test_reference_preset_matches_utility_h(alego):
    p = alego.default_params(alego.PRESET_REFERENCE)  # include/alego/utility.h:50-65
    assert (p.n_scan, p.horizon_scan, p.ground_scan_id) == (16, 4000, 10)
    assert (p.ang_res_x, p.ang_res_y, p.ang_bottom) == (0.09, 2.0, 15.0)
    assert (p.seg_theta, p.seg_valid_point_num, p.seg_valid_line_num, p.nearest_feature_dist) == (1.047, 5, 3, 25.0)


def test_null_and_bad_arguments_do_not_crash(alego):
    L = ctypes.CDLL(alego.LIB_PATH)
    L.alego_last_error.restype = ctypes.c_char_p
    L.alego_last_error.argtypes = [ctypes.c_void_p]
    assert L.alego_last_error(None) == b"null handle"
    L.alego_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    h = ctypes.c_void_p()
    assert L.alego_create(None, 0, 1, 10, ctypes.byref(h)) == alego.BAD_ARG
    p = alego.default_params(0)
    assert L.alego_create(ctypes.byref(p), 0, 0, 10, ctypes.byref(h)) == alego.BAD_ARG
    p.n_scan = 100000
    assert L.alego_create(ctypes.byref(p), 0, 1, 10, ctypes.byref(h)) == alego.BAD_ARG
    L.alego_ip_run.argtypes = [ctypes.c_void_p]
    assert L.alego_ip_run(None) == alego.BAD_ARG
    L.alego_destroy.argtypes = [ctypes.c_void_p]
    L.alego_destroy(None)
    # the "next row" entry points refuse a null handle before touching CUDA
    L.alego_lc_icp.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_double,
                               ctypes.c_int32, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    res = alego.AlegoIcpResult()
    assert L.alego_lc_icp(None, None, 0, None, 0, 100.0, 100, 1e-6, 1e-6, ctypes.byref(res), None) == alego.BAD_ARG
    L.alego_lo_adjust_distortion.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
    assert L.alego_lo_adjust_distortion(None, None, None, 0.2, None) == alego.BAD_ARG


def test_next_row_struct_layouts(alego):
    # AlegoImuQueue: 4 int32 + 10 pointers; AlegoIcpResult: 16 floats + double + 4 int32 (include/alego_b200.h)
    assert ctypes.sizeof(alego.AlegoImuQueue) == 4 * 4 + 10 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(alego.AlegoIcpResult) == 16 * 4 + 8 + 4 * 4
    assert alego.AlegoIcpResult.fitness_score.offset == 64 and alego.AlegoIcpResult.has_converged.offset == 72


def test_every_entry_point_refuses_a_null_handle(alego):
    """Errors are return codes, nothing throws or crashes across the boundary: every exported function called with a
    null handle and zeroed arguments returns ALEGO_BAD_ARG (alego_last_error: its fixed string)."""
    L = alego.lib()
    skip = {"alego_host_alloc", "alego_host_free", "alego_destroy", "alego_default_params", "alego_create"}
    for name in alego.EXPORTED_SYMBOLS:
        if name in skip:
            continue
        fn = getattr(L, name)
        args = []
        for t in fn.argtypes:
            if t in (ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t):
                args.append(0)
            elif t in (ctypes.c_double, ctypes.c_float):
                args.append(0.0)
            else:
                args.append(None)
        r = fn(*args)
        if name == "alego_last_error":
            assert r == b"null handle"
        else:
            assert r == alego.BAD_ARG, (name, r)
