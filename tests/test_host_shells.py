"""The ROS-free C++ host cores (a-lego-loam_b200/host: alego::ImageProjection / LaserOdometry / LaserMapping, driven by
alego_run) give the same poses as the ctypes path: both are thin callers of the same C ABI."""
import os
import struct
import subprocess

import numpy as np
import pytest


def test_host_library_builds_and_links(alego):
    assert os.path.exists(alego.HOST_PATH), "libalego_host.so missing: run __graft_entry__.build()"
    run = os.path.join(alego.HOST_DIR, "alego_run")
    assert os.path.exists(run) and os.access(run, os.X_OK)
    hdr = open(os.path.join(alego.HOST_DIR, "alego_host.h")).read()
    for cls in ("class ImageProjection", "class LaserOdometry", "class LaserMapping", "int onInit()"):
        assert cls in hdr
    r = subprocess.run([run], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def test_pointcloud2_wire_format_helpers(alego):
    """sensor_msgs/PointCloud2 <-> sweep buffers (the wire format either side of the path, SURVEY §8f N3 without ROS):
    arbitrary field offsets / point_step / row padding, both byte orders, packed-xyz and xyzi outputs, NaNs passed through."""
    import ctypes as C
    L = C.CDLL(alego.HOST_PATH)
    L.alego_host_decode_pointcloud2.restype = C.c_long
    L.alego_host_decode_pointcloud2.argtypes = [C.c_void_p, C.c_size_t] + [C.c_uint32] * 7 + [C.c_int32, C.c_int, C.c_void_p, C.c_int, C.c_size_t]
    L.alego_host_encode_pointcloud2_xyzi.restype = C.c_size_t
    L.alego_host_encode_pointcloud2_xyzi.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    rng = np.random.default_rng(1)
    W, H, step, pad = 37, 3, 22, 5     # velodyne-like record: x y z @0,4,8, intensity @16, ring uint16 @20; rows padded by 5 bytes
    xyzi = rng.normal(0, 20, (H, W, 4)).astype(np.float32)
    xyzi[1, 4, 1] = np.nan
    xyzi[2, 0, 0] = np.inf
    for big in (False, True):
        raw = np.zeros((H, W * step + pad), np.uint8)
        dt = ">f4" if big else "<f4"
        for r in range(H):
            rec = raw[r, :W * step].reshape(W, step)
            for k, off in enumerate((0, 4, 8, 16)):
                rec[:, off:off + 4] = xyzi[r, :, k].astype(dt).view(np.uint8).reshape(W, 4)
            rec[:, 20:22] = 7
        for stride in (3, 4):
            out = np.zeros((W * H, stride), np.float32)
            n = L.alego_host_decode_pointcloud2(raw.ctypes.data, raw.nbytes, W, H, step, W * step + pad, 0, 4, 8, 16, int(big), out.ctypes.data,
                                                stride, len(out))
            assert n == W * H
            assert np.array_equal(out, xyzi.reshape(-1, 4)[:, :stride], equal_nan=True)
        out = np.zeros((W * H, 4), np.float32)   # a message without an intensity field decodes to intensity 0
        assert L.alego_host_decode_pointcloud2(raw.ctypes.data, raw.nbytes, W, H, step, W * step + pad, 0, 4, 8, -1, int(big), out.ctypes.data, 4, len(out)) == W * H
        assert (out[:, 3] == 0).all()
        # inconsistent views are refused: capacity too small, point_step smaller than a field, row_step smaller than a row
        assert L.alego_host_decode_pointcloud2(raw.ctypes.data, raw.nbytes, W, H, step, W * step + pad, 0, 4, 8, 16, int(big), out.ctypes.data, 4, 5) == -1
        assert L.alego_host_decode_pointcloud2(raw.ctypes.data, raw.nbytes, W, H, 10, W * step + pad, 0, 4, 8, 16, int(big), out.ctypes.data, 4, len(out)) == -1
        assert L.alego_host_decode_pointcloud2(raw.ctypes.data, raw.nbytes, W, H, step, 10, 0, 4, 8, 16, int(big), out.ctypes.data, 4, len(out)) == -1
        # truncated message (data shorter than height * row_step says), and headers crafted to wrap 32-bit arithmetic
        assert L.alego_host_decode_pointcloud2(raw.ctypes.data, raw.nbytes - pad - 1, W, H, step, W * step + pad, 0, 4, 8, 16, int(big), out.ctypes.data, 4, len(out)) == -1
        assert L.alego_host_decode_pointcloud2(raw.ctypes.data, raw.nbytes - pad, W, H, step, W * step + pad, 0, 4, 8, 16, int(big), out.ctypes.data, 4, len(out)) == W * H
        assert L.alego_host_decode_pointcloud2(raw.ctypes.data, raw.nbytes, 2 ** 31, 2, 2, 0, 0, 4, 8, 16, int(big), out.ctypes.data, 4, 2 ** 40) == -1
        assert L.alego_host_decode_pointcloud2(raw.ctypes.data, raw.nbytes, W, H, step, W * step + pad, 2 ** 32 - 4, 4, 8, 16, int(big), out.ctypes.data, 4, len(out)) == -1
    # publisher side: PCL's PointXYZI layout (x y z @0,4,8, intensity @16, 32-byte records), and it decodes back
    flat = np.ascontiguousarray(xyzi.reshape(-1, 4))
    msg = np.zeros(len(flat) * 32, np.uint8)
    assert L.alego_host_encode_pointcloud2_xyzi(flat.ctypes.data, len(flat), msg.ctypes.data) == len(flat) * 32
    back = np.zeros_like(flat)
    assert L.alego_host_decode_pointcloud2(msg.ctypes.data, msg.nbytes, len(flat), 1, 32, 0, 0, 4, 8, 16, 0, back.ctypes.data, 4, len(back)) == len(flat)
    assert np.array_equal(back, flat, equal_nan=True)


@pytest.mark.gpu
def test_alego_run_matches_ctypes_pipeline(alego, tmp_path):
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seeds, n_sweeps = [3, 4], 4
    worlds = [alego.SynthWorld(seed=s) for s in seeds]
    maps = [w.make_map(4000, 20000, seed=s, radius=60.0) for w, s in zip(worlds, seeds)]
    sweeps = [[w.render(P, alego.trajectory_pose(t, seed=s), noise_seed=31 * s + t) for w, s in zip(worlds, seeds)] for t in range(n_sweeps)]
    fin, fout = str(tmp_path / "sweeps.bin"), str(tmp_path / "poses.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<4i", alego.PRESET_VLP16_1800, len(seeds), n_sweeps, 1))
        for cm, sm in maps:
            f.write(struct.pack("<2i", len(cm), len(sm)))
            f.write(cm.astype("<f4").tobytes())
            f.write(sm.astype("<f4").tobytes())
        for t in range(n_sweeps):
            for s in sweeps[t]:
                f.write(struct.pack("<i", len(s)))
                f.write(s.astype("<f4").tobytes())
    r = subprocess.run([os.path.join(alego.HOST_DIR, "alego_run"), fin, fout], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(fout, np.float64).reshape(n_sweeps, len(seeds), 12)
    a = alego.Alego(P, n_seq=len(seeds))
    for b, (cm, sm) in enumerate(maps):
        a.lm_set_map(b, cm, sm)
    a.pipeline_config(lm_every=1)
    for t in range(n_sweeps):
        buf, n = a.pack_scans(sweeps[t])
        poses = a.pipeline_step(buf, n)
        # alego_run layout: LM params[6], LO t_w_cur[3], LM t_map2laser[3]; pipeline_step: t_map2laser[3], params[6], t_w[3]
        assert np.array_equal(got[t][:, 0:6], poses[:, 3:9]), t
        assert np.array_equal(got[t][:, 6:9], poses[:, 9:12]), t
        assert np.array_equal(got[t][:, 9:12], poses[:, 0:3]), t
    a.close()


@pytest.mark.gpu
def test_alego_run_closed_loop_keyframes(alego, tmp_path):
    """alego_run with keyframe_every=1: the host core's keyframe deque + device local-map assembly (N1) gives the same
    trajectory as the same steps driven through ctypes (get downsampled clouds -> alego_lm_assemble_map)."""
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed, n_sweeps = 5, 5
    w = alego.SynthWorld(seed=seed)
    cm, sm = w.make_map(4000, 20000, seed=seed, radius=60.0)
    sweeps = [w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=17 * seed + t) for t in range(n_sweeps)]
    fin, fout = str(tmp_path / "sweeps.bin"), str(tmp_path / "poses.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<4i", alego.PRESET_VLP16_1800, 1, n_sweeps, 1))
        f.write(struct.pack("<2i", len(cm), len(sm)))
        f.write(cm.astype("<f4").tobytes())
        f.write(sm.astype("<f4").tobytes())
        for s in sweeps:
            f.write(struct.pack("<i", len(s)))
            f.write(s.astype("<f4").tobytes())
    r = subprocess.run([os.path.join(alego.HOST_DIR, "alego_run"), fin, fout, "0", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(fout, np.float64).reshape(n_sweeps, 1, 12)
    a = alego.Alego(P, n_seq=1)
    a.lm_set_map(0, cm, sm)
    kfs, poses6 = [], []
    moved = False
    for t in range(n_sweeps):
        buf, n = a.pack_scans([sweeps[t]])
        a.ip_process(buf, n)
        a.lo_extract()
        a.lo_scan2scan()
        a.lm_scan2map()
        prm = a.lm_get_state(0)["params"]
        assert np.array_equal(got[t, 0, 0:6], prm), t
        kfs.append(a.lm_get_downsampled(0))
        poses6.append(np.asarray(prm, np.float32))
        a.lm_assemble_map(0, [k[0] for k in kfs], [k[1] for k in kfs], [k[2] for k in kfs], np.stack(poses6))
        nc, ns = (len(x) for x in a.lm_get_map(0))
        assert nc > 50 and ns > 500
        moved = moved or abs(prm[0]) > 0.05
    assert moved
    a.close()


def test_pointcloud2_decoder_properties(alego):
    """Property test (hypothesis) of alego::decode_pointcloud2: for any record layout the decoded floats are exactly the bytes at the
    field offsets; any view whose geometry does not fit (point_step too small for a field, row_step shorter than a row, capacity too
    small) is refused with -1 and nothing is written."""
    import ctypes as C
    from hypothesis import given, settings, strategies as st
    L = C.CDLL(alego.HOST_PATH)
    L.alego_host_decode_pointcloud2.restype = C.c_long
    L.alego_host_decode_pointcloud2.argtypes = [C.c_void_p, C.c_size_t] + [C.c_uint32] * 7 + [C.c_int32, C.c_int, C.c_void_p, C.c_int, C.c_size_t]

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 40), st.integers(0, 4), st.integers(1, 48), st.integers(0, 9), st.lists(st.integers(0, 44), min_size=4, max_size=4),
           st.booleans(), st.sampled_from([3, 4]), st.integers(0, 2 ** 32 - 1), st.integers(-3, 3))
    def run(width, height, step, pad, offs, big, stride, seed, cap_delta):
        rng = np.random.default_rng(seed)
        row_step = width * step + pad
        raw = rng.integers(0, 256, max(height * row_step, 1), dtype=np.uint8)
        # random bytes include NaN / inf / denormal patterns: they must pass through unchanged
        n = width * height
        cap = max(n + cap_delta, 0)
        out = np.full((max(cap, 1), stride), -7.0, np.float32)
        off_i = offs[3] if seed % 3 else -1
        need = (height - 1) * row_step + width * step if n else 0
        size = max(need - (seed % 5 == 0), 0)  # every fifth example: one byte short of what the geometry needs
        got = L.alego_host_decode_pointcloud2(raw.ctypes.data, size, width, height, step, row_step, offs[0], offs[1], offs[2], off_i, int(big),
                                              out.ctypes.data, stride, cap)
        fits = step >= max(offs[0], offs[1], offs[2], max(off_i, 0)) + 4 and size >= need
        if n == 0:
            assert got == 0
            return
        if not fits or n > cap:
            assert got == -1 and (out == -7.0).all()
            return
        assert got == n
        dt = ">f4" if big else "<f4"
        for r in range(height):
            for c in (0, width - 1):
                base = r * row_step + c * step
                for k in range(3):
                    want = raw[base + offs[k]:base + offs[k] + 4].view(dt)[0]
                    assert out[r * width + c, k].tobytes() == np.float32(want).tobytes()
                if stride == 4:
                    want = raw[base + off_i:base + off_i + 4].view(dt)[0] if off_i >= 0 else np.float32(0)
                    assert out[r * width + c, 3].tobytes() == np.float32(want).tobytes()

    run()
