"""File / --fast mode over several devices (vp_run_segments, vp_run_segments_cb, Upscaler::upscaleSegments; reference
src/main.cpp:222-355, :627-629): contiguous frame segments with temporal warm-up frames must reproduce the one-device,
frame-by-frame run bit for bit.  A device may be listed more than once (one context per entry), so the segment logic
is exercised on a one-GPU box too; with two or more GPUs the same test runs across real devices."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import ref_chain as RC
from oracle import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from tests import gpu_util as G  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SW, SH = 160, 96


def _one_device(vp, cfg, frames):
    outs = []
    with vp.Context(cfg) as ctx:
        for f in frames:
            src = G.to_dev(f)
            dst = G.dev_empty(cfg.dst_h, cfg.dst_w)
            ctx.call("vp_process_device", src.data_ptr(), SW * 3, dst.data_ptr(), cfg.dst_w * 3, None)
            ctx.call("vp_synchronize")
            outs.append(dst.cpu().numpy())
    return outs


def _device_lists():
    n = torch.cuda.device_count()
    lists = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [[0, 1], list(range(n))]
    return lists


def _cfg(vp, temporal, **kw):
    ocfg = RC.ChainConfig(dst_w=2 * SW, dst_h=2 * SH, temporal=temporal, **kw)
    return G.chain_cfg(vp, ocfg, SW, SH)


@pytest.mark.parametrize("temporal", [RC.TEMPORAL_BLEND_FRAMES, RC.TEMPORAL_MAIN_BLEND, RC.TEMPORAL_NONE])
def test_run_segments_bit_identical_to_one_device(vp, temporal):
    n = 13
    frames = synth.clip("scene", SW, SH, n)
    cfg = _cfg(vp, temporal)
    ref = _one_device(vp, cfg, frames)
    for devs in _device_lists():
        outs = [np.full((2 * SH, 2 * SW, 3), 0xEE, np.uint8) for _ in range(n)]
        sp = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        dp = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        rep = (vp.SegmentReport * len(devs))()
        rc = vp.lib.vp_run_segments(C.byref(cfg), (C.c_int * len(devs))(*devs), len(devs), n, sp, SW * 3, dp, 2 * SW * 3, rep)
        assert rc == 0, vp.lib.vp_segments_last_error()
        for a, b in zip(ref, outs):
            assert np.array_equal(a, b), devs
        warm = {RC.TEMPORAL_NONE: 0, RC.TEMPORAL_MAIN_BLEND: 1, RC.TEMPORAL_BLEND_FRAMES: 3}[temporal]
        for k, r in enumerate(rep):
            assert (r.start, r.stop) == (k * n // len(devs), (k + 1) * n // len(devs)) and r.status == 0
            assert r.first_frame == r.start - min(warm, r.start) and r.device == devs[k]


def test_run_segments_with_flow_and_scene_cuts(vp):
    """TemporalConsistency::process with its flow front end and a scene cut inside a segment's warm-up frames."""
    n = 12
    frames = synth.clip("scene", SW, SH, n)
    frames[5:] = synth.clip("noise", SW, SH, n - 5, 99)          # a hard cut at frame 5
    cfg = _cfg(vp, RC.TEMPORAL_BLEND_FRAMES, temporal_cfg=RC.TemporalConfig(use_flow=True, scene_detect=True))
    ref = _one_device(vp, cfg, frames)
    devs = [0, 0, 0] if torch.cuda.device_count() < 2 else [0, 1, 0]
    outs = [np.zeros((2 * SH, 2 * SW, 3), np.uint8) for _ in range(n)]
    sp = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
    dp = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    rc = vp.lib.vp_run_segments(C.byref(cfg), (C.c_int * 3)(*devs), 3, n, sp, SW * 3, dp, 2 * SW * 3, None)
    assert rc == 0, vp.lib.vp_segments_last_error()
    for i, (a, b) in enumerate(zip(ref, outs)):
        assert np.array_equal(a, b), i


def test_run_segments_callbacks(vp):
    n = 10
    frames = synth.clip("scene", SW, SH, n)
    cfg = _cfg(vp, RC.TEMPORAL_BLEND_FRAMES)
    ref = _one_device(vp, cfg, frames)
    got, order = {}, []

    def source(user, i, buf, step):
        C.memmove(buf, frames[i].ctypes.data, SH * SW * 3)
        return 0

    def sink(user, i, buf, step):
        got[i] = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), shape=(2 * SH, 2 * SW, 3)).copy()
        order.append(i)
        return 0

    devs = [0, 0] if torch.cuda.device_count() < 2 else [0, 1]
    rc = vp.lib.vp_run_segments_cb(C.byref(cfg), (C.c_int * 2)(*devs), 2, n, vp.FRAME_SOURCE(source), vp.FRAME_SINK(sink), None, None)
    assert rc == 0, vp.lib.vp_segments_last_error()
    assert sorted(got) == list(range(n))
    for seg in ([i for i in order if i < n // 2], [i for i in order if i >= n // 2]):
        assert seg == sorted(seg)                  # in order within a segment
    for i in range(n):
        assert np.array_equal(got[i], ref[i]), i
    # a failing source aborts the run with an error instead of hanging or crashing
    rc = vp.lib.vp_run_segments_cb(C.byref(cfg), (C.c_int * 2)(*devs), 2, n, vp.FRAME_SOURCE(lambda u, i, b, s: 1 if i == 3 else source(u, i, b, s)),
                                   vp.FRAME_SINK(sink), None, None)
    assert rc != 0 and b"source" in vp.lib.vp_segments_last_error()


def test_run_segments_rejects_bad_arguments(vp):
    cfg = _cfg(vp, RC.TEMPORAL_NONE)
    assert vp.lib.vp_run_segments(None, (C.c_int * 1)(0), 1, 0, None, 0, None, 0, None) == vp.VP_ERR_INVALID_ARG
    assert vp.lib.vp_run_segments(C.byref(cfg), (C.c_int * 1)(0), 1, 2, None, 0, None, 0, None) == vp.VP_ERR_INVALID_ARG
    assert vp.lib.vp_run_segments(C.byref(cfg), (C.c_int * 1)(0), 1, 0, None, 0, None, 0, None) == 0       # an empty clip is fine
    f = synth.clip("scene", SW, SH, 1)
    o = np.zeros((2 * SH, 2 * SW, 3), np.uint8)
    rc = vp.lib.vp_run_segments(C.byref(cfg), (C.c_int * 1)(64), 1, 1, (C.c_void_p * 1)(f[0].ctypes.data), SW * 3,
                                (C.c_void_p * 1)(o.ctypes.data), 2 * SW * 3, None)
    assert rc == vp.VP_ERR_INVALID_ARG and b"device" in vp.lib.vp_segments_last_error()


def test_upscaler_upscale_segments_cpp(vp, tmp_path):
    """Upscaler::upscaleSegments (the C++ face) against Upscaler::upscale frame by frame, through tests/cpp/api_driver."""
    exe = str(tmp_path / "api_driver")
    libdir = os.path.join(ROOT, "video_processor_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-DVP_FORCE_MAT_SHIM", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "api_driver.cpp"), "-o", exe, "-L", libdir, "-lvpb200",
                    f"-Wl,-rpath,{libdir}", "-lpthread"], check=True)
    n, sw, sh = 9, 128, 72
    frames = synth.clip("scene", sw, sh, n)
    inp = tmp_path / "in.bin"
    inp.write_bytes(b"".join(f.tobytes() for f in frames))

    def run(mode, env=None):
        outp = tmp_path / f"{mode}.bin"
        r = subprocess.run([exe, mode, str(inp), str(sw), str(sh), str(n), str(2 * sw), str(2 * sh), str(outp)],
                           capture_output=True, text=True, timeout=300, env={**os.environ, **(env or {})})
        assert r.returncode == 0, r.stdout + r.stderr
        return np.frombuffer(outp.read_bytes(), np.uint8)

    one = run("chain")
    devs = "0,0,0" if torch.cuda.device_count() < 2 else "0,1,0"
    assert np.array_equal(one, run("segments", {"VP_SEGMENT_DEVICES": devs}))
