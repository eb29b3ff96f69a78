"""The C++ API (include/*.h, the classes pipeline.cpp / main.cpp / test_enhancements.cpp use) driven by
tests/cpp/api_driver.cpp on the GPU and compared with the oracle chain."""
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

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SW, SH, N = 128, 72, 5


@pytest.fixture(scope="module")
def driver(vp, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("bin") / "api_driver")
    libdir = os.path.join(ROOT, "video_processor_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-DVP_FORCE_MAT_SHIM", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "api_driver.cpp"), "-o", exe, "-L", libdir, "-lvpb200",
                    f"-Wl,-rpath,{libdir}", "-lpthread"], check=True)
    return exe


def run(driver, tmp_path, mode, frames, tw, th):
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    inp.write_bytes(b"".join(f.tobytes() for f in frames))
    r = subprocess.run([driver, mode, str(inp), str(SW), str(SH), str(len(frames)), str(tw), str(th), str(outp)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.frombuffer(outp.read_bytes(), np.uint8).reshape(len(frames), th, tw, 3)


def ref_defaults(**kw):
    """Upscaler::initializeEnhancements as the reference writes it (src/upscaler.cpp:666-758): selective and
    multi-scale on for both bilateral stages, adaptive_sigma on for the sharpening (all with repaired semantics), and
    TemporalConsistency::process with its Farneback flow / warp / reliability-mask front end."""
    return RC.ChainConfig(
        pre=RC.BilateralConfig(RC.PRE_PROCESSING, True, True, 7, 30.0, 30.0, True, 30.0, 1.5, 2.0, True, 3),
        post=RC.BilateralConfig(RC.POST_PROCESSING, True, True, 5, 20.0, 20.0, True, 30.0, 1.2, 1.5, True, 2),
        sharpen=RC.SharpenConfig(0.8, 1.2, 0.4, 30.0, 1.5, 5, True, True, True),
        temporal_cfg=RC.TemporalConfig(use_flow=True, scene_detect=True), **kw)


def close(o, r):
    d = np.abs(o.astype(int) - r.astype(int))
    return RC.psnr(o, r) >= 50.0 and d.max() <= 4


@pytest.mark.parametrize("mode", ["chain", "pipelined", "processor", "stages"])
def test_upscaler_chain_modes(driver, tmp_path, mode):
    frames = synth.clip("scene", SW, SH, N)
    outs = run(driver, tmp_path, mode, frames, 2 * SW, 2 * SH)
    if mode == "stages":      # the stage objects of "stages" carry plain configs; TemporalConsistency always runs its flow front end
        ocfg = RC.ChainConfig(dst_w=2 * SW, dst_h=2 * SH, temporal=RC.TEMPORAL_BLEND_FRAMES,
                              temporal_cfg=RC.TemporalConfig(use_flow=True, scene_detect=True))
    else:
        ocfg = ref_defaults(dst_w=2 * SW, dst_h=2 * SH, temporal=RC.TEMPORAL_BLEND_FRAMES)
    oracle = RC.Chain(RC.NumpyOps(), ocfg)
    for o, f in zip(outs, frames):
        assert close(o, oracle.process(f))


def test_resize_cubic_bit_exact(driver, tmp_path):
    import cv2
    frames = synth.clip("noise", SW, SH, 3)
    outs = run(driver, tmp_path, "resize", frames, 4 * SW, 4 * SH)
    for o, f in zip(outs, frames):
        assert np.array_equal(o, cv2.resize(f, (4 * SW, 4 * SH), interpolation=cv2.INTER_CUBIC))


def test_upscaler_bicubic_is_the_reference_wrapper(driver, tmp_path):
    """Upscaler(BICUBIC) == CPUImpl::upscale: pre-blur, INTER_CUBIC, enhanceBicubicResult (src/upscaler.cpp:75-148)."""
    frames = synth.clip("scene", SW, SH, 3)
    outs = run(driver, tmp_path, "bicubic", frames, 2 * SW, 2 * SH)
    for o, f in zip(outs, frames):
        r = RC.cpuimpl_bicubic(RC.Cv2Ops(), f, 2 * SW, 2 * SH)
        d = np.abs(o.astype(int) - r.astype(int))
        assert d.max() <= 1 and (d > 0).sum() <= o.size // 500


def test_upscaler_toggles(driver, tmp_path):
    frames = synth.clip("scene", SW, SH, 3)
    outs = run(driver, tmp_path, "toggles", frames, 2 * SW, 2 * SH)
    for i, (o, f) in enumerate(zip(outs, frames)):
        cfg = ref_defaults(dst_w=2 * SW, dst_h=2 * SH, temporal=RC.TEMPORAL_NONE, use_pre=i < 1, use_post=i < 1, use_sharpen=i < 2)
        assert close(o, RC.Chain(RC.NumpyOps(), cfg).process(f)), i


def test_video_enhancer_facade(driver, tmp_path):
    frames = synth.clip("scene", SW, SH, 3)
    outs = run(driver, tmp_path, "enhancer", frames, SW, SH)
    # the facade runs the north-star chain (plain branches) at the input resolution
    oracle = RC.Chain(RC.NumpyOps(), RC.ChainConfig(dst_w=SW, dst_h=SH, temporal=RC.TEMPORAL_BLEND_FRAMES))
    for o, f in zip(outs, frames):
        assert close(o, oracle.process(f))
