"""Edge cases and full-size properties: tiny / ragged frames, non-integer ratios, BASELINE config C3
(4K->8K), and size-independent invariants of every stage at full size."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref_chain as RC
from oracle import restate as R
from oracle import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from tests import gpu_util as G  # noqa: E402


@pytest.fixture(scope="module")
def ctx(vp):
    c = vp.Context()
    yield c
    c.close()


@pytest.mark.parametrize("size,scale", [((1, 1), 2), ((2, 1), 2), ((1, 3), 4), ((3, 2), 4), ((5, 4), 2), ((129, 33), 2), ((127, 31), 4)])
def test_bicubic_tiny_and_ragged(vp, ctx, size, scale):
    import cv2
    w, h = size
    img = synth.frame("noise", w, h, 9)
    out = G.run_bicubic(vp, ctx, img, w * scale, h * scale)
    assert np.array_equal(out, R.resize_cubic(img, w * scale, h * scale))
    assert np.array_equal(out, cv2.resize(img, (w * scale, h * scale), interpolation=cv2.INTER_CUBIC))


@pytest.mark.parametrize("src,dst", [((100, 60), (300, 180)), ((77, 50), (100, 65)), ((64, 48), (64, 48)), ((90, 40), (91, 97))])
def test_bicubic_other_ratios_match_restatement(vp, ctx, src, dst):
    """x3, non-integer, identity and anisotropic ratios against OpenCV's own C++ path as restated
    (path='default' == the float vertical pass; stock cv2 may route these through IPP)."""
    img = synth.frame("scene", src[0], src[1], 3)
    out = G.run_bicubic(vp, ctx, img, dst[0], dst[1])
    assert np.array_equal(out, R.resize_cubic(img, dst[0], dst[1]))


@pytest.mark.parametrize("size,d", [((3, 3), 5), ((2, 7), 7), ((9, 1), 3), ((16, 16), 15), ((65, 17), 11)])
def test_bilateral_tiny(vp, ctx, size, d):
    w, h = size
    img = synth.frame("noise", w, h, 2)
    out = G.run_bilateral(vp, ctx, img, d, 40.0, 40.0)
    mx, n = G.diff_stats(out, R.bilateral_u8c3(img, d, 40.0, 40.0))
    assert mx <= 1, (mx, n)


def test_sharpen_minimum_size_and_rejects_smaller(vp, ctx):
    ocfg = RC.SharpenConfig(adaptive_sigma=False)
    cfg = G.sharpen_cfg(vp, ocfg)
    img = synth.frame("scene", 8, 8, 1)
    src, dst = G.to_dev(img), G.dev_empty(8, 8)
    ctx.call("vp_sharpen", C.byref(cfg), src.data_ptr(), G.pitch(src), dst.data_ptr(), G.pitch(dst), 8, 8, None)
    G.sync()
    mx, _ = G.diff_stats(dst.cpu().numpy(), RC.sharpen(RC.NumpyOps(), img, ocfg))
    assert mx <= 1
    assert vp.lib.vp_sharpen(ctx.handle, C.byref(cfg), src.data_ptr(), 24, dst.data_ptr(), 24, 7, 8, None) == vp.VP_ERR_UNSUPPORTED
    bad = G.sharpen_cfg(vp, RC.SharpenConfig(kernel_size=4, adaptive_sigma=False))
    assert vp.lib.vp_sharpen(ctx.handle, C.byref(bad), src.data_ptr(), 24, dst.data_ptr(), 24, 8, 8, None) == vp.VP_ERR_UNSUPPORTED


def test_c3_4k_to_8k(vp, ctx):
    """BASELINE config C3: bicubic bit-exact at 3840x2160 -> 7680x4320, then the whole chain (with the
    temporal stage) on two frames against the real OpenCV kernels."""
    import cv2
    sw, sh = 3840, 2160
    f0 = synth.frame("scene", sw, sh, 0)
    out = G.run_bicubic(vp, ctx, f0, 2 * sw, 2 * sh)
    assert np.array_equal(out, cv2.resize(f0, (2 * sw, 2 * sh), interpolation=cv2.INTER_CUBIC))
    del out
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_BLEND_FRAMES)
    oracle = RC.Chain(RC.Cv2Ops(), ocfg)
    cfg = G.chain_cfg(vp, ocfg, sw, sh)
    with vp.Context(cfg) as c:
        for t in range(2):
            f = synth.frame("scene", sw, sh, t)
            src = G.to_dev(f)
            dst = G.dev_empty(2 * sh, 2 * sw)
            c.call("vp_process_device", src.data_ptr(), G.pitch(src), dst.data_ptr(), G.pitch(dst), None)
            c.call("vp_synchronize")
            o, r = dst.cpu().numpy(), oracle.process(f)
            d = np.abs(o.astype(np.int16) - r.astype(np.int16))
            assert RC.psnr(o, r) >= 50.0 and d.max() <= 4, (RC.psnr(o, r), int(d.max()))


def test_full_size_invariants(vp, ctx):
    """Size-independent properties at the 4K working size: constant frames are fixed points of every
    stage, blending a frame with itself returns it, the bilateral commutes with a horizontal flip."""
    w, h = 3840, 2160
    const = np.full((h, w, 3), (40, 130, 221), np.uint8)
    assert np.array_equal(G.run_bilateral(vp, ctx, const, 5, 24.0, 24.0), const)
    scfg = G.sharpen_cfg(vp, RC.SharpenConfig(adaptive_sigma=False, preserve_tone=False))
    src, dst = G.to_dev(const), G.dev_empty(h, w)
    ctx.call("vp_sharpen", C.byref(scfg), src.data_ptr(), G.pitch(src), dst.data_ptr(), G.pitch(dst), w, h, None)
    G.sync()
    assert np.array_equal(dst.cpu().numpy(), const)
    img = synth.frame("scene", w, h, 1)
    d_img = G.to_dev(img)
    ptrs = (C.c_void_p * 3)(d_img.data_ptr(), d_img.data_ptr(), d_img.data_ptr())
    ctx.call("vp_temporal_blend", ptrs, 3, G.pitch(d_img), 0.6, dst.data_ptr(), G.pitch(dst), w, h, None)
    G.sync()
    mx, _ = G.diff_stats(dst.cpu().numpy(), img)
    assert mx <= 1        # sum(w_i*x)/sum(w_i) in float32
    ctx.call("vp_add_weighted", d_img.data_ptr(), G.pitch(d_img), 0.6, d_img.data_ptr(), G.pitch(d_img), 0.4, dst.data_ptr(), G.pitch(dst), w, h, None)
    G.sync()
    assert np.array_equal(dst.cpu().numpy(), img)
    a = G.run_bilateral(vp, ctx, img, 5, 24.0, 24.0)
    b = G.run_bilateral(vp, ctx, np.ascontiguousarray(img[:, ::-1]), 5, 24.0, 24.0)[:, ::-1]
    mx, n = G.diff_stats(a, b)
    assert mx <= 1 and n <= a.size // 1000      # tap order reverses -> float sums may differ in the last bit
    sums = torch.zeros(6, dtype=torch.int64, device="cuda")
    ctx.call("vp_mean_stddev_sums", d_img.data_ptr(), G.pitch(d_img), w, h, sums.data_ptr(), None)
    G.sync()
    x = img.reshape(-1, 3).astype(np.int64)
    assert sums.cpu().tolist() == x.sum(0).tolist() + (x * x).sum(0).tolist()


def test_context_lifecycle_does_not_leak(vp):
    """create / run / destroy every kind of chain context repeatedly: device memory returns to where it started"""
    import torch
    sw, sh = 96, 64
    variants = []
    for ms, asg, flow in [(False, False, False), (True, True, True), (False, True, False), (True, False, True)]:
        t = RC.TemporalConfig(use_flow=flow, scene_detect=flow)
        o = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_BLEND_FRAMES, temporal_cfg=t,
                           pre=RC.BilateralConfig(RC.PRE_PROCESSING, True, True, 7, 30.0, 30.0, ms, 30.0, 1.5, 2.0, ms, 3),
                           sharpen=RC.SharpenConfig(0.8, 1.2, 0.4, 30.0, 1.5, 5, True, True, asg),
                           post=RC.BilateralConfig(RC.POST_PROCESSING, True, True, 5, 20.0, 20.0, ms, 30.0, 1.2, 1.5, ms, 2))
        variants.append(o)
    frames = [G.to_dev(f) for f in synth.clip("scene", sw, sh, 3)]
    dst = G.dev_empty(2 * sh, 2 * sw)

    def cycle():
        for o in variants:
            with vp.Context(G.chain_cfg(vp, o, sw, sh)) as ctx:
                for f in frames:
                    ctx.call("vp_process_device", f.data_ptr(), sw * 3, dst.data_ptr(), G.pitch(dst), None)
                ctx.call("vp_synchronize")
    cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(5):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < (8 << 20), (free0, free1)
