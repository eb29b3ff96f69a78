"""Guard-band tests: every entry point that writes a caller buffer is run with its output inside a larger canary-filled
allocation, with padded (and, for byte frames, unaligned) rows.  The bytes before the first row, after the last row and
in every row's padding must be untouched, and the result must be identical to the one written into a tight buffer -
a tile-wide store (TMA bulk stores, 16-byte vector stores) that ran past a ragged edge would show up here.
(compute-sanitizer is not available on the GPU pool.)"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ref_chain as RC
from oracle import synth
from tests import gpu_util as G

pytestmark = pytest.mark.gpu

CANARY = 0xA5


class Guarded:
    """h rows of row_bytes inside a canary buffer: lead bytes | (row | pad) * h | lead bytes"""

    def __init__(self, h, row_bytes, pad, lead):
        self.h, self.row_bytes, self.pitch, self.lead = h, row_bytes, row_bytes + pad, lead
        self.buf = torch.full((2 * lead + h * self.pitch,), CANARY, dtype=torch.uint8, device="cuda")
        self.ptr = self.buf.data_ptr() + lead

    def body(self):
        return self.buf[self.lead:self.lead + self.h * self.pitch].view(self.h, self.pitch)

    def data(self, dtype=np.uint8):
        raw = self.body()[:, :self.row_bytes].contiguous().cpu().numpy()
        return raw.view(dtype)

    def check(self, what):
        G.sync()
        assert bool((self.buf[:self.lead] == CANARY).all()), f"{what}: wrote before the first row"
        assert bool((self.buf[self.lead + self.h * self.pitch:] == CANARY).all()), f"{what}: wrote after the last row"
        assert bool((self.body()[:, self.row_bytes:] == CANARY).all()), f"{what}: wrote into the row padding"


# (pad, lead): aligned rows; rows that break 4/16-byte alignment (byte frames only)
U8_LAYOUTS = [(64, 4096), (7, 4099)]
F32_LAYOUTS = [(64, 4096), (12, 4100)]
SIZES = [(67, 45), (130, 70), (256, 64)]


def frame(w, h, t=0):
    return synth.frame("scene", w, h, t)


@pytest.fixture(scope="module")
def ctx(vp):
    with vp.Context() as c:
        yield c


@pytest.mark.parametrize("pad,lead", U8_LAYOUTS)
@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("scale", [2, 4, 1.5])
def test_bicubic_stays_inside(vp, ctx, w, h, scale, pad, lead):
    img = frame(w, h)
    dw, dh = int(w * scale), int(h * scale)
    tight = G.run_bicubic(vp, ctx, img, dw, dh)
    src = G.to_dev(img)
    g = Guarded(dh, dw * 3, pad, lead)
    ctx.call("vp_bicubic", src.data_ptr(), G.pitch(src), w, h, g.ptr, g.pitch, dw, dh, None)
    g.check("vp_bicubic")
    assert np.array_equal(g.data().reshape(dh, dw, 3), tight)


@pytest.mark.parametrize("pad,lead", U8_LAYOUTS)
@pytest.mark.parametrize("w,h", SIZES)
def test_u8_stages_stay_inside(vp, ctx, w, h, pad, lead):
    img = frame(w, h)
    src = G.to_dev(img)

    def run(name, *args_before_dst, tail=(w, h, None), oh=h, ow=w):
        t = G.dev_empty(oh, ow)
        ctx.call(name, *args_before_dst, t.data_ptr(), G.pitch(t), *tail)
        G.sync()
        g = Guarded(oh, ow * 3, pad, lead)
        ctx.call(name, *args_before_dst, g.ptr, g.pitch, *tail)
        g.check(name)
        assert np.array_equal(g.data().reshape(oh, ow, 3), t.cpu().numpy()), name

    for d in (5, 9):
        t = G.dev_empty(h, w)
        ctx.call("vp_bilateral", src.data_ptr(), G.pitch(src), t.data_ptr(), G.pitch(t), w, h, d, 30.0, 30.0, None)
        g = Guarded(h, w * 3, pad, lead)
        ctx.call("vp_bilateral", src.data_ptr(), G.pitch(src), g.ptr, g.pitch, w, h, d, 30.0, 30.0, None)
        g.check("vp_bilateral")
        assert np.array_equal(g.data().reshape(h, w, 3), t.cpu().numpy())
    for sel, ms in ((False, False), (True, False), (False, True)):
        cfg = G.bilateral_cfg(vp, RC.BilateralConfig(selective=sel, use_multiscale=ms))
        run("vp_selective_bilateral", cfg, src.data_ptr(), G.pitch(src))
    for kw in ({"adaptive_sigma": False}, {"adaptive_sigma": True}, {"adaptive_sigma": False, "kernel_size": 3}):
        cfg = G.sharpen_cfg(vp, RC.SharpenConfig(**kw))
        run("vp_sharpen", cfg, src.data_ptr(), G.pitch(src))
    run("vp_add_weighted", src.data_ptr(), G.pitch(src), 0.6, src.data_ptr(), G.pitch(src), 0.4)
    for n in (2, 4):
        ptrs = (C.c_void_p * n)(*[src.data_ptr()] * n)
        run("vp_temporal_blend", ptrs, n, G.pitch(src), 0.6)
    dw, dh = (w + 1) // 2, (h + 1) // 2
    run("vp_pyr_down", src.data_ptr(), G.pitch(src), w, h, tail=(None,), oh=dh, ow=dw)
    small = G.to_dev(frame(dw, dh, 3))
    run("vp_pyr_up", small.data_ptr(), G.pitch(small), dw, dh, tail=(w, h, None))


@pytest.mark.parametrize("pad,lead", F32_LAYOUTS)
@pytest.mark.parametrize("w,h", SIZES)
def test_float_planes_stay_inside(vp, ctx, w, h, pad, lead):
    img = frame(w, h)
    src = G.to_dev(img)

    def plane():
        return torch.empty((h, w), dtype=torch.float32, device="cuda")

    def run(name, *args_before):
        t = plane()
        ctx.call(name, *args_before, t.data_ptr(), G.pitch(t), w, h, None)
        G.sync()
        g = Guarded(h, w * 4, pad, lead)
        ctx.call(name, *args_before, g.ptr, g.pitch, w, h, None)
        g.check(name)
        assert np.array_equal(g.data(np.float32), t.cpu().numpy()), name

    run("vp_detail_mask", 30.0, src.data_ptr(), G.pitch(src))
    run("vp_edge_mask", G.sharpen_cfg(vp, RC.SharpenConfig()), src.data_ptr(), G.pitch(src))
    run("vp_sigma_map", src.data_ptr(), G.pitch(src))
    rng = np.random.default_rng(5)
    fx = G.to_dev((rng.standard_normal((h, w)) * 12).astype(np.float32))
    fy = G.to_dev((rng.standard_normal((h, w)) * 12).astype(np.float32))
    # vp_flow_reliability: the flow planes share the mask's pitch
    t = plane()
    ctx.call("vp_flow_reliability", fx.data_ptr(), fy.data_ptr(), G.pitch(fx), 15.0, t.data_ptr(), G.pitch(t), w, h, None)
    gfx, gfy, g = (Guarded(h, w * 4, pad, lead) for _ in range(3))
    gfx.body()[:, :w * 4] = fx.view(torch.uint8).view(h, w * 4)
    gfy.body()[:, :w * 4] = fy.view(torch.uint8).view(h, w * 4)
    ctx.call("vp_flow_reliability", gfx.ptr, gfy.ptr, gfx.pitch, 15.0, g.ptr, g.pitch, w, h, None)
    g.check("vp_flow_reliability"); gfx.check("flow x"); gfy.check("flow y")
    assert np.array_equal(g.data(np.float32), t.cpu().numpy())
    # warp: byte output driven by float planes
    t = G.dev_empty(h, w)
    ctx.call("vp_warp_flow", src.data_ptr(), G.pitch(src), fx.data_ptr(), fy.data_ptr(), G.pitch(fx), t.data_ptr(), G.pitch(t), w, h, None)
    g = Guarded(h, w * 3, 7, 4099)
    ctx.call("vp_warp_flow", src.data_ptr(), G.pitch(src), fx.data_ptr(), fy.data_ptr(), G.pitch(fx), g.ptr, g.pitch, w, h, None)
    g.check("vp_warp_flow")
    assert np.array_equal(g.data().reshape(h, w, 3), t.cpu().numpy())


@pytest.mark.parametrize("pad,lead", F32_LAYOUTS)
@pytest.mark.parametrize("w,h", [(130, 70), (256, 144)])
def test_optical_flow_stays_inside(vp, ctx, w, h, pad, lead):
    a, b = G.to_dev(frame(w, h, 0)), G.to_dev(frame(w, h, 1))
    tc = G.temporal_cfg(vp, RC.TemporalConfig())
    tx = torch.empty((h, w), dtype=torch.float32, device="cuda")
    ty = torch.empty_like(tx)
    ctx.call("vp_optical_flow", tc, a.data_ptr(), b.data_ptr(), G.pitch(a), tx.data_ptr(), ty.data_ptr(), G.pitch(tx), w, h, None)
    gx, gy = Guarded(h, w * 4, pad, lead), Guarded(h, w * 4, pad, lead)
    ctx.call("vp_optical_flow", tc, a.data_ptr(), b.data_ptr(), G.pitch(a), gx.ptr, gy.ptr, gx.pitch, w, h, None)
    gx.check("vp_optical_flow x"); gy.check("vp_optical_flow y")
    assert np.array_equal(gx.data(np.float32), tx.cpu().numpy()) and np.array_equal(gy.data(np.float32), ty.cpu().numpy())


@pytest.mark.parametrize("pad,lead", U8_LAYOUTS)
@pytest.mark.parametrize("variant", ["main_blend", "blend_frames", "reference_defaults"])
@pytest.mark.parametrize("w,h,scale", [(131, 67, 2), (96, 54, 4)])
def test_chain_stays_inside(vp, w, h, scale, variant, pad, lead):
    ocfg = RC.ChainConfig(dst_w=w * scale, dst_h=h * scale)
    if variant == "blend_frames":
        ocfg.temporal = RC.TEMPORAL_BLEND_FRAMES
    if variant == "reference_defaults":       # src/upscaler.cpp:666-758 as written (tests/test_gpu_cpp_api.py::ref_defaults)
        ocfg = RC.ChainConfig(
            dst_w=w * scale, dst_h=h * scale, temporal=RC.TEMPORAL_BLEND_FRAMES,
            pre=RC.BilateralConfig(RC.PRE_PROCESSING, True, True, 7, 30.0, 30.0, True, 30.0, 1.5, 2.0, True, 3),
            post=RC.BilateralConfig(RC.POST_PROCESSING, True, True, 5, 20.0, 20.0, True, 30.0, 1.2, 1.5, True, 2),
            sharpen=RC.SharpenConfig(0.8, 1.2, 0.4, 30.0, 1.5, 5, True, True, True),
            temporal_cfg=RC.TemporalConfig(use_flow=True, scene_detect=True))
    frames = synth.clip("scene", w, h, 5)
    dw, dh = ocfg.dst_w, ocfg.dst_h
    cfg = G.chain_cfg(vp, ocfg, w, h)
    tight = []
    with vp.Context(cfg) as c:
        for f in frames:
            src, dst = G.to_dev(f), G.dev_empty(dh, dw)
            c.call("vp_process_device", src.data_ptr(), G.pitch(src), dst.data_ptr(), G.pitch(dst), None)
            c.call("vp_synchronize")
            tight.append(dst.cpu().numpy())
    with vp.Context(cfg) as c:
        for i, f in enumerate(frames):
            gs = Guarded(h, w * 3, pad, lead)                 # padded / unaligned source rows as well
            gs.body()[:, :w * 3] = torch.from_numpy(f.reshape(h, w * 3)).cuda()
            g = Guarded(dh, dw * 3, pad, lead)
            c.call("vp_process_device", gs.ptr, gs.pitch, g.ptr, g.pitch, None)
            c.call("vp_synchronize")
            g.check(f"vp_process_device frame {i}")
            gs.check("source")
            assert np.array_equal(g.data().reshape(dh, dw, 3), tight[i]), i
