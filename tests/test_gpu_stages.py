"""Stage-wise parity of the CUDA kernels (through the C ABI) against the oracle.
Bars: bicubic / meanStdDev sums / addWeighted bit-exact; float stages +-1 LSB."""
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


SIZES = [(131, 67), (160, 90), (64, 32), (17, 9), (240, 135)]


@pytest.mark.parametrize("kind", ["noise", "scene"])
@pytest.mark.parametrize("scale", [2, 4])
@pytest.mark.parametrize("size", SIZES)
def test_bicubic_bit_exact(vp, ctx, kind, scale, size):
    import cv2
    w, h = size
    img = synth.frame(kind, w, h, 1)
    out = G.run_bicubic(vp, ctx, img, w * scale, h * scale)
    ref = R.resize_cubic(img, w * scale, h * scale)
    assert np.array_equal(out, ref)
    assert np.array_equal(out, cv2.resize(img, (w * scale, h * scale), interpolation=cv2.INTER_CUBIC))


def test_bicubic_padded_pitch(vp, ctx):
    img = synth.frame("noise", 96, 40, 2)
    src = G.padded_dev(img, 7)
    dst = torch.zeros((80, 192 * 3 + 5), dtype=torch.uint8, device="cuda")
    ctx.call("vp_bicubic", src.data_ptr(), G.pitch(src), 96, 40, dst.data_ptr(), G.pitch(dst), 192, 80, None)
    G.sync()
    out = dst.cpu().numpy()
    assert np.array_equal(out[:, :576].reshape(80, 192, 3), R.resize_cubic(img, 192, 80))
    assert not out[:, 576:].any()


@pytest.mark.parametrize("size", [(131, 67), (256, 64), (1000, 3)])
def test_mean_stddev_sums_exact(vp, ctx, size):
    w, h = size
    img = synth.frame("noise", w, h, 5)
    src = G.to_dev(img)
    sums = torch.zeros(6, dtype=torch.int64, device="cuda")
    ctx.call("vp_mean_stddev_sums", src.data_ptr(), G.pitch(src), w, h, sums.data_ptr(), None)
    G.sync()
    _, _, s, sq = R.mean_stddev(img)
    assert sums.cpu().tolist() == list(s) + list(sq)


@pytest.mark.parametrize("kind", ["noise", "scene"])
@pytest.mark.parametrize("d,sc,ss", [(3, 16, 16), (5, 24, 24), (7, 30, 30), (9, 36, 36), (11, 45, 45), (13, 45, 30), (15, 45, 45), (4, 10, 50)])
def test_bilateral_within_1lsb(vp, ctx, kind, d, sc, ss):
    img = synth.frame(kind, 150, 70, 3)
    out = G.run_bilateral(vp, ctx, img, d, sc, ss)
    ref = R.bilateral_u8c3(img, d, sc, ss)
    mx, n = G.diff_stats(out, ref)
    assert mx <= 1 and n <= img.size // 1000, (mx, n)


@pytest.mark.parametrize("kind,stage", [("noise", 0), ("scene", 1), ("flat", 0), ("flat", 1)])
def test_selective_bilateral_adaptive(vp, ctx, kind, stage):
    img = synth.frame(kind, 144, 80, 0)
    ocfg = RC.BilateralConfig(stage, True, True, 7, 30.0, 30.0, False, 30.0, 1.5, 2.0, False, 3)
    ref = RC.bilateral_plain(RC.NumpyOps(), img, ocfg)
    src = G.to_dev(img)
    dst = G.dev_empty(80, 144)
    cfg = G.bilateral_cfg(vp, ocfg)
    ctx.call("vp_selective_bilateral", C.byref(cfg), src.data_ptr(), G.pitch(src), dst.data_ptr(), G.pitch(dst), 144, 80, None)
    G.sync()
    mx, n = G.diff_stats(dst.cpu().numpy(), ref)
    assert mx <= 1 and n <= img.size // 1000, (mx, n)


@pytest.mark.parametrize("kind", ["noise", "scene"])
@pytest.mark.parametrize("size", [(150, 70), (64, 32), (70, 41)])
def test_edge_mask(vp, ctx, kind, size):
    w, h = size
    img = synth.frame(kind, w, h, 4)
    ocfg = RC.SharpenConfig(adaptive_sigma=False)
    ref = RC.edge_mask(RC.NumpyOps(), img, ocfg)
    src = G.to_dev(img)
    mask = torch.zeros((h, w), dtype=torch.float32, device="cuda")
    cfg = G.sharpen_cfg(vp, ocfg)
    ctx.call("vp_edge_mask", C.byref(cfg), src.data_ptr(), G.pitch(src), mask.data_ptr(), G.pitch(mask), w, h, None)
    G.sync()
    assert np.abs(mask.cpu().numpy() - ref).max() <= 1e-6
    import cv2
    refcv = RC.edge_mask(RC.Cv2Ops(), img, ocfg)
    assert np.abs(mask.cpu().numpy() - refcv).max() <= 1e-5


@pytest.mark.parametrize("kind", ["noise", "scene"])
@pytest.mark.parametrize("tone", [True, False])
@pytest.mark.parametrize("ks,sigma", [(5, 1.5), (3, 0.8), (7, 2.5)])
def test_unsharp_with_oracle_mask(vp, ctx, kind, tone, ks, sigma):
    w, h = 150, 70
    img = synth.frame(kind, w, h, 6)
    ocfg = RC.SharpenConfig(0.8, 1.2, 0.4, 30.0, sigma, ks, tone, True, False)
    m = RC.edge_mask(RC.NumpyOps(), img, ocfg)
    ref = RC.unsharp_mask(RC.NumpyOps(), img, m, ocfg)
    src, dm = G.to_dev(img), G.to_dev(m)
    dst = G.dev_empty(h, w)
    cfg = G.sharpen_cfg(vp, ocfg)
    ctx.call("vp_unsharp", C.byref(cfg), src.data_ptr(), G.pitch(src), dm.data_ptr(), G.pitch(dm), dst.data_ptr(), G.pitch(dst), w, h, None)
    G.sync()
    assert np.array_equal(dst.cpu().numpy(), ref)   # same mask in -> integer/explicitly rounded float path is exact


@pytest.mark.parametrize("kind", ["noise", "scene", "flat"])
def test_sharpen_within_1lsb(vp, ctx, kind):
    w, h = 200, 96
    img = synth.frame(kind, w, h, 7)
    ocfg = RC.SharpenConfig(adaptive_sigma=False)
    ref = RC.sharpen(RC.NumpyOps(), img, ocfg)
    refcv = RC.sharpen(RC.Cv2Ops(), img, ocfg)
    src = G.to_dev(img)
    dst = G.dev_empty(h, w)
    cfg = G.sharpen_cfg(vp, ocfg)
    ctx.call("vp_sharpen", C.byref(cfg), src.data_ptr(), G.pitch(src), dst.data_ptr(), G.pitch(dst), w, h, None)
    G.sync()
    out = dst.cpu().numpy()
    mx, n = G.diff_stats(out, ref)
    assert mx <= 1 and n <= out.size // 1000, (mx, n)
    mx, n = G.diff_stats(out, refcv)
    assert mx <= 1, (mx, n)


@pytest.mark.parametrize("alpha,beta", [(0.6, 0.4), (0.5, 0.5), (0.7, 0.3), (float(np.float32(0.6)), float(np.float32(1) - np.float32(0.6)))])
def test_add_weighted_exact(vp, ctx, alpha, beta):
    import cv2
    a = synth.frame("noise", 133, 31, 1)
    b = synth.frame("noise", 133, 31, 2)
    da, db = G.to_dev(a), G.to_dev(b)
    dst = G.dev_empty(31, 133)
    ctx.call("vp_add_weighted", da.data_ptr(), G.pitch(da), alpha, db.data_ptr(), G.pitch(db), beta, dst.data_ptr(), G.pitch(dst), 133, 31, None)
    G.sync()
    out = dst.cpu().numpy()
    assert np.array_equal(out, R.add_weighted_u8(a, alpha, b, beta))
    assert np.array_equal(out, cv2.addWeighted(a, alpha, b, beta, 0))


def test_add_weighted_all_pairs(vp, ctx):
    a = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    b = np.tile(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    a3 = np.stack([a, a, a], -1)[:, :, :][:, :255]   # width 255 -> 765 bytes, exercises the byte tail
    b3 = np.stack([b, b, b], -1)[:, :255]
    da, db = G.to_dev(a3), G.to_dev(b3)
    dst = G.dev_empty(256, 255)
    ctx.call("vp_add_weighted", da.data_ptr(), G.pitch(da), 0.6, db.data_ptr(), G.pitch(db), 0.4, dst.data_ptr(), G.pitch(dst), 255, 256, None)
    G.sync()
    assert np.array_equal(dst.cpu().numpy(), R.add_weighted_u8(a3, 0.6, b3, 0.4))


@pytest.mark.parametrize("nframes", [1, 2, 3, 4])
def test_temporal_blend(vp, ctx, nframes):
    w, h = 176, 50
    frames = [synth.frame("noise", w, h, t) for t in range(nframes)]
    ref = RC.blend_frames(frames, 0.6) if nframes > 1 else frames[0]
    devs = [G.to_dev(f) for f in frames]
    ptrs = (C.c_void_p * nframes)(*[d.data_ptr() for d in devs])
    dst = G.dev_empty(h, w)
    ctx.call("vp_temporal_blend", ptrs, nframes, G.pitch(devs[0]), 0.6, dst.data_ptr(), G.pitch(dst), w, h, None)
    G.sync()
    assert np.array_equal(dst.cpu().numpy(), ref)


def test_error_codes(vp, ctx):
    src = G.dev_empty(8, 8)
    assert vp.lib.vp_bicubic(ctx.handle, None, 24, 8, 8, src.data_ptr(), 48, 16, 16, None) == vp.VP_ERR_INVALID_ARG
    assert vp.lib.vp_bicubic(ctx.handle, src.data_ptr(), 24, 0, 8, src.data_ptr(), 48, 16, 16, None) == vp.VP_ERR_INVALID_ARG
    assert vp.lib.vp_bicubic(ctx.handle, src.data_ptr(), 23, 8, 8, src.data_ptr(), 48, 16, 16, None) == vp.VP_ERR_INVALID_ARG
    assert vp.lib.vp_bicubic(ctx.handle, src.data_ptr(), 24, 8, 8, src.data_ptr(), 12, 4, 4, None) == vp.VP_ERR_UNSUPPORTED
    assert vp.lib.vp_bilateral(ctx.handle, src.data_ptr(), 24, src.data_ptr(), 24, 8, 8, 17, 10.0, 10.0, None) == vp.VP_ERR_UNSUPPORTED
    assert b"radius" in vp.lib.vp_last_error(ctx.handle)
    assert vp.lib.vp_process_device(ctx.handle, src.data_ptr(), 24, src.data_ptr(), 24, None) == vp.VP_ERR_STATE


def test_stage_calls_on_two_streams_share_tables_safely(vp, ctx):
    """The stage-wise entry points share per-context tables (here vp_bilateral's candidate table, re-uploaded whenever the
    parameters change).  Calls alternating between two streams with alternating parameters, with no host synchronisation in
    between, must each see their own table: the context orders a call behind the previous call's kernels on the device."""
    w, h = 512, 288
    img = synth.frame("scene", w, h, 3)
    want = {p: G.run_bilateral(vp, ctx, img, *p) for p in [(5, 30.0, 30.0), (9, 12.0, 20.0)]}
    src = G.to_dev(img)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for i in range(8):
        p = (5, 30.0, 30.0) if i % 2 == 0 else (9, 12.0, 20.0)
        st = s1 if i % 2 == 0 else s2
        dst = G.dev_empty(h, w)
        ctx.call("vp_bilateral", src.data_ptr(), G.pitch(src), dst.data_ptr(), G.pitch(dst), w, h, p[0], p[1], p[2], st.cuda_stream)
        outs.append((p, dst))
    G.sync()
    for p, dst in outs:
        assert np.array_equal(dst.cpu().numpy(), want[p]), p
