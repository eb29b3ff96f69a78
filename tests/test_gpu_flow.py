"""TemporalConsistency's motion-compensated front end through the C ABI against the oracle:
Farneback flow within 2e-3 px of cv2.calcOpticalFlowFarneback (float / double summation order), remap bit-exact,
reliability mask within 1e-5, the whole TemporalConsistency::process within +-1 LSB on a small fraction."""
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


def gpu_flow(vp, ctx, f0, f1, tcfg):
    h, w = f0.shape[:2]
    a, b = G.to_dev(f0), G.to_dev(f1)
    fx = torch.empty((h, w), dtype=torch.float32, device="cuda")
    fy = torch.empty((h, w), dtype=torch.float32, device="cuda")
    ctx.call("vp_optical_flow", G.temporal_cfg(vp, tcfg), a.data_ptr(), b.data_ptr(), G.pitch(a), fx.data_ptr(), fy.data_ptr(), G.pitch(fx), w, h, None)
    G.sync()
    return np.stack([fx.cpu().numpy(), fy.cpu().numpy()], axis=-1)


def shifted_pair(w, h, dx, dy):
    """a textured frame and the same texture moved by (dx, dy): a flow field with known sign and size"""
    big = synth.frame("scene", w + 32, h + 32, 5)
    import cv2
    big = cv2.GaussianBlur(big, (7, 7), 2.0)
    return np.ascontiguousarray(big[16:16 + h, 16:16 + w]), np.ascontiguousarray(big[16 - dy:16 - dy + h, 16 - dx:16 - dx + w])


@pytest.mark.parametrize("size", [(160, 96), (131, 67), (320, 180), (64, 48)])
@pytest.mark.parametrize("levels,iters,win", [(3, 3, 15), (0, 1, 15), (2, 2, 9)])
def test_farneback_matches_cv2(vp, ctx, size, levels, iters, win):
    import cv2
    w, h = size
    f0, f1 = synth.frame("scene", w, h, 3), synth.frame("scene", w, h, 4)
    t = RC.TemporalConfig(levels=levels, iterations=iters, winsize=win)
    flow = gpu_flow(vp, ctx, f0, f1, t)
    g0, g1 = cv2.cvtColor(f0, cv2.COLOR_BGR2GRAY), cv2.cvtColor(f1, cv2.COLOR_BGR2GRAY)
    ref = cv2.calcOpticalFlowFarneback(g0, g1, None, t.pyr_scale, levels, win, iters, t.poly_n, t.poly_sigma, 0)
    assert np.abs(flow - ref).max() <= 2e-3, np.abs(flow - ref).max()
    mine = R.farneback(R.bgr2gray(f0), R.bgr2gray(f1), t.pyr_scale, levels, win, iters, t.poly_n, t.poly_sigma)
    assert np.abs(flow - mine).max() <= 2e-3


def test_farneback_recovers_a_translation(vp, ctx):
    f0, f1 = shifted_pair(256, 144, 3, -2)
    flow = gpu_flow(vp, ctx, f0, f1, RC.TemporalConfig())
    core = flow[32:-32, 32:-32]
    assert abs(np.median(core[..., 0]) - 3.0) < 0.5 and abs(np.median(core[..., 1]) + 2.0) < 0.5


@pytest.mark.parametrize("size", [(160, 96), (131, 67), (17, 9)])
def test_warp_bit_exact(vp, ctx, size):
    import cv2
    w, h = size
    img = synth.frame("noise", w, h, 1)
    rng = np.random.default_rng(0)
    flow = (rng.standard_normal((h, w, 2)) * 6).astype(np.float32)
    flow[0:3] *= 40            # far outside the frame: BORDER_REPLICATE
    flow[5] = np.rint(flow[5])  # integer displacements
    src = G.to_dev(img)
    fx, fy = G.to_dev(flow[..., 0].copy()), G.to_dev(flow[..., 1].copy())
    dst = G.dev_empty(h, w)
    ctx.call("vp_warp_flow", src.data_ptr(), G.pitch(src), fx.data_ptr(), fy.data_ptr(), G.pitch(fx), dst.data_ptr(), G.pitch(dst), w, h, None)
    G.sync()
    mx, my = R.flow_to_maps(flow)
    assert np.array_equal(dst.cpu().numpy(), cv2.remap(img, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE))
    assert np.array_equal(dst.cpu().numpy(), R.remap_linear_replicate_u8(img, mx, my))


@pytest.mark.parametrize("size", [(160, 96), (131, 67), (40, 20)])
def test_reliability_mask(vp, ctx, size):
    w, h = size
    rng = np.random.default_rng(1)
    flow = (rng.standard_normal((h, w, 2)) * 14).astype(np.float32)      # magnitudes on both sides of the threshold
    fx, fy = G.to_dev(flow[..., 0].copy()), G.to_dev(flow[..., 1].copy())
    mask = torch.empty((h, w), dtype=torch.float32, device="cuda")
    ctx.call("vp_flow_reliability", fx.data_ptr(), fy.data_ptr(), G.pitch(fx), 15.0, mask.data_ptr(), G.pitch(mask), w, h, None)
    G.sync()
    for ops in (RC.NumpyOps(), RC.Cv2Ops()):
        ref = RC.flow_reliability_mask(ops, flow, 15.0)
        assert np.abs(mask.cpu().numpy() - ref).max() <= 1e-5


def run_tc(vp, frames, tcfg, sw, sh):
    ocfg = RC.ChainConfig(dst_w=sw, dst_h=sh, use_pre=False, use_sharpen=False, use_post=False, temporal=RC.TEMPORAL_BLEND_FRAMES, temporal_cfg=tcfg)
    outs = []
    with vp.Context(G.chain_cfg(vp, ocfg, sw, sh)) as ctx:
        for f in frames:
            src = G.to_dev(f)
            dst = G.dev_empty(sh, sw)
            ctx.call("vp_process_device", src.data_ptr(), sw * 3, dst.data_ptr(), G.pitch(dst), None)
            G.sync()
            outs.append(dst.cpu().numpy())
    return outs


@pytest.mark.parametrize("kind", ["scene", "noise"])
@pytest.mark.parametrize("buffer_size", [3, 2])
def test_temporal_consistency_full(vp, kind, buffer_size):
    """TemporalConsistency::process as written: flow, warp, masks, blendFrames over every buffered frame"""
    sw, sh = 160, 96
    frames = synth.clip(kind, sw, sh, 7)
    t = RC.TemporalConfig(buffer_size=buffer_size, use_flow=True, scene_detect=True)
    outs = run_tc(vp, frames, t, sw, sh)
    for ops in (RC.NumpyOps(), RC.Cv2Ops()):
        oracle = RC.TemporalConsistencyFull(t, ops)
        for i, (o, f) in enumerate(zip(outs, frames)):
            ref = oracle.process(f)
            mx, n = G.diff_stats(o, ref)
            assert mx <= 1 and n <= o.size * 0.01, (ops.name, i, mx, n)


def test_temporal_consistency_full_scene_cut(vp):
    sw, sh = 128, 72
    frames = synth.clip("flat", sw, sh, 4) + synth.clip("scene", sw, sh, 4, t0=3)      # a flat -> scene cut at frame 4
    t = RC.TemporalConfig(use_flow=True, scene_detect=True)
    outs = run_tc(vp, frames, t, sw, sh)
    oracle = RC.TemporalConsistencyFull(t, RC.NumpyOps())
    refs = [oracle.process(f) for f in frames]
    assert np.array_equal(refs[4], frames[4])          # the oracle sees the cut
    for i, (o, r) in enumerate(zip(outs, refs)):
        mx, n = G.diff_stats(o, r)
        assert mx <= 1 and n <= o.size * 0.01, (i, mx, n)
    assert np.array_equal(outs[4], frames[4])


def test_chain_with_flow_temporal(vp):
    """the whole chain in front of the flow-based temporal stage, batch API included"""
    import ctypes as C
    sw, sh = 128, 72
    t = RC.TemporalConfig(use_flow=True, scene_detect=True)
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_BLEND_FRAMES, temporal_cfg=t)
    frames = synth.clip("scene", sw, sh, 5)
    oracle = RC.Chain(RC.NumpyOps(), ocfg)
    refs = [oracle.process(f) for f in frames]
    for batch in (False, True):
        with vp.Context(G.chain_cfg(vp, ocfg, sw, sh)) as ctx:
            srcs = [G.to_dev(f) for f in frames]
            dsts = [G.dev_empty(2 * sh, 2 * sw) for _ in frames]
            if batch:
                sp = (C.c_void_p * len(frames))(*[s.data_ptr() for s in srcs])
                dp = (C.c_void_p * len(frames))(*[d.data_ptr() for d in dsts])
                ctx.call("vp_process_batch_device", len(frames), sp, sw * 3, dp, G.pitch(dsts[0]), None)
            else:
                for s, d in zip(srcs, dsts):
                    ctx.call("vp_process_device", s.data_ptr(), sw * 3, d.data_ptr(), G.pitch(d), None)
            ctx.call("vp_synchronize")
            for d, r in zip(dsts, refs):
                o = d.cpu().numpy()
                mx, _ = G.diff_stats(o, r)
                assert RC.psnr(o, r) >= 50.0 and mx <= 4, (batch, RC.psnr(o, r), mx)


@pytest.mark.parametrize("size", [(20, 12), (131, 67), (33, 200)])
def test_temporal_consistency_full_odd_sizes(vp, size):
    """ragged and tiny frames: every pyramid level smaller than 32 px is dropped, windows larger than the frame clamp"""
    sw, sh = size
    frames = synth.clip("scene", sw, sh, 5)
    t = RC.TemporalConfig(use_flow=True, scene_detect=True)
    outs = run_tc(vp, frames, t, sw, sh)
    oracle = RC.TemporalConsistencyFull(t, RC.Cv2Ops())
    for i, (o, f) in enumerate(zip(outs, frames)):
        ref = oracle.process(f)
        mx, n = G.diff_stats(o, ref)
        assert mx <= 1 and n <= max(o.size * 0.01, 8), (i, mx, n)


def test_flow_temporal_through_the_host_api(vp):
    """vp_process_host and the pipelined vp_submit_host / vp_wait drive the same device state as vp_process_device"""
    sw, sh = 96, 64
    t = RC.TemporalConfig(use_flow=True, scene_detect=True)
    ocfg = RC.ChainConfig(dst_w=sw, dst_h=sh, use_pre=False, use_sharpen=False, use_post=False, temporal=RC.TEMPORAL_BLEND_FRAMES, temporal_cfg=t)
    frames = synth.clip("scene", sw, sh, 6)
    ref = run_tc(vp, frames, t, sw, sh)
    with vp.Context(G.chain_cfg(vp, ocfg, sw, sh)) as ctx:
        outs = [np.empty((sh, sw, 3), np.uint8) for _ in frames]
        for f, o in zip(frames, outs):
            ctx.call("vp_process_host", f.ctypes.data, sw * 3, o.ctypes.data, sw * 3)
        for o, r in zip(outs, ref):
            assert np.array_equal(o, r)
    with vp.Context(G.chain_cfg(vp, ocfg, sw, sh)) as ctx:
        outs = [np.empty((sh, sw, 3), np.uint8) for _ in frames]
        depth = vp.lib.vp_pipeline_depth(ctx.handle)
        inflight = 0
        for f, o in zip(frames, outs):
            if inflight == depth:
                ctx.call("vp_wait"); inflight -= 1
            ctx.call("vp_submit_host", f.ctypes.data, sw * 3, o.ctypes.data, sw * 3)
            inflight += 1
        while inflight:
            ctx.call("vp_wait"); inflight -= 1
        for o, r in zip(outs, ref):
            assert np.array_equal(o, r)


def test_flow_temporal_reset(vp):
    sw, sh = 96, 64
    t = RC.TemporalConfig(use_flow=True, scene_detect=True)
    ocfg = RC.ChainConfig(dst_w=sw, dst_h=sh, use_pre=False, use_sharpen=False, use_post=False, temporal=RC.TEMPORAL_BLEND_FRAMES, temporal_cfg=t)
    frames = synth.clip("scene", sw, sh, 4)
    ref = run_tc(vp, frames, t, sw, sh)
    with vp.Context(G.chain_cfg(vp, ocfg, sw, sh)) as ctx:
        for rep in range(2):
            for f, r in zip(frames, ref):
                src = G.to_dev(f)
                dst = G.dev_empty(sh, sw)
                ctx.call("vp_process_device", src.data_ptr(), sw * 3, dst.data_ptr(), G.pitch(dst), None)
                G.sync()
                assert np.array_equal(dst.cpu().numpy(), r), rep
            ctx.call("vp_reset_temporal")


def test_against_frozen_opencv_vectors(vp, ctx):
    """tests/golden/golden_widened.npz = output of the real OpenCV kernels (minted by make_golden_widened.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_widened.npz"))
    img, nxt = synth.frame("scene", 96, 64, 5), synth.frame("scene", 96, 64, 6)
    t = RC.TemporalConfig(use_flow=True, scene_detect=True)
    assert np.abs(gpu_flow(vp, ctx, img, nxt, t) - g["farneback_flow"]).max() <= 2e-3
    src = G.to_dev(img)
    fx, fy = G.to_dev(g["farneback_flow"][..., 0].copy()), G.to_dev(g["farneback_flow"][..., 1].copy())
    dst = G.dev_empty(64, 96)
    ctx.call("vp_warp_flow", src.data_ptr(), G.pitch(src), fx.data_ptr(), fy.data_ptr(), G.pitch(fx), dst.data_ptr(), G.pitch(dst), 96, 64, None)
    G.sync()
    assert np.array_equal(dst.cpu().numpy(), g["warp"])
    outs = run_tc(vp, synth.clip("scene", 96, 64, 5), t, 96, 64)
    for k, o in enumerate(outs):
        mx, n = G.diff_stats(o, g[f"tc_full_t{k}"])
        assert mx <= 1 and n <= o.size * 0.01, (k, mx, n)
    # pyramids and the detail / sigma planes of the other widened rows
    down = G.dev_empty(32, 48)
    ctx.call("vp_pyr_down", src.data_ptr(), G.pitch(src), 96, 64, down.data_ptr(), G.pitch(down), None)
    up = G.dev_empty(64, 96)
    ctx.call("vp_pyr_up", down.data_ptr(), G.pitch(down), 48, 32, up.data_ptr(), G.pitch(up), 96, 64, None)
    mask = torch.empty((64, 96), dtype=torch.float32, device="cuda")
    ctx.call("vp_detail_mask", 30.0, src.data_ptr(), G.pitch(src), mask.data_ptr(), G.pitch(mask), 96, 64, None)
    sig = torch.empty((64, 96), dtype=torch.float32, device="cuda")
    ctx.call("vp_sigma_map", src.data_ptr(), G.pitch(src), sig.data_ptr(), G.pitch(sig), 96, 64, None)
    G.sync()
    assert np.array_equal(down.cpu().numpy(), g["pyr_down"]) and np.array_equal(up.cpu().numpy(), g["pyr_up"])
    assert np.abs(mask.cpu().numpy() - g["detail_mask"]).max() <= 5e-6
    assert np.abs(sig.cpu().numpy() - g["sigma_map"]).max() <= 5e-6


def test_full_size_1080p_against_opencv(vp, ctx):
    """BASELINE's frame size: Farneback, the temporal stage and the repaired branches against the real OpenCV kernels"""
    import cv2
    w, h = 1920, 1080
    f0, f1, f2 = synth.clip("scene", w, h, 3)
    t = RC.TemporalConfig(use_flow=True, scene_detect=True)
    flow = gpu_flow(vp, ctx, f0, f1, t)
    g0, g1 = cv2.cvtColor(f0, cv2.COLOR_BGR2GRAY), cv2.cvtColor(f1, cv2.COLOR_BGR2GRAY)
    ref = cv2.calcOpticalFlowFarneback(g0, g1, None, t.pyr_scale, t.levels, t.winsize, t.iterations, t.poly_n, t.poly_sigma, 0)
    assert np.abs(flow - ref).max() <= 5e-3, np.abs(flow - ref).max()
    outs = run_tc(vp, [f0, f1, f2], t, w, h)
    oracle = RC.TemporalConsistencyFull(t, RC.Cv2Ops())
    for i, (o, f) in enumerate(zip(outs, (f0, f1, f2))):
        r = oracle.process(f)
        mx, n = G.diff_stats(o, r)
        assert mx <= 1 and n <= o.size * 0.01, (i, mx, n)
    # SelectiveBilateral multi-scale and AdaptiveSharpening adaptive_sigma at the same size
    ops = RC.Cv2Ops()
    src = G.to_dev(f0)
    dst = G.dev_empty(h, w)
    bcfg = RC.BilateralConfig(use_multiscale=True, num_scales=3)
    ctx.call("vp_selective_bilateral", G.bilateral_cfg(vp, bcfg), src.data_ptr(), G.pitch(src), dst.data_ptr(), G.pitch(dst), w, h, None)
    G.sync()
    r = RC.selective_bilateral_process(ops, f0, bcfg)
    mx, _ = G.diff_stats(dst.cpu().numpy(), r)
    assert mx <= 4 and RC.psnr(dst.cpu().numpy(), r) >= 50.0
    scfg = RC.SharpenConfig(adaptive_sigma=True)
    ctx.call("vp_sharpen", G.sharpen_cfg(vp, scfg), src.data_ptr(), G.pitch(src), dst.data_ptr(), G.pitch(dst), w, h, None)
    G.sync()
    r = RC.sharpen_variable_sigma(ops, f0, scfg)
    mx, n = G.diff_stats(dst.cpu().numpy(), r)
    assert mx <= 2 and n <= r.size * 0.02, (mx, n)
