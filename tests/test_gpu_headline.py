"""Parity of exactly what bench.py times, at BASELINE's full size (C4: 1920x1080 -> 3840x2160, plain
SelectiveBilateral PRE/POST, AdaptiveSharpening, TemporalConsistency::blendFrames over the last three frames):

  * vp_process_batch_device with four stream lanes, the first three frames fed as temporal warm-up (d_dst = NULL, what
    rank k > 0 of the file-mode sharding does), enough frames for the steady four-frame blend and the state ring's wrap;
  * vp_submit_host / vp_wait with three pinned frames in flight (the `e2e` leg);

both against the reference's own OpenCV CPU path (oracle/ref_chain.py over the real cv2 kernels;
reference src/upscaler.cpp:772-829, src/temporal_consistency.cpp:127-148), plus the two bilateral filters of that
chain stage-wise against cv2.bilateralFilter at the parameters calculateAdaptiveParams picks on the `scene` clip.
Bars: chain PSNR >= 50 dB and max |d| <= 4; bilateral max |d| <= 1.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import ref_chain as RC
from oracle import synth

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from tests import gpu_util as G  # noqa: E402

SW, SH, DW, DH = 1920, 1080, 3840, 2160
NFRAMES, WARM = 9, 3


def _ocfg():
    return RC.ChainConfig(dst_w=DW, dst_h=DH, temporal=RC.TEMPORAL_BLEND_FRAMES)


@pytest.fixture(scope="module")
def clip():
    return synth.clip("scene", SW, SH, NFRAMES, 20260)


@pytest.fixture(scope="module")
def cv2_refs(clip):
    """The reference chain over the real OpenCV kernels on all NFRAMES frames (about 1 s per frame on the box's cores)."""
    oracle = RC.Chain(RC.Cv2Ops(), _ocfg())
    return [oracle.process(f) for f in clip]


def _report(tag, outs, refs):
    worst_psnr, worst_max = 1e9, 0
    hist = np.zeros(6, np.int64)
    for o, r in zip(outs, refs):
        d = np.abs(o.astype(np.int16) - r.astype(np.int16))
        hist += np.bincount(np.minimum(d, 5).ravel(), minlength=6)
        worst_psnr = min(worst_psnr, RC.psnr(o, r))
        worst_max = max(worst_max, int(d.max()))
    tot = hist.sum()
    print(f"\n{tag}: worst PSNR {worst_psnr:.2f} dB, max |d| {worst_max}, |d| histogram 0..5+ = "
          + ", ".join(f"{100.0 * h / tot:.4f}%" for h in hist))
    return worst_psnr, worst_max


def test_c4_full_size_batch_api_vs_cv2(vp, clip, cv2_refs):
    cfg = G.chain_cfg(vp, _ocfg(), SW, SH)
    with vp.Context(cfg) as ctx:
        ctx.call("vp_set_batch_lanes", 4)
        srcs = [G.to_dev(f) for f in clip]
        dsts = [None] * WARM + [G.dev_empty(DH, DW) for _ in range(NFRAMES - WARM)]
        sp = (C.c_void_p * NFRAMES)(*[t.data_ptr() for t in srcs])
        dp = (C.c_void_p * NFRAMES)(*[d.data_ptr() if d is not None else None for d in dsts])
        ctx.call("vp_process_batch_device", NFRAMES, sp, SW * 3, dp, DW * 3, None)
        ctx.call("vp_synchronize")
        assert ctx.stage_params(0)[0] == 11 and ctx.stage_params(1)[0] == 5
        outs = [d.cpu().numpy() for d in dsts[WARM:]]
    psnr, mx = _report("C4 batch API (4 lanes, 3 warm-up frames)", outs, cv2_refs[WARM:])
    assert psnr >= 50.0 and mx <= 4


def test_c4_full_size_submit_host_vs_cv2(vp, clip, cv2_refs):
    cfg = G.chain_cfg(vp, _ocfg(), SW, SH)
    with vp.Context(cfg) as ctx:
        depth = vp.lib.vp_pipeline_depth(ctx.handle)
        assert depth == 3
        pin_in = [torch.from_numpy(f).pin_memory() for f in clip]
        pin_out = [torch.zeros((DH, DW, 3), dtype=torch.uint8).pin_memory() for _ in clip]
        inflight = 0
        for i in range(NFRAMES):
            if inflight == depth:
                ctx.call("vp_wait")
                inflight -= 1
            ctx.call("vp_submit_host", pin_in[i].data_ptr(), SW * 3, pin_out[i].data_ptr(), DW * 3)
            inflight += 1
        while inflight:
            ctx.call("vp_wait")
            inflight -= 1
        outs = [o.numpy() for o in pin_out]
    psnr, mx = _report("C4 vp_submit_host/vp_wait (depth 3)", outs, cv2_refs)
    assert psnr >= 50.0 and mx <= 4


def test_c4_batch_and_host_paths_agree_at_full_size(vp, clip):
    """The two timed APIs and the per-frame call produce the same bytes at full size (the small-size identities of
    test_gpu_chain.py, repeated where bench.py runs)."""
    cfg = G.chain_cfg(vp, _ocfg(), SW, SH)
    n = 6
    with vp.Context(cfg) as ctx:
        srcs = [G.to_dev(f) for f in clip[:n]]
        per = []
        for s in srcs:
            d = G.dev_empty(DH, DW)
            ctx.call("vp_process_device", s.data_ptr(), SW * 3, d.data_ptr(), DW * 3, None)
            per.append(d)
        ctx.call("vp_synchronize")
        ctx.call("vp_reset_temporal")
        bat = [G.dev_empty(DH, DW) for _ in range(n)]
        sp = (C.c_void_p * n)(*[t.data_ptr() for t in srcs])
        dp = (C.c_void_p * n)(*[d.data_ptr() for d in bat])
        ctx.call("vp_process_batch_device", n, sp, SW * 3, dp, DW * 3, None)
        ctx.call("vp_synchronize")
        for a, b in zip(per, bat):
            assert torch.equal(a, b)
        ctx.call("vp_reset_temporal")
        pin_in = [torch.from_numpy(f).pin_memory() for f in clip[:n]]
        pin_out = [torch.zeros((DH, DW, 3), dtype=torch.uint8).pin_memory() for _ in range(n)]
        inflight = 0
        for i in range(n):
            if inflight == 3:
                ctx.call("vp_wait")
                inflight -= 1
            ctx.call("vp_submit_host", pin_in[i].data_ptr(), SW * 3, pin_out[i].data_ptr(), DW * 3)
            inflight += 1
        while inflight:
            ctx.call("vp_wait")
            inflight -= 1
        for a, b in zip(per, pin_out):
            assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("name,w,h,d,sigma", [("PRE d=11 sigma 45 @1080p", 1920, 1080, 11, 45.0),
                                              ("POST d=5 sigma 24 @4K", 3840, 2160, 5, 24.0)])
def test_bilateral_full_size_vs_cv2(vp, name, w, h, d, sigma):
    """The chain's two bilateral launches stage-wise, at the sizes and parameters of the headline run
    (calculateAdaptiveParams: noise_factor 1.5 on the `scene` clip), against cv2.bilateralFilter."""
    import cv2
    if w == 1920:
        img = synth.frame("scene", w, h, 3, 20260)
    else:   # a realistic POST input: the upscaled frame
        img = cv2.resize(synth.frame("scene", w // 2, h // 2, 3, 20260), (w, h), interpolation=cv2.INTER_CUBIC)
    ref = cv2.bilateralFilter(img, d, sigma, sigma)
    with vp.Context() as ctx:
        out = G.run_bilateral(vp, ctx, img, d, sigma, sigma)
    diff = np.abs(out.astype(np.int16) - ref.astype(np.int16))
    print(f"\n{name}: max |d| {int(diff.max())}, values off by one {int((diff == 1).sum())} of {diff.size}")
    assert diff.max() <= 1
    assert (diff > 0).sum() <= diff.size // 1000


@pytest.mark.parametrize("toggles", [(1, 0, 0), (0, 0, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1)])
def test_same_size_chain_without_temporal_writes_dst(vp, toggles):
    """Every stage combination at x1 without a temporal stage writes the caller's frame - including PRE alone, whose
    result lives in a lane buffer (the PRE-only chain once returned VP_OK without touching d_dst)."""
    sw, sh = 96, 64
    use_pre, use_sharpen, use_post = toggles
    ocfg = RC.ChainConfig(dst_w=sw, dst_h=sh, use_pre=bool(use_pre), use_sharpen=bool(use_sharpen), use_post=bool(use_post),
                          temporal=RC.TEMPORAL_NONE)
    frames = synth.clip("scene", sw, sh, 2)
    oracle = RC.Chain(RC.NumpyOps(), ocfg)
    cfg = G.chain_cfg(vp, ocfg, sw, sh)
    with vp.Context(cfg) as ctx:
        for f in frames:
            ref = oracle.process(f)
            src = G.to_dev(f)
            dst = torch.full((sh, sw, 3), 0xA5, dtype=torch.uint8, device="cuda")
            ctx.call("vp_process_device", src.data_ptr(), sw * 3, dst.data_ptr(), sw * 3, None)
            ctx.call("vp_synchronize")
            out = dst.cpu().numpy()
            mx, nd = G.diff_stats(out, ref)
            assert RC.psnr(out, ref) >= 50.0 and mx <= 4, (toggles, mx, nd)
            if toggles == (0, 0, 0):
                assert mx == 0
            # and through the synchronous host call (its D2H reads d_dst)
            hout = np.full((sh, sw, 3), 0x5A, np.uint8)
            ctx.call("vp_process_host", f.ctypes.data, f.strides[0], hout.ctypes.data, hout.strides[0])
            assert np.array_equal(hout, out)
