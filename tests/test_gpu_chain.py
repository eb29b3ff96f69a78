"""End-to-end parity of the fused chain (vp_process_device / vp_process_host / submit-wait) against the
oracle chain, incl. the temporal stage, adaptive-parameter selection and multi-GPU segment equivalence."""
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


def run_chain_device(vp, ocfg, frames, sw, sh, outputs=None):
    cfg = G.chain_cfg(vp, ocfg, sw, sh)
    outs = []
    with vp.Context(cfg) as ctx:
        for i, f in enumerate(frames):
            src = G.to_dev(f)
            want = outputs is None or outputs[i]
            dst = G.dev_empty(ocfg.dst_h, ocfg.dst_w) if want else None
            ctx.call("vp_process_device", src.data_ptr(), G.pitch(src), dst.data_ptr() if want else None,
                     G.pitch(dst) if want else 0, None)
            ctx.call("vp_synchronize")
            outs.append(dst.cpu().numpy() if want else None)
        params = (ctx.stage_params(0), ctx.stage_params(1))
    return outs, params


@pytest.mark.parametrize("kind", ["noise", "scene", "flat"])
@pytest.mark.parametrize("temporal", [RC.TEMPORAL_NONE, RC.TEMPORAL_MAIN_BLEND, RC.TEMPORAL_BLEND_FRAMES])
def test_chain_vs_oracle(vp, kind, temporal):
    sw, sh, n = 160, 96, 4
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=temporal)
    frames = synth.clip(kind, sw, sh, n)
    oracle = RC.Chain(RC.NumpyOps(), ocfg)
    refs = [oracle.process(f) for f in frames]
    outs, params = run_chain_device(vp, ocfg, frames, sw, sh)
    # adaptive parameter selection agrees with calculateAdaptiveParams
    assert params[0][0] == RC.adaptive_params(RC.NumpyOps(), frames[-1], ocfg.pre)[0]
    for o, r in zip(outs, refs):
        mx, nd = G.diff_stats(o, r)
        assert RC.psnr(o, r) >= 50.0, (mx, nd)
        assert mx <= 4, (mx, nd)


@pytest.mark.parametrize("buffer_size", [1, 2, 3])
def test_chain_temporal_buffer_sizes(vp, buffer_size):
    sw, sh, n = 96, 64, 5
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_BLEND_FRAMES)
    ocfg.temporal_cfg = RC.TemporalConfig(buffer_size=buffer_size, blend_strength=0.45)
    frames = synth.clip("scene", sw, sh, n)
    oracle = RC.Chain(RC.NumpyOps(), ocfg)
    refs = [oracle.process(f) for f in frames]
    outs, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    for o, r in zip(outs, refs):
        mx, nd = G.diff_stats(o, r)
        assert RC.psnr(o, r) >= 50.0 and mx <= 4, (buffer_size, mx, nd)


@pytest.mark.parametrize("dst,ks,sigma,adaptive", [((384, 256), 3, 0.8, True), ((384, 256), 7, 2.5, False), ((96, 64), 5, 1.5, False),
                                                   ((144, 100), 5, 1.5, True), ((288, 192), 5, 1.0, True)])
def test_chain_other_ratios_and_configs(vp, dst, ks, sigma, adaptive):
    """x4, x1 (no resize), x1.5-ish and x3 chains; 3/7-tap unsharp Gaussians; fixed (non-adaptive) bilateral parameters."""
    sw, sh = 96, 64
    ocfg = RC.ChainConfig(dst_w=dst[0], dst_h=dst[1], temporal=RC.TEMPORAL_MAIN_BLEND)
    ocfg.sharpen = RC.SharpenConfig(0.7, 1.3, 0.3, 25.0, sigma, ks, True, True, False)
    ocfg.pre = RC.BilateralConfig(RC.PRE_PROCESSING, True, adaptive, 5, 25.0, 12.0, False, 30.0, 1.5, 2.0, False, 3)
    ocfg.post = RC.BilateralConfig(RC.POST_PROCESSING, True, adaptive, 7, 18.0, 40.0, False, 30.0, 1.2, 1.5, False, 2)
    frames = synth.clip("scene", sw, sh, 3)
    oracle = RC.Chain(RC.NumpyOps(), ocfg)
    refs = [oracle.process(f) for f in frames]
    outs, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    for o, r in zip(outs, refs):
        mx, nd = G.diff_stats(o, r)
        assert RC.psnr(o, r) >= 50.0 and mx <= 4, (dst, ks, mx, nd)


@pytest.mark.parametrize("use_post", [True, False])
def test_scene_change_detection(vp, use_post):
    """detectSceneChange on the device: a flat -> scene cut clears the history (the cut frame passes through
    unblended) and blending resumes afterwards; frames without a cut blend as usual."""
    sw, sh = 96, 64
    frames = [synth.frame("flat", sw, sh, 0), synth.frame("flat", sw, sh, 1), synth.frame("flat", sw, sh, 2),
              synth.frame("scene", sw, sh, 3), synth.frame("scene", sw, sh, 4), synth.frame("scene", sw, sh, 5)]
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_BLEND_FRAMES, use_post=use_post)
    ocfg.temporal_cfg = RC.TemporalConfig(scene_detect=True)
    oracle = RC.Chain(RC.NumpyOps(), ocfg)
    refs = [oracle.process(f) for f in frames]
    assert len(oracle.temporal.frames) == 3
    spatial3 = RC.Chain(RC.NumpyOps(), RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_NONE, use_post=use_post)).process(frames[3])
    assert np.array_equal(refs[3], spatial3)          # the oracle itself saw the cut at frame 3
    outs, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    for i, (o, r) in enumerate(zip(outs, refs)):
        mx, nd = G.diff_stats(o, r)
        assert RC.psnr(o, r) >= 50.0 and mx <= 4, (i, mx, nd)
    # and with the detector off the same clip blends across the cut
    ocfg.temporal_cfg = RC.TemporalConfig(scene_detect=False)
    outs2, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    assert RC.psnr(outs2[3], refs[3]) < 45.0


def test_chain_stage_toggles(vp):
    sw, sh = 96, 64
    frames = synth.clip("scene", sw, sh, 3)
    for use_pre, use_sharpen, use_post in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1)]:
        ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, use_pre=bool(use_pre), use_sharpen=bool(use_sharpen),
                              use_post=bool(use_post), temporal=RC.TEMPORAL_MAIN_BLEND)
        oracle = RC.Chain(RC.NumpyOps(), ocfg)
        refs = [oracle.process(f) for f in frames]
        outs, _ = run_chain_device(vp, ocfg, frames, sw, sh)
        for o, r in zip(outs, refs):
            mx, nd = G.diff_stats(o, r)
            assert RC.psnr(o, r) >= 50.0 and mx <= 4, (use_pre, use_sharpen, use_post, mx, nd)
            if not use_pre and not use_post and not use_sharpen:
                assert mx == 0   # bicubic + addWeighted only: bit-exact


@pytest.mark.parametrize("temporal,warm", [(RC.TEMPORAL_MAIN_BLEND, 1), (RC.TEMPORAL_BLEND_FRAMES, 3)])
def test_segment_sharding_is_bit_identical(vp, temporal, warm):
    """file/--fast mode partitioning (DESIGN.md multi-GPU): a segment started with `warm` warm-up frames
    (d_dst = NULL) reproduces the single-context outputs exactly."""
    sw, sh, n = 128, 72, 7
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=temporal)
    frames = synth.clip("scene", sw, sh, n)
    full, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    start = 4
    seg_frames = frames[start - warm:]
    flags = [False] * warm + [True] * (n - start)
    seg, _ = run_chain_device(vp, ocfg, seg_frames, sw, sh, outputs=flags)
    for i in range(start, n):
        assert np.array_equal(full[i], seg[warm + i - start])


@pytest.mark.parametrize("temporal", [RC.TEMPORAL_NONE, RC.TEMPORAL_MAIN_BLEND, RC.TEMPORAL_BLEND_FRAMES])
def test_batch_api_is_bit_identical_to_per_frame(vp, temporal):
    """vp_process_batch_device (two stream lanes, events for the temporal dependency) == one
    vp_process_device call per frame, including warm-up frames (d_dst = NULL) and repeated batches."""
    sw, sh, n = 128, 72, 9
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=temporal)
    frames = synth.clip("scene", sw, sh, n)
    flags = [False, False] + [True] * (n - 2)
    ref, _ = run_chain_device(vp, ocfg, frames, sw, sh, outputs=flags)
    cfg = G.chain_cfg(vp, ocfg, sw, sh)
    with vp.Context(cfg) as ctx:
        srcs = [G.to_dev(f) for f in frames]
        dsts = [G.dev_empty(ocfg.dst_h, ocfg.dst_w) if fl else None for fl in flags]
        for lo, hi in ((0, 5), (5, n)):       # two consecutive batches continue the same temporal state
            sp = (C.c_void_p * (hi - lo))(*[t.data_ptr() for t in srcs[lo:hi]])
            dp = (C.c_void_p * (hi - lo))(*[d.data_ptr() if d is not None else None for d in dsts[lo:hi]])
            ctx.call("vp_process_batch_device", hi - lo, sp, sw * 3, dp, ocfg.dst_w * 3, None)
        ctx.call("vp_synchronize")
        for i in range(n):
            if flags[i]:
                assert np.array_equal(dsts[i].cpu().numpy(), ref[i]), i


@pytest.mark.parametrize("temporal", [RC.TEMPORAL_NONE, RC.TEMPORAL_MAIN_BLEND, RC.TEMPORAL_BLEND_FRAMES])
def test_batch_graph_replay_is_bit_identical(vp, temporal):
    """The CUDA-graph form of vp_process_batch_device: the same batch (same frame buffers, same ring position) runs
    eagerly, is captured on its second sighting and replayed afterwards; every pass must equal one vp_process_device call
    per frame - also after the input buffers' CONTENTS changed (the graph bakes pointers, not pixels) and with graphs
    switched off again."""
    sw, sh, n = 128, 72, 12
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=temporal)
    clips = [synth.clip("scene", sw, sh, n), synth.clip("noise", sw, sh, n)]
    flags = [False] * 3 + [True] * (n - 3)
    refs = [run_chain_device(vp, ocfg, fr, sw, sh, outputs=flags)[0] for fr in clips]
    cfg = G.chain_cfg(vp, ocfg, sw, sh)
    with vp.Context(cfg) as ctx:
        srcs = [G.to_dev(f) for f in clips[0]]
        dsts = [G.dev_empty(ocfg.dst_h, ocfg.dst_w) if fl else None for fl in flags]
        sp = (C.c_void_p * n)(*[t.data_ptr() for t in srcs])
        dp = (C.c_void_p * n)(*[d.data_ptr() if d is not None else None for d in dsts])

        def run_and_check(ref):
            ctx.call("vp_reset_temporal")
            for d in dsts:
                if d is not None:
                    d.zero_()
            ctx.call("vp_process_batch_device", n, sp, sw * 3, dp, ocfg.dst_w * 3, None)
            ctx.call("vp_synchronize")
            for i in range(n):
                if flags[i]:
                    assert np.array_equal(dsts[i].cpu().numpy(), ref[i]), i

        for _ in range(4):                    # eager, capture + launch, replay, replay
            run_and_check(refs[0])
        assert vp.lib.vp_batch_graph_replays(ctx.handle) == 3
        for t, f in zip(srcs, clips[1]):      # new pixels behind the same pointers
            t.copy_(G.to_dev(f))
        run_and_check(refs[1])
        assert vp.lib.vp_batch_graph_replays(ctx.handle) == 4
        ctx.call("vp_set_batch_graphs", 0)
        run_and_check(refs[1])
        assert vp.lib.vp_batch_graph_replays(ctx.handle) == 4
        # another lane count is another batch: eager once, captured on the next sighting, same bytes
        ctx.call("vp_set_batch_graphs", 1)
        ctx.call("vp_set_batch_lanes", 2)
        run_and_check(refs[1])
        run_and_check(refs[1])
        assert vp.lib.vp_batch_graph_replays(ctx.handle) == 5
        ctx.call("vp_set_batch_lanes", 4)
        ctx.call("vp_set_batch_graphs", 0)
        # switched on again: the cache was dropped, so one eager pass, then a fresh capture
        ctx.call("vp_set_batch_graphs", 1)
        run_and_check(refs[1])
        run_and_check(refs[1])
        assert vp.lib.vp_batch_graph_replays(ctx.handle) == 6


@pytest.mark.parametrize("size", [(160, 96), (131, 67), (320, 200)])
@pytest.mark.parametrize("kind", ["scene", "noise"])
def test_lane_buffer_format_is_transparent(vp, monkeypatch, size, kind):
    """The plain chain hands k_upsharp's output to the POST bilateral as 32-bit (B, G, R, 1) pixels whose halo tile arrives by
    one TMA box (interior tiles untouched, border tiles mirrored in place).  With VPB200_NO_BGRX the same chain uses the packed
    3-byte lane buffer and the expansion pass; both must produce the same bytes - odd sizes put every kind of border tile
    (left / right / top / bottom, partial last tile) on the 32-bit path."""
    sw, sh = size
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_BLEND_FRAMES)
    frames = synth.clip(kind, sw, sh, 5)
    monkeypatch.delenv("VPB200_NO_BGRX", raising=False)
    words, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    monkeypatch.setenv("VPB200_NO_BGRX", "1")
    packed, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    for i, (a, b) in enumerate(zip(words, packed)):
        assert np.array_equal(a, b), (i, G.diff_stats(a, b))


@pytest.mark.parametrize("kind", ["scene", "noise", "flat"])
@pytest.mark.parametrize("size,scale", [((160, 96), 2), ((131, 67), 2), ((96, 54), 4)])
def test_bicubic_wrapper_chain(vp, kind, size, scale):
    """CPUImpl's BICUBIC wrapper (GaussianBlur 3x3 -> INTER_CUBIC -> enhanceBicubicResult with Canny) followed by
    main.cpp's 0.6/0.4 smoothing.  Blur, resize, gray, Canny, dilate and the 5-point sharpen are integer and
    bit-exact; only the bilateral-filtered (non-edge) pixels may differ, by at most 1 LSB."""
    sw, sh = size
    ocfg = RC.ChainConfig(dst_w=scale * sw, dst_h=scale * sh, temporal=RC.TEMPORAL_MAIN_BLEND, bicubic_enhance=True)
    frames = synth.clip(kind, sw, sh, 3)
    oracle = RC.Chain(RC.NumpyOps(), ocfg)
    refs = [oracle.process(f) for f in frames]
    outs, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    for o, r in zip(outs, refs):
        mx, nd = G.diff_stats(o, r)
        assert mx <= 1 and nd <= o.size // 500, (mx, nd)


def test_bicubic_wrapper_full_size_against_cv2(vp):
    """BASELINE config C1 (1280x720 -> 2560x1440) through the wrapper, against the real OpenCV kernels."""
    sw, sh = 1280, 720
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_MAIN_BLEND, bicubic_enhance=True)
    frames = synth.clip("scene", sw, sh, 2)
    oracle = RC.Chain(RC.Cv2Ops(), ocfg)
    refs = [oracle.process(f) for f in frames]
    outs, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    for o, r in zip(outs, refs):
        mx, nd = G.diff_stats(o, r)
        assert mx <= 1 and nd <= o.size // 500, (mx, nd)


def test_host_paths_match_device_path(vp):
    sw, sh, n = 128, 72, 6
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_BLEND_FRAMES)
    frames = synth.clip("noise", sw, sh, n)
    dev, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    cfg = G.chain_cfg(vp, ocfg, sw, sh)
    # synchronous host call with pageable numpy buffers
    with vp.Context(cfg) as ctx:
        for f, d in zip(frames, dev):
            out = np.zeros((ocfg.dst_h, ocfg.dst_w, 3), np.uint8)
            ctx.call("vp_process_host", f.ctypes.data, f.strides[0], out.ctypes.data, out.strides[0])
            assert np.array_equal(out, d)
    # pipelined submit/wait with pinned buffers, depth frames in flight
    with vp.Context(cfg) as ctx:
        depth = vp.lib.vp_pipeline_depth(ctx.handle)
        pin_in = [torch.from_numpy(f).pin_memory() for f in frames]
        pin_out = [torch.zeros((ocfg.dst_h, ocfg.dst_w, 3), dtype=torch.uint8).pin_memory() for _ in frames]
        inflight = 0
        for i in range(n):
            if inflight == depth:
                ctx.call("vp_wait")
                inflight -= 1
            ctx.call("vp_submit_host", pin_in[i].data_ptr(), sw * 3, pin_out[i].data_ptr(), ocfg.dst_w * 3)
            inflight += 1
        assert vp.lib.vp_submit_host(ctx.handle, pin_in[0].data_ptr(), sw * 3, pin_out[0].data_ptr(), ocfg.dst_w * 3) in (vp.VP_OK, vp.VP_ERR_STATE)
        while True:
            rc = vp.lib.vp_wait(ctx.handle)
            if rc != vp.VP_OK:
                assert rc == vp.VP_ERR_STATE
                break
        for o, d in zip(pin_out, dev):
            assert np.array_equal(o.numpy(), d)


@pytest.mark.parametrize("flow", [False, True])
def test_host_pipeline_lanes_mixed_with_device_calls(vp, flow):
    """vp_submit_host runs consecutive frames on different compute lanes; device entry points issued while host frames
    are still in flight (no vp_wait in between) must see the temporal state those frames leave, and the other way round."""
    sw, sh, n = 640, 360, 13
    ocfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_BLEND_FRAMES)
    if flow:
        ocfg.temporal_cfg = RC.TemporalConfig(use_flow=True, scene_detect=True)
    frames = synth.clip("scene", sw, sh, n)
    dev, _ = run_chain_device(vp, ocfg, frames, sw, sh)
    cfg = G.chain_cfg(vp, ocfg, sw, sh)
    dw, dh = ocfg.dst_w, ocfg.dst_h
    with vp.Context(cfg) as ctx:
        pin_in = [torch.from_numpy(f).pin_memory() for f in frames]
        pin_out = [torch.zeros((dh, dw, 3), dtype=torch.uint8).pin_memory() for _ in frames]
        d_in = [G.to_dev(f) for f in frames]
        d_out = [G.dev_empty(dh, dw) for _ in frames]
        inflight = 0

        def submit(i):
            nonlocal inflight
            if inflight == vp.lib.vp_pipeline_depth(ctx.handle):
                ctx.call("vp_wait"); inflight -= 1
            ctx.call("vp_submit_host", pin_in[i].data_ptr(), sw * 3, pin_out[i].data_ptr(), dw * 3)
            inflight += 1

        for i in range(0, 4):
            submit(i)
        for i in range(4, 6):                                    # device calls right behind the frames in flight
            ctx.call("vp_process_device", d_in[i].data_ptr(), G.pitch(d_in[i]), d_out[i].data_ptr(), G.pitch(d_out[i]), None)
        for i in range(6, 9):
            submit(i)
        sp = (C.c_void_p * 4)(*[t.data_ptr() for t in d_in[9:13]])
        dp = (C.c_void_p * 4)(*[t.data_ptr() for t in d_out[9:13]])
        ctx.call("vp_process_batch_device", 4, sp, sw * 3, dp, dw * 3, None)
        while inflight:
            ctx.call("vp_wait"); inflight -= 1
        ctx.call("vp_synchronize")
        for i in range(n):
            got = d_out[i].cpu().numpy() if i in (4, 5) or i >= 9 else pin_out[i].numpy()
            assert np.array_equal(got, dev[i]), i


def test_full_size_c2_against_cv2(vp):
    """BASELINE config C2 at full size against the real OpenCV kernels (cv2) on one scene frame pair."""
    sw, sh = 1920, 1080
    ocfg = RC.ChainConfig(dst_w=3840, dst_h=2160, temporal=RC.TEMPORAL_MAIN_BLEND)
    frames = synth.clip("scene", sw, sh, 2)
    oracle = RC.Chain(RC.Cv2Ops(), ocfg)
    refs = [oracle.process(f) for f in frames]
    outs, params = run_chain_device(vp, ocfg, frames, sw, sh)
    assert params[0][0] == 11 and params[1][0] == 5
    for o, r in zip(outs, refs):
        mx, nd = G.diff_stats(o, r)
        assert RC.psnr(o, r) >= 50.0 and mx <= 4, (mx, nd)


@pytest.mark.parametrize("cfgname,sw,sh,scale", [("C1", 1280, 720, 2), ("C2", 1920, 1080, 2), ("C5", 960, 540, 4)])
def test_full_size_bicubic_bit_exact(vp, cfgname, sw, sh, scale):
    import cv2
    img = synth.frame("noise", sw, sh, 0)
    with vp.Context() as ctx:
        out = G.run_bicubic(vp, ctx, img, sw * scale, sh * scale)
    assert np.array_equal(out, cv2.resize(img, (sw * scale, sh * scale), interpolation=cv2.INTER_CUBIC))
