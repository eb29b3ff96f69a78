"""Helpers for the -m gpu parity tests: device buffers come from torch (plumbing only), every
compute call goes through the C ABI of libvpb200.so."""
import ctypes as C

import numpy as np
import torch


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def dev_empty(h, w, c=3, dtype=torch.uint8):
    shape = (h, w, c) if c else (h, w)
    return torch.empty(shape, dtype=dtype, device="cuda")


def pitch(t):
    return t.stride(0) * t.element_size()


def padded_dev(a, pad_bytes):
    """Copy HxWx3 uint8 `a` into a device buffer whose rows are pad_bytes longer (odd pitch tests)."""
    h, w, c = a.shape
    buf = torch.zeros((h, w * c + pad_bytes), dtype=torch.uint8, device="cuda")
    buf[:, : w * c] = torch.from_numpy(a.reshape(h, w * c)).cuda()
    return buf


def sync():
    torch.cuda.synchronize()


def diff_stats(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()), int((d > 0).sum())


def run_bicubic(vp, ctx, img, dw, dh):
    src = to_dev(img)
    dst = dev_empty(dh, dw)
    ctx.call("vp_bicubic", src.data_ptr(), pitch(src), img.shape[1], img.shape[0], dst.data_ptr(), pitch(dst), dw, dh, None)
    sync()
    return dst.cpu().numpy()


def run_bilateral(vp, ctx, img, d, sc, ss):
    src = to_dev(img)
    dst = dev_empty(*img.shape[:2])
    ctx.call("vp_bilateral", src.data_ptr(), pitch(src), dst.data_ptr(), pitch(dst), img.shape[1], img.shape[0], d, float(sc), float(ss), None)
    sync()
    return dst.cpu().numpy()


def sharpen_cfg(vp, o):
    return vp.SharpenConfig(o.strength, o.edge_strength, o.smooth_strength, o.edge_threshold, o.sigma,
                            o.kernel_size, int(o.preserve_tone), int(o.adaptive_sigma))


def bilateral_cfg(vp, o):
    return vp.BilateralConfig(o.stage, int(o.adaptive_params), o.diameter, 0, o.sigma_color, o.sigma_space,
                              int(o.selective), int(o.use_multiscale), o.num_scales, 0, o.detail_threshold, o.edge_preserve)


def chain_cfg(vp, o, sw, sh):
    """oracle.ref_chain.ChainConfig -> capi.ChainConfig"""
    c = vp.ChainConfig()
    c.src_w, c.src_h, c.dst_w, c.dst_h = sw, sh, o.dst_w, o.dst_h
    c.use_pre, c.use_sharpen, c.use_post = int(o.use_pre), int(o.use_sharpen), int(o.use_post)
    c.bicubic_enhance = int(getattr(o, "bicubic_enhance", False))
    c.pre = bilateral_cfg(vp, o.pre)
    c.sharpen = sharpen_cfg(vp, o.sharpen)
    c.post = bilateral_cfg(vp, o.post)
    t = o.temporal_cfg
    c.temporal = vp.TemporalConfig(o.temporal, t.buffer_size, t.blend_strength, o.main_blend_weight,
                                   int(getattr(t, "scene_detect", False)), t.scene_change_threshold,
                                   int(getattr(t, "use_flow", False)), t.motion_threshold, t.pyr_scale, t.poly_sigma,
                                   t.levels, t.winsize, t.iterations, t.poly_n, t.flags, 0)
    return c


def temporal_cfg(vp, t):
    """oracle TemporalConfig -> capi.TemporalConfig (flow parameters for the stage-wise flow entry points)"""
    return vp.TemporalConfig(vp.VP_TEMPORAL_BLEND_FRAMES, t.buffer_size, t.blend_strength, 0.6, 1, t.scene_change_threshold,
                             1, t.motion_threshold, t.pyr_scale, t.poly_sigma, t.levels, t.winsize, t.iterations, t.poly_n, t.flags, 0)
