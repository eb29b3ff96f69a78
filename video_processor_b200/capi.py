"""ctypes binding of include/vpb200.h (the C ABI of libvpb200.so).

This is the thin FFI stub tests, bench.py and __graft_entry__ use to reach the library; the product's
host side is the C++ in csrc/host (the reference is C++).  There is no fallback: if the shared object
is missing or fails to load, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VPB200_LIB: a development build of the same library (video_processor_b200/build.py --variant) for A/B timing
LIB_PATH = os.environ.get("VPB200_LIB") or os.path.join(_HERE, "libvpb200.so")

VP_OK = 0
VP_ERR_INVALID_ARG, VP_ERR_UNSUPPORTED, VP_ERR_CUDA, VP_ERR_NO_DEVICE, VP_ERR_STATE, VP_ERR_NOMEM = -1, -2, -3, -4, -5, -6
VP_STAGE_PRE, VP_STAGE_POST = 0, 1
VP_TEMPORAL_NONE, VP_TEMPORAL_MAIN_BLEND, VP_TEMPORAL_BLEND_FRAMES = 0, 1, 2


class BilateralConfig(C.Structure):
    _fields_ = [("stage", C.c_int32), ("adaptive_params", C.c_int32), ("diameter", C.c_int32),
                ("reserved_", C.c_int32), ("sigma_color", C.c_double), ("sigma_space", C.c_double),
                ("selective", C.c_int32), ("use_multiscale", C.c_int32), ("num_scales", C.c_int32),
                ("reserved2_", C.c_int32), ("detail_threshold", C.c_double), ("edge_preserve", C.c_double)]


class SharpenConfig(C.Structure):
    _fields_ = [("strength", C.c_float), ("edge_strength", C.c_float), ("smooth_strength", C.c_float),
                ("edge_threshold", C.c_float), ("sigma", C.c_float), ("kernel_size", C.c_int32),
                ("preserve_tone", C.c_int32), ("adaptive_sigma", C.c_int32)]


class TemporalConfig(C.Structure):
    _fields_ = [("mode", C.c_int32), ("buffer_size", C.c_int32), ("blend_strength", C.c_float),
                ("main_weight", C.c_float), ("scene_detect", C.c_int32), ("scene_change_threshold", C.c_float),
                ("use_flow", C.c_int32), ("motion_threshold", C.c_float), ("pyr_scale", C.c_double),
                ("poly_sigma", C.c_double), ("levels", C.c_int32), ("winsize", C.c_int32), ("iterations", C.c_int32),
                ("poly_n", C.c_int32), ("flags", C.c_int32), ("reserved3_", C.c_int32)]


class SegmentReport(C.Structure):
    _fields_ = [("device", C.c_int32), ("first_frame", C.c_int32), ("start", C.c_int32), ("stop", C.c_int32),
                ("status", C.c_int32), ("reserved_", C.c_int32), ("seconds", C.c_double)]


FRAME_SOURCE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t)
FRAME_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t)


class ChainConfig(C.Structure):
    _fields_ = [("src_w", C.c_int32), ("src_h", C.c_int32), ("dst_w", C.c_int32), ("dst_h", C.c_int32),
                ("use_pre", C.c_int32), ("use_sharpen", C.c_int32), ("use_post", C.c_int32), ("bicubic_enhance", C.c_int32),
                ("pre", BilateralConfig), ("sharpen", SharpenConfig), ("post", BilateralConfig),
                ("temporal", TemporalConfig)]


_vp = C.c_void_p
_u8p = C.c_void_p      # device or host address passed as an integer
_sz = C.c_size_t

# name -> (restype, argtypes); every symbol include/vpb200.h declares
PROTOTYPES = {
    "vp_abi_version": (C.c_int, []),
    "vp_device_count": (C.c_int, []),
    "vp_default_chain_config": (C.c_int, [C.POINTER(ChainConfig), C.c_int, C.c_int, C.c_int, C.c_int]),
    "vp_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.POINTER(ChainConfig)]),
    "vp_destroy": (None, [_vp]),
    "vp_last_error": (C.c_char_p, [_vp]),
    "vp_launch_count": (C.c_uint64, [_vp]),
    "vp_synchronize": (C.c_int, [_vp]),
    "vp_profile_enable": (C.c_int, [_vp, C.c_int]),
    "vp_profile_read": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "vp_process_device": (C.c_int, [_vp, _u8p, _sz, _u8p, _sz, C.c_void_p]),
    "vp_set_batch_lanes": (C.c_int, [_vp, C.c_int]),
    "vp_set_batch_graphs": (C.c_int, [_vp, C.c_int]),
    "vp_batch_graph_replays": (C.c_uint64, [_vp]),
    "vp_process_batch_device": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_void_p), _sz, C.POINTER(C.c_void_p), _sz, C.c_void_p]),
    "vp_process_host": (C.c_int, [_vp, _u8p, _sz, _u8p, _sz]),
    "vp_pipeline_depth": (C.c_int, [_vp]),
    "vp_submit_host": (C.c_int, [_vp, _u8p, _sz, _u8p, _sz]),
    "vp_wait": (C.c_int, [_vp]),
    "vp_plan_segment": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(TemporalConfig), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vp_run_segments": (C.c_int, [C.POINTER(ChainConfig), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p), _sz,
                                  C.POINTER(C.c_void_p), _sz, C.POINTER(SegmentReport)]),
    "vp_run_segments_cb": (C.c_int, [C.POINTER(ChainConfig), C.POINTER(C.c_int), C.c_int, C.c_int, FRAME_SOURCE, FRAME_SINK,
                                     C.c_void_p, C.POINTER(SegmentReport)]),
    "vp_segments_last_error": (C.c_char_p, []),
    "vp_reset_temporal": (C.c_int, [_vp]),
    "vp_get_stage_params": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "vp_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), _sz]),
    "vp_host_free": (C.c_int, [C.c_void_p]),
    "vp_bicubic": (C.c_int, [_vp, _u8p, _sz, C.c_int, C.c_int, _u8p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_bilateral": (C.c_int, [_vp, _u8p, _sz, _u8p, _sz, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]),
    "vp_selective_bilateral": (C.c_int, [_vp, C.POINTER(BilateralConfig), _u8p, _sz, _u8p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_detail_mask": (C.c_int, [_vp, C.c_double, _u8p, _sz, C.c_void_p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_pyr_down": (C.c_int, [_vp, _u8p, _sz, C.c_int, C.c_int, _u8p, _sz, C.c_void_p]),
    "vp_pyr_up": (C.c_int, [_vp, _u8p, _sz, C.c_int, C.c_int, _u8p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_mean_stddev_sums": (C.c_int, [_vp, _u8p, _sz, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vp_edge_mask": (C.c_int, [_vp, C.POINTER(SharpenConfig), _u8p, _sz, C.c_void_p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_unsharp": (C.c_int, [_vp, C.POINTER(SharpenConfig), _u8p, _sz, C.c_void_p, _sz, _u8p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_sharpen": (C.c_int, [_vp, C.POINTER(SharpenConfig), _u8p, _sz, _u8p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_sigma_map": (C.c_int, [_vp, _u8p, _sz, C.c_void_p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_optical_flow": (C.c_int, [_vp, C.POINTER(TemporalConfig), _u8p, _u8p, _sz, C.c_void_p, C.c_void_p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_warp_flow": (C.c_int, [_vp, _u8p, _sz, C.c_void_p, C.c_void_p, _sz, _u8p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_flow_reliability": (C.c_int, [_vp, C.c_void_p, C.c_void_p, _sz, C.c_float, C.c_void_p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_add_weighted": (C.c_int, [_vp, _u8p, _sz, C.c_double, _u8p, _sz, C.c_double, _u8p, _sz, C.c_int, C.c_int, C.c_void_p]),
    "vp_temporal_blend": (C.c_int, [_vp, C.POINTER(C.c_void_p), C.c_int, _sz, C.c_float, _u8p, _sz, C.c_int, C.c_int, C.c_void_p]),
}


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m video_processor_b200.build` "
            "(there is no CPU or PyTorch fallback for the enhance chain)")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class VpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vpb200 error {code}: {msg}")
        self.code = code


def check(rc, ctx=None):
    if rc != VP_OK:
        msg = lib.vp_last_error(ctx)
        raise VpError(rc, msg.decode() if msg else "")
    return rc


def default_chain_config(src_w, src_h, dst_w, dst_h):
    cfg = ChainConfig()
    check(lib.vp_default_chain_config(C.byref(cfg), src_w, src_h, dst_w, dst_h))
    return cfg


class Context:
    """Owns one vp_ctx.  Frames are passed as integer addresses + pitches (device or host)."""

    def __init__(self, cfg=None, device=0):
        self.handle = _vp()
        rc = lib.vp_create(C.byref(self.handle), device, C.byref(cfg) if cfg is not None else None)
        if rc != VP_OK:
            msg = lib.vp_last_error(None)
            raise VpError(rc, msg.decode() if msg else "")
        self.cfg = cfg

    def close(self):
        if self.handle:
            lib.vp_destroy(self.handle)
            self.handle = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def call(self, name, *args):
        return check(getattr(lib, name)(self.handle, *args), self.handle)

    @property
    def launches(self):
        return int(lib.vp_launch_count(self.handle))

    def profile_read(self):
        ms = (C.c_double * 6)()
        n = (C.c_uint64 * 6)()
        self.call("vp_profile_read", ms, n, 6)
        names = ["stats", "bilateral_pre", "upsharp", "bilateral_post", "blend", "bilateral_other"]
        return {k: (ms[i], int(n[i])) for i, k in enumerate(names)}

    def stage_params(self, which):
        d, sc, ss = C.c_int(), C.c_double(), C.c_double()
        self.call("vp_get_stage_params", which, C.byref(d), C.byref(sc), C.byref(ss))
        return d.value, sc.value, ss.value
