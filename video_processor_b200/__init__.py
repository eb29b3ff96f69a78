"""video_processor_b200 - B200 (sm_100a) implementation of video-processor's per-frame enhance chain.

The product is `libvpb200.so` (hand-written CUDA kernels + the C ABI of include/vpb200.h + the C++
Upscaler / SelectiveBilateral / AdaptiveSharpening / TemporalConsistency / FrameBuffer classes that
mirror the reference headers).  This package only carries the build script and the ctypes stub
(`capi`) that tests and bench.py use; importing `capi` fails loudly when the library is absent.
"""
__all__ = ["build", "capi"]
