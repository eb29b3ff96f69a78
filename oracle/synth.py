"""Deterministic synthetic clips (TEST INFRASTRUCTURE - part of the oracle, never shipped).

Integer-only generators so Python / C++ / CUDA agree byte-for-byte.
  noise : uniform 8-bit hash noise (std-dev ~73.9)   -> noise_factor 1.5 in calculateAdaptiveParams
  scene : moving gradients + checker edges + light noise (std-dev ~50) -> noise_factor 1.5
  flat  : 128 + (noise & 3) (std-dev < 5)            -> noise_factor 0.7
All frames are HxWx3 uint8, BGR interleaved, C-contiguous (a continuous CV_8UC3 cv::Mat).
"""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def _mix32(h):
    h = h & _M32
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & _M32
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & _M32
    h ^= h >> np.uint64(15)
    return h


def noise_frame(w, h, t=0, seed=20260):
    x = np.arange(w, dtype=np.uint64)[None, :, None]
    y = np.arange(h, dtype=np.uint64)[:, None, None]
    c = np.arange(3, dtype=np.uint64)[None, None, :]
    k = ((x * np.uint64(0x9E3779B1)) & _M32) ^ ((y * np.uint64(0x85EBCA77)) & _M32) \
        ^ ((c * np.uint64(0xC2B2AE3D)) & _M32) ^ ((np.uint64(t + seed) * np.uint64(0x27D4EB2F)) & _M32)
    return (_mix32(k) & np.uint64(0xFF)).astype(np.uint8)


def _tri(v, p):
    return np.abs((v % (2 * p)) - p) * 255 // p


def scene_frame(w, h, t=0, seed=20260):
    x = np.arange(w, dtype=np.int64)[None, :, None]
    y = np.arange(h, dtype=np.int64)[:, None, None]
    c = np.arange(3, dtype=np.int64)[None, None, :]
    n = noise_frame(w, h, t, seed).astype(np.int64)
    v = _tri(x + 2 * t + 13 * c, 96) * 3 // 8 + _tri(y + t, 64) * 3 // 8 \
        + 64 * (((x >> 5) ^ (y >> 5) ^ (t >> 4)) & 1) + (n & 15)
    return np.clip(v, 0, 255).astype(np.uint8) + np.zeros((h, w, 3), np.uint8)


def flat_frame(w, h, t=0, seed=20260):
    return (128 + (noise_frame(w, h, t, seed) & 3)).astype(np.uint8)


KINDS = {"noise": noise_frame, "scene": scene_frame, "flat": flat_frame}


def frame(kind, w, h, t=0, seed=20260):
    return np.ascontiguousarray(KINDS[kind](w, h, t, seed))


def clip(kind, w, h, n, seed=20260, t0=0):
    return [frame(kind, w, h, t0 + t, seed) for t in range(n)]
