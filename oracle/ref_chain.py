"""CPU restatement of the reference's enhance-chain glue (TEST INFRASTRUCTURE ONLY).

Every function follows one body of /root/reference (file:line cited in each docstring) and calls an
`ops` backend for the OpenCV primitives it uses:

  Cv2Ops   - the REAL OpenCV kernels through cv2 (what the reference links against; 4.13.0 here,
             the author linked 4.12.0).  Strongest oracle; also the CPU baseline bench.py times.
  NumpyOps - oracle/restate.py, the CPU-independent numpy restatement of the same arithmetic
             (what the CUDA kernels are held bit-exact / +-1 LSB against, and what generates
             tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path never does: it fails loudly if libvpb200.so is missing.

Scope decisions where the literal reference code is undefined behaviour (SURVEY.md section 8(c),
Appendix B) -- each is an oracle decision, stated here once:
  * SelectiveBilateral: the plain branch (use_multiscale=false, selective=false) is the north-star path;
    createDetailMask calls cv::sqrt on a CV_8U Mat (selective_bilateral.cpp:276), so the selective and
    multi-scale branches are restated with REPAIRED semantics (local_var -> float32 before the sqrt, a zero
    maximum normalises to 0) - see create_detail_mask below.
  * AdaptiveSharpening: only adaptive_sigma=false is in scope; calculateTextureMap has the same
    defect (adaptive_sharpening.cpp:386).
  * TemporalConsistency::blendFrames writes Vec3f into a CV_8UC3 Mat (temporal_consistency.cpp:367
    vs :408); the oracle uses the evidently intended float32 accumulator.  The north-star blend takes
    previous frames unwarped with mask == 1 (TemporalBlendState); the motion-compensated front end
    (Farneback flow, remap warp, reliability masks) is TemporalConsistencyFull.
Parity status: the reference has no golden vectors and cannot be compiled here, so this glue is
"parity unpinned" against a reference binary; the arithmetic underneath is pinned against real
OpenCV (tests/test_oracle_restate.py) and frozen in tests/golden/.
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import restate as R

F32 = np.float32

PRE_PROCESSING = 0
POST_PROCESSING = 1


# ------------------------------------------------------------------------------------------------
# Config mirrors (field order == the reference's aggregate structs, callers brace-init positionally)
# ------------------------------------------------------------------------------------------------
@dataclass
class BilateralConfig:  # include/selective_bilateral.h:20-40
    stage: int = PRE_PROCESSING
    use_gpu: bool = True
    adaptive_params: bool = True
    diameter: int = 7
    sigma_color: float = 30.0
    sigma_space: float = 30.0
    selective: bool = True
    detail_threshold: float = 30.0
    texture_boost: float = 1.5
    edge_preserve: float = 2.0
    use_multiscale: bool = True
    num_scales: int = 3


@dataclass
class SharpenConfig:  # include/adaptive_sharpening.h:14-24
    strength: float = 0.8
    edge_strength: float = 1.2
    smooth_strength: float = 0.4
    edge_threshold: float = 30.0
    sigma: float = 1.5
    kernel_size: int = 5
    preserve_tone: bool = True
    use_gpu: bool = True
    adaptive_sigma: bool = True


@dataclass
class TemporalConfig:  # include/temporal_consistency.h:18-33
    buffer_size: int = 3
    blend_strength: float = 0.6
    motion_threshold: float = 15.0
    scene_change_threshold: float = 100.0
    use_gpu: bool = True
    pyr_scale: float = 0.5
    levels: int = 3
    winsize: int = 15
    iterations: int = 3
    poly_n: int = 5
    poly_sigma: float = 1.2
    flags: int = 0
    # oracle switches (not reference fields).  scene_detect: run detectSceneChange (temporal_consistency.cpp:283-323) and
    # reset the buffers on a cut (:100-105); off reproduces the blend alone.  use_flow: the motion-compensated front end
    # (Farneback flow, remap warp, reliability masks, :94-150) - TemporalConsistencyFull below; off = the north-star blend
    # of unwarped frames with unit masks (TemporalBlendState).
    scene_detect: bool = False
    use_flow: bool = False


# ------------------------------------------------------------------------------------------------
# primitive backends
# ------------------------------------------------------------------------------------------------
class NumpyOps:
    name = "numpy-restatement"

    def resize_cubic(self, img, dw, dh):
        return R.resize_cubic(img, dw, dh, "default")

    def bilateral(self, img, d, sc, ss):
        return R.bilateral_u8c3(img, d, sc, ss)

    def mean_stddev(self, img):
        m, s, _, _ = R.mean_stddev(img)
        return m, s

    def gaussian_u8(self, img, k, sigma):
        return R.gaussian_blur_u8(img, k, sigma)

    def gaussian_f32(self, img, k, sigma):
        return R.gaussian_blur_f32(img, k, sigma)

    def bgr2gray(self, img):
        return R.bgr2gray(img)

    def bgr2ycrcb(self, img):
        return R.bgr2ycrcb(img)

    def ycrcb2bgr(self, img):
        return R.ycrcb2bgr(img)

    def sobel_abs(self, gray, dx, dy):
        return R.abs_sat_u8(R.sobel_x(gray) if dx else R.sobel_y(gray))

    def laplacian_abs(self, gray):
        return R.abs_sat_u8(R.laplacian3(gray))

    def add_weighted(self, a, alpha, b, beta):
        return R.add_weighted_u8(a, alpha, b, beta)

    def subtract(self, a, b):
        return R.subtract_sat_u8(a, b)

    def canny(self, gray, low, high):
        return R.canny_u8(gray, low, high)

    def scene_metrics(self, prev_gray, cur_gray):
        return R.scene_metrics(prev_gray, cur_gray)

    def dilate_2x2(self, edges):
        return R.dilate_2x2(edges)

    def box_blur5(self, img):
        return R.box_blur5_u8(img)

    def box_mean7(self, img):
        return R.box_mean7_u8(img)

    def pyr_down(self, img):
        return R.pyr_down_u8(img)

    def pyr_up(self, img, dw, dh):
        return R.pyr_up_u8(img, dw, dh)

    def sobel_magnitude(self, gray):
        return R.sobel_magnitude_f32(gray)

    def add_weighted_f32(self, a, alpha, b, beta):
        return (a.astype(np.float64) * alpha + b.astype(np.float64) * beta).astype(F32)

    def farneback(self, prev_gray, cur_gray, t):
        return R.farneback(prev_gray, cur_gray, t.pyr_scale, t.levels, t.winsize, t.iterations, t.poly_n, t.poly_sigma)

    def remap_flow(self, frame, flow):
        mx, my = R.flow_to_maps(flow)
        return R.remap_linear_replicate_u8(frame, mx, my)


class Cv2Ops:
    name = "cv2"

    def __init__(self, threads=None):
        import cv2
        self.cv2 = cv2
        cv2.setUseOptimized(True)
        if threads is not None:
            cv2.setNumThreads(int(threads))

    def resize_cubic(self, img, dw, dh):
        return self.cv2.resize(img, (dw, dh), interpolation=self.cv2.INTER_CUBIC)

    def bilateral(self, img, d, sc, ss):
        return self.cv2.bilateralFilter(img, d, sc, ss)

    def mean_stddev(self, img):
        m, s = self.cv2.meanStdDev(img)
        return m.ravel()[:3], s.ravel()[:3]

    def gaussian_u8(self, img, k, sigma):
        return self.cv2.GaussianBlur(img, (k, k), sigma)

    def gaussian_f32(self, img, k, sigma):
        return self.cv2.GaussianBlur(img, (k, k), sigma)

    def bgr2gray(self, img):
        return self.cv2.cvtColor(img, self.cv2.COLOR_BGR2GRAY)

    def bgr2ycrcb(self, img):
        return self.cv2.cvtColor(img, self.cv2.COLOR_BGR2YCrCb)

    def ycrcb2bgr(self, img):
        return self.cv2.cvtColor(img, self.cv2.COLOR_YCrCb2BGR)

    def sobel_abs(self, gray, dx, dy):
        g = self.cv2.Sobel(gray, self.cv2.CV_16S, dx, dy, ksize=3)
        return self.cv2.convertScaleAbs(g)

    def laplacian_abs(self, gray):
        return self.cv2.convertScaleAbs(self.cv2.Laplacian(gray, self.cv2.CV_16S, ksize=3))

    def add_weighted(self, a, alpha, b, beta):
        return self.cv2.addWeighted(a, alpha, b, beta, 0)

    def subtract(self, a, b):
        return self.cv2.subtract(a, b)

    def canny(self, gray, low, high):
        return self.cv2.Canny(gray, low, high)

    def scene_metrics(self, prev_gray, cur_gray):
        cv2 = self.cv2
        ph = cv2.normalize(cv2.calcHist([prev_gray], [0], None, [64], [0, 256]), None, 0, 1, cv2.NORM_MINMAX)
        ch = cv2.normalize(cv2.calcHist([cur_gray], [0], None, [64], [0, 256]), None, 0, 1, cv2.NORM_MINMAX)
        return cv2.compareHist(ph, ch, cv2.HISTCMP_CORREL), cv2.mean(cv2.absdiff(prev_gray, cur_gray))[0]

    def dilate_2x2(self, edges):
        return self.cv2.dilate(edges, self.cv2.getStructuringElement(self.cv2.MORPH_RECT, (2, 2)))

    def box_blur5(self, img):
        return self.cv2.blur(img, (5, 5))

    def box_mean7(self, img):
        return self.cv2.filter2D(img, -1, np.ones((7, 7), F32) / F32(49.0))

    def pyr_down(self, img):
        return self.cv2.pyrDown(img)

    def pyr_up(self, img, dw, dh):
        return self.cv2.pyrUp(img, dstsize=(dw, dh))

    def sobel_magnitude(self, gray):
        cv2 = self.cv2
        return cv2.magnitude(cv2.Sobel(gray, cv2.CV_32F, 1, 0, ksize=3), cv2.Sobel(gray, cv2.CV_32F, 0, 1, ksize=3))

    def add_weighted_f32(self, a, alpha, b, beta):
        return self.cv2.addWeighted(a, alpha, b, beta, 0)

    def farneback(self, prev_gray, cur_gray, t):
        return self.cv2.calcOpticalFlowFarneback(prev_gray, cur_gray, None, t.pyr_scale, t.levels, t.winsize, t.iterations,
                                                 t.poly_n, t.poly_sigma, t.flags)

    def remap_flow(self, frame, flow):
        mx, my = R.flow_to_maps(flow)
        return self.cv2.remap(frame, mx, my, self.cv2.INTER_LINEAR, borderMode=self.cv2.BORDER_REPLICATE)


# ------------------------------------------------------------------------------------------------
# SelectiveBilateral (plain branch)
# ------------------------------------------------------------------------------------------------
def adaptive_params(ops, img, cfg):
    """SelectiveBilateral::calculateAdaptiveParams, selective_bilateral.cpp:315-371."""
    _, std = ops.mean_stddev(img)
    avg = 0.0
    for i in range(3):
        avg += float(std[i])
    avg /= 3
    return adaptive_params_from_std(avg, cfg)


def adaptive_params_from_std(avg_stddev, cfg):
    nf = 1.0
    if avg_stddev < 5.0:
        nf = 0.7
    elif avg_stddev > 15.0:
        nf = 1.5
    d = max(5, int(cfg.diameter * nf))
    if d % 2 == 0:
        d += 1
    d = min(d, 15)
    sc = cfg.sigma_color * nf
    ss = cfg.sigma_space * nf
    if cfg.stage == POST_PROCESSING:
        d = max(3, d - 2)
        sc *= 0.8
        ss *= 0.8
    return d, sc, ss


def bilateral_plain(ops, img, cfg):
    """SelectiveBilateral::applyBilateralFilter, selective_bilateral.cpp:84-119 (CPU branch :110)."""
    d, sc, ss = cfg.diameter, cfg.sigma_color, cfg.sigma_space
    if cfg.adaptive_params:
        d, sc, ss = adaptive_params(ops, img, cfg)
    return ops.bilateral(img, d, sc, ss)


# ------------------------------------------------------------------------------------------------
# SelectiveBilateral, selective and multiscale branches with REPAIRED semantics.
# The literal code calls cv::sqrt on the CV_8U `local_var` (selective_bilateral.cpp:276, undefined
# behaviour: the float kernel runs over a byte buffer); the repair converts `local_var` to float32 first
# and keeps everything else as written, including the saturating u8 subtract / multiply (:268-269).
# "Parity unpinned": no run of the reference can produce these numbers.
# ------------------------------------------------------------------------------------------------
def create_detail_mask(ops, img, cfg):
    """SelectiveBilateral::createDetailMask, selective_bilateral.cpp:239-313."""
    gray = ops.bgr2gray(img)                                        # :245
    magnitude = ops.sobel_magnitude(gray)                           # :250-256
    local_mean = ops.box_blur5(gray)                                # :263
    diff = ops.subtract(gray, local_mean)                           # :267-268 (u8, saturating)
    d16 = diff.astype(np.uint16)
    diff_sq = np.minimum(d16 * d16, 255).astype(np.uint8)           # :269 cv::multiply on u8 saturates
    local_var = ops.box_blur5(diff_sq)                              # :272
    texture = np.sqrt(local_var.astype(F32)).astype(F32)            # :276 REPAIRED: float32 before cv::sqrt
    # :281-285  Mat / double == convertTo(alpha = 1/max): a float32 multiply by (float)(1.0 / max)
    # REPAIRED as well: a zero maximum (flat frame) would make this 0 * inf = NaN and the output black; it yields 0
    mm, tm = float(magnitude.max()), float(texture.max())
    norm_mag = (magnitude * F32(1.0 / mm)).astype(F32) if mm > 0 else np.zeros_like(magnitude)
    norm_tex = (texture * F32(1.0 / tm)).astype(F32) if tm > 0 else np.zeros_like(texture)
    detail = ops.add_weighted_f32(norm_mag, 0.7, norm_tex, 0.3)    # :288
    mask = R.detail_sigmoid_f32(detail, cfg.detail_threshold)       # :291-303
    return ops.gaussian_f32(mask, 5, 1.0)                           # :306


def joint_blend(img, filtered, mask, edge_preserve):
    """applyJointBilateral's per-pixel loop, selective_bilateral.cpp:412-428: float32 arithmetic
    `orig * d * p + filt * (1 - d * p)`, saturate_cast<uchar>."""
    d = mask.astype(F32)[:, :, None]
    # `float preservation = 1.0f + detail_value * (m_config.edge_preserve - 1.0f)`: edge_preserve is a double, so the
    # right-hand side is evaluated in double and rounded once on the store
    p = (1.0 + d.astype(np.float64) * (float(edge_preserve) - 1.0)).astype(F32)
    dp = (d * p).astype(F32)
    a = ((img.astype(F32) * d).astype(F32) * p).astype(F32)
    b = (filtered.astype(F32) * (F32(1.0) - dp).astype(F32)).astype(F32)
    return np.clip(np.rint((a + b).astype(F32)), 0, 255).astype(np.uint8)


def bilateral_selective(ops, img, cfg):
    """SelectiveBilateral::applySelectiveBilateral, selective_bilateral.cpp:121-136 -> applyJointBilateral :373-439."""
    mask = create_detail_mask(ops, img, cfg)
    filtered = bilateral_plain(ops, img, cfg)
    return joint_blend(img, filtered, mask, cfg.edge_preserve)


def multiscale_blend(fine, coarse_up, mask):
    """selective_bilateral.cpp:198-228: result = saturate_cast<uchar>(fine * d + coarse * (1 - d)), float32."""
    d = mask.astype(F32)[:, :, None]
    cw = (F32(1.0) - d).astype(F32)
    v = ((fine.astype(F32) * d).astype(F32) + (coarse_up.astype(F32) * cw).astype(F32)).astype(F32)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def bilateral_multiscale(ops, img, cfg):
    """SelectiveBilateral::applyMultiscaleBilateral, selective_bilateral.cpp:138-237, CPU branch (:178): the
    config's diameter at every scale (no adaptive parameters here), sigmas x (1 + 0.5 i)."""
    n = max(1, min(int(cfg.num_scales), 5))                         # initialize(), :40-43
    scales = [img]
    for i in range(1, n):
        scales.append(ops.pyr_down(scales[i - 1]))                  # :146-151
    done = [ops.bilateral(s, cfg.diameter, cfg.sigma_color * (1.0 + 0.5 * i), cfg.sigma_space * (1.0 + 0.5 * i))
            for i, s in enumerate(scales)]                          # :155-182
    for i in range(n - 1, 0, -1):                                   # :185-230
        h, w = done[i - 1].shape[:2]
        up = ops.pyr_up(done[i], w, h)
        mask = create_detail_mask(ops, done[i - 1], cfg)
        done[i - 1] = multiscale_blend(done[i - 1], up, mask)
    return done[0]


def selective_bilateral_process(ops, img, cfg):
    """SelectiveBilateral::process dispatch, selective_bilateral.cpp:57-82."""
    if cfg.use_multiscale:
        return bilateral_multiscale(ops, img, cfg)
    if cfg.selective:
        return bilateral_selective(ops, img, cfg)
    return bilateral_plain(ops, img, cfg)


# ------------------------------------------------------------------------------------------------
# AdaptiveSharpening (adaptive_sigma = false)
# ------------------------------------------------------------------------------------------------
def edge_strength_u8(ops, img):
    """adaptive_sharpening.cpp:114-198: gray, Sobel x/y -> convertScaleAbs -> addWeighted(.5,.5),
    Laplacian -> convertScaleAbs, addWeighted(.6,.4).  Returns the CV_8U `combined_edges`."""
    gray = ops.bgr2gray(img)
    ax = ops.sobel_abs(gray, 1, 0)
    ay = ops.sobel_abs(gray, 0, 1)
    sobel = ops.add_weighted(ax, 0.5, ay, 0.5)
    lap = ops.laplacian_abs(gray)
    return ops.add_weighted(sobel, 0.6, lap, 0.4)


def edge_mask(ops, img, cfg):
    """AdaptiveSharpening::createEdgeMask, adaptive_sharpening.cpp:107-230.  The per-pixel sigmoid
    loop (:212-219) has a 256-value domain, so it is applied as the exact LUT of restate.sigmoid_lut."""
    comb = edge_strength_u8(ops, img)
    lut = R.sigmoid_lut(cfg.edge_threshold)
    sig = lut[comb]
    return ops.gaussian_f32(sig, 5, 1.5)  # :222 -- literal Size(5,5), 1.5 (not the config's)


def unsharp_mask(ops, img, mask, cfg):
    """AdaptiveSharpening::applyUnsharpMask, adaptive_sharpening.cpp:232-355 (3-channel branch).
    All float32, no FMA contraction (the reference is built without -O / -mfma)."""
    blurred = ops.gaussian_u8(img, cfg.kernel_size, float(F32(cfg.sigma)))
    return unsharp_from_blur(ops, img, blurred, mask, cfg)


def unsharp_from_blur(ops, img, blurred, mask, cfg):
    """The part applyUnsharpMask (:262-347) and applyVariableSigmaUnsharpMask (:503-585) share: one-sided
    unsharp mask, edge-adaptive strength, tone preservation."""
    um = ops.subtract(img, blurred)  # saturating: negative lobes dropped (:265, :505)
    e = mask.astype(F32)
    st, es, ss = F32(cfg.strength), F32(cfg.edge_strength), F32(cfg.smooth_strength)
    t1 = (e * es).astype(F32)
    t2 = ((F32(1.0) - e).astype(F32) * ss).astype(F32)
    strength = (st * (t1 + t2).astype(F32)).astype(F32)
    prod = (strength[..., None] * um.astype(F32)).astype(F32)
    sharp = (img.astype(F32) + prod).astype(F32)
    out = np.clip(np.rint(sharp), 0, 255).astype(np.uint8)
    if cfg.preserve_tone:  # :326-347
        yi = ops.bgr2ycrcb(img)
        yo = ops.bgr2ycrcb(out)
        yo[..., 1] = yi[..., 1]
        yo[..., 2] = yi[..., 2]
        out = ops.ycrcb2bgr(np.ascontiguousarray(yo))
    return out


# ---- adaptive_sigma = true with REPAIRED semantics -------------------------------------------
# calculateTextureMap calls cv::sqrt on the CV_8U `local_var` (adaptive_sharpening.cpp:386, undefined behaviour);
# the repair converts it to float32 first.  Degenerate ranges (max == min, a constant frame) divide by zero in the
# reference and send NaN through static_cast<uchar>; the repair maps them to 0.  "Parity unpinned".
def texture_map(ops, img):
    """AdaptiveSharpening::calculateTextureMap, adaptive_sharpening.cpp:357-403."""
    gray = ops.bgr2gray(img)
    local_mean = ops.box_mean7(gray)                                # :373 filter2D(gray, -1, ones(7,7)/49)
    diff = ops.subtract(gray, local_mean)                           # :377 u8, saturating
    d16 = diff.astype(np.uint16)
    diff_sq = np.minimum(d16 * d16, 255).astype(np.uint8)           # :379 cv::multiply on u8 saturates
    local_var = ops.box_mean7(diff_sq)                              # :383
    tex = np.sqrt(local_var.astype(F32)).astype(F32)                # :386 REPAIRED: float32 before cv::sqrt
    mn, mx = float(tex.min()), float(tex.max())                     # :390
    if mx <= mn:
        return np.zeros_like(tex)                                   # REPAIRED: 0/0
    # :391 (t - min) / (max - min) is one MatExpr: convertTo(alpha = 1/(max-min), beta = -min/(max-min)) in float32
    a, b = F32(1.0 / (mx - mn)), F32(-mn * (1.0 / (mx - mn)))
    return (tex.astype(np.float64) * np.float64(a) + np.float64(b)).astype(F32)


def adaptive_sigma_map(ops, tex):
    """AdaptiveSharpening::calculateAdaptiveSigma, adaptive_sharpening.cpp:405-440."""
    min_sigma, max_sigma = F32(0.8), F32(2.5)
    span = F32(max_sigma - min_sigma)
    sigma = (max_sigma - (tex * span).astype(F32)).astype(F32)      # :423
    return ops.gaussian_f32(sigma, 5, 1.0)                          # :431


SIGMA_LEVELS = 5


def variable_sigma_blur(ops, img, sigma_map, kernel_size):
    """The blur of applyVariableSigmaUnsharpMask, adaptive_sharpening.cpp:442-500: five Gaussian levels between the
    sigma map's extremes, per-pixel linear interpolation in float32, static_cast<uchar> (truncation)."""
    L = SIGMA_LEVELS
    mn, mx = F32(sigma_map.min()), F32(sigma_map.max())             # :457-460
    span = F32(mx - mn)
    levels = [F32(mn + F32(F32(span * F32(i)) / F32(L - 1))) for i in range(L)]                                           # :465
    blurred = [ops.gaussian_u8(img, kernel_size, float(lv)) for lv in levels]                                               # :466-468
    s = sigma_map.astype(F32)
    if span > 0:
        r = (((s - mn).astype(F32) / span).astype(F32) * F32(L - 1)).astype(F32)
        idx = np.clip(np.trunc(r).astype(np.int64), 0, L - 2)       # :477-478
    else:
        idx = np.zeros(s.shape, np.int64)                           # REPAIRED: 0/0
    lo = (mn + ((span * idx.astype(F32)).astype(F32) / F32(L - 1)).astype(F32)).astype(F32)          # :482
    hi = (mn + ((span * (idx + 1).astype(F32)).astype(F32) / F32(L - 1)).astype(F32)).astype(F32)    # :483
    den = (hi - lo).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = np.where(den > 0, ((s - lo).astype(F32) / den).astype(F32), F32(0.0)).astype(F32)    # :486 (REPAIRED 0/0)
    stack = np.stack(blurred)                                       # L x H x W x 3
    yy, xx = np.indices(s.shape)
    p_lo = stack[idx, yy, xx].astype(F32)
    p_hi = stack[idx + 1, yy, xx].astype(F32)
    a3 = alpha[..., None]
    v = ((p_lo * (F32(1.0) - a3).astype(F32)).astype(F32) + (p_hi * a3).astype(F32)).astype(F32)     # :496-498
    return np.clip(np.trunc(v), 0, 255).astype(np.uint8)            # static_cast<uchar>


def sharpen_variable_sigma(ops, img, cfg):
    """AdaptiveSharpening::process with adaptive_sigma=true, adaptive_sharpening.cpp:70-90."""
    mask = edge_mask(ops, img, cfg)
    sigma_map = adaptive_sigma_map(ops, texture_map(ops, img))
    blurred = variable_sigma_blur(ops, img, sigma_map, cfg.kernel_size)
    return unsharp_from_blur(ops, img, blurred, mask, cfg)


def sharpen(ops, img, cfg, honour_adaptive_sigma=False):
    """AdaptiveSharpening::process, adaptive_sharpening.cpp:51-105.  The north-star branch is adaptive_sigma=false;
    the adaptive_sigma=true body is the repaired restatement above and is used only when asked for."""
    if honour_adaptive_sigma and cfg.adaptive_sigma:
        return sharpen_variable_sigma(ops, img, cfg)
    return unsharp_mask(ops, img, edge_mask(ops, img, cfg), cfg)


# ------------------------------------------------------------------------------------------------
# CPUImpl BICUBIC wrapper (what Upscaler(BICUBIC) runs on the reference's CPU path)
# ------------------------------------------------------------------------------------------------
def enhance_bicubic_result(ops, image):
    """CPUImpl::enhanceBicubicResult, upscaler.cpp:98-148: bilateral(5,30,30) everywhere, 5-point sharpen
    on the dilated Canny(50,150) edges (interior pixels only, :118-133), per-pixel select (:138-147)."""
    blurred = ops.bilateral(image, 5, 30.0, 30.0)
    gray = ops.bgr2gray(image)
    edges = ops.canny(gray, 50, 150)
    mask = ops.dilate_2x2(edges) > 0          # convertTo(CV_8U, 1/255): 255 -> 1, then `> 0`
    im = image.astype(np.int32)
    sharp = image.copy()
    val = 5 * im[1:-1, 1:-1] - im[:-2, 1:-1] - im[2:, 1:-1] - im[1:-1, :-2] - im[1:-1, 2:]
    inner = np.clip(val, 0, 255).astype(np.uint8)
    sel = mask[1:-1, 1:-1]
    sharp[1:-1, 1:-1][sel] = inner[sel]
    return np.where(mask[..., None], sharp, blurred)


def cpuimpl_bicubic(ops, img, dst_w, dst_h):
    """CPUImpl::upscale for Upscaler::BICUBIC, upscaler.cpp:75-84: GaussianBlur 3x3 sigma 0.5 ->
    cv::resize(INTER_CUBIC) -> enhanceBicubicResult."""
    pre = ops.gaussian_u8(img, 3, 0.5)
    up = ops.resize_cubic(pre, dst_w, dst_h)
    return enhance_bicubic_result(ops, up)


# ------------------------------------------------------------------------------------------------
# temporal stages
# ------------------------------------------------------------------------------------------------
def blend_frames(frames, blend_strength, masks=None):
    """TemporalConsistency::blendFrames, temporal_consistency.cpp:355-444, with the intended float32
    accumulator.  frames[0] = current, frames[i] = i-th previous (already aligned); masks default 1.
    weight = exp(-(float)i/2.0f) * mask * blend_factor in float32 (:376,:405), accumulation in frame
    order (:411-418), `pixel /= weight` (cv::Vec /= float multiplies by 1.f/weight), convertTo RNE."""
    h, w = frames[0].shape[:2]
    acc = np.zeros((h, w, 3), dtype=F32)
    wsum = np.zeros((h, w), dtype=F32)
    for i, f in enumerate(frames):
        fw = F32(math.exp(float(F32(-float(i)) / F32(2.0))))  # std::exp(float) == expf, correctly rounded
        m = np.ones((h, w), dtype=F32) if masks is None or masks[i] is None else masks[i].astype(F32)
        bf = F32(1.0) if i == 0 else F32(blend_strength)
        wt = ((fw * m).astype(F32) * bf).astype(F32)
        acc = (acc + (wt[..., None] * f.astype(F32)).astype(F32)).astype(F32)
        wsum = (wsum + wt).astype(F32)
    inv = (F32(1.0) / wsum).astype(F32)
    res = (acc * inv[..., None]).astype(F32)
    res = np.where(wsum[..., None] > 0, res, frames[0].astype(F32))
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


def detect_scene_change(ops, prev_gray, cur_gray, threshold):
    """TemporalConsistency::detectSceneChange, temporal_consistency.cpp:283-323."""
    corr, mad = ops.scene_metrics(prev_gray, cur_gray)
    combined = (1.0 - corr) * 100.0 + mad * 0.5
    return combined > float(F32(threshold))


class TemporalBlendState:
    """TemporalConsistency::process buffer policy (temporal_consistency.cpp:61-172) with the
    flow/warp front end out of scope: first frame passes through; afterwards the current frame is
    blended with EVERY buffered *unblended* input - up to buffer_size of them: m_frame_buffer keeps the
    last buffer_size inputs (:162-165) and m_flow_buffer, trimmed to buffer_size-1 only after the blend
    (:167-169), holds buffer_size flows once the new one is pushed (:97), so the loop bound of :127 is
    min(frames, flows) = len(frames); the unblended input is what is stored (:158).  With
    cfg.scene_detect a detected cut clears the buffers and passes the frame through (:100-112)."""

    def __init__(self, cfg=None, ops=None):
        self.cfg = cfg or TemporalConfig()
        self.ops = ops
        self.frames = []

    def reset(self):
        self.frames = []

    def process(self, cur):
        if not self.frames:
            self.frames.append(cur.copy())
            return cur.copy()
        if self.cfg.scene_detect:
            ops = self.ops or NumpyOps()
            if detect_scene_change(ops, ops.bgr2gray(self.frames[-1]), ops.bgr2gray(cur), self.cfg.scene_change_threshold):
                self.frames = [cur.copy()]
                return cur.copy()
        nprev = min(len(self.frames), self.cfg.buffer_size)
        fr = [cur] + [self.frames[-1 - i] for i in range(nprev)]
        out = blend_frames(fr, self.cfg.blend_strength) if len(fr) > 1 else cur.copy()
        self.frames.append(cur.copy())
        while len(self.frames) > self.cfg.buffer_size:
            self.frames.pop(0)
        return out


def flow_reliability_mask(ops, flow, motion_threshold):
    """TemporalConsistency::calculateFlowReliabilityMask, temporal_consistency.cpp:325-353."""
    return ops.gaussian_f32(R.flow_reliability_raw(flow, motion_threshold), 15, 5.0)   # :347


class TemporalConsistencyFull:
    """TemporalConsistency::process exactly as written, temporal_consistency.cpp:61-172 (CPU branches): gray buffer,
    scene-change reset, Farneback flow prev->cur, every buffered frame warped with ITS OWN stored flow (frame t-1-i
    with the flow computed when frame t-i arrived, :127-131 - not a flow composed up to the current frame) and weighted
    by that flow's reliability mask, blendFrames with the float32 accumulator."""

    def __init__(self, cfg=None, ops=None):
        self.cfg = cfg or TemporalConfig()
        self.ops = ops or NumpyOps()
        self.reset()

    def reset(self):
        self.frames, self.grays, self.flows = [], [], []

    def process(self, cur, debug=None):
        ops, cfg = self.ops, self.cfg
        gray = ops.bgr2gray(cur)                                           # :74
        if not self.frames:                                                # :79-86
            self.frames.append(cur.copy()); self.grays.append(gray)
            return cur.copy()
        if detect_scene_change(ops, self.grays[-1], gray, cfg.scene_change_threshold):   # :89-92, :100-112
            self.frames, self.grays, self.flows = [cur.copy()], [gray], []
            return cur.copy()
        flow = ops.farneback(self.grays[-1], gray, cfg)                    # :95-97
        self.flows.append(flow)
        frames, masks = [cur], [None]                                      # :121-124 (identity mask)
        for i in range(min(len(self.flows), len(self.frames))):            # :127
            fl = self.flows[-1 - i]
            frames.append(ops.remap_flow(self.frames[-1 - i], fl))         # :133-135
            masks.append(flow_reliability_mask(ops, fl, cfg.motion_threshold))   # :137-138
        if debug is not None:
            debug.update(flow=flow, frames=frames, masks=masks)
        out = blend_frames(frames, cfg.blend_strength, masks)              # :147-148
        self.frames.append(cur.copy()); self.grays.append(gray)            # :158-159
        while len(self.frames) > cfg.buffer_size:                          # :162-165
            self.frames.pop(0); self.grays.pop(0)
        while len(self.flows) >= cfg.buffer_size:                          # :167-169
            self.flows.pop(0)
        return out


class MainBlendState:
    """main.cpp:269-292: history of processed (unblended) frames; from the second frame on
    out = addWeighted(cur, w, prev, 1-w), w = 0.6 (bicubic) / 0.7 (super-res)."""

    def __init__(self, ops, current_weight=0.6):
        self.ops = ops
        self.cw = F32(current_weight)
        self.pw = F32(1.0) - self.cw
        self.prev = None

    def reset(self):
        self.prev = None

    def process(self, cur):
        out = cur.copy() if self.prev is None else self.ops.add_weighted(cur, float(self.cw), self.prev, float(self.pw))
        self.prev = cur.copy()
        return out


# ------------------------------------------------------------------------------------------------
# the chain: Upscaler::upscale REAL_ESRGAN branch with the bicubic substitute
# ------------------------------------------------------------------------------------------------
TEMPORAL_NONE, TEMPORAL_MAIN_BLEND, TEMPORAL_BLEND_FRAMES = 0, 1, 2


@dataclass
class ChainConfig:
    dst_w: int = 0
    dst_h: int = 0
    use_pre: bool = True
    use_sharpen: bool = True
    use_post: bool = True
    temporal: int = TEMPORAL_MAIN_BLEND
    # in-scope variants of Upscaler::initializeEnhancements (upscaler.cpp:666-758): same numbers,
    # with the UB branches (multiscale / selective / adaptive_sigma) switched off.
    pre: BilateralConfig = field(default_factory=lambda: BilateralConfig(
        PRE_PROCESSING, True, True, 7, 30.0, 30.0, False, 30.0, 1.5, 2.0, False, 3))
    sharpen: SharpenConfig = field(default_factory=lambda: SharpenConfig(
        0.8, 1.2, 0.4, 30.0, 1.5, 5, True, True, False))
    post: BilateralConfig = field(default_factory=lambda: BilateralConfig(
        POST_PROCESSING, True, True, 5, 20.0, 20.0, False, 30.0, 1.2, 1.5, False, 2))
    temporal_cfg: TemporalConfig = field(default_factory=TemporalConfig)
    main_blend_weight: float = 0.6
    # True: the spatial part is CPUImpl's BICUBIC wrapper (upscaler.cpp:75-148) instead of the
    # enhancement chain; main.cpp then applies its 0.6/0.4 smoothing (temporal = TEMPORAL_MAIN_BLEND)
    bicubic_enhance: bool = False


class Chain:
    """Upscaler::upscale enhancement chain, upscaler.cpp:772-829 == test_enhancements.cpp:333-377,
    with step 2 = cv::resize(INTER_CUBIC) (test_enhancements.cpp:346-349)."""

    def __init__(self, ops, cfg):
        self.ops, self.cfg = ops, cfg
        if cfg.temporal == TEMPORAL_MAIN_BLEND:
            self.temporal = MainBlendState(ops, cfg.main_blend_weight)
        elif cfg.temporal == TEMPORAL_BLEND_FRAMES:
            self.temporal = (TemporalConsistencyFull if cfg.temporal_cfg.use_flow else TemporalBlendState)(cfg.temporal_cfg, ops)
        else:
            self.temporal = None

    def reset(self):
        if self.temporal:
            self.temporal.reset()

    def spatial(self, frame, stages=None):
        """Everything before the temporal stage; optionally records the stage outputs."""
        ops, cfg = self.ops, self.cfg
        if cfg.bicubic_enhance:
            return cpuimpl_bicubic(ops, frame, cfg.dst_w, cfg.dst_h)
        x = selective_bilateral_process(ops, frame, cfg.pre) if cfg.use_pre else frame
        if stages is not None:
            stages["pre"] = x
        x = ops.resize_cubic(x, cfg.dst_w, cfg.dst_h)
        if stages is not None:
            stages["up"] = x
        if cfg.use_sharpen:
            x = sharpen(ops, x, cfg.sharpen, honour_adaptive_sigma=True)
        if stages is not None:
            stages["sharp"] = x
        if cfg.use_post:
            x = selective_bilateral_process(ops, x, cfg.post)
        if stages is not None:
            stages["post"] = x
        return x

    def process(self, frame, stages=None):
        x = self.spatial(frame, stages)
        return self.temporal.process(x) if self.temporal else x


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * math.log10(255.0 * 255.0 / mse)
