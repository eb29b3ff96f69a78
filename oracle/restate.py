"""Pure-numpy restatement of the OpenCV arithmetic on the enhance-chain hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (video_processor_b200/, include/) may import
this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

The reference (skandvarma/video-processor) has no arithmetic of its own on this path: every
number comes out of OpenCV (third-party, NOT vendored; the author linked 4.12.0, see
/root/reference/build/CMakeFiles/video_processor.dir/link.txt).  This file restates the published
OpenCV 4.x algorithms (modules/imgproc/src/{resize,smooth,color_yuv,deriv,bilateral_filter}.cpp,
modules/core/src/{arithm,mean}.cpp) so parity does not depend on which SIMD path a given host CPU
dispatches.  Each function names the reference call site it serves.  tests/test_oracle_*.py pin
every function here against the real OpenCV kernels through `cv2` 4.13.0 and against the
committed golden vectors (tests/golden/).

Pinning status: the reference ships no golden vectors / KATs, and its C++
cannot be compiled here (OpenCV C++ absent), so the *glue* in oracle/ref_chain.py is "parity
unpinned" against the reference binary; the OpenCV arithmetic restated here IS pinned, against the
real library.
"""
import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------
# borders
# --------------------------------------------------------------------------------------------
def reflect101(i, n):
    """cv::BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba), valid for any offset."""
    i = np.asarray(i)
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def pad_reflect101(img, r):
    h, w = img.shape[:2]
    ys = reflect101(np.arange(-r, h + r), h)
    xs = reflect101(np.arange(-r, w + r), w)
    return img[ys][:, xs]


# --------------------------------------------------------------------------------------------
# cv::resize(INTER_CUBIC) on CV_8UC3  -- reference call sites upscaler.cpp:81,
# test_enhancements.cpp:348, main.cpp:265
# --------------------------------------------------------------------------------------------
def cubic_coeffs(src_len, dst_len):
    """Per-destination-index source offset and 11-bit fixed-point taps.

    Follows cv::resize's table build (resize.cpp, `interpolateCubic` + the xofs/ialpha loop):
    scale = 1/(dst/src) in double, f = (float)((d+0.5)*scale-0.5), s=floor(f), A=-0.75 in float32,
    taps = saturate_cast<short>(c*2048) (round-half-even).  Returns (ofs[int32 dst_len] = s-1,
    coef[int16 dst_len,4]); tap k reads source index clamp(ofs+k, 0, src_len-1).
    """
    inv_scale = float(dst_len) / float(src_len)
    scale = 1.0 / inv_scale
    d = np.arange(dst_len, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(F32)
    s = np.floor(f).astype(np.int32)
    x = (f - s.astype(F32)).astype(F32)
    A = F32(-0.75)
    one = F32(1.0)
    xp1 = (x + one).astype(F32)
    omx = (one - x).astype(F32)
    c0 = ((A * xp1 - F32(5) * A) * xp1 + F32(8) * A) * xp1 - F32(4) * A
    c1 = ((A + F32(2)) * x - (A + F32(3))) * x * x + one
    c2 = ((A + F32(2)) * omx - (A + F32(3))) * omx * omx + one
    c3 = one - c0 - c1 - c2
    c = np.stack([c0, c1, c2, c3], axis=1).astype(F32)
    coef = np.rint(c * F32(2048)).astype(np.int16)
    return (s - 1).astype(np.int32), coef


GENERIC_VEC_LANES = 8  # VTraits<v_int16>::vlanes() of the SSE-baseline VResizeCubicVec_32s8u


def resize_cubic(src, dst_w, dst_h, path="default"):
    """Bit-exact model of cv::resize(src8U, dst, Size(dst_w,dst_h), 0, 0, INTER_CUBIC).

    Horizontal pass: int32 with 11-bit short taps (|H| <= 595,680 < 2^24, exact in float32).
    Vertical pass, b_k = (float)beta_k * 2^-22, S_k = (float)H[ytap_k]:
      path="default" - what a stock OpenCV build (IPP on, useOptimized) returns for x2 and x4, the
          only scales on the hot path; verified here against cv2 4.13.0 incl. exact .5 ties and odd
          row lengths:  t = fma(S0,b0, S1*b1) + fma(S2,b2, S3*b3);  dst = sat_u8(rint(t)) for
          EVERY element (no separate row tail).
      path="generic" - OpenCV's own C++ path (cv2.ipp.setUseIPP(False) / setUseOptimized(False)),
          valid for any scale: vector body (first (width*cn)//8*8 elements of a row)
          t = S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), every op rounded to float32, no FMA,
          dst = sat_u8(rint(t)); row tail uses the integer FixedPtCast
          dst = sat_u8((sum_k H_k*beta_k + 2^21) >> 22).
    At x2 the two paths agree everywhere except on the generic path's integer row tail at exact ties.
    """
    src = np.asarray(src)
    if src.ndim == 2:
        src = src[:, :, None]
    sh, sw, cn = src.shape
    xofs, alpha = cubic_coeffs(sw, dst_w)
    yofs, beta = cubic_coeffs(sh, dst_h)
    H = np.zeros((sh, dst_w, cn), dtype=np.int64)
    s = src.astype(np.int64)
    for k in range(4):
        xi = np.clip(xofs + k, 0, sw - 1)
        H += s[:, xi, :] * alpha[:, k].astype(np.int64)[None, :, None]
    H = H.reshape(sh, dst_w * cn)
    rows = [H[np.clip(yofs + k, 0, sh - 1)] for k in range(4)]  # each (dst_h, dst_w*cn)
    n = dst_w * cn
    b = [(beta[:, k].astype(F32) * F32(1.0 / 4194304.0))[:, None] for k in range(4)]
    S = [rows[k].astype(F32) for k in range(4)]
    if path == "default":
        # fma evaluated through float64: the product of two 24-bit significands is exact in
        # float64 and the following add of a float32 is correctly rounded to float64; the double
        # rounding float64->float32 cannot differ from a true fma here because |S*b| < 2^9 carries
        # at most 48 significant bits + alignment < 53 (checked exhaustively in tests on ties).
        def fma32(a, bb, c):
            return (a.astype(np.float64) * bb.astype(np.float64) + c.astype(np.float64)).astype(F32)
        p1 = (S[1] * b[1]).astype(F32)
        p3 = (S[3] * b[3]).astype(F32)
        t = (fma32(S[0], b[0], p1) + fma32(S[2], b[2], p3)).astype(F32)
        out = np.clip(np.rint(t), 0, 255).astype(np.uint8)
    elif path == "generic":
        t = ((S[2] * b[2]).astype(F32) + (S[3] * b[3]).astype(F32)).astype(F32)
        t = ((S[1] * b[1]).astype(F32) + t).astype(F32)
        t = ((S[0] * b[0]).astype(F32) + t).astype(F32)
        out = np.clip(np.rint(t), 0, 255).astype(np.uint8)
        nvec = (n // GENERIC_VEC_LANES) * GENERIC_VEC_LANES
        if nvec < n:
            acc = sum(rows[k] * beta[:, k].astype(np.int64)[:, None] for k in range(4))
            tail = np.clip((acc + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
            out[:, nvec:] = tail[:, nvec:]
    else:
        raise ValueError(path)
    return out.reshape(dst_h, dst_w, cn)


# --------------------------------------------------------------------------------------------
# cv::GaussianBlur on CV_8U (fixed-point path)  -- upscaler.cpp:78, adaptive_sharpening.cpp:258
# --------------------------------------------------------------------------------------------
def gaussian_kernel_f64(ksize, sigma):
    """cv::getGaussianKernel(ksize, sigma, CV_64F)."""
    small = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
             7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}
    if sigma <= 0 and ksize in small:
        return np.array(small[ksize], dtype=np.float64)
    sx = sigma if sigma > 0 else ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8
    scale2x = -0.5 / (sx * sx)
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(scale2x * x * x)
    return k / k.sum()


def gaussian_kernel_q8(ksize, sigma):
    """Q0.8 kernel used by the 8-bit fixed-point GaussianBlur (getGaussianKernelFixedPoint_ED):
    side taps outside-in with error diffusion, centre = 256 - 2*sum(sides)."""
    g = gaussian_kernel_f64(ksize, sigma)
    n = ksize // 2
    q = np.zeros(ksize, dtype=np.int64)
    err = 0.0
    for i in range(n):
        v = 256.0 * g[i] + err
        qi = int(np.rint(v))  # cvRound: half-to-even
        err = v - qi
        q[i] = q[ksize - 1 - i] = qi
    q[n] = 256 - 2 * int(q[:n].sum())
    return q


def gaussian_blur_u8(img, ksize, sigma):
    """cv::GaussianBlur(img8U, out, Size(k,k), sigma) -- BORDER_REFLECT_101, Q8.8 row pass,
    Q8.16 column pass, dst = (v + 2^15) >> 16."""
    img = np.asarray(img)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    q = gaussian_kernel_q8(ksize, sigma)
    r = ksize // 2
    p = pad_reflect101(img, r).astype(np.int64)
    h, w = img.shape[:2]
    row = sum(p[:, k:k + w] * q[k] for k in range(ksize))
    col = sum(row[k:k + h] * q[k] for k in range(ksize))
    out = ((col + (1 << 15)) >> 16).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def gaussian_blur_f32(img, ksize, sigma):
    """cv::GaussianBlur on CV_32F (adaptive_sharpening.cpp:222): separable float32, kernel =
    (float)getGaussianKernel; row pass then column pass, left-to-right accumulation."""
    k = gaussian_kernel_f64(ksize, sigma).astype(F32)
    r = ksize // 2
    p = pad_reflect101(np.asarray(img, dtype=F32), r)
    h, w = img.shape[:2]
    row = np.zeros((h + 2 * r, w), dtype=F32)
    for i in range(ksize):
        row = (row + p[:, i:i + w] * k[i]).astype(F32)
    col = np.zeros((h, w), dtype=F32)
    for i in range(ksize):
        col = (col + row[i:i + h] * k[i]).astype(F32)
    return col


# --------------------------------------------------------------------------------------------
# cv::cvtColor 8-bit  -- adaptive_sharpening.cpp:114,329-330,346
# --------------------------------------------------------------------------------------------
def bgr2gray(img):
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((9798 * r + 19235 * g + 3735 * b + (1 << 14)) >> 15).astype(np.uint8)


def bgr2ycrcb(img):
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    y = (4899 * r + 9617 * g + 1868 * b + (1 << 13)) >> 14
    cr = ((r - y) * 11682 + (128 << 14) + (1 << 13)) >> 14
    cb = ((b - y) * 9241 + (128 << 14) + (1 << 13)) >> 14
    return np.stack([y, np.clip(cr, 0, 255), np.clip(cb, 0, 255)], axis=-1).astype(np.uint8)


def ycrcb2bgr(img):
    y, cr, cb = (img[..., i].astype(np.int64) for i in range(3))
    cr = cr - 128
    cb = cb - 128
    b = y + ((cb * 29049 + (1 << 13)) >> 14)
    g = y + ((cb * -5636 + cr * -11698 + (1 << 13)) >> 14)
    r = y + ((cr * 22987 + (1 << 13)) >> 14)
    return np.clip(np.stack([b, g, r], axis=-1), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# derivatives / elementwise  -- adaptive_sharpening.cpp:159-165,192-198,265 ; main.cpp:286-290
# --------------------------------------------------------------------------------------------
def _corr3(gray, k):
    p = pad_reflect101(gray, 1).astype(np.int64)
    h, w = gray.shape
    out = np.zeros((h, w), dtype=np.int64)
    for i in range(3):
        for j in range(3):
            if k[i][j]:
                out += k[i][j] * p[i:i + h, j:j + w]
    return out


def sobel_x(gray):
    return _corr3(gray, [[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]])


def sobel_y(gray):
    return _corr3(gray, [[-1, -2, -1], [0, 0, 0], [1, 2, 1]])


def laplacian3(gray):
    """cv::Laplacian(gray, CV_16S, 3): aperture 3 kernel [[2,0,2],[0,-8,0],[2,0,2]]."""
    return _corr3(gray, [[2, 0, 2], [0, -8, 0], [2, 0, 2]])


def abs_sat_u8(x):
    """cv::convertScaleAbs on CV_16S: sat_u8(|x|)."""
    return np.clip(np.abs(x), 0, 255).astype(np.uint8)


def add_weighted_u8(a, alpha, b, beta):
    """cv::addWeighted(a8U, alpha, b8U, beta, 0): sat_u8(rint(fmaf(a, (float)alpha, b*(float)beta)))
    (the AVX2 form; for 0.5/0.5 and 0.6/0.4 every formulation agrees)."""
    fa, fb = F32(alpha), F32(beta)
    pb = (b.astype(F32) * fb).astype(F32)
    t = (a.astype(np.float64) * np.float64(fa) + pb.astype(np.float64)).astype(F32)
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


def subtract_sat_u8(a, b):
    return np.clip(a.astype(np.int16) - b.astype(np.int16), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# cv::meanStdDev  -- selective_bilateral.cpp:322
# --------------------------------------------------------------------------------------------
def mean_stddev(img):
    """Exact integer sums then double: mean = S/N, std = sqrt(max(SQ/N - mean^2, 0)) per channel."""
    x = img.reshape(-1, img.shape[-1]).astype(np.int64)
    n = x.shape[0]
    s = x.sum(axis=0)
    sq = (x * x).sum(axis=0)
    mean = s.astype(np.float64) / n
    var = np.maximum(sq.astype(np.float64) / n - mean * mean, 0.0)
    return mean, np.sqrt(var), s, sq


# --------------------------------------------------------------------------------------------
# cv::bilateralFilter CV_8UC3  -- selective_bilateral.cpp:110
# --------------------------------------------------------------------------------------------
def bilateral_tables(d, sigma_color, sigma_space):
    """radius, tap offsets (i-major) and float32 weight tables exactly as bilateralFilter_8u builds
    them (double exp, stored as float)."""
    if sigma_color <= 0:
        sigma_color = 1.0
    if sigma_space <= 0:
        sigma_space = 1.0
    gc = -0.5 / (sigma_color * sigma_color)
    gs = -0.5 / (sigma_space * sigma_space)
    radius = d // 2 if d > 0 else int(np.rint(sigma_space * 1.5))
    radius = max(radius, 1)
    color_w = np.array([math.exp(i * i * gc) for i in range(256 * 3)], dtype=np.float64).astype(F32)
    taps, space_w = [], []
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            rr = math.sqrt(float(i) * i + float(j) * j)
            if rr > radius:
                continue
            taps.append((i, j))
            space_w.append(math.exp(rr * rr * gs))
    return radius, taps, np.array(space_w, dtype=np.float64).astype(F32), color_w


def bilateral_u8c3(img, d, sigma_color, sigma_space, fma=True):
    """cv::bilateralFilter(img8UC3, out, d, sc, ss) with BORDER_REFLECT_101.
    w = space_w[k]*color_w[|db|+|dg|+|dr|]; wsum += w; sum_c = fma(p_c, w, sum_c) (the SIMD build)
    or sum_c += p_c*w (scalar build); dst_c = rint(sum_c * (1/wsum))."""
    radius, taps, space_w, color_w = bilateral_tables(d, sigma_color, sigma_space)
    h, w = img.shape[:2]
    p = pad_reflect101(img, radius)
    c0 = img.astype(np.int32)
    wsum = np.zeros((h, w), dtype=F32)
    acc = np.zeros((h, w, 3), dtype=F32)
    for (i, j), sw in zip(taps, space_w):
        q = p[radius + i:radius + i + h, radius + j:radius + j + w]
        n = np.abs(q.astype(np.int32) - c0).sum(axis=2)
        wt = (sw * color_w[n]).astype(F32)
        wsum = (wsum + wt).astype(F32)
        qf = q.astype(F32)
        if fma:
            acc = (qf.astype(np.float64) * wt[..., None].astype(np.float64) + acc.astype(np.float64)).astype(F32)
        else:
            acc = (acc + (qf * wt[..., None]).astype(F32)).astype(F32)
    inv = (F32(1.0) / wsum).astype(F32)
    return np.clip(np.rint((acc * inv[..., None]).astype(F32)), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# sigmoid LUT of createEdgeMask  -- adaptive_sharpening.cpp:202-219
# --------------------------------------------------------------------------------------------
def sigmoid_lut(edge_threshold):
    """256-entry table: edge_value = ((float)u * (float)(1/255)) * 255.0f  (convertTo(CV_32F,1/255)
    multiplies in float32, then `* 255.0f`); threshold is a double copy of the float config;
    sigmoid = (float)(1.0 / (1.0 + exp(-((double)edge_value - thr) * (double)0.1f)))."""
    thr = float(F32(edge_threshold))
    lut = np.zeros(256, dtype=F32)
    s = float(F32(0.1))
    inv255 = F32(1.0 / 255.0)
    for u in range(256):
        ev = F32(F32(F32(u) * inv255) * F32(255.0))
        lut[u] = F32(1.0 / (1.0 + math.exp(-(float(ev) - thr) * s)))
    return lut


# --------------------------------------------------------------------------------------------
# cv::Canny(gray8U, edges, low, high)  (aperture 3, L1 gradient)  -- upscaler.cpp:108
# cv::dilate(edges, mask, getStructuringElement(MORPH_RECT, Size(2,2)))  -- upscaler.cpp:112
# --------------------------------------------------------------------------------------------
CANNY_TG22 = 13573  # (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)


def canny_nms_map(gray, low, high):
    """Sobel (BORDER_REPLICATE) -> |dx|+|dy| -> non-maximum suppression exactly as cv::Canny's integer
    tangent test.  Returns 0 (no edge), 1 (weak: low < m <= high) or 2 (strong: m > high) per pixel."""
    g = gray.astype(np.int64)
    p = np.pad(g, 1, mode="edge")
    h, w = g.shape
    dx = (p[0:h, 2:w + 2] + 2 * p[1:h + 1, 2:w + 2] + p[2:h + 2, 2:w + 2]) - (p[0:h, 0:w] + 2 * p[1:h + 1, 0:w] + p[2:h + 2, 0:w])
    dy = (p[2:h + 2, 0:w] + 2 * p[2:h + 2, 1:w + 1] + p[2:h + 2, 2:w + 2]) - (p[0:h, 0:w] + 2 * p[0:h, 1:w + 1] + p[0:h, 2:w + 2])
    m = np.abs(dx) + np.abs(dy)
    mp = np.pad(m, 1, mode="constant")           # magnitude is 0 outside the image

    def nb(oy, ox):
        return mp[1 + oy:1 + oy + h, 1 + ox:1 + ox + w]
    x = np.abs(dx)
    y = np.abs(dy) << 15
    tg22x = x * CANNY_TG22
    tg67x = tg22x + (x << 16)
    horiz = y < tg22x
    vert = (~horiz) & (y > tg67x)
    diag = (~horiz) & (~vert)
    s_pos = (dx ^ dy) >= 0
    lm_h = (m > nb(0, -1)) & (m >= nb(0, 1))
    lm_v = (m > nb(-1, 0)) & (m >= nb(1, 0))
    lm_d = (m > np.where(s_pos, nb(-1, -1), nb(-1, 1))) & (m > np.where(s_pos, nb(1, 1), nb(1, -1)))
    lm = (m > low) & ((horiz & lm_h) | (vert & lm_v) | (diag & lm_d))
    return np.where(lm, np.where(m > high, 2, 1), 0).astype(np.uint8)


def canny_hysteresis(nms):
    """Weak pixels 8-connected to a strong pixel become edges: fixed point of a 3x3 dilation masked by
    the candidate set (the order cv::Canny's stack visits pixels in does not change the result)."""
    cand = nms > 0
    edge = nms == 2
    while True:
        p = np.pad(edge, 1, mode="constant")
        h, w = edge.shape
        grown = np.zeros_like(edge)
        for oy in range(3):
            for ox in range(3):
                grown |= p[oy:oy + h, ox:ox + w]
        grown &= cand
        if np.array_equal(grown, edge):
            return edge
        edge = grown


def canny_u8(gray, low, high):
    return (canny_hysteresis(canny_nms_map(gray, low, high)) * 255).astype(np.uint8)


def dilate_2x2(edges):
    """cv::dilate with a 2x2 rectangle, default anchor (1,1): max over (y-1..y, x-1..x); outside ignored."""
    e = edges > 0
    p = np.pad(e, ((1, 0), (1, 0)), mode="constant")
    h, w = e.shape
    out = p[0:h, 0:w] | p[0:h, 1:w + 1] | p[1:h + 1, 0:w] | p[1:h + 1, 1:w + 1]
    return (out * 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# TemporalConsistency::detectSceneChange primitives  -- temporal_consistency.cpp:283-323
# --------------------------------------------------------------------------------------------
def hist64(gray):
    """cv::calcHist(gray8U, 64 bins over [0,256)): bin = v >> 2, counts as float32."""
    return np.bincount((gray.reshape(-1) >> 2).astype(np.int64), minlength=64).astype(F32)


def normalize_minmax01(h):
    """cv::normalize(h, h, 0, 1, NORM_MINMAX) on CV_32F: scale/shift formed in double, applied in float32."""
    smin, smax = float(h.min()), float(h.max())
    scale = 1.0 / (smax - smin) if (smax - smin) > 2.220446049250313e-16 else 0.0
    shift = 0.0 - smin * scale
    return (h * F32(scale) + F32(shift)).astype(F32)


def hist_correlation(h1, h2):
    """cv::compareHist(HISTCMP_CORREL): double sums; 1 when a histogram is constant."""
    a, b = h1.astype(np.float64), h2.astype(np.float64)
    n = a.size
    s1, s2, s11, s22, s12 = a.sum(), b.sum(), (a * a).sum(), (b * b).sum(), (a * b).sum()
    num = s12 - s1 * s2 / n
    den = (s11 - s1 * s1 / n) * (s22 - s2 * s2 / n)
    return num / math.sqrt(den) if abs(den) > 2.220446049250313e-16 else 1.0


def scene_metrics(prev_gray, cur_gray):
    corr = hist_correlation(normalize_minmax01(hist64(prev_gray)), normalize_minmax01(hist64(cur_gray)))
    mad = float(np.abs(prev_gray.astype(np.int64) - cur_gray.astype(np.int64)).sum()) / prev_gray.size
    return corr, mad


# --------------------------------------------------------------------------------------------
# SelectiveBilateral's multiscale / selective branches
# -- selective_bilateral.cpp:138-313 (pyrDown/pyrUp :148-150,:188; blur :263,:272; Sobel/magnitude :251-256)
# --------------------------------------------------------------------------------------------
def box_blur5_u8(img):
    """cv::blur(img8U, out, Size(5,5)): normalised box filter, BORDER_REFLECT_101, dst = round(sum / 25)
    (sum / 25 never lands on .5, so every rounding mode agrees)."""
    img = np.asarray(img)
    h, w = img.shape[:2]
    p = pad_reflect101(img, 2).astype(np.int64)
    s = sum(p[dy:dy + h, dx:dx + w] for dy in range(5) for dx in range(5))
    return ((2 * s + 25) // 50).astype(np.uint8)


def box_mean7_u8(img):
    """cv::filter2D(img8U, out, -1, ones(7,7)/49.f) (adaptive_sharpening.cpp:369-373): float32 accumulation then
    saturate_cast<uchar>; sum/49 is never within 0.01 of .5, so it equals round(sum / 49) whatever the summation order."""
    img = np.asarray(img)
    h, w = img.shape[:2]
    p = pad_reflect101(img, 3).astype(np.int64)
    s = sum(p[dy:dy + h, dx:dx + w] for dy in range(7) for dx in range(7))
    return ((2 * s + 49) // 98).astype(np.uint8)


def pyr_down_u8(img):
    """cv::pyrDown on CV_8U: [1 4 6 4 1] x [1 4 6 4 1], BORDER_REFLECT_101, (v + 128) >> 8 at even (y, x);
    dst = ((W+1)/2, (H+1)/2)."""
    img = np.asarray(img)
    k = (1, 4, 6, 4, 1)
    h, w = img.shape[:2]
    p = pad_reflect101(img, 2).astype(np.int64)
    row = sum(p[:, i:i + w] * k[i] for i in range(5))
    col = sum(row[i:i + h] * k[i] for i in range(5))
    return ((col + 128) >> 8)[::2, ::2].astype(np.uint8)


def pyr_up_u8(img, dst_w, dst_h):
    """cv::pyrUp(img8U, out, Size(dst_w, dst_h)) for dst = 2*src or 2*src - 1 per axis (the sizes
    applyMultiscaleBilateral produces, selective_bilateral.cpp:188).  Even outputs: s[i-1] + 6 s[i] + s[i+1],
    odd outputs: 4 (s[i] + s[i+1]); the low neighbour is reflect-101 (s[-1] = s[1]), the high neighbour
    replicates (s[n] = s[n-1]); both axes, then (v + 32) >> 6."""
    src = np.asarray(img)
    H, W = src.shape[:2]
    assert dst_w in (2 * W, 2 * W - 1) and dst_h in (2 * H, 2 * H - 1)
    s = src.astype(np.int64)
    xm = np.arange(W) - 1
    xm[0] = 1 if W > 1 else 0
    xp = np.minimum(np.arange(W) + 1, W - 1)
    row = np.empty((H, 2 * W) + s.shape[2:], dtype=np.int64)
    row[:, 0::2] = s[:, xm] + 6 * s + s[:, xp]
    row[:, 1::2] = 4 * (s + s[:, xp])
    ym = np.arange(H) - 1
    ym[0] = 1 if H > 1 else 0
    yp = np.minimum(np.arange(H) + 1, H - 1)
    out = np.empty((2 * H, 2 * W) + s.shape[2:], dtype=np.int64)
    out[0::2] = row[ym] + 6 * row + row[yp]
    out[1::2] = 4 * (row + row[yp])
    return ((out + 32) >> 6).astype(np.uint8)[:dst_h, :dst_w]


def sobel_magnitude_f32(gray):
    """cv::Sobel(gray, CV_32F, 1,0,3) / (0,1,3) + cv::magnitude (selective_bilateral.cpp:251-256): the
    derivatives are exact integers, gx^2 + gy^2 < 2^24 is exact, so this is one correctly rounded sqrt
    (cv2's own magnitude deviates from it by at most 1 ulp)."""
    gx = sobel_x(gray).astype(np.int64)
    gy = sobel_y(gray).astype(np.int64)
    return np.sqrt((gx * gx + gy * gy).astype(F32)).astype(F32)


def detail_sigmoid_f32(detail, detail_threshold):
    """selective_bilateral.cpp:293-303: mask = 1.0f / (1.0f + std::exp(-(v - threshold) * 10.0f)) with
    `threshold` a double (detail_threshold / 255.0), so the expression is evaluated in double and rounded
    to float32 on the store."""
    thr = float(detail_threshold) / 255.0
    v = np.asarray(detail, dtype=F32).astype(np.float64)
    return (1.0 / (1.0 + np.exp(-(v - thr) * 10.0))).astype(F32)


# --------------------------------------------------------------------------------------------
# TemporalConsistency's motion-compensated front end:
# cv::calcOpticalFlowFarneback (temporal_consistency.cpp:211-216), cv::remap (:261-273).  The Farneback
# restatement follows OpenCV's published algorithm (modules/video/src/optflowgf.cpp: FarnebackPrepareGaussian,
# FarnebackPolyExp, FarnebackUpdateMatrices, FarnebackUpdateFlow_Blur and the pyramid loop of
# FarnebackOpticalFlowImpl::calc) including its float / double mix; tests pin it against cv2 to 1e-4 px.
# --------------------------------------------------------------------------------------------
def cv_round(x):
    return int(np.rint(x))

def resize_linear_f32(src, dw, dh):
    """cv::resize(INTER_LINEAR) on CV_32F (any channel count)."""
    sh, sw = src.shape[:2]
    def tabs(sn, dn):
        scale = float(sn) / float(dn)
        ofs = np.zeros(dn, np.int64); a = np.zeros(dn, F32)
        for d in range(dn):
            fx = F32((d + 0.5) * scale - 0.5)
            sx = int(math.floor(fx)); fx = F32(fx - F32(sx))
            if sx < 0: fx = F32(0); sx = 0
            if sx >= sn - 1: fx = F32(0); sx = sn - 1
            ofs[d] = sx; a[d] = fx
        return ofs, a
    xo, xa = tabs(sw, dw); yo, ya = tabs(sh, dh)
    s = src.astype(F32)
    x1 = np.minimum(xo + 1, sw - 1)
    a1 = xa.reshape((1, dw) + (1,) * (s.ndim - 2)); a0 = (F32(1) - a1).astype(F32)
    h = ((s[:, xo] * a0).astype(F32) + (s[:, x1] * a1).astype(F32)).astype(F32)
    y1 = np.minimum(yo + 1, sh - 1)
    b1 = ya.reshape((dh, 1) + (1,) * (s.ndim - 2)); b0 = (F32(1) - b1).astype(F32)
    return ((h[yo] * b0).astype(F32) + (h[y1] * b1).astype(F32)).astype(F32)

def prepare_gaussian(n, sigma):
    if sigma < np.finfo(np.float32).eps: sigma = n * 0.3
    xs = np.arange(-n, n + 1)
    g = np.exp(-xs * xs / (2 * sigma * sigma)).astype(F32)
    s = 1.0 / float(g.astype(np.float64).sum())
    g = (g.astype(np.float64) * s).astype(F32)
    xg = (xs.astype(F32) * g).astype(F32)
    xxg = ((xs * xs).astype(F32) * g).astype(F32)
    G = np.zeros((6, 6))
    for y in range(-n, n + 1):
        for x in range(-n, n + 1):
            w = float(F32(g[y + n] * g[x + n]))     # float*float in float
            G[0, 0] += w; G[1, 1] += w * x * x; G[3, 3] += w * x * x * x * x; G[5, 5] += w * x * x * y * y
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]; G[3, 4] = G[4, 3] = G[5, 5]
    inv = np.linalg.inv(G)
    return g, xg, xxg, inv[1, 1], inv[0, 3], inv[3, 3], inv[5, 5]

def poly_exp(src, n, sigma):
    h, w = src.shape
    g, xg, xxg, ig11, ig03, ig33, ig55 = prepare_gaussian(n, sigma)
    c = n  # centre index
    s = src.astype(F32)
    ys = np.arange(h)
    row0 = (s * g[c]).astype(F32); row1 = np.zeros_like(s); row2 = np.zeros_like(s)
    for k in range(1, n + 1):
        a = s[np.maximum(ys - k, 0)]; b = s[np.minimum(ys + k, h - 1)]
        p = (a + b).astype(F32)
        row0 = (row0 + (g[c + k] * p).astype(F32)).astype(F32)
        row1 = (row1 + (xg[c + k] * (b - a).astype(F32)).astype(F32)).astype(F32)
        row2 = (row2 + (xxg[c + k] * p).astype(F32)).astype(F32)
    xs = np.arange(w)
    def at(r, k):
        return r[:, np.clip(xs + k, 0, w - 1)]
    b1 = (row0 * g[c]).astype(F32).astype(np.float64); b3 = (row1 * g[c]).astype(F32).astype(np.float64); b5 = (row2 * g[c]).astype(F32).astype(np.float64)
    b2 = np.zeros((h, w)); b4 = np.zeros((h, w)); b6 = np.zeros((h, w))
    for k in range(1, n + 1):
        tg = (at(row0, k) + at(row0, -k)).astype(F32).astype(np.float64)
        g0 = float(g[c + k])
        b1 += tg * g0; b4 += tg * float(xxg[c + k])
        b2 += ((at(row0, k) - at(row0, -k)).astype(F32) * xg[c + k]).astype(F32).astype(np.float64)
        b3 += ((at(row1, k) + at(row1, -k)).astype(F32) * g[c + k]).astype(F32).astype(np.float64)
        b6 += ((at(row1, k) - at(row1, -k)).astype(F32) * xg[c + k]).astype(F32).astype(np.float64)
        b5 += ((at(row2, k) + at(row2, -k)).astype(F32) * g[c + k]).astype(F32).astype(np.float64)
    out = np.empty((h, w, 5), F32)
    out[..., 1] = (b2 * ig11).astype(F32); out[..., 0] = (b3 * ig11).astype(F32)
    out[..., 3] = (b1 * ig03 + b4 * ig33).astype(F32); out[..., 2] = (b1 * ig03 + b5 * ig33).astype(F32)
    out[..., 4] = (b6 * ig55).astype(F32)
    return out

BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], F32)

def update_matrices(R0, R1, flow):
    h, w = flow.shape[:2]
    xs = np.arange(w, dtype=F32)[None, :]; ys = np.arange(h, dtype=F32)[:, None]
    dx = flow[..., 0]; dy = flow[..., 1]
    fx = (xs + dx).astype(F32); fy = (ys + dy).astype(F32)
    x1 = np.floor(fx).astype(np.int64); y1 = np.floor(fy).astype(np.int64)
    fx = (fx - x1.astype(F32)).astype(F32); fy = (fy - y1.astype(F32)).astype(F32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    xc = np.clip(x1, 0, w - 2); yc = np.clip(y1, 0, h - 2)
    a00 = ((F32(1) - fx) * (F32(1) - fy)).astype(F32); a01 = (fx * (F32(1) - fy)).astype(F32)
    a10 = ((F32(1) - fx) * fy).astype(F32); a11 = (fx * fy).astype(F32)
    def interp(c):
        p = R1[..., c]
        v = (a00 * p[yc, xc]).astype(F32)
        v = (v + (a01 * p[yc, xc + 1]).astype(F32)).astype(F32)
        v = (v + (a10 * p[yc + 1, xc]).astype(F32)).astype(F32)
        v = (v + (a11 * p[yc + 1, xc + 1]).astype(F32)).astype(F32)
        return v
    r2 = interp(0); r3 = interp(1); r4 = interp(2); r5 = interp(3); r6 = interp(4)
    r4 = ((R0[..., 2] + r4).astype(F32) * F32(0.5)).astype(F32)
    r5 = ((R0[..., 3] + r5).astype(F32) * F32(0.5)).astype(F32)
    r6 = ((R0[..., 4] + r6).astype(F32) * F32(0.25)).astype(F32)
    r2 = np.where(inside, r2, F32(0)); r3 = np.where(inside, r3, F32(0))
    r4 = np.where(inside, r4, R0[..., 2]); r5 = np.where(inside, r5, R0[..., 3])
    r6 = np.where(inside, r6, (R0[..., 4] * F32(0.5)).astype(F32))
    r2 = ((R0[..., 0] - r2).astype(F32) * F32(0.5)).astype(F32)
    r3 = ((R0[..., 1] - r3).astype(F32) * F32(0.5)).astype(F32)
    r2 = (r2 + ((r4 * dy).astype(F32) + (r6 * dx).astype(F32)).astype(F32)).astype(F32)
    r3 = (r3 + ((r6 * dy).astype(F32) + (r5 * dx).astype(F32)).astype(F32)).astype(F32)
    # :BORDER scale factors of the five outermost rows / columns
    sx = np.ones(w, F32); sy = np.ones(h, F32)
    for k in range(5):
        if k < w: sx[k] = BORDER[k]
    sxr = np.ones(w, F32)
    for k in range(5):
        if w - 1 - k >= 0: sxr[w - 1 - k] = BORDER[k]
    for k in range(5):
        if k < h: sy[k] = BORDER[k]
    syr = np.ones(h, F32)
    for k in range(5):
        if h - 1 - k >= 0: syr[h - 1 - k] = BORDER[k]
    scale = (((sx * sxr).astype(F32)[None, :] * sy[:, None]).astype(F32) * syr[:, None]).astype(F32)
    r2 = (r2 * scale).astype(F32); r3 = (r3 * scale).astype(F32); r4 = (r4 * scale).astype(F32); r5 = (r5 * scale).astype(F32); r6 = (r6 * scale).astype(F32)
    M = np.empty((h, w, 5), F32)
    M[..., 0] = ((r4 * r4).astype(F32) + (r6 * r6).astype(F32)).astype(F32)
    M[..., 1] = ((r4 + r5).astype(F32) * r6).astype(F32)
    M[..., 2] = ((r5 * r5).astype(F32) + (r6 * r6).astype(F32)).astype(F32)
    M[..., 3] = ((r4 * r2).astype(F32) + (r6 * r3).astype(F32)).astype(F32)
    M[..., 4] = ((r6 * r2).astype(F32) + (r5 * r3).astype(F32)).astype(F32)
    return M

def update_flow_blur(M, block):
    h, w = M.shape[:2]
    m = block // 2
    ys = np.arange(h); xs = np.arange(w)
    Md = M.astype(np.float64)
    v = np.zeros((h, w, 5))
    for k in range(-m, m + 1):
        v += Md[np.clip(ys + k, 0, h - 1)]
    s = np.zeros((h, w, 5))
    for k in range(-m, m + 1):
        s += v[:, np.clip(xs + k, 0, w - 1)]
    s *= 1.0 / (block * block)
    g11, g12, g22, h1, h2 = [s[..., i] for i in range(5)]
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    flow = np.empty((h, w, 2), F32)
    flow[..., 0] = ((g11 * h2 - g12 * h1) * idet).astype(F32)
    flow[..., 1] = ((g22 * h1 - g12 * h2) * idet).astype(F32)
    return flow

def farneback(prev, nxt, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2):
    h0, w0 = prev.shape
    min_size = 32
    k = 0; scale = 1.0
    while k < levels:
        scale *= pyr_scale
        if w0 * scale < min_size or h0 * scale < min_size: break
        k += 1
    levels = k
    prev_flow = None
    for k in range(levels, -1, -1):
        scale = 1.0
        for i in range(k): scale *= pyr_scale
        sigma = (1.0 / scale - 1) * 0.5
        smooth_sz = cv_round(sigma * 5) | 1
        smooth_sz = max(smooth_sz, 3)
        width = cv_round(w0 * scale); height = cv_round(h0 * scale)
        if prev_flow is None:
            flow = np.zeros((height, width, 2), F32)
        else:
            flow = (resize_linear_f32(prev_flow, width, height) * F32(1.0 / pyr_scale)).astype(F32)
        Rs = []
        for img in (prev, nxt):
            f = img.astype(F32)
            f = gaussian_blur_f32(f, smooth_sz, sigma)
            I = resize_linear_f32(f, width, height)
            Rs.append(poly_exp(I, poly_n, poly_sigma))
        M = update_matrices(Rs[0], Rs[1], flow)
        for i in range(iterations):
            flow = update_flow_blur(M, winsize)
            if i < iterations - 1:
                M = update_matrices(Rs[0], Rs[1], flow)
        prev_flow = flow
    return flow


def remap_linear_replicate_u8(img, map_x, map_y):
    """cv::remap(img8U, out, map_x, map_y, INTER_LINEAR, BORDER_REPLICATE) (temporal_consistency.cpp:273): the maps are
    rounded to 1/32 pixel (cvRound(v * 32)), the four taps are clamped into the image, integer weights
    32 * (32-fx) * (32-fy) ... summing to 2^15, dst = (sum + 2^14) >> 15.  Bit-exact against cv2."""
    h, w = img.shape[:2]
    sx = np.rint((np.asarray(map_x, F32) * F32(32)).astype(F32)).astype(np.int64)
    sy = np.rint((np.asarray(map_y, F32) * F32(32)).astype(F32)).astype(np.int64)
    fx, fy = sx & 31, sy & 31
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)
    x0, x1 = np.clip(ix, 0, w - 1), np.clip(ix + 1, 0, w - 1)
    y0, y1 = np.clip(iy, 0, h - 1), np.clip(iy + 1, 0, h - 1)
    w00 = (32 * (32 - fx) * (32 - fy))[..., None]
    w01 = (32 * fx * (32 - fy))[..., None]
    w10 = (32 * (32 - fx) * fy)[..., None]
    w11 = (32 * fx * fy)[..., None]
    p = img.astype(np.int64)
    v = p[y0, x0] * w00 + p[y0, x1] * w01 + p[y1, x0] * w10 + p[y1, x1] * w11
    return np.clip((v + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def flow_to_maps(flow):
    """temporal_consistency.cpp:262-270: map_x = x + f[0], map_y = y + f[1] in float32."""
    h, w = flow.shape[:2]
    xs = np.arange(w, dtype=F32)[None, :]
    ys = np.arange(h, dtype=F32)[:, None]
    return (xs + flow[..., 0]).astype(F32), (ys + flow[..., 1]).astype(F32)


def flow_reliability_raw(flow, motion_threshold):
    """calculateFlowReliabilityMask before its blur, temporal_consistency.cpp:325-345: float32 magnitude (no FMA),
    1 where |f| <= threshold, else clamp(expf(-(|f| - thr) / 10.f), 0, 1)."""
    fx, fy = flow[..., 0].astype(F32), flow[..., 1].astype(F32)
    mag = np.sqrt(((fx * fx).astype(F32) + (fy * fy).astype(F32)).astype(F32)).astype(F32)
    thr = F32(motion_threshold)
    rel = np.exp((-(mag - thr).astype(F32) / F32(10.0)).astype(F32).astype(np.float64)).astype(F32)
    rel = np.clip(rel, F32(0), F32(1))
    return np.where(mag > thr, rel, F32(1.0)).astype(F32)
