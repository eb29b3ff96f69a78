"""Pins oracle/restate.py (the numpy restatement) against the REAL OpenCV kernels through cv2, and
against the committed golden vectors.  CPU only."""
import os

import numpy as np
import pytest

from oracle import ref_chain as RC
from oracle import restate as R
from oracle import synth

cv2 = pytest.importorskip("cv2")
cv2.setUseOptimized(True)
GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_small.npz")


@pytest.mark.parametrize("kind", ["noise", "scene", "flat"])
@pytest.mark.parametrize("scale", [2, 4])
@pytest.mark.parametrize("size", [(131, 67), (64, 48), (17, 9)])
def test_resize_cubic_bit_exact(kind, scale, size):
    w, h = size
    img = synth.frame(kind, w, h, 2)
    ref = cv2.resize(img, (w * scale, h * scale), interpolation=cv2.INTER_CUBIC)
    assert np.array_equal(R.resize_cubic(img, w * scale, h * scale), ref)


def test_resize_cubic_taps():
    _, c2 = R.cubic_coeffs(100, 200)
    assert c2[1].tolist() == [-216, 1800, 536, -72] and c2[2].tolist() == [-72, 536, 1800, -216]
    _, c4 = R.cubic_coeffs(100, 400)
    assert sorted(map(tuple, c4[4:8].tolist())) == sorted([(-135, 873, 1535, -225), (-21, 235, 1981, -147),
                                                           (-147, 1981, 235, -21), (-225, 1535, 873, -135)])
    assert (c4.sum(axis=1) == 2048).all()


@pytest.mark.parametrize("k,sigma,q", [(3, 0.5, [27, 202, 27]), (5, 1.5, [31, 60, 74, 60, 31]), (5, 1.0, [14, 62, 104, 62, 14]),
                                       (5, 0.8, [6, 58, 128, 58, 6]), (5, 2.5, [43, 55, 60, 55, 43]), (3, 0, [64, 128, 64])])
def test_gaussian_q8_kernels(k, sigma, q):
    assert R.gaussian_kernel_q8(k, sigma).tolist() == q


@pytest.mark.parametrize("k,sigma", [(3, 0.5), (5, 1.5), (7, 2.5), (3, 0.8)])
def test_gaussian_u8_bit_exact(k, sigma):
    img = synth.frame("noise", 97, 53, 1)
    assert np.array_equal(R.gaussian_blur_u8(img, k, sigma), cv2.GaussianBlur(img, (k, k), sigma))


def test_color_conversions_bit_exact():
    img = synth.frame("noise", 211, 97, 3)
    assert np.array_equal(R.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    ycc = cv2.cvtColor(img, cv2.COLOR_BGR2YCrCb)
    assert np.array_equal(R.bgr2ycrcb(img), ycc)
    assert np.array_equal(R.ycrcb2bgr(ycc), cv2.cvtColor(ycc, cv2.COLOR_YCrCb2BGR))
    assert np.array_equal(R.ycrcb2bgr(img), cv2.cvtColor(img, cv2.COLOR_YCrCb2BGR))   # arbitrary triples


def test_derivatives_bit_exact():
    g = synth.frame("noise", 120, 77, 4)[..., 0].copy()
    assert np.array_equal(R.abs_sat_u8(R.sobel_x(g)), cv2.convertScaleAbs(cv2.Sobel(g, cv2.CV_16S, 1, 0, ksize=3)))
    assert np.array_equal(R.abs_sat_u8(R.sobel_y(g)), cv2.convertScaleAbs(cv2.Sobel(g, cv2.CV_16S, 0, 1, ksize=3)))
    assert np.array_equal(R.abs_sat_u8(R.laplacian3(g)), cv2.convertScaleAbs(cv2.Laplacian(g, cv2.CV_16S, ksize=3)))


@pytest.mark.parametrize("alpha,beta", [(0.6, 0.4), (0.5, 0.5), (0.7, 0.3)])
def test_add_weighted_all_pairs(alpha, beta):
    a = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    b = np.tile(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    assert np.array_equal(R.add_weighted_u8(a, alpha, b, beta), cv2.addWeighted(a, alpha, b, beta, 0))


def test_edge_combination_integer_form():
    """The kernel evaluates RNE(fma(e1, 0.6f, lap*0.4f)) as (6*e1 + 4*lap + 5)//10 and
    RNE(fma(ax,.5,ay*.5)) as (s + ((s>>1)&1))>>1: exhaustive over all 65,536 pairs."""
    a = np.repeat(np.arange(256, dtype=np.int64), 256)
    b = np.tile(np.arange(256, dtype=np.int64), 256)
    f1 = R.add_weighted_u8(a.astype(np.uint8), 0.6, b.astype(np.uint8), 0.4).astype(np.int64)
    assert np.array_equal(f1, ((6 * a + 4 * b + 5) * 6554) >> 16)
    assert np.array_equal(f1, (6 * a + 4 * b + 5) // 10)
    f2 = R.add_weighted_u8(a.astype(np.uint8), 0.5, b.astype(np.uint8), 0.5).astype(np.int64)
    s = a + b
    assert np.array_equal(f2, (s + ((s >> 1) & 1)) >> 1)


@pytest.mark.parametrize("kind", ["scene", "noise", "flat"])
def test_canny_and_dilate_bit_exact(kind):
    g = cv2.cvtColor(synth.frame(kind, 257, 131, 3), cv2.COLOR_BGR2GRAY)
    e = cv2.Canny(g, 50, 150)
    assert np.array_equal(R.canny_u8(g, 50, 150), e)
    assert np.array_equal(R.dilate_2x2(e), cv2.dilate(e, cv2.getStructuringElement(cv2.MORPH_RECT, (2, 2))))


@pytest.mark.parametrize("kind", ["scene", "noise"])
def test_cpuimpl_bicubic_wrapper_numpy_vs_cv2(kind):
    img = synth.frame(kind, 120, 68, 1)
    assert np.array_equal(RC.cpuimpl_bicubic(RC.NumpyOps(), img, 240, 136), RC.cpuimpl_bicubic(RC.Cv2Ops(), img, 240, 136))


@pytest.mark.parametrize("a,b", [(("scene", 0), ("scene", 1)), (("scene", 15), ("scene", 16)), (("flat", 0), ("scene", 3)),
                                 (("scene", 0), ("noise", 0)), (("flat", 1), ("flat", 2))])
def test_scene_metrics_match_cv2(a, b):
    ga = cv2.cvtColor(synth.frame(a[0], 320, 180, a[1]), cv2.COLOR_BGR2GRAY)
    gb = cv2.cvtColor(synth.frame(b[0], 320, 180, b[1]), cv2.COLOR_BGR2GRAY)
    c1, m1 = R.scene_metrics(ga, gb)
    c2, m2 = RC.Cv2Ops().scene_metrics(ga, gb)
    assert abs(c1 - c2) < 1e-7 and m1 == m2
    assert np.array_equal(R.normalize_minmax01(R.hist64(ga)),
                          cv2.normalize(cv2.calcHist([ga], [0], None, [64], [0, 256]), None, 0, 1, cv2.NORM_MINMAX).ravel())


def test_scene_change_resets_blend_state():
    cfg = RC.TemporalConfig(scene_detect=True)
    st = RC.TemporalBlendState(cfg, RC.NumpyOps())
    f = [synth.frame("flat", 64, 48, 0), synth.frame("flat", 64, 48, 1), synth.frame("scene", 64, 48, 3), synth.frame("scene", 64, 48, 4)]
    outs = [st.process(x) for x in f]
    assert np.array_equal(outs[2], f[2])                       # cut flat -> scene: passthrough, history cleared
    assert np.array_equal(outs[3], RC.blend_frames([f[3], f[2]], cfg.blend_strength))


def test_mean_stddev_matches_cv2():
    img = synth.frame("scene", 160, 90, 0)
    m, s, _, _ = R.mean_stddev(img)
    mc, sc = cv2.meanStdDev(img)
    assert np.allclose(m, mc.ravel(), rtol=0, atol=1e-9) and np.allclose(s, sc.ravel(), rtol=0, atol=1e-9)


@pytest.mark.parametrize("d,sc,ss", [(5, 24, 24), (11, 45, 45), (3, 16, 16)])
def test_bilateral_within_1lsb_of_cv2(d, sc, ss):
    img = synth.frame("scene", 120, 70, 1)
    diff = np.abs(R.bilateral_u8c3(img, d, sc, ss).astype(int) - cv2.bilateralFilter(img, d, sc, ss).astype(int))
    assert diff.max() <= 1 and (diff > 0).sum() <= img.size // 1000


def test_sigmoid_lut_known_values():
    lut = R.sigmoid_lut(30.0)
    for u, v in [(0, 0.047425870), (15, 0.18242553), (30, 0.50000006), (45, 0.81757450), (60, 0.95257413), (255, 1.0)]:
        assert abs(float(lut[u]) - v) < 5e-8


@pytest.mark.parametrize("kind", ["noise", "scene"])
def test_chain_numpy_vs_cv2_backend(kind):
    """The whole oracle chain over the numpy restatement vs over the real cv2 kernels."""
    sw, sh = 96, 64
    cfg = RC.ChainConfig(dst_w=2 * sw, dst_h=2 * sh, temporal=RC.TEMPORAL_BLEND_FRAMES)
    a, b = RC.Chain(RC.NumpyOps(), cfg), RC.Chain(RC.Cv2Ops(), cfg)
    for f in synth.clip(kind, sw, sh, 3):
        x, y = a.process(f), b.process(f)
        assert RC.psnr(x, y) >= 60.0 and np.abs(x.astype(int) - y.astype(int)).max() <= 4


def test_golden_vectors():
    """Frozen outputs of the oracle (tests/golden/make_golden.py): guards the oracle itself against drift."""
    g = np.load(GOLD)
    img = synth.frame("scene", 48, 32, 5)
    assert np.array_equal(g["scene_48x32_t5"], img)
    assert np.array_equal(g["bicubic_x2"], R.resize_cubic(img, 96, 64))
    assert np.array_equal(g["bicubic_x4"], R.resize_cubic(img, 192, 128))
    assert np.array_equal(g["bilateral_d5"], R.bilateral_u8c3(img, 5, 24.0, 24.0))
    scfg = RC.SharpenConfig(adaptive_sigma=False)
    assert np.allclose(g["edge_mask"], RC.edge_mask(RC.NumpyOps(), img, scfg), rtol=0, atol=1e-7)
    assert np.array_equal(g["sharpen"], RC.sharpen(RC.NumpyOps(), img, scfg))
    cfg = RC.ChainConfig(dst_w=96, dst_h=64, temporal=RC.TEMPORAL_BLEND_FRAMES)
    ch = RC.Chain(RC.NumpyOps(), cfg)
    for t in range(3):
        assert np.array_equal(g[f"chain_t{t}"], ch.process(synth.frame("scene", 48, 32, t)))
    # the same vectors out of the real OpenCV kernels
    assert np.array_equal(g["bicubic_x2"], cv2.resize(img, (96, 64), interpolation=cv2.INTER_CUBIC))
    assert np.array_equal(g["bicubic_x4"], cv2.resize(img, (192, 128), interpolation=cv2.INTER_CUBIC))


# ---- SelectiveBilateral's selective / multi-scale branches --------------------------------
@pytest.mark.parametrize("size", [(37, 53), (64, 64), (5, 7), (3, 3), (1, 9), (120, 200)])
def test_box_blur5_bit_exact(size):
    g = np.random.default_rng(1).integers(0, 256, size, dtype=np.uint8)
    assert np.array_equal(R.box_blur5_u8(g), cv2.blur(g, (5, 5)))


@pytest.mark.parametrize("size", [(37, 53), (64, 64), (5, 7), (3, 3), (120, 200), (135, 240), (2, 2)])
def test_pyr_down_up_bit_exact(size):
    h, w = size
    img = np.random.default_rng(2).integers(0, 256, (h, w, 3), dtype=np.uint8)
    down = R.pyr_down_u8(img)
    assert np.array_equal(down, cv2.pyrDown(img))
    assert np.array_equal(R.pyr_up_u8(down, w, h), cv2.pyrUp(down, dstsize=(w, h)))


def test_sobel_magnitude_within_1ulp():
    g = np.random.default_rng(3).integers(0, 256, (97, 131), dtype=np.uint8)
    ref = cv2.magnitude(cv2.Sobel(g, cv2.CV_32F, 1, 0, ksize=3), cv2.Sobel(g, cv2.CV_32F, 0, 1, ksize=3))
    mine = R.sobel_magnitude_f32(g)
    assert np.abs(mine.view(np.int32) - ref.view(np.int32)).max() <= 1


@pytest.mark.parametrize("kind", ["noise", "scene", "flat"])
def test_detail_mask_numpy_vs_cv2(kind):
    img = synth.frame(kind, 160, 96, 1)
    cfg = RC.BilateralConfig()
    a = RC.create_detail_mask(RC.NumpyOps(), img, cfg)
    b = RC.create_detail_mask(RC.Cv2Ops(), img, cfg)
    assert a.dtype == np.float32 and np.abs(a - b).max() <= 2e-6 and 0.0 <= a.min() and a.max() <= 1.0


@pytest.mark.parametrize("mode", [dict(use_multiscale=False, selective=True), dict(use_multiscale=True, num_scales=3)])
def test_selective_modes_numpy_vs_cv2(mode):
    img = synth.frame("scene", 160, 96, 1)
    cfg = RC.BilateralConfig(**mode)
    a = RC.selective_bilateral_process(RC.NumpyOps(), img, cfg)
    b = RC.selective_bilateral_process(RC.Cv2Ops(), img, cfg)
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() <= 2 and (d > 0).mean() < 0.02


def test_detail_mask_of_a_constant_frame_is_finite():
    img = np.full((20, 30, 3), 77, np.uint8)
    m = RC.create_detail_mask(RC.NumpyOps(), img, RC.BilateralConfig())
    assert np.isfinite(m).all()


# ---- AdaptiveSharpening, adaptive_sigma = true ----------------------------------------------
@pytest.mark.parametrize("size", [(37, 53), (64, 64), (7, 9), (3, 3), (120, 200)])
def test_box_mean7_bit_exact(size):
    g = np.random.default_rng(5).integers(0, 256, size, dtype=np.uint8)
    assert np.array_equal(R.box_mean7_u8(g), cv2.filter2D(g, -1, np.ones((7, 7), np.float32) / np.float32(49.0)))


@pytest.mark.parametrize("kind", ["noise", "scene", "flat"])
def test_variable_sigma_numpy_vs_cv2(kind):
    img = synth.frame(kind, 160, 96, 1)
    cfg = RC.SharpenConfig()
    ta, tb = RC.texture_map(RC.NumpyOps(), img), RC.texture_map(RC.Cv2Ops(), img)
    assert np.array_equal(ta, tb)
    sa, sb = RC.adaptive_sigma_map(RC.NumpyOps(), ta), RC.adaptive_sigma_map(RC.Cv2Ops(), tb)
    assert np.abs(sa - sb).max() <= 2e-6 and 0.8 - 1e-5 <= sa.min() and sa.max() <= 2.5 + 1e-5
    a = RC.sharpen_variable_sigma(RC.NumpyOps(), img, cfg)
    b = RC.sharpen_variable_sigma(RC.Cv2Ops(), img, cfg)
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() <= 2 and (d > 0).mean() < 0.01


def test_variable_sigma_constant_frame_is_passthrough_safe():
    img = np.full((20, 30, 3), 77, np.uint8)
    out = RC.sharpen_variable_sigma(RC.NumpyOps(), img, RC.SharpenConfig())
    assert out.shape == img.shape and np.abs(out.astype(int) - 77).max() <= 1


def test_golden_vectors_widened_rows():
    """tests/golden/golden_widened.npz holds output of the REAL OpenCV kernels (cv2 backend) for the SURVEY 8(f) rows;
    the numpy restatement must reproduce it: integer stages exactly, float stages / truncating stages within the bars
    the GPU tests use."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_widened.npz"))
    n = RC.NumpyOps()
    img, nxt = synth.frame("scene", 96, 64, 5), synth.frame("scene", 96, 64, 6)
    assert np.array_equal(g["pyr_down"], R.pyr_down_u8(img))
    assert np.array_equal(g["pyr_up"], R.pyr_up_u8(g["pyr_down"], 96, 64))
    assert np.abs(g["detail_mask"] - RC.create_detail_mask(n, img, RC.BilateralConfig())).max() <= 2e-6

    def lsb(a, b, mx, frac):
        d = np.abs(a.astype(int) - b.astype(int))
        return d.max() <= mx and (d > 0).mean() <= frac
    assert lsb(g["selective"], RC.selective_bilateral_process(n, img, RC.BilateralConfig(selective=True, use_multiscale=False)), 1, 0.01)
    assert lsb(g["multiscale"], RC.selective_bilateral_process(n, img, RC.BilateralConfig(use_multiscale=True, num_scales=3)), 2, 0.02)
    assert np.abs(g["sigma_map"] - RC.adaptive_sigma_map(n, RC.texture_map(n, img))).max() <= 2e-6
    assert lsb(g["variable_sigma_sharpen"], RC.sharpen_variable_sigma(n, img, RC.SharpenConfig(adaptive_sigma=True)), 2, 0.01)
    t = RC.TemporalConfig(use_flow=True, scene_detect=True)
    flow = n.farneback(R.bgr2gray(img), R.bgr2gray(nxt), t)
    assert np.abs(g["farneback_flow"] - flow).max() <= 1e-4
    assert np.array_equal(g["warp"], n.remap_flow(img, g["farneback_flow"]))
    big = (g["farneback_flow"] * np.float32(12.0)).astype(np.float32)
    assert np.abs(g["reliability_mask"] - RC.flow_reliability_mask(n, big, t.motion_threshold)).max() <= 2e-6
    tc = RC.TemporalConsistencyFull(t, n)
    for k in range(5):
        assert lsb(g[f"tc_full_t{k}"], tc.process(synth.frame("scene", 96, 64, k)), 1, 0.01)
