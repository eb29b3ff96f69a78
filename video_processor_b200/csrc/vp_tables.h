// vp_tables.h - host-side construction of the coefficient tables the kernels consume, evaluated the
// way OpenCV evaluates them (oracle/restate.py restates and pins the same formulas):
//   cv::resize cubic taps       (imgproc resize.cpp: interpolateCubic + the xofs/ialpha loop)
//   cv::getGaussianKernel, its 8-bit fixed-point form (smooth.dispatch.cpp, error diffusion)
//   cv::bilateralFilter weight tables (bilateral_filter.dispatch.cpp)
//   the sigmoid of AdaptiveSharpening::createEdgeMask (src/adaptive_sharpening.cpp:202-219)
// Compile with -ffp-contract=off: the float32 expressions must not be fused or widened.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace vp_host {

struct CubicTable {
    std::vector<int> ofs;          // first tap index (s-1), unclamped
    std::vector<short> coef;       // 4 per destination index, sums to 2048
};

inline CubicTable cubic_table(int src_len, int dst_len) {
    CubicTable t;
    t.ofs.resize(dst_len);
    t.coef.resize((size_t)dst_len * 4);
    const double inv_scale = (double)dst_len / (double)src_len;
    const double scale = 1.0 / inv_scale;
    const float A = -0.75f;
    for (int d = 0; d < dst_len; ++d) {
        volatile float f = (float)((d + 0.5) * scale - 0.5);
        const int s = (int)std::floor((float)f);
        volatile float x = (float)f - (float)s;
        volatile float xp1 = x + 1.0f, omx = 1.0f - x;
        volatile float c0 = ((A * xp1 - 5.0f * A) * xp1 + 8.0f * A) * xp1 - 4.0f * A;
        volatile float c1 = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
        volatile float c2 = ((A + 2.0f) * omx - (A + 3.0f)) * omx * omx + 1.0f;
        volatile float c3 = 1.0f - c0 - c1 - c2;
        const float c[4] = {c0, c1, c2, c3};
        t.ofs[d] = s - 1;
        for (int k = 0; k < 4; ++k) {
            volatile float v = c[k] * 2048.0f;
            long r = std::lrintf(v);   // round-half-even (default rounding mode) == cvRound
            if (r > 32767) r = 32767;
            if (r < -32768) r = -32768;
            t.coef[(size_t)d * 4 + k] = (short)r;
        }
    }
    return t;
}

inline std::vector<double> gaussian_kernel_f64(int ksize, double sigma) {
    static const double small[4][7] = {{1.0}, {0.25, 0.5, 0.25}, {0.0625, 0.25, 0.375, 0.25, 0.0625},
                                       {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125}};
    std::vector<double> k(ksize);
    if (sigma <= 0 && (ksize & 1) && ksize <= 7) {
        for (int i = 0; i < ksize; ++i) k[i] = small[ksize >> 1][i];
        return k;
    }
    const double sx = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double scale2x = -0.5 / (sx * sx);
    double sum = 0;
    for (int i = 0; i < ksize; ++i) {
        const double x = i - (ksize - 1) * 0.5;
        k[i] = std::exp(scale2x * x * x);
        sum += k[i];
    }
    for (int i = 0; i < ksize; ++i) k[i] /= sum;
    return k;
}

// numpy sums pairwise; for <= 7 terms that is plain left-to-right, same as the loop above.
inline void gaussian_kernel_q8(int ksize, double sigma, int q[7]) {
    const std::vector<double> g = gaussian_kernel_f64(ksize, sigma);
    const int n = ksize / 2;
    double err = 0.0;
    int side = 0;
    for (int i = 0; i < 7; ++i) q[i] = 0;
    for (int i = 0; i < n; ++i) {
        const double v = 256.0 * g[i] + err;
        const int qi = (int)std::lrint(v);
        err = v - qi;
        q[i] = q[ksize - 1 - i] = qi;
        side += qi;
    }
    q[n] = 256 - 2 * side;
}

struct BilateralTables {
    int radius, d;
    std::vector<float> space_w;    // dense (2r+1)^2, zero outside the circle
    std::vector<float> color_w;    // 768
};

inline BilateralTables bilateral_tables(int d, double sigma_color, double sigma_space) {
    BilateralTables t;
    if (sigma_color <= 0) sigma_color = 1.0;
    if (sigma_space <= 0) sigma_space = 1.0;
    const double gc = -0.5 / (sigma_color * sigma_color);
    const double gs = -0.5 / (sigma_space * sigma_space);
    int radius = d > 0 ? d / 2 : (int)std::lrint(sigma_space * 1.5);
    if (radius < 1) radius = 1;
    t.radius = radius;
    t.d = radius * 2 + 1;
    t.color_w.resize(768);
    for (int i = 0; i < 768; ++i) t.color_w[i] = (float)std::exp((double)i * i * gc);
    const int D = 2 * radius + 1;
    t.space_w.assign((size_t)D * D, 0.f);
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            const double rr = std::sqrt((double)i * i + (double)j * j);
            if (rr > radius) continue;
            t.space_w[(size_t)(i + radius) * D + (j + radius)] = (float)std::exp(rr * rr * gs);
        }
    return t;
}

inline void sigmoid_lut(float edge_threshold, float lut[256]) {
    const double thr = (double)edge_threshold;
    const double s = (double)0.1f;
    const float inv255 = (float)(1.0 / 255.0);
    for (int u = 0; u < 256; ++u) {
        volatile float a = (float)u * inv255;
        volatile float ev = a * 255.0f;
        lut[u] = (float)(1.0 / (1.0 + std::exp(-((double)ev - thr) * s)));
    }
}

// SelectiveBilateral::calculateAdaptiveParams for a given noise class (0: 0.7, 1: 1.0, 2: 1.5)
inline void adaptive_params_for_class(int cls, int stage_post, int cfg_d, double cfg_sc, double cfg_ss,
                                      int* d, double* sc, double* ss) {
    const double nf = cls == 0 ? 0.7 : (cls == 2 ? 1.5 : 1.0);
    int dd = (int)(cfg_d * nf);
    if (dd < 5) dd = 5;
    if (dd % 2 == 0) dd++;
    if (dd > 15) dd = 15;
    double c = cfg_sc * nf, s = cfg_ss * nf;
    if (stage_post) {
        dd = dd - 2 < 3 ? 3 : dd - 2;
        c *= 0.8;
        s *= 0.8;
    }
    *d = dd; *sc = c; *ss = s;
}

}  // namespace vp_host
