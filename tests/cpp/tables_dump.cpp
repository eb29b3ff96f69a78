// tables_dump.cpp - prints the host-built coefficient tables (csrc/vp_tables.h) as JSON so
// tests/test_host_tables.py can hold them against the oracle's restatement of OpenCV's formulas.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "vp_tables.h"

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    const std::string what = argv[1];
    if (what == "cubic") {
        const int s = std::atoi(argv[2]), d = std::atoi(argv[3]);
        vp_host::CubicTable t = vp_host::cubic_table(s, d);
        std::printf("{\"ofs\":[");
        for (int i = 0; i < d; ++i) std::printf("%s%d", i ? "," : "", t.ofs[i]);
        std::printf("],\"coef\":[");
        for (int i = 0; i < 4 * d; ++i) std::printf("%s%d", i ? "," : "", (int)t.coef[i]);
        std::printf("]}\n");
    } else if (what == "gauss") {
        const int k = std::atoi(argv[2]); const double sigma = std::atof(argv[3]);
        int q[7]; vp_host::gaussian_kernel_q8(k, sigma, q);
        std::vector<double> g = vp_host::gaussian_kernel_f64(k, sigma);
        std::printf("{\"q\":[");
        for (int i = 0; i < k; ++i) std::printf("%s%d", i ? "," : "", q[i]);
        std::printf("],\"f32\":[");
        for (int i = 0; i < k; ++i) { float f = (float)g[i]; unsigned u; std::memcpy(&u, &f, 4); std::printf("%s%u", i ? "," : "", u); }
        std::printf("]}\n");
    } else if (what == "bilateral") {
        vp_host::BilateralTables t = vp_host::bilateral_tables(std::atoi(argv[2]), std::atof(argv[3]), std::atof(argv[4]));
        std::printf("{\"radius\":%d,\"space\":[", t.radius);
        for (size_t i = 0; i < t.space_w.size(); ++i) { unsigned u; std::memcpy(&u, &t.space_w[i], 4); std::printf("%s%u", i ? "," : "", u); }
        std::printf("],\"color\":[");
        for (int i = 0; i < 768; ++i) { unsigned u; std::memcpy(&u, &t.color_w[i], 4); std::printf("%s%u", i ? "," : "", u); }
        std::printf("]}\n");
    } else if (what == "sigmoid") {
        float lut[256]; vp_host::sigmoid_lut((float)std::atof(argv[2]), lut);
        std::printf("{\"lut\":[");
        for (int i = 0; i < 256; ++i) { unsigned u; std::memcpy(&u, &lut[i], 4); std::printf("%s%u", i ? "," : "", u); }
        std::printf("]}\n");
    } else if (what == "adaptive") {
        int d; double sc, ss;
        vp_host::adaptive_params_for_class(std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]), std::atof(argv[5]), std::atof(argv[6]), &d, &sc, &ss);
        std::printf("{\"d\":%d,\"sc\":%.17g,\"ss\":%.17g}\n", d, sc, ss);
    }
    return 0;
}
