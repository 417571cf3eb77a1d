// Exercises aus-rust_b200/cpp/aus.hpp the way the reference's own tests exercise the crate (mp.rs:127-136,
// fft.rs:277-299): read audio, analyse, round-trip.  Prints values for tests/test_cpp_mirror.py to compare.
#include <cmath>
#include <cstdio>
#include <fstream>
#include "../aus-rust_b200/cpp/aus.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<double> audio;
    double v;
    while (in.read(reinterpret_cast<char *>(&v), sizeof v)) audio.push_back(v);
    try {
        auto an = aus::mp::stft_analysis(audio, 2048, 44100, 8);
        std::printf("frames %zu\n", an.size());
        for (size_t i = 0; i < an.size(); i += an.size() / 3 + 1) {
            const double *r = reinterpret_cast<const double *>(&an[i]);
            std::printf("row %zu", i);
            for (int k = 0; k < 14; ++k) std::printf(" %.17g", r[k]);
            std::printf("\n");
        }
        auto S = aus::spectrum::rstft(audio, 2048, 1024, aus::WindowType::Hamming);
        auto y = aus::spectrum::irstft(S, 2048, 1024, aus::WindowType::Hamming);
        double mx = 0;
        for (double s : y) mx = std::fmax(mx, s);
        std::printf("irstft %zu %.17g\n", y.size(), mx);
        {   // single-feature wrappers with non-analyzer parameters (spectral_analysis_tools.rs:385, :447)
            std::vector<double> mag(S[1].size());
            for (size_t k = 0; k < mag.size(); ++k) mag[k] = std::abs(S[1][k]);
            const auto freqs = aus::spectrum::rfftfreq(2048, 44100);
            std::printf("wrappers %.17g %.17g %.17g\n", aus::analysis::spectral_roll_off_point(mag, freqs, 0.85),
                        aus::analysis::spectral_slope_region(mag, freqs, 105.0, 852.0, 44100), aus::analysis::spectral_centroid(mag, freqs));
        }
        S[1].pop_back();
        try {
            aus::spectrum::irstft(S, 2048, 1024, aus::WindowType::Hamming);
            std::printf("error none\n");
        } catch (const aus::spectrum::SpectrumError &e) {
            std::printf("error %s\n", e.error_msg.c_str());
        }
    } catch (const std::exception &e) {
        std::printf("exception %s\n", e.what());
        return 1;
    }
    return 0;
}
