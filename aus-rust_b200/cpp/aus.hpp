// aus.hpp -- C++ host mirror of the `aus` crate's public API for the STFT + spectral-feature path, header-only, over the
// C ABI of include/aus_b200.h.  Same names, argument order and error behaviour as the Rust functions cited (paths
// under the reference's src/): Vec<f64> <-> std::vector<double>, Vec<Vec<Complex<f64>>> <-> vector<vector<complex>>,
// Result<_, SpectrumError> <-> throws aus::spectrum::SpectrumError, Rust panics <-> throws std::runtime_error.
#pragma once
#include <complex>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/aus_b200.h"

namespace aus {

enum class WindowType { Bartlett = 0, Blackman, Hanning, Hamming, Rectangular };   // window.rs:9-15

namespace detail {
struct Ctx {
    aus_ctx *h = nullptr;
    Ctx() {
        if (aus_ctx_create(0, &h) != AUS_OK) throw std::runtime_error(std::string("aus_b200: ") + aus_last_error());
    }
    ~Ctx() { aus_ctx_destroy(h); }
};
inline aus_ctx *ctx() {
    thread_local Ctx c;            // one context per calling thread: the reference's functions are re-entrant
    return c.h;
}
inline void ok(int rc) {
    if (rc != AUS_OK) throw std::runtime_error(std::string("aus_b200: ") + aus_last_error());
}
}  // namespace detail

inline std::vector<double> generate_window(WindowType t, size_t n) {               // window.rs:110
    std::vector<double> w(n);
    detail::ok(aus_generate_window((int)t, n, w.data()));
    return w;
}

namespace spectrum {
using cplx = std::complex<double>;
struct SpectrumError : std::runtime_error {                                          // fft.rs:16-19
    std::string error_msg;
    explicit SpectrumError(const std::string &m) : std::runtime_error(m), error_msg(m) {}
};
inline void ok_or_spectrum_error(int rc) {
    if (rc == AUS_ERR_SIZE) throw SpectrumError(aus_last_error());
    detail::ok(rc);
}

inline std::vector<cplx> rfft(const std::vector<double> &audio, size_t fft_size) {    // fft.rs:37
    if (audio.size() != fft_size) throw std::runtime_error("rfft: audio must have fft_size samples");
    std::vector<cplx> out(fft_size / 2 + 1);
    detail::ok(aus_rfft(detail::ctx(), audio.data(), fft_size, AUS_F64, reinterpret_cast<double *>(out.data())));
    return out;
}
inline std::vector<double> irfft(const std::vector<cplx> &spectrum, size_t fft_size) {   // fft.rs:66
    std::vector<double> out(fft_size);
    ok_or_spectrum_error(aus_irfft(detail::ctx(), reinterpret_cast<const double *>(spectrum.data()), spectrum.size(),
                                   fft_size, AUS_F64, out.data()));
    return out;
}
inline std::vector<std::vector<cplx>> rstft(const std::vector<double> &audio, size_t fft_size, size_t hop_size,
                                            WindowType window_type) {                 // fft.rs:119
    const size_t K = fft_size / 2 + 1, F = aus_num_frames(audio.size(), fft_size, hop_size);
    std::vector<cplx> flat(F * K);
    detail::ok(aus_rstft(detail::ctx(), audio.data(), audio.size(), fft_size, hop_size, (int)window_type, AUS_F64,
                         reinterpret_cast<double *>(flat.data())));
    std::vector<std::vector<cplx>> out(F);
    for (size_t f = 0; f < F; ++f) out[f].assign(flat.begin() + f * K, flat.begin() + (f + 1) * K);
    return out;
}
inline std::vector<double> irstft(const std::vector<std::vector<cplx>> &spectrogram, size_t fft_size, size_t hop_size,
                                  WindowType window_type) {                           // fft.rs:204
    const size_t K = fft_size / 2 + 1;
    for (size_t i = 0; i < spectrogram.size(); ++i)                                    // fft.rs:217-219
        if (spectrogram[i].size() != K)
            throw SpectrumError("The spectrum at frame " + std::to_string(i) + " has an incorrect length. It should be " +
                                std::to_string(K) + ", but it is actually " + std::to_string(spectrogram[i].size()) + ".");
    std::vector<cplx> flat;
    flat.reserve(spectrogram.size() * K);
    for (auto &r : spectrogram) flat.insert(flat.end(), r.begin(), r.end());
    std::vector<double> out(aus_irstft_len(spectrogram.size(), fft_size, hop_size));
    ok_or_spectrum_error(aus_irstft(detail::ctx(), reinterpret_cast<const double *>(flat.data()), spectrogram.size(), K,
                                    fft_size, hop_size, (int)window_type, AUS_F64, 1, out.data()));
    return out;
}
inline std::pair<std::vector<std::vector<double>>, std::vector<std::vector<double>>> complex_to_polar_rstft(
    const std::vector<std::vector<cplx>> &spectrogram) {                              // fft_tools.rs:41
    const size_t F = spectrogram.size(), K = F ? spectrogram[0].size() : 0;
    std::vector<cplx> flat;
    for (auto &r : spectrogram) flat.insert(flat.end(), r.begin(), r.end());
    std::vector<double> m(F * K), p(F * K);
    detail::ok(aus_complex_to_polar(detail::ctx(), reinterpret_cast<const double *>(flat.data()), F * K, m.data(), p.data()));
    std::vector<std::vector<double>> mm(F), pp(F);
    for (size_t f = 0; f < F; ++f) {
        mm[f].assign(m.begin() + f * K, m.begin() + (f + 1) * K);
        pp[f].assign(p.begin() + f * K, p.begin() + (f + 1) * K);
    }
    return {mm, pp};
}
inline std::vector<double> rfftfreq(size_t fft_size, uint32_t sample_rate) {          // fft_tools.rs:199
    std::vector<double> f(fft_size / 2 + 1);
    detail::ok(aus_rfftfreq(fft_size, sample_rate, f.data()));
    return f;
}
inline std::vector<double> overlap_add(const std::vector<std::vector<double>> &audio_chunks, size_t fft_size, size_t hop_size) {   // fft_tools.rs:69
    if (audio_chunks.empty()) return {};
    std::vector<double> flat;
    flat.reserve(audio_chunks.size() * fft_size);
    for (auto &ch : audio_chunks) {
        if (ch.size() != fft_size) throw std::runtime_error("overlap_add: every chunk must hold fft_size samples");
        flat.insert(flat.end(), ch.begin(), ch.end());
    }
    std::vector<double> out(fft_size + hop_size * (audio_chunks.size() - 1));
    detail::ok(aus_overlap_add(detail::ctx(), flat.data(), audio_chunks.size(), fft_size, hop_size, out.data()));
    return out;
}
}  // namespace spectrum

namespace analysis {
using Analysis = aus_analysis;                                                        // analyzer.rs:9-24
inline Analysis analyzer(const std::vector<double> &magnitude_spectrum, size_t fft_size, uint32_t sample_rate) {   // analyzer.rs:41
    if (magnitude_spectrum.size() != fft_size / 2 + 1) throw std::runtime_error("analyzer: wrong spectrum length");
    Analysis a;
    detail::ok(aus_analyzer_batch(detail::ctx(), magnitude_spectrum.data(), 1, fft_size, sample_rate, AUS_F64,
                                  reinterpret_cast<double *>(&a)));
    return a;
}
// The single-feature wrappers of spectral_analysis_tools.rs:306-471 with the reference's full parameter space (rows of any
// length, any frequency table, any roll-off fraction, any slope band) through the generic device entry point:
// [centroid, variance, skewness, kurtosis, entropy, flatness, slope, roll-offs..., band slopes...].
namespace detail {
inline std::vector<double> one(const std::vector<double> &m, const std::vector<double> *freqs, const std::vector<double> &roll,
                               const std::vector<double> &lo, const std::vector<double> &hi, uint32_t sample_rate) {
    std::vector<double> idx;
    if (freqs == nullptr) {
        idx.resize(m.size());
        for (size_t i = 0; i < m.size(); ++i) idx[i] = (double)i;
        freqs = &idx;
    }
    if (freqs->size() != m.size()) throw std::runtime_error("rfft_freqs and magnitude_spectrum differ in length");
    std::vector<double> out(7 + roll.size() + lo.size());
    aus::detail::ok(aus_spectral_features_generic(aus::detail::ctx(), m.data(), 1, m.size(), freqs->data(), roll.data(), roll.size(),
                                                  lo.data(), hi.data(), lo.size(), sample_rate, out.data()));
    return out;
}
}  // namespace detail
// The public building blocks of spectral_analysis_tools.rs:79-290 with the caller's own intermediates (aus_compute_kind).
namespace detail {
inline double compute(int what, const std::vector<double> &a, const std::vector<double> *b, double s0, double s1, uint32_t sr = 44100) {
    if (b && b->size() != a.size()) throw std::runtime_error("the two vectors differ in length");
    double out = 0.0;
    aus::detail::ok(aus_spectral_compute(aus::detail::ctx(), what, a.data(), b ? b->data() : nullptr, a.size(), s0, s1, sr, &out));
    return out;
}
inline std::vector<double> compute_vec(int what, const std::vector<double> &a, double s0) {
    std::vector<double> out(a.size());
    aus::detail::ok(aus_spectral_compute(aus::detail::ctx(), what, a.data(), nullptr, a.size(), s0, 0.0, 44100, out.data()));
    return out;
}
}  // namespace detail
inline double compute_spectral_centroid(const std::vector<double> &m, const std::vector<double> &f, double m_sum) { return detail::compute(AUS_COMPUTE_CENTROID, m, &f, m_sum, 0.0); }   // :79
inline double compute_spectral_entropy(const std::vector<double> &pmf) { return detail::compute(AUS_COMPUTE_ENTROPY, pmf, nullptr, 0.0, 0.0); }                                      // :93
inline double compute_spectral_flatness(const std::vector<double> &m, double m_sum) { return detail::compute(AUS_COMPUTE_FLATNESS, m, nullptr, m_sum, 0.0); }                        // :107
inline double compute_spectral_kurtosis(const std::vector<double> &pmf, const std::vector<double> &f, double c, double var) { return detail::compute(AUS_COMPUTE_KURTOSIS, pmf, &f, c, var); }   // :122
inline double compute_spectral_roll_off_point(const std::vector<double> &p, const std::vector<double> &f, double p_sum, double n) { return detail::compute(AUS_COMPUTE_ROLL_OFF, p, &f, p_sum, n); }   // :137
inline double compute_spectral_skewness(const std::vector<double> &pmf, const std::vector<double> &f, double c, double var) { return detail::compute(AUS_COMPUTE_SKEWNESS, pmf, &f, c, var); }   // :153
inline double compute_spectral_slope(const std::vector<double> &p, double p_sum) { return detail::compute(AUS_COMPUTE_SLOPE, p, nullptr, p_sum, 0.0); }                              // :167
inline double compute_spectral_slope_region(const std::vector<double> &p, const std::vector<double> &f, double f_lower, double f_upper, uint32_t sample_rate) {                      // :190
    return detail::compute(AUS_COMPUTE_SLOPE_REGION, p, &f, f_lower, f_upper, sample_rate);
}
inline double compute_spectral_variance(const std::vector<double> &pmf, const std::vector<double> &f, double c) { return detail::compute(AUS_COMPUTE_VARIANCE, pmf, &f, c, 0.0); }     // :238
inline double dot_product(const std::vector<double> &a, const std::vector<double> &b) { return detail::compute(AUS_COMPUTE_DOT_PRODUCT, a, &b, 0.0, 0.0); }                          // :253
inline std::vector<double> make_power_spectrum(const std::vector<double> &m) { return detail::compute_vec(AUS_COMPUTE_POWER_SPECTRUM, m, 0.0); }                                     // :273
inline std::vector<double> make_spectrum_pmf(const std::vector<double> &p, double p_sum) { return detail::compute_vec(AUS_COMPUTE_SPECTRUM_PMF, p, p_sum); }                          // :284

inline double spectral_centroid(const std::vector<double> &m, const std::vector<double> &f) { return detail::one(m, &f, {}, {}, {}, 44100)[0]; }   // :306
inline double spectral_variance(const std::vector<double> &m, const std::vector<double> &f) { return detail::one(m, &f, {}, {}, {}, 44100)[1]; }   // :466
inline double spectral_skewness(const std::vector<double> &m, const std::vector<double> &f) { return detail::one(m, &f, {}, {}, {}, 44100)[2]; }   // :405
inline double spectral_kurtosis(const std::vector<double> &m, const std::vector<double> &f) { return detail::one(m, &f, {}, {}, {}, 44100)[3]; }   // :362
inline double spectral_entropy(const std::vector<double> &m) { return detail::one(m, nullptr, {}, {}, {}, 44100)[4]; }                             // :324
inline double spectral_flatness(const std::vector<double> &m) { return detail::one(m, nullptr, {}, {}, {}, 44100)[5]; }                            // :343
inline double spectral_slope(const std::vector<double> &m) { return detail::one(m, nullptr, {}, {}, {}, 44100)[6]; }                               // :426
inline double spectral_roll_off_point(const std::vector<double> &m, const std::vector<double> &f, double n) {                                    // :385
    return detail::one(m, &f, {n}, {}, {}, 44100)[7];
}
inline double spectral_slope_region(const std::vector<double> &m, const std::vector<double> &f, double f_lower, double f_upper,
                                    uint32_t sample_rate) {                                                                                       // :447
    return detail::one(m, &f, {}, {f_lower}, {f_upper}, sample_rate)[7];
}
}  // namespace analysis

namespace mp {
// mp.rs:26 -- hop = fft_size/2, Hamming; max_num_threads accepted and ignored
inline std::vector<analysis::Analysis> stft_analysis(std::vector<double> &audio, size_t fft_size, uint32_t sample_rate,
                                                     std::optional<size_t> max_num_threads) {
    std::vector<analysis::Analysis> out(aus_num_frames(audio.size(), fft_size, fft_size / 2));
    detail::ok(aus_stft_analysis(detail::ctx(), audio.data(), audio.size(), fft_size, sample_rate,
                                 max_num_threads.value_or(0), out.data()));
    return out;
}
// general hop / window (new entry point)
inline std::vector<analysis::Analysis> stft_features(const std::vector<double> &audio, size_t fft_size, size_t hop_size,
                                                     WindowType window_type, uint32_t sample_rate) {
    std::vector<analysis::Analysis> out(aus_num_frames(audio.size(), fft_size, hop_size));
    detail::ok(aus_stft_features(detail::ctx(), audio.data(), audio.size(), fft_size, hop_size, (int)window_type,
                                 sample_rate, AUS_F64, reinterpret_cast<double *>(out.data())));
    return out;
}
}  // namespace mp

}  // namespace aus
