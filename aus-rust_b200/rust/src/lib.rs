//! Drop-in replacement for the STFT + spectral-feature path of `aus` (fleximeter/aus-rust 0.1.7) on B200.
//!
//! Every public function below has the *same signature* as the reference function it replaces (cited by
//! `file:line` of the reference's `src/`), and forwards to the C ABI of `include/aus_b200.h`.  All arithmetic
//! happens in the CUDA library; this file only reshapes `Vec`s.  There is no CPU fallback: if no B200 is present the
//! functions panic with the library's error message (the reference's own failure mode on this path is a panic too).
//!
//! NOT compiled in the repository's build environment (no Rust toolchain there); kept deliberately thin.
#![allow(clippy::needless_range_loop)]

use num::Complex;
use std::cell::RefCell;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

// ------------------------------------------------------------------------------------------------ FFI (include/aus_b200.h)
#[repr(C)]
pub struct AusCtx { _private: [u8; 0] }

/// `#[repr(C)]` mirror of `aus_analysis`; `Analysis` itself has Rust-unspecified layout (analyzer.rs:9-24).
#[repr(C)]
#[derive(Copy, Clone, Default)]
struct AusAnalysis { v: [f64; 14] }

extern "C" {
    fn aus_last_error() -> *const c_char;
    fn aus_ctx_create(device: c_int, out: *mut *mut AusCtx) -> c_int;
    fn aus_ctx_destroy(ctx: *mut AusCtx);
    fn aus_num_frames(n_samples: usize, fft_size: usize, hop: usize) -> usize;
    fn aus_irstft_len(n_frames: usize, fft_size: usize, hop: usize) -> usize;
    fn aus_generate_window(window_type: c_int, window_length: usize, out: *mut f64) -> c_int;
    fn aus_rfftfreq(fft_size: usize, sample_rate: u32, out: *mut f64) -> c_int;
    fn aus_stft_analysis(ctx: *mut AusCtx, audio: *const f64, n: usize, fft_size: usize, sample_rate: u32,
                         max_num_threads: usize, out: *mut AusAnalysis) -> c_int;
    fn aus_stft_features(ctx: *mut AusCtx, audio: *const f64, n: usize, fft_size: usize, hop: usize, window: c_int,
                         sample_rate: u32, precision: c_int, features: *mut f64) -> c_int;
    fn aus_rstft(ctx: *mut AusCtx, audio: *const f64, n: usize, fft_size: usize, hop: usize, window: c_int,
                 precision: c_int, spectra: *mut f64) -> c_int;
    fn aus_rfft(ctx: *mut AusCtx, audio: *const f64, fft_size: usize, precision: c_int, spectrum: *mut f64) -> c_int;
    fn aus_irstft(ctx: *mut AusCtx, spectra: *const f64, n_frames: usize, frame_len: usize, fft_size: usize, hop: usize,
                  window: c_int, precision: c_int, normalize: c_int, audio: *mut f64) -> c_int;
    fn aus_irfft(ctx: *mut AusCtx, spectrum: *const f64, spectrum_len: usize, fft_size: usize, precision: c_int,
                 audio: *mut f64) -> c_int;
    fn aus_complex_to_polar(ctx: *mut AusCtx, spectra: *const f64, n: usize, mag: *mut f64, ph: *mut f64) -> c_int;
    fn aus_polar_to_complex(ctx: *mut AusCtx, mag: *const f64, ph: *const f64, n: usize, spectra: *mut f64) -> c_int;
    fn aus_analyzer_batch(ctx: *mut AusCtx, mags: *const f64, n_frames: usize, fft_size: usize, sample_rate: u32,
                          precision: c_int, features: *mut f64) -> c_int;
    fn aus_spectral_features_generic(ctx: *mut AusCtx, mags: *const f64, n_rows: usize, n_bins: usize, rfft_freqs: *const f64,
                                     roll_fractions: *const f64, n_roll: usize, band_lo: *const f64, band_hi: *const f64,
                                     n_bands: usize, sample_rate: u32, out: *mut f64) -> c_int;
    // round 2
    fn aus_irstft_batch(ctx: *mut AusCtx, spectra: *const f64, n_frames: *const u64, n_spectrograms: usize, frame_len: usize,
                        fft_size: usize, hop: usize, window: c_int, precision: c_int, normalize: c_int, audio: *mut f64) -> c_int;
    fn aus_irstft_multi(devices: *const c_int, n_devices: c_int, spectra: *const f64, n_frames: *const u64, n_spectrograms: usize,
                        frame_len: usize, fft_size: usize, hop: usize, window: c_int, normalize: c_int, audio: *mut f64) -> c_int;
    fn aus_stft_features_multi(devices: *const c_int, n_devices: c_int, audio: *const f64, offsets: *const u64, lengths: *const u64,
                               n_files: usize, fft_size: usize, hop: usize, window: c_int, sample_rate: u32, precision: c_int,
                               features: *mut f64) -> c_int;
    fn aus_spectral_compute(ctx: *mut AusCtx, what: c_int, a: *const f64, b: *const f64, n: usize, s0: f64, s1: f64, sample_rate: u32,
                            out: *mut f64) -> c_int;
    fn aus_overlap_add(ctx: *mut AusCtx, chunks: *const f64, n_chunks: usize, fft_size: usize, hop: usize, audio: *mut f64) -> c_int;
    fn aus_autocorrelation(ctx: *mut AusCtx, audio: *const f64, audio_len: usize, fft_size: usize, out: *mut f64) -> c_int;
    fn aus_time_features_batch(ctx: *mut AusCtx, audio: *const f64, offsets: *const u64, lengths: *const u64, n_files: usize,
                               sample_rate: u32, out: *mut f64) -> c_int;
    fn aus_time_features_frames(ctx: *mut AusCtx, audio: *const f64, offsets: *const u64, lengths: *const u64, n_files: usize,
                                fft_size: usize, hop: usize, sample_rate: u32, out: *mut f64) -> c_int;
    fn aus_mixdown(ctx: *mut AusCtx, channels: *const f64, n_frames: usize, n_channels: c_int, out: *mut f64) -> c_int;
    fn aus_pcm_to_f64(ctx: *mut AusCtx, pcm: *const c_void, sample_bits: c_int, n_frames: usize, n_channels: c_int,
                      interleaved: c_int, mixdown: c_int, out: *mut f64) -> c_int;
}

const AUS_ERR_SIZE: c_int = -4;
const F64: c_int = 0;

struct Ctx(*mut AusCtx);
impl Drop for Ctx { fn drop(&mut self) { unsafe { aus_ctx_destroy(self.0) } } }
thread_local! { static CTX: RefCell<Option<Ctx>> = RefCell::new(None); }

fn last_error() -> String { unsafe { CStr::from_ptr(aus_last_error()).to_string_lossy().into_owned() } }

/// One context per calling thread (the reference's functions are stateless and re-entrant).
fn with_ctx<R>(f: impl FnOnce(*mut AusCtx) -> R) -> R {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() {
            let dev = std::env::var("AUS_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            let mut h: *mut AusCtx = std::ptr::null_mut();
            let rc = unsafe { aus_ctx_create(dev, &mut h) };
            if rc != 0 { panic!("aus_b200: {}", last_error()); }
            *c = Some(Ctx(h));
        }
        f(c.as_ref().unwrap().0)
    })
}
fn ok(rc: c_int) { if rc != 0 { panic!("aus_b200: {}", last_error()); } }

// ------------------------------------------------------------------------------------------------ window.rs
/// window.rs:9-15
#[derive(Copy, Clone)]
pub enum WindowType { Bartlett, Blackman, Hanning, Hamming, Rectangular }
fn wcode(w: WindowType) -> c_int { w as c_int }

/// window.rs:110
pub fn generate_window(window_type: WindowType, window_length: usize) -> Vec<f64> {
    let mut w = vec![0.0; window_length];
    ok(unsafe { aus_generate_window(wcode(window_type), window_length, w.as_mut_ptr()) });
    w
}
pub fn generate_window_bartlett(n: usize) -> Vec<f64> { generate_window(WindowType::Bartlett, n) }       // window.rs:26
pub fn generate_window_blackman(n: usize) -> Vec<f64> { generate_window(WindowType::Blackman, n) }       // window.rs:43
pub fn generate_window_hanning(n: usize) -> Vec<f64> { generate_window(WindowType::Hanning, n) }         // window.rs:61
pub fn generate_window_hamming(n: usize) -> Vec<f64> { generate_window(WindowType::Hamming, n) }         // window.rs:78
pub fn generate_window_rectangular(n: usize) -> Vec<f64> { generate_window(WindowType::Rectangular, n) } // window.rs:95

// ------------------------------------------------------------------------------------------------ spectrum
pub mod spectrum {
    use super::*;

    /// spectrum/fft.rs:16-19
    #[derive(Debug, Clone)]
    pub struct SpectrumError { pub error_msg: String }

    fn rows(flat: Vec<f64>, k: usize) -> Vec<Vec<Complex<f64>>> {
        flat.chunks_exact(2 * k).map(|r| r.chunks_exact(2).map(|c| Complex::new(c[0], c[1])).collect()).collect()
    }
    fn flat(s: &[Vec<Complex<f64>>]) -> Vec<f64> {
        let mut v = Vec::with_capacity(s.len() * s.first().map_or(0, |r| r.len()) * 2);
        for r in s { for c in r { v.push(c.re); v.push(c.im); } }
        v
    }

    /// spectrum/fft.rs:37
    pub fn rfft(audio: &[f64], fft_size: usize) -> Vec<Complex<f64>> {
        assert_eq!(audio.len(), fft_size, "rfft: zero-pad the audio to fft_size first (fft.rs:23-24)");
        let k = fft_size / 2 + 1;
        let mut out = vec![0.0; 2 * k];
        with_ctx(|c| ok(unsafe { aus_rfft(c, audio.as_ptr(), fft_size, F64, out.as_mut_ptr()) }));
        out.chunks_exact(2).map(|c| Complex::new(c[0], c[1])).collect()
    }

    /// spectrum/fft.rs:66
    pub fn irfft(spectrum: &[Complex<f64>], fft_size: usize) -> Result<Vec<f64>, SpectrumError> {
        let s: Vec<f64> = spectrum.iter().flat_map(|c| [c.re, c.im]).collect();
        let mut out = vec![0.0; fft_size];
        let rc = with_ctx(|c| unsafe { aus_irfft(c, s.as_ptr(), spectrum.len(), fft_size, F64, out.as_mut_ptr()) });
        if rc == AUS_ERR_SIZE { return Err(SpectrumError { error_msg: last_error() }); }
        ok(rc);
        Ok(out)
    }

    /// spectrum/fft.rs:119
    pub fn rstft(audio: &Vec<f64>, fft_size: usize, hop_size: usize, window_type: WindowType) -> Vec<Vec<Complex<f64>>> {
        let k = fft_size / 2 + 1;
        let f = unsafe { aus_num_frames(audio.len(), fft_size, hop_size) };
        let mut out = vec![0.0; f * k * 2];
        with_ctx(|c| ok(unsafe {
            aus_rstft(c, audio.as_ptr(), audio.len(), fft_size, hop_size, wcode(window_type), F64, out.as_mut_ptr())
        }));
        rows(out, k)
    }

    /// spectrum/fft.rs:204
    pub fn irstft(spectrogram: &Vec<Vec<Complex<f64>>>, fft_size: usize, hop_size: usize, window_type: WindowType)
                  -> Result<Vec<f64>, SpectrumError> {
        let k = fft_size / 2 + 1;
        for (i, r) in spectrogram.iter().enumerate() {      // fft.rs:217-219: per-frame length check
            if r.len() != k {
                return Err(SpectrumError { error_msg: format!(
                    "The spectrum at frame {} has an incorrect length. It should be {}, but it is actually {}.", i, k, r.len()) });
            }
        }
        let s = flat(spectrogram);
        let mut out = vec![0.0; unsafe { aus_irstft_len(spectrogram.len(), fft_size, hop_size) }];
        let rc = with_ctx(|c| unsafe {
            aus_irstft(c, s.as_ptr(), spectrogram.len(), k, fft_size, hop_size, wcode(window_type), F64, 1, out.as_mut_ptr())
        });
        if rc == AUS_ERR_SIZE { return Err(SpectrumError { error_msg: last_error() }); }
        ok(rc);                                             // empty spectrogram: the reference panics too (fft.rs:208)
        Ok(out)
    }

    /// spectrum/fft_tools.rs:21
    pub fn complex_to_polar_rfft(spectrum: &Vec<Complex<f64>>) -> (Vec<f64>, Vec<f64>) {
        let s: Vec<f64> = spectrum.iter().flat_map(|c| [c.re, c.im]).collect();
        let (mut m, mut p) = (vec![0.0; spectrum.len()], vec![0.0; spectrum.len()]);
        with_ctx(|c| ok(unsafe { aus_complex_to_polar(c, s.as_ptr(), spectrum.len(), m.as_mut_ptr(), p.as_mut_ptr()) }));
        (m, p)
    }

    /// spectrum/fft_tools.rs:41
    pub fn complex_to_polar_rstft(spectrogram: &Vec<Vec<Complex<f64>>>) -> (Vec<Vec<f64>>, Vec<Vec<f64>>) {
        let k = spectrogram.first().map_or(0, |r| r.len());
        let s = flat(spectrogram);
        let n = s.len() / 2;
        let (mut m, mut p) = (vec![0.0; n], vec![0.0; n]);
        with_ctx(|c| ok(unsafe { aus_complex_to_polar(c, s.as_ptr(), n, m.as_mut_ptr(), p.as_mut_ptr()) }));
        (m.chunks(k.max(1)).map(|r| r.to_vec()).collect(), p.chunks(k.max(1)).map(|r| r.to_vec()).collect())
    }

    /// spectrum/fft_tools.rs:131
    pub fn polar_to_complex_rfft(magnitude_spectrum: &Vec<f64>, phase_spectrum: &Vec<f64>) -> Result<Vec<Complex<f64>>, SpectrumError> {
        if magnitude_spectrum.len() != phase_spectrum.len() {
            return Err(SpectrumError { error_msg: format!(
                "The magnitude spectrum and phase spectrum do not have the same length. The magnitude spectrum has len {} and the phase spectrum has len {}.",
                magnitude_spectrum.len(), phase_spectrum.len()) });
        }
        let mut out = vec![0.0; 2 * magnitude_spectrum.len()];
        with_ctx(|c| ok(unsafe {
            aus_polar_to_complex(c, magnitude_spectrum.as_ptr(), phase_spectrum.as_ptr(), magnitude_spectrum.len(), out.as_mut_ptr())
        }));
        Ok(out.chunks_exact(2).map(|c| Complex::new(c[0], c[1])).collect())
    }

    /// spectrum/fft_tools.rs:164
    pub fn polar_to_complex_rstft(magnitude_spectrogram: &Vec<Vec<f64>>, phase_spectrogram: &Vec<Vec<f64>>)
                                  -> Result<Vec<Vec<Complex<f64>>>, SpectrumError> {
        if magnitude_spectrogram.len() != phase_spectrogram.len() {
            return Err(SpectrumError { error_msg: format!(
                "The magnitude spectrogram and phase spectrogram do not have the same length. The magnitude spectrogram has len {} and the phase spectrogram has len {}.",
                magnitude_spectrogram.len(), phase_spectrogram.len()) });
        }
        magnitude_spectrogram.iter().zip(phase_spectrogram.iter()).map(|(m, p)| polar_to_complex_rfft(m, p)).collect()
    }

    /// spectrum/fft_tools.rs:69
    pub fn overlap_add(audio_chunks: &Vec<Vec<f64>>, fft_size: usize, hop_size: usize) -> Vec<f64> {
        if audio_chunks.is_empty() { return Vec::new(); }
        let mut flat: Vec<f64> = Vec::with_capacity(audio_chunks.len() * fft_size);
        for ch in audio_chunks { assert_eq!(ch.len(), fft_size); flat.extend_from_slice(ch); }
        let mut out = vec![0.0f64; fft_size + hop_size * (audio_chunks.len() - 1)];
        with_ctx(|c| ok(unsafe { aus_overlap_add(c, flat.as_ptr(), audio_chunks.len(), fft_size, hop_size, out.as_mut_ptr()) }));
        out
    }

    /// spectrum/fft_tools.rs:199
    pub fn rfftfreq(fft_size: usize, sample_rate: u32) -> Vec<f64> {
        let mut f = vec![0.0; fft_size / 2 + 1];
        ok(unsafe { aus_rfftfreq(fft_size, sample_rate, f.as_mut_ptr()) });
        f
    }
}

// ------------------------------------------------------------------------------------------------ analysis
pub mod analysis {
    use super::*;

    /// analysis/analyzer.rs:9-24
    #[derive(Copy, Clone)]
    pub struct Analysis {
        pub spectral_centroid: f64,
        pub spectral_variance: f64,
        pub spectral_skewness: f64,
        pub spectral_kurtosis: f64,
        pub spectral_entropy: f64,
        pub spectral_flatness: f64,
        pub spectral_roll_off_50: f64,
        pub spectral_roll_off_75: f64,
        pub spectral_roll_off_90: f64,
        pub spectral_roll_off_95: f64,
        pub spectral_slope: f64,
        pub spectral_slope_0_1_khz: f64,
        pub spectral_slope_1_5_khz: f64,
        pub spectral_slope_0_5_khz: f64,
    }
    pub(crate) fn from_row(v: &[f64]) -> Analysis {     // field-by-field copy from the repr(C) row
        Analysis {
            spectral_centroid: v[0], spectral_variance: v[1], spectral_skewness: v[2], spectral_kurtosis: v[3],
            spectral_entropy: v[4], spectral_flatness: v[5], spectral_roll_off_50: v[6], spectral_roll_off_75: v[7],
            spectral_roll_off_90: v[8], spectral_roll_off_95: v[9], spectral_slope: v[10],
            spectral_slope_0_1_khz: v[11], spectral_slope_1_5_khz: v[12], spectral_slope_0_5_khz: v[13],
        }
    }

    /// analysis/analyzer.rs:41
    pub fn analyzer(magnitude_spectrum: &Vec<f64>, fft_size: usize, sample_rate: u32) -> Analysis {
        assert_eq!(magnitude_spectrum.len(), fft_size / 2 + 1);
        let mut out = [0.0f64; 14];
        with_ctx(|c| ok(unsafe {
            aus_analyzer_batch(c, magnitude_spectrum.as_ptr(), 1, fft_size, sample_rate, F64, out.as_mut_ptr())
        }));
        from_row(&out)
    }

    // The public building blocks of spectral_analysis_tools.rs:79-290: same signatures, and -- like the reference -- they use
    // whatever intermediates (sums, pmf, centroid, variance) the caller passes.  `what` = aus_compute_kind of aus_b200.h.
    fn compute(what: c_int, a: &[f64], b: Option<&[f64]>, s0: f64, s1: f64, sample_rate: u32) -> f64 {
        if let Some(b) = b { assert_eq!(a.len(), b.len()); }
        let mut out = 0.0f64;
        with_ctx(|c| ok(unsafe {
            aus_spectral_compute(c, what, a.as_ptr(), b.map_or(std::ptr::null(), |b| b.as_ptr()), a.len(), s0, s1, sample_rate, &mut out)
        }));
        out
    }
    fn compute_vec(what: c_int, a: &[f64], s0: f64) -> Vec<f64> {
        let mut out = vec![0.0f64; a.len()];
        with_ctx(|c| ok(unsafe { aus_spectral_compute(c, what, a.as_ptr(), std::ptr::null(), a.len(), s0, 0.0, 44100, out.as_mut_ptr()) }));
        out
    }
    pub fn compute_spectral_centroid(magnitude_spectrum: &Vec<f64>, rfft_freqs: &Vec<f64>, magnitude_spectrum_sum: f64) -> f64 {   // :79
        compute(0, magnitude_spectrum, Some(rfft_freqs.as_slice()), magnitude_spectrum_sum, 0.0, 44100)
    }
    pub fn compute_spectral_entropy(spectrum_pmf: &Vec<f64>) -> f64 { compute(1, spectrum_pmf, None, 0.0, 0.0, 44100) }             // :93
    pub fn compute_spectral_flatness(magnitude_spectrum: &Vec<f64>, magnitude_spectrum_sum: f64) -> f64 {                          // :107
        compute(2, magnitude_spectrum, None, magnitude_spectrum_sum, 0.0, 44100)
    }
    pub fn compute_spectral_kurtosis(spectrum_pmf: &Vec<f64>, rfft_freqs: &Vec<f64>, spectral_centroid: f64, spectral_variance: f64) -> f64 {   // :122
        compute(3, spectrum_pmf, Some(rfft_freqs.as_slice()), spectral_centroid, spectral_variance, 44100)
    }
    pub fn compute_spectral_roll_off_point(power_spectrum: &Vec<f64>, rfft_freqs: &Vec<f64>, power_spectrum_sum: f64, n: f64) -> f64 {          // :137
        compute(4, power_spectrum, Some(rfft_freqs.as_slice()), power_spectrum_sum, n, 44100)
    }
    pub fn compute_spectral_skewness(spectrum_pmf: &Vec<f64>, rfft_freqs: &Vec<f64>, spectral_centroid: f64, spectral_variance: f64) -> f64 {   // :153
        compute(5, spectrum_pmf, Some(rfft_freqs.as_slice()), spectral_centroid, spectral_variance, 44100)
    }
    pub fn compute_spectral_slope(power_spectrum: &Vec<f64>, power_spectrum_sum: f64) -> f64 {                                     // :167
        compute(6, power_spectrum, None, power_spectrum_sum, 0.0, 44100)
    }
    pub fn compute_spectral_slope_region(power_spectrum: &Vec<f64>, rfft_freqs: &Vec<f64>, f_lower: f64, f_upper: f64, sample_rate: u32) -> f64 {  // :190
        compute(7, power_spectrum, Some(rfft_freqs.as_slice()), f_lower, f_upper, sample_rate)
    }
    pub fn compute_spectral_variance(spectrum_pmf: &Vec<f64>, rfft_freqs: &Vec<f64>, spectral_centroid: f64) -> f64 {               // :238
        compute(8, spectrum_pmf, Some(rfft_freqs.as_slice()), spectral_centroid, 0.0, 44100)
    }
    pub fn dot_product(vec1: &[f64], vec2: &[f64]) -> f64 { compute(9, vec1, Some(vec2), 0.0, 0.0, 44100) }                        // :253
    pub fn make_power_spectrum(magnitude_spectrum: &Vec<f64>) -> Vec<f64> { compute_vec(10, magnitude_spectrum, 0.0) }              // :273
    pub fn make_spectrum_pmf(power_spectrum: &Vec<f64>, power_spectrum_sum: f64) -> Vec<f64> {                                      // :284
        compute_vec(11, power_spectrum, power_spectrum_sum)
    }

    // The single-feature wrappers (spectral_analysis_tools.rs:306-471) keep the reference's full parameter space -- rows of
    // any length, any frequency table, any roll-off fraction, any slope band -- through the generic device entry point:
    // [centroid, variance, skewness, kurtosis, entropy, flatness, slope, roll-offs..., band slopes...].
    fn one(m: &Vec<f64>, freqs: Option<&Vec<f64>>, roll: &[f64], band: Option<(f64, f64)>, sample_rate: u32) -> Vec<f64> {
        let idx: Vec<f64>;
        let f: &Vec<f64> = match freqs { Some(f) => { assert_eq!(f.len(), m.len()); f }, None => { idx = (0..m.len()).map(|i| i as f64).collect(); &idx } };
        let (lo, hi): (Vec<f64>, Vec<f64>) = band.map_or((vec![], vec![]), |(l, h)| (vec![l], vec![h]));
        let mut out = vec![0.0f64; 7 + roll.len() + lo.len()];
        with_ctx(|c| ok(unsafe {
            aus_spectral_features_generic(c, m.as_ptr(), 1, m.len(), f.as_ptr(), roll.as_ptr(), roll.len(), lo.as_ptr(), hi.as_ptr(),
                                          lo.len(), sample_rate, out.as_mut_ptr())
        }));
        out
    }
    pub fn spectral_centroid(m: &Vec<f64>, rfft_freqs: &Vec<f64>) -> f64 { one(m, Some(rfft_freqs), &[], None, 44100)[0] }   // :306
    pub fn spectral_variance(m: &Vec<f64>, rfft_freqs: &Vec<f64>) -> f64 { one(m, Some(rfft_freqs), &[], None, 44100)[1] }   // :466
    pub fn spectral_skewness(m: &Vec<f64>, rfft_freqs: &Vec<f64>) -> f64 { one(m, Some(rfft_freqs), &[], None, 44100)[2] }   // :405
    pub fn spectral_kurtosis(m: &Vec<f64>, rfft_freqs: &Vec<f64>) -> f64 { one(m, Some(rfft_freqs), &[], None, 44100)[3] }   // :362
    pub fn spectral_entropy(m: &Vec<f64>) -> f64 { one(m, None, &[], None, 44100)[4] }                                      // :324
    pub fn spectral_flatness(m: &Vec<f64>) -> f64 { one(m, None, &[], None, 44100)[5] }                                     // :343
    pub fn spectral_slope(m: &Vec<f64>) -> f64 { one(m, None, &[], None, 44100)[6] }                                        // :426
    /// :385
    pub fn spectral_roll_off_point(m: &Vec<f64>, rfft_freqs: &Vec<f64>, n: f64) -> f64 { one(m, Some(rfft_freqs), &[n], None, 44100)[7] }
    /// :447 (a band that reaches the last bin panics, as in the reference)
    pub fn spectral_slope_region(m: &Vec<f64>, rfft_freqs: &Vec<f64>, f_lower: f64, f_upper: f64, sample_rate: u32) -> f64 {
        one(m, Some(rfft_freqs), &[], Some((f_lower, f_upper)), sample_rate)[7]
    }
}

// ------------------------------------------------------------------------------------------------ mp
pub mod mp {
    use super::*;
    use crate::analysis::Analysis;

    /// mp.rs:26 -- same signature; `audio` is never mutated (as in the reference), `max_num_threads` is ignored.
    pub fn stft_analysis(audio: &mut Vec<f64>, fft_size: usize, sample_rate: u32, max_num_threads: Option<usize>) -> Vec<Analysis> {
        let f = unsafe { aus_num_frames(audio.len(), fft_size, fft_size / 2) };
        let mut rows = vec![AusAnalysis::default(); f];
        with_ctx(|c| ok(unsafe {
            aus_stft_analysis(c, audio.as_ptr(), audio.len(), fft_size, sample_rate, max_num_threads.unwrap_or(0), rows.as_mut_ptr())
        }));
        rows.iter().map(|r| analysis::from_row(&r.v)).collect()
    }

    /// New general entry point (hop and window are hard-coded inside mp.rs:42).
    pub fn stft_features(audio: &[f64], fft_size: usize, hop_size: usize, window_type: WindowType, sample_rate: u32) -> Vec<Analysis> {
        let f = unsafe { aus_num_frames(audio.len(), fft_size, hop_size) };
        let mut out = vec![0.0f64; f * 14];
        with_ctx(|c| ok(unsafe {
            aus_stft_features(c, audio.as_ptr(), audio.len(), fft_size, hop_size, wcode(window_type), sample_rate, F64, out.as_mut_ptr())
        }));
        out.chunks_exact(14).map(analysis::from_row).collect()
    }
}

// ------------------------------------------------------------------------------------------------ round-2 rows
/// audiofile.rs:168-178 on the samples of an `AudioFile` (`samples: Vec<Vec<f64>>`, one Vec per channel): channel 0
/// accumulates every channel including itself, then divides by the channel count; the other channels are dropped.
pub fn mixdown_samples(samples: &mut Vec<Vec<f64>>) {
    if samples.len() > 1 {
        let n = samples[0].len();
        let flat: Vec<f64> = samples.iter().flat_map(|c| c.iter().copied()).collect();      // planar [channel][frame]
        let mut out = vec![0.0f64; n];
        with_ctx(|c| ok(unsafe { aus_mixdown(c, flat.as_ptr(), n, samples.len() as c_int, out.as_mut_ptr()) }));
        samples[0] = out;
        samples.truncate(1);
    }
}

/// audiofile.rs:296-329 for one decoded packet of interleaved 16-bit PCM: `sample as f64 / INTMAX16` per channel.
pub fn widen_s16_interleaved(pcm: &[i16], n_channels: usize) -> Vec<Vec<f64>> {
    let n = pcm.len() / n_channels;
    let mut flat = vec![0.0f64; n * n_channels];
    with_ctx(|c| ok(unsafe {
        aus_pcm_to_f64(c, pcm.as_ptr() as *const c_void, 16, n, n_channels as c_int, 1, 0, flat.as_mut_ptr())
    }));
    flat.chunks_exact(n).map(|c| c.to_vec()).collect()
}

/// analysis/spectral_analysis_tools.rs:31 -- same signature
pub fn autocorrelation(audio: &[f64], fft_size: usize) -> Result<Vec<f64>, spectrum::SpectrumError> {
    let mut out = vec![0.0f64; fft_size];
    let rc = with_ctx(|c| unsafe { aus_autocorrelation(c, audio.as_ptr(), audio.len(), fft_size, out.as_mut_ptr()) });
    if rc == AUS_ERR_SIZE { return Err(spectrum::SpectrumError { error_msg: last_error() }); }
    ok(rc);
    Ok(out)
}

fn time_scalars(audio: &[f64], sample_rate: u32) -> [f64; 4] {
    let (off, len) = ([0u64], [audio.len() as u64]);
    let mut out = [0.0f64; 4];
    with_ctx(|c| ok(unsafe { aus_time_features_batch(c, audio.as_ptr(), off.as_ptr(), len.as_ptr(), 1, sample_rate, out.as_mut_ptr()) }));
    out
}
pub fn dc_bias(audio: &[f64]) -> f64 { time_scalars(audio, 44100)[0] }                                   // audio_analysis_tools.rs:22
pub fn dbfs_max(audio: &[f64]) -> f64 { time_scalars(audio, 44100)[1] }                                  // :55
pub fn energy(audio: &[f64]) -> f64 { time_scalars(audio, 44100)[2] }                                    // :74
pub fn zero_crossing_rate(audio: &[f64], sample_rate: u32) -> f64 { time_scalars(audio, sample_rate)[3] } // :98

/// The four scalars above for every STFT frame of `audio` (new: rows aligned with `mp::stft_features`).
pub fn time_features_frames(audio: &[f64], fft_size: usize, hop_size: usize, sample_rate: u32) -> Vec<[f64; 4]> {
    let f = unsafe { aus_num_frames(audio.len(), fft_size, hop_size) };
    let (off, len) = ([0u64], [audio.len() as u64]);
    let mut out = vec![0.0f64; f * 4];
    with_ctx(|c| ok(unsafe {
        aus_time_features_frames(c, audio.as_ptr(), off.as_ptr(), len.as_ptr(), 1, fft_size, hop_size, sample_rate, out.as_mut_ptr())
    }));
    out.chunks_exact(4).map(|r| [r[0], r[1], r[2], r[3]]).collect()
}

/// Multi-GPU forms (no counterpart in the reference): a batch of files / one long file over `devices`, and the inverse of
/// several spectrograms (a stereo pair) by frame range with the host-side max of fft.rs:255-258 as the only exchange.
pub fn stft_features_multi(devices: &[i32], audio: &[f64], offsets: &[u64], lengths: &[u64], fft_size: usize, hop_size: usize,
                           window_type: WindowType, sample_rate: u32) -> Vec<analysis::Analysis> {
    let total: usize = lengths.iter().map(|&l| unsafe { aus_num_frames(l as usize, fft_size, hop_size) }).sum();
    let mut out = vec![0.0f64; total * 14];
    ok(unsafe {
        aus_stft_features_multi(devices.as_ptr(), devices.len() as c_int, audio.as_ptr(), offsets.as_ptr(), lengths.as_ptr(),
                                lengths.len(), fft_size, hop_size, wcode(window_type), sample_rate, F64, out.as_mut_ptr())
    });
    out.chunks_exact(14).map(analysis::from_row).collect()
}
pub fn irstft_multi(devices: &[i32], spectrograms: &[Vec<Vec<Complex<f64>>>], fft_size: usize, hop_size: usize,
                    window_type: WindowType) -> Result<Vec<Vec<f64>>, spectrum::SpectrumError> {
    let k = spectrograms[0][0].len();
    let nf: Vec<u64> = spectrograms.iter().map(|s| s.len() as u64).collect();
    let mut flat = Vec::with_capacity(nf.iter().sum::<u64>() as usize * k * 2);
    for s in spectrograms { for row in s { for c in row { flat.push(c.re); flat.push(c.im); } } }
    let lens: Vec<usize> = nf.iter().map(|&f| unsafe { aus_irstft_len(f as usize, fft_size, hop_size) }).collect();
    let mut out = vec![0.0f64; lens.iter().sum()];
    let rc = unsafe {
        aus_irstft_multi(devices.as_ptr(), devices.len() as c_int, flat.as_ptr(), nf.as_ptr(), nf.len(), k, fft_size, hop_size,
                         wcode(window_type), 1, out.as_mut_ptr())
    };
    if rc == AUS_ERR_SIZE { return Err(spectrum::SpectrumError { error_msg: last_error() }); }
    ok(rc);
    let mut res = Vec::new();
    let mut pos = 0;
    for l in lens { res.push(out[pos..pos + l].to_vec()); pos += l; }
    Ok(res)
}

#[allow(dead_code)]
fn _unused(_: *mut c_void, _: *const u64) { let _ = aus_irstft_batch; }
