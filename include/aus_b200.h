/* aus_b200.h -- C ABI of the B200-native (sm_100a) STFT + spectral-feature path of `aus`.
 *
 * This is the drop-in boundary: the functions a Rust shim (aus-rust_b200/rust/, `extern "C"` block) or any
 * FFI (ctypes, cgo, JNI) binds in place of the reference's CPU code.  The reference has no FFI of its own;
 * its boundary is the crate's public Rust API (src/lib.rs:8-21), so each entry point below cites the Rust
 * function it replaces (paths relative to the reference's src/).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; nothing throws across the boundary.
 *   - every call returns AUS_OK (0) or a negative aus_status; aus_last_error() returns a thread-local
 *     message the shim can wrap into `SpectrumError { error_msg }` (spectrum/fft.rs:16-19).
 *   - outputs are caller-allocated, flat and row-major, sized with aus_num_frames():
 *       spectra   : frames x (fft_size/2+1) x {re, im}
 *       features  : frames x 14, in the declaration order of `Analysis` (analysis/analyzer.rs:10-23)
 *   - `precision` selects the arithmetic type on the device: AUS_F64 (reference-exact API, default) or AUS_F32
 *     (opt-in).  Host-pointer entry points always take and return f64 and convert on the device; `_dev` entry
 *     points take device buffers already in the compute type (double* / float*).
 *   - `_dev` entry points are stream-ordered on the context's stream and do NOT synchronise; host-pointer entry
 *     points copy H2D, run, copy D2H and return when the result is in the caller's buffer.
 *   - there is no CPU fallback: without a usable CUDA device every compute call fails with AUS_ERR_CUDA.
 *   - fft_size: powers of two in [64, 16384] run on the tuned radix-32 kernels; every other size in [2, 8192] -- rustfft
 *     plans any length, spectrum/fft.rs:120-121 -- runs as a chirp-z transform on top of them (same results to rounding,
 *     about 20x the work; the inverse needs an even size, like the reference).  Anything else is AUS_ERR_UNSUPPORTED.
 *   - 1 <= hop <= fft_size.  (The reference loops forever on hop == 0 and panics on hop > fft_size,
 *     spectrum/fft.rs:126-144; here both are AUS_ERR_INVALID.)
 */
#ifndef AUS_B200_H
#define AUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum aus_status {
    AUS_OK = 0,
    AUS_ERR_INVALID = -1,      /* bad argument (hop, fft_size, null pointer, band above Nyquist, ...) */
    AUS_ERR_CUDA = -2,         /* CUDA runtime error or no device */
    AUS_ERR_ALLOC = -3,        /* host or device allocation failed */
    AUS_ERR_SIZE = -4,         /* spectrum length mismatch (irfft / irstft `Err(SpectrumError)`, fft.rs:70-72, :217-219) */
    AUS_ERR_UNSUPPORTED = -5   /* valid for the reference but not built here (a non power-of-two fft_size above 8192) */
} aus_status;

/* window.rs:9-15 `WindowType`, same order */
typedef enum aus_window {
    AUS_WINDOW_BARTLETT = 0,
    AUS_WINDOW_BLACKMAN = 1,
    AUS_WINDOW_HANNING = 2,
    AUS_WINDOW_HAMMING = 3,
    AUS_WINDOW_RECTANGULAR = 4
} aus_window;

typedef enum aus_precision { AUS_F64 = 0, AUS_F32 = 1 } aus_precision;

#define AUS_NUM_FEATURES 14

/* analysis/analyzer.rs:9-24 `Analysis`, #[repr(C)] mirror (field order = declaration order) */
typedef struct aus_analysis {
    double spectral_centroid;
    double spectral_variance;
    double spectral_skewness;
    double spectral_kurtosis;
    double spectral_entropy;
    double spectral_flatness;
    double spectral_roll_off_50;
    double spectral_roll_off_75;
    double spectral_roll_off_90;
    double spectral_roll_off_95;
    double spectral_slope;
    double spectral_slope_0_1_khz;
    double spectral_slope_1_5_khz;
    double spectral_slope_0_5_khz;
} aus_analysis;

/* Opaque per-device context: device id, stream, twiddle/window tables, scratch.  Not shared between threads
 * concurrently; create one per caller thread (the reference's functions are stateless and re-entrant). */
typedef struct aus_ctx aus_ctx;

/* ---- library / context ------------------------------------------------------------------------------- */
const char *aus_version(void);
const char *aus_last_error(void);                       /* thread-local, never NULL */
int aus_device_count(void);                             /* < 0 on CUDA error */
int aus_ctx_create(int device, aus_ctx **out);
void aus_ctx_destroy(aus_ctx *ctx);
int aus_ctx_device(const aus_ctx *ctx);
void *aus_ctx_stream(const aus_ctx *ctx);               /* cudaStream_t, for callers that enqueue their own work */
int aus_ctx_set_stream(aus_ctx *ctx, void *cuda_stream);/* run on a caller-owned stream (e.g. torch's current stream) */
int aus_ctx_synchronize(aus_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t aus_ctx_launch_count(const aus_ctx *ctx);
/* input staging of the forward kernels for full, 16-byte aligned frames: 0 (default) = TMA bulk copy into the frame
 * group's shared buffer; 1 = coalesced 16-byte loads global -> registers (no shared staging).  Results are identical. */
int aus_ctx_set_direct_load(aus_ctx *ctx, int on);
/* forward-kernel variant: -1 (default) = chosen per fft_size / precision / outputs from B200 measurements,
 * 0 = shared-memory tail, 1 = register-resident tail (last FFT pass, real-FFT split and |X|^2 stay in registers; spectra
 * are stored straight from them), 2 = producer/consumer warps for f64 features at fft_size 2048 (FFT warps hand the power
 * row to feature warps; measured slower than mode 0, kept as a documented experiment).  Results agree to rounding (same
 * arithmetic, different summation grouping).  The environment variable AUS_B200_RR=0|1|2 sets the initial value. */
int aus_ctx_set_forward_variant(aus_ctx *ctx, int mode);
/* granularity (samples) of the host-pointer feature path's copy/compute pipeline; default 32 Mi samples (256 MB f64) */
int aus_ctx_set_chunk_samples(aus_ctx *ctx, uint64_t n_samples);
/* host-pointer feature calls on PAGEABLE input (mp.rs:26 hands over a `Vec<f64>`): 0 = give it to cudaMemcpyAsync (the
 * driver stages it, ~11 GB/s), 1 = page-lock it for the duration of the call (cudaHostRegister), 2 = copy it chunk by chunk
 * into the pipeline's own pinned buffers with several host threads while the previous chunk is in flight, -1 (default) =
 * mode 2 for inputs of at least 64 MB.  Already pinned inputs are used as they are (bench.py e2e vs e2e_pageable). */
int aus_ctx_set_host_register(aus_ctx *ctx, int mode);

/* ---- pure host helpers ------------------------------------------------------------------------------- */
/* frame count of the loop in spectrum/fft.rs:126-176 (strict `<` at :131; len <= fft_size -> 1 frame) */
size_t aus_num_frames(size_t n_samples, size_t fft_size, size_t hop);
/* irstft output length, spectrum/fft.rs:208: fft_size + hop * (frames - 1) */
size_t aus_irstft_len(size_t n_frames, size_t fft_size, size_t hop);
/* window.rs:110-118 `generate_window` (and the five generators at :26-98), f64, host */
int aus_generate_window(int window_type, size_t window_length, double *out);
/* spectrum/fft_tools.rs:199-206 `rfftfreq` */
int aus_rfftfreq(size_t fft_size, uint32_t sample_rate, double *out /* fft_size/2+1 */);

/* ---- fused STFT -> magnitude -> 14 features (kernel K1) ---------------------------------------------- */
/* replaces the loop rstft -> complex_to_polar_rstft -> analyzer per frame (analysis/analyzer.rs:99-104,
 * mp.rs:42-88) for a general hop and window. */
int aus_stft_features(aus_ctx *ctx, const double *audio, size_t n_samples, size_t fft_size, size_t hop,
                      int window_type, uint32_t sample_rate, int precision, double *features /* F x 14 */);
/* mp.rs:26 `stft_analysis(audio, fft_size, sample_rate, max_num_threads)`: hop = fft_size/2, Hamming (mp.rs:42).
 * `max_num_threads` is accepted and ignored (the GPU replaces the thread pool). */
int aus_stft_analysis(aus_ctx *ctx, const double *audio, size_t n_samples, size_t fft_size, uint32_t sample_rate,
                      size_t max_num_threads, aus_analysis *out /* F */);
/* batch of files in one launch: file i is audio[offsets[i] .. offsets[i]+lengths[i]); features of file i start at
 * row frame_starts[i] (frame_starts has n_files+1 entries, filled by the call; may be NULL). */
int aus_stft_features_batch(aus_ctx *ctx, const double *audio, const uint64_t *offsets, const uint64_t *lengths,
                            size_t n_files, size_t fft_size, size_t hop, int window_type, uint32_t sample_rate,
                            int precision, double *features, uint64_t *frame_starts);
/* device-resident variant: d_audio is T[...], d_features is T[total_frames x 14] (T = double or float per
 * `precision`); offsets/lengths are HOST arrays.  Optionally also writes spectra (d_spectra: T[frames x K x 2],
 * may be NULL).  Stream-ordered, no sync.  Alignment of device pointers (all *_dev entry points): sizeof(T) for real
 * arrays, 2 sizeof(T) for complex spectra (AUS_ERR_INVALID otherwise); 16-byte aligned bases take the fastest staging. */
int aus_stft_features_batch_dev(aus_ctx *ctx, const void *d_audio, const uint64_t *offsets, const uint64_t *lengths,
                                size_t n_files, size_t fft_size, size_t hop, int window_type, uint32_t sample_rate,
                                int precision, void *d_features, void *d_spectra);
/* SURVEY 8(f) item 4 -- decoded integer PCM straight from the codec: the samples are uploaded as they are (1, 2 or 4
 * bytes each) and widened on the device exactly like the reader does, `sample as f64 / INTMAX` with INTMAX = 2^(bits-1)-1
 * (audiofile.rs:12-15, :296-329).  sample_bits 8: `pcm` is int8_t[]; 16: int16_t[]; 24 or 32: int32_t[].  Same batch
 * layout and output as aus_stft_features_batch. */
int aus_stft_features_batch_pcm(aus_ctx *ctx, const void *pcm, int sample_bits, const uint64_t *offsets,
                                const uint64_t *lengths, size_t n_files, size_t fft_size, size_t hop, int window_type,
                                uint32_t sample_rate, int precision, double *features, uint64_t *frame_starts);
/* total frames of a batch (sum of aus_num_frames over files) */
uint64_t aus_batch_num_frames(const uint64_t *lengths, size_t n_files, size_t fft_size, size_t hop);

/* one file / batch sharded over several GPUs (contiguous frame ranges or whole files per device), host gather in
 * shard order like mp.rs:100-116.  devices == NULL -> devices 0..n_devices-1. */
int aus_stft_features_multi(const int *devices, int n_devices, const double *audio, const uint64_t *offsets,
                            const uint64_t *lengths, size_t n_files, size_t fft_size, size_t hop, int window_type,
                            uint32_t sample_rate, int precision, double *features);

/* the multi-device entry points (aus_stft_features_multi, aus_irstft_multi) cache one context per (device, slot) and hold
 * its lock for the duration of a call; this frees them (call it when no multi-device call is in flight). */
int aus_multi_shutdown(void);

/* pure host: the shard plan aus_stft_features_multi uses (also for callers that drive the devices themselves, e.g. one
 * process per GPU).  Virtual file v = samples [vf_off[v], vf_off[v]+vf_len[v]) of the caller's buffer; its frames are
 * rows [vf_row[v], ...) of the whole batch; vf_dev[v] in 0..n_devices-1.  A single long file is cut into contiguous
 * frame ranges overlapping by fft_size-hop samples; only the range ending at the true end of the file can produce the
 * short last frame (fft.rs:148-159).  *n_out = number of virtual files (<= n_files * n_devices). */
int aus_plan_shards(const uint64_t *offsets, const uint64_t *lengths, size_t n_files, size_t fft_size, size_t hop,
                    int n_devices, uint64_t *vf_off, uint64_t *vf_len, uint64_t *vf_row, int *vf_dev, size_t cap,
                    size_t *n_out);

/* ---- STFT with spectra written out (kernel K2) ------------------------------------------------------- */
/* spectrum/fft.rs:119 `rstft` */
int aus_rstft(aus_ctx *ctx, const double *audio, size_t n_samples, size_t fft_size, size_t hop, int window_type,
              int precision, double *spectra /* F x K x 2 */);
int aus_rstft_dev(aus_ctx *ctx, const void *d_audio, size_t n_samples, size_t fft_size, size_t hop,
                  int window_type, int precision, void *d_spectra);
/* spectrum/fft.rs:37 `rfft` (one frame, no window; audio has fft_size samples) */
int aus_rfft(aus_ctx *ctx, const double *audio, size_t fft_size, int precision, double *spectrum /* K x 2 */);

/* ---- polar conversion on device (spectrum/fft_tools.rs:21-55, :131-188) ------------------------------- */
int aus_complex_to_polar(aus_ctx *ctx, const double *spectra, size_t n_bins_total, double *magnitudes, double *phases);
int aus_polar_to_complex(aus_ctx *ctx, const double *magnitudes, const double *phases, size_t n_bins_total, double *spectra);
/* device-pointer variants (f64, stream-ordered, no sync; d_phases may be NULL in the forward direction): the edit loop of
 * spectrum/fft.rs:288-291 -- rstft -> complex_to_polar_rstft -> (edit) -> polar_to_complex_rstft -> irstft -- stays in HBM
 * between aus_rstft_dev and aus_irstft_dev. */
int aus_complex_to_polar_dev(aus_ctx *ctx, const void *d_spectra, size_t n_bins_total, void *d_magnitudes, void *d_phases);
int aus_polar_to_complex_dev(aus_ctx *ctx, const void *d_magnitudes, const void *d_phases, size_t n_bins_total, void *d_spectra);

/* ---- features from caller-supplied magnitude spectra (kernel K3) ------------------------------------- */
/* analysis/analyzer.rs:41 `analyzer` over a batch of rows; each row has fft_size/2+1 magnitudes */
int aus_analyzer_batch(aus_ctx *ctx, const double *magnitudes, size_t n_frames, size_t fft_size,
                       uint32_t sample_rate, int precision, double *features /* n_frames x 14 */);
int aus_analyzer_batch_dev(aus_ctx *ctx, const void *d_magnitudes, size_t n_frames, size_t fft_size,
                           uint32_t sample_rate, int precision, void *d_features);

/* analysis::spectral_centroid / _variance / _skewness / _kurtosis / _entropy / _flatness / _slope / _roll_off_point /
 * _slope_region (analysis/spectral_analysis_tools.rs:306-471) with the reference's full parameter space: rows of ANY
 * length n_bins, the caller's own rfft_freqs table (need not be uniform), any roll-off fractions and any slope bands
 * (band i = [band_lo[i], band_hi[i]] Hz at sample_rate).  One warp per row (coalesced lane-strided passes, fixed-order
 * tree sums); a roll-off threshold within rounding distance of a running sum is decided by the reference's own sequential
 * `cum += p[i] / P` arithmetic, so exact ties give the reference's index.  The single-feature wrappers' path, not the
 * hot path (aus_analyzer_batch is).
 * out: n_rows x (7 + n_roll + n_bands) = centroid, variance, skewness, kurtosis, entropy, flatness, slope, roll-offs, band
 * slopes.  A band whose upper edge reaches the last bin (the reference indexes out of bounds) gives AUS_ERR_INVALID. */
int aus_spectral_features_generic(aus_ctx *ctx, const double *magnitudes, size_t n_rows, size_t n_bins,
                                  const double *rfft_freqs, const double *roll_fractions, size_t n_roll,
                                  const double *band_lo, const double *band_hi, size_t n_bands, uint32_t sample_rate,
                                  double *out);

/* The public building blocks of analysis/spectral_analysis_tools.rs:79-290, which take the CALLER'S intermediates (sums,
 * pmf, centroid, variance) instead of deriving them -- a drop-in must honour whatever the caller passes.  One call = one
 * vector of n elements; a, b and the two scalars mean, per `what`:
 *   CENTROID      a = magnitudes, b = rfft_freqs, s0 = magnitude sum                      (:79-85)
 *   ENTROPY       a = pmf                                                                  (:93-99)
 *   FLATNESS      a = magnitudes, s0 = magnitude sum                                       (:107-114)
 *   KURTOSIS      a = pmf, b = rfft_freqs, s0 = centroid, s1 = variance                    (:122-129)
 *   ROLL_OFF      a = power spectrum, b = rfft_freqs, s0 = power sum, s1 = n (fraction)    (:137-145)
 *   SKEWNESS      a = pmf, b = rfft_freqs, s0 = centroid, s1 = variance                    (:153-160)
 *   SLOPE         a = power spectrum, s0 = power sum                                       (:167-182)
 *   SLOPE_REGION  a = power spectrum, b = rfft_freqs, s0 = f_lower, s1 = f_upper, sample_rate   (:190-230)
 *   VARIANCE      a = pmf, b = rfft_freqs, s0 = centroid                                   (:238-244)
 *   DOT_PRODUCT   a, b                                                                     (:253-259)
 *   POWER_SPECTRUM  a = magnitudes -> out[n]                                               (:273-279)
 *   SPECTRUM_PMF    a = power spectrum, s0 = power sum -> out[n]                           (:284-290)
 * out: one double (the last two kinds: n doubles).  b may be NULL where unused. */
typedef enum aus_compute_kind {
    AUS_COMPUTE_CENTROID = 0, AUS_COMPUTE_ENTROPY = 1, AUS_COMPUTE_FLATNESS = 2, AUS_COMPUTE_KURTOSIS = 3,
    AUS_COMPUTE_ROLL_OFF = 4, AUS_COMPUTE_SKEWNESS = 5, AUS_COMPUTE_SLOPE = 6, AUS_COMPUTE_SLOPE_REGION = 7,
    AUS_COMPUTE_VARIANCE = 8, AUS_COMPUTE_DOT_PRODUCT = 9, AUS_COMPUTE_POWER_SPECTRUM = 10, AUS_COMPUTE_SPECTRUM_PMF = 11
} aus_compute_kind;
int aus_spectral_compute(aus_ctx *ctx, int what, const double *a, const double *b, size_t n, double s0, double s1,
                         uint32_t sample_rate, double *out);

/* ---- inverse (kernels K4/K5) ------------------------------------------------------------------------- */
/* spectrum/fft_tools.rs:69 `overlap_add`: n_chunks rows of fft_size samples, row h added at offset h * hop, every output
 * sample summing its covering rows in ascending row order -> fft_size + hop (n_chunks - 1) samples.  f64. */
int aus_overlap_add(aus_ctx *ctx, const double *chunks, size_t n_chunks, size_t fft_size, size_t hop, double *audio);
int aus_overlap_add_dev(aus_ctx *ctx, const void *d_chunks, size_t n_chunks, size_t fft_size, size_t hop, void *d_audio);
/* spectrum/fft.rs:204 `irstft`: Hermitian extend, un-normalised inverse, window, overlap-add, window-norm divide,
 * peak-normalise to -6 dBFS.  frame_len must be fft_size/2+1, else AUS_ERR_SIZE with the reference's message.
 * normalize == 0 skips the peak normalisation (new knob; the reference always normalises -> pass 1). */
int aus_irstft(aus_ctx *ctx, const double *spectra, size_t n_frames, size_t frame_len, size_t fft_size, size_t hop,
               int window_type, int precision, int normalize, double *audio /* aus_irstft_len */);
int aus_irstft_dev(aus_ctx *ctx, const void *d_spectra, size_t n_frames, size_t fft_size, size_t hop,
                   int window_type, int precision, int normalize, void *d_audio);
/* several spectrograms of the same fft_size / hop / window in one launch (config 2's stereo pair): rows and output signals
 * are concatenated in order; spectrogram i has n_frames[i] rows and yields aus_irstft_len(n_frames[i]) samples, each
 * peak-normalised on its own (fft.rs:255-265). */
int aus_irstft_batch(aus_ctx *ctx, const double *spectra, const uint64_t *n_frames, size_t n_spectrograms, size_t frame_len,
                     size_t fft_size, size_t hop, int window_type, int precision, int normalize, double *audio);
int aus_irstft_batch_dev(aus_ctx *ctx, const void *d_spectra, const uint64_t *n_frames, size_t n_spectrograms, size_t fft_size,
                         size_t hop, int window_type, int precision, int normalize, void *d_audio);
/* irstft sharded over several GPUs of one box: contiguous frame ranges (whole spectrograms when there are at least as many
 * as devices); a range recomputes the ceil(fft_size/hop)-1 frames before it instead of exchanging partial sums, and the
 * only exchange is the host-side max over the ranges' peaks before the scale pass (fft.rs:255-265).  f64.
 * devices == NULL -> devices 0..n_devices-1.  Same layout as aus_irstft_batch. */
int aus_irstft_multi(const int *devices, int n_devices, const double *spectra, const uint64_t *n_frames, size_t n_spectrograms,
                     size_t frame_len, size_t fft_size, size_t hop, int window_type, int normalize, double *audio);
/* spectrum/fft.rs:66 `irfft` (output is fft_size x the signal: rustfft does not normalise) */
int aus_irfft(aus_ctx *ctx, const double *spectrum, size_t spectrum_len, size_t fft_size, int precision,
              double *audio /* fft_size */);

/* ---- SURVEY 8(f) "next" rows ------------------------------------------------------------------------ */
/* analysis/spectral_analysis_tools.rs:31-70 `autocorrelation` (audio_len != fft_size -> AUS_ERR_SIZE with the reference's
 * message); the batch form takes n_chunks consecutive chunks of fft_size samples.  2*fft_size must be <= 16384. */
int aus_autocorrelation(aus_ctx *ctx, const double *audio, size_t audio_len, size_t fft_size, double *out /* fft_size */);
int aus_autocorrelation_batch(aus_ctx *ctx, const double *audio, size_t n_chunks, size_t fft_size, double *out);
/* analysis/audio_analysis_tools.rs:22-108 over every file of a batch: out[f] = {dc_bias, dbfs_max, energy,
 * zero_crossing_rate} */
int aus_time_features_batch(aus_ctx *ctx, const double *audio, const uint64_t *offsets, const uint64_t *lengths,
                            size_t n_files, uint32_t sample_rate, double *out /* n_files x 4 */);

/* the same four scalars of every STFT frame of a batch -- the slice audio[h*hop .. min(h*hop + fft_size, len)) that the
 * frame loop of spectrum/fft.rs:126-176 cuts -- in the row order of the feature matrix: out[frame] = {dc_bias, dbfs_max,
 * energy, zero_crossing_rate}. */
int aus_time_features_frames(aus_ctx *ctx, const double *audio, const uint64_t *offsets, const uint64_t *lengths, size_t n_files,
                             size_t fft_size, size_t hop, uint32_t sample_rate, double *out /* total_frames x 4 */);
int aus_time_features_frames_dev(aus_ctx *ctx, const void *d_audio, const uint64_t *offsets, const uint64_t *lengths,
                                 size_t n_files, size_t fft_size, size_t hop, uint32_t sample_rate, void *d_out);
/* SURVEY 8(f) item 4, input side: decoded integer PCM -> f64 channels as the reader does (audiofile.rs:296-329,
 * `sample as f64 / INTMAX`; sample_bits 8 / 16 / 24 / 32 as above), from interleaved ([frame][channel]) or planar
 * ([channel][frame]) samples, and `mixdown` (audiofile.rs:168-178) fused in when mixdown != 0: channel 0 accumulates
 * every channel INCLUDING ITSELF, in channel order, then divides by the channel count (the reference's arithmetic, bit
 * for bit; a single channel is left untouched).  out: planar [n_channels][n_frames], or [n_frames] after mixdown. */
int aus_pcm_to_f64(aus_ctx *ctx, const void *pcm, int sample_bits, size_t n_frames, int n_channels, int interleaved,
                   int mixdown, double *out);
int aus_pcm_to_f64_dev(aus_ctx *ctx, const void *d_pcm, int sample_bits, size_t n_frames, int n_channels, int interleaved,
                       int mixdown, void *d_out);
/* audiofile.rs:168-178 `mixdown` on f64 channels, planar [n_channels][n_frames] -> [n_frames] */
int aus_mixdown(aus_ctx *ctx, const double *channels, size_t n_frames, int n_channels, double *out);
int aus_mixdown_dev(aus_ctx *ctx, const void *d_channels, size_t n_frames, int n_channels, void *d_out);

/* ---- synthetic input generator used by the benchmark (synthesis.rs:50-59 sine + counter noise) --- */
int aus_synth_dev(aus_ctx *ctx, void *d_audio, uint64_t n_samples, uint64_t n0, uint32_t sample_rate, double freq,
                  uint64_t seed, uint64_t file_id, int precision);

#ifdef __cplusplus
}
#endif
#endif /* AUS_B200_H */
