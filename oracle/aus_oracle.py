"""CPU oracle for the aus STFT + spectral-feature path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker.

PARITY STATUS: **parity unpinned.**  The reference (fleximeter/aus-rust, ``aus`` 0.1.7) is Rust;
there is no cargo/rustc in this image, its FFT lives in the un-vendored crate ``rustfft ^6.2.0``
(Cargo.toml:20, no Cargo.lock), and none of its tests hold a golden vector for this path
(fft.rs:277-299, mp.rs:127-136, analyzer.rs:88-105 read absent WAV files and assert nothing).
This file therefore restates the reference's arithmetic line by line and is pinned only against
(a) analytic known-answer tests and (b) an mpmath 50-digit evaluation of the same formulas
(tests/test_oracle.py).  rustfft's documented contract is used for the DFT: forward sign
e^{-2*pi*i*jk/N}, no normalisation in either direction.

Every function cites the reference lines it follows (paths relative to /root/reference/src).
Sums are sequential in ascending index like the reference's ``for`` loops (np.cumsum is a
sequential left-to-right accumulation; np.sum is pairwise and is NOT used for those).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, fields

import numpy as np

# window.rs:9-15 -- enum order is the C-ABI int code.
BARTLETT, BLACKMAN, HANNING, HAMMING, RECTANGULAR = range(5)
WINDOW_NAMES = ("bartlett", "blackman", "hanning", "hamming", "rectangular")

FEATURE_NAMES = (
    "spectral_centroid", "spectral_variance", "spectral_skewness", "spectral_kurtosis",
    "spectral_entropy", "spectral_flatness", "spectral_roll_off_50", "spectral_roll_off_75",
    "spectral_roll_off_90", "spectral_roll_off_95", "spectral_slope", "spectral_slope_0_1_khz",
    "spectral_slope_1_5_khz", "spectral_slope_0_5_khz",
)  # analyzer.rs:10-23 declaration order


def _seqsum(a) -> float:
    """Sequential left-to-right f64 sum (Rust ``iter().sum()`` / ``sum += x`` loops)."""
    a = np.asarray(a, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.cumsum(a)[-1])


# --------------------------------------------------------------------------- window.rs
def generate_window(window_type: int, window_length: int) -> np.ndarray:
    """window.rs:26-118.  Expressions are evaluated in the reference's operation order."""
    L = int(window_length)
    i = np.arange(L, dtype=np.float64)
    Lf = np.float64(L)
    with np.errstate(divide="ignore", invalid="ignore"):
        if window_type == BARTLETT:      # window.rs:29
            return (2.0 / (Lf - 1.0)) * ((Lf - 1.0) / 2.0 - np.abs(i - (Lf - 1.0) / 2.0))
        if window_type == BLACKMAN:      # window.rs:46-47 (periodic: denominator L)
            return 0.42 - 0.5 * np.cos((2.0 * math.pi * i) / Lf) + 0.08 * np.cos((4.0 * math.pi * i) / Lf)
        if window_type == HANNING:       # window.rs:64
            return 0.5 - 0.5 * np.cos((2.0 * math.pi * i) / (Lf - 1.0))
        if window_type == HAMMING:       # window.rs:81
            return 0.54 - 0.46 * np.cos((2.0 * math.pi * i) / (Lf - 1.0))
        if window_type == RECTANGULAR:   # window.rs:96
            return np.ones(L, dtype=np.float64)
    raise ValueError("unknown window type")


# --------------------------------------------------------------------------- spectrum/fft.rs
def rfft(audio, fft_size: int) -> np.ndarray:
    """fft.rs:37-47: complex N-point forward DFT of the real input, bins 0..=N/2, un-normalised.
    (The reference transforms ``audio.len()`` points; callers pass len == fft_size.)"""
    x = np.asarray(audio, dtype=np.float64)
    return np.fft.fft(x.astype(np.complex128))[: fft_size // 2 + 1].copy()


class SpectrumError(Exception):
    """fft.rs:16-19."""


def irfft(spectrum, fft_size: int) -> np.ndarray:
    """fft.rs:66-93: Hermitian-extend, un-normalised inverse (N x the signal), keep .re."""
    s = np.asarray(spectrum, dtype=np.complex128)
    K = fft_size // 2 + 1
    if s.shape[0] != K:
        raise SpectrumError(f"The spectrum is the wrong size. It should be {K}, but it is actually {s.shape[0]}.")
    full = np.concatenate([s, np.conj(s[fft_size // 2 - 1:0:-1])])   # fft.rs:76-83
    return (np.fft.ifft(full) * fft_size).real.copy()


def num_frames(n_samples: int, fft_size: int, hop_size: int) -> int:
    """Frame count of the loop at fft.rs:126-176 (strict '<' at :131)."""
    if n_samples <= fft_size:
        return 1
    return -(-(n_samples - fft_size) // hop_size) + 1


def stft_frames(audio, fft_size: int, hop_size: int, window_type: int) -> np.ndarray:
    """Windowed, zero-padded frames exactly as built at fft.rs:128-168 (before the FFT)."""
    x = np.asarray(audio, dtype=np.float64)
    n = x.shape[0]
    if hop_size <= 0 or hop_size > fft_size:
        raise ValueError("hop_size must be in 1..=fft_size (the reference hangs / panics otherwise)")
    window = generate_window(window_type, fft_size)          # fft.rs:123
    F = num_frames(n, fft_size, hop_size)
    out = np.zeros((F, fft_size), dtype=np.float64)
    for h in range(F):
        start = h * hop_size
        end = start + fft_size if start + fft_size < n else n  # fft.rs:130-135
        if end == n:
            assert h == F - 1
        L = end - start
        if L < fft_size:                                     # num_zeros > 0, fft.rs:148-159
            out[h, :L] = x[start:end] * generate_window(window_type, L)
        else:
            out[h, :] = x[start:end] * window
    return out


def rstft(audio, fft_size: int, hop_size: int, window_type: int) -> np.ndarray:
    """fft.rs:119-178 -> complex128 [F, N/2+1]."""
    fr = stft_frames(audio, fft_size, hop_size, window_type)
    return np.fft.fft(fr.astype(np.complex128), axis=1)[:, : fft_size // 2 + 1].copy()


def overlap_add(chunks: np.ndarray, fft_size: int, hop_size: int) -> np.ndarray:
    """fft_tools.rs:69-112: per output sample, add covering frames in ascending frame order."""
    F = chunks.shape[0]
    out = np.zeros(fft_size + hop_size * (F - 1), dtype=np.float64)
    for h in range(F):           # ascending h == the reference's lower..=upper inner loop order
        out[h * hop_size: h * hop_size + fft_size] += chunks[h]
    return out


def irstft(spectrogram, fft_size: int, hop_size: int, window_type: int, normalize: bool = True) -> np.ndarray:
    """fft.rs:204-268 including the -6 dBFS peak normalisation and its NaN/inf quirks.
    ``normalize=False`` (not in the reference) stops after the window-norm division (fft.rs:250-252)."""
    S = np.asarray(spectrogram, dtype=np.complex128)
    if S.ndim != 2 or S.shape[0] == 0:
        raise SpectrumError("empty spectrogram (the reference underflows at fft.rs:208)")
    K = fft_size // 2 + 1
    if S.shape[1] != K:
        raise SpectrumError(f"The spectrum at frame 0 has an incorrect length. It should be {K}, but it is actually {S.shape[1]}.")
    F = S.shape[0]
    w = generate_window(window_type, fft_size)
    full = np.concatenate([S, np.conj(S[:, fft_size // 2 - 1:0:-1])], axis=1)   # fft.rs:221-229
    y = (np.fft.ifft(full, axis=1) * fft_size).real                           # fft.rs:231
    chunks = y * w[None, :]                                                   # fft.rs:234-236
    norm = np.zeros(fft_size + hop_size * (F - 1), dtype=np.float64)
    w2 = w ** 2.0
    for h in range(F):                                                        # fft.rs:239-243
        norm[h * hop_size: h * hop_size + fft_size] += w2
    audio = overlap_add(chunks, fft_size, hop_size)                           # fft.rs:247
    with np.errstate(divide="ignore", invalid="ignore"):
        audio = audio / norm                                                  # fft.rs:250-252
        if not normalize:
            return audio
        maxval = 0.0
        m = np.fmax.reduce(audio, initial=0.0) if audio.size else 0.0         # f64::max ignores NaN, fft.rs:255-258
        maxval = float(m)
        scaler = np.float64(10.0 ** (-6.0 / 20.0)) / np.float64(maxval)       # fft.rs:261-262
        return audio * scaler


# --------------------------------------------------------------------------- spectrum/fft_tools.rs
def complex_to_polar_rstft(spectrogram):
    """fft_tools.rs:41-55 (and :21-29 for one frame)."""
    S = np.asarray(spectrogram, dtype=np.complex128)
    mag = np.sqrt(S.real ** 2.0 + S.imag ** 2.0)
    ph = np.arctan2(S.imag, S.real)
    return mag, ph


complex_to_polar_rfft = complex_to_polar_rstft


def polar_to_complex_rstft(mag, ph):
    """fft_tools.rs:131-188."""
    mag = np.asarray(mag, dtype=np.float64)
    ph = np.asarray(ph, dtype=np.float64)
    if mag.shape != ph.shape:
        raise SpectrumError("The magnitude spectrum and phase spectrum do not have the same length.")
    return (np.cos(ph) * mag) + 1j * (np.sin(ph) * mag)


polar_to_complex_rfft = polar_to_complex_rstft


def rfftfreq(fft_size: int, sample_rate: int) -> np.ndarray:
    """fft_tools.rs:199-206: f0 = sr/N, f[i] = f0*i, f[0] = 0.0."""
    f0 = np.float64(sample_rate) / np.float64(fft_size)
    f = f0 * np.arange(fft_size // 2 + 1, dtype=np.float64)
    f[0] = 0.0
    return f


# --------------------------------------------------------------------------- analysis/spectral_analysis_tools.rs
def make_power_spectrum(m):                       # :273-279
    return np.asarray(m, dtype=np.float64) ** 2.0


def make_spectrum_pmf(p, p_sum):                  # :284-290
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.asarray(p, dtype=np.float64) / np.float64(p_sum)


def compute_spectral_centroid(m, freqs, m_sum):   # :79-85
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(_seqsum(m * freqs)) / np.float64(m_sum))


def compute_spectral_entropy(pmf):                # :93-99
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(-np.float64(_seqsum(pmf * np.log2(pmf))))


def compute_spectral_flatness(m, m_sum):          # :107-114
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.float64(_seqsum(np.log(m))) / np.float64(len(m))
        return float(np.exp(s) * np.float64(len(m)) / np.float64(m_sum))


def compute_spectral_variance(pmf, freqs, c):     # :238-244
    with np.errstate(invalid="ignore", over="ignore"):
        return _seqsum((freqs - c) ** 2.0 * pmf)


def compute_spectral_skewness(pmf, freqs, c, var):  # :153-160
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        s = np.float64(_seqsum((freqs - c) ** 3.0 * pmf))
        return float(s / np.float64(var) ** 1.5)


def compute_spectral_kurtosis(pmf, freqs, c, var):  # :122-129
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        s = np.float64(_seqsum((freqs - c) ** 4.0 * pmf))
        return float(s / np.float64(var) ** 2.0)


def compute_spectral_roll_off_point(p, freqs, p_sum, n):  # :137-145
    if not n > 0.0:
        raise IndexError("n <= 0 indexes rfft_freqs[-1] in the reference (panic)")
    with np.errstate(divide="ignore", invalid="ignore"):
        cum = np.cumsum(np.asarray(p, dtype=np.float64) / np.float64(p_sum))
    K = len(freqs)
    # loop: i=-1; while cum < n and i < K-1: i += 1; cum += p[i]/P  ->  first i with cum_i >= n
    # (a NaN cum also stops the loop, leaving i at the first NaN index), capped at K-1.
    stop = ~(cum < n)
    idx = int(np.argmax(stop)) if stop.any() else K - 1
    return float(freqs[idx])


def compute_spectral_slope(p, p_sum):             # :167-182
    n = np.float64(len(p))
    x = np.arange(len(p), dtype=np.float64)
    sum_x = n * (n - 1.0) / 2.0
    sum_x_2 = n * (n - 1.0) * (2.0 * n - 1.0) / 6.0
    with np.errstate(invalid="ignore"):
        return float((n * np.float64(_seqsum(p * x)) - sum_x * np.float64(p_sum)) / (n * sum_x_2 - sum_x ** 2.0))


def compute_spectral_slope_region(p, freqs, f_lower, f_upper, sample_rate):  # :190-230
    p = np.asarray(p, dtype=np.float64)
    K = len(p)
    fund = np.float64(sample_rate) / (np.float64(K - 1) * 2.0)
    m_fl = np.float64(f_lower) / fund
    m_fu = np.float64(f_upper) / fund
    n = np.float64(K)
    a_c, a_f = int(math.ceil(m_fl)), int(math.floor(m_fl))
    b_c, b_f = int(math.ceil(m_fu)), int(math.floor(m_fu))
    if b_c >= K or a_c > b_f:
        raise IndexError("band outside the spectrum (the reference panics on the slice / index)")
    fr, pr = freqs[a_c:b_f], p[a_c:b_f]
    with np.errstate(invalid="ignore", over="ignore"):
        yl = p[a_f] + (m_fl - a_f) * (p[a_c] - p[a_f])
        yu = p[b_f] + (m_fu - b_f) * (p[b_c] - p[b_f])
        sum_x = np.float64(f_lower) + np.float64(_seqsum(fr)) + np.float64(f_upper)       # :208
        sum_y = yl + np.float64(_seqsum(pr)) + yu                                          # :211-215
        sum_x_2 = np.float64(f_lower) ** 2.0 + np.float64(_seqsum(fr * fr)) + np.float64(f_upper) ** 2.0  # :218-221
        sum_xy = np.float64(f_lower) * yl + np.float64(_seqsum(pr * fr)) + np.float64(f_upper) * yu      # :224-226
        return float((n * sum_xy - sum_x * sum_y) / (n * sum_x_2 - sum_x ** 2.0))


# public one-feature wrappers, spectral_analysis_tools.rs:306-471
def spectral_centroid(m, freqs):
    m = np.asarray(m, dtype=np.float64)
    return compute_spectral_centroid(m, freqs, _seqsum(m))


def _pmf_of(m):
    p = make_power_spectrum(m)
    return p, make_spectrum_pmf(p, _seqsum(p))


def spectral_entropy(m):
    return compute_spectral_entropy(_pmf_of(m)[1])


def spectral_flatness(m):
    m = np.asarray(m, dtype=np.float64)
    return compute_spectral_flatness(m, _seqsum(m))


def spectral_variance(m, freqs):
    m = np.asarray(m, dtype=np.float64)
    _, pmf = _pmf_of(m)
    return compute_spectral_variance(pmf, freqs, spectral_centroid(m, freqs))


def spectral_skewness(m, freqs):
    m = np.asarray(m, dtype=np.float64)
    _, pmf = _pmf_of(m)
    c = spectral_centroid(m, freqs)
    return compute_spectral_skewness(pmf, freqs, c, compute_spectral_variance(pmf, freqs, c))


def spectral_kurtosis(m, freqs):
    m = np.asarray(m, dtype=np.float64)
    _, pmf = _pmf_of(m)
    c = spectral_centroid(m, freqs)
    return compute_spectral_kurtosis(pmf, freqs, c, compute_spectral_variance(pmf, freqs, c))


def spectral_roll_off_point(m, freqs, n):
    p = make_power_spectrum(m)
    return compute_spectral_roll_off_point(p, freqs, _seqsum(p), n)


def spectral_slope(m):
    p = make_power_spectrum(m)
    return compute_spectral_slope(p, _seqsum(p))


def spectral_slope_region(m, freqs, f_lower, f_upper, sample_rate):
    return compute_spectral_slope_region(make_power_spectrum(m), freqs, f_lower, f_upper, sample_rate)


# --------------------------------------------------------------------------- analysis/analyzer.rs
@dataclass
class Analysis:
    """analyzer.rs:9-24."""
    spectral_centroid: float
    spectral_variance: float
    spectral_skewness: float
    spectral_kurtosis: float
    spectral_entropy: float
    spectral_flatness: float
    spectral_roll_off_50: float
    spectral_roll_off_75: float
    spectral_roll_off_90: float
    spectral_roll_off_95: float
    spectral_slope: float
    spectral_slope_0_1_khz: float
    spectral_slope_1_5_khz: float
    spectral_slope_0_5_khz: float

    def as_array(self) -> np.ndarray:
        return np.array([getattr(self, f.name) for f in fields(self)], dtype=np.float64)


def analyzer(magnitude_spectrum, fft_size: int, sample_rate: int) -> Analysis:
    """analyzer.rs:41-81."""
    m = np.asarray(magnitude_spectrum, dtype=np.float64)
    p = make_power_spectrum(m)
    m_sum = _seqsum(m)
    p_sum = _seqsum(p)
    pmf = make_spectrum_pmf(p, p_sum)
    freqs = rfftfreq(fft_size, sample_rate)
    c = compute_spectral_centroid(m, freqs, m_sum)
    var = compute_spectral_variance(pmf, freqs, c)
    return Analysis(
        spectral_centroid=c,
        spectral_variance=var,
        spectral_skewness=compute_spectral_skewness(pmf, freqs, c, var),
        spectral_kurtosis=compute_spectral_kurtosis(pmf, freqs, c, var),
        spectral_entropy=compute_spectral_entropy(pmf),
        spectral_flatness=compute_spectral_flatness(m, m_sum),
        spectral_roll_off_50=compute_spectral_roll_off_point(p, freqs, p_sum, 0.5),
        spectral_roll_off_75=compute_spectral_roll_off_point(p, freqs, p_sum, 0.75),
        spectral_roll_off_90=compute_spectral_roll_off_point(p, freqs, p_sum, 0.9),
        spectral_roll_off_95=compute_spectral_roll_off_point(p, freqs, p_sum, 0.95),
        spectral_slope=compute_spectral_slope(p, p_sum),
        spectral_slope_0_1_khz=compute_spectral_slope_region(p, freqs, 0.0, 1000.0, sample_rate),
        spectral_slope_1_5_khz=compute_spectral_slope_region(p, freqs, 1000.0, 5000.0, sample_rate),
        spectral_slope_0_5_khz=compute_spectral_slope_region(p, freqs, 0.0, 5000.0, sample_rate),
    )


def analyze_spectrogram(mag: np.ndarray, fft_size: int, sample_rate: int) -> np.ndarray:
    """The loop at analyzer.rs:99-104 / mp.rs:86-88 -> float64 [F, 14] in FEATURE_NAMES order."""
    return np.stack([analyzer(row, fft_size, sample_rate).as_array() for row in np.asarray(mag)])


def stft_features(audio, fft_size: int, hop_size: int, window_type: int, sample_rate: int) -> np.ndarray:
    """rstft -> complex_to_polar_rstft -> analyzer per frame (the general-hop/window form of mp.rs)."""
    mag, _ = complex_to_polar_rstft(rstft(audio, fft_size, hop_size, window_type))
    return analyze_spectrogram(mag, fft_size, sample_rate)


# --------------------------------------------------------------------------- mp.rs
def stft_analysis(audio, fft_size: int, sample_rate: int, max_num_threads=None) -> np.ndarray:
    """mp.rs:26-119: hop = fft_size/2, Hamming (mp.rs:42); output in frame order.  The thread pool only
    changes who computes which contiguous chunk (mp.rs:47-57), never the values or their order."""
    return stft_features(audio, fft_size, fft_size // 2, HAMMING, sample_rate)


# --------------------------------------------------------------------------- SURVEY 8(f) "next" rows
def autocorrelation(audio, fft_size: int) -> np.ndarray:
    """spectral_analysis_tools.rs:31-70: N/2 zeros left and right, 2N-point FFT, |X|^2, un-normalised inverse,
    samples [N, 2N)."""
    a = np.asarray(audio, dtype=np.float64)
    if a.shape[0] != fft_size:
        raise SpectrumError("The audio length does not match the FFT size.")
    padded = np.concatenate([np.zeros(fft_size // 2), a, np.zeros(fft_size // 2)]).astype(np.complex128)
    spec = np.fft.fft(padded)
    spec = spec * np.conj(spec)
    back = np.fft.ifft(spec) * (2 * fft_size)
    return back[fft_size:2 * fft_size].real.copy()


DBFS_EPSILON = 1e-20          # audio_analysis_tools.rs:10


def dc_bias(audio) -> float:
    """audio_analysis_tools.rs:22-28."""
    a = np.asarray(audio, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(np.float64(_seqsum(a)) / np.float64(a.size))


def dbfs(val: float, epsilon: float = DBFS_EPSILON) -> float:
    """audio_analysis_tools.rs:42-48 (note: log10 of the signed value, as written)."""
    if abs(val) < epsilon:
        return -math.inf
    return 20.0 * math.log10(val) if val > 0 else math.nan


def dbfs_max(audio) -> float:
    """audio_analysis_tools.rs:55-65."""
    a = np.asarray(audio, dtype=np.float64)
    m = 0.0
    if a.size:
        ab = np.abs(a)
        ab = ab[~np.isnan(ab)]          # `sample_abs > maxval` is false for NaN
        m = float(ab.max()) if ab.size else 0.0
    return dbfs(m, DBFS_EPSILON)


def energy(audio) -> float:
    """audio_analysis_tools.rs:74-84."""
    a = np.asarray(audio, dtype=np.float64)
    if a.size < 1:
        return 0.0
    return float(np.sqrt(1.0 / a.size * np.float64(_seqsum(a ** 2.0))))


def zero_crossing_rate(audio, sample_rate: int) -> float:
    """audio_analysis_tools.rs:98-108."""
    a = np.asarray(audio, dtype=np.float64)
    n = a.size
    if n == 0:
        return math.nan                  # 0 * sr / 0
    prod_neg = a[:-1] * a[1:] < 0.0                      # i = 1 .. n-1
    zero_then_neg = np.zeros(n - 1, dtype=bool)
    if n > 2:
        zero_then_neg[:-1] = (a[1:-1] == 0.0) & (a[2:] < 0.0)     # i < n-1 and a[i+1] < 0 and a[i] == 0
    num = int(np.count_nonzero(prod_neg | (~prod_neg & zero_then_neg)))
    return float(num) * float(sample_rate) / float(n)


# ------------------------------------------------------------------------------------------------ audiofile.rs (input side)
INTMAX = {8: float(2 ** 7 - 1), 16: float(2 ** 15 - 1), 24: float(2 ** 23 - 1), 32: float(2 ** 31 - 1)}   # audiofile.rs:12-15


def pcm_to_f64(pcm, sample_bits: int) -> np.ndarray:
    """audiofile.rs:296-329: ``sample as f64 / INTMAX`` (IEEE division), any array shape."""
    return np.asarray(pcm).astype(np.float64) / INTMAX[int(sample_bits)]


def mixdown(channels) -> np.ndarray:
    """audiofile.rs:168-178 on ``[n_channels][n_frames]``: with more than one channel, channel 0 accumulates EVERY channel
    in channel order -- itself included, so it is counted twice -- and is then divided by the channel count."""
    ch = np.asarray(channels, dtype=np.float64)
    if ch.ndim == 1:
        ch = ch[None, :]
    out = ch[0].copy()
    if ch.shape[0] > 1:                                   # :169
        for c in range(ch.shape[0]):                      # :171-173 (c = 0 adds channel 0 to itself)
            out = out + ch[c]
        out = out / float(ch.shape[0])                    # :174
    return out


def time_features_frames(audio, fft_size: int, hop_size: int, sample_rate: int) -> np.ndarray:
    """dc_bias, dbfs_max, energy, zero_crossing_rate (audio_analysis_tools.rs:22-108) of every slice the frame loop of
    fft.rs:126-176 cuts: ``audio[h*hop : min(h*hop + fft_size, len)]`` -> ``[frames, 4]``."""
    x = np.asarray(audio, dtype=np.float64)
    F = num_frames(x.size, fft_size, hop_size)
    out = np.empty((F, 4))
    with np.errstate(all="ignore"):
        for h in range(F):
            seg = x[h * hop_size: h * hop_size + fft_size]
            out[h] = (dc_bias(seg) if seg.size else np.nan, dbfs_max(seg), energy(seg),
                      zero_crossing_rate(seg, sample_rate) if seg.size else np.nan)
    return out

# --------------------------------------------------------------------------- synthetic inputs
_M64 = (1 << 64) - 1


def counter_uniform(seed: int, file_id: int, n0: int, count: int) -> np.ndarray:
    """u[n] in [-1, 1): splitmix64 finaliser of a (seed, file_id, n) counter; bit-identical to the
    device generator in csrc (aus_synth kernel) and to oracle/aus_oracle.c."""
    n = np.arange(n0, n0 + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(file_id) * np.uint64(0xD1B54A32D192ED03)
             + n * np.uint64(0x94D049BB133111EB) + np.uint64(0x2545F4914F6CDD1D))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0


def synth_signal(n_samples: int, sample_rate: int = 44100, freq: float = 440.0, seed: int = 0,
                 file_id: int = 0, n0: int = 0) -> np.ndarray:
    """x[n] = 0.5*sin(c*n) + 0.1*u[n], c = 2*pi*f/sr as in synthesis.rs:50-59 (initial phase 0)."""
    n = np.arange(n0, n0 + n_samples, dtype=np.float64)
    sine_coef = 2.0 * math.pi * freq / float(sample_rate)
    return 0.5 * np.sin(sine_coef * n) + 0.1 * counter_uniform(seed, file_id, n0, n_samples)


def file_freq(file_id: int) -> float:
    """Per-file sine frequency for corpus runs: 110 Hz * 2^(k/12), k = file_id mod 61 (110 Hz .. 3.5 kHz)."""
    return 110.0 * 2.0 ** ((file_id % 61) / 12.0)
