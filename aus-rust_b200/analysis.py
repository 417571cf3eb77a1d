"""Mirror of src/analysis/{analyzer.rs, spectral_analysis_tools.rs:306-471}."""
from __future__ import annotations

from dataclasses import dataclass, fields

import numpy as np

from .api import default_context
from ._lib import FEATURE_NAMES


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

    @classmethod
    def from_row(cls, row) -> "Analysis":
        return cls(*[float(v) for v in row])

    def as_array(self) -> np.ndarray:
        return np.array([getattr(self, f.name) for f in fields(self)], dtype=np.float64)


assert tuple(f.name for f in fields(Analysis)) == FEATURE_NAMES


def analyzer(magnitude_spectrum, fft_size: int, sample_rate: int) -> Analysis:
    """analyzer.rs:41 -- all 14 features of one magnitude spectrum."""
    return Analysis.from_row(default_context().analyzer_batch(magnitude_spectrum, fft_size, sample_rate)[0])


def analyzer_batch(magnitude_spectrogram, fft_size: int, sample_rate: int) -> np.ndarray:
    """The loop at analyzer.rs:99-104 in one launch -> float64 [frames, 14]."""
    return default_context().analyzer_batch(magnitude_spectrogram, fft_size, sample_rate)


def make_power_spectrum(magnitude_spectrum) -> np.ndarray:
    """spectral_analysis_tools.rs:273 (trivial, host)."""
    return np.asarray(magnitude_spectrum, dtype=np.float64) ** 2.0


def _one(magnitude_spectrum, rfft_freqs, sample_rate=None) -> np.ndarray:
    """Single-feature wrappers take the bin frequencies as an argument; the device path needs them to be the
    rfftfreq grid f[k] = f0*k, which is what the reference's own callers pass (analyzer.rs:46)."""
    m = np.asarray(magnitude_spectrum, dtype=np.float64)
    fft_size = 2 * (m.shape[-1] - 1)
    if rfft_freqs is not None:
        f = np.asarray(rfft_freqs, dtype=np.float64)
        if f.shape[-1] != m.shape[-1]:
            raise IndexError("rfft_freqs and magnitude_spectrum differ in length (the reference panics)")
        sr = float(f[1]) * fft_size
        if not np.allclose(f, f[1] * np.arange(f.shape[-1]), rtol=1e-12, atol=0.0):
            raise ValueError("rfft_freqs must be the uniform rfftfreq grid")
        if sample_rate is None:
            sample_rate = int(round(sr))
        if abs(sr - sample_rate) > 1e-6 * max(1.0, sr):
            raise ValueError("rfft_freqs does not correspond to an integer sample rate")
    elif sample_rate is None:
        sample_rate = 44100   # features that do not depend on frequency
    return default_context().analyzer_batch(m, fft_size, int(sample_rate))[0]


def spectral_centroid(magnitude_spectrum, rfft_freqs) -> float:      # s_a_t.rs:306
    return float(_one(magnitude_spectrum, rfft_freqs)[0])


def spectral_variance(magnitude_spectrum, rfft_freqs) -> float:      # :466
    return float(_one(magnitude_spectrum, rfft_freqs)[1])


def spectral_skewness(magnitude_spectrum, rfft_freqs) -> float:      # :405
    return float(_one(magnitude_spectrum, rfft_freqs)[2])


def spectral_kurtosis(magnitude_spectrum, rfft_freqs) -> float:      # :362
    return float(_one(magnitude_spectrum, rfft_freqs)[3])


def spectral_entropy(magnitude_spectrum) -> float:                   # :324
    return float(_one(magnitude_spectrum, None)[4])


def spectral_flatness(magnitude_spectrum) -> float:                  # :343
    return float(_one(magnitude_spectrum, None)[5])


def spectral_roll_off_point(magnitude_spectrum, rfft_freqs, n: float) -> float:   # :385
    idx = {0.5: 6, 0.75: 7, 0.9: 8, 0.95: 9}.get(float(n))
    if idx is None:
        raise ValueError("the device path provides the analyzer's roll-off points 0.5 / 0.75 / 0.9 / 0.95 (analyzer.rs:53-56)")
    return float(_one(magnitude_spectrum, rfft_freqs)[idx])


def spectral_slope(magnitude_spectrum) -> float:                     # :426
    return float(_one(magnitude_spectrum, None)[10])


def spectral_slope_region(magnitude_spectrum, rfft_freqs, f_lower: float, f_upper: float, sample_rate: int) -> float:  # :447
    idx = {(0.0, 1000.0): 11, (1000.0, 5000.0): 12, (0.0, 5000.0): 13}.get((float(f_lower), float(f_upper)))
    if idx is None:
        raise ValueError("the device path provides the analyzer's three bands 0-1k / 1-5k / 0-5k Hz (analyzer.rs:60-62)")
    return float(_one(magnitude_spectrum, rfft_freqs, sample_rate)[idx])


# ---- SURVEY 8(f) "next" rows: analysis::autocorrelation and the time-domain scalars of audio_analysis_tools.rs
def autocorrelation(audio, fft_size: int) -> np.ndarray:
    """spectral_analysis_tools.rs:31 -- raises SpectrumError when len(audio) != fft_size."""
    return default_context().autocorrelation(np.asarray(audio, dtype=np.float64).reshape(-1), fft_size)


def _time(audio, sample_rate=44100):
    a = np.ascontiguousarray(audio, dtype=np.float64).reshape(-1)
    return default_context().time_features_batch(a, [0], [a.size], sample_rate)[0]


def dc_bias(audio) -> float:                          # audio_analysis_tools.rs:22
    return float(_time(audio)[0])


def dbfs_max(audio) -> float:                         # audio_analysis_tools.rs:55
    return float(_time(audio)[1])


def energy(audio) -> float:                           # audio_analysis_tools.rs:74
    return float(_time(audio)[2])


def zero_crossing_rate(audio, sample_rate: int) -> float:   # audio_analysis_tools.rs:98
    return float(_time(audio, sample_rate)[3])
