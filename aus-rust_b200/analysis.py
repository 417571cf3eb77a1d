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


def _one(magnitude_spectrum, rfft_freqs, roll=(), bands=(), sample_rate=None) -> np.ndarray:
    """One row through the generic device path (any length, any frequency table, any parameters):
    [centroid, variance, skewness, kurtosis, entropy, flatness, slope, roll-offs..., band slopes...]."""
    m = np.asarray(magnitude_spectrum, dtype=np.float64).reshape(-1)
    if rfft_freqs is None:
        f = np.arange(m.size, dtype=np.float64)          # features that do not depend on the frequency table
    else:
        f = np.asarray(rfft_freqs, dtype=np.float64).reshape(-1)
        if f.size != m.size:
            raise IndexError("rfft_freqs and magnitude_spectrum differ in length (the reference panics)")
    return default_context().spectral_features_generic(m, f, roll, bands, sample_rate if sample_rate is not None else 44100)[0]


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
    """Any roll-off fraction 0 < n <= 1 and any frequency table."""
    return float(_one(magnitude_spectrum, rfft_freqs, roll=(float(n),))[7])


def spectral_slope(magnitude_spectrum) -> float:                     # :426
    return float(_one(magnitude_spectrum, None)[6])


def spectral_slope_region(magnitude_spectrum, rfft_freqs, f_lower: float, f_upper: float, sample_rate: int) -> float:  # :447
    """Any band inside the spectrum (the reference panics when the upper edge reaches the last bin: AusError here)."""
    return float(_one(magnitude_spectrum, rfft_freqs, bands=((float(f_lower), float(f_upper)),), sample_rate=sample_rate)[7])


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
