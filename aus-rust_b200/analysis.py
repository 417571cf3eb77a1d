"""Mirror of src/analysis/{analyzer.rs, spectral_analysis_tools.rs:306-471}."""
from __future__ import annotations

from dataclasses import dataclass, fields

import numpy as np

from . import _lib
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


def _c(what, a, b=None, s0=0.0, s1=0.0, sample_rate=44100):
    return default_context().spectral_compute(what, a, b, s0, s1, sample_rate)


# the public building blocks of spectral_analysis_tools.rs:79-290: same names, arguments and meaning -- they use whatever
# intermediates (sums, pmf, centroid, variance) the caller passes, like the reference
def make_power_spectrum(magnitude_spectrum) -> np.ndarray:                                          # :273
    return _c(_lib.COMPUTE_POWER_SPECTRUM, magnitude_spectrum)


def make_spectrum_pmf(power_spectrum, power_spectrum_sum: float) -> np.ndarray:                     # :284
    return _c(_lib.COMPUTE_SPECTRUM_PMF, power_spectrum, None, power_spectrum_sum)


def dot_product(vec1, vec2) -> float:                                                               # :253
    return _c(_lib.COMPUTE_DOT_PRODUCT, vec1, vec2)


def compute_spectral_centroid(magnitude_spectrum, rfft_freqs, magnitude_spectrum_sum: float) -> float:        # :79
    return _c(_lib.COMPUTE_CENTROID, magnitude_spectrum, rfft_freqs, magnitude_spectrum_sum)


def compute_spectral_entropy(spectrum_pmf) -> float:                                                # :93
    return _c(_lib.COMPUTE_ENTROPY, spectrum_pmf)


def compute_spectral_flatness(magnitude_spectrum, magnitude_spectrum_sum: float) -> float:          # :107
    return _c(_lib.COMPUTE_FLATNESS, magnitude_spectrum, None, magnitude_spectrum_sum)


def compute_spectral_kurtosis(spectrum_pmf, rfft_freqs, spectral_centroid: float, spectral_variance: float) -> float:   # :122
    return _c(_lib.COMPUTE_KURTOSIS, spectrum_pmf, rfft_freqs, spectral_centroid, spectral_variance)


def compute_spectral_roll_off_point(power_spectrum, rfft_freqs, power_spectrum_sum: float, n: float) -> float:          # :137
    return _c(_lib.COMPUTE_ROLL_OFF, power_spectrum, rfft_freqs, power_spectrum_sum, n)


def compute_spectral_skewness(spectrum_pmf, rfft_freqs, spectral_centroid: float, spectral_variance: float) -> float:   # :153
    return _c(_lib.COMPUTE_SKEWNESS, spectrum_pmf, rfft_freqs, spectral_centroid, spectral_variance)


def compute_spectral_slope(power_spectrum, power_spectrum_sum: float) -> float:                     # :167
    return _c(_lib.COMPUTE_SLOPE, power_spectrum, None, power_spectrum_sum)


def compute_spectral_slope_region(power_spectrum, rfft_freqs, f_lower: float, f_upper: float, sample_rate: int) -> float:  # :190
    return _c(_lib.COMPUTE_SLOPE_REGION, power_spectrum, rfft_freqs, f_lower, f_upper, sample_rate)


def compute_spectral_variance(spectrum_pmf, rfft_freqs, spectral_centroid: float) -> float:         # :238
    return _c(_lib.COMPUTE_VARIANCE, spectrum_pmf, rfft_freqs, spectral_centroid)


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
