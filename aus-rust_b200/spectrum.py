"""Mirror of src/spectrum/{fft.rs, fft_tools.rs}: same function names and argument meaning as the reference."""
from __future__ import annotations

import numpy as np

from ._lib import SpectrumError  # noqa: F401  (fft.rs:16-19)
from .api import default_context
from . import _lib


def rfft(audio, fft_size: int) -> np.ndarray:
    """fft.rs:37 -- complex spectrum, bins 0..=fft_size/2, un-normalised."""
    return default_context().rfft(audio, fft_size)


def irfft(spectrum, fft_size: int) -> np.ndarray:
    """fft.rs:66 -- raises SpectrumError when len(spectrum) != fft_size/2+1; output is fft_size x the signal."""
    return default_context().irfft(spectrum, fft_size)


def rstft(audio, fft_size: int, hop_size: int, window_type) -> np.ndarray:
    """fft.rs:119 -- complex128 [frames, fft_size/2+1]."""
    return default_context().rstft(audio, fft_size, hop_size, int(window_type))


def irstft(spectrogram, fft_size: int, hop_size: int, window_type) -> np.ndarray:
    """fft.rs:204 -- raises SpectrumError on a wrong frame length."""
    return default_context().irstft(spectrogram, fft_size, hop_size, int(window_type))


def complex_to_polar_rfft(spectrum):
    """fft_tools.rs:21 -> (magnitude, phase)."""
    return default_context().complex_to_polar(spectrum)


def complex_to_polar_rstft(spectrogram):
    """fft_tools.rs:41 -> (magnitude, phase), each [frames, bins]."""
    return default_context().complex_to_polar(spectrogram)


def polar_to_complex_rfft(magnitude_spectrum, phase_spectrum) -> np.ndarray:
    """fft_tools.rs:131."""
    m, p = np.asarray(magnitude_spectrum), np.asarray(phase_spectrum)
    if m.shape != p.shape:
        raise SpectrumError(_lib.AUS_ERR_SIZE, "The magnitude spectrum and phase spectrum do not have the same length. "
                            f"The magnitude spectrum has len {m.shape[0]} and the phase spectrum has len {p.shape[0]}.")
    return default_context().polar_to_complex(m, p)


def polar_to_complex_rstft(magnitude_spectrogram, phase_spectrogram) -> np.ndarray:
    """fft_tools.rs:164."""
    m, p = np.asarray(magnitude_spectrogram), np.asarray(phase_spectrogram)
    if m.shape != p.shape:
        raise SpectrumError(_lib.AUS_ERR_SIZE, "The magnitude spectrogram and phase spectrogram do not have the same "
                            f"length. The magnitude spectrogram has len {m.shape[0]} and the phase spectrogram has len {p.shape[0]}.")
    return default_context().polar_to_complex(m, p)


def overlap_add(audio_chunks, fft_size: int, hop_size: int) -> np.ndarray:
    """fft_tools.rs:69 -- chunk h added at offset h * hop_size, ascending chunk order per sample."""
    return default_context().overlap_add(audio_chunks, fft_size, hop_size)


def rfftfreq(fft_size: int, sample_rate: int) -> np.ndarray:
    """fft_tools.rs:199."""
    out = np.empty(fft_size // 2 + 1, dtype=np.float64)
    _lib.check(_lib.load_library().aus_rfftfreq(int(fft_size), int(sample_rate), out.ctypes.data_as(_lib._p_d)))
    return out
