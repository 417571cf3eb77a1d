"""Mirror of src/mp.rs: the threaded analyzer, with the GPU in place of the thread pool."""
from __future__ import annotations

import numpy as np

from .analysis import Analysis
from .api import default_context, stft_features_multi


def stft_analysis_array(audio, fft_size: int, sample_rate: int, max_num_threads=None) -> np.ndarray:
    """mp.rs:26 as a float64 [frames, 14] array (hop = fft_size/2, Hamming, mp.rs:42)."""
    return default_context().stft_analysis(audio, fft_size, sample_rate, max_num_threads)


def stft_analysis(audio, fft_size: int, sample_rate: int, max_num_threads=None):
    """mp.rs:26 ``stft_analysis(audio, fft_size, sample_rate, max_num_threads) -> Vec<Analysis>`` in frame order.
    ``max_num_threads`` is accepted for signature compatibility and ignored."""
    return [Analysis.from_row(r) for r in stft_analysis_array(audio, fft_size, sample_rate, max_num_threads)]


def stft_features(audio, fft_size: int, hop_size: int, window_type, sample_rate: int, precision="f64") -> np.ndarray:
    """New general entry point (hop and window are fixed inside mp.rs:42): fused rstft -> magnitude -> analyzer."""
    return default_context().stft_features(audio, fft_size, hop_size, int(window_type), sample_rate, precision)


def stft_features_sharded(devices, audio, fft_size: int, hop_size: int, window_type, sample_rate: int,
                          precision="f64") -> np.ndarray:
    """One long file cut into contiguous frame ranges over several GPUs (the analogue of mp.rs:47-57)."""
    a = np.ascontiguousarray(audio, dtype=np.float64)
    return stft_features_multi(devices, a, [0], [a.size], fft_size, hop_size, int(window_type), sample_rate, precision)
