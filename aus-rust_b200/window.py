"""Mirror of src/window.rs: ``WindowType`` and the window generators (host, f64, the reference's exact formulas)."""
from __future__ import annotations

import enum

import numpy as np

from . import _lib


class WindowType(enum.IntEnum):
    """window.rs:9-15 (same order == C-ABI code)."""
    Bartlett = 0
    Blackman = 1
    Hanning = 2
    Hamming = 3
    Rectangular = 4


def generate_window(window_type, window_length: int) -> np.ndarray:
    """window.rs:110-118."""
    out = np.empty(int(window_length), dtype=np.float64)
    lib = _lib.load_library()
    _lib.check(lib.aus_generate_window(int(window_type), int(window_length), out.ctypes.data_as(_lib._p_d)))
    return out


def generate_window_bartlett(n: int) -> np.ndarray:      # window.rs:26
    return generate_window(WindowType.Bartlett, n)


def generate_window_blackman(n: int) -> np.ndarray:      # window.rs:43
    return generate_window(WindowType.Blackman, n)


def generate_window_hanning(n: int) -> np.ndarray:       # window.rs:61
    return generate_window(WindowType.Hanning, n)


def generate_window_hamming(n: int) -> np.ndarray:       # window.rs:78
    return generate_window(WindowType.Hamming, n)


def generate_window_rectangular(n: int) -> np.ndarray:   # window.rs:95
    return generate_window(WindowType.Rectangular, n)
