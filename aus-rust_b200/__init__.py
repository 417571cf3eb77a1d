"""aus_rust_b200 -- B200-native (sm_100a) STFT + spectral-feature path of the `aus` crate.

Host-side mirror of the reference's public API for this path (same module and function names as
fleximeter/aus-rust: ``window``, ``spectrum``, ``analysis``, ``mp``), implemented as a thin ctypes binding over the
C ABI in ``include/aus_b200.h`` (``csrc/libaus_b200.so``).  There is no CPU fallback: importing works anywhere, but
every compute call needs the CUDA library and a B200, and fails loudly otherwise.
"""

from ._lib import AusError, SpectrumError, lib_path, load_library, FEATURE_NAMES, NUM_FEATURES  # noqa: F401
from .api import (Context, default_context, num_frames, irstft_len, batch_num_frames, plan_shards,  # noqa: F401
                  stft_features_multi, irstft_multi, multi_shutdown)
from .window import (WindowType, generate_window, generate_window_bartlett, generate_window_blackman,  # noqa: F401
                     generate_window_hanning, generate_window_hamming, generate_window_rectangular)
from . import spectrum, analysis, mp  # noqa: F401

__all__ = ["Context", "default_context", "WindowType", "generate_window", "spectrum", "analysis", "mp",
           "AusError", "SpectrumError", "num_frames", "irstft_len", "batch_num_frames", "FEATURE_NAMES"]
