"""ctypes binding of include/aus_b200.h.  Loading fails loudly when the CUDA library has not been built."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_size_t, c_uint32, c_uint64, c_void_p

NUM_FEATURES = 14
FEATURE_NAMES = (
    "spectral_centroid", "spectral_variance", "spectral_skewness", "spectral_kurtosis", "spectral_entropy",
    "spectral_flatness", "spectral_roll_off_50", "spectral_roll_off_75", "spectral_roll_off_90",
    "spectral_roll_off_95", "spectral_slope", "spectral_slope_0_1_khz", "spectral_slope_1_5_khz",
    "spectral_slope_0_5_khz",
)  # analysis/analyzer.rs:10-23

# aus_compute_kind (include/aus_b200.h)
(COMPUTE_CENTROID, COMPUTE_ENTROPY, COMPUTE_FLATNESS, COMPUTE_KURTOSIS, COMPUTE_ROLL_OFF, COMPUTE_SKEWNESS, COMPUTE_SLOPE,
 COMPUTE_SLOPE_REGION, COMPUTE_VARIANCE, COMPUTE_DOT_PRODUCT, COMPUTE_POWER_SPECTRUM, COMPUTE_SPECTRUM_PMF) = range(12)

AUS_OK, AUS_ERR_INVALID, AUS_ERR_CUDA, AUS_ERR_ALLOC, AUS_ERR_SIZE, AUS_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5


class AusError(RuntimeError):
    """Any non-zero status of the C ABI (the reference would panic or hang in most of these cases)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"aus_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class SpectrumError(AusError):
    """spectrum/fft.rs:16-19 ``SpectrumError { error_msg }``: the only ``Err`` the reference returns on this path."""

    @property
    def error_msg(self) -> str:
        return self.msg


def lib_path() -> str:
    # AUS_B200_LIB: load another build of the same C ABI (kernel experiments, tools/kbench.py); default = the in-tree build
    return os.environ.get("AUS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libaus_b200.so")


_LIB = None

_p_d = POINTER(c_double)
_p_u64 = POINTER(c_uint64)

# name -> (restype, argtypes); every symbol include/aus_b200.h declares
SIGNATURES = {
    "aus_version": (c_char_p, []),
    "aus_last_error": (c_char_p, []),
    "aus_device_count": (c_int, []),
    "aus_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "aus_ctx_destroy": (None, [c_void_p]),
    "aus_ctx_device": (c_int, [c_void_p]),
    "aus_ctx_stream": (c_void_p, [c_void_p]),
    "aus_ctx_set_stream": (c_int, [c_void_p, c_void_p]),
    "aus_ctx_synchronize": (c_int, [c_void_p]),
    "aus_ctx_launch_count": (c_uint64, [c_void_p]),
    "aus_ctx_set_chunk_samples": (c_int, [c_void_p, c_uint64]),
    "aus_ctx_set_direct_load": (c_int, [c_void_p, c_int]),
    "aus_ctx_set_forward_variant": (c_int, [c_void_p, c_int]),
    "aus_ctx_set_host_register": (c_int, [c_void_p, c_int]),
    "aus_multi_shutdown": (c_int, []),
    "aus_num_frames": (c_size_t, [c_size_t, c_size_t, c_size_t]),
    "aus_irstft_len": (c_size_t, [c_size_t, c_size_t, c_size_t]),
    "aus_generate_window": (c_int, [c_int, c_size_t, _p_d]),
    "aus_rfftfreq": (c_int, [c_size_t, c_uint32, _p_d]),
    "aus_stft_features": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, c_size_t, c_int, c_uint32, c_int, _p_d]),
    "aus_stft_analysis": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, c_uint32, c_size_t, c_void_p]),
    "aus_stft_features_batch": (c_int, [c_void_p, _p_d, _p_u64, _p_u64, c_size_t, c_size_t, c_size_t, c_int,
                                        c_uint32, c_int, _p_d, _p_u64]),
    "aus_stft_features_batch_pcm": (c_int, [c_void_p, c_void_p, c_int, _p_u64, _p_u64, c_size_t, c_size_t, c_size_t, c_int,
                                            c_uint32, c_int, _p_d, _p_u64]),
    "aus_stft_features_batch_dev": (c_int, [c_void_p, c_void_p, _p_u64, _p_u64, c_size_t, c_size_t, c_size_t, c_int,
                                            c_uint32, c_int, c_void_p, c_void_p]),
    "aus_batch_num_frames": (c_uint64, [_p_u64, c_size_t, c_size_t, c_size_t]),
    "aus_stft_features_multi": (c_int, [POINTER(c_int), c_int, _p_d, _p_u64, _p_u64, c_size_t, c_size_t, c_size_t,
                                        c_int, c_uint32, c_int, _p_d]),
    "aus_plan_shards": (c_int, [_p_u64, _p_u64, c_size_t, c_size_t, c_size_t, c_int, _p_u64, _p_u64, _p_u64,
                                POINTER(c_int), c_size_t, POINTER(c_size_t)]),
    "aus_rstft": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, c_size_t, c_int, c_int, _p_d]),
    "aus_rstft_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_int, c_void_p]),
    "aus_rfft": (c_int, [c_void_p, _p_d, c_size_t, c_int, _p_d]),
    "aus_complex_to_polar": (c_int, [c_void_p, _p_d, c_size_t, _p_d, _p_d]),
    "aus_polar_to_complex": (c_int, [c_void_p, _p_d, _p_d, c_size_t, _p_d]),
    "aus_complex_to_polar_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "aus_polar_to_complex_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "aus_spectral_features_generic": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, _p_d, _p_d, c_size_t, _p_d, _p_d, c_size_t,
                                              c_uint32, _p_d]),
    "aus_spectral_compute": (c_int, [c_void_p, c_int, _p_d, _p_d, c_size_t, c_double, c_double, c_uint32, _p_d]),
    "aus_overlap_add": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, c_size_t, _p_d]),
    "aus_overlap_add_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    "aus_analyzer_batch": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, c_uint32, c_int, _p_d]),
    "aus_analyzer_batch_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_uint32, c_int, c_void_p]),
    "aus_irstft": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, c_size_t, c_size_t, c_int, c_int, c_int, _p_d]),
    "aus_irstft_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_int, c_int, c_void_p]),
    "aus_irfft": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, c_int, _p_d]),
    "aus_irstft_batch": (c_int, [c_void_p, _p_d, _p_u64, c_size_t, c_size_t, c_size_t, c_size_t, c_int, c_int, c_int, _p_d]),
    "aus_irstft_batch_dev": (c_int, [c_void_p, c_void_p, _p_u64, c_size_t, c_size_t, c_size_t, c_int, c_int, c_int, c_void_p]),
    "aus_irstft_multi": (c_int, [POINTER(c_int), c_int, _p_d, _p_u64, c_size_t, c_size_t, c_size_t, c_size_t, c_int, c_int, _p_d]),
    "aus_autocorrelation": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, _p_d]),
    "aus_autocorrelation_batch": (c_int, [c_void_p, _p_d, c_size_t, c_size_t, _p_d]),
    "aus_time_features_batch": (c_int, [c_void_p, _p_d, _p_u64, _p_u64, c_size_t, c_uint32, _p_d]),
    "aus_time_features_frames": (c_int, [c_void_p, _p_d, _p_u64, _p_u64, c_size_t, c_size_t, c_size_t, c_uint32, _p_d]),
    "aus_time_features_frames_dev": (c_int, [c_void_p, c_void_p, _p_u64, _p_u64, c_size_t, c_size_t, c_size_t, c_uint32, c_void_p]),
    "aus_pcm_to_f64": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_int, c_int, c_int, _p_d]),
    "aus_pcm_to_f64_dev": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_int, c_int, c_int, c_void_p]),
    "aus_mixdown": (c_int, [c_void_p, _p_d, c_size_t, c_int, _p_d]),
    "aus_mixdown_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "aus_synth_dev": (c_int, [c_void_p, c_void_p, c_uint64, c_uint64, c_uint32, c_double, c_uint64, c_uint64, c_int]),
}


def load_library():
    """Load csrc/libaus_b200.so (built by ``__graft_entry__.build()`` / ``make -C aus-rust_b200/csrc``)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`make -C aus-rust_b200/csrc`.  aus_rust_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int):
    if rc == AUS_OK:
        return
    msg = load_library().aus_last_error().decode("utf-8", "replace")
    if rc == AUS_ERR_SIZE:
        raise SpectrumError(rc, msg)
    raise AusError(rc, msg)
