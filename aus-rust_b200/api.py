"""Context object over the C ABI: numpy (host-pointer) entry points and torch (device-pointer) entry points.

torch is used only for device memory and streams (plumbing); all arithmetic happens in csrc/libaus_b200.so.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_int, c_uint64, c_void_p

import numpy as np

from . import _lib
from ._lib import NUM_FEATURES

F64, F32 = 0, 1


def _prec(precision) -> int:
    if precision in (0, "f64", "float64", np.float64):
        return F64
    if precision in (1, "f32", "float32", np.float32):
        return F32
    raise ValueError(f"unknown precision {precision!r}")


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _pd(a: np.ndarray):
    return a.ctypes.data_as(_lib._p_d)


def _pu64(a: np.ndarray):
    return a.ctypes.data_as(_lib._p_u64)


def _batch_desc(offsets, lengths, n_available: int):
    """offsets / lengths as uint64 arrays, checked against the buffer they index (the C ABI takes no total size: a bad
    descriptor would make the H2D copy read past the caller's buffer)."""
    off = np.ascontiguousarray(offsets, dtype=np.uint64).reshape(-1)
    ln = np.ascontiguousarray(lengths, dtype=np.uint64).reshape(-1)
    if off.size != ln.size:
        raise ValueError("offsets and lengths differ in length")
    if ln.size and int((off.astype(object) + ln.astype(object)).max()) > int(n_available):
        raise ValueError("a file descriptor (offset + length) reaches past the end of the audio buffer")
    return off, ln


def _check_out(out, total_rows: int, width: int = NUM_FEATURES):
    if out.dtype != np.float64 or not out.flags.c_contiguous or out.size != total_rows * width:
        raise ValueError(f"out must be a C-contiguous float64 array of {total_rows} x {width}")


def _pcm_bits(p: np.ndarray, sample_bits=None) -> int:
    """int8 -> 8, int16 -> 16, int32 -> 24 or 32 (``sample_bits``, default 32): the widths audiofile.rs:296-329 reads."""
    if p.dtype == np.int8:
        bits = 8
    elif p.dtype == np.int16:
        bits = 16
    elif p.dtype == np.int32:
        bits = int(sample_bits or 32)
        if bits not in (24, 32):
            raise ValueError("int32 PCM holds 24- or 32-bit samples")
    else:
        raise ValueError("pcm must be int8, int16 or int32")
    if sample_bits is not None and int(sample_bits) != bits:
        raise ValueError(f"sample_bits={sample_bits} does not match dtype {p.dtype}")
    return bits


def num_frames(n_samples: int, fft_size: int, hop: int) -> int:
    """Frame count of the loop in spectrum/fft.rs:126-176."""
    return int(_lib.load_library().aus_num_frames(int(n_samples), int(fft_size), int(hop)))


def irstft_len(n_frames: int, fft_size: int, hop: int) -> int:
    """spectrum/fft.rs:208."""
    return int(_lib.load_library().aus_irstft_len(int(n_frames), int(fft_size), int(hop)))


def batch_num_frames(lengths, fft_size: int, hop: int) -> int:
    ln = np.ascontiguousarray(lengths, dtype=np.uint64)
    return int(_lib.load_library().aus_batch_num_frames(_pu64(ln), ln.size, int(fft_size), int(hop)))


class Context:
    """One per device and caller thread (opaque ``aus_ctx``): stream, twiddle/window tables, scratch."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load_library()
        h = c_void_p()
        _lib.check(self._lib.aus_ctx_create(int(device), ctypes.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.aus_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing
    def synchronize(self):
        _lib.check(self._lib.aus_ctx_synchronize(self._h))

    def set_stream(self, cuda_stream_ptr: int):
        _lib.check(self._lib.aus_ctx_set_stream(self._h, c_void_p(int(cuda_stream_ptr))))

    def use_torch_stream(self):
        """Launch on torch's current stream of this device, so torch.cuda.Event timing sees the kernels."""
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def set_direct_load(self, on: bool):
        """Input staging of full aligned frames: False = TMA bulk copy into shared memory, True = global -> registers."""
        _lib.check(self._lib.aus_ctx_set_direct_load(self._h, 1 if on else 0))

    def set_forward_variant(self, mode: int):
        """Forward kernel for fft_size <= 2048: -1 auto, 0 shared-memory tail, 1 register-resident tail."""
        _lib.check(self._lib.aus_ctx_set_forward_variant(self._h, int(mode)))

    def set_chunk_samples(self, n: int):
        """Granularity of the copy/compute pipeline of the host-pointer feature calls (default 32 Mi samples)."""
        _lib.check(self._lib.aus_ctx_set_chunk_samples(self._h, int(n)))

    @property
    def launch_count(self) -> int:
        return int(self._lib.aus_ctx_launch_count(self._h))

    # ---- host-pointer API (numpy in, numpy out; H2D + kernels + D2H inside the call)
    def stft_features(self, audio, fft_size: int, hop: int, window_type: int, sample_rate: int, precision="f64"):
        """rstft -> |X| -> analyzer per frame, fused (analysis/analyzer.rs:99-104 with general hop/window)."""
        a = _f64(audio)
        out = np.empty((num_frames(a.size, fft_size, hop), NUM_FEATURES), dtype=np.float64)
        _lib.check(self._lib.aus_stft_features(self._h, _pd(a), a.size, int(fft_size), int(hop), int(window_type),
                                               int(sample_rate), _prec(precision), _pd(out)))
        return out

    def stft_analysis(self, audio, fft_size: int, sample_rate: int, max_num_threads=None):
        """mp.rs:26 ``stft_analysis``: hop = fft_size/2, Hamming."""
        a = _f64(audio)
        out = np.empty((num_frames(a.size, fft_size, max(1, fft_size // 2)), NUM_FEATURES), dtype=np.float64)
        _lib.check(self._lib.aus_stft_analysis(self._h, _pd(a), a.size, int(fft_size), int(sample_rate),
                                               int(max_num_threads or 0), out.ctypes.data_as(c_void_p)))
        return out

    def stft_features_batch(self, audio, offsets, lengths, fft_size: int, hop: int, window_type: int,
                            sample_rate: int, precision="f64", out=None):
        """Batch of files in one launch.  ``audio`` / ``out`` may be views of pinned host memory (then the copies
        inside the call run at full PCIe rate)."""
        a = _f64(audio)
        off, ln = _batch_desc(offsets, lengths, a.size)
        total = batch_num_frames(ln, fft_size, hop)
        if out is None:
            out = np.empty((total, NUM_FEATURES), dtype=np.float64)
        else:
            _check_out(out, total)
        starts = np.empty(ln.size + 1, dtype=np.uint64)
        _lib.check(self._lib.aus_stft_features_batch(self._h, _pd(a), _pu64(off), _pu64(ln), ln.size, int(fft_size),
                                                     int(hop), int(window_type), int(sample_rate), _prec(precision),
                                                     _pd(out), _pu64(starts)))
        return out, starts.astype(np.int64)

    def stft_features_batch_pcm(self, pcm, offsets, lengths, fft_size: int, hop: int, window_type: int,
                                sample_rate: int, precision="f64", sample_bits=None, out=None):
        """Decoded integer PCM (int16, or int32 holding 24- or 32-bit samples) uploaded as-is and widened on the device
        like the reference's reader: ``sample as f64 / INTMAX`` (audiofile.rs:296-329)."""
        p = np.ascontiguousarray(pcm)
        bits = _pcm_bits(p, sample_bits)
        off, ln = _batch_desc(offsets, lengths, p.size)
        total = batch_num_frames(ln, fft_size, hop)
        if out is None:
            out = np.empty((total, NUM_FEATURES), dtype=np.float64)
        else:
            _check_out(out, total)
        starts = np.empty(ln.size + 1, dtype=np.uint64)
        _lib.check(self._lib.aus_stft_features_batch_pcm(self._h, p.ctypes.data_as(c_void_p), bits, _pu64(off), _pu64(ln),
                                                         ln.size, int(fft_size), int(hop), int(window_type),
                                                         int(sample_rate), _prec(precision), _pd(out), _pu64(starts)))
        return out, starts.astype(np.int64)

    def rstft(self, audio, fft_size: int, hop: int, window_type: int, precision="f64") -> np.ndarray:
        """spectrum/fft.rs:119 -> complex128 [F, fft_size/2+1]."""
        a = _f64(audio)
        out = np.empty((num_frames(a.size, fft_size, hop), fft_size // 2 + 1), dtype=np.complex128)
        _lib.check(self._lib.aus_rstft(self._h, _pd(a), a.size, int(fft_size), int(hop), int(window_type),
                                       _prec(precision), out.ctypes.data_as(_lib._p_d)))
        return out

    def rfft(self, audio, fft_size: int, precision="f64") -> np.ndarray:
        """spectrum/fft.rs:37."""
        a = _f64(audio)
        if a.size != fft_size:
            raise ValueError("rfft: audio must have exactly fft_size samples (zero-pad first, fft.rs:23-24)")
        out = np.empty(fft_size // 2 + 1, dtype=np.complex128)
        _lib.check(self._lib.aus_rfft(self._h, _pd(a), int(fft_size), _prec(precision), out.ctypes.data_as(_lib._p_d)))
        return out

    def irstft(self, spectrogram, fft_size: int, hop: int, window_type: int, precision="f64", normalize=True):
        """spectrum/fft.rs:204."""
        s = np.ascontiguousarray(spectrogram, dtype=np.complex128)
        if s.ndim != 2:
            raise ValueError("spectrogram must be [frames, bins]")
        out = np.empty(irstft_len(s.shape[0], fft_size, hop), dtype=np.float64)
        _lib.check(self._lib.aus_irstft(self._h, s.ctypes.data_as(_lib._p_d), s.shape[0], s.shape[1], int(fft_size),
                                        int(hop), int(window_type), _prec(precision), 1 if normalize else 0, _pd(out)))
        return out

    def irfft(self, spectrum, fft_size: int, precision="f64") -> np.ndarray:
        """spectrum/fft.rs:66."""
        s = np.ascontiguousarray(spectrum, dtype=np.complex128).reshape(-1)
        out = np.empty(int(fft_size), dtype=np.float64)
        _lib.check(self._lib.aus_irfft(self._h, s.ctypes.data_as(_lib._p_d), s.size, int(fft_size), _prec(precision),
                                       _pd(out)))
        return out

    def complex_to_polar(self, spectra, want_phase=True):
        """spectrum/fft_tools.rs:21-55."""
        s = np.ascontiguousarray(spectra, dtype=np.complex128)
        mag = np.empty(s.shape, dtype=np.float64)
        ph = np.empty(s.shape, dtype=np.float64) if want_phase else None
        _lib.check(self._lib.aus_complex_to_polar(self._h, s.ctypes.data_as(_lib._p_d), s.size, _pd(mag),
                                                  _pd(ph) if want_phase else None))
        return mag, ph

    def polar_to_complex(self, mag, ph) -> np.ndarray:
        """spectrum/fft_tools.rs:131-188."""
        m, p = _f64(mag), _f64(ph)
        out = np.empty(m.shape, dtype=np.complex128)
        _lib.check(self._lib.aus_polar_to_complex(self._h, _pd(m), _pd(p), m.size, out.ctypes.data_as(_lib._p_d)))
        return out

    def analyzer_batch(self, magnitudes, fft_size: int, sample_rate: int, precision="f64") -> np.ndarray:
        """analysis/analyzer.rs:41 over rows."""
        m = _f64(magnitudes)
        if m.ndim == 1:
            m = m[None, :]
        if m.shape[1] != fft_size // 2 + 1:
            raise IndexError("magnitude rows must have fft_size/2+1 bins (the reference indexes out of bounds otherwise)")
        out = np.empty((m.shape[0], NUM_FEATURES), dtype=np.float64)
        _lib.check(self._lib.aus_analyzer_batch(self._h, _pd(m), m.shape[0], int(fft_size), int(sample_rate),
                                                _prec(precision), _pd(out)))
        return out

    def spectral_features_generic(self, magnitudes, rfft_freqs, roll_fractions=(), bands=(), sample_rate: int = 44100) -> np.ndarray:
        """The single-feature wrappers of spectral_analysis_tools.rs:306-471 with arbitrary parameters: rows of any
        length, any frequency table, any roll-off fractions, any slope bands [(f_lower, f_upper), ...].
        -> [rows, 7 + len(roll_fractions) + len(bands)]: centroid, variance, skewness, kurtosis, entropy, flatness, slope, ..."""
        m = _f64(magnitudes)
        if m.ndim == 1:
            m = m[None, :]
        f = _f64(rfft_freqs).reshape(-1)
        if f.size != m.shape[1]:
            raise IndexError("rfft_freqs and magnitude_spectrum differ in length (the reference panics)")
        r = _f64(np.asarray(roll_fractions, dtype=np.float64).reshape(-1))
        b = np.asarray(bands, dtype=np.float64).reshape(-1, 2)
        lo, hi = _f64(b[:, 0].copy()), _f64(b[:, 1].copy())
        out = np.empty((m.shape[0], 7 + r.size + lo.size), dtype=np.float64)
        _lib.check(self._lib.aus_spectral_features_generic(self._h, _pd(m), m.shape[0], m.shape[1], _pd(f), _pd(r), r.size,
                                                           _pd(lo), _pd(hi), lo.size, int(sample_rate), _pd(out)))
        return out

    def spectral_compute(self, what: int, a, b=None, s0: float = 0.0, s1: float = 0.0, sample_rate: int = 44100):
        """One of the `compute_*` / `make_*` / `dot_product` helpers of spectral_analysis_tools.rs:79-290 on one vector, with
        the caller's own intermediates (see aus_compute_kind in include/aus_b200.h) -> float, or an array for the two
        vector-valued kinds."""
        av = _f64(a).reshape(-1)
        bv = None if b is None else _f64(b).reshape(-1)
        if bv is not None and bv.size != av.size:
            raise IndexError("the two vectors differ in length (the reference panics)")
        vec = what in (_lib.COMPUTE_POWER_SPECTRUM, _lib.COMPUTE_SPECTRUM_PMF)
        out = np.empty(av.size if vec else 1, dtype=np.float64)
        _lib.check(self._lib.aus_spectral_compute(self._h, int(what), _pd(av), None if bv is None else _pd(bv), av.size,
                                                  float(s0), float(s1), int(sample_rate), _pd(out)))
        return out if vec else float(out[0])

    def overlap_add(self, audio_chunks, fft_size: int, hop: int) -> np.ndarray:
        """spectrum/fft_tools.rs:69: [n_chunks, fft_size] rows -> fft_size + hop (n_chunks - 1) samples."""
        ch = _f64(audio_chunks)
        if ch.ndim != 2 or ch.shape[1] != fft_size:
            raise IndexError("every chunk must hold fft_size samples (the reference indexes out of bounds)")
        out = np.empty(int(fft_size) + int(hop) * (ch.shape[0] - 1) if ch.shape[0] else 0, dtype=np.float64)
        _lib.check(self._lib.aus_overlap_add(self._h, _pd(ch), ch.shape[0], int(fft_size), int(hop), _pd(out)))
        return out

    def overlap_add_dev(self, d_chunks, n_chunks: int, fft_size: int, hop: int, d_audio):
        _lib.check(self._lib.aus_overlap_add_dev(self._h, self._check_dev(d_chunks, "f64", "d_chunks"), int(n_chunks), int(fft_size),
                                                 int(hop), self._check_dev(d_audio, "f64", "d_audio")))

    def autocorrelation(self, audio, fft_size: int) -> np.ndarray:
        """analysis/spectral_analysis_tools.rs:31 (one chunk, or [n_chunks, fft_size] rows)."""
        a = _f64(audio)
        if a.ndim == 1:
            out = np.empty(int(fft_size), dtype=np.float64)
            _lib.check(self._lib.aus_autocorrelation(self._h, _pd(a), a.size, int(fft_size), _pd(out)))
            return out
        if a.shape[1] != fft_size:
            raise _lib.SpectrumError(_lib.AUS_ERR_SIZE, "The audio length does not match the FFT size.")
        out = np.empty(a.shape, dtype=np.float64)
        _lib.check(self._lib.aus_autocorrelation_batch(self._h, _pd(a), a.shape[0], int(fft_size), _pd(out)))
        return out

    def time_features_batch(self, audio, offsets, lengths, sample_rate: int) -> np.ndarray:
        """dc_bias, dbfs_max, energy, zero_crossing_rate per file (analysis/audio_analysis_tools.rs:22-108)."""
        a = _f64(audio)
        off, ln = _batch_desc(offsets, lengths, a.size)
        out = np.empty((ln.size, 4), dtype=np.float64)
        _lib.check(self._lib.aus_time_features_batch(self._h, _pd(a), _pu64(off), _pu64(ln), ln.size, int(sample_rate), _pd(out)))
        return out

    def time_features_frames(self, audio, offsets, lengths, fft_size: int, hop: int, sample_rate: int) -> np.ndarray:
        """dc_bias, dbfs_max, energy, zero_crossing_rate of every STFT frame (same rows as the feature matrix)."""
        a = _f64(audio)
        off, ln = _batch_desc(offsets, lengths, a.size)
        out = np.empty((batch_num_frames(ln, fft_size, hop), 4), dtype=np.float64)
        _lib.check(self._lib.aus_time_features_frames(self._h, _pd(a), _pu64(off), _pu64(ln), ln.size, int(fft_size), int(hop),
                                                      int(sample_rate), _pd(out)))
        return out

    def pcm_to_f64(self, pcm, n_channels: int = 1, interleaved: bool = True, mixdown: bool = False, sample_bits=None) -> np.ndarray:
        """Decoded integer PCM -> f64 as audiofile.rs:296-329 does (``sample as f64 / INTMAX``), optionally followed by
        ``mixdown`` (audiofile.rs:168-178).  -> ``[n_channels, n_frames]``, or ``[n_frames]`` after mixdown."""
        p = np.ascontiguousarray(pcm)
        bits = _pcm_bits(p, sample_bits)
        if p.size % int(n_channels):
            raise ValueError("sample count is not a multiple of the channel count")
        nf = p.size // int(n_channels)
        out = np.empty(nf if mixdown else (int(n_channels), nf), dtype=np.float64)
        _lib.check(self._lib.aus_pcm_to_f64(self._h, p.ctypes.data_as(c_void_p), bits, nf, int(n_channels), 1 if interleaved else 0,
                                            1 if mixdown else 0, _pd(out)))
        return out

    def mixdown(self, channels) -> np.ndarray:
        """audiofile.rs:168-178 on f64 channels ``[n_channels, n_frames]`` (channel 0 is counted twice, as in the reference)."""
        ch = _f64(channels)
        if ch.ndim == 1:
            ch = ch[None, :]
        out = np.empty(ch.shape[1], dtype=np.float64)
        _lib.check(self._lib.aus_mixdown(self._h, _pd(ch), ch.shape[1], ch.shape[0], _pd(out)))
        return out

    def irstft_batch(self, spectrograms, fft_size: int, hop: int, window_type: int, precision="f64", normalize=True):
        """Several spectrograms (list of ``[frames_i, bins]``) in one launch -> list of signals (fft.rs:204 each)."""
        specs = [np.ascontiguousarray(s, dtype=np.complex128) for s in spectrograms]
        nf = np.array([s.shape[0] for s in specs], dtype=np.uint64)
        cat = np.ascontiguousarray(np.concatenate(specs, axis=0))
        lens = [irstft_len(int(f), fft_size, hop) for f in nf]
        out = np.empty(int(sum(lens)), dtype=np.float64)
        _lib.check(self._lib.aus_irstft_batch(self._h, cat.ctypes.data_as(_lib._p_d), _pu64(nf), nf.size, cat.shape[1], int(fft_size),
                                              int(hop), int(window_type), _prec(precision), 1 if normalize else 0, _pd(out)))
        return list(np.split(out, np.cumsum(lens)[:-1]))

    def set_host_register(self, mode: int):
        """Pageable inputs of the host-pointer feature calls: 0 driver staging, 1 cudaHostRegister around the call, 2 own
        pinned staging ring filled by several host threads, -1 auto (mode 2 from 64 MB)."""
        _lib.check(self._lib.aus_ctx_set_host_register(self._h, int(mode)))

    # ---- device-pointer API (torch CUDA tensors; stream-ordered, no sync)
    def _check_dev(self, t, precision, name):
        import torch
        want = torch.float64 if _prec(precision) == F64 else torch.float32
        if t is None:
            return None
        if not t.is_cuda or t.device.index != self.device or t.dtype != want or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous {want} CUDA tensor on device {self.device}")
        return c_void_p(t.data_ptr())

    def stft_features_batch_dev(self, d_audio, offsets, lengths, fft_size: int, hop: int, window_type: int,
                                sample_rate: int, precision="f64", d_features=None, d_spectra=None):
        off, ln = _batch_desc(offsets, lengths, d_audio.numel())
        total = batch_num_frames(ln, fft_size, hop)
        if d_features is not None and d_features.numel() < total * NUM_FEATURES:
            raise ValueError("d_features is smaller than total_frames x 14")
        if d_spectra is not None and d_spectra.numel() < total * (fft_size // 2 + 1) * 2:
            raise ValueError("d_spectra is smaller than total_frames x (fft_size/2+1) x 2")
        _lib.check(self._lib.aus_stft_features_batch_dev(
            self._h, self._check_dev(d_audio, precision, "d_audio"), _pu64(off), _pu64(ln), ln.size, int(fft_size),
            int(hop), int(window_type), int(sample_rate), _prec(precision),
            self._check_dev(d_features, precision, "d_features"), self._check_dev(d_spectra, precision, "d_spectra")))

    def analyzer_batch_dev(self, d_mags, n_frames: int, fft_size: int, sample_rate: int, precision, d_features):
        _lib.check(self._lib.aus_analyzer_batch_dev(self._h, self._check_dev(d_mags, precision, "d_mags"), int(n_frames),
                                                    int(fft_size), int(sample_rate), _prec(precision),
                                                    self._check_dev(d_features, precision, "d_features")))

    def irstft_dev(self, d_spectra, n_frames: int, fft_size: int, hop: int, window_type: int, precision, d_audio,
                   normalize=True):
        _lib.check(self._lib.aus_irstft_dev(self._h, self._check_dev(d_spectra, precision, "d_spectra"), int(n_frames),
                                            int(fft_size), int(hop), int(window_type), _prec(precision),
                                            1 if normalize else 0, self._check_dev(d_audio, precision, "d_audio")))

    def complex_to_polar_dev(self, d_spectra, d_mag, d_ph=None):
        """fft_tools.rs:41-55 on device tensors (f64; d_spectra is ``[..., 2]`` re/im pairs)."""
        _lib.check(self._lib.aus_complex_to_polar_dev(self._h, self._check_dev(d_spectra, "f64", "d_spectra"), d_mag.numel(),
                                                      self._check_dev(d_mag, "f64", "d_mag"), self._check_dev(d_ph, "f64", "d_ph")))

    def polar_to_complex_dev(self, d_mag, d_ph, d_spectra):
        """fft_tools.rs:164-188 on device tensors (f64)."""
        _lib.check(self._lib.aus_polar_to_complex_dev(self._h, self._check_dev(d_mag, "f64", "d_mag"), self._check_dev(d_ph, "f64", "d_ph"),
                                                      d_mag.numel(), self._check_dev(d_spectra, "f64", "d_spectra")))

    def irstft_batch_dev(self, d_spectra, n_frames, fft_size: int, hop: int, window_type: int, precision, d_audio, normalize=True):
        nf = np.ascontiguousarray(n_frames, dtype=np.uint64)
        _lib.check(self._lib.aus_irstft_batch_dev(self._h, self._check_dev(d_spectra, precision, "d_spectra"), _pu64(nf), nf.size,
                                                  int(fft_size), int(hop), int(window_type), _prec(precision), 1 if normalize else 0,
                                                  self._check_dev(d_audio, precision, "d_audio")))

    def time_features_frames_dev(self, d_audio, offsets, lengths, fft_size: int, hop: int, sample_rate: int, d_out):
        off, ln = _batch_desc(offsets, lengths, d_audio.numel())
        _lib.check(self._lib.aus_time_features_frames_dev(self._h, self._check_dev(d_audio, "f64", "d_audio"), _pu64(off), _pu64(ln),
                                                          ln.size, int(fft_size), int(hop), int(sample_rate),
                                                          self._check_dev(d_out, "f64", "d_out")))

    def mixdown_dev(self, d_channels, n_frames: int, n_channels: int, d_out):
        _lib.check(self._lib.aus_mixdown_dev(self._h, self._check_dev(d_channels, "f64", "d_channels"), int(n_frames), int(n_channels),
                                             self._check_dev(d_out, "f64", "d_out")))

    def synth_dev(self, d_audio, n0: int = 0, sample_rate: int = 44100, freq: float = 440.0, seed: int = 0,
                  file_id: int = 0, precision="f64"):
        _lib.check(self._lib.aus_synth_dev(self._h, self._check_dev(d_audio, precision, "d_audio"), d_audio.numel(),
                                           int(n0), int(sample_rate), float(freq), int(seed), int(file_id),
                                           _prec(precision)))


def stft_features_multi(devices, audio, offsets, lengths, fft_size: int, hop: int, window_type: int, sample_rate: int,
                        precision="f64") -> np.ndarray:
    """Shard a batch (or one long file) over several GPUs of this box; host threads, ordered gather."""
    lib = _lib.load_library()
    a = _f64(audio)
    off, ln = _batch_desc(offsets, lengths, a.size)
    devs = (c_int * len(devices))(*[int(d) for d in devices])
    out = np.empty((batch_num_frames(ln, fft_size, hop), NUM_FEATURES), dtype=np.float64)
    _lib.check(lib.aus_stft_features_multi(devs, len(devices), _pd(a), _pu64(off), _pu64(ln), ln.size, int(fft_size),
                                           int(hop), int(window_type), int(sample_rate), _prec(precision), _pd(out)))
    return out


def irstft_multi(devices, spectrograms, fft_size: int, hop: int, window_type: int, normalize=True):
    """irstft sharded by frame range (or by spectrogram) over several GPUs; list of ``[frames_i, bins]`` -> list of signals."""
    lib = _lib.load_library()
    specs = [np.ascontiguousarray(s, dtype=np.complex128) for s in spectrograms]
    nf = np.array([s.shape[0] for s in specs], dtype=np.uint64)
    cat = np.ascontiguousarray(np.concatenate(specs, axis=0))
    lens = [irstft_len(int(f), fft_size, hop) for f in nf]
    out = np.empty(int(sum(lens)), dtype=np.float64)
    devs = (c_int * len(devices))(*[int(d) for d in devices])
    _lib.check(lib.aus_irstft_multi(devs, len(devices), cat.ctypes.data_as(_lib._p_d), _pu64(nf), nf.size, cat.shape[1], int(fft_size),
                                    int(hop), int(window_type), 1 if normalize else 0, _pd(out)))
    return list(np.split(out, np.cumsum(lens)[:-1]))


def multi_shutdown():
    """Free the contexts cached by the multi-device entry points."""
    _lib.check(_lib.load_library().aus_multi_shutdown())


def plan_shards(offsets, lengths, fft_size: int, hop: int, n_devices: int):
    """The shard plan of aus_stft_features_multi as numpy arrays (vf_off, vf_len, vf_row, vf_dev); pure host."""
    lib = _lib.load_library()
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lengths, dtype=np.uint64)
    cap = ln.size * max(int(n_devices), 1) + 1
    vo, vl, vr = (np.empty(cap, dtype=np.uint64) for _ in range(3))
    vd = np.empty(cap, dtype=np.int32)
    n = ctypes.c_size_t(0)
    _lib.check(lib.aus_plan_shards(_pu64(off), _pu64(ln), ln.size, int(fft_size), int(hop), int(n_devices), _pu64(vo),
                                   _pu64(vl), _pu64(vr), vd.ctypes.data_as(ctypes.POINTER(c_int)), cap, ctypes.byref(n)))
    k = n.value
    return vo[:k].astype(np.int64), vl[:k].astype(np.int64), vr[:k].astype(np.int64), vd[:k].copy()


_default = {}
_default_lock = threading.Lock()


def default_context() -> Context:
    """Process-wide context used by the reference-named free functions (device = LOCAL_RANK or 0), one per thread."""
    key = threading.get_ident()
    with _default_lock:
        ctx = _default.get(key)
        if ctx is None:
            ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
            _default[key] = ctx
        return ctx
