"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Tolerances are the north star's:
relative 1e-9 for f64 features and spectra, 1e-4 for f32."""
import numpy as np
import pytest

from helpers import ROLL, assert_features, assert_spectra, cancellation_scales, rel_err

pytestmark = pytest.mark.gpu

RTOL64 = 1e-9
RTOL32 = 1e-4
SR = 44100


@pytest.mark.parametrize("fft_size,hop", [(64, 16), (128, 64), (256, 64), (512, 128), (1024, 256), (2048, 512),
                                          (2048, 1024), (4096, 1024), (8192, 2048), (16384, 4096)])
def test_features_f64_all_sizes(ctx, orc, fft_size, hop):
    n = max(6 * fft_size + 37, 30000)
    x = orc.synth_signal(n, SR, seed=1, file_id=fft_size)
    ref = orc.stft_features(x, fft_size, hop, orc.HANNING, SR)
    got = ctx.stft_features(x, fft_size, hop, orc.HANNING, SR, "f64")
    assert_features(got, ref, RTOL64, f"N={fft_size}")


@pytest.mark.parametrize("window", [0, 1, 2, 3, 4])
def test_features_f64_all_windows(ctx, orc, window):
    x = orc.synth_signal(50000, SR, seed=2, file_id=window)
    ref = orc.stft_features(x, 2048, 512, window, SR)
    got = ctx.stft_features(x, 2048, 512, window, SR, "f64")
    assert_features(got, ref, RTOL64, orc.WINDOW_NAMES[window])


def test_config1_full_length(ctx, orc):
    """BASELINE config 1 at full size: 60 s mono, fft 2048, hop 512, Hann -> 5165 frames."""
    x = orc.synth_signal(60 * SR)
    ref = orc.stft_features(x, 2048, 512, orc.HANNING, SR)
    got = ctx.stft_features(x, 2048, 512, orc.HANNING, SR, "f64")
    assert got.shape == (5165, 14)
    assert_features(got, ref, RTOL64, "config 1")


def test_stft_analysis_signature_path(ctx, orc):
    """mp::stft_analysis: hop = N/2, Hamming (mp.rs:42)."""
    x = orc.synth_signal(3 * SR, SR, seed=3)
    ref = orc.stft_analysis(x, 2048, SR, 8)
    got = ctx.stft_analysis(x, 2048, SR, 8)
    assert_features(got, ref, RTOL64, "stft_analysis")


@pytest.mark.parametrize("fft_size,hop,window", [(256, 64, 2), (1024, 256, 3), (2048, 512, 2), (4096, 1024, 1),
                                                 (16384, 4096, 0)])
def test_rstft_spectra(ctx, orc, fft_size, hop, window):
    x = orc.synth_signal(5 * fft_size + 11, SR, seed=4, file_id=fft_size)
    ref = orc.rstft(x, fft_size, hop, window)
    got = ctx.rstft(x, fft_size, hop, window)
    assert_spectra(got, ref, RTOL64, f"rstft N={fft_size}")


@pytest.mark.parametrize("n", [0, 1, 2, 5, 100, 2047, 2048, 2049, 2560, 2561, 4095, 4096, 4097])
def test_ragged_lengths_and_short_last_frame(ctx, orc, n):
    """Frame-count rule (strict '<', fft.rs:131) and the short last window (fft.rs:148-159), incl. the L = 1 NaN."""
    x = orc.synth_signal(n, SR, seed=5, file_id=n)
    for window in (orc.HANNING, orc.BLACKMAN):
        ref_s = orc.rstft(x, 2048, 512, window)
        got_s = ctx.rstft(x, 2048, 512, window)
        assert got_s.shape == ref_s.shape
        assert np.array_equal(np.isnan(got_s), np.isnan(ref_s))
        assert_spectra(np.nan_to_num(got_s), np.nan_to_num(ref_s), RTOL64, f"n={n}")
        ref = orc.stft_features(x, 2048, 512, window, SR)
        got = ctx.stft_features(x, 2048, 512, window, SR)
        # a 1- or 2-sample frame has a flat spectrum: its slopes are pure cancellation (see cancellation_scales)
        sc = cancellation_scales(np.abs(np.nan_to_num(ref_s)), 2048, SR)
        assert_features(got, ref, RTOL64, f"n={n} w={window}", scales=sc)


def test_silent_frames_nan_pattern(ctx, orc):
    x = np.zeros(8000)
    x[5000:] = orc.synth_signal(3000, SR, seed=6)
    ref = orc.stft_features(x, 2048, 512, orc.HANNING, SR)
    got = ctx.stft_features(x, 2048, 512, orc.HANNING, SR)
    assert np.isnan(ref[0, :6]).all() and (ref[0, 6:] == 0).all()
    assert_features(got, ref, RTOL64, "silence")


def test_batch_ragged_files(ctx, orc):
    lens = [0, 1, 300, 2048, 5000, 2049, 12345, 777, 40000]
    offs, pieces, pos = [], [], 0
    for i, ln in enumerate(lens):
        offs.append(pos)
        pieces.append(orc.synth_signal(ln, SR, seed=7, file_id=i))
        pos += ln + (i % 3)          # odd gaps: some files start on an unaligned sample -> non-TMA path
        pieces.append(np.zeros(i % 3))
    audio = np.concatenate(pieces)
    got, starts = ctx.stft_features_batch(audio, offs, lens, 1024, 256, orc.HAMMING, SR)
    for i, ln in enumerate(lens):
        ref = orc.stft_features(audio[offs[i]:offs[i] + ln], 1024, 256, orc.HAMMING, SR)
        assert starts[i + 1] - starts[i] == ref.shape[0]
        assert_features(got[starts[i]:starts[i + 1]], ref, RTOL64, f"file {i}")


@pytest.mark.parametrize("fft_size", [256, 2048, 16384])
def test_analyzer_batch(ctx, orc, fft_size):
    rng = np.random.default_rng(fft_size)
    mags = np.abs(rng.standard_normal((37, fft_size // 2 + 1))) + 1e-3
    mags[3] = 0.0                       # silent row
    mags[5, 17] = 0.0                   # one zero bin: entropy NaN, flatness 0
    ref = orc.analyze_spectrogram(mags, fft_size, SR)
    got = ctx.analyzer_batch(mags, fft_size, SR)
    assert_features(got, ref, RTOL64, "analyzer_batch")


@pytest.mark.parametrize("fft_size,hop,window", [(2048, 1024, 3), (4096, 1024, 2), (1024, 256, 3), (256, 128, 0),
                                                 (16384, 4096, 3), (512, 512, 4), (64, 16, 3), (8192, 4096, 2)])
def test_irstft(ctx, orc, fft_size, hop, window):
    x = orc.synth_signal(20 * fft_size + 123, SR, seed=8, file_id=fft_size)
    S = orc.rstft(x, fft_size, hop, window)
    ref = orc.irstft(S, fft_size, hop, window)
    got = ctx.irstft(S, fft_size, hop, window)
    assert got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))          # Hanning / Bartlett: 0/0 at both ends
    fin = np.isfinite(ref)
    scale = np.max(np.abs(ref[fin]))
    assert np.max(np.abs(got[fin] - ref[fin])) <= RTOL64 * scale


def test_irstft_blackman_unnormalised(ctx, orc):
    """Blackman's w[0] is a rounding residue (~ -1.4e-17), so out[0] = (FFT rounding noise) / 1e-17 is arbitrary in
    ANY implementation and may even become the normalisation peak.  Parity is therefore checked before the peak
    normalisation and away from samples whose window norm is ~0."""
    fft_size, hop = 1024, 256
    x = orc.synth_signal(20 * fft_size + 123, SR, seed=8, file_id=1)
    S = orc.rstft(x, fft_size, hop, orc.BLACKMAN)
    ref = orc.irstft(S, fft_size, hop, orc.BLACKMAN, normalize=False)
    got = ctx.irstft(S, fft_size, hop, orc.BLACKMAN, normalize=False)
    w2 = orc.generate_window(orc.BLACKMAN, fft_size) ** 2
    norm = np.zeros(ref.size)
    for h in range(S.shape[0]):
        norm[h * hop:h * hop + fft_size] += w2
    ok = norm > 1e-12
    assert ok.sum() >= ref.size - 2
    assert np.max(np.abs(got[ok] - ref[ok])) <= RTOL64 * np.max(np.abs(ref[ok]))


def test_rfft_irfft(ctx, orc):
    for fft_size in (64, 2048, 8192):
        x = orc.synth_signal(fft_size, SR, seed=9, file_id=fft_size)
        ref = orc.rfft(x, fft_size)
        got = ctx.rfft(x, fft_size)
        assert_spectra(got, ref, RTOL64, "rfft")
        back = ctx.irfft(got, fft_size)
        assert np.max(np.abs(back - fft_size * x)) <= RTOL64 * fft_size * np.max(np.abs(x))
        assert np.max(np.abs(back - orc.irfft(ref, fft_size))) <= RTOL64 * fft_size


def test_f32_mode(ctx, orc):
    for fft_size, hop in ((1024, 256), (2048, 512), (16384, 4096)):
        x = orc.synth_signal(12 * fft_size, SR, seed=10, file_id=fft_size)
        ref = orc.stft_features(x, fft_size, hop, orc.HANNING, SR)
        got = ctx.stft_features(x, fft_size, hop, orc.HANNING, SR, "f32")
        r = rel_err(got, ref)
        r[:, ROLL] = 0.0
        assert r.max() <= RTOL32, f"f32 N={fft_size}: {r.max():.3e}"
        # roll-off points are exact bin frequencies; an f32 tie may land one bin away
        f0 = SR / fft_size
        assert np.all(np.abs(got[:, ROLL] - ref[:, ROLL]) <= 1.5 * f0)
        assert np.mean(got[:, ROLL] == ref[:, ROLL]) >= 0.97
        Sref = orc.rstft(x, fft_size, hop, orc.HANNING)
        assert_spectra(ctx.rstft(x, fft_size, hop, orc.HANNING, "f32"), Sref, RTOL32, "f32 spectra")


def test_polar_roundtrip(ctx, orc):
    x = orc.synth_signal(9000, SR, seed=11)
    S = orc.rstft(x, 1024, 512, orc.HAMMING)
    mag, ph = ctx.complex_to_polar(S)
    mref, pref = orc.complex_to_polar_rstft(S)
    assert np.max(np.abs(mag - mref) / np.maximum(mref, 1e-300)) <= RTOL64
    assert np.max(np.abs(ph - pref)) <= 1e-12
    back = ctx.polar_to_complex(mag, ph)
    assert_spectra(back, S, RTOL64, "polar round trip")


def test_error_codes(ctx, aus, orc):
    x = orc.synth_signal(5000)
    with pytest.raises(aus.AusError):
        ctx.rstft(x, 2048, 0, 2)                      # hop = 0: the reference never terminates
    with pytest.raises(aus.AusError):
        ctx.rstft(x, 2048, 4096, 2)                   # hop > N: the reference panics
    with pytest.raises(aus.AusError):
        ctx.rstft(x, 1000, 250, 2)                    # rustfft takes any N; this build is power-of-two only
    with pytest.raises(aus.AusError):
        ctx.stft_features(x, 2048, 512, 2, 8000)      # 5 kHz band above Nyquist: the reference panics
    S = orc.rstft(x, 2048, 1024, 3)
    with pytest.raises(aus.SpectrumError) as e:
        ctx.irstft(S[:, :-1], 2048, 1024, 3)
    assert "incorrect length" in e.value.error_msg and "1025" in e.value.error_msg
    with pytest.raises(aus.SpectrumError):
        ctx.irfft(S[0, :-2], 2048)
