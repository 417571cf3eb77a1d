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
        ctx.rstft(x, 10000, 2500, 2)                  # rustfft takes any N; here: powers of two to 16384, anything to 8192
    with pytest.raises(aus.AusError):
        ctx.stft_features(x, 2048, 512, 2, 8000)      # 5 kHz band above Nyquist: the reference panics
    S = orc.rstft(x, 2048, 1024, 3)
    with pytest.raises(aus.SpectrumError) as e:
        ctx.irstft(S[:, :-1], 2048, 1024, 3)
    assert "incorrect length" in e.value.error_msg and "1025" in e.value.error_msg
    with pytest.raises(aus.SpectrumError):
        ctx.irfft(S[0, :-2], 2048)


# ------------------------------------------------------------------------------------------------ golden fixtures
def test_golden_fixtures(ctx, orc):
    """CUDA path vs the committed oracle outputs (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
    assert_features(ctx.stft_features(g["c1_x"], 2048, 512, orc.HANNING, SR), g["c1_feat"], RTOL64, "golden c1")
    assert_spectra(ctx.rstft(g["c1_x"], 2048, 512, orc.HANNING)[[0, 7, -1]], g["c1_spec"], RTOL64, "golden c1 spectra")
    assert_features(ctx.stft_analysis(g["mp_x"], 2048, SR, 8), g["mp_feat"], RTOL64, "golden mp")
    for ch in range(2):
        S = ctx.rstft(g[f"c2_x{ch}"], 4096, 1024, orc.HAMMING)
        assert_spectra(S[0], g[f"c2_spec{ch}_first"], RTOL64, "golden c2 spectrum")
        y = ctx.irstft(S, 4096, 1024, orc.HAMMING)
        assert np.max(np.abs(y - g[f"c2_audio{ch}"])) <= RTOL64 * np.max(np.abs(g[f"c2_audio{ch}"]))
    lens = g["c3_lens"]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    got, _ = ctx.stft_features_batch(g["c3_x"], offs, lens, 1024, 256, orc.HANNING, SR)
    assert_features(got, g["c3_feat"], RTOL64, "golden c3")
    assert_features(ctx.stft_features(g["c4_x"], 16384, 4096, orc.HANNING, SR), g["c4_feat"], RTOL64, "golden c4")
    assert_features(ctx.analyzer_batch(g["an_mag"], 2048, SR), g["an_feat"], RTOL64, "golden analyzer")


def test_reference_named_free_functions(aus, orc):
    """The mirror modules (same names as the crate) run on the default context."""
    x = orc.synth_signal(6000, SR, seed=13)
    S = aus.spectrum.rstft(x, 1024, 512, aus.WindowType.Hamming)
    assert_spectra(S, orc.rstft(x, 1024, 512, orc.HAMMING), RTOL64, "spectrum.rstft")
    mag, ph = aus.spectrum.complex_to_polar_rstft(S)
    a = aus.analysis.analyzer(mag[2], 1024, SR)
    ref = orc.analyzer(np.abs(S[2]), 1024, SR)
    assert_features(a.as_array()[None], ref.as_array()[None], RTOL64, "analysis.analyzer")
    f = aus.spectrum.rfftfreq(1024, SR)
    assert aus.analysis.spectral_centroid(mag[2], f) == pytest.approx(ref.spectral_centroid, rel=RTOL64)
    assert aus.analysis.spectral_roll_off_point(mag[2], f, 0.9) == ref.spectral_roll_off_90
    assert aus.analysis.spectral_slope_region(mag[2], f, 1000.0, 5000.0, SR) == pytest.approx(ref.spectral_slope_1_5_khz, rel=RTOL64)
    res = aus.mp.stft_analysis(x, 1024, SR, 4)
    assert len(res) == orc.num_frames(x.size, 1024, 512)
    assert res[1].spectral_entropy == pytest.approx(orc.stft_analysis(x, 1024, SR)[1, 4], rel=RTOL64)
    back = aus.spectrum.irstft(aus.spectrum.polar_to_complex_rstft(mag, ph), 1024, 512, aus.WindowType.Hamming)
    assert np.max(np.abs(back - orc.irstft(S, 1024, 512, orc.HAMMING))) <= 1e-9


# ------------------------------------------------------------------------------------------------ sharded path
@pytest.mark.parametrize("n_dev", [2, 3])
def test_multi_device_entry_point(aus, orc, n_dev):
    """aus_stft_features_multi: host threads + shard plan.  On a 1-GPU box every slot maps to device 0."""
    import torch
    have = torch.cuda.device_count()
    devices = [i % have for i in range(n_dev)]
    x = orc.synth_signal(70000, SR, seed=14)
    got = aus.stft_features_multi(devices, x, [0], [x.size], 2048, 512, orc.HANNING, SR)
    assert_features(got, orc.stft_features(x, 2048, 512, orc.HANNING, SR), RTOL64, "multi single file")
    lens = [5000, 0, 12000, 2048, 30000]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    audio = orc.synth_signal(sum(lens), SR, seed=15)
    got = aus.stft_features_multi(devices, audio, offs, lens, 1024, 256, orc.HAMMING, SR)
    ref = np.concatenate([orc.stft_features(audio[o:o + n], 1024, 256, orc.HAMMING, SR) for o, n in zip(offs, lens)])
    assert_features(got, ref, RTOL64, "multi batch")


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_device_resident(ctx, orc):
    """BASELINE-sized inputs, checked through size-independent properties plus oracle spot checks:
    config 3 (1000 x 30 s, fft 1024 / hop 256) and config 2 (10 min, fft 4096 / hop 1024 round trip)."""
    import torch
    dev = torch.device("cuda", 0)
    ctx.use_torch_stream()          # torch ops below must be stream-ordered with the library's kernels
    # ---- config 3: one launch for 1000 clips
    n_clips, n = 1000, 30 * SR
    d = torch.empty(n_clips * n, dtype=torch.float64, device=dev)
    for i in range(n_clips):
        ctx.synth_dev(d[i * n:(i + 1) * n], 0, SR, orc.file_freq(i), 0, i, "f64")
    lens = np.full(n_clips, n, dtype=np.uint64)
    offs = np.arange(n_clips, dtype=np.uint64) * np.uint64(n)
    F1 = orc.num_frames(n, 1024, 256)
    out = torch.empty((n_clips * F1, 14), dtype=torch.float64, device=dev)
    ctx.stft_features_batch_dev(d, offs, lens, 1024, 256, orc.HANNING, SR, "f64", out)
    out2 = torch.empty_like(out)
    ctx.stft_features_batch_dev(d, offs, lens, 1024, 256, orc.HANNING, SR, "f64", out2)
    torch.cuda.synchronize()
    assert out.shape[0] == 5165000
    assert torch.equal(out, out2)                                   # deterministic: no atomics, fixed reduction order
    assert bool(torch.isfinite(out).all())
    roll = out[:, 6:10]
    assert bool((roll[:, 1:] >= roll[:, :-1]).all())                # roll-off points are ordered
    bins = roll / (SR / 1024)
    assert float((bins - bins.round()).abs().max()) < 1e-9          # ... and sit on the rfftfreq grid
    assert bool((out[:, 5] > 0).all()) and bool((out[:, 5] <= 1 + 1e-12).all())     # flatness in (0, 1]
    assert bool((out[:, 4] > 0).all()) and bool((out[:, 4] <= np.log2(513) + 1e-9).all())   # 0 < entropy <= log2 K
    for clip in (0, 417, 999):                                      # spot checks against the oracle (device-made input)
        x = d[clip * n:(clip + 1) * n].cpu().numpy()
        ref = orc.stft_features(x, 1024, 256, orc.HANNING, SR)
        assert_features(out[clip * F1:(clip + 1) * F1].cpu().numpy(), ref, RTOL64, f"clip {clip}")
    del d, out, out2
    # ---- config 2: linearity of the STFT and the round trip at full length
    n, N, hop = 600 * SR, 4096, 1024
    F, K = orc.num_frames(n, N, hop), N // 2 + 1
    xa, xb = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(2))
    ctx.synth_dev(xa, 0, SR, 440.0, 0, 100, "f64")
    ctx.synth_dev(xb, 0, SR, 880.0, 0, 101, "f64")
    l1, o1 = np.array([n], dtype=np.uint64), np.array([0], dtype=np.uint64)
    Sa, Sb, Sc = (torch.empty((F, K, 2), dtype=torch.float64, device=dev) for _ in range(3))
    ctx.stft_features_batch_dev(xa, o1, l1, N, hop, orc.HAMMING, SR, "f64", None, Sa)
    ctx.stft_features_batch_dev(xb, o1, l1, N, hop, orc.HAMMING, SR, "f64", None, Sb)
    xc = (0.75 * xa - 1.5 * xb).contiguous()
    ctx.stft_features_batch_dev(xc, o1, l1, N, hop, orc.HAMMING, SR, "f64", None, Sc)
    torch.cuda.synchronize()
    peak = float(Sc.abs().max())
    assert float((Sc - (0.75 * Sa - 1.5 * Sb)).abs().max()) <= 1e-12 * peak        # linearity
    y = torch.empty(N + hop * (F - 1), dtype=torch.float64, device=dev)
    ctx.irstft_dev(Sa, F, N, hop, orc.HAMMING, "f64", y, True)
    torch.cuda.synchronize()
    assert float(y.max()) == pytest.approx(10 ** (-6 / 20), rel=1e-12)               # peak-normalised to -6 dBFS
    body = slice(N, n - 2 * N)
    gain = float((y[body] * xa[body]).sum() / (xa[body] * xa[body]).sum())
    err = float((y[body] / gain - xa[body]).abs().max())
    assert 0.0 < err < 1e-9, err                                                     # reconstructs up to the gain
    ref0 = orc.rstft(xa[:3 * N].cpu().numpy(), N, hop, orc.HAMMING)[:2]              # first frames against the oracle
    got0 = torch.view_as_complex(Sa[:2].contiguous()).cpu().numpy()
    assert_spectra(got0, ref0, RTOL64, "config 2 first frames")


def test_determinism_all_kernels(ctx, orc):
    """No atomics on the data path and fixed reduction orders: repeated runs are bit-identical (also a race detector)."""
    for N, hop in ((64, 16), (1024, 256), (2048, 512), (4096, 1024), (16384, 4096)):
        x = orc.synth_signal(40 * hop + N + 7, SR, seed=17, file_id=N)
        for prec in ("f64", "f32"):
            f1, f2 = (ctx.stft_features(x, N, hop, orc.HAMMING, SR, prec) for _ in range(2))
            assert np.array_equal(f1, f2, equal_nan=True)
            S1, S2 = (ctx.rstft(x, N, hop, orc.HAMMING, prec) for _ in range(2))
            assert np.array_equal(S1, S2)
            y1, y2 = (ctx.irstft(S1, N, hop, orc.HAMMING, prec) for _ in range(2))
            assert np.array_equal(y1, y2, equal_nan=True)
            a1, a2 = (ctx.analyzer_batch(np.abs(S1), N, SR, prec) for _ in range(2))
            assert np.array_equal(a1, a2, equal_nan=True)


def test_features_and_spectra_in_one_call(ctx, orc):
    """aus_stft_features_batch_dev with both outputs (the mixed path) agrees with the two single-output paths."""
    import torch
    ctx.use_torch_stream()
    x = orc.synth_signal(30000, SR, seed=18)
    d = torch.from_numpy(x).cuda()
    F = orc.num_frames(x.size, 2048, 512)
    feat = torch.empty((F, 14), dtype=torch.float64, device="cuda")
    spec = torch.empty((F, 1025, 2), dtype=torch.float64, device="cuda")
    ctx.stft_features_batch_dev(d, [0], [x.size], 2048, 512, orc.HANNING, SR, "f64", feat, spec)
    torch.cuda.synchronize()
    assert_features(feat.cpu().numpy(), orc.stft_features(x, 2048, 512, orc.HANNING, SR), RTOL64, "mixed features")
    assert_spectra(torch.view_as_complex(spec).cpu().numpy(), orc.rstft(x, 2048, 512, orc.HANNING), RTOL64, "mixed spectra")


def test_host_pipeline_chunking(aus, orc):
    """The host-pointer feature path cuts its input into chunks (copy/compute overlap).  With a tiny chunk size every
    code path of the cut is exercised: long files split into frame ranges with halo, short files grouped, empty files."""
    c = aus.Context(0)
    x = orc.synth_signal(90000, SR, seed=19)
    ref = orc.stft_features(x, 2048, 512, orc.HANNING, SR)
    for chunk in (1000, 5000, 33333, 10 ** 9):
        c.set_chunk_samples(chunk)
        assert np.array_equal(c.stft_features(x, 2048, 512, orc.HANNING, SR), c.stft_features(x, 2048, 512, orc.HANNING, SR))
        assert_features(c.stft_features(x, 2048, 512, orc.HANNING, SR), ref, RTOL64, f"chunk {chunk}")
        assert_features(c.stft_features(x, 2048, 512, orc.HANNING, SR, "f32")[:, :6], ref[:, :6], RTOL32, f"chunk {chunk} f32")
    lens = [0, 7000, 100, 0, 20000, 2048, 15000, 1]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    audio = orc.synth_signal(sum(lens), SR, seed=20)
    refb = np.concatenate([orc.stft_features(audio[o:o + n], 1024, 256, orc.HAMMING, SR) for o, n in zip(offs, lens)])
    for chunk in (3000, 9000, 10 ** 9):
        c.set_chunk_samples(chunk)
        got, starts = c.stft_features_batch(audio, offs, lens, 1024, 256, orc.HAMMING, SR)
        assert_features(got, refb, RTOL64, f"batch chunk {chunk}")
    c.close()


def test_rolloff_knife_edges_match_reference_arithmetic(ctx, orc):
    """Spectra built so that a roll-off threshold coincides with a running sum in exact arithmetic: the answer then
    depends on how `cum += p[i] / P` rounds (s_a_t.rs:137-145).  The device path must reproduce that, not merely be
    'close': e.g. 40 equal bins -> the 20th partial sum of 1/40 is 0.5000000000000001 or 0.49999999999999994."""
    N, K = 2048, 1025
    rows = []
    rng = np.random.default_rng(5)
    for nz in (2, 4, 10, 20, 40, 100, 200, 400, 1000):
        for amp in (1.0, 3.0, 0.1):
            m = np.zeros(K)
            idx = np.sort(rng.choice(K, nz, replace=False))
            m[idx] = amp
            rows.append(m)
    rows.append(np.full(K, 2.0))
    m = np.zeros(K); m[::2] = 1.0; rows.append(m)
    mags = np.array(rows)
    ref = orc.analyze_spectrogram(mags, N, SR)
    got = ctx.analyzer_batch(mags, N, SR)
    assert np.array_equal(got[:, ROLL], ref[:, ROLL]), np.argwhere(got[:, ROLL] != ref[:, ROLL])
    # the remaining features of these degenerate rows: zero bins -> entropy NaN, flatness 0 (same NaN pattern)
    assert np.array_equal(np.isnan(got), np.isnan(ref))


def test_pcm_input_matches_reader_then_analyzer(ctx, orc):
    """integer PCM widened on the device as `sample as f64 / INTMAX` (audiofile.rs:296-329)."""
    rng = np.random.default_rng(21)
    lens = [9000, 0, 30000, 2049]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    x = orc.synth_signal(sum(lens), SR, seed=22)
    for dtype, bits, intmax in ((np.int16, 16, 2 ** 15 - 1), (np.int32, 24, 2 ** 23 - 1), (np.int32, 32, 2 ** 31 - 1)):
        pcm = np.round(x * intmax).astype(dtype)
        wide = pcm.astype(np.float64) / float(intmax)
        ref = np.concatenate([orc.stft_features(wide[o:o + n], 2048, 512, orc.HANNING, SR) for o, n in zip(offs, lens)])
        got, starts = ctx.stft_features_batch_pcm(pcm, offs, lens, 2048, 512, orc.HANNING, SR, sample_bits=bits)
        assert_features(got, ref, RTOL64, f"pcm{bits}")
        # identical to uploading the widened f64 samples (the conversion itself is bit-exact)
        got64, _ = ctx.stft_features_batch(wide, offs, lens, 2048, 512, orc.HANNING, SR)
        assert np.array_equal(got, got64, equal_nan=True)


def test_next_rows_autocorrelation_and_time_features(ctx, aus, orc):
    """analysis::autocorrelation and dc_bias / dbfs_max / energy / zero_crossing_rate."""
    for N in (64, 1024, 2048, 8192):
        x = orc.synth_signal(3 * N, SR, seed=23, file_id=N).reshape(3, N)
        ref = np.stack([orc.autocorrelation(r, N) for r in x])
        got = ctx.autocorrelation(x, N)
        assert np.max(np.abs(got - ref)) <= RTOL64 * np.max(np.abs(ref))
        assert np.max(np.abs(ctx.autocorrelation(x[1], N) - ref[1])) <= RTOL64 * np.max(np.abs(ref))
    with pytest.raises(aus.SpectrumError) as e:
        ctx.autocorrelation(np.zeros(100), 128)
    assert e.value.error_msg == "The audio length does not match the FFT size."
    lens = [70001, 1, 2, 3, 200000, 4097]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    a = orc.synth_signal(sum(lens), SR, seed=24)
    a[offs[4] + 10:offs[4] + 20] = 0.0                    # exact zeros: the second zero-crossing rule
    a[offs[4] + 20] = -0.3
    a[offs[4] + 65535:offs[4] + 65538] = [0.2, 0.0, -0.1]    # ... across a partial-sum chunk boundary
    got = ctx.time_features_batch(a, offs, lens, SR)
    for i, (o, n) in enumerate(zip(offs, lens)):
        seg = a[o:o + n]
        ref = np.array([orc.dc_bias(seg), orc.dbfs_max(seg), orc.energy(seg), orc.zero_crossing_rate(seg, SR)])
        assert got[i, 3] == ref[3], (i, got[i], ref)                      # a count: exact
        assert np.all(np.abs(got[i, :3] - ref[:3]) <= RTOL64 * np.abs(ref[:3])), (i, got[i], ref)
    assert aus.analysis.energy(a[:5000]) == pytest.approx(orc.energy(a[:5000]), rel=RTOL64)
    assert aus.analysis.dbfs_max(np.zeros(10)) == -np.inf


def test_randomised_parameter_sweep(ctx, orc):
    """Random fft sizes, hops (odd, 1, = fft_size), windows and lengths: features (f64), spectra and inverse (f64, f32)."""
    import os
    rng = np.random.default_rng(int(os.environ.get("AUS_SWEEP_SEED", "2026")))     # soak runs: more trials, other seeds
    sizes = [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
    for trial in range(int(os.environ.get("AUS_SWEEP_TRIALS", "36"))):
        N = int(sizes[trial % len(sizes)])
        hop = int(rng.choice([1, 3, N // 4, N // 2, N, int(rng.integers(1, N + 1)), N // 3 + 1]))
        if hop < 8 and N > 512:
            hop = N // 8 + int(rng.integers(0, 5))          # keep the oracle's frame count small
        window = int(rng.integers(0, 5))
        n = int(rng.integers(0, 4 * N + 50)) if hop >= 8 else int(rng.integers(0, N + 40))
        x = orc.synth_signal(n, SR, freq=float(rng.uniform(50, 8000)), seed=100 + trial, file_id=trial)
        what = f"trial {trial}: N={N} hop={hop} w={window} n={n}"
        Sref = orc.rstft(x, N, hop, window)
        S = ctx.rstft(x, N, hop, window)
        assert S.shape == Sref.shape, what
        assert np.array_equal(np.isnan(S), np.isnan(Sref)), what
        assert_spectra(np.nan_to_num(S), np.nan_to_num(Sref), RTOL64, what)
        S32 = ctx.rstft(x, N, hop, window, "f32")
        assert np.array_equal(np.isnan(S32), np.isnan(Sref)), what
        assert_spectra(np.nan_to_num(S32), np.nan_to_num(Sref), RTOL32, what + " f32")
        if not np.isnan(Sref).any():
            sc = cancellation_scales(np.abs(Sref), N, SR)
            assert_features(ctx.stft_features(x, N, hop, window, SR), orc.stft_features(x, N, hop, window, SR), RTOL64, what, scales=sc)
            if window in (3, 4) and Sref.shape[0] * N < 400000:          # windows without (near-)zero end points
                ref = orc.irstft(Sref, N, hop, window, normalize=False)
                got = ctx.irstft(Sref, N, hop, window, normalize=False)
                assert got.shape == ref.shape, what
                assert np.max(np.abs(got - ref)) <= RTOL64 * max(np.max(np.abs(ref)), 1e-300), what
                got32 = ctx.irstft(Sref, N, hop, window, "f32", normalize=False)
                assert np.max(np.abs(got32 - ref)) <= RTOL32 * max(np.max(np.abs(ref)), 1e-300), what + " f32"


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("fft_size,hop", [(64, 16), (128, 32), (256, 64), (512, 128), (1024, 256), (2048, 512), (2048, 300),
                                          (4096, 1024), (8192, 2048), (16384, 4096)])
def test_forward_kernel_variants(ctx, orc, variant, fft_size, hop):
    """The forward kernels -- shared-memory tail (0), register-resident tail (1, aus_stft_r.cuh) and the producer/consumer
    build (2, aus_stft_pc.cuh: f64 features-only launches at fft_size 2048, the default choice elsewhere) -- against the
    oracle, f64 and f32, features + spectra in one call, ragged files incl. silence and a short last frame."""
    import torch
    ctx.use_torch_stream()
    ctx.set_forward_variant(variant)
    try:
        lens = [5 * fft_size + 13, 0, fft_size, 3 * fft_size + hop + 1, 7]
        sig = [orc.synth_signal(n, SR, seed=31, file_id=i) for i, n in enumerate(lens)]
        sig[3][: 2 * fft_size] = 0.0                     # silent frames: NaN pattern of the reference
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
        audio = np.concatenate(sig)
        K = fft_size // 2 + 1
        refs_f = [orc.stft_features(s, fft_size, hop, orc.HAMMING, SR) for s in sig]
        refs_s = [orc.rstft(s, fft_size, hop, orc.HAMMING) for s in sig]
        F = sum(r.shape[0] for r in refs_f)
        for prec, tdt, tol in (("f64", torch.float64, RTOL64), ("f32", torch.float32, RTOL32)):
            d = torch.from_numpy(audio).to(tdt).cuda()
            feat = torch.empty((F, 14), dtype=tdt, device="cuda")
            spec = torch.empty((F, K, 2), dtype=tdt, device="cuda")
            ctx.stft_features_batch_dev(d, offs, lens, fft_size, hop, orc.HAMMING, SR, prec, feat, spec)
            torch.cuda.synchronize()
            got_f = feat.cpu().numpy().astype(np.float64)
            got_s = torch.view_as_complex(spec).cpu().numpy().astype(np.complex128)
            # the single-output launches (features only; spectra only = the build that copies the next frame in early)
            feat1 = torch.zeros_like(feat)
            spec1 = torch.zeros_like(spec)
            ctx.stft_features_batch_dev(d, offs, lens, fft_size, hop, orc.HAMMING, SR, prec, feat1, None)
            ctx.stft_features_batch_dev(d, offs, lens, fft_size, hop, orc.HAMMING, SR, prec, None, spec1)
            torch.cuda.synchronize()
            rt = 1e-11 if prec == "f64" else 2e-5
            pk = np.abs(got_s).max(axis=1, keepdims=True) + 1e-300
            s1 = torch.view_as_complex(spec1).cpu().numpy().astype(np.complex128)
            assert np.array_equal(np.isnan(s1), np.isnan(got_s))
            assert np.nanmax(np.abs(np.nan_to_num(s1) - np.nan_to_num(got_s)) / pk, initial=0.0) <= rt
            f1 = feat1.cpu().numpy().astype(np.float64)
            assert np.array_equal(np.isnan(f1), np.isnan(got_f))
            if prec == "f64":
                assert_features(f1, got_f, 1e-9, f"variant {variant} {prec} features-only vs mixed launch",
                                scales=cancellation_scales(np.abs(np.nan_to_num(got_s)), fft_size, SR))
            pos = 0
            for i, (rf, rs) in enumerate(zip(refs_f, refs_s)):
                gs, gf = got_s[pos:pos + rf.shape[0]], got_f[pos:pos + rf.shape[0]]
                pos += rf.shape[0]
                assert np.array_equal(np.isnan(gs), np.isnan(rs))
                assert_spectra(np.nan_to_num(gs), np.nan_to_num(rs), tol, f"variant {variant} {prec} file {i}")
                if prec == "f64":
                    sc = cancellation_scales(np.abs(np.nan_to_num(rs)), fft_size, SR)
                    assert_features(gf, rf, tol, f"variant {variant} f64 file {i}", scales=sc)
                else:
                    ok = np.isfinite(rf).all(axis=1)
                    cols = [c for c in range(14) if c not in range(6, 10)]
                    sc = cancellation_scales(np.abs(np.nan_to_num(rs)), fft_size, SR)
                    assert_features(gf[ok][:, cols], rf[ok][:, cols], tol, f"variant {variant} f32 file {i}",
                                    scales=sc[ok][:, cols])
    finally:
        ctx.set_forward_variant(-1)


def test_single_feature_wrappers_generic_parameters(ctx, aus, orc):
    """spectral_analysis_tools.rs:306-471 with the reference's full parameter space: rows of any length, a non-uniform
    frequency table, arbitrary roll-off fractions and slope bands (the doc example's 105-852 Hz), zero bins."""
    rng = np.random.default_rng(11)
    for K, sr in ((1025, 44100), (513, 48000), (100, 8000), (2, 44100), (4097, 22050)):
        m = np.abs(rng.normal(size=K)) + 1e-3
        f_uniform = orc.rfftfreq(2 * (K - 1), sr) if K > 2 else np.array([0.0, sr / 2.0])
        f_warped = np.sort(rng.uniform(0.0, sr / 2.0, K))               # any table is legal in the reference
        for f in (f_uniform, f_warped):
            A = aus.analysis
            assert A.spectral_centroid(m, f) == pytest.approx(orc.spectral_centroid(m, f), rel=RTOL64)
            assert A.spectral_variance(m, f) == pytest.approx(orc.spectral_variance(m, f), rel=RTOL64)
            assert A.spectral_skewness(m, f) == pytest.approx(orc.spectral_skewness(m, f), rel=1e-7, abs=1e-9)
            assert A.spectral_kurtosis(m, f) == pytest.approx(orc.spectral_kurtosis(m, f), rel=RTOL64)
            for n in (0.3, 0.5, 0.85, 0.999, 1.0):
                assert A.spectral_roll_off_point(m, f, n) == orc.spectral_roll_off_point(m, f, n)
        assert aus.analysis.spectral_entropy(m) == pytest.approx(orc.spectral_entropy(m), rel=RTOL64)
        assert aus.analysis.spectral_flatness(m) == pytest.approx(orc.spectral_flatness(m), rel=RTOL64)
        assert aus.analysis.spectral_slope(m) == pytest.approx(orc.spectral_slope(m), rel=1e-7, abs=1e-12)
        if K >= 100:
            fund = sr / (2.0 * (K - 1))
            for fl, fu in ((105.0, 852.0), (0.0, 3.5 * fund), (2 * fund, 40 * fund), (1000.0, 1000.0 + 10.25 * fund)):
                if fu / fund >= K - 1:
                    continue
                got = aus.analysis.spectral_slope_region(m, f_uniform, fl, fu, sr)
                ref = orc.spectral_slope_region(m, f_uniform, fl, fu, sr)
                assert got == pytest.approx(ref, rel=1e-7, abs=1e-9 * float(np.max(m * m))), (K, fl, fu)
            with pytest.raises(aus.AusError):                                  # the reference indexes out of bounds
                aus.analysis.spectral_slope_region(m, f_uniform, 0.0, sr / 2.0 + 1.5 * fund, sr)
            got = aus.analysis.spectral_slope_region(m, f_uniform, 0.0, sr / 2.0, sr)     # up to the last bin is legal
            assert got == pytest.approx(orc.spectral_slope_region(m, f_uniform, 0.0, sr / 2.0, sr), rel=1e-7,
                                        abs=1e-9 * float(np.max(m * m)))
    # zero bins: entropy NaN (0 * -inf), flatness 0 (exp(-inf)), like the reference
    m = np.abs(rng.normal(size=1025)) + 1e-3
    m[7] = 0.0
    assert np.isnan(aus.analysis.spectral_entropy(m)) and np.isnan(orc.spectral_entropy(m))
    assert aus.analysis.spectral_flatness(m) == 0.0 == orc.spectral_flatness(m)
    # exact ties: flat rows put thresholds exactly on (or within rounding of) a running sum; the reference's sequential
    # arithmetic decides and the device must reproduce its index, not a neighbour's
    for K in (64, 96, 1000, 1025):
        m = np.full(K, 0.37)
        f = np.arange(K, dtype=np.float64) * 3.0
        for n in (1.0 / 64, 0.25, 0.5, 0.75, 1.0 - 1.0 / 64, 1.0):
            assert aus.analysis.spectral_roll_off_point(m, f, n) == orc.spectral_roll_off_point(m, f, n), (K, n)
    # many rows in one call
    M = np.abs(rng.normal(size=(300, 513))) + 1e-3
    f = orc.rfftfreq(1024, 44100)
    got = ctx.spectral_features_generic(M, f, (0.25, 0.9), ((105.0, 852.0),), 44100)
    assert got.shape == (300, 10)
    for r in (0, 150, 299):
        assert got[r, 0] == pytest.approx(orc.spectral_centroid(M[r], f), rel=RTOL64)
        assert got[r, 7] == orc.spectral_roll_off_point(M[r], f, 0.25) and got[r, 8] == orc.spectral_roll_off_point(M[r], f, 0.9)
        assert got[r, 9] == pytest.approx(orc.spectral_slope_region(M[r], f, 105.0, 852.0, 44100), rel=1e-7)
    # more rows than one grid pass covers (the kernel walks rows warp by warp), every row checked on two columns
    M = np.abs(rng.normal(size=(40000, 65))) + 1e-3
    f = orc.rfftfreq(128, 8000)
    got = ctx.spectral_features_generic(M, f, (0.5,), (), 8000)
    assert got.shape == (40000, 8)
    cen = (M @ f) / M.sum(axis=1)
    np.testing.assert_allclose(got[:, 0], cen, rtol=1e-12)
    P = M * M
    cum = np.cumsum(P / P.sum(axis=1, keepdims=True), axis=1)
    idx = np.minimum((cum < 0.5).sum(axis=1), 64)
    assert np.array_equal(got[:, 7], f[idx])


@pytest.mark.parametrize("fft_size,hop", [(1024, 256), (2048, 512), (4096, 1024)])
def test_direct_load_knob_is_bit_identical(ctx, orc, fft_size, hop):
    """aus_ctx_set_direct_load: TMA staging vs global->register loads of pass 1 give identical bits (same arithmetic),
    for every forward variant, features and spectra."""
    import torch
    ctx.use_torch_stream()
    x = orc.synth_signal(9 * fft_size + 5, SR, seed=41, file_id=fft_size)
    d = torch.from_numpy(x).cuda()
    F, K = orc.num_frames(x.size, fft_size, hop), fft_size // 2 + 1
    res = {}
    try:
        for variant in (0, 1):
            ctx.set_forward_variant(variant)
            for direct in (False, True):
                ctx.set_direct_load(direct)
                feat = torch.empty((F, 14), dtype=torch.float64, device="cuda")
                spec = torch.empty((F, K, 2), dtype=torch.float64, device="cuda")
                spec_only = torch.empty((F, K, 2), dtype=torch.float64, device="cuda")
                ctx.stft_features_batch_dev(d, [0], [x.size], fft_size, hop, orc.HANNING, SR, "f64", feat, spec)
                ctx.stft_features_batch_dev(d, [0], [x.size], fft_size, hop, orc.HANNING, SR, "f64", None, spec_only)
                torch.cuda.synchronize()
                res[(variant, direct)] = (feat.cpu().numpy(), spec.cpu().numpy(), spec_only.cpu().numpy())
            for a, b in zip(res[(variant, False)], res[(variant, True)]):
                assert np.array_equal(a, b, equal_nan=True)
        assert_spectra(res[(1, True)][2][..., 0] + 1j * res[(1, True)][2][..., 1], orc.rstft(x, fft_size, hop, orc.HANNING),
                       RTOL64, "direct load, spectra only")
    finally:
        ctx.set_direct_load(False)
        ctx.set_forward_variant(-1)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("fft_size,hop", [(256, 64), (2048, 512), (4096, 1000)])
def test_device_entry_points_stay_inside_their_buffers(ctx, orc, prec, fft_size, hop):
    """Every device-pointer entry point, with its inputs and outputs carved out of one arena at 0-, 1- and 3-element
    offsets (16-byte aligned, 8-byte aligned, odd: the bulk-copy, vector and scalar staging paths): the guard bands
    around every output keep their fill value, and the results do not depend on the alignment."""
    import torch
    ctx.use_torch_stream()
    dt = torch.float64 if prec == "f64" else torch.float32
    N, K, FILL = fft_size, fft_size // 2 + 1, 1234.5
    x = orc.synth_signal(37 * hop + N + 11, SR, seed=23, file_id=N)
    n, F = x.size, orc.num_frames(x.size, N, hop)
    L = N + hop * (F - 1)
    sizes = {"audio": n, "feat": F * 14, "spec": F * K * 2, "mags": F * K, "feat2": F * 14, "rec": L}
    results = []
    for shift in (0, 1, 3):
        G = 64 + shift
        arena = torch.full((sum(sizes.values()) + G * (len(sizes) + 1),), FILL, dtype=dt, device="cuda")
        view, guards, pos = {}, [], 0
        for name, sz in sizes.items():
            guards.append(arena[pos:pos + G + ((pos + G) & 1 if name == "spec" else 0)])   # spectra: whole (re, im) pairs
            pos += guards[-1].numel()
            view[name] = arena[pos:pos + sz]
            pos += sz
        guards.append(arena[pos:])
        view["audio"].copy_(torch.from_numpy(x).to(dt))
        ctx.stft_features_batch_dev(view["audio"], [0], [n], N, hop, orc.HAMMING, SR, prec, view["feat"], view["spec"])
        view["mags"].copy_(torch.view_as_complex(view["spec"].clone().view(F * K, 2)).abs())
        ctx.analyzer_batch_dev(view["mags"], F, N, SR, prec, view["feat2"])
        ctx.irstft_dev(view["spec"], F, N, hop, orc.HAMMING, prec, view["rec"], True)
        torch.cuda.synchronize()
        for i, g in enumerate(guards):
            assert bool((g == FILL).all()), f"guard band {i} overwritten (shift {shift})"
        results.append({k: view[k].cpu().numpy().copy() for k in ("feat", "spec", "feat2", "rec")})
    tol = 1e-12 if prec == "f64" else 2e-6
    for r in results[1:]:
        assert np.array_equal(r["feat"], results[0]["feat"], equal_nan=True)          # staging differs, arithmetic does not
        assert np.array_equal(r["spec"], results[0]["spec"])
        assert np.array_equal(r["feat2"], results[0]["feat2"], equal_nan=True)
        # irstft: the aligned fast path multiplies by a reciprocal norm where the generic path divides
        assert np.max(np.abs(r["rec"] - results[0]["rec"])) <= tol * max(1.0, np.max(np.abs(results[0]["rec"])))
    odd = torch.empty(2 * F * K + 1, dtype=dt, device="cuda")[1:]                    # spectra not on a (re, im) boundary
    with pytest.raises(Exception, match="aligned"):
        ctx.irstft_dev(odd, F, N, hop, orc.HAMMING, prec, view["rec"], True)
    with pytest.raises(Exception, match="aligned"):
        ctx.stft_features_batch_dev(view["audio"], [0], [n], N, hop, orc.HAMMING, SR, prec, None, odd)


# ------------------------------------------------------------------------------------------------ round 2 additions
def test_producer_consumer_kernel_edge_cases(aus, orc):
    """aus_stft_pc.cuh on its own context (variant 2): every window, odd hops / unaligned starts (cooperative staging),
    silence, zero bins, a short last frame, one-frame files, and roll-off knife edges through the fused path."""
    c = aus.Context(0)
    c.set_forward_variant(2)
    try:
        for window in range(5):
            x = orc.synth_signal(40000 + 17 * window, SR, seed=51, file_id=window)
            assert_features(c.stft_features(x, 2048, 512, window, SR), orc.stft_features(x, 2048, 512, window, SR), RTOL64,
                            f"pc window {window}")
        for hop, n in ((511, 30001), (1, 2300), (2048, 10000), (300, 2047), (512, 2049)):
            x = orc.synth_signal(n, SR, seed=52, file_id=hop)
            S = orc.rstft(x, 2048, hop, orc.HAMMING)
            sc = cancellation_scales(np.abs(np.nan_to_num(S)), 2048, SR)
            assert_features(c.stft_features(x, 2048, hop, orc.HAMMING, SR), orc.stft_features(x, 2048, hop, orc.HAMMING, SR), RTOL64,
                            f"pc hop {hop}", scales=sc)
        x = np.zeros(9000)
        x[6000:] = orc.synth_signal(3000, SR, seed=53)
        assert_features(c.stft_features(x, 2048, 512, orc.HANNING, SR), orc.stft_features(x, 2048, 512, orc.HANNING, SR), RTOL64, "pc silence")
        # a tone on a bin centre with a rectangular window over a -50 dB floor: one bin carries nearly all the power (entropy
        # ~1e-4 bits: this kernel forms it as log2 P - S / P, whose absolute error is ~4e-15, see aus_stft_pc.cuh)
        n = np.arange(8192)
        tone = np.sin(2 * np.pi * 64 * n / 2048) + 3e-2 * orc.synth_signal(8192, SR, seed=54)
        ref = orc.stft_features(tone, 2048, 512, orc.RECTANGULAR, SR)
        got = c.stft_features(tone, 2048, 512, orc.RECTANGULAR, SR)
        sc = cancellation_scales(np.abs(orc.rstft(tone, 2048, 512, orc.RECTANGULAR)), 2048, SR)
        assert np.array_equal(got[:, ROLL], ref[:, ROLL])
        assert_features(got, ref, RTOL64, "pc tone over a floor", scales=sc)
        lens = [2048, 0, 1, 2049, 5000, 2047]
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
        a = orc.synth_signal(sum(lens), SR, seed=55)
        got, starts = c.stft_features_batch(a, offs, lens, 2048, 512, orc.BLACKMAN, SR)
        for i, (o, ln) in enumerate(zip(offs, lens)):
            S = orc.rstft(a[o:o + ln], 2048, 512, orc.BLACKMAN)
            sc = cancellation_scales(np.abs(np.nan_to_num(S)), 2048, SR)
            assert_features(got[starts[i]:starts[i + 1]], orc.stft_features(a[o:o + ln], 2048, 512, orc.BLACKMAN, SR), RTOL64,
                            f"pc file {i}", scales=sc)
    finally:
        c.close()


def test_irstft_batch_and_frame_range_shards(ctx, aus, orc):
    """aus_irstft_batch (several spectrograms, one launch; config 2's stereo pair) and aus_irstft_multi (frame-range shards
    with recomputed halo + host max over the shard peaks).  On a 1-GPU box every slot maps to device 0; the
    shard arithmetic is the same."""
    import torch
    have = torch.cuda.device_count()
    for N, hop, window in ((4096, 1024, orc.HAMMING), (2048, 512, orc.HANNING), (1024, 1000, orc.HAMMING), (256, 64, orc.BARTLETT)):
        xs = [orc.synth_signal(n, SR, seed=61, file_id=i) for i, n in enumerate((37 * hop + N + 5, N, 9 * N + 1))]
        S = [orc.rstft(x, N, hop, window) for x in xs]
        ref = [orc.irstft(s, N, hop, window) for s in S]
        got = ctx.irstft_batch(S, N, hop, window)
        for g, r in zip(got, ref):
            assert g.shape == r.shape and np.array_equal(np.isnan(g), np.isnan(r))
            fin = np.isfinite(r)
            assert np.max(np.abs(g[fin] - r[fin])) <= RTOL64 * np.max(np.abs(r[fin]))
        for n_dev in (1, 2, 3, 5):
            devices = [i % have for i in range(n_dev)]
            for specs, refs in ((S[:1], ref[:1]), (S, ref)):               # one long file by frame range; a batch
                got = aus.irstft_multi(devices, specs, N, hop, window)
                for g, r in zip(got, refs):
                    assert g.shape == r.shape and np.array_equal(np.isnan(g), np.isnan(r)), (N, n_dev)
                    fin = np.isfinite(r)
                    assert np.max(np.abs(g[fin] - r[fin])) <= RTOL64 * np.max(np.abs(r[fin])), (N, n_dev)
        # the un-normalised output of a sharded run is BIT-identical to the single-device one (same frames, same order)
        one = aus.irstft_multi([0], S[:1], N, hop, window, normalize=False)[0]
        four = aus.irstft_multi([i % have for i in range(4)], S[:1], N, hop, window, normalize=False)[0]
        assert np.array_equal(one, four, equal_nan=True)
    aus.multi_shutdown()


def test_multi_gpu_entry_points_on_distinct_devices(aus, orc):
    """The product's own multi-GPU entry points with every slot on a DIFFERENT device (skipped on a 1-GPU box): row order
    across devices, frame-range halo, host gather."""
    import torch
    have = torch.cuda.device_count()
    if have < 2:
        pytest.skip("needs at least 2 GPUs")
    devices = list(range(min(have, 8)))
    x = orc.synth_signal(40 * 4096 + 1234, SR, seed=71)
    got = aus.stft_features_multi(devices, x, [0], [x.size], 16384, 4096, orc.HANNING, SR)            # config-4 shape, by frame range
    assert_features(got, orc.stft_features(x, 16384, 4096, orc.HANNING, SR), RTOL64, "multi single file, distinct devices")
    lens = [30 * 441, 0, 12000, 2048, 30000, 5, 44100, 7000, 1024, 9999]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    audio = orc.synth_signal(sum(lens), SR, seed=72)
    got = aus.stft_features_multi(devices, audio, offs, lens, 1024, 256, orc.HANNING, SR)                  # config-3 shape, by file
    ref = np.concatenate([orc.stft_features(audio[o:o + n], 1024, 256, orc.HANNING, SR) for o, n in zip(offs, lens)])
    assert_features(got, ref, RTOL64, "multi batch, distinct devices")
    S = [orc.rstft(orc.synth_signal(50 * 1024 + 4096, SR, seed=73, file_id=ch), 4096, 1024, orc.HAMMING) for ch in range(2)]
    for specs in (S[:1], S):
        got = aus.irstft_multi(devices, specs, 4096, 1024, orc.HAMMING)
        for g, s in zip(got, specs):
            r = orc.irstft(s, 4096, 1024, orc.HAMMING)
            assert np.max(np.abs(g - r)) <= RTOL64 * np.max(np.abs(r))
    aus.multi_shutdown()


def test_polar_edit_loop_stays_on_device(ctx, orc):
    """rstft -> complex_to_polar -> edit -> polar_to_complex -> irstft (fft.rs:288-291) with device pointers
    only; the one host copy is the final comparison."""
    import torch
    ctx.use_torch_stream()
    N, hop = 1024, 256
    x = orc.synth_signal(30000, SR, seed=81)
    F, K = orc.num_frames(x.size, N, hop), N // 2 + 1
    d_x = torch.from_numpy(x).cuda()
    spec = torch.empty((F, K, 2), dtype=torch.float64, device="cuda")
    mag, ph = (torch.empty((F, K), dtype=torch.float64, device="cuda") for _ in range(2))
    ctx.stft_features_batch_dev(d_x, [0], [x.size], N, hop, orc.HAMMING, SR, "f64", None, spec)
    ctx.complex_to_polar_dev(spec, mag, ph)
    gain = torch.linspace(1.0, 0.25, K, dtype=torch.float64, device="cuda")
    mag *= gain                                                        # the "edit": a spectral tilt, on the device
    spec2 = torch.empty_like(spec)
    ctx.polar_to_complex_dev(mag, ph, spec2)
    y = torch.empty(N + hop * (F - 1), dtype=torch.float64, device="cuda")
    ctx.irstft_dev(spec2, F, N, hop, orc.HAMMING, "f64", y, True)
    torch.cuda.synchronize()
    S = orc.rstft(x, N, hop, orc.HAMMING)
    m_ref, p_ref = orc.complex_to_polar_rstft(S)
    assert np.max(np.abs(ph.cpu().numpy() - p_ref)) <= 1e-12
    ref = orc.irstft(orc.polar_to_complex_rstft(m_ref * np.linspace(1.0, 0.25, K), p_ref), N, hop, orc.HAMMING)
    assert np.max(np.abs(y.cpu().numpy() - ref)) <= RTOL64 * np.max(np.abs(ref))
    mag_only = torch.empty_like(mag)
    ctx.complex_to_polar_dev(spec, mag_only, None)                     # phases are optional (the analyzer path drops them)
    torch.cuda.synchronize()
    assert np.max(np.abs(mag_only.cpu().numpy() - m_ref)) <= RTOL64 * np.max(m_ref)


def test_pcm_widths_and_mixdown_are_bit_exact(ctx, orc):
    """`sample as f64 / INTMAX` for S8 / S16 / S24 / S32 (audiofile.rs:296-329) and `mixdown` with the
    reference's double count of channel 0 (audiofile.rs:168-178): integer -> f64 division and the adds are exact ops, so
    the device must reproduce the oracle bit for bit."""
    rng = np.random.default_rng(91)
    for dtype, bits in ((np.int8, 8), (np.int16, 16), (np.int32, 24), (np.int32, 32)):
        lim = 2 ** (bits - 1)
        for n_ch in (1, 2, 3, 6):
            pcm = rng.integers(-lim, lim, size=(4097, n_ch)).astype(dtype)           # interleaved [frame][channel]
            wide = orc.pcm_to_f64(pcm.T, bits)
            assert np.array_equal(ctx.pcm_to_f64(pcm, n_ch, True, False, sample_bits=bits), wide)
            assert np.array_equal(ctx.pcm_to_f64(np.ascontiguousarray(pcm.T), n_ch, False, False, sample_bits=bits), wide)
            assert np.array_equal(ctx.pcm_to_f64(pcm, n_ch, True, True, sample_bits=bits), orc.mixdown(wide))
            assert np.array_equal(ctx.mixdown(wide), orc.mixdown(wide))
    two = np.array([[0.25, -0.5, 1.0], [0.5, 0.25, -1.0]])
    assert np.array_equal(ctx.mixdown(two), (two[0] + two[0] + two[1]) / 2.0)          # channel 0 twice, as in the reference
    assert np.array_equal(ctx.mixdown(two[0]), two[0])                                  # a single channel is left alone
    # S8 through the fused feature path
    x = orc.synth_signal(20000, SR, seed=92)
    pcm8 = np.round(x * 127).astype(np.int8)
    got, _ = ctx.stft_features_batch_pcm(pcm8, [0], [pcm8.size], 1024, 256, orc.HANNING, SR)
    assert_features(got, orc.stft_features(orc.pcm_to_f64(pcm8, 8), 1024, 256, orc.HANNING, SR), RTOL64, "pcm8")


def test_time_features_per_frame(ctx, orc):
    """dc_bias / dbfs_max / energy / zero_crossing_rate of every STFT frame (the slice the frame loop cuts),
    rows aligned with the feature matrix; ragged batch, short last frames, exact zeros, silence, many files."""
    lens = [9000, 0, 2048, 2049, 1, 30000, 777]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    a = orc.synth_signal(sum(lens), SR, seed=95)
    a[offs[5] + 100:offs[5] + 140] = 0.0                     # exact zeros: second zero-crossing rule, dbfs of silence
    a[offs[5] + 140] = -0.2
    a[offs[5] + 4096:offs[5] + 4096 + 2048] = 0.0            # a silent frame: dbfs_max = -inf
    for N, hop in ((2048, 512), (1024, 1024), (4096, 1000)):
        got = ctx.time_features_frames(a, offs, lens, N, hop, SR)
        ref = np.concatenate([orc.time_features_frames(a[o:o + n], N, hop, SR) for o, n in zip(offs, lens)])
        assert got.shape == ref.shape
        assert np.array_equal(got[:, 3], ref[:, 3], equal_nan=True)              # counts: exact
        assert np.array_equal(np.isinf(got[:, 1]), np.isinf(ref[:, 1]))
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        fin = np.isfinite(ref)
        assert np.all(np.abs(got[fin] - ref[fin]) <= RTOL64 * np.maximum(np.abs(ref[fin]), 1e-300) + 1e-18)
    many = 3000
    ln = np.full(many, 300, dtype=np.uint64)
    of = np.arange(many, dtype=np.uint64) * np.uint64(300)
    b = orc.synth_signal(many * 300, SR, seed=96)
    got = ctx.time_features_batch(b, of, ln, SR)                                   # binary search over 3000 files
    for i in (0, 1, 1499, 2999):
        seg = b[i * 300:(i + 1) * 300]
        assert got[i, 3] == orc.zero_crossing_rate(seg, SR) and got[i, 2] == pytest.approx(orc.energy(seg), rel=RTOL64)


def test_wrapper_rejects_descriptors_past_the_buffer(ctx, aus, orc):
    """the batch ABI takes offsets / lengths but no total size, so the wrappers check them (a bad descriptor
    would make the H2D copy read past the caller's array), and `out` arrays are validated for every batch call."""
    x = orc.synth_signal(5000, SR, seed=97)
    with pytest.raises(ValueError):
        ctx.stft_features_batch(x, [0, 4000], [4000, 1001], 1024, 256, 2, SR)
    with pytest.raises(ValueError):
        ctx.time_features_batch(x, [10], [4991], SR)
    with pytest.raises(ValueError):
        aus.stft_features_multi([0], x, [1], [5000], 1024, 256, 2, SR)
    pcm = np.zeros(5000, dtype=np.int16)
    with pytest.raises(ValueError):
        ctx.stft_features_batch_pcm(pcm, [0], [5001], 1024, 256, 2, SR)
    with pytest.raises(ValueError):
        ctx.stft_features_batch_pcm(pcm, [0], [5000], 1024, 256, 2, SR, out=np.empty((3, 14), dtype=np.float32))
    with pytest.raises(ValueError):
        ctx.stft_features_batch_pcm(pcm.astype(np.int32), [0], [5000], 1024, 256, 2, SR, sample_bits=20)
    with pytest.raises(aus.AusError):                              # generic wrapper: ceil(m_fl) > floor(m_fu) panics upstream
        aus.analysis.spectral_slope_region(np.ones(1025), orc.rfftfreq(2048, SR), 100.0, 101.0, SR)


def test_host_register_modes_agree(aus, orc):
    """aus_ctx_set_host_register: pageable input page-locked for the call (1), never (0), auto (-1): identical bits."""
    c = aus.Context(0)
    x = orc.synth_signal(300000, SR, seed=98)
    res = []
    for mode in (0, 1, 2, -1):
        c.set_chunk_samples(100000 if mode == 2 else 32 << 20)       # mode 2 with several chunks: both staging buffers cycle
        c.set_host_register(mode)
        res.append(c.stft_features(x, 2048, 512, orc.HANNING, SR))
    assert all(np.array_equal(res[0], r) for r in res[1:])
    with pytest.raises(aus.AusError):
        c.stft_features(x, 2048, 512, orc.HANNING, 8000)            # band above Nyquist: rejected before any copy is queued
    assert np.array_equal(c.stft_features(x, 2048, 512, orc.HANNING, SR), res[0])      # ... and the context is still usable
    c.close()
