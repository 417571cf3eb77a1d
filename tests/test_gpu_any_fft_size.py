"""fft sizes outside the radix-32 kernels' range -- anything that is not a power of two in [64, 16384].  rustfft plans any
length (spectrum/fft.rs:120-121, :205-206), so the reference accepts them; here they run as a chirp-z transform on top of
the power-of-two kernels (aus-rust_b200/csrc/aus_anyn.cuh).  Same oracle, same tolerances as the native sizes."""
import numpy as np
import pytest

from helpers import assert_features, assert_spectra

pytestmark = pytest.mark.gpu

RTOL64 = 1e-9
SR = 44100


@pytest.mark.parametrize("fft_size,hop", [(1000, 250), (1001, 333), (1536, 384), (3000, 1000), (8191, 4096), (6000, 1500),
                                          (48, 12), (32, 8), (6, 3), (3, 1), (2, 1), (100, 100)])
def test_rstft_any_size(ctx, orc, fft_size, hop):
    """Composite, prime, odd, tiny and small power-of-two sizes; ragged length (short last frame with its own window)."""
    x = orc.synth_signal(7 * fft_size + 11, SR, seed=21, file_id=fft_size)
    for window in (orc.HANNING, orc.HAMMING, orc.BLACKMAN):
        ref = orc.rstft(x, fft_size, hop, window)
        got = ctx.rstft(x, fft_size, hop, window)
        assert got.shape == ref.shape == (orc.num_frames(x.size, fft_size, hop), fft_size // 2 + 1)
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert_spectra(np.nan_to_num(got), np.nan_to_num(ref), RTOL64, f"N={fft_size} w={window}")


@pytest.mark.parametrize("n", [0, 1, 2, 5, 999, 1000, 1001, 1250, 1251])
def test_any_size_ragged_lengths(ctx, orc, n):
    """Frame-count rule (strict '<', fft.rs:131) and the short last window (fft.rs:148-159) on the chirp-z path."""
    x = orc.synth_signal(n, SR, seed=5, file_id=n)
    ref = orc.rstft(x, 1000, 250, orc.HANNING)
    got = ctx.rstft(x, 1000, 250, orc.HANNING)
    assert got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert_spectra(np.nan_to_num(got), np.nan_to_num(ref), RTOL64, f"n={n}")


@pytest.mark.parametrize("fft_size,hop,sr", [(1000, 250, 44100), (1001, 500, 44100), (1920, 480, 48000), (3000, 750, 44100),
                                             (441, 147, 44100), (32, 8, 44100)])
def test_features_any_size(ctx, orc, fft_size, hop, sr):
    x = orc.synth_signal(40 * fft_size + 37, sr, seed=22, file_id=fft_size)
    ref = orc.stft_features(x, fft_size, hop, orc.HANNING, sr)
    got = ctx.stft_features(x, fft_size, hop, orc.HANNING, sr, "f64")
    assert_features(got, ref, RTOL64, f"N={fft_size}")
    got32 = ctx.stft_features(x, fft_size, hop, orc.HANNING, sr, "f32")      # f64 arithmetic inside, f32 at the boundary
    assert_features(got32, ref, 1e-4, f"N={fft_size} f32")


def test_stft_analysis_any_size(ctx, orc):
    """mp::stft_analysis(audio, 1000, ..): hop = 500, Hamming (mp.rs:42)."""
    x = orc.synth_signal(2 * SR, SR, seed=23)
    ref = orc.stft_analysis(x, 1000, SR, 8)
    got = ctx.stft_analysis(x, 1000, SR, 8)
    assert_features(got, ref, RTOL64, "stft_analysis N=1000")


def test_batch_and_analyzer_any_size(ctx, aus, orc):
    """Several ragged files in one call, and analyzer() on caller-supplied magnitude rows of K = 501 bins."""
    lens = [5000, 0, 1000, 999, 12345, 1]
    xs = [orc.synth_signal(n, SR, seed=30 + i, file_id=i) for i, n in enumerate(lens)]
    offs = np.cumsum([0] + lens[:-1])
    feats, starts = ctx.stft_features_batch(np.concatenate(xs), offs, lens, 1000, 250, orc.HAMMING, SR)
    assert starts[-1] == sum(orc.num_frames(n, 1000, 250) for n in lens)
    for i, x in enumerate(xs):
        if x.size < 4:
            continue          # 0..3-sample files: degenerate windows, covered by the ragged test on spectra
        ref = orc.stft_features(x, 1000, 250, orc.HAMMING, SR)
        assert_features(feats[starts[i]:starts[i + 1]], ref, RTOL64, f"file {i}")
    rng = np.random.default_rng(5)
    mags = np.abs(rng.normal(size=(37, 501))) + 1e-3
    mags[3, :] = 0.0
    mags[4, 17] = 0.0
    ref = orc.analyze_spectrogram(mags, 1000, SR)
    got = ctx.analyzer_batch(mags, 1000, SR)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=RTOL64, atol=1e-300)


@pytest.mark.parametrize("fft_size,hop,window", [(1000, 250, 3), (1000, 500, 3), (1536, 384, 3), (48, 12, 3), (1000, 250, 2),
                                                 (600, 200, 0)])
def test_irstft_any_size(ctx, orc, fft_size, hop, window):
    x = orc.synth_signal(20 * fft_size + 123, SR, seed=8, file_id=fft_size)
    S = orc.rstft(x, fft_size, hop, window)
    ref = orc.irstft(S, fft_size, hop, window, normalize=False)
    got = ctx.irstft(S, fft_size, hop, window, normalize=False)
    assert got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))          # Hanning / Bartlett: 0/0 at both ends
    # samples whose window norm is tiny amplify ANY implementation's rounding by 1/w (see test_irstft_blackman_unnormalised)
    w2 = orc.generate_window(window, fft_size) ** 2
    norm = np.zeros(ref.size)
    for h in range(S.shape[0]):
        norm[h * hop:h * hop + fft_size] += w2
    ok = norm > 1e-6
    assert ok.sum() >= ref.size - 64
    assert np.max(np.abs(got[ok] - ref[ok])) <= RTOL64 * np.max(np.abs(ref[ok]))
    if window == 3:                                              # Hamming never vanishes: the normalised signal too
        refn = orc.irstft(S, fft_size, hop, window)
        gotn = ctx.irstft(S, fft_size, hop, window)
        assert np.max(np.abs(gotn - refn)) <= RTOL64 * np.max(np.abs(refn))
        body = slice(fft_size, x.size - fft_size)
        assert np.max(np.abs(got[body] / fft_size - x[body])) < 1e-9 * np.max(np.abs(x))   # round trip (rustfft's inverse is unnormalised)


def test_rfft_irfft_and_errors_any_size(ctx, aus, orc):
    rng = np.random.default_rng(9)
    for N in (1000, 30, 2, 4410):
        x = rng.normal(size=N)
        X = ctx.rfft(x, N)
        np.testing.assert_allclose(X, np.fft.rfft(x), rtol=0, atol=1e-11 * np.abs(np.fft.rfft(x)).max())
        y = ctx.irfft(X, N)
        np.testing.assert_allclose(y, x * N, rtol=0, atol=1e-10 * N)            # rustfft's inverse is unnormalised (fft.rs:86)
    x = rng.normal(size=1001)
    X = ctx.rfft(x, 1001)                                                       # odd size: (N + 1) / 2 bins, forward only
    assert X.shape == (501,)
    np.testing.assert_allclose(X, np.fft.fft(x)[:501], rtol=0, atol=1e-11 * np.abs(np.fft.fft(x)).max())
    with pytest.raises(aus.AusError):
        ctx.irstft(np.zeros((3, 501), dtype=np.complex128), 1001, 250, 3)        # odd size: the reference panics in rustfft
    with pytest.raises(aus.AusError):
        ctx.rstft(rng.normal(size=50000), 10000, 2500, 2)                        # beyond the chirp-z range
    with pytest.raises(aus.AusError):
        ctx.rstft(rng.normal(size=5000), 1000, 0, 2)                             # hop = 0: the reference never terminates
    with pytest.raises(aus.SpectrumError):
        ctx.irstft(np.zeros((3, 500), dtype=np.complex128), 1000, 250, 3)        # wrong row length (fft.rs:217-219)
    # autocorrelation of a 500-sample chunk runs on the 1000-point transform (s_a_t.rs:31-65)
    a = rng.normal(size=500)
    np.testing.assert_allclose(ctx.autocorrelation(a, 500), orc.autocorrelation(a, 500), rtol=0, atol=1e-9 * np.abs(a @ a))


def test_any_size_many_frames_and_device_pointers(ctx, aus, orc):
    """More frames than one workspace pass holds (the path is chunked), through the device-pointer entry points."""
    torch = pytest.importorskip("torch")
    N, hop = 1000, 100
    x = orc.synth_signal(40 * SR, SR, seed=41)                  # 17 631 frames; one pass holds 8 190
    F = orc.num_frames(x.size, N, hop)
    dx = torch.from_numpy(x).cuda()
    dfeat = torch.empty((F, 14), dtype=torch.float64, device="cuda")
    dspec = torch.empty((F, N // 2 + 1, 2), dtype=torch.float64, device="cuda")
    ctx.use_torch_stream()
    ctx.stft_features_batch_dev(dx, [0], [x.size], N, hop, orc.HANNING, SR, "f64", dfeat, dspec)
    ctx.synchronize()
    sel = np.r_[0:40, 8170:8230, F - 40:F]                       # both ends and the seam between two passes
    ref_s = orc.rstft(x, N, hop, orc.HANNING)[sel]
    got_s = dspec.cpu().numpy()[sel]
    assert_spectra(got_s[..., 0] + 1j * got_s[..., 1], ref_s, RTOL64, "device spectra")
    mag = np.abs(ref_s)
    ref_f = orc.analyze_spectrogram(mag, N, SR)
    assert_features(dfeat.cpu().numpy()[sel], ref_f, RTOL64, "device features")


def test_any_size_multi_device_entry_points(aus, orc):
    """Frame-range shards with recomputed halo (features and irstft) work the same for a chirp-z size: four shards over the
    visible devices (all on device 0 when there is one)."""
    torch = pytest.importorskip("torch")
    devs = [i % torch.cuda.device_count() for i in range(4)]
    N, hop = 1000, 250
    x = orc.synth_signal(60 * N + 77, SR, seed=51)
    try:
        got = aus.stft_features_multi(devs, x, [0], [x.size], N, hop, orc.HAMMING, SR)
        assert_features(got, orc.stft_features(x, N, hop, orc.HAMMING, SR), RTOL64, "multi features N=1000")
        S = orc.rstft(x, N, hop, orc.HAMMING)
        outs = aus.irstft_multi(devs, [S, S[:7]], N, hop, orc.HAMMING)
        for o, s in zip(outs, (S, S[:7])):
            ref = orc.irstft(s, N, hop, orc.HAMMING)
            assert o.shape == ref.shape and np.max(np.abs(o - ref)) <= RTOL64 * np.max(np.abs(ref))
    finally:
        aus.multi_shutdown()


def test_any_size_chunked_host_pipeline_pcm_and_f32_inverse(ctx, aus, orc):
    """The chunked copy/compute pipeline (virtual files with a recomputed halo), the PCM entry and the f32 inverse all reach
    the chirp-z path through the same funnels as the native sizes."""
    N, hop = 1000, 250
    x = orc.synth_signal(70000, SR, seed=61)
    ref = orc.stft_features(x, N, hop, orc.HANNING, SR)
    try:
        ctx.set_chunk_samples(9000)                               # eight chunks
        got = ctx.stft_features(x, N, hop, orc.HANNING, SR)
    finally:
        ctx.set_chunk_samples(32 << 20)
    assert_features(got, ref, RTOL64, "chunked N=1000")
    pcm = np.round(x * 20000).astype(np.int16)
    xf = orc.pcm_to_f64(pcm, 16)
    got_p, _ = ctx.stft_features_batch_pcm(pcm, [0], [pcm.size], N, hop, orc.HANNING, SR)
    assert_features(got_p, orc.stft_features(xf, N, hop, orc.HANNING, SR), RTOL64, "pcm N=1000")
    S = orc.rstft(x, N, hop, orc.HAMMING)
    ref_i = orc.irstft(S, N, hop, orc.HAMMING)
    got_i = ctx.irstft(S, N, hop, orc.HAMMING, precision="f32")
    assert np.max(np.abs(got_i - ref_i)) <= 1e-4 * np.max(np.abs(ref_i))


def test_any_size_randomised_sweep(ctx, orc):
    """Seeded sweep over arbitrary sizes, hops, windows and lengths: spectra against the oracle, frame count included."""
    rng = np.random.default_rng(2026)
    for case in range(40):
        N = int(rng.integers(2, 3001))
        if N >= 64 and (N & (N - 1)) == 0:
            N += 1                                               # keep it off the native path
        hop = int(rng.integers(1, N + 1))
        window = int(rng.integers(0, 5))
        n = int(rng.integers(0, 12 * N))
        x = orc.synth_signal(n, SR, seed=100 + case, file_id=case)
        ref = orc.rstft(x, N, hop, window)
        if ref.shape[0] > 4000:                                  # hop = 1 on a long signal: keep the case small
            x = x[:N + 300 * hop]
            ref = orc.rstft(x, N, hop, window)
        got = ctx.rstft(x, N, hop, window)
        assert got.shape == ref.shape, (N, hop, window, n)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (N, hop, window, n)
        assert_spectra(np.nan_to_num(got), np.nan_to_num(ref), RTOL64, f"N={N} hop={hop} w={window} n={n}")


def test_any_size_device_entry_points_stay_inside_their_buffers(ctx, aus, orc):
    """The memcheck stand-in (compute-sanitizer is closed on the pool) for the chirp-z path and overlap_add: every device
    output sits between guard bands that must keep their fill value."""
    torch = pytest.importorskip("torch")
    ctx.use_torch_stream()
    N, hop, G = 1000, 250, 4096
    x = orc.synth_signal(9 * N + 77, SR, seed=71)
    F, K = orc.num_frames(x.size, N, hop), N // 2 + 1

    def guarded(n):
        a = torch.full((n + 2 * G,), -7.25, dtype=torch.float64, device="cuda")
        return a, a[G:G + n]

    def intact(a, n):
        return bool((a[:G] == -7.25).all()) and bool((a[G + n:] == -7.25).all())

    dx = torch.from_numpy(x).cuda()
    fa, feat = guarded(F * 14)
    sa, spec = guarded(F * K * 2)
    ctx.stft_features_batch_dev(dx, [0], [x.size], N, hop, orc.HAMMING, SR, "f64", feat, spec)
    n_out = aus.irstft_len(F, N, hop)
    oa, out = guarded(n_out)
    ctx.irstft_dev(spec, F, N, hop, orc.HAMMING, "f64", out)
    ch = torch.from_numpy(np.random.default_rng(3).normal(size=(F, N))).cuda()
    la, ola = guarded(N + hop * (F - 1))
    ctx.overlap_add_dev(ch, F, N, hop, ola)
    ctx.synchronize()
    assert intact(fa, F * 14) and intact(sa, F * K * 2) and intact(oa, n_out) and intact(la, N + hop * (F - 1))
    ref = orc.stft_features(x, N, hop, orc.HAMMING, SR)
    assert_features(feat.cpu().numpy().reshape(F, 14), ref, RTOL64, "guarded features")
    S = orc.rstft(x, N, hop, orc.HAMMING)
    refo = orc.irstft(S, N, hop, orc.HAMMING)
    assert np.max(np.abs(out.cpu().numpy() - refo)) <= RTOL64 * np.max(np.abs(refo))
