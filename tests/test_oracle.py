"""Pins the CPU oracle: analytic known-answer tests, an mpmath 50-digit evaluation of the same formulas, the C
restatement against the numpy one, and the committed golden fixtures.  (The reference holds no golden vector for
this path and cannot be built here -- SURVEY 4, 8c -- so these are what 'parity' is anchored on.)"""
import ctypes
import math
import os

import numpy as np
import pytest

SR = 44100
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------- windows (window.rs)
def test_window_known_values(orc):
    L = 17
    han, ham, bla, bar, rec = (orc.generate_window(t, L) for t in (orc.HANNING, orc.HAMMING, orc.BLACKMAN, orc.BARTLETT, orc.RECTANGULAR))
    assert han[0] == 0.0 and abs(han[-1]) < 1e-15 and han[L // 2] == pytest.approx(1.0, abs=1e-15)
    assert ham[0] == pytest.approx(0.08, abs=1e-15) and ham[L // 2] == pytest.approx(1.0, abs=1e-15)
    assert abs(bla[0]) < 1e-16                       # periodic Blackman: 0.42 - 0.5 + 0.08
    assert bla[-1] > 1e-3 and bla[1] == pytest.approx(bla[-1], abs=1e-15)   # periodic (denominator L): only w[0] is 0
    assert bar[0] == 0.0 and bar[L // 2] == pytest.approx(1.0, abs=1e-15) and bar[-1] == pytest.approx(0.0, abs=1e-15)
    assert np.all(rec == 1.0)
    # closed forms, evaluated independently
    i = np.arange(L)
    assert np.allclose(han, 0.5 - 0.5 * np.cos(2 * np.pi * i / (L - 1)), rtol=0, atol=1e-15)
    assert np.allclose(bla, 0.42 - 0.5 * np.cos(2 * np.pi * i / L) + 0.08 * np.cos(4 * np.pi * i / L), rtol=0, atol=1e-15)
    # L = 1: 0/0 inside the cosine (window.rs:64, :81, :29)
    for t in (orc.HANNING, orc.HAMMING, orc.BARTLETT):
        assert np.isnan(orc.generate_window(t, 1)).all()
    assert orc.generate_window(orc.RECTANGULAR, 1)[0] == 1.0


# ---------------------------------------------------------------------------------------------- framing (fft.rs:126-176)
@pytest.mark.parametrize("n,N,hop,F", [(0, 2048, 512, 1), (1, 2048, 512, 1), (2048, 2048, 512, 1), (2049, 2048, 512, 2),
                                       (2560, 2048, 512, 2), (2561, 2048, 512, 3), (2646000, 2048, 512, 5165),
                                       (2646000, 2048, 1024, 2583), (26460000, 4096, 1024, 25837),
                                       (1323000, 1024, 256, 5165), (158760000, 16384, 4096, 38757)])
def test_frame_count_rule(orc, n, N, hop, F):
    assert orc.num_frames(n, N, hop) == F          # SURVEY 8 shape table
    if n <= 30000:
        assert orc.rstft(np.zeros(n), N, hop, orc.RECTANGULAR).shape == (F, N // 2 + 1)


def test_exact_fit_uses_full_window_and_short_frame_uses_short_window(orc):
    N, hop = 256, 64
    x = orc.synth_signal(N + 2 * hop)                # len = h*hop + N exactly -> last frame full, normal window
    fr = orc.stft_frames(x, N, hop, orc.HANNING)
    assert fr.shape[0] == 3
    assert np.array_equal(fr[2], x[2 * hop:] * orc.generate_window(orc.HANNING, N))
    x = orc.synth_signal(N + 2 * hop + 10)           # 10 samples spill into a 4th frame of length 202... = len - 3*hop
    fr = orc.stft_frames(x, N, hop, orc.HANNING)
    L = x.size - 3 * hop
    assert fr.shape[0] == 4 and L < N
    assert np.array_equal(fr[3, :L], x[3 * hop:] * orc.generate_window(orc.HANNING, L))   # fft.rs:149: window of length L
    assert np.all(fr[3, L:] == 0.0)


# ---------------------------------------------------------------------------------------------- transform KATs
def test_bin_centred_cosine(orc):
    N, k0, A = 2048, 37, 0.7
    x = A * np.cos(2 * np.pi * k0 * np.arange(N) / N)
    X = orc.rfft(x, N)
    mag = np.abs(X)
    assert mag[k0] == pytest.approx(A * N / 2, rel=1e-12)
    assert np.max(np.delete(mag, k0)) < 1e-9
    S = orc.rstft(x, N, N, orc.RECTANGULAR)
    assert S.shape == (1, N // 2 + 1) and np.allclose(S[0], X, rtol=0, atol=1e-9)


def test_irfft_is_n_times_signal(orc):
    x = orc.synth_signal(512)
    assert np.allclose(orc.irfft(orc.rfft(x, 512), 512), 512 * x, rtol=0, atol=1e-9)
    with pytest.raises(orc.SpectrumError):
        orc.irfft(np.zeros(100, complex), 512)


def test_dft_against_mpmath(orc):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    N = 128
    x = orc.synth_signal(N, seed=3)
    w = orc.generate_window(orc.HAMMING, N)
    X = orc.rstft(x, N, N, orc.HAMMING)[0]
    xs = [mp.mpf(float(a)) * mp.mpf(float(b)) for a, b in zip(x, w)]
    for k in (0, 1, 17, N // 2):
        ref = sum(xs[n] * mp.exp(-2j * mp.pi * k * n / N) for n in range(N))
        assert abs(complex(ref) - X[k]) <= 1e-13 * max(abs(complex(ref)), 1.0)


# ---------------------------------------------------------------------------------------------- feature KATs
def test_flat_spectrum_features(orc):
    N = 2048
    K = N // 2 + 1
    a = orc.analyzer(np.full(K, 3.0), N, SR)
    f0 = SR / N
    assert a.spectral_flatness == pytest.approx(1.0, rel=1e-13)
    assert a.spectral_entropy == pytest.approx(math.log2(K), rel=1e-13)
    assert abs(a.spectral_slope) < 1e-18 * 9.0 * K
    assert a.spectral_centroid == pytest.approx(f0 * (K - 1) / 2, rel=1e-13)
    # cumulative i/K crosses theta at the first i with (i+1)/K >= theta
    for theta, got in ((0.5, a.spectral_roll_off_50), (0.75, a.spectral_roll_off_75), (0.9, a.spectral_roll_off_90), (0.95, a.spectral_roll_off_95)):
        assert got == f0 * (math.ceil(theta * K) - 1)


def test_silent_and_zero_bin_patterns(orc):
    a = orc.analyzer(np.zeros(1025), 2048, SR).as_array()
    assert np.isnan(a[:6]).all() and np.all(a[6:] == 0.0)        # NaN x 6, roll-offs f[0], slopes 0
    m = np.abs(np.random.default_rng(0).standard_normal(1025)) + 0.1
    m[100] = 0.0
    a = orc.analyzer(m, 2048, SR)
    assert math.isnan(a.spectral_entropy)                          # 0 * log2(0) = NaN poisons the sum (s_a_t.rs:96)
    assert a.spectral_flatness == 0.0                              # exp(-inf) = 0 (s_a_t.rs:110-113)
    assert math.isfinite(a.spectral_centroid)


def test_single_feature_wrappers_equal_analyzer_fields(orc):
    m = np.abs(np.random.default_rng(1).standard_normal(513)) + 0.01
    N = 1024
    f = orc.rfftfreq(N, SR)
    a = orc.analyzer(m, N, SR)
    assert orc.spectral_centroid(m, f) == a.spectral_centroid
    assert orc.spectral_variance(m, f) == a.spectral_variance
    assert orc.spectral_skewness(m, f) == a.spectral_skewness
    assert orc.spectral_kurtosis(m, f) == a.spectral_kurtosis
    assert orc.spectral_entropy(m) == a.spectral_entropy
    assert orc.spectral_flatness(m) == a.spectral_flatness
    assert orc.spectral_roll_off_point(m, f, 0.9) == a.spectral_roll_off_90
    assert orc.spectral_slope(m) == a.spectral_slope
    assert orc.spectral_slope_region(m, f, 1000.0, 5000.0, SR) == a.spectral_slope_1_5_khz


def test_features_against_mpmath(orc):
    """The oracle's 14 features of one frame vs a 50-digit evaluation of Appendix A of SURVEY.md."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    N = 256
    K = N // 2 + 1
    sr = 44100
    x = orc.synth_signal(N, seed=11)
    mag = orc.complex_to_polar_rstft(orc.rstft(x, N, N, orc.HANNING))[0][0]
    got = orc.analyzer(mag, N, sr).as_array()
    m = [mp.mpf(float(v)) for v in mag]
    p = [v * v for v in m]
    M, P = sum(m), sum(p)
    f0 = mp.mpf(sr) / N
    f = [f0 * k for k in range(K)]
    q = [v / P for v in p]
    c = sum(a * b for a, b in zip(m, f)) / M
    V = sum((fk - c) ** 2 * qk for fk, qk in zip(f, q))
    ref = [c, V, sum((fk - c) ** 3 * qk for fk, qk in zip(f, q)) / V ** mp.mpf(1.5),
           sum((fk - c) ** 4 * qk for fk, qk in zip(f, q)) / V ** 2,
           -sum(qk * mp.log(qk, 2) for qk in q), mp.exp(sum(mp.log(v) for v in m) / K) * K / M]
    for theta in (0.5, 0.75, 0.9, 0.95):
        cum, idx = mp.mpf(0), K - 1
        for i in range(K):
            cum += q[i]
            if cum >= theta:
                idx = i
                break
        ref.append(f[idx])
    n = mp.mpf(K)
    Sx, Sxx = n * (n - 1) / 2, n * (n - 1) * (2 * n - 1) / 6
    ref.append((n * sum(pk * k for k, pk in enumerate(p)) - Sx * P) / (n * Sxx - Sx ** 2))
    g = mp.mpf(sr) / ((K - 1) * 2)
    for fl, fu in ((0, 1000), (1000, 5000), (0, 5000)):
        a, b = mp.mpf(fl) / g, mp.mpf(fu) / g
        af, ac, bf, bc = int(mp.floor(a)), int(mp.ceil(a)), int(mp.floor(b)), int(mp.ceil(b))
        yl = p[af] + (a - af) * (p[ac] - p[af])
        yu = p[bf] + (b - bf) * (p[bc] - p[bf])
        sx = fl + sum(f[ac:bf]) + fu
        sy = yl + sum(p[ac:bf]) + yu
        sxx = mp.mpf(fl) ** 2 + sum(v * v for v in f[ac:bf]) + mp.mpf(fu) ** 2
        sxy = fl * yl + sum(a_ * b_ for a_, b_ in zip(p[ac:bf], f[ac:bf])) + fu * yu
        ref.append((n * sxy - sx * sy) / (n * sxx - sx ** 2))
    ref = np.array([float(v) for v in ref])
    assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref)), (got, ref)


# ---------------------------------------------------------------------------------------------- inverse
def test_irstft_quirks(orc):
    N, hop = 512, 256
    x = orc.synth_signal(40 * hop + 77, seed=4)
    y = orc.irstft(orc.rstft(x, N, hop, orc.HANNING), N, hop, orc.HANNING)
    F = orc.num_frames(x.size, N, hop)
    assert y.size == N + hop * (F - 1)                                  # fft.rs:208
    assert math.isnan(y[0]) and math.isnan(y[-1])                       # w[0] = w[N-1] = 0 -> 0/0
    assert np.nanmax(y) == pytest.approx(10 ** (-6 / 20), rel=1e-12)    # -6 dBFS peak (fft.rs:260-265)
    body = slice(N, x.size - 2 * N)
    gain = np.dot(y[body], x[body]) / np.dot(x[body], x[body])
    assert np.max(np.abs(y[body] / gain - x[body])) < 1e-10             # reconstructs up to a data-dependent gain
    with pytest.raises(orc.SpectrumError):
        orc.irstft(np.zeros((3, 100), complex), N, hop, orc.HANNING)


# ---------------------------------------------------------------------------------------------- C restatement
def _clib():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(HERE, "..", "oracle", "libaus_oracle.so"))
    lib.aus_oracle_num_frames.restype = ctypes.c_size_t
    return lib


def _pd(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@pytest.mark.parametrize("N,hop,window,mode,threads", [(2048, 512, 2, 0, 3), (2048, 1024, 3, 0, 8), (1024, 256, 1, 1, 2),
                                                       (4096, 1024, 0, 1, 5), (256, 100, 4, 0, 1)])
def test_c_port_matches_numpy_oracle(orc, N, hop, window, mode, threads):
    lib = _clib()
    x = orc.synth_signal(7 * N + 13, seed=5, file_id=N)
    F = orc.num_frames(x.size, N, hop)
    assert lib.aus_oracle_num_frames(ctypes.c_size_t(x.size), ctypes.c_size_t(N), ctypes.c_size_t(hop)) == F
    out = np.empty((F, 14))
    rc = lib.aus_oracle_stft_features_mt(_pd(x), ctypes.c_size_t(x.size), ctypes.c_size_t(N), ctypes.c_size_t(hop),
                                         ctypes.c_int(window), ctypes.c_uint32(SR), ctypes.c_size_t(threads),
                                         ctypes.c_int(mode), _pd(out))
    assert rc == 0
    ref = orc.stft_features(x, N, hop, window, SR)
    assert np.all(np.abs(out - ref) <= 1e-10 * np.abs(ref))
    spec = np.empty((F, N // 2 + 1), dtype=np.complex128)
    assert lib.aus_oracle_rstft(_pd(x), ctypes.c_size_t(x.size), ctypes.c_size_t(N), ctypes.c_size_t(hop),
                                ctypes.c_int(window), spec.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    S = orc.rstft(x, N, hop, window)
    assert np.max(np.abs(spec - S)) <= 1e-12 * np.max(np.abs(S))
    if window in (3, 4):
        y = np.empty(N + hop * (F - 1))
        assert lib.aus_oracle_irstft(spec.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_size_t(F),
                                     ctypes.c_size_t(N), ctypes.c_size_t(hop), ctypes.c_int(window), _pd(y)) == 0
        assert np.max(np.abs(y - orc.irstft(S, N, hop, window))) <= 1e-10


def test_c_synth_is_bit_identical_noise(orc):
    lib = _clib()
    n = 5000
    x = np.empty(n)
    lib.aus_oracle_synth(_pd(x), ctypes.c_size_t(n), ctypes.c_size_t(123), ctypes.c_uint32(SR), ctypes.c_double(440.0),
                         ctypes.c_uint64(7), ctypes.c_uint64(9))
    ref = orc.synth_signal(n, SR, 440.0, seed=7, file_id=9, n0=123)
    assert np.max(np.abs(x - ref)) <= 2e-16          # sin may differ by an ulp between libm and numpy
    u = orc.counter_uniform(7, 9, 123, n)
    assert u.min() >= -1.0 and u.max() < 1.0 and abs(u.mean()) < 0.05


# ---------------------------------------------------------------------------------------------- golden fixtures
def test_oracle_reproduces_golden_fixtures(orc):
    g = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    tight = dict(rtol=1e-12, atol=0.0, equal_nan=True)
    assert np.allclose(orc.stft_features(g["c1_x"], 2048, 512, orc.HANNING, SR), g["c1_feat"], **tight)
    assert np.allclose(orc.rstft(g["c1_x"], 2048, 512, orc.HANNING)[[0, 7, -1]], g["c1_spec"], rtol=1e-12, atol=1e-12)
    assert np.allclose(orc.stft_analysis(g["mp_x"], 2048, SR, 8), g["mp_feat"], **tight)
    for ch in range(2):
        S = orc.rstft(g[f"c2_x{ch}"], 4096, 1024, orc.HAMMING)
        assert np.allclose(S[0], g[f"c2_spec{ch}_first"], rtol=1e-12, atol=1e-12)
        assert np.allclose(orc.irstft(S, 4096, 1024, orc.HAMMING), g[f"c2_audio{ch}"], rtol=1e-10, atol=1e-13)
    lens = g["c3_lens"]
    pos, feats = 0, []
    for n in lens:
        feats.append(orc.stft_features(g["c3_x"][pos:pos + n], 1024, 256, orc.HANNING, SR))
        pos += n
    assert np.allclose(np.concatenate(feats), g["c3_feat"], **tight)
    assert np.allclose(orc.stft_features(g["c4_x"], 16384, 4096, orc.HANNING, SR), g["c4_feat"], **tight)
    assert np.allclose(orc.analyze_spectrogram(g["an_mag"], 2048, SR), g["an_feat"], **tight)
    for w in range(5):
        assert np.array_equal(orc.generate_window(w, 16), g[f"win{w}_16"]) or np.allclose(orc.generate_window(w, 16), g[f"win{w}_16"], rtol=0, atol=1e-16)


# ---------------------------------------------------------------------------------------------- SURVEY 8(f) rows
def test_autocorrelation_is_reversed_scaled_linear_autocorrelation(orc):
    """s_a_t.rs:31-70 returns samples [N, 2N) of the un-normalised inverse: 2N x the linear autocorrelation at lag N - i."""
    N = 256
    x = orc.synth_signal(N, seed=31)
    ac = orc.autocorrelation(x, N)
    full = np.correlate(x, x, "full")                      # lag l at index N - 1 + l
    i = np.arange(1, N)
    assert np.max(np.abs(ac[i] - 2 * N * full[N - 1 + (N - i)])) <= 1e-12 * np.max(np.abs(ac))
    assert abs(ac[0]) <= 1e-9                              # lag N: no overlap
    with pytest.raises(orc.SpectrumError):
        orc.autocorrelation(x[:-1], N)


def test_time_domain_scalars(orc):
    a = np.array([0.5, -0.25, 0.0, -0.5, 0.0, 0.75, 0.0])
    assert orc.dc_bias(a) == pytest.approx(0.5 / 7)
    assert orc.energy(a) == pytest.approx(np.sqrt(np.mean(a ** 2)))
    assert orc.dbfs_max(a) == pytest.approx(20 * np.log10(0.75))
    assert orc.dbfs_max(np.zeros(4)) == -np.inf and orc.dbfs(1e-21) == -np.inf
    # i=1: 0.5*-0.25<0 (+1); i=2: prod 0, a[3]<0 and a[2]==0 (+1); i=3: prod 0, a[4]<0? no; i=4: a[5]<0? no; i=5, 6: no
    assert orc.zero_crossing_rate(a, 7) == 2.0
    assert orc.energy(np.array([])) == 0.0


# ---------------------------------------------------------------------------------------------- independent library
def test_oracle_against_scipy(orc):
    """A second, unrelated implementation (scipy.signal) of the pieces whose conventions scipy shares: the symmetric
    Hann / Hamming / Bartlett windows, the periodic Blackman, the un-normalised one-sided STFT of full frames
    (scipy's 'spectrum' scaling undone), and the weighted-overlap-add inverse of the interior samples."""
    sig = pytest.importorskip("scipy.signal")
    L = 256
    assert np.allclose(orc.generate_window(orc.HANNING, L), sig.windows.hann(L, sym=True), rtol=0, atol=1e-15)
    assert np.allclose(orc.generate_window(orc.HAMMING, L), sig.windows.hamming(L, sym=True), rtol=0, atol=1e-15)
    assert np.allclose(orc.generate_window(orc.BARTLETT, L), sig.windows.bartlett(L, sym=True), rtol=0, atol=1e-15)
    assert np.allclose(orc.generate_window(orc.BLACKMAN, L), sig.windows.blackman(L, sym=False), rtol=0, atol=1e-15)
    for N, hop, wt in ((512, 128, orc.HANNING), (1024, 300, orc.HAMMING), (256, 256, orc.BLACKMAN)):
        x = orc.synth_signal(9 * N + 3, 44100, seed=5, file_id=N)
        w = orc.generate_window(wt, N)
        S = orc.rstft(x, N, hop, wt)
        _, _, Z = sig.stft(x, window=w, nperseg=N, noverlap=N - hop, nfft=N, boundary=None, padded=False,
                           return_onesided=True, scaling="spectrum")
        Z = Z.T * w.sum()                                        # scipy keeps full frames only
        nfull = Z.shape[0]
        assert nfull >= S.shape[0] - 1 and nfull <= S.shape[0]
        peak = np.abs(S[:nfull]).max(axis=1, keepdims=True)
        assert np.max(np.abs(S[:nfull] - Z) / peak) < 1e-13, (N, hop, wt)
    # inverse: interior samples of irstft (before peak normalisation) equal the signal when sum w^2 is the divisor
    N, hop = 512, 128
    x = orc.synth_signal(20 * N, 44100, seed=6)
    y = orc.irstft(orc.rstft(x, N, hop, orc.HAMMING), N, hop, orc.HAMMING, normalize=False)
    w = orc.generate_window(orc.HAMMING, N)
    _, _, Z = sig.stft(x, window=w, nperseg=N, noverlap=N - hop, boundary=None, padded=False, scaling="spectrum")
    _, ys = sig.istft(Z, window=w, nperseg=N, noverlap=N - hop, boundary=False, scaling="spectrum")
    body = slice(N, min(ys.size, y.size, x.size) - N)
    # the reference's irfft is un-normalised (N x the signal, fft.rs:66-93) and irstft keeps that factor
    assert np.max(np.abs(y[body] / N - x[body])) < 1e-12
    assert np.max(np.abs(ys[body] - x[body])) < 1e-12


def test_non_power_of_two_sizes_against_the_dft_definition(orc):
    """rustfft plans any length (fft.rs:120-121); the oracle's rstft / irstft for such sizes are pinned against the DFT sum
    itself (mpmath, 40 digits) and against the overlap-add identity, so the chirp-z path on the GPU is checked against
    something that does not share its algorithm."""
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(17)
    for N in (12, 45, 97):                                   # composite, odd composite, prime
        x = rng.normal(size=N)
        X = orc.rfft(x, N)
        assert X.shape == (N // 2 + 1,)
        for k in range(N // 2 + 1):
            ref = sum(mp.mpf(float(x[n])) * mp.e ** (-2j * mp.pi * k * n / N) for n in range(N))
            assert abs(complex(X[k]) - complex(ref)) < 1e-12 * max(1.0, abs(complex(ref)))
    # rstft -> irstft at a non-power-of-two size reproduces N x the signal away from the ends (rustfft's inverse is unnormalised)
    N, hop = 1000, 250
    x = orc.synth_signal(12 * N + 5, 44100, seed=3)
    S = orc.rstft(x, N, hop, orc.HAMMING)
    assert S.shape == (orc.num_frames(x.size, N, hop), N // 2 + 1)
    y = orc.irstft(S, N, hop, orc.HAMMING, normalize=False)
    body = slice(N, x.size - N)
    assert np.max(np.abs(y[body] / N - x[body])) < 1e-12
