"""The public building blocks of the path that the fused kernels absorb -- `compute_spectral_*`, `make_power_spectrum`,
`make_spectrum_pmf`, `dot_product` (analysis/spectral_analysis_tools.rs:79-290) and `overlap_add` (spectrum/fft_tools.rs:69) --
as stand-alone calls with the reference's signatures: they must honour whatever intermediates the caller passes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL64 = 1e-9


@pytest.mark.parametrize("K,sr", [(1025, 44100), (513, 48000), (100, 8000), (8193, 44100), (2, 44100), (700, 22050)])
def test_compute_helpers_match_the_oracle(aus, orc, K, sr):
    A = aus.analysis
    rng = np.random.default_rng(K)
    m = np.abs(rng.normal(size=K)) + 1e-3
    f = orc.rfftfreq(2 * (K - 1), sr) if K > 2 else np.array([0.0, sr / 2.0])
    p = A.make_power_spectrum(m)
    assert np.array_equal(p, orc.make_power_spectrum(m))                                    # one multiply per bin: exact
    p_sum, m_sum = float(np.sum(p)), float(np.sum(m))
    pmf = A.make_spectrum_pmf(p, p_sum)
    assert np.array_equal(pmf, orc.make_spectrum_pmf(p, p_sum))                             # one IEEE division per bin: exact
    assert A.dot_product(m, f) == pytest.approx(float(np.dot(m, f)), rel=1e-12)
    c = A.compute_spectral_centroid(m, f, m_sum)
    assert c == pytest.approx(orc.compute_spectral_centroid(m, f, m_sum), rel=RTOL64)
    var = A.compute_spectral_variance(pmf, f, c)
    assert var == pytest.approx(orc.compute_spectral_variance(pmf, f, c), rel=RTOL64)
    assert A.compute_spectral_skewness(pmf, f, c, var) == pytest.approx(orc.compute_spectral_skewness(pmf, f, c, var), rel=1e-7, abs=1e-9)
    assert A.compute_spectral_kurtosis(pmf, f, c, var) == pytest.approx(orc.compute_spectral_kurtosis(pmf, f, c, var), rel=RTOL64)
    assert A.compute_spectral_entropy(pmf) == pytest.approx(orc.compute_spectral_entropy(pmf), rel=RTOL64)
    assert A.compute_spectral_flatness(m, m_sum) == pytest.approx(orc.compute_spectral_flatness(m, m_sum), rel=RTOL64)
    assert A.compute_spectral_slope(p, p_sum) == pytest.approx(orc.compute_spectral_slope(p, p_sum), rel=1e-7, abs=1e-12)
    for n in (0.3, 0.5, 0.85, 0.999, 1.0):
        assert A.compute_spectral_roll_off_point(p, f, p_sum, n) == orc.compute_spectral_roll_off_point(p, f, p_sum, n), n
    if K >= 100:
        fund = sr / (2.0 * (K - 1))
        for fl, fu in ((105.0, 852.0), (0.0, 3.5 * fund), (2 * fund, 40 * fund)):
            got = A.compute_spectral_slope_region(p, f, fl, fu, sr)
            assert got == pytest.approx(orc.compute_spectral_slope_region(p, f, fl, fu, sr), rel=1e-7, abs=1e-9 * float(p.max()))
        with pytest.raises(aus.AusError):
            A.compute_spectral_slope_region(p, f, 0.0, sr / 2.0 + 1.5 * fund, sr)           # the reference indexes out of bounds
    # the caller's intermediates are used as given, not re-derived: a deliberately wrong sum / centroid changes the result
    assert A.compute_spectral_centroid(m, f, 2.0 * m_sum) == pytest.approx(0.5 * c, rel=1e-12)
    assert A.compute_spectral_variance(pmf, f, c + 100.0) == pytest.approx(orc.compute_spectral_variance(pmf, f, c + 100.0), rel=RTOL64)
    assert A.compute_spectral_roll_off_point(p, f, 2.0 * p_sum, 0.4) == orc.compute_spectral_roll_off_point(p, f, 2.0 * p_sum, 0.4)


def test_compute_helpers_edge_cases(aus, orc):
    A = aus.analysis
    K = 64
    flat = np.full(K, 0.37)
    f = np.arange(K, dtype=np.float64) * 3.0
    p = flat * flat
    ps = float(np.sum(p))
    for n in (1.0 / 64, 0.25, 0.5, 0.75, 1.0):                                             # exact ties
        assert A.compute_spectral_roll_off_point(p, f, ps, n) == orc.compute_spectral_roll_off_point(p, f, ps, n), n
    with pytest.raises(aus.AusError):
        A.compute_spectral_roll_off_point(p, f, ps, 0.0)                                    # the reference indexes [-1]
    z = np.zeros(K)
    assert np.isnan(A.compute_spectral_entropy(z)) and np.isnan(orc.compute_spectral_entropy(z))       # 0 * -inf
    assert np.isnan(A.compute_spectral_centroid(z, f, 0.0))                                 # 0 / 0
    assert A.compute_spectral_roll_off_point(z, f, 0.0, 0.5) == orc.compute_spectral_roll_off_point(z, f, 0.0, 0.5) == 0.0
    zz = flat.copy()
    zz[7] = 0.0
    assert A.compute_spectral_flatness(zz, float(zz.sum())) == 0.0 == orc.compute_spectral_flatness(zz, float(zz.sum()))
    with pytest.raises(IndexError):
        A.dot_product(np.ones(4), np.ones(5))


@pytest.mark.parametrize("F,N,hop", [(1, 16, 4), (9, 16, 4), (7, 1000, 250), (40, 2048, 512), (5, 8, 8), (5, 8, 11), (3, 7, 1)])
def test_overlap_add_is_bit_identical(ctx, aus, orc, F, N, hop):
    """Same additions in the same (ascending chunk) order as fft_tools.rs:102-105 -> identical bits; hop > fft_size leaves gaps
    of zeros like the reference's empty lower..=upper range."""
    rng = np.random.default_rng(F * N + hop)
    ch = rng.normal(size=(F, N))
    ref = np.zeros(N + hop * (F - 1))
    for h in range(F):                                                                      # per sample: ascending h
        ref[h * hop:h * hop + N] += ch[h]
    got = aus.spectrum.overlap_add(ch, N, hop)
    assert np.array_equal(got, ref)
    if hop <= N:
        assert np.array_equal(got, orc.overlap_add(ch, N, hop))
    torch = pytest.importorskip("torch")
    ctx.use_torch_stream()
    d_in = torch.from_numpy(ch).cuda()
    d_out = torch.empty(ref.size, dtype=torch.float64, device="cuda")
    ctx.overlap_add_dev(d_in, F, N, hop, d_out)
    ctx.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), ref)
