import numpy as np

ROLL = slice(6, 10)


def rel_err(got, ref):
    """Element-wise relative error with NaN == NaN and inf == inf counted as exact."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    both_nan = np.isnan(got) & np.isnan(ref)
    same = (got == ref) | both_nan
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(got - ref) / np.abs(ref)
    r = np.where(same, 0.0, r)
    r = np.where(np.isnan(r), np.inf, r)       # NaN on one side only
    return r


def cancellation_scales(mag, fft_size, sample_rate):
    """Per-feature magnitude of the largest terms that cancel inside the reference's formula, [F, 14].  A feature
    that is a difference of large terms (slope of a flat spectrum, skewness of a symmetric one) carries an
    absolute error of (machine eps x those terms) in ANY implementation, the reference included, so parity is
    |got - ref| <= rtol * max(|ref|, scale).  Zero for features that are sums of same-signed terms."""
    mag = np.atleast_2d(np.asarray(mag, dtype=np.float64))
    F, K = mag.shape
    n = float(K)
    f0 = sample_rate / fft_size
    f = f0 * np.arange(K)
    p = mag ** 2
    P = p.sum(axis=1)
    sc = np.zeros((F, 14))
    with np.errstate(all="ignore"):
        c = (mag * f).sum(axis=1) / mag.sum(axis=1)
        q = p / P[:, None]
        var = ((f - c[:, None]) ** 2 * q).sum(axis=1)
        sc[:, 2] = (np.abs(f - c[:, None]) ** 3 * q).sum(axis=1) / var ** 1.5
        sx = n * (n - 1) / 2
        sxx = n * (n - 1) * (2 * n - 1) / 6
        sc[:, 10] = (n * (p * np.arange(K)).sum(axis=1) + sx * P) / (n * sxx - sx * sx)
        g = sample_rate / ((K - 1) * 2.0)
        for b, (fl, fu) in enumerate(((0.0, 1000.0), (1000.0, 5000.0), (0.0, 5000.0))):
            a, e = int(np.ceil(fl / g)), int(np.floor(fu / g))
            ee = min(int(np.ceil(fu / g)), K - 1)
            Sx = fl + f[a:e].sum() + fu
            Sxx = fl * fl + (f[a:e] ** 2).sum() + fu * fu
            Sy = p[:, a:ee + 1].sum(axis=1)
            Sxy = (p[:, a:ee + 1] * f[a:ee + 1]).sum(axis=1) + fu * p[:, ee]
            sc[:, 11 + b] = (n * Sxy + Sx * Sy) / abs(n * Sxx - Sx * Sx)
    return np.nan_to_num(sc, nan=0.0, posinf=0.0)


def assert_features(got, ref, rtol, what="", scales=None):
    r = rel_err(got, ref)
    if scales is not None:
        with np.errstate(invalid="ignore", divide="ignore"):
            r2 = np.abs(np.asarray(got) - np.asarray(ref)) / np.maximum(np.abs(ref), scales)
        r = np.where(np.isfinite(r2), np.minimum(r, r2), r)
    bad = np.argwhere(r > rtol)
    assert bad.size == 0, (f"{what}: {len(bad)} feature values beyond rtol={rtol:g}; worst {r.max():.3e} at "
                           f"{tuple(bad[np.argmax(r[tuple(bad.T)])])}; got {got[tuple(bad[0])]!r} ref {ref[tuple(bad[0])]!r}")
    return float(r.max()) if r.size else 0.0


def assert_spectra(got, ref, rtol, what=""):
    """|delta| <= rtol * max|X| per frame (bins far below the frame peak carry no relative information)."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    peak = np.max(np.abs(ref), axis=-1, keepdims=True)
    peak = np.where(peak == 0, 1.0, peak)
    err = np.abs(got - ref) / peak
    err = np.where(np.isnan(got) & np.isnan(ref), 0.0, err)
    assert np.all(err <= rtol), f"{what}: max scaled spectrum error {np.nanmax(err):.3e} > {rtol:g}"
    return float(np.nanmax(err)) if err.size else 0.0
