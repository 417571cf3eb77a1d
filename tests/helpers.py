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


def assert_features(got, ref, rtol, what=""):
    r = rel_err(got, ref)
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
