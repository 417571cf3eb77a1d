"""Minimax (Lawson-weighted least squares) fit of the degree-2 polynomial in log2(1 + r) = r * poly(r), |r| <= 2^-8,
used by log2_d3() in aus-rust_b200/csrc/aus_stft_pc.cuh; prints the coefficients and the true error (mpmath)."""
import numpy as np
from mpmath import mp, mpf, log
mp.dps = 40
R = 2.0**-8
# Remez for g(r) = log2(1+r)/r approximated by a + b r + c r^2 minimising max |r*(poly - g)|  (error of log2 value)
def g(r):
    r = mpf(r)
    return (log(1+r)/log(2))/r if r != 0 else 1/log(2)
# weighted least squares on dense Chebyshev grid then a few Remez-like reweighting iterations (Lawson)
xs = np.cos(np.linspace(0, np.pi, 2001)) * R
xs = xs[xs != 0]
G = np.array([float(g(x)) for x in xs])
W = np.ones_like(xs)
A = np.vstack([np.ones_like(xs), xs, xs**2]).T
for it in range(200):
    sw = np.sqrt(W) * np.abs(xs)      # error of interest = r * (poly - g)
    coef, *_ = np.linalg.lstsq(A * sw[:, None], G * sw, rcond=None)
    err = np.abs(xs * (A @ coef - G))
    W = W * (err / err.max() + 1e-3)
    W /= W.sum()
print("coef (r^0, r^1, r^2):", [repr(c) for c in coef])
# evaluate true error in high precision
xs2 = np.linspace(-R, R, 20001)
errs = []
for x in xs2:
    if x == 0: continue
    val = mpf(coef[0]) + mpf(coef[1])*mpf(x) + mpf(coef[2])*mpf(x)**2
    errs.append(float(mpf(x)*val - log(1+mpf(x))/log(2)))
errs = np.array(errs)
print("max abs err", np.abs(errs).max(), "mean err", errs.mean())
t = [1/np.log(2), -1/(2*np.log(2)), 1/(3*np.log(2))]
errs_t = np.array([float(mpf(x)*(mpf(t[0]) + mpf(t[1])*mpf(x) + mpf(t[2])*mpf(x)**2) - log(1+mpf(x))/log(2)) for x in xs2 if x != 0])
print("taylor: max", np.abs(errs_t).max(), "mean", errs_t.mean())
