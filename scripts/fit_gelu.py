"""Fit of the erf-GELU approximation used by the GEGLU epilogue (rcdms_b200/csrc/common.cuh: gelu_erf_f).
erfc(t) = 2^-p(t), p = c1 t + ... + c5 t^5 (weighted least squares of -log2 erfc on [0, 6], p(0) = 0).
Prints the coefficients in t and, with the 1/sqrt(2) of gelu folded in, in |x|, plus the max abs errors."""
import math

import numpy as np
from scipy.special import erf, erfc

x = np.linspace(0, 6.0, 60001)
y = -np.log2(np.maximum(erfc(x), 1e-300))
w = np.sqrt(erfc(x)) + 1e-6
V = np.vander(x, 6, increasing=True)[:, 1:]
coef, *_ = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)
print("coefficients in t        :", coef)
print("coefficients in |x|      :", [coef[k] / math.sqrt(2) ** (k + 1) for k in range(5)])
print("erf  max abs error       : %.2e" % np.abs(1 - np.exp2(-(V @ coef)) - erf(x)).max())
g = np.linspace(-12, 12, 240001)
ax = np.abs(g) / math.sqrt(2)
e = np.copysign(1 - np.exp2(-(np.vander(ax, 6, increasing=True)[:, 1:] @ coef)), g)
print("gelu max abs error       : %.2e" % np.abs(0.5 * g * (1 + e) - 0.5 * g * (1 + erf(g / math.sqrt(2)))).max())
