"""Fit of the division-free GELU used by the CUDA epilogues (sky_common.cuh::gelu_erf):
gelu(x) = relu(x) - |x| h(|x|), h(a) = exp(-a^2/2) P7(a), P7 ~ 0.5 erfcx(a / sqrt 2) on [0, 6]
(least squares under the weight exp(-a^2/2), i.e. minimising the error of h itself)."""
import numpy as np
from numpy.polynomial import chebyshev as Ch, polynomial as P
from scipy.special import erf, erfcx

A, deg = 6.0, 7
a = np.linspace(0, A, 40001)
z = a / np.sqrt(2)
g, wgt = 0.5 * erfcx(z), np.exp(-z * z)
T = Ch.chebvander(2 * a / A - 1, deg)
c = np.linalg.lstsq(T * wgt[:, None], g * wgt, rcond=None)[0]
coef = Ch.Chebyshev(c, domain=[0, A]).convert(kind=P.Polynomial, domain=[-1, 1], window=[-1, 1]).coef
print("coefficients a^0..a^7:", ", ".join("%.9ef" % v for v in coef))
x = np.linspace(-8, 8, 200001).astype(np.float32)
ax = np.abs(x)
p = np.float32(coef[7]) * np.ones_like(x)
for k in range(6, -1, -1):
    p = p * ax + np.float32(coef[k])
e = np.exp2((x * x * np.float32(-0.5 * 1.4426950408889634)).astype(np.float32))
gel = np.maximum(x, 0) - ax * (p * e)
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("max abs error vs exact erf-GELU: %.2e" % np.abs(gel - ref).max())
