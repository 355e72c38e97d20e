"""Known-answer cases restated from the reference's own unit tests.

Each entry: (kind, sigma, ybar, sens, obs, expected score, expected gradient
or None).  Sources: chi/tests/test_error_models.py:23-153, 167-262 (CAM),
411-511, 525-610 (Gaussian), 689-811, 825-921 (LogNormal), 1001-1101,
1115-1202 (Multiplicative).  The expected values are the closed forms the
reference's tests assert (X = X^m cases and unit-sigma cases).
"""
import numpy as np

L2PI = np.log(2 * np.pi)
_S = np.array([[1.0] * 10, [2.0] * 10]).T


def error_model_kats():
    out = []
    # --- ConstantAndMultiplicativeGaussian, X = X^m -------------------------
    for x in (1.0, 10.0):
        st = 1 + 0.1 * x
        out.append(('cam', [1, 0.1], [x] * 10, _S, [x] * 10,
                    -5 * L2PI - 10 * np.log(st),
                    [-10 * 0.1 / st * 1, -10 * 0.1 / st * 2, -10 / st,
                     -10 / st * x]))
    # sigma_tot = 1: score = -log(2pi)/2 - (X^m - X)^2 / 2
    out.append(('cam', [0.9, 0.1], [1] * 10, _S, [2] * 10,
                -5 * L2PI - 5 * 1.0, None))
    out.append(('cam', [0.9, 0.1], [1] * 10, _S, [10] * 10,
                -5 * L2PI - 5 * 81.0, None))
    # --- Gaussian --------------------------------------------------------------
    out.append(('gaussian', [1.0], [1] * 10, _S, [1] * 10, -5 * L2PI,
                [0.0, 0.0, -10.0]))
    out.append(('gaussian', [2.0], [1] * 10, _S, [3] * 10,
                -5 * L2PI - 10 * np.log(2) - 10 * 4 / 8,
                [10 * 2 / 4 * 1, 10 * 2 / 4 * 2, 10 * 4 / 8 - 10 / 2]))
    # --- LogNormal: log y = log ybar - sigma^2/2  -> error term vanishes -----
    s = 0.5
    y = np.exp(np.log(2.0) - s * s / 2)
    out.append(('lognormal', [s], [2.0] * 10, _S, [y] * 10,
                -5 * L2PI - 10 * np.log(s) - 10 * np.log(y),
                [0.0, 0.0, -10 / s]))
    # --- Multiplicative, X = X^m ---------------------------------------------------
    for x in (1.0, 10.0):
        st = 0.1 * x
        out.append(('multiplicative', [0.1], [x] * 10, _S, [x] * 10,
                    -5 * L2PI - 10 * np.log(st),
                    [-10 * 0.1 / st * 1, -10 * 0.1 / st * 2,
                     -10 / st * x]))
    return out
