"""Coefficients of the linear-model Rosenbrock methods ROSL(s) of
chi_b200/csrc/integrator_stream.cuh.

For a LINEAR system z' = J z + g (g constant over a step) every Rosenbrock
method with a single gamma reduces to

    z_1 = R(hJ) z_0 + h Q(hJ) g,     R(x) = 1 + x Q(x) = P(x) / (1 - gamma x)^s,

and only the linear order conditions R(x) = exp(x) + O(x^(p+1)) remain.  ROSL(s)
takes P of degree s-1 from the series of exp(x) (1 - gamma x)^s: the restricted
Pade approximant (Norsett) of order p = s-1 with R(inf) = 0 (L-stable when
A-stable).  With tau = 1/(1 - gamma x), i.e. T = (I - gamma h J)^-1, and
E = T - I = gamma h J T (what the kernel calls E = W^-1 J):

    Q(x) = sum_{k=0}^{s-1} c_k tau (tau - 1)^k
    z_1  = z_0 + sum_k c_k w_k,    w_0 = T (h f(z_0)),   w_k = E w_{k-1}.

(tau - 1)^k = O(x^k): the terms decay, the coefficients are O(1), and dropping
the last two terms leaves a method of order s-2, so

    err = c_{s-2} w_{s-2} + c_{s-1} w_{s-1}

is the embedded error estimate -- for free.  This script computes the c_k in
60-digit arithmetic, checks order and A-stability, and prints the tables.
"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60


def coefficients(s, gamma):
    g = mp.mpf(gamma)
    den = [mp.binomial(s, k) * (-g) ** k for k in range(s + 1)]
    ex = [1 / mp.factorial(k) for k in range(s)]
    P = [sum(ex[i] * den[k - i] for i in range(k + 1)) for k in range(s)]
    num = [(P[k] if k < s else 0) - den[k] for k in range(s + 1)]
    assert abs(num[0]) < mp.mpf(10) ** -50
    N = num[1:]
    B = mp.zeros(s, s)
    for k in range(s):
        for j in range(s - k):
            B[k + j, k] = g ** k * mp.binomial(s - 1 - k, j) * (-g) ** j
    c = mp.lu_solve(B, mp.matrix(N))
    return [c[k] for k in range(s)], P, den


def R(c, gamma, x):
    tau = 1 / (1 - gamma * x)
    return 1 + x * sum(ck * tau * (tau - 1) ** k for k, ck in enumerate(c))


def check(s, gamma):
    c, P, den = coefficients(s, gamma)
    # order: R(x) - exp(x) = C x^s + ...
    x = mp.mpf(1) / 64
    e1 = R(c, gamma, x) - mp.e ** x
    e2 = R(c, gamma, x / 2) - mp.e ** (x / 2)
    order = mp.log(abs(e1 / e2), 2) - 1
    # embedded: drop the last two terms
    e1 = R(c[:-2], gamma, x) - mp.e ** x
    e2 = R(c[:-2], gamma, x / 2) - mp.e ** (x / 2)
    order_hat = mp.log(abs(e1 / e2), 2) - 1
    y = np.concatenate([np.linspace(0, 80, 80001), np.logspace(1.9, 8, 8000)])
    cf = [float(v) for v in c]
    gamma = float(gamma)
    tau = 1 / (1 - gamma * 1j * y)
    Riy = 1 + 1j * y * sum(ck * tau * (tau - 1) ** k
                           for k, ck in enumerate(cf))
    xr = -np.concatenate([np.linspace(0, 300, 60001), np.logspace(2.5, 9, 4000)])
    tau = 1 / (1 - gamma * xr)
    Rx = 1 + xr * sum(ck * tau * (tau - 1) ** k for k, ck in enumerate(cf))
    return c, float(order), float(order_hat), float(np.abs(Riy).max()), \
        float(np.abs(Rx).max()), float(abs(Rx[-1]))


if __name__ == '__main__':
    methods = [(9, '0.2'), (11, '0.16')]
    if len(sys.argv) > 2:
        methods = [(int(sys.argv[1]), sys.argv[2])]
    for s, gamma in methods:
        c, order, order_hat, amax, rmax, rinf = check(s, mp.mpf(gamma))
        print('// ROSL%d: gamma = %s, order %.2f (embedded %.2f), '
              'max |R(iy)| = %.12f, max |R(x<0)| = %.6f, |R(-1e9)| = %.1e'
              % (s, gamma, order, order_hat, amax, rmax, rinf))
        for k, ck in enumerate(c):
            print('  %s,' % mp.nstr(ck, 20))
