"""Derives the beta = alpha + gamma_offdiag tables of the Rosenbrock methods in
their original (k-stage) form from the transformed (Hairer-Wanner) tables the
solver carries (a_ij, C_ij, gamma):

    U = Gamma k,   a = alpha Gamma^-1,   C = diag(1/gamma) - Gamma^-1

For a LINEAR system  y' = J y + g  the stage equations collapse to
    (I - gamma h J) k_i = h J (y + sum_{j<i} beta_ij k_j) + h g,
    beta = alpha + (Gamma - gamma I),
i.e. ONE combination per stage; stiffly accurate methods then have
    y_new = arg_s + gamma k_s,   y_hat = arg_{s-1} + gamma k_{s-1}.

usage: python tools/derive_kform.py    (prints C++ constants)
"""
import re
import sys
from fractions import Fraction
from decimal import Decimal

import mpmath as mp

mp.mp.dps = 50


def read_namespace(path, name):
    text = open(path).read()
    body = text[text.index('namespace %s {' % name):]
    body = body[:body.index('}  // namespace %s' % name)]
    vals = {}
    for m in re.finditer(r'constexpr double (\w+) = ([-+0-9.eE]+);', body):
        vals[m.group(1)] = mp.mpf(m.group(2))
    return vals


def derive(vals, s):
    g = vals['GAMMA']
    a = mp.zeros(s, s)
    C = mp.zeros(s, s)
    for i in range(1, s):
        have = ('A%d1' % (i + 1)) in vals
        for j in range(i):
            key = '%d%d' % (i + 1, j + 1)
            C[i, j] = vals['C' + key]
            if have:
                a[i, j] = vals['A' + key]
            else:
                # incremental row: a_i = a_{i-1} + e_{i-1}
                a[i, j] = a[i - 1, j] if j < i - 1 else mp.mpf(1)
    Ginv = mp.diag([1 / g] * s) - C
    G = Ginv ** -1
    alpha = a * G
    beta = alpha + G - g * mp.eye(s)
    return g, alpha, G, beta


def main():
    root = __file__.rsplit('/tools/', 1)[0]
    for name, path, s in (
            ('rodas4', root + '/chi_b200/csrc/integrator.cuh', 6),
            ('rodas5', root + '/chi_b200/csrc/integrator_stream.cuh', 8)):
        vals = read_namespace(path, name)
        g, alpha, G, beta = derive(vals, s)
        # sanity: row sums of alpha are the abscissae, last rows sum to 1
        print('// %s: alpha row sums' % name,
              [mp.nstr(sum(alpha[i, j] for j in range(i)), 8) for i in range(s)])
        print('// %s: gamma row sums' % name,
              [mp.nstr(sum(G[i, j] for j in range(i + 1)), 8) for i in range(s)])
        for i in range(1, s):
            for j in range(i):
                print('constexpr double B%d%d = %s;' % (
                    i + 1, j + 1, mp.nstr(beta[i, j], 20)))


if __name__ == '__main__':
    main()
