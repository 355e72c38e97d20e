"""Runs the host build of the per-system solve (the same integrator headers
the kernels are compiled from) under AddressSanitizer over the model kinds,
methods, observation paths and modes.  Executed by tests/test_port_sanitizer.py
in a subprocess with libasan preloaded; argv[1] is the instrumented library."""
import ctypes
import sys

import numpy as np

import oracle.cport as cp

lib = ctypes.CDLL(sys.argv[1])
lib.chi_port_solve.argtypes = [
    ctypes.c_char_p, ctypes.POINTER(cp.SolveArgs), ctypes.c_int]
lib.chi_port_solve2.argtypes = [
    ctypes.c_char_p, ctypes.POINTER(cp.SolveArgs), ctypes.c_int, ctypes.c_int]
lib.chi_port_snap_width.argtypes = [
    ctypes.c_char_p, ctypes.POINTER(cp.SolveArgs)]
lib.chi_port_fill_integrated.argtypes = [
    ctypes.c_char_p, ctypes.POINTER(cp.SolveArgs)]
assert lib.chi_port_sizeof_args() == ctypes.sizeof(cp.SolveArgs)
cp._lib = lib

rng = np.random.default_rng(0)
CASES = [
    ('two_compartment_pk_model', True, 'central.drug_concentration',
     [0, 0, 2, 1, 5, 1, 10.0]),
    ('two_compartment_pk_model', False, 'central.drug_concentration',
     [0, 0, 0, 2, 1.5, 1, 5, 1, 10.0]),
    ('one_compartment_pk_model', True, 'central.drug_concentration',
     [0, 2, 1.0]),
    ('erlotinib_tumour_growth_inhibition_model', True, 'global.tumour_volume',
     [0.0, 0.2, 2.0, 1.0, 0.5, 0.3, 0.1]),
]
for name, direct, output, psi0 in CASES:
    key = cp.find_key(name, 'central', direct)
    n = 8
    # ragged records, one individual without any, records on dose edges
    times = [np.sort(rng.uniform(0.25, 24, rng.integers(0, 7)))
             for _ in range(n)]
    times[0] = np.array([1.0, 8.0, 9.0])
    regimens = [[(10., 0., 1., 8., 3.)]] * n
    obs = [np.abs(rng.normal(1, 0.3, len(t))) + 0.05 for t in times]
    for kind, sigma in (('lognormal', [0.2]), ('cam', [0.1, 0.2])):
        pop = cp.Population(key, [output], [kind], times, obs, None, regimens)
        x = np.tile(np.concatenate([psi0, sigma]), (n, 1))
        for method in (5, 1, 0):
            for deferred in (True, False):
                if method == 0 and deferred:
                    continue
                for with_grad in (True, False):
                    res = pop.loglik(x, stream_cb=method, deferred=deferred,
                                     with_grad=with_grad, n_threads=2)
                    assert np.all(res[2] == 0)
        pop.simulate(np.tile(psi0, (n, 1)), stream_cb=5)
        pop.simulate(np.tile(psi0, (n, 1)), stream_cb=0, with_sens=False)
print('SANITIZER RUN OK')
