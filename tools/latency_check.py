"""Latency of a single-vector evaluateS1 (what pints calls) of the
1000-individual C2 model for every kernel variant (dev helper, GPU)."""
import time
import numpy as np
from chi_b200 import _lib, _synthetic

lib = _lib.lib()
for n_ids in (100, 1000, 10000):
    prob = _synthetic.two_compartment_problem(
        n_ids, seed=0, n_doses=3, rtol=1e-6, atol=1e-8)
    post = prob['log_posterior']
    ll = prob['log_likelihood']
    x = prob['x0']
    h = ll._device.problem.handle
    model_id = lib.chi_b200_model_find(prob['model'].model_key().encode())
    variants = np.zeros((2, 8), dtype=np.int32)
    nv = np.zeros(1, dtype=np.int32)
    lib.chi_b200_model_variants(
        model_id, nv.ctypes.data_as(_lib.c_i32p),
        variants[0].ctypes.data_as(_lib.c_i32p),
        variants[1].ctypes.data_as(_lib.c_i32p), 8)
    ref = None
    for v in range(int(nv[0])):
        _lib.check(lib.chi_b200_problem_set_variant(h, v))
        for _ in range(3):
            s, g = post.evaluateS1(x)
        t0 = time.perf_counter()
        for _ in range(20):
            s, g = post.evaluateS1(x)
        dt = (time.perf_counter() - t0) / 20
        t0 = time.perf_counter()
        for _ in range(20):
            f = post(x)
        dtf = (time.perf_counter() - t0) / 20
        if ref is None:
            ref = s
        print('n_ids %5d variant %d (lanes %d, cols/method %d): evaluateS1 '
              '%.3f ms, __call__ %.3f ms, score rel diff %.1e' % (
                  n_ids, v, variants[0][v], variants[1][v], dt * 1e3,
                  dtf * 1e3, abs(s - ref) / abs(ref)))
