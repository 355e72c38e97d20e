"""Dev helper (GPU): where the time of a single-vector evaluateS1 goes."""
import time, sys
import numpy as np
sys.path.insert(0, '.')
from chi_b200 import _lib, _synthetic

def bench(f, n=300):
    for _ in range(20): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6

for n_ids in ([int(v) for v in sys.argv[1:]] or [1000]):
    prob = _synthetic.two_compartment_problem(n_ids, seed=0, n_doses=3, rtol=1e-8, atol=1e-10)
    post, ll, prior = prob['log_posterior'], prob['log_likelihood'], prob['log_prior']
    x = prob['x0']
    nb = ll._n_bottom
    print('n_ids', n_ids)
    print('  post.evaluateS1        %.1f us' % bench(lambda: post.evaluateS1(x)))
    print('  ll.evaluateS1          %.1f us' % bench(lambda: ll.evaluateS1(x)))
    print('  prior.evaluateS1       %.1f us' % bench(lambda: prior.evaluateS1(x[nb:])))
    print('  post.__call__          %.1f us' % bench(lambda: post(x)))
    print('  ll.__call__            %.1f us' % bench(lambda: ll(x)))
    lib = _lib.lib()
    X = _lib.pinned_empty((1, len(x))); X[0] = x
    score, status, grad = ll._host_buffers(1, True)
    h = ll._device.problem.handle
    pX = X.ctypes.data_as(_lib.c_f64p); pS = score.ctypes.data_as(_lib.c_f64p)
    pG = grad.ctypes.data_as(_lib.c_f64p); pT = status.ctypes.data_as(_lib.c_i32p)
    print('  raw C call with grad   %.1f us' % bench(lambda: lib.chi_b200_hier_logpdf(h, 1, pX, 1, pS, pG, pT)))
    print('  raw C call forward     %.1f us' % bench(lambda: lib.chi_b200_hier_logpdf(h, 1, pX, 0, pS, None, pT)))
    ll.set_timing(True)
    for _ in range(50): lib.chi_b200_hier_logpdf(h, 1, pX, 1, pS, pG, pT)
    ms, n = ll.timings(); print('  solve kernel with grad %.1f us (events)' % (ms / n * 1e3))
    ll.set_timing(True)
    for _ in range(50): lib.chi_b200_hier_logpdf(h, 1, pX, 0, pS, None, pT)
    ms, n = ll.timings(); print('  solve kernel forward   %.1f us (events)' % (ms / n * 1e3))
    ll.set_timing(False)
    print('  steps/system', ll.counters())
