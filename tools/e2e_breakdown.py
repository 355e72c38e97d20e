"""Dev helper (GPU): where the end-to-end time of a batched evaluateS1 goes (north-star size)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from chi_b200 import _lib, _synthetic
n_ids = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
C = 1024
prob = _synthetic.two_compartment_problem(n_ids, seed=0, n_doses=3, rtol=1e-8, atol=1e-10)
post, ll = prob['log_posterior'], prob['log_likelihood']
X = _synthetic.spread_parameter_vectors(prob, C, 0.1, seed=0)
Xp = _lib.pinned_empty(X.shape); Xp[:] = X
lib = _lib.lib(); h = ll._device.problem.handle
score, status, grad = ll._host_buffers(C, True)
pX = Xp.ctypes.data_as(_lib.c_f64p); pS = score.ctypes.data_as(_lib.c_f64p)
pG = grad.ctypes.data_as(_lib.c_f64p); pT = status.ctypes.data_as(_lib.c_i32p)
def bench(f, n=5):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print('post.evaluateS1_batch(copy=False)  %.2f ms' % bench(lambda: post.evaluateS1_batch(Xp, copy=False)))
print('ll.evaluateS1_batch(copy=False)    %.2f ms' % bench(lambda: ll.evaluateS1_batch(Xp, copy=False)))
print('raw C call with grad               %.2f ms' % bench(lambda: lib.chi_b200_hier_logpdf(h, C, pX, 1, pS, pG, pT)))
print('raw C call with grad, pageable X   %.2f ms' % bench(lambda: lib.chi_b200_hier_logpdf(h, C, X.ctypes.data_as(_lib.c_f64p), 1, pS, pG, pT)))
print('raw C call forward                 %.2f ms' % bench(lambda: lib.chi_b200_hier_logpdf(h, C, pX, 0, pS, None, pT)))
for ch in (1, 2, 4, 12):
    ll.set_batching(ch, -1, -1)
    print('chunks %2d: raw C call with grad     %.2f ms' % (ch, bench(lambda: lib.chi_b200_hier_logpdf(h, C, pX, 1, pS, pG, pT))))
ll.set_batching(-1, -1, -1)
ll.set_timing(True)
for _ in range(3): lib.chi_b200_hier_logpdf(h, C, pX, 1, pS, pG, pT)
ms, n = ll.timings(); print('solve kernel (timing mode, 1 chunk) %.2f ms' % (ms / n))
