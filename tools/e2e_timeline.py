"""Dev helper (GPU): per-chunk timeline of the host-buffer entry point (CHI_B200_TIMELINE=1)."""
import sys, os
import numpy as np
sys.path.insert(0, '.')
from chi_b200 import _lib, _synthetic
prob = _synthetic.two_compartment_problem(10000, seed=0, n_doses=3, rtol=1e-8, atol=1e-10)
ll = prob['log_likelihood']
C = 1024
X = _synthetic.spread_parameter_vectors(prob, C, 0.1, seed=0)
Xp = _lib.pinned_empty(X.shape); Xp[:] = X
lib = _lib.lib(); h = ll._device.problem.handle
score, status, grad = ll._host_buffers(C, True)
a = (Xp.ctypes.data_as(_lib.c_f64p), 1, score.ctypes.data_as(_lib.c_f64p), grad.ctypes.data_as(_lib.c_f64p), status.ctypes.data_as(_lib.c_i32p))
for ch in (int(v) for v in sys.argv[1:]):
    ll.set_batching(ch, -1, -1)
    for i in range(3):
        sys.stderr.write('--- chunks %d, call %d\n' % (ch, i)); sys.stderr.flush()
        lib.chi_b200_hier_logpdf(h, C, *a)
