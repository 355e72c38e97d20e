"""Dev helper (GPU): a few single-vector evaluations (for ncu)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from chi_b200 import _synthetic
prob = _synthetic.two_compartment_problem(1000, seed=0, n_doses=3, rtol=1e-8, atol=1e-10)
post = prob['log_posterior']
x = prob['x0']
fwd = len(sys.argv) > 1 and sys.argv[1] == 'fwd'
for _ in range(6):
    r = post(x) if fwd else post.evaluateS1(x)
print('ok', r if fwd else r[0])
