"""Dev helper (GPU): which vectors fail / go NaN under widely spread vectors."""
import sys
import numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from chi_b200 import _synthetic
import test_gpu_bench_tolerance as T

prob = _synthetic.two_compartment_problem(300, seed=13, n_doses=3, error='cam', covariates=True, rtol=1e-6, atol=1e-8)
X = _synthetic.spread_parameter_vectors(prob, 8, 0.1, seed=3)
s, g = prob['log_posterior'].evaluateS1_batch(X)
ref, gref = T.oracle_scores(prob, X, 'cam', True)
print('gpu', s); print('ref', ref)
print('nan in X', np.isnan(X).any(), 'prior', [prob['log_prior'](x[1500:]) for x in X])

for n_ids, C in ((10000, 1024), (10000, 128), (1000, 1024)):
    prob = _synthetic.two_compartment_problem(n_ids, seed=0, n_doses=3, rtol=1e-6, atol=1e-8)
    ll = prob['log_likelihood']
    X = _synthetic.spread_parameter_vectors(prob, C, 0.1, seed=100)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        score, grad, status = ll._evaluate(X, True)
    bad = np.flatnonzero(status != 0)
    print('n_ids', n_ids, 'C', C, 'bad vectors', len(bad), 'status', status[bad][:10], 'finite', np.isfinite(score).all())
    if len(bad):
        c = bad[0]
        from oracle import stats
        lay = stats.PopLayout([('pooled', 2), ('lognormal', 5, False), ('pooled', 1)], n_ids)
        x = X[c]
        eta = stats.pop_individual_parameters(lay, x[lay.n_bottom:], x[:lay.n_bottom], None, return_eta=True)
        psi = stats.pop_individual_parameters(lay, x[lay.n_bottom:], eta, None)
        g_ll, g_grad, st = ll._device.loglik(psi, ids=np.arange(n_ids, dtype=np.int32), with_grad=True)
        b2 = np.flatnonzero(st != 0)
        print('  per-system path: bad', len(b2), st[b2][:5], psi[b2][:3], 'top', x[lay.n_bottom:])
