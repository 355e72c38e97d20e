"""experiment (CPU, host port): error of the summed log-likelihood and its gradient against the exact
solution as a function of the integrator tolerance, all-columns form and split form"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cport, ode_exact, stats
import bench
n_ids, n_vec = 60, 2
times, events = bench.host_design(n_ids, 0, 10, 3)
key = cport.find_key('two_compartment_pk_model', 'central', True)
rng = np.random.default_rng(5)
omodel = ode_exact.two_compartment(True); omodel.set_outputs(['central.drug_concentration'])
psi = np.zeros((n_vec * n_ids, 7))
psi[:, 2:] = bench.TWO_CMT_TYPICAL * np.exp(0.3 * rng.normal(size=(n_vec * n_ids, 5)))
psi[:, :2] = np.abs(0.01 * rng.normal(size=(n_vec * n_ids, 2)))
ids = (np.arange(n_vec * n_ids) % n_ids).astype(np.int32)
# observations from the exact solution with 10 % log-normal noise
obs = []
for i in range(n_ids):
    y = ode_exact.simulate(omodel, psi[i], times[i], events[i], sensitivities=False)
    obs.append(np.asarray(y).reshape(-1) * rng.lognormal(0, 0.1, size=len(times[i])))
pop = cport.Population(key, ['central.drug_concentration'], ['lognormal'], times, obs, None, events)
x = np.column_stack([psi, np.full(len(psi), 0.1)])
o_ll = np.empty(len(x)); o_g = np.empty((len(x), 8))
for k in range(len(x)):
    i = ids[k]
    sim = lambda p, s, i=i: ode_exact.simulate(omodel, p, times[i], events[i], sensitivities=s)
    o_ll[k], o_g[k] = stats.loglikelihood_s1(sim, 7, ['lognormal'], [obs[i]], np.ones((1, 10), bool), x[k], True)
for rtol in (1e-8, 2e-8, 3e-8, 5e-8, 1e-7):
    for name, kw in (('all', dict(stream_cb=11, deferred=False)), ('split', dict(stream_cb=111, split=True, deferred=False))):
        ll, g, st, nst, nrj = pop.loglik(x, ids=ids, rtol=rtol, atol=rtol * 1e-2, **kw)
        print('rtol %.0e %-5s score rel %.2e  grad rel %.2e  ll max abs %.2e  attempts %.1f' % (
            rtol, name, abs(ll.sum() - o_ll.sum()) / abs(o_ll.sum()), np.max(np.abs(g - o_g)) / np.max(np.abs(o_g)),
            np.max(np.abs(ll - o_ll)), (nst + nrj).mean()))
