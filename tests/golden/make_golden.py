"""Generates tests/golden/*.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The reference's statistical modules (``chi/_error_models.py``,
``_population_models.py``, ``_covariate_models.py``, ``_log_pdfs.py``) are
loaded by path through ``oracle/refload.py`` (stub ``myokit``/``pints``) and
evaluated on seeded random inputs.  Everything above the ODE solve therefore
comes from chi's own code.  The ODE solve itself (myokit/CVODES, not
installed) is replaced by a ``chi.MechanisticModel`` subclass that returns the
exact solution from ``oracle/ode_exact.py`` -- the same seam the reference's
own tests use for their toy models (chi/tests/test_log_pdfs.py:18-113).

Files written:
  stats_golden.json  error models and population models
  hier_golden.json   LogLikelihood / HierarchicalLogPosterior (configs C1,
                     C2, C5 at small sizes), keyed by parameter name
"""
import copy as copy_module
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ode_exact, refload  # noqa: E402

warnings.simplefilter('ignore')
chi = refload.load()
import pints  # noqa: E402  (the stub registered by refload)


def tolist(a):
    return np.asarray(a, dtype=float).tolist()


# --------------------------------------------------------------------------
# error models
# --------------------------------------------------------------------------

def error_model_cases(rng):
    models = [
        ('gaussian', chi.GaussianErrorModel(), 1),
        ('lognormal', chi.LogNormalErrorModel(), 1),
        ('cam', chi.ConstantAndMultiplicativeGaussianErrorModel(), 2),
        ('multiplicative', chi.MultiplicativeGaussianErrorModel(), 1)]
    out = []
    for kind, em, n_sig in models:
        for n_obs, n_par in [(1, 1), (6, 3), (25, 9)]:
            ybar = rng.uniform(0.5, 3.0, n_obs)
            obs = rng.uniform(0.5, 3.0, n_obs)
            sens = rng.normal(size=(n_obs, n_par))
            sigma = rng.uniform(0.1, 1.0, n_sig)
            ll, grad = em.compute_sensitivities(sigma, ybar, sens, obs)
            out.append({
                'kind': kind, 'sigma': tolist(sigma), 'ybar': tolist(ybar),
                'obs': tolist(obs), 'sens': tolist(sens), 'll': float(ll),
                'grad': tolist(grad),
                'll_call': float(em.compute_log_likelihood(sigma, ybar, obs)),
                'pointwise': tolist(
                    em.compute_pointwise_ll(sigma, ybar, obs))})
    return out


# --------------------------------------------------------------------------
# population models
# --------------------------------------------------------------------------

POP_SPECS = [
    [['pooled', 2], ['lognormal', 5, False], ['pooled', 1]],
    [['lognormal', 3, True]],
    [['gaussian', 2, True], ['pooled', 1], ['gaussian', 2, False]],
    [['hetero', 1], ['lognormal', 2, False], ['pooled', 2]],
    [['pooled', 2], ['covariate', ['lognormal', 5, False], 2], ['pooled', 2]],
    [['covariate', ['gaussian', 2, True], 1], ['lognormal', 1, True]],
    [['covariate', ['lognormal', 2, True], 3], ['hetero', 2]],
    [['covariate', ['gaussian', 2, False], 2]],
]


def reduced_error_model_cases(rng):
    """ReducedErrorModel (chi/_error_models.py:1583-1906) over the CAM model
    with 'Sigma rel.' fixed, and over the Gaussian model with nothing fixed."""
    out = []
    n_obs, n_par = 8, 4
    ybar = rng.uniform(0.5, 3.0, n_obs)
    obs = rng.uniform(0.5, 3.0, n_obs)
    sens = rng.normal(size=(n_obs, n_par))
    for kind, em, fixed, free in (
            ('cam', chi.ConstantAndMultiplicativeGaussianErrorModel(),
             {'Sigma rel.': 0.15}, [0.4]),
            ('cam', chi.ConstantAndMultiplicativeGaussianErrorModel(),
             {'Sigma base': 0.2}, [0.3]),
            ('gaussian', chi.GaussianErrorModel(), {}, [0.7])):
        rem = chi.ReducedErrorModel(em)
        if fixed:
            rem.fix_parameters(fixed)
        ll, grad = rem.compute_sensitivities(
            np.array(free), ybar, sens, obs)
        out.append({
            'kind': kind, 'fixed': fixed, 'free': free,
            'names': rem.get_parameter_names(),
            'n_parameters': rem.n_parameters(),
            'n_fixed': rem.n_fixed_parameters(),
            'ybar': tolist(ybar), 'obs': tolist(obs), 'sens': tolist(sens),
            'll': float(ll), 'grad': tolist(grad),
            'll_call': float(rem.compute_log_likelihood(
                np.array(free), ybar, obs)),
            'sample_seed5': tolist(rem.sample(
                np.array(free), ybar, n_samples=2, seed=5))})
    return out


def truncated_gaussian_cases(rng):
    """TruncatedGaussianModel stand-alone (the reference does not implement
    compute_individual_parameters for it, so it cannot sit inside a
    ComposedPopulationModel there): log-likelihood, sensitivities in both
    shapes, moments and seeded samples."""
    out = []
    for n_dim, n_ids in ((1, 7), (3, 5)):
        pm = chi.TruncatedGaussianModel(n_dim=n_dim)
        theta = np.concatenate([rng.uniform(-0.5, 2.0, n_dim),
                                rng.uniform(0.3, 1.5, n_dim)])
        obs = rng.uniform(0.05, 3.0, (n_ids, n_dim))
        dlp = rng.normal(size=(n_ids, n_dim))
        ll = pm.compute_log_likelihood(theta, obs)
        s_r, g_r = pm.compute_sensitivities(
            theta, obs, dlogp_dpsi=dlp.copy(), reduce=True)
        s_f, dpsi, dtheta = pm.compute_sensitivities(
            theta, obs, dlogp_dpsi=dlp.copy(), flattened=True)
        neg = obs.copy()
        neg[0, 0] = -0.1
        bad = theta.copy()
        bad[n_dim] = 0.0
        out.append({
            'n_dim': n_dim, 'n_ids': n_ids, 'names': pm.get_parameter_names(),
            'theta': tolist(theta), 'obs': tolist(obs),
            'dlogp_dpsi': tolist(dlp), 'll': float(ll),
            # the reference allocates d theta with n_parameters = 2 n_dim
            # rows per individual and fills the first two (mu, sigma); only
            # those are kept
            'score_reduced': float(s_r),
            'grad_reduced': tolist(g_r[:n_ids * n_dim + 2 * n_dim]),
            'score_flat': float(s_f), 'dpsi': tolist(dpsi),
            'dtheta': tolist(np.asarray(dtheta).reshape(
                -1, n_dim)[:2].flatten()),
            'll_negative_obs': float(pm.compute_log_likelihood(theta, neg)),
            'll_zero_sigma': float(pm.compute_log_likelihood(bad, obs)),
            # the reference allocates (n_parameters, n_dim) and fills two rows
            'mean_and_std': tolist(pm.get_mean_and_std(theta)[:2]),
            'samples_seed3': tolist(pm.sample(theta, n_samples=4, seed=3))})
    return out


def build_ref_population(spec):
    models = []
    for b in spec:
        if b[0] == 'pooled':
            models.append(chi.PooledModel(n_dim=b[1]))
        elif b[0] == 'hetero':
            models.append(chi.HeterogeneousModel(n_dim=b[1]))
        elif b[0] == 'gaussian':
            models.append(chi.GaussianModel(n_dim=b[1], centered=b[2]))
        elif b[0] == 'lognormal':
            models.append(chi.LogNormalModel(n_dim=b[1], centered=b[2]))
        else:
            inner = build_ref_population([b[1]])[0]
            models.append(chi.CovariatePopulationModel(
                inner, chi.LinearCovariateModel(n_cov=b[2])))
    return models


def population_cases(rng):
    out = []
    n_ids = 5
    for spec in POP_SPECS:
        pm = chi.ComposedPopulationModel(build_ref_population(spec))
        pm.set_n_ids(n_ids)
        n_b, n_t = pm.n_hierarchical_parameters(n_ids)
        n_cov = pm.n_covariates()
        cov = rng.normal(size=(n_ids, n_cov)) if n_cov else None
        names = pm.get_parameter_names()
        top = rng.uniform(0.2, 1.5, n_t)
        for i, name in enumerate(names):
            if 'Cov.' in name:
                top[i] *= 0.05
        bottom = rng.uniform(0.3, 2.0, n_b)
        eta = pm.compute_individual_parameters(
            top, bottom, cov, return_eta=True)
        psi = pm.compute_individual_parameters(top, eta, cov)
        score = pm.compute_log_likelihood(top, eta, cov)
        dlp = rng.normal(size=(n_ids, pm.n_dim()))
        s1, grad = pm.compute_sensitivities(
            top, eta, dlogp_dpsi=dlp, reduce=True, covariates=cov)
        selections = []
        for m in pm.get_population_models():
            if isinstance(m, chi.CovariatePopulationModel):
                pidx, didx = \
                    m._covariate_model.get_set_population_parameters()
                selections.append([tolist(pidx), tolist(didx)])
        out.append({
            'spec': spec, 'n_ids': n_ids, 'names': names,
            'selections': selections,
            'covariates': None if cov is None else tolist(cov),
            'top': tolist(top), 'bottom': tolist(bottom),
            'eta': tolist(eta), 'psi': tolist(psi), 'score': float(score),
            'dlogp_dpsi': tolist(dlp), 'score_s1': float(s1),
            'grad': tolist(grad), 'n_bottom': int(n_b), 'n_top': int(n_t),
            'special_dims': [list(map(lambda v: int(v) if not isinstance(
                v, bool) else v, e)) for e in pm.get_special_dims()[0]]})
    return out


# --------------------------------------------------------------------------
# log-pdfs on top of the exact ODE solution
# --------------------------------------------------------------------------

class ExactModel(chi.MechanisticModel):
    """chi.MechanisticModel whose simulate returns the exact ODE solution."""

    def __init__(self, omodel, events):
        super(ExactModel, self).__init__()
        self._m = omodel
        self._events = events
        self._sens = False

    def copy(self):
        return ExactModel(self._m, self._events)

    def enable_sensitivities(self, enabled, parameter_names=None):
        self._sens = bool(enabled)
        self._cols = None
        if enabled and parameter_names is not None:
            self._cols = [self._m.parameter_names.index(n)
                          for n in self._m.parameter_names
                          if n in parameter_names]

    def has_sensitivities(self):
        return self._sens

    def n_outputs(self):
        return len(self._m.outputs)

    def n_parameters(self):
        return self._m.P

    def outputs(self):
        return list(self._m.outputs)

    def parameters(self):
        return list(self._m.parameter_names)

    def simulate(self, parameters, times):
        out = ode_exact.simulate(
            self._m, parameters, times, self._events,
            sensitivities=self._sens)
        if self._sens and getattr(self, '_cols', None) is not None:
            return out[0], out[1][:, :, self._cols]
        return out


class IndependentPrior(pints.LogPrior):
    """Independent Gaussian (kind 0) / log-normal (kind 1) priors."""

    def __init__(self, kinds, locs, scales):
        self.kinds, self.locs, self.scales = kinds, locs, scales

    def n_parameters(self):
        return len(self.kinds)

    def evaluateS1(self, x):
        score = 0.0
        grad = np.zeros(len(x))
        for k, (kind, m, s) in enumerate(
                zip(self.kinds, self.locs, self.scales)):
            if kind == 0:
                z = (x[k] - m) / s
                score += -0.5 * np.log(2 * np.pi) - np.log(s) - 0.5 * z * z
                grad[k] = -z / s
            else:
                z = (np.log(x[k]) - m) / s
                score += -0.5 * np.log(2 * np.pi) - np.log(s) \
                    - np.log(x[k]) - 0.5 * z * z
                grad[k] = -(1 + z / s) / x[k]
        return score, grad

    def __call__(self, x):
        return self.evaluateS1(x)[0]


def c1_case(rng):
    """config C1: one-compartment PK, bolus, Gaussian error, 100 times."""
    omodel = ode_exact.one_compartment(True)
    events = [(10 / 0.01, 0.0, 0.01, 0.0, 0.0)]
    times = np.linspace(0.1, 10, 100)
    psi = np.array([0.0, 2.0, 1.0])
    y = ode_exact.simulate(omodel, psi, times, events, False)[0]
    obs = y + 0.1 * rng.normal(size=len(y))
    ll = chi.LogLikelihood(
        ExactModel(omodel, events), chi.GaussianErrorModel(), obs, times)
    prior = IndependentPrior([1, 1, 1, 1], [0.0, np.log(2), 0.0, np.log(.1)],
                             [1.0, 0.5, 0.5, 0.5])
    post = chi.LogPosterior(ll, prior)
    x = np.array([0.3, 2.2, 0.9, 0.12])
    score, grad = post.evaluateS1(x)
    return {
        'model': 'one_compartment_pk_model', 'direct': True,
        'output': 'central.drug_concentration', 'error': 'gaussian',
        'events': [list(e) for e in events], 'times': tolist(times),
        'obs': tolist(obs), 'names': ll.get_parameter_names(),
        'prior': [prior.kinds, tolist(prior.locs), tolist(prior.scales)],
        'x': tolist(x), 'score': float(score), 'grad': tolist(grad),
        'score_call': float(post(x))}


def hier_case(rng, n_ids, error, covariates):
    """configs C2 / C5 at small size."""
    omodel = ode_exact.two_compartment(True)
    mu_log = np.log([2.0, 1.0, 5.0, 1.0, 10.0])
    sigma_log = np.full(5, 0.3)
    times, events, obs_all = [], [], []
    eta = rng.normal(size=(n_ids, 5))
    if error == 'lognormal':
        em = chi.LogNormalErrorModel()
        sig = [0.1]
    else:
        em = chi.ConstantAndMultiplicativeGaussianErrorModel()
        sig = [0.05, 0.1]
    lls = []
    for i in range(n_ids):
        t = np.sort(rng.uniform(0.25, 24.0, 8))
        ev = [(rng.uniform(5, 15) / 1.0, 0.0, 1.0, 8.0, 3.0)]
        psi = np.concatenate(
            [[0.0, 0.0], np.exp(mu_log + sigma_log * eta[i])])
        y = ode_exact.simulate(omodel, psi, t, ev, False)[0]
        if error == 'lognormal':
            obs = y * rng.lognormal(-0.005, 0.1, size=len(y))
        else:
            obs = np.abs(y + (0.05 + 0.1 * y) * rng.normal(size=len(y)))
        ll = chi.LogLikelihood(ExactModel(omodel, ev), em, obs, t)
        ll.set_id('id %d' % i)
        lls.append(ll)
        times.append(tolist(t))
        events.append([list(e) for e in ev])
        obs_all.append(tolist(obs))
    names = lls[0].get_parameter_names()
    if covariates:
        middle = chi.CovariatePopulationModel(
            chi.LogNormalModel(n_dim=5, centered=False),
            chi.LinearCovariateModel(n_cov=2))
        X = rng.normal(size=(n_ids, 2))
    else:
        middle = chi.LogNormalModel(n_dim=5, centered=False)
        X = None
    pm = chi.ComposedPopulationModel(
        [chi.PooledModel(2), middle, chi.PooledModel(len(sig))])
    pm.set_dim_names(names)
    hll = chi.HierarchicalLogLikelihood(lls, pm, covariates=X)
    top_names = hll.get_parameter_names(exclude_bottom_level=True)
    n_top = len(top_names)
    top = np.zeros(n_top)
    kinds, locs, scales = [], [], []
    for k, name in enumerate(top_names):
        if 'Cov.' in name:
            top[k] = 0.03 * rng.normal()
            kinds.append(0), locs.append(0.0), scales.append(0.5)
        elif name.startswith('Log mean'):
            d = names.index(name[len('Log mean '):]) - 2
            top[k] = mu_log[d]
            kinds.append(0), locs.append(mu_log[d]), scales.append(1.0)
        elif name.startswith('Log std.'):
            top[k] = 0.3
            kinds.append(1), locs.append(np.log(0.3)), scales.append(0.5)
        elif 'Sigma' in name:
            s = sig[0] if ('log' in name or 'base' in name) else sig[-1]
            top[k] = s
            kinds.append(1), locs.append(np.log(s)), scales.append(0.5)
        else:  # pooled initial amounts
            top[k] = 0.05
            kinds.append(0), locs.append(0.0), scales.append(1.0)
    prior = IndependentPrior(kinds, locs, scales)
    post = chi.HierarchicalLogPosterior(hll, prior)
    x = np.concatenate([eta.reshape(-1), top])
    x = x * (1 + 0.02 * rng.normal(size=len(x)))
    score, grad = post.evaluateS1(x)
    return {
        'model': 'two_compartment_pk_model', 'direct': True,
        'output': 'central.drug_concentration', 'error': error,
        'n_ids': n_ids, 'times': times, 'events': events, 'obs': obs_all,
        'covariates': None if X is None else tolist(X),
        'll_names': names,
        'names': post.get_parameter_names(include_ids=True),
        'top_names': top_names,
        'prior': [kinds, tolist(locs), tolist(scales)],
        'x': tolist(x), 'score': float(score), 'grad': tolist(grad),
        'score_call': float(post(x))}


def predictive_case(rng, name):
    """configs C3 / C5 forward sampling at small size:
    PopulationPredictiveModel.sample(return_df=False)."""
    if name == 'c3':
        omodel = ode_exact.tumour_growth_inhibition_pkpd(True)
        events = [(5 / 0.01, 0.0, 0.01, 1.0, 30.0)]
        em = chi.LogNormalErrorModel()
        # [A0, V_T0, size, critical_volume, k_e, kappa, lambda | sigma]
        pm = chi.ComposedPopulationModel([
            chi.PooledModel(1), chi.LogNormalModel(1), chi.PooledModel(2),
            chi.LogNormalModel(3, centered=False), chi.PooledModel(1)])
        theta = [0.0, np.log(0.2), 0.2, 2.0, 1.0,
                 np.log(0.5), np.log(0.3), np.log(0.1), 0.2, 0.3, 0.1, 0.1]
        times = np.arange(0.0, 31.0)
        X = None
        n_samples = 6
        model, direct, output = (
            'erlotinib_tumour_growth_inhibition_model', True,
            'global.tumour_volume')
        error = 'lognormal'
        spec = [['pooled', 1], ['lognormal', 1, True], ['pooled', 2],
                ['lognormal', 3, False], ['pooled', 1]]
    else:
        omodel = ode_exact.two_compartment(True)
        events = [(10.0, 0.0, 1.0, 8.0, 3.0)]
        em = chi.ConstantAndMultiplicativeGaussianErrorModel()
        pm = chi.ComposedPopulationModel([
            chi.PooledModel(2),
            chi.CovariatePopulationModel(
                chi.LogNormalModel(5, centered=False),
                chi.LinearCovariateModel(n_cov=2)),
            chi.PooledModel(2)])
        n_top = pm.n_parameters()
        theta = np.zeros(n_top)
        names = pm.get_parameter_names()
        mu = np.log([2.0, 1.0, 5.0, 1.0, 10.0])
        for k, nm in enumerate(names):
            if 'Cov.' in nm:
                theta[k] = 0.05 * rng.normal() if nm.startswith(
                    'Log mean') else 0.0
            elif nm.startswith('Log mean'):
                theta[k] = mu[int(nm.split('Dim. ')[1]) - 3]
            elif nm.startswith('Log std.'):
                theta[k] = 0.3
        theta[-2:] = [0.05, 0.1]
        times = np.array([0.5, 1.0, 2.0, 4.0, 8.5, 12.0, 20.0, 24.0])
        n_samples = 5
        X = rng.normal(size=(n_samples, 2))
        model, direct, output = (
            'two_compartment_pk_model', True, 'central.drug_concentration')
        error = 'cam'
        spec = [['pooled', 2], ['covariate', ['lognormal', 5, False], 2],
                ['pooled', 2]]
        theta = list(theta)
    omodel.set_outputs([output])
    pred = chi.PredictiveModel(ExactModel(omodel, events), [em])
    ppm = chi.PopulationPredictiveModel(pred, pm)
    kwargs = {} if X is None else {'covariates': X}
    samples = ppm.sample(
        theta, times, n_samples=n_samples, seed=11, return_df=False,
        **kwargs)
    return {
        'model': model, 'direct': direct, 'output': output, 'error': error,
        'events': [list(e) for e in events], 'spec': spec,
        'theta_names': pm.get_parameter_names(), 'theta': tolist(theta),
        'times': tolist(times), 'n_samples': n_samples, 'seed': 11,
        'covariates': None if X is None else tolist(X),
        'samples': tolist(samples)}


# --------------------------------------------------------------------------
# round 2: stand-alone covariate model, pointwise log-likelihood, reduced
# models inside a log-likelihood, a 120-individual hierarchical case,
# population filters and their posterior, DataFrame ingestion
# --------------------------------------------------------------------------

def linear_covariate_cases(rng):
    out = []
    for n_cov, n_pop, n_dim, sel, n_ids in (
            (1, 2, 1, None, 4), (2, 2, 5, [[0, 1], [1, 3], [0, 4], [1, 1]], 7),
            (3, 2, 3, [[1, 2], [0, 0], [0, 2], [1, 2]], 5)):
        cm = chi.LinearCovariateModel(n_cov=n_cov)
        if sel is not None:
            cm.set_population_parameters(sel)
        pidx, didx = cm.get_set_population_parameters()
        beta = rng.normal(size=cm.n_parameters())
        pop = rng.uniform(0.2, 2.0, (n_pop, n_dim))
        X = rng.normal(size=(n_ids, n_cov))
        vartheta = cm.compute_population_parameters(beta, pop, X)
        dl = rng.normal(size=(n_ids, n_pop, n_dim))
        dpop, dbeta = cm.compute_sensitivities(beta, pop, X, dl)
        out.append({
            'n_cov': n_cov, 'selection': sel, 'pidx': tolist(pidx),
            'didx': tolist(didx), 'names': cm.get_parameter_names(),
            'beta': tolist(beta), 'pop': tolist(pop), 'covariates': tolist(X),
            'vartheta': tolist(vartheta), 'dlogp_dvartheta': tolist(dl),
            'dpop': tolist(dpop), 'dbeta': tolist(dbeta)})
    return out


def pointwise_and_reduced_cases(rng):
    """LogLikelihood.compute_pointwise_ll (chi/_log_pdfs.py:932-971) and a
    log-likelihood over a ReducedMechanisticModel / ReducedErrorModel
    (chi/_mechanistic_models.py:867-1206, chi/_error_models.py:1583-1906)."""
    omodel = ode_exact.two_compartment(True)
    omodel.set_outputs(
        ['central.drug_concentration', 'peripheral.drug_concentration'])
    events = [(8.0, 0.0, 1.0, 6.0, 3.0)]
    t0 = np.sort(rng.uniform(0.3, 20, 9))
    t1 = np.sort(rng.uniform(0.3, 20, 6))
    psi = np.array([0.1, 0.05, 2.0, 1.0, 5.0, 1.0, 10.0])
    y = ode_exact.simulate(omodel, psi, np.union1d(t0, t1), events, False)
    grid = np.union1d(t0, t1)
    obs0 = np.interp(t0, grid, y[0]) * rng.lognormal(0, 0.1, len(t0))
    obs1 = np.abs(np.interp(t1, grid, y[1])
                  + 0.05 * rng.normal(size=len(t1))) + 1e-3
    ems = [chi.LogNormalErrorModel(),
           chi.ConstantAndMultiplicativeGaussianErrorModel()]
    ll = chi.LogLikelihood(
        ExactModel(omodel, events), ems, [obs0, obs1], [t0, t1])
    x = np.concatenate([psi * (1 + 0.03 * rng.normal(size=7)),
                        [0.12, 0.04, 0.15]])
    x[:2] = [0.1, 0.05]
    pointwise = ll.compute_pointwise_ll(x)
    score, grad = ll.evaluateS1(x)
    case = {
        'model': 'two_compartment_pk_model', 'direct': True,
        'outputs': list(omodel.outputs), 'errors': ['lognormal', 'cam'],
        'events': [list(e) for e in events],
        'times': [tolist(t0), tolist(t1)], 'obs': [tolist(obs0), tolist(obs1)],
        'names': ll.get_parameter_names(), 'x': tolist(x),
        'pointwise': tolist(pointwise), 'score': float(score),
        'grad': tolist(grad)}
    # the same data under reduced models: two mechanistic parameters and
    # one error parameter fixed
    rmm = chi.ReducedMechanisticModel(ExactModel(omodel, events))
    fixed_mech = {'central.drug_amount': 0.1, 'global.transition_rate_p2c': 1.1}
    rmm.fix_parameters(fixed_mech)
    rem = chi.ReducedErrorModel(
        chi.ConstantAndMultiplicativeGaussianErrorModel())
    rem.fix_parameters({'Sigma rel.': 0.15})
    llr = chi.LogLikelihood(
        rmm, [chi.LogNormalErrorModel(), rem], [obs0, obs1], [t0, t1])
    names = llr.get_parameter_names()
    xr = np.array([0.05, 2.1, 0.9, 5.2, 9.5, 0.11, 0.05])
    score_r, grad_r = llr.evaluateS1(xr)
    case['reduced'] = {
        'fixed_mechanistic': fixed_mech, 'fixed_error': {'Sigma rel.': 0.15},
        'names': names, 'x': tolist(xr), 'score': float(score_r),
        'grad': tolist(grad_r), 'score_call': float(llr(xr))}
    return case


def filter_cases(rng):
    obs = rng.uniform(0.5, 3.0, (7, 2, 4))
    obs[2, 1, 3] = np.nan
    obs[0, 0, 0] = np.nan
    y = rng.uniform(0.5, 3.0, (12, 2, 4))
    out = []
    for name, kw in (('GaussianFilter', {}), ('LogNormalFilter', {}),
                     ('GaussianKDEFilter', {}), ('LogNormalKDEFilter', {}),
                     ('GaussianMixtureFilter', {'n_kernels': 3})):
        f = getattr(chi, name)(obs, **kw)
        s, g = f.compute_sensitivities(y)
        out.append({'filter': name, 'kwargs': kw, 'score': float(s),
                    'score_call': float(f.compute_log_likelihood(y)),
                    'grad': tolist(g)})
    comp = chi.ComposedPopulationFilter(
        [chi.GaussianFilter(obs[:, :, :2]),
         chi.LogNormalKDEFilter(obs[:, :, 2:])])
    order = np.array([2, 0, 3, 1])
    comp.sort_times(order)
    s, g = comp.compute_sensitivities(y)
    out.append({'filter': 'composed', 'order': tolist(order),
                'score': float(s), 'grad': tolist(g)})
    return {'observations': [[[None if np.isnan(v) else float(v) for v in r]
                              for r in o] for o in obs],
            'simulated': tolist(y), 'cases': out}


class SamplingPrior(IndependentPrior):
    def sample(self, n=1):
        return np.ones((n, len(self.kinds)))


def filter_posterior_case(rng, log_scale, sigma_fixed):
    """PopulationFilterLogPosterior (chi/_log_pdfs.py:1325-2075) on the
    exact solution of the one-compartment model."""
    omodel = ode_exact.one_compartment(True)
    events = [(10.0 / 0.5, 0.0, 0.5, 0.0, 0.0)]
    times = np.array([4.0, 0.5, 1.5, 8.0])
    n_data, n_sim = 9, 6
    data = np.empty((n_data, 1, len(times)))
    for i in range(n_data):
        psi = [0.0, 2.0 * np.exp(0.2 * rng.normal()),
               1.0 * np.exp(0.2 * rng.normal())]
        data[i, 0] = ode_exact.simulate(
            omodel, psi, np.sort(times), events, False)[0][
                np.argsort(np.argsort(times))] * rng.lognormal(0, 0.05, 4)
    data[3, 0, 1] = np.nan
    filt = chi.GaussianKDEFilter(data) if not log_scale \
        else chi.LogNormalFilter(data)
    pm = chi.ComposedPopulationModel([
        chi.PooledModel(1), chi.LogNormalModel(1, centered=False),
        chi.GaussianModel(1, centered=True)])
    n_top = pm.n_parameters() + (0 if sigma_fixed else 1)
    prior = SamplingPrior([0] * n_top, [0.5] * n_top, [2.0] * n_top)
    post = chi.PopulationFilterLogPosterior(
        filt, times, ExactModel(omodel, events), pm, prior,
        sigma=[0.07] if sigma_fixed else None,
        error_on_log_scale=log_scale, n_samples=n_sim)
    theta = [0.05, np.log(2.0), 0.2, 1.0, 0.15]
    if not sigma_fixed:
        theta = theta + [0.08]
    eta = np.column_stack([rng.normal(size=n_sim),
                           rng.normal(1.0, 0.15, size=n_sim)])
    eps = rng.normal(size=(n_sim, 1, len(times)))
    x = np.concatenate([theta, eta.reshape(-1), eps.reshape(-1)])
    score, grad = post.evaluateS1(x)
    return {
        'model': 'one_compartment_pk_model', 'direct': True,
        'output': 'central.drug_concentration',
        'events': [list(e) for e in events], 'times': tolist(times),
        'data': [[[None if np.isnan(v) else float(v) for v in r] for r in o]
                 for o in data],
        'filter': 'LogNormalFilter' if log_scale else 'GaussianKDEFilter',
        'population': [['pooled', 1], ['lognormal', 1, False],
                       ['gaussian', 1, True]],
        'sigma': [0.07] if sigma_fixed else None, 'log_scale': log_scale,
        'n_samples': n_sim, 'prior': [prior.kinds, prior.locs, prior.scales],
        'names': post.get_parameter_names(include_ids=True),
        'n_parameters': post.n_parameters(),
        'n_top': post.n_parameters(exclude_bottom_level=True),
        'x': tolist(x), 'score': float(score), 'grad': tolist(grad),
        'score_call': float(post(x))}


class DosedStub(chi.MechanisticModel):
    """Just enough of a PKPD model for ProblemModellingController.set_data:
    two outputs, supports dosing, remembers the protocol it is handed."""

    def __init__(self):
        super(DosedStub, self).__init__()
        self.protocol = None

    def copy(self):
        return copy_module.deepcopy(self)

    def n_outputs(self):
        return 2

    def n_parameters(self):
        return 3

    def outputs(self):
        return ['central.drug_concentration', 'global.tumour_volume']

    def parameters(self):
        return ['a', 'b', 'c']

    def supports_dosing(self):
        return True

    def has_sensitivities(self):
        return False

    def enable_sensitivities(self, enabled, parameter_names=None):
        pass

    def set_dosing_regimen(self, regimen, *args, **kwargs):
        self.protocol = regimen

    def simulate(self, parameters, times):
        return np.ones((2, len(times)))


def ingestion_case(rng):
    """chi.ProblemModellingController.set_data and what it builds per
    individual (chi/_problems.py:246-376, 623-734): measurement arrays per
    output, dose events, covariates."""
    import pandas as pd
    rows = []
    ids = [3, 'b7', 11, 4]
    for k, i in enumerate(ids):
        for t in np.sort(rng.uniform(0, 10, 5 + k)):
            rows.append([i, t, 'Conc', rng.uniform(0.5, 2), np.nan, np.nan])
        for t in np.sort(rng.uniform(0, 10, 3)):
            rows.append([i, t, 'Volume', rng.uniform(0.5, 2), np.nan, np.nan])
        rows.append([i, 2.5, 'Conc', np.nan, np.nan, np.nan])   # no value
        rows.append([i, np.nan, 'Age', 30.0 + k, np.nan, np.nan])
        rows.append([i, np.nan, 'Weight', 60.0 + 2 * k, np.nan, np.nan])
        for d in range(k + 1):
            rows.append([i, 1.0 * d, np.nan, np.nan, 5.0 + k,
                         0.5 if d % 2 == 0 else np.nan])
    data = pd.DataFrame(rows, columns=[
        'ID', 'Time', 'Observable', 'Value', 'Dose', 'Duration'])
    model = DosedStub()
    ctrl = chi.ProblemModellingController(
        model, [chi.GaussianErrorModel(), chi.LogNormalErrorModel()])
    pm = chi.ComposedPopulationModel([
        chi.PooledModel(2),
        chi.CovariatePopulationModel(
            chi.GaussianModel(1), chi.LinearCovariateModel(n_cov=2)),
        chi.PooledModel(2)])
    ctrl.set_population_model(pm)
    cov_names = ctrl.get_covariate_names()
    ctrl.set_data(
        data,
        output_observable_dict={'central.drug_concentration': 'Conc',
                                'global.tumour_volume': 'Volume'},
        covariate_dict={cov_names[0]: 'Age', cov_names[1]: 'Weight'})
    per_id = []
    for individual in ctrl._ids:
        ll = ctrl._create_log_likelihood(individual)
        per_id.append({
            'id': str(individual),
            'times': [tolist(ll._times[m]) for m in ll._obs_masks],
            'observations': [tolist(o) for o in ll._observations],
            'events': [list(e) for e in
                       ctrl._dosing_regimens[individual].events]})
    cov = ctrl._extract_covariates(cov_names)
    return {'columns': list(data.columns),
            'rows': [[None if (isinstance(v, float) and np.isnan(v)) else
                      (v if isinstance(v, str) else float(v)) for v in r]
                     for r in rows],
            'row_id_is_str': [isinstance(r[0], str) for r in rows],
            'output_observable_dict': {
                'central.drug_concentration': 'Conc',
                'global.tumour_volume': 'Volume'},
            'individuals': per_id, 'covariates': tolist(cov)}


def main():
    rng = np.random.default_rng(20240607)
    stats_golden = {
        'error_models': error_model_cases(rng),
        'population_models': population_cases(rng)}
    with open(os.path.join(HERE, 'stats_golden.json'), 'w') as f:
        json.dump(stats_golden, f)
    hier_golden = {
        'c1': c1_case(rng),
        'c2': hier_case(rng, 6, 'lognormal', False),
        'c5': hier_case(rng, 5, 'cam', True),
        'predictive_c3': predictive_case(rng, 'c3'),
        'predictive_c5': predictive_case(rng, 'c5')}
    with open(os.path.join(HERE, 'hier_golden.json'), 'w') as f:
        json.dump(hier_golden, f)
    rng2 = np.random.default_rng(20241020)
    with open(os.path.join(HERE, 'stats_golden_extra.json'), 'w') as f:
        json.dump({'truncated_gaussian': truncated_gaussian_cases(rng2),
                   'reduced_error_model': reduced_error_model_cases(rng2)}, f)
    rng3 = np.random.default_rng(20261020)
    round2 = {
        'linear_covariate': linear_covariate_cases(rng3),
        'pointwise_and_reduced': pointwise_and_reduced_cases(rng3),
        'c2_120': hier_case(rng3, 120, 'lognormal', False),
        'filters': filter_cases(rng3),
        'filter_posterior': [
            filter_posterior_case(rng3, False, True),
            filter_posterior_case(rng3, True, False)],
        'ingestion': ingestion_case(rng3)}
    with open(os.path.join(HERE, 'round2_golden.json'), 'w') as f:
        json.dump(round2, f)
    print('wrote', os.path.join(HERE, 'stats_golden.json'),
          os.path.join(HERE, 'hier_golden.json'),
          os.path.join(HERE, 'stats_golden_extra.json'))


if __name__ == '__main__':
    main()
