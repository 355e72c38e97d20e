"""The CUDA path against the reference's golden vectors and known answers.

``tests/golden/hier_golden.json`` holds outputs of the UNMODIFIED reference
log-pdf classes (driving the exact ODE solution); parameters are matched by
NAME because the order of the covariate coefficients in the reference is
platform dependent.  Integrator at rtol 1e-11: score
within 1e-8 relative, gradient within 1e-6.
"""
import json
import os

import numpy as np
import pytest

import chi_b200
from tests import kats

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

CLASSES = {
    'gaussian': chi_b200.GaussianErrorModel,
    'lognormal': chi_b200.LogNormalErrorModel,
    'cam': chi_b200.ConstantAndMultiplicativeGaussianErrorModel,
    'multiplicative': chi_b200.MultiplicativeGaussianErrorModel}


def _load(name):
    with open(os.path.join(HERE, 'golden', name)) as f:
        return json.load(f)


def test_error_model_known_answers_on_device():
    for kind, sigma, ybar, sens, obs, score, grad in kats.error_model_kats():
        em = CLASSES[kind]()
        ll, g = em.compute_sensitivities(sigma, ybar, sens, obs)
        assert ll == pytest.approx(score, rel=1e-12, abs=1e-12)
        assert em.compute_log_likelihood(sigma, ybar, obs) == pytest.approx(
            score, rel=1e-12)
        pw = em.compute_pointwise_ll(sigma, ybar, obs)
        assert pw.shape == (len(obs),)
        assert np.sum(pw) == pytest.approx(score, rel=1e-12)
        if grad is not None:
            np.testing.assert_allclose(g, grad, rtol=1e-12, atol=1e-12)


def test_error_models_match_reference_golden_on_device():
    for case in _load('stats_golden.json')['error_models']:
        em = CLASSES[case['kind']]()
        ll, g = em.compute_sensitivities(
            case['sigma'], case['ybar'], np.array(case['sens']), case['obs'])
        assert ll == pytest.approx(case['ll'], rel=1e-12)
        np.testing.assert_allclose(g, case['grad'], rtol=1e-11, atol=1e-13)


def _build_population(spec):
    models = []
    for b in spec:
        if b[0] == 'pooled':
            models.append(chi_b200.PooledModel(n_dim=b[1]))
        elif b[0] == 'hetero':
            models.append(chi_b200.HeterogeneousModel(n_dim=b[1]))
        elif b[0] == 'gaussian':
            models.append(chi_b200.GaussianModel(n_dim=b[1], centered=b[2]))
        elif b[0] == 'lognormal':
            models.append(chi_b200.LogNormalModel(n_dim=b[1], centered=b[2]))
        else:
            inner = _build_population([b[1]])[0]
            models.append(chi_b200.CovariatePopulationModel(
                inner, chi_b200.LinearCovariateModel(n_cov=b[2])))
    return models


def test_population_models_match_reference_golden_on_device():
    for case in _load('stats_golden.json')['population_models']:
        pm = chi_b200.ComposedPopulationModel(_build_population(case['spec']))
        pm.set_n_ids(case['n_ids'])
        names = pm.get_parameter_names()
        assert sorted(names) == sorted(case['names'])
        # map the reference's top-level vector onto this platform's order
        perm = [case['names'].index(n) for n in names]
        top = np.array(case['top'])[perm]
        cov = None if case['covariates'] is None else \
            np.array(case['covariates'])
        assert pm.n_hierarchical_parameters(case['n_ids']) == (
            case['n_bottom'], case['n_top'])
        assert pm.get_special_dims()[0] == case['special_dims']
        eta = pm.compute_individual_parameters(
            top, np.array(case['bottom']), cov, return_eta=True)
        np.testing.assert_allclose(eta, case['eta'], rtol=1e-14)
        psi = pm.compute_individual_parameters(top, eta, cov)
        np.testing.assert_allclose(psi, case['psi'], rtol=1e-13)
        assert pm.compute_log_likelihood(top, eta, covariates=cov) == \
            pytest.approx(case['score'], rel=1e-12)
        s, g = pm.compute_sensitivities(
            top, eta, dlogp_dpsi=np.array(case['dlogp_dpsi']), reduce=True,
            covariates=cov)
        assert s == pytest.approx(case['score_s1'], rel=1e-12)
        nb = case['n_bottom']
        ref = np.array(case['grad'])
        np.testing.assert_allclose(g[:nb], ref[:nb], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(
            g[nb:], ref[nb:][perm], rtol=1e-11,
            atol=1e-13 * max(1.0, np.max(np.abs(ref))))


def test_truncated_gaussian_matches_reference_golden_on_device():
    """TruncatedGaussianModel (chi/_population_models.py:3500-3939) against
    the reference's own outputs, stand-alone as the reference supports it."""
    for case in _load('stats_golden_extra.json')['truncated_gaussian']:
        d, n_ids = case['n_dim'], case['n_ids']
        pm = chi_b200.TruncatedGaussianModel(n_dim=d)
        assert pm.get_parameter_names() == case['names']
        theta = np.array(case['theta'])
        obs = np.array(case['obs'])
        dlp = np.array(case['dlogp_dpsi'])
        assert pm.compute_log_likelihood(theta, obs) == pytest.approx(
            case['ll'], rel=1e-12)
        s, g = pm.compute_sensitivities(
            theta, obs, dlogp_dpsi=dlp, reduce=True)
        assert s == pytest.approx(case['score_reduced'], rel=1e-12)
        np.testing.assert_allclose(
            g, case['grad_reduced'], rtol=1e-11,
            atol=1e-13 * np.max(np.abs(case['grad_reduced'])))
        neg = obs.copy()
        neg[0, 0] = -0.1
        assert pm.compute_log_likelihood(theta, neg) == -np.inf
        bad = theta.copy()
        bad[d] = 0.0
        assert pm.compute_log_likelihood(bad, obs) == -np.inf
        np.testing.assert_allclose(
            pm.get_mean_and_std(theta), case['mean_and_std'], rtol=1e-13)
        np.testing.assert_allclose(
            pm.sample(theta, n_samples=4, seed=3), case['samples_seed3'],
            rtol=1e-13)


def _prior(spec):
    kinds, locs, scales = spec
    priors = []
    for kind, m, s in zip(kinds, locs, scales):
        priors.append(chi_b200.GaussianLogPrior(m, s) if kind == 0
                      else chi_b200.LogNormalLogPrior(m, s))
    return priors


def test_c1_log_posterior_matches_reference_golden():
    """BASELINE.json configs[0]: one-compartment PK, single bolus, one
    individual, GaussianErrorModel, LogPosterior.evaluateS1 on 100 times."""
    case = _load('hier_golden.json')['c1']
    model = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    model.set_administration('central', direct=True)
    level, start, duration, _, _ = case['events'][0]
    model.set_dosing_regimen(
        dose=level * duration, start=start, duration=duration)
    model.set_tolerance(abs_tol=1e-13, rel_tol=1e-11)
    ll = chi_b200.LogLikelihood(
        model, chi_b200.GaussianErrorModel(), case['obs'], case['times'])
    assert ll.get_parameter_names() == case['names']
    post = chi_b200.LogPosterior(
        ll, chi_b200.ComposedLogPrior(*_prior(case['prior'])))
    score, grad = post.evaluateS1(np.array(case['x']))
    assert score == pytest.approx(case['score'], rel=1e-8)
    np.testing.assert_allclose(grad, case['grad'], rtol=1e-6)
    assert post(np.array(case['x'])) == pytest.approx(
        case['score_call'], rel=1e-8)


@pytest.mark.parametrize('name', ['c2', 'c5', 'c2_120'])
def test_hierarchical_posterior_matches_reference_golden(name):
    """C2 / C5 at 5-6 individuals and C2 at 120 individuals
    (round2_golden.json)."""
    case = _load('round2_golden.json')[name] if name == 'c2_120' \
        else _load('hier_golden.json')[name]
    n_ids = case['n_ids']
    model = chi_b200.library.ModelLibrary().two_compartment_pk_model()
    model.set_administration('central', direct=True)
    model.set_outputs([case['output']])
    model.set_tolerance(abs_tol=1e-13, rel_tol=1e-11)
    em = CLASSES[case['error']]()
    lls = []
    for i in range(n_ids):
        m = model.copy()
        m.set_tolerance(abs_tol=1e-13, rel_tol=1e-11)
        reg = chi_b200.DosingRegimen()
        for ev in case['events'][i]:
            reg.schedule(*ev)
        m.set_dosing_regimen(reg)
        ll = chi_b200.LogLikelihood(m, em, case['obs'][i], case['times'][i])
        ll.set_id('id %d' % i)
        lls.append(ll)
    assert lls[0].get_parameter_names() == case['ll_names']
    n_sig = len(case['ll_names']) - 7
    if case['covariates'] is not None:
        middle = chi_b200.CovariatePopulationModel(
            chi_b200.LogNormalModel(n_dim=5, centered=False),
            chi_b200.LinearCovariateModel(n_cov=2))
    else:
        middle = chi_b200.LogNormalModel(n_dim=5, centered=False)
    pm = chi_b200.ComposedPopulationModel(
        [chi_b200.PooledModel(2), middle, chi_b200.PooledModel(n_sig)])
    pm.set_dim_names(case['ll_names'])
    hll = chi_b200.HierarchicalLogLikelihood(
        lls, pm, covariates=case['covariates'])
    names = hll.get_parameter_names(include_ids=True)
    assert sorted(names) == sorted(case['names'])
    perm = [case['names'].index(n) for n in names]
    top_names = hll.get_parameter_names(exclude_bottom_level=True)
    tperm = [case['top_names'].index(n) for n in top_names]
    kinds, locs, scales = case['prior']
    prior = chi_b200.ComposedLogPrior(*_prior(
        [[kinds[k] for k in tperm], [locs[k] for k in tperm],
         [scales[k] for k in tperm]]))
    post = chi_b200.HierarchicalLogPosterior(hll, prior)
    x = np.array(case['x'])[perm]
    score, grad = post.evaluateS1(x)
    gref = np.array(case['grad'])[perm]
    assert score == pytest.approx(case['score'], rel=1e-8)
    np.testing.assert_allclose(
        grad, gref, rtol=1e-6, atol=1e-6 * np.max(np.abs(gref)))
    assert post(x) == pytest.approx(case['score_call'], rel=1e-8)
    assert post.n_ids() == n_ids
    assert post.get_id(unique=True) == ['id %d' % i for i in range(n_ids)]


@pytest.mark.parametrize('name', ['predictive_c3', 'predictive_c5'])
def test_population_predictive_sample_matches_reference_golden(name):
    """BASELINE.json configs[2] / [4] forward path at small size:
    PopulationPredictiveModel.sample(return_df=False), same seed as the
    reference.  The noise stream is reproduced exactly; the trajectories are
    integrated at rtol 1e-11, so samples agree to ~1e-9."""
    case = _load('hier_golden.json')[name]
    lib = chi_b200.library.ModelLibrary()
    model = getattr(lib, case['model'])()
    model.set_administration('central', direct=case['direct'])
    model.set_outputs([case['output']])
    reg = chi_b200.DosingRegimen()
    for ev in case['events']:
        reg.schedule(*ev)
    model.set_dosing_regimen(reg)
    model.set_tolerance(abs_tol=1e-13, rel_tol=1e-11)
    pred = chi_b200.PredictiveModel(model, [CLASSES[case['error']]()])
    pm = chi_b200.ComposedPopulationModel(_build_population(case['spec']))
    names = pm.get_parameter_names()
    assert sorted(names) == sorted(case['theta_names'])
    theta = np.array(case['theta'])[
        [case['theta_names'].index(n) for n in names]]
    ppm = chi_b200.PopulationPredictiveModel(pred, pm)
    kwargs = {}
    if case['covariates'] is not None:
        kwargs['covariates'] = np.array(case['covariates'])
    out = ppm.sample(theta, case['times'], n_samples=case['n_samples'],
                     seed=case['seed'], return_df=False, **kwargs)
    ref = np.array(case['samples'])
    assert out.shape == ref.shape
    # (values that decayed below the integrator's atol = 1e-13 are compared
    # absolutely)
    np.testing.assert_allclose(out, ref, rtol=1e-8, atol=1e-12)
    df = ppm.sample(theta, case['times'], n_samples=case['n_samples'],
                    seed=case['seed'], **kwargs)
    assert list(df.columns) == ['ID', 'Time', 'Observable', 'Value']
    values = df[df['Observable'] == case['output']]['Value'].to_numpy()
    np.testing.assert_allclose(values, out.flatten(), rtol=1e-15)


def test_predictive_model_sample_single_individual():
    """PredictiveModel.sample: one solve, seed-pinned noise
    (chi/_predictive_models.py:659-736)."""
    model = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    model.set_administration('central')
    pred = chi_b200.PredictiveModel(model, chi_b200.GaussianErrorModel())
    pred.set_dosing_regimen(dose=10, start=0, duration=0.01)
    assert pred.get_parameter_names()[-1] == 'Sigma'
    times = [3.0, 1.0, 2.0]
    out = pred.sample([0, 2, 1, 0.1], times, n_samples=4, seed=3,
                      return_df=False)
    assert out.shape == (1, 3, 4)
    ybar = pred.get_submodels()['Mechanistic model'].simulate(
        [0, 2, 1], np.sort(times))
    assert np.all(ybar > 0.1)
    noise = np.random.default_rng(3).normal(0, 0.1, size=(3, 4))
    np.testing.assert_allclose(out[0], ybar[0][:, None] + noise, rtol=1e-13)
    df = pred.sample([0, 2, 1, 0.1], times, n_samples=2, seed=3,
                     include_regimen=True)
    assert set(df.columns) >= {'ID', 'Time', 'Observable', 'Value'}
    reg = pred.get_dosing_regimen()
    assert list(reg['Dose']) == [10.0] and list(reg['Duration']) == [0.01]


def test_reduced_error_model_matches_reference_golden_on_device():
    """ReducedErrorModel (chi/_error_models.py:1583-1906): fixed parameters
    are scattered into the full vector and dropped from the gradient."""
    ctor = {'cam': chi_b200.ConstantAndMultiplicativeGaussianErrorModel,
            'gaussian': chi_b200.GaussianErrorModel}
    for case in _load('stats_golden_extra.json')['reduced_error_model']:
        rem = chi_b200.ReducedErrorModel(ctor[case['kind']]())
        if case['fixed']:
            rem.fix_parameters(case['fixed'])
        assert rem.get_parameter_names() == case['names']
        assert rem.n_parameters() == case['n_parameters']
        assert rem.n_fixed_parameters() == case['n_fixed']
        free = np.array(case['free'])
        ybar, obs = np.array(case['ybar']), np.array(case['obs'])
        sens = np.array(case['sens'])
        ll, grad = rem.compute_sensitivities(free, ybar, sens, obs)
        assert ll == pytest.approx(case['ll'], rel=1e-12)
        np.testing.assert_allclose(
            grad, case['grad'], rtol=1e-11,
            atol=1e-13 * np.max(np.abs(case['grad'])))
        assert rem.compute_log_likelihood(free, ybar, obs) == pytest.approx(
            case['ll_call'], rel=1e-12)
        np.testing.assert_allclose(
            rem.sample(free, ybar, n_samples=2, seed=5),
            case['sample_seed5'], rtol=1e-13)
        if case['fixed']:
            rem.fix_parameters({k: None for k in case['fixed']})
            assert rem.n_fixed_parameters() == 0
