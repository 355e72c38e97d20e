"""Round-2 additions against golden vectors of the UNMODIFIED reference
(tests/golden/round2_golden.json, written by tests/golden/make_golden.py):
stand-alone LinearCovariateModel, LogLikelihood.compute_pointwise_ll,
Reduced{Mechanistic,Error}Model inside a device log-likelihood, the
population-filter log-posterior, and the averaged predictive models."""
import json
import os

import numpy as np
import pytest

import chi_b200

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TIGHT = dict(abs_tol=1e-13, rel_tol=1e-11)


def _golden():
    with open(os.path.join(HERE, 'golden', 'round2_golden.json')) as f:
        return json.load(f)


def _nan(a):
    return np.array([[[np.nan if v is None else v for v in r] for r in o]
                     for o in a], dtype=float)


def test_linear_covariate_model_stand_alone():
    """chi/_covariate_models.py:257-320 through chi_b200_covariate_linear."""
    for case in _golden()['linear_covariate']:
        cm = chi_b200.LinearCovariateModel(n_cov=case['n_cov'])
        if case['selection'] is not None:
            cm.set_population_parameters(case['selection'])
        pidx, didx = cm.get_set_population_parameters()
        assert list(pidx) == [int(v) for v in case['pidx']]
        assert list(didx) == [int(v) for v in case['didx']]
        assert cm.get_parameter_names() == case['names']
        vartheta = cm.compute_population_parameters(
            case['beta'], np.array(case['pop']), np.array(case['covariates']))
        np.testing.assert_allclose(vartheta, case['vartheta'], rtol=1e-13)
        dpop, dbeta = cm.compute_sensitivities(
            case['beta'], np.array(case['pop']), np.array(case['covariates']),
            np.array(case['dlogp_dvartheta']))
        np.testing.assert_allclose(dpop, case['dpop'], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(
            dbeta, case['dbeta'], rtol=1e-12, atol=1e-14)


def _two_output_model(case):
    model = chi_b200.library.ModelLibrary().two_compartment_pk_model()
    model.set_administration('central', direct=True)
    model.set_outputs(case['outputs'])
    reg = chi_b200.DosingRegimen()
    for ev in case['events']:
        reg.schedule(*ev)
    model.set_dosing_regimen(reg)
    model.set_tolerance(**TIGHT)
    return model


def test_pointwise_log_likelihood_two_outputs():
    """LogLikelihood.compute_pointwise_ll, chi/_log_pdfs.py:932-971."""
    case = _golden()['pointwise_and_reduced']
    ll = chi_b200.LogLikelihood(
        _two_output_model(case),
        [chi_b200.LogNormalErrorModel(),
         chi_b200.ConstantAndMultiplicativeGaussianErrorModel()],
        case['obs'], case['times'])
    assert ll.get_parameter_names() == case['names']
    x = np.array(case['x'])
    pw = ll.compute_pointwise_ll(x)
    ref = np.array(case['pointwise'])
    assert pw.shape == ref.shape
    np.testing.assert_allclose(pw, ref, rtol=1e-7, atol=1e-9)
    score, grad = ll.evaluateS1(x)
    assert np.sum(pw) == pytest.approx(score, rel=1e-10)
    assert score == pytest.approx(case['score'], rel=1e-8)
    np.testing.assert_allclose(
        grad, case['grad'], rtol=1e-6,
        atol=1e-6 * np.max(np.abs(case['grad'])))
    with pytest.raises(NotImplementedError):
        # the reference does not implement it either
        # (chi/_log_pdfs.py:161-198)
        chi_b200.HierarchicalLogLikelihood.compute_pointwise_ll(None, x)


def test_reduced_models_inside_a_device_log_likelihood():
    """ReducedMechanisticModel (chi/_mechanistic_models.py:1160-1186) and
    ReducedErrorModel (chi/_error_models.py:1704-1750) as the models of a
    LogLikelihood: their fixed entries go to the device once
    (chi_b200_problem_set_fixed_parameters), sensitivities are integrated
    for the free parameters only."""
    case = _golden()['pointwise_and_reduced']
    red = case['reduced']
    rmm = chi_b200.ReducedMechanisticModel(_two_output_model(case))
    rmm.fix_parameters(red['fixed_mechanistic'])
    assert rmm.n_fixed_parameters() == 2 and rmm.n_parameters() == 5
    rem = chi_b200.ReducedErrorModel(
        chi_b200.ConstantAndMultiplicativeGaussianErrorModel())
    rem.fix_parameters(red['fixed_error'])
    ll = chi_b200.LogLikelihood(
        rmm, [chi_b200.LogNormalErrorModel(), rem], case['obs'],
        case['times'])
    assert ll.get_parameter_names() == red['names']
    assert ll.n_parameters() == len(red['x'])
    x = np.array(red['x'])
    score, grad = ll.evaluateS1(x)
    assert score == pytest.approx(red['score'], rel=1e-8)
    np.testing.assert_allclose(
        grad, red['grad'], rtol=1e-6, atol=1e-6 * np.max(np.abs(red['grad'])))
    assert ll(x) == pytest.approx(red['score_call'], rel=1e-8)
    # fixing one more parameter on the log-likelihood itself composes with
    # the wrappers' fixed entries, and the wrappers' entries stay locked
    ll.fix_parameters({'central.size': x[1], 'central.drug_amount': 7.0})
    assert 'central.size' not in ll.get_parameter_names()
    keep = [k for k, n in enumerate(red['names']) if n != 'central.size']
    s2, g2 = ll.evaluateS1(x[keep])
    assert s2 == pytest.approx(red['score'], rel=1e-8)
    np.testing.assert_allclose(
        g2, np.array(red['grad'])[keep], rtol=1e-6,
        atol=1e-6 * np.max(np.abs(red['grad'])))
    # the stand-alone simulate of the reduced model scatters the free values
    y_full = rmm.mechanistic_model().simulate(
        [0.1, 0.05, 2.1, 0.9, 5.2, 1.1, 9.5], case['times'][0])
    y_red = rmm.simulate([0.05, 2.1, 0.9, 5.2, 9.5], case['times'][0])
    np.testing.assert_allclose(y_red, y_full, rtol=1e-13)


def _prior(spec):
    kinds, locs, scales = spec
    return chi_b200.ComposedLogPrior(*[
        chi_b200.GaussianLogPrior(m, s) if k == 0
        else chi_b200.LogNormalLogPrior(m, s)
        for k, m, s in zip(kinds, locs, scales)])


@pytest.mark.parametrize('index', [0, 1])
def test_population_filter_log_posterior_matches_reference(index):
    """chi/_log_pdfs.py:1325-2075: one batched solve (with sensitivities for
    evaluateS1) of the virtual individuals + the filter's reductions."""
    case = _golden()['filter_posterior'][index]
    model = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    model.set_administration('central', direct=True)
    model.set_outputs([case['output']])
    level, start, duration, _, _ = case['events'][0]
    model.set_dosing_regimen(
        dose=level * duration, start=start, duration=duration)
    model.set_tolerance(**TIGHT)
    data = _nan(case['data'])
    filt = getattr(chi_b200, case['filter'])(data)
    pm = chi_b200.ComposedPopulationModel([
        chi_b200.PooledModel(1), chi_b200.LogNormalModel(1, centered=False),
        chi_b200.GaussianModel(1, centered=True)])
    post = chi_b200.PopulationFilterLogPosterior(
        filt, case['times'], model, pm, _prior(case['prior']),
        sigma=case['sigma'], error_on_log_scale=case['log_scale'],
        n_samples=case['n_samples'])
    assert post.n_parameters() == case['n_parameters']
    assert post.n_parameters(exclude_bottom_level=True) == case['n_top']
    assert post.get_parameter_names(include_ids=True) == case['names']
    x = np.array(case['x'])
    score, grad = post.evaluateS1(x)
    assert score == pytest.approx(case['score'], rel=1e-8)
    np.testing.assert_allclose(
        grad, case['grad'], rtol=1e-6,
        atol=1e-7 * np.max(np.abs(case['grad'])))
    assert post(x) == pytest.approx(case['score_call'], rel=1e-8)
    positive = chi_b200.ComposedLogPrior(*[
        chi_b200.LogNormalLogPrior(0.0, 0.1)] * case['n_top'])
    post2 = chi_b200.PopulationFilterLogPosterior(
        filt, case['times'], model, pm, positive, sigma=case['sigma'],
        error_on_log_scale=case['log_scale'], n_samples=case['n_samples'])
    init = post2.sample_initial_parameters(n_samples=2, seed=3)
    assert init.shape == (2, case['n_parameters'])
    assert np.all(np.isfinite(init))
    assert np.isfinite(post2(init[0]))


def test_posterior_and_prior_predictive_models():
    """chi/_predictive_models.py:130-344, 1121-1230: parameter sets drawn
    from posterior samples / a prior, all series in one batched launch."""
    model = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    model.set_administration('central', direct=True)
    model.set_outputs(['central.drug_concentration'])
    pred = chi_b200.PredictiveModel(model, [chi_b200.GaussianErrorModel()])
    pred.set_dosing_regimen(dose=10, start=0, duration=0.5)
    names = pred.get_parameter_names()
    rng = np.random.default_rng(0)
    truth = {'central.drug_amount': 0.0, 'central.size': 2.0,
             'global.elimination_rate': 1.0, 'Sigma': 1e-3}
    posterior = {}
    for n in names:
        draws = truth[n] * (1 + 0.01 * rng.normal(size=(2, 50, 3)))
        draws[1, 40:, :] = np.nan          # chains of unequal length
        posterior[n] = draws
    posterior['individual'] = ['a', 'b', 'c']
    ppm = chi_b200.PosteriorPredictiveModel(pred, posterior)
    times = np.array([3.0, 0.5, 1.0, 2.0])
    df = ppm.sample(times, n_samples=40, individual='b', seed=1)
    assert list(df.columns) == ['ID', 'Time', 'Observable', 'Value']
    assert len(df) == 40 * 4 and set(df['ID']) == set(range(1, 41))
    arr = ppm.sample(times, n_samples=40, individual='b', seed=1,
                     return_df=False)
    assert arr.shape == (1, 4, 40)
    exact = pred._mechanistic_model.simulate(
        [0.0, 2.0, 1.0], np.sort(times))
    np.testing.assert_allclose(
        np.mean(arr[0], axis=1), exact[0], rtol=0.02)
    with pytest.raises(ValueError):
        ppm.sample(times, individual='z')
    with pytest.raises(ValueError):
        chi_b200.PosteriorPredictiveModel(pred, {'central.size': np.ones(
            (1, 2))})
    prior = chi_b200.ComposedLogPrior(
        chi_b200.UniformLogPrior(0, 1e-9), chi_b200.LogNormalLogPrior(
            np.log(2), 0.01), chi_b200.LogNormalLogPrior(0, 0.01),
        chi_b200.UniformLogPrior(1e-4, 2e-4))
    prm = chi_b200.PriorPredictiveModel(pred, prior)
    arr = prm.sample(times, n_samples=30, seed=2, return_df=False)
    assert arr.shape == (1, 4, 30)
    np.testing.assert_allclose(np.mean(arr[0], axis=1), exact[0], rtol=0.02)
    # population-level predictive model: one virtual individual per draw
    pm = chi_b200.ComposedPopulationModel([
        chi_b200.PooledModel(1), chi_b200.LogNormalModel(2),
        chi_b200.PooledModel(1)])
    popm = chi_b200.PopulationPredictiveModel(pred, pm)
    pnames = popm.get_parameter_names()
    ptruth = [0.0, np.log(2.0), np.log(1.0), 0.01, 0.01, 1e-3]
    pposterior = {n: v * np.ones((2, 10)) for n, v in zip(pnames, ptruth)}
    ppp = chi_b200.PosteriorPredictiveModel(popm, pposterior)
    arr = ppp.sample(times, n_samples=25, seed=4, return_df=False)
    assert arr.shape == (1, 4, 25)
    np.testing.assert_allclose(np.mean(arr[0], axis=1), exact[0], rtol=0.05)


@pytest.mark.gpu
def test_small_launches_one_warp_per_system_against_the_oracle():
    """Single parameter vectors (what pints asks for) run one warp per system
    with the sensitivity columns side by side on its lanes
    (chi_b200_problem_set_small_batch): log-likelihood, gradient, the -inf
    convention and the forward-only call against the oracle (exact solution +
    numpy error / population models) and against the thread-per-system
    kernel, at the default tolerances and for two error models."""
    from chi_b200 import _synthetic
    from tests.test_gpu_bench_tolerance import oracle_scores
    for error in ('lognormal', 'cam'):
        prob = _synthetic.two_compartment_problem(
            60, seed=21, n_doses=3, error=error,
            rtol=_synthetic.BENCH_RTOL, atol=_synthetic.BENCH_ATOL)
        post, ll = prob['log_posterior'], prob['log_likelihood']
        rng = np.random.default_rng(3)
        x = prob['x0'] * (1 + 0.05 * rng.normal(size=len(prob['x0'])))
        ref, gref = oracle_scores(prob, x[None, :], error, True)
        ll.set_small_batch(-1)
        s, g = post.evaluateS1(x)
        f = post(x)
        ll.set_small_batch(0)
        s_t, g_t = post.evaluateS1(x)
        f_t = post(x)
        ll.set_small_batch(-1)
        assert abs(s - ref[0]) <= 1e-8 * abs(ref[0])
        assert np.max(np.abs(g - gref[0])) <= 1e-6 * np.max(np.abs(gref[0]))
        assert abs(f - ref[0]) <= 1e-8 * abs(ref[0])
        assert s == pytest.approx(s_t, rel=3e-9)
        assert f == pytest.approx(f_t, rel=3e-9)
        np.testing.assert_allclose(g, g_t, rtol=1e-6,
                                   atol=1e-7 * np.max(np.abs(g_t)))
        # invalid sigma: -inf / inf from every lane of the warp
        xb = x.copy()
        xb[-1] = -0.1
        with np.errstate(all='ignore'):
            sb, gb = post.evaluateS1(xb)
        assert np.isneginf(sb)
