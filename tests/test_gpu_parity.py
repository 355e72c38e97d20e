"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star):
  * trajectories and sensitivities: rtol 1e-6 (scaled by the column maximum)
    at the default integrator tolerances, and ~1e-9 at rtol 1e-11;
  * summed log-pdf: 1e-8 relative, gradient 1e-6 relative, both with the
    integrator at rtol 1e-11 / atol 1e-13;
  * statistical stages fed identical inputs: 1e-12.
"""
import warnings

import numpy as np
import pytest

import chi_b200
from oracle import ode_exact, stats

pytestmark = pytest.mark.gpu

TIGHT = dict(abs_tol=1e-13, rel_tol=1e-11)


def _events(model):
    reg = model.dosing_regimen()
    if reg is None:
        return []
    return [tuple(row) for row in reg.as_array()]


def _scaled_err(a, b, axis=0):
    scale = np.max(np.abs(b), axis=axis, keepdims=True)
    scale[scale == 0] = 1.0
    return np.max(np.abs(a - b) / scale)


CASES = {
    # name: (library ctor, oracle ctor, direct, output, psi, regimen kwargs)
    'one_cmt_direct_bolus': (
        'one_compartment_pk_model', ode_exact.one_compartment, True,
        'central.drug_concentration', [0.0, 2.0, 1.0],
        dict(dose=10, start=0, duration=0.01)),
    'one_cmt_indirect': (
        'one_compartment_pk_model', ode_exact.one_compartment, False,
        'central.drug_concentration', [0.3, 0.1, 2.0, 3.0, 1.0],
        dict(dose=4, start=0.5, duration=0.01, period=1.5, num=3)),
    'two_cmt_direct_infusion': (
        'two_compartment_pk_model', ode_exact.two_compartment, True,
        'central.drug_concentration', [0.0, 0.0, 2.0, 1.0, 5.0, 1.0, 10.0],
        dict(dose=10, start=0, duration=1.0, period=8, num=3)),
    'two_cmt_indirect_indefinite': (
        'two_compartment_pk_model', ode_exact.two_compartment, False,
        'central.drug_concentration',
        [0.1, 0.2, 0.3, 2.0, 10.0, 1.0, 5.0, 1.0, 10.0],
        dict(dose=2, period=1)),
    'tgi_direct_daily': (
        'erlotinib_tumour_growth_inhibition_model',
        ode_exact.tumour_growth_inhibition_pkpd, True,
        'global.tumour_volume', [0.0, 0.2, 2.0, 1.0, 0.5, 0.3, 0.1],
        dict(dose=5, start=0, duration=0.01, period=1, num=30)),
    'tgi_indirect_daily': (
        'erlotinib_tumour_growth_inhibition_model',
        ode_exact.tumour_growth_inhibition_pkpd, False,
        'global.tumour_volume',
        [0.0, 0.0, 0.2, 2.0, 3.0, 1.0, 0.5, 0.05, 0.1],
        dict(dose=5, start=0, duration=0.01, period=1, num=30)),
}

TIMES = {
    'one_cmt_direct_bolus': np.linspace(0.1, 10, 100),
    'one_cmt_indirect': np.array([0.0, 0.25, 0.5, 0.505, 1.0, 2.0, 2.7, 6.0]),
    'two_cmt_direct_infusion': np.sort(
        np.random.default_rng(1).uniform(0.25, 24, 10)),
    'two_cmt_indirect_indefinite': np.array([0.5, 1, 1.5, 2, 2.5, 3]),
    'tgi_direct_daily': np.arange(0.0, 31.0),
    'tgi_indirect_daily': np.arange(0.0, 31.0),
}


def _build(name):
    lib_name, oracle_ctor, direct, output, psi, reg = CASES[name]
    model = getattr(chi_b200.library.ModelLibrary(), lib_name)()
    model.set_administration('central', direct=direct)
    model.set_outputs([output])
    model.set_dosing_regimen(**reg)
    omodel = oracle_ctor(direct)
    omodel.set_outputs([output])
    assert model.parameters() == omodel.parameter_names
    return model, omodel, np.array(psi, float), TIMES[name]


@pytest.mark.parametrize('name', sorted(CASES))
def test_simulate_matches_exact_solution(name):
    model, omodel, psi, times = _build(name)
    y_ref, s_ref = ode_exact.simulate(omodel, psi, times, _events(model))

    # forward only, default tolerances: 1e-6 contract
    y = model.simulate(parameters=psi, times=times)
    assert y.shape == (1, len(times))
    assert _scaled_err(y, y_ref, axis=1) < 1e-6

    # with sensitivities, default tolerances: 1e-6 contract
    model.enable_sensitivities(True)
    y, s = model.simulate(parameters=psi, times=times)
    assert s.shape == (len(times), 1, len(psi))
    assert _scaled_err(y, y_ref, axis=1) < 1e-6
    assert _scaled_err(s[:, 0, :], s_ref[:, 0, :]) < 1e-6

    # tight tolerances: converges to the exact solution
    model.set_tolerance(**TIGHT)
    y, s = model.simulate(parameters=psi, times=times)
    assert _scaled_err(y, y_ref, axis=1) < 2e-10
    assert _scaled_err(s[:, 0, :], s_ref[:, 0, :]) < 2e-9


def test_simulate_multiple_outputs_and_subset_sensitivities():
    model, omodel, psi, times = _build('two_cmt_direct_infusion')
    outs = ['central.drug_amount', 'peripheral.drug_concentration',
            'central.drug_concentration']
    model.set_outputs(outs)
    omodel.set_outputs(outs)
    model.set_tolerance(**TIGHT)
    names = ['global.elimination_rate', 'peripheral.size',
             'central.drug_amount']
    model.enable_sensitivities(True, names)
    y, s = model.simulate(parameters=psi, times=times)
    y_ref, s_ref = ode_exact.simulate(omodel, psi, times, _events(model))
    cols = [model.parameters().index(n) for n in sorted(
        names, key=model.parameters().index)]
    assert y.shape == (3, len(times)) and s.shape == (len(times), 3, 3)
    assert _scaled_err(y, y_ref, axis=1) < 2e-10
    for o in range(3):
        assert _scaled_err(s[:, o, :], s_ref[:, o, cols]) < 2e-9


def test_simulate_no_dose_and_koch_models():
    lib = chi_b200.library.ModelLibrary()
    model = lib.tumour_growth_inhibition_model_koch()
    assert model.parameters() == [
        'global.tumour_volume', 'global.drug_concentration', 'global.kappa',
        'global.lambda_0', 'global.lambda_1']
    omodel = ode_exact.tumour_growth_koch()
    psi = np.array([0.3, 1.2, 0.2, 0.4, 0.7])
    times = np.linspace(0, 20, 15)
    model.set_tolerance(**TIGHT)
    model.enable_sensitivities(True)
    y, s = model.simulate(psi, times)
    y_ref, s_ref = ode_exact.simulate(omodel, psi, times, [])
    assert _scaled_err(y, y_ref, axis=1) < 2e-10
    assert _scaled_err(s[:, 0, :], s_ref[:, 0, :]) < 2e-9


ERR = [
    ('gaussian', chi_b200.GaussianErrorModel, [0.3]),
    ('lognormal', chi_b200.LogNormalErrorModel, [0.2]),
    ('cam', chi_b200.ConstantAndMultiplicativeGaussianErrorModel, [0.1, 0.2]),
    ('multiplicative', chi_b200.MultiplicativeGaussianErrorModel, [0.25]),
]


@pytest.mark.parametrize('kind,cls,sigma', ERR)
def test_error_models_match_oracle(kind, cls, sigma):
    rng = np.random.default_rng(3)
    em = cls()
    n_sig, f_ll, f_s1 = stats.ERROR_MODELS[kind]
    for n_obs, n_par in [(1, 1), (7, 3), (100, 9), (1000, 4)]:
        ybar = rng.uniform(0.5, 3.0, n_obs)
        obs = rng.uniform(0.5, 3.0, n_obs)
        sens = rng.normal(size=(n_obs, n_par))
        ll, grad = em.compute_sensitivities(sigma, ybar, sens, obs)
        ll_ref, grad_ref = f_s1(sigma, ybar, sens, obs)
        assert ll == pytest.approx(ll_ref, rel=1e-12)
        np.testing.assert_allclose(grad, grad_ref, rtol=1e-11, atol=1e-12)
        assert em.compute_log_likelihood(sigma, ybar, obs) == pytest.approx(
            f_ll(sigma, ybar, obs), rel=1e-12)
        np.testing.assert_allclose(
            em.compute_pointwise_ll(sigma, ybar, obs),
            stats.error_pointwise_ll(kind, sigma, ybar, obs), rtol=1e-12)
    # sigma <= 0 -> -inf / inf (chi/_error_models.py:238-240, 297-300, ...)
    bad = [-s for s in sigma]
    ll, grad = em.compute_sensitivities(bad, [1.0, 2.0], np.ones((2, 2)),
                                        [1.0, 2.0])
    assert ll == -np.inf and np.all(np.isinf(grad)) and np.all(grad > 0)
    assert em.compute_log_likelihood(bad, [1.0], [1.0]) == -np.inf


def test_lognormal_error_model_nonpositive_output():
    em = chi_b200.LogNormalErrorModel()
    ll, grad = em.compute_sensitivities(
        [0.2], [1.0, -0.5], np.ones((2, 2)), [1.0, 2.0])
    assert ll == -np.inf and np.all(grad == np.inf)


def _oracle_individual(omodel, kind, times, obs, events, n_mech):
    def simulate(psi, with_sens):
        return ode_exact.simulate(
            omodel, psi, times, events, sensitivities=with_sens)

    masks = np.ones((1, len(times)), bool)

    def s1(x):
        return stats.loglikelihood_s1(
            simulate, n_mech, [kind], [obs], masks, x, True)

    def ll(x):
        return stats.loglikelihood_s1(
            simulate, n_mech, [kind], [obs], masks, x, False)
    return ll, s1


@pytest.mark.parametrize('kind,cls,sigma', ERR)
def test_log_likelihood_matches_oracle(kind, cls, sigma):
    """config C1-like: LogLikelihood.__call__ / evaluateS1."""
    model, omodel, psi, times = _build('one_cmt_direct_bolus')
    model.set_tolerance(**TIGHT)
    rng = np.random.default_rng(1)
    y_ref, _ = ode_exact.simulate(omodel, psi, times, _events(model))
    obs = np.abs(y_ref[0] * (1 + 0.1 * rng.normal(size=len(times)))) + 1e-3
    ll = chi_b200.LogLikelihood(model, cls(), obs, times)
    assert ll.get_parameter_names()[:3] == model.parameters()
    x = np.array(list(psi) + list(sigma))
    x[0] = 0.5
    f_ll, f_s1 = _oracle_individual(
        omodel, kind, times, obs, _events(model), len(psi))
    score, grad = ll.evaluateS1(x)
    score_ref, grad_ref = f_s1(x)
    assert score == pytest.approx(score_ref, rel=1e-8)
    np.testing.assert_allclose(
        grad, grad_ref, rtol=1e-6, atol=1e-6 * np.max(np.abs(grad_ref)))
    assert ll(x) == pytest.approx(f_ll(x), rel=1e-8)


def test_log_likelihood_two_outputs_ragged_times():
    """Two outputs on different grids (chi/tests/test_log_pdfs.py:1674-1697
    style fixture): union grid + masks."""
    model, omodel, psi, _ = _build('one_cmt_indirect')
    outs = ['central.drug_amount', 'dose.drug_amount']
    model.set_outputs(outs)
    omodel.set_outputs(outs)
    model.set_tolerance(**TIGHT)
    t0 = np.array([0.5, 1.0, 2.0, 4.0])
    t1 = np.array([1.0, 3.0, 4.0])
    obs0 = np.array([0.4, 0.9, 1.1, 0.6])
    obs1 = np.array([2.0, 1.0, 0.7])
    ll = chi_b200.LogLikelihood(
        model, [chi_b200.GaussianErrorModel(),
                chi_b200.LogNormalErrorModel()], [obs0, obs1], [t0, t1])
    names = ll.get_parameter_names()
    assert names[-2:] == ['central.drug_amount Sigma',
                          'dose.drug_amount Sigma log']
    x = np.array(list(psi) + [0.3, 0.4])
    union, masks = stats.union_times([t0, t1])

    def simulate(p, with_sens):
        return ode_exact.simulate(
            omodel, p, union, _events(model), sensitivities=with_sens)
    ref, gref = stats.loglikelihood_s1(
        simulate, len(psi), ['gaussian', 'lognormal'], [obs0, obs1], masks,
        x, True)
    score, grad = ll.evaluateS1(x)
    assert score == pytest.approx(ref, rel=1e-8)
    np.testing.assert_allclose(
        grad, gref, rtol=1e-6, atol=1e-6 * np.max(np.abs(gref)))


def test_log_likelihood_invalid_sigma_and_failure_convention():
    model, omodel, psi, times = _build('one_cmt_direct_bolus')
    obs = np.ones(len(times))
    ll = chi_b200.LogLikelihood(
        model, chi_b200.GaussianErrorModel(), obs, times)
    score, grad = ll.evaluateS1(list(psi) + [-1.0])
    assert score == -np.inf and np.all(grad == np.inf)
    # integrator failure -> RuntimeWarning, -inf, +inf gradient
    model.set_tolerance(abs_tol=1e-13, rel_tol=1e-11, max_steps=5)
    ll = chi_b200.LogLikelihood(
        model, chi_b200.GaussianErrorModel(), obs, times)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        score, grad = ll.evaluateS1(list(psi) + [0.1])
    assert score == -np.inf and np.all(grad == np.inf)
    assert any(issubclass(x.category, RuntimeWarning) for x in w)


def _pop_cases(n_ids):
    G, L, P, H = (chi_b200.GaussianModel, chi_b200.LogNormalModel,
                  chi_b200.PooledModel, chi_b200.HeterogeneousModel)
    T = chi_b200.TruncatedGaussianModel
    C = chi_b200.CovariatePopulationModel
    Lin = chi_b200.LinearCovariateModel
    return [
        ([P(2), L(5, centered=False), P(1)],
         [('pooled', 2), ('lognormal', 5, False), ('pooled', 1)]),
        ([L(3, centered=True)], [('lognormal', 3, True)]),
        ([G(2, centered=True), P(1), G(2, centered=False)],
         [('gaussian', 2, True), ('pooled', 1), ('gaussian', 2, False)]),
        ([H(1, n_ids=n_ids), L(2, centered=False), P(2)],
         [('hetero', 1), ('lognormal', 2, False), ('pooled', 2)]),
        ([P(2), C(L(5, centered=False), Lin(n_cov=2)), P(2)],
         [('pooled', 2), ('covariate', ('lognormal', 5, False), 2),
          ('pooled', 2)]),
        ([C(G(2, centered=True), Lin(n_cov=1)), L(1, centered=True)],
         [('covariate', ('gaussian', 2, True), 1), ('lognormal', 1, True)]),
        ([C(L(2, centered=True), Lin(n_cov=3)), H(2, n_ids=n_ids)],
         [('covariate', ('lognormal', 2, True), 3), ('hetero', 2)]),
        # TruncatedGaussianModel inside a composed model (the reference only
        # supports it stand-alone; here it is a centered block like any other)
        ([P(1), T(3), L(1, centered=False)],
         [('pooled', 1), ('truncgauss', 3, True), ('lognormal', 1, False)]),
        ([T(2)], [('truncgauss', 2, True)]),
    ]


def _oracle_layout(models, spec, n_ids):
    blocks = []
    for m, b in zip(models, spec):
        if b[0] == 'covariate':
            pidx, didx = m._covariate_model.get_set_population_parameters()
            blocks.append(('covariate', b[1], b[2], pidx, didx))
        else:
            blocks.append(b)
    return stats.PopLayout(blocks, n_ids)


@pytest.mark.parametrize('n_ids', [1, 7, 300])
def test_population_models_match_oracle(n_ids):
    rng = np.random.default_rng(n_ids)
    for models, spec in _pop_cases(n_ids):
        pm = chi_b200.ComposedPopulationModel(models)
        pm.set_n_ids(n_ids)
        lay = _oracle_layout(models, spec, n_ids)
        n_b, n_t = pm.n_hierarchical_parameters(n_ids)
        assert (n_b, n_t) == (lay.n_bottom, lay.n_top)
        n_cov = pm.n_covariates()
        cov = rng.normal(size=(n_ids, n_cov)) if n_cov else None
        top = rng.uniform(0.2, 1.5, n_t)
        for i, name in enumerate(pm.get_parameter_names()):
            if 'Cov.' in name:
                top[i] *= 0.05
        bottom = rng.uniform(0.3, 2.0, n_b)
        eta = pm.compute_individual_parameters(
            top, bottom, cov, return_eta=True)
        eta_ref = stats.pop_individual_parameters(
            lay, top, bottom, cov, return_eta=True)
        np.testing.assert_allclose(eta, eta_ref, rtol=1e-14)
        psi = pm.compute_individual_parameters(top, eta, cov)
        psi_ref = stats.pop_individual_parameters(lay, top, eta_ref, cov)
        np.testing.assert_allclose(psi, psi_ref, rtol=1e-13)
        score = pm.compute_log_likelihood(top, eta, covariates=cov)
        assert score == pytest.approx(
            stats.pop_log_likelihood(lay, top, eta_ref, cov), rel=1e-12)
        dlp = rng.normal(size=(n_ids, lay.n_dim))
        score, grad = pm.compute_sensitivities(
            top, eta, dlogp_dpsi=dlp, reduce=True, covariates=cov)
        s_ref, g_ref = stats.pop_sensitivities_reduced(
            lay, top, eta_ref, dlp, cov)
        assert score == pytest.approx(s_ref, rel=1e-12)
        np.testing.assert_allclose(
            grad, g_ref, rtol=1e-11, atol=1e-12 * np.max(np.abs(g_ref)))


def test_population_negative_sigma_is_minus_inf():
    pm = chi_b200.ComposedPopulationModel(
        [chi_b200.LogNormalModel(2, centered=False), chi_b200.PooledModel(1)])
    pm.set_n_ids(3)
    top = np.array([0.1, 0.2, 0.3, -0.4, 1.0])
    bottom = np.linspace(0.1, 0.6, 6)
    eta = pm.compute_individual_parameters(top, bottom, return_eta=True)
    score, _ = pm.compute_sensitivities(
        top, eta, dlogp_dpsi=np.ones((3, 3)), reduce=True)
    assert score == -np.inf


def _hier_oracle(prob, n_ids, kind):
    ll = prob['log_likelihood']
    pm = prob['population_model']
    models = pm.get_population_models()
    spec = []
    for m in models:
        if isinstance(m, chi_b200.PooledModel):
            spec.append(('pooled', m.n_dim()))
        elif isinstance(m, chi_b200.CovariatePopulationModel):
            spec.append(('covariate', ('lognormal', 5, False), 2))
        else:
            spec.append(('lognormal', m.n_dim(), False))
    lay = _oracle_layout(models, spec, n_ids)
    omodel = ode_exact.two_compartment(True)
    omodel.set_outputs(['central.drug_concentration'])
    fns = []
    for i in range(n_ids):
        ev = [tuple(r) for r in prob['regimens'][i].as_array()]
        fns.append(_oracle_individual(
            omodel, kind, prob['times'][i], prob['observations'][i], ev, 7))
    return lay, fns


@pytest.mark.parametrize('error,covariates', [
    ('lognormal', False), ('cam', True), ('gaussian', False)])
def test_hierarchical_log_posterior_matches_oracle(error, covariates):
    """configs C2 / C5 at a size the oracle finishes in seconds."""
    from chi_b200 import _synthetic
    n_ids = 12
    prob = _synthetic.two_compartment_problem(
        n_ids, seed=5, n_doses=3, error=error, covariates=covariates,
        rtol=1e-11, atol=1e-13)
    post = prob['log_posterior']
    lay, fns = _hier_oracle(prob, n_ids, error)
    rng = np.random.default_rng(0)
    x = prob['x0'] * (1 + 0.02 * rng.normal(size=len(prob['x0'])))
    assert len(x) == post.n_parameters() == lay.n_parameters
    ref, gref = stats.posterior_s1(
        lay, lambda i, p: fns[i][1](p), prob['log_prior'], x,
        prob['covariates'])
    score, grad = post.evaluateS1(x)
    assert score == pytest.approx(ref, rel=1e-8)
    np.testing.assert_allclose(
        grad, gref, rtol=1e-6, atol=1e-6 * np.max(np.abs(gref)))
    ref0 = stats.posterior_call(
        lay, lambda i, p: fns[i][0](p), prob['log_prior'], x,
        prob['covariates'])
    score0 = post(x)
    assert score0 == pytest.approx(ref0, rel=1e-8)
    # batched entry points agree with the per-vector calls (the forward-only
    # and the sensitivity kernels take different steps, so they agree with
    # one another only to the integrator tolerance)
    X = np.vstack([x, prob['x0'], x * 1.01])
    scores = post.evaluate_batch(X)
    assert scores[0] == pytest.approx(score0, rel=1e-13)
    s1, g1 = post.evaluateS1_batch(X)
    assert s1[0] == pytest.approx(score, rel=1e-13)
    np.testing.assert_allclose(g1[0], grad, rtol=1e-13)
    np.testing.assert_allclose(s1, scores, rtol=1e-9)


def test_hierarchical_parameter_layout():
    """chi/tests/test_log_pdfs.py:740-767 and SURVEY.md section 8a."""
    from chi_b200 import _synthetic
    prob = _synthetic.two_compartment_problem(3, seed=1)
    ll = prob['log_likelihood']
    names = ll.get_parameter_names()
    assert len(names) == 28 and ll.n_parameters() == 28
    assert names[0:5] == [
        'central.size', 'global.elimination_rate',
        'global.transition_rate_c2p', 'global.transition_rate_p2c',
        'peripheral.size']
    assert names[15:17] == ['Pooled central.drug_amount',
                            'Pooled peripheral.drug_amount']
    assert names[27] == 'Pooled Sigma log'
    sd, n_pooled, n_het = prob['population_model'].get_special_dims()
    assert sd == [[0, 2, 0, 2, True], [7, 8, 12, 13, True]]
    assert (n_pooled, n_het) == (3, 0)
    ids = ll.get_id()
    assert ids[0] == ids[4] and ids[5] != ids[4] and ids[-1] is None


def test_gpu_kernel_matches_cpu_port():
    """The host port (oracle/csolver) runs the same algorithm; the two
    must agree to rounding."""
    from oracle import cport
    model, omodel, psi, times = _build('two_cmt_direct_infusion')
    model.enable_sensitivities(True)
    y, s = model.simulate(psi, times)
    key = cport.find_key('two_compartment_pk_model', 'central', True)
    assert key == model.model_key()
    pop = cport.Population(
        key, ['central.drug_concentration'], None, [times], None, None,
        [_events(model)])
    # same routine as the default kernel variant of a linear model: ROSL(11)
    yc, sc, st, _, _ = pop.simulate(psi, rtol=1e-8, atol=1e-10, stream_cb=11)
    assert st[0] == 0
    np.testing.assert_allclose(y[0], yc, rtol=1e-9)
    np.testing.assert_allclose(
        s[:, 0, :], sc, rtol=1e-7, atol=1e-9 * np.max(np.abs(sc)))


def test_all_kernel_variants_agree_with_exact_solution():
    """Every compiled sensitivity-kernel variant (ROSL(11) / ROSL(9) for
    linear models, RODAS5 / RODAS4) integrates to the same exact solution."""
    model, omodel, psi, times = _build('two_cmt_direct_infusion')
    y_ref, s_ref = ode_exact.simulate(omodel, psi, times, _events(model))
    model.set_tolerance(**TIGHT)
    model.enable_sensitivities(True)
    variants = model.model_code().variants()
    assert variants[0] == (0, 11) and (0, 5) in variants and (0, 1) in variants
    for v in range(len(variants)):
        model.set_kernel_variant(v)
        y, s = model.simulate(psi, times)
        assert _scaled_err(y, y_ref, axis=1) < 2e-10, variants[v]
        assert _scaled_err(s[:, 0, :], s_ref[:, 0, :]) < 2e-9, variants[v]
    # and on a nonlinear model with an indirect route (N = 3)
    model, omodel, psi, times = _build('tgi_indirect_daily')
    y_ref, s_ref = ode_exact.simulate(omodel, psi, times, _events(model))
    model.set_tolerance(**TIGHT)
    model.enable_sensitivities(True)
    for v in range(len(model.model_code().variants())):
        model.set_kernel_variant(v)
        y, s = model.simulate(psi, times)
        assert _scaled_err(y, y_ref, axis=1) < 2e-10
        assert _scaled_err(s[:, 0, :], s_ref[:, 0, :]) < 2e-9


def test_user_defined_model_is_compiled_on_first_use(tmp_path):
    """A model that is not built in (Michaelis-Menten elimination, written
    as SBML here) goes through the SBML reader, the code generator and the
    run-time module build, and matches the converged solution."""
    sbml = """<?xml version="1.0" encoding="UTF-8"?>
<sbml xmlns="http://www.sbml.org/sbml/level3/version2/core" level="3" version="2">
  <model id="mm_pk" timeUnits="day">
    <listOfCompartments>
      <compartment id="central" name="central" size="1"/>
    </listOfCompartments>
    <listOfSpecies>
      <species id="drug" name="drug" compartment="central" initialAmount="0"
               hasOnlySubstanceUnits="false"/>
    </listOfSpecies>
    <listOfParameters>
      <parameter id="v_max" value="1" constant="true"/>
      <parameter id="k_m" value="1" constant="true"/>
    </listOfParameters>
    <listOfReactions>
      <reaction id="elimination" reversible="false">
        <listOfReactants><speciesReference species="drug"/></listOfReactants>
        <kineticLaw>
          <math xmlns="http://www.w3.org/1998/Math/MathML">
            <apply><divide/>
              <apply><times/><ci> central </ci><ci> v_max </ci><ci> drug </ci></apply>
              <apply><plus/><ci> k_m </ci><ci> drug </ci></apply>
            </apply>
          </math>
        </kineticLaw>
      </reaction>
    </listOfReactions>
  </model>
</sbml>
"""
    path = tmp_path / 'mm_pk.xml'
    path.write_text(sbml)
    model = chi_b200.PKPDModel(str(path))
    assert model.parameters() == [
        'central.drug_amount', 'central.size', 'global.k_m', 'global.v_max']
    model.set_administration('central', direct=False)
    assert model.parameters() == [
        'central.drug_amount', 'dose.drug_amount', 'central.size',
        'dose.absorption_rate', 'global.k_m', 'global.v_max']
    model.set_outputs(['central.drug_concentration'])
    model.set_dosing_regimen(dose=8, start=0.25, duration=0.5, period=2, num=3)
    model.set_tolerance(**TIGHT)
    model.enable_sensitivities(True)
    psi = np.array([0.5, 0.1, 2.0, 3.0, 1.5, 2.5])
    times = np.array([0.0, 0.3, 0.75, 1.0, 2.5, 4.0, 6.0, 9.0])
    y, s = model.simulate(psi, times)

    cc = '(central__drug_amount / central__size)'
    omodel = ode_exact.OdeModel(
        ['central.drug_amount', 'dose.drug_amount'],
        ['central.size', 'global.v_max', 'global.k_m', 'dose.absorption_rate'],
        {'central.drug_amount':
            '-central__size * global__v_max * %s / (global__k_m + %s)'
            ' + dose__absorption_rate * dose__drug_amount' % (cc, cc),
         'dose.drug_amount': '-dose__absorption_rate * dose__drug_amount'},
        {'central.drug_concentration': cc}, 'dose.drug_amount')
    omodel.set_outputs(['central.drug_concentration'])
    assert omodel.parameter_names == model.parameters()
    y_ref, s_ref = ode_exact.simulate(omodel, psi, times, _events(model))
    assert _scaled_err(y, y_ref, axis=1) < 2e-10
    assert _scaled_err(s[:, 0, :], s_ref[:, 0, :]) < 2e-9


def test_user_defined_linear_model_with_dense_column_terms(tmp_path):
    """Transit compartments sharing a rate constant: a linear model whose
    column term D_k has two non-zero columns (the k-stage form with the dense
    (W^-1 D_k) V product; every built-in linear model takes the rank-one
    path).  Compiled on first use, compared with the exact solution."""
    from tests.test_host import TRANSIT_SBML
    path = tmp_path / 'transit.xml'
    path.write_text(TRANSIT_SBML)
    model = chi_b200.PKPDModel(str(path))
    model.set_administration('gut', amount_var='a1_amount', direct=True)
    model.set_outputs(['central.drug_amount'])
    model.set_dosing_regimen(dose=3, start=0, duration=0.5, period=3, num=3)
    model.set_tolerance(**TIGHT)
    model.enable_sensitivities(True)
    psi = np.array([0.3, 0.1, 0.2, 2.0, 0.7, 1.9, 1.0, 1.0])
    times = np.array([0.2, 0.5, 1.0, 1.7, 2.5, 4.0, 6.5, 9.0])
    y, s = model.simulate(psi, times)
    omodel = ode_exact.OdeModel(
        ['central.drug_amount', 'gut.a1_amount', 'transit.a2_amount'],
        ['central.size', 'global.k_e', 'global.k_tr', 'gut.size',
         'transit.size'],
        {'central.drug_amount':
            'global__k_tr * transit__a2_amount'
            ' - global__k_e * central__drug_amount',
         'gut.a1_amount': '-global__k_tr * gut__a1_amount',
         'transit.a2_amount':
            'global__k_tr * gut__a1_amount'
            ' - global__k_tr * transit__a2_amount'},
        {'central.drug_amount': 'central__drug_amount'}, 'gut.a1_amount')
    omodel.set_outputs(['central.drug_amount'])
    assert omodel.parameter_names == model.parameters()
    y_ref, s_ref = ode_exact.simulate(omodel, psi, times, _events(model))
    assert _scaled_err(y, y_ref, axis=1) < 2e-10
    assert _scaled_err(s[:, 0, :], s_ref[:, 0, :]) < 2e-9


def test_dataframe_ingestion_and_batched_sampler():
    """Long-format table -> device problem (chi/_problems.py:246-376) and a
    short batched Haario-Bardenet run over it (chi/_inference.py:702-760)."""
    import pandas as pd
    from chi_b200 import _synthetic
    n_ids = 6
    prob = _synthetic.two_compartment_problem(n_ids, seed=2, n_doses=2)
    rows = []
    for i in range(n_ids):
        for t, v in zip(prob['times'][i], prob['observations'][i]):
            rows.append({'ID': 'p%d' % i, 'Time': t,
                         'Observable': 'central.drug_concentration',
                         'Value': v, 'Dose': np.nan, 'Duration': np.nan})
        ev = prob['regimens'][i].as_array()[0]
        for k in range(2):
            rows.append({'ID': 'p%d' % i, 'Time': ev[1] + k * ev[3],
                         'Observable': np.nan, 'Value': np.nan,
                         'Dose': ev[0] * ev[2], 'Duration': ev[2]})
    data = pd.DataFrame(rows)
    pm = prob['population_model']
    hll = chi_b200.hierarchical_log_likelihood_from_dataframe(
        data, prob['model'], [prob['error_model']], pm)
    assert hll.get_id(unique=True) == ['p%d' % i for i in range(n_ids)]
    x = prob['x0']
    a = hll.evaluateS1(x)
    b = prob['log_likelihood'].evaluateS1(x)
    assert a[0] == pytest.approx(b[0], rel=1e-12)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-10, atol=1e-12)

    post = chi_b200.HierarchicalLogPosterior(hll, prob['log_prior'])
    ctrl = chi_b200.BatchedSamplingController(post, seed=1)
    ctrl.set_n_runs(16)
    ctrl.set_initial_parameters(x)
    out = ctrl.run(n_iterations=300)
    assert out['samples'].shape == (16, 300, post.n_parameters())
    assert np.all(np.isfinite(out['log_pdf']))
    assert np.all(out['acceptance_rate'] > 0.01)
    assert out['parameter_names'][-1] == 'Pooled Sigma log'
    # HMC uses the gradients; a handful of iterations must keep the chains
    # in a region of finite posterior density
    ctrl.set_sampler(chi_b200.HamiltonianMCMC)
    out = ctrl.run(n_iterations=5, hyperparameters={
        'n_leapfrog': 3, 'step_size': 0.02})
    assert np.all(np.isfinite(out['log_pdf']))


def test_fix_parameters_matches_unfixed_model():
    """LogLikelihood.fix_parameters (chi/_log_pdfs.py:1029-1073) and its
    hierarchical counterpart: fixing the initial amounts at 0 must give the
    score and the gradient of the free parameters of the full model."""
    from chi_b200 import _synthetic
    n_ids = 9
    prob = _synthetic.two_compartment_problem(
        n_ids, seed=4, n_doses=2, rtol=1e-10, atol=1e-12)
    full = prob['log_likelihood']
    x = prob['x0'].copy()
    n_bottom = 5 * n_ids
    x[n_bottom:n_bottom + 2] = 0.0          # pooled initial amounts
    s_full, g_full = full.evaluateS1(x)

    # single individual
    model = prob['model']
    reg0 = prob['regimens'][0]
    m0 = model.copy()
    m0.set_dosing_regimen(reg0)
    ll = chi_b200.LogLikelihood(
        m0, chi_b200.LogNormalErrorModel(), prob['observations'][0],
        prob['times'][0])
    psi0 = np.concatenate([[0.0, 0.0], [2.1, 0.9, 4.0, 1.2, 9.0], [0.11]])
    ref_s, ref_g = ll.evaluateS1(psi0)
    ll.fix_parameters({'central.drug_amount': 0.0,
                       'peripheral.drug_amount': 0})
    assert ll.n_parameters() == 6 and ll.n_fixed_parameters() == 2
    assert ll.get_parameter_names()[0] == 'central.size'
    s, g = ll.evaluateS1(psi0[2:])
    assert s == pytest.approx(ref_s, rel=1e-10)
    np.testing.assert_allclose(g, ref_g[2:], rtol=1e-7, atol=1e-9)
    assert ll(psi0[2:]) == pytest.approx(ref_s, rel=1e-8)
    ll.fix_parameters({'central.drug_amount': None,
                       'peripheral.drug_amount': None})
    assert ll.n_parameters() == 8

    # hierarchical: population over the six free parameters only
    pm = chi_b200.ComposedPopulationModel([
        chi_b200.LogNormalModel(n_dim=5, centered=False),
        chi_b200.PooledModel(n_dim=1)])
    hll = chi_b200.HierarchicalLogLikelihood.from_arrays(
        model, [chi_b200.LogNormalErrorModel()], prob['times'],
        prob['observations'], pm, regimens=prob['regimens'],
        fixed_parameters={'central.drug_amount': 0.0,
                          'peripheral.drug_amount': 0.0})
    assert hll.n_parameters() == full.n_parameters() - 2
    x_fixed = np.concatenate([x[:n_bottom], x[n_bottom + 2:]])
    s_fix, g_fix = hll.evaluateS1(x_fixed)
    assert s_fix == pytest.approx(s_full, rel=1e-10)
    g_ref = np.concatenate([g_full[:n_bottom], g_full[n_bottom + 2:]])
    np.testing.assert_allclose(
        g_fix, g_ref, rtol=1e-7, atol=1e-8 * np.max(np.abs(g_ref)))
    assert hll(x_fixed) == pytest.approx(full(x), rel=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize('error', ['lognormal', 'cam', 'gaussian'])
def test_deferred_and_fused_observation_agree(error):
    """The streamed kernel either evaluates the error model inside the solve
    (fused) or stores snapshots at the records and evaluates them in a second
    kernel (deferred, the default): same arithmetic, same results -- also for
    failed systems and non-positive outputs."""
    from chi_b200 import _synthetic
    prob = _synthetic.two_compartment_problem(
        n_ids=37, seed=3, n_obs=7, n_doses=2, error=error)
    ll = prob['log_likelihood']
    rng = np.random.default_rng(5)
    X = prob['x0'][None, :] * (1 + 0.02 * rng.normal(size=(6, len(prob['x0']))))
    X[4, -1] = -0.1                      # invalid sigma -> -inf
    res = {}
    # (thread-per-system kernel: the warp-per-system form of small launches
    # always observes inside the solve)
    ll.set_small_batch(0)
    for name, budget in (('deferred', -1), ('fused', 0)):
        ll.set_snapshot_budget(budget)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            s, g = ll.evaluateS1_batch(X)
            res[name] = (np.array(s), np.array(g), np.array(
                ll.evaluate_batch(X)))
        launches = int(ll.counters()[3])
        res[name + '_launches'] = launches
    ll.set_snapshot_budget(-1)
    ll.set_small_batch(-1)
    assert res['deferred_launches'] == res['fused_launches'] + 1
    sd, gd, fd = res['deferred']
    sf, gf, ff = res['fused']
    assert sd[4] == -np.inf and sf[4] == -np.inf
    ok = np.isfinite(sf)
    assert ok.sum() == 5
    np.testing.assert_allclose(sd[ok], sf[ok], rtol=1e-13)
    np.testing.assert_allclose(
        gd[ok], gf[ok], rtol=1e-11, atol=1e-11 * np.max(np.abs(gf[ok])))
    np.testing.assert_allclose(fd[ok], ff[ok], rtol=1e-13)
    assert np.array_equal(gd[4], gf[4], equal_nan=True)


@pytest.mark.gpu
def test_individual_shards_sum_to_the_full_evaluation():
    """Multi-GPU individuals-sharded mode on one device: the partial scores
    and top-level gradients of two shards add up to the unsharded result, each
    shard returns its own block of the bottom-level gradient and leaves the
    other individuals' entries zero (chi_b200_problem_set_shard; only the
    shard's column blocks are copied)."""
    from chi_b200 import _synthetic
    from chi_b200._distributed import shard_bounds
    prob = _synthetic.two_compartment_problem(
        n_ids=23, seed=11, n_obs=6, n_doses=2, error='cam', covariates=True)
    ll = prob['log_likelihood']
    n_params = len(prob['x0'])
    n_top = ll.n_parameters(exclude_bottom_level=True)
    n_bottom = n_params - n_top
    n_hdim = n_bottom // 23
    rng = np.random.default_rng(2)
    X = prob['x0'][None, :] * (1 + 0.02 * rng.normal(size=(3, n_params)))
    full_s, full_g = ll.evaluateS1_batch(X)
    full_f = ll.evaluate_batch(X)
    tot_s = np.zeros(3)
    tot_f = np.zeros(3)
    tot_g = np.zeros_like(full_g)
    for rank in range(2):
        b, e = shard_bounds(23, rank, 2)
        ll.set_shard(b, e)
        s, g = ll.evaluateS1_batch(X)
        tot_f += ll.evaluate_batch(X)
        outside = np.ones(n_bottom, dtype=bool)
        outside[b * n_hdim:e * n_hdim] = False
        assert np.all(g[:, :n_bottom][:, outside] == 0.0)
        tot_s += s
        tot_g += g
    ll.set_shard(0, 23)
    np.testing.assert_allclose(tot_s, full_s, rtol=1e-12)
    np.testing.assert_allclose(tot_f, full_f, rtol=1e-12)
    np.testing.assert_allclose(
        tot_g, full_g, rtol=1e-10, atol=1e-10 * np.max(np.abs(full_g)))
    s, g = ll.evaluateS1_batch(X)
    np.testing.assert_allclose(g, full_g, rtol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize('C', [77, 120])
def test_results_do_not_depend_on_how_the_batch_is_worked_through(C):
    """Chunks of the host entry point on two streams, individual-major work
    order and whole-warp refill of the persistent solve kernel
    (chi_b200_problem_set_batching) only change the schedule: every system
    runs the same arithmetic and every vector's reduction keeps its order, so
    scores and gradients are bit-identical to the plain pass -- with and
    without a shard of individuals, with and without gradient."""
    from chi_b200 import _synthetic
    n_ids = 37
    prob = _synthetic.two_compartment_problem(
        n_ids=n_ids, seed=5, n_obs=7, n_doses=3, rtol=1e-6, atol=1e-8)
    ll = prob['log_likelihood']
    n_params = len(prob['x0'])
    rng = np.random.default_rng(4)
    # 77: not a multiple of the warp size or of the chunk count (general
    # kernel in individual-major order); 120: four warps per individual, the
    # last one padded with ghost lanes, vectors handed out in the stiffness
    # order (warp-synchronous kernel)
    X = prob['x0'][None, :] * (1 + 0.05 * rng.normal(size=(C, n_params)))
    X[5, -1] = -0.1       # invalid sigma: -inf score, inf gradient
    # (the warp-per-system form of small launches is a different
    # discretisation -- a step-size control per sensitivity column -- and is
    # compared within the tolerance contract elsewhere)
    ll.set_small_batch(0)
    ll.set_batching(n_chunks=1, order_min_vectors=0, refill_wait=0)
    with np.errstate(all='ignore'):
        ref_s, ref_g = ll.evaluateS1_batch(X)
        ref_f = ll.evaluate_batch(X)
    assert np.isneginf(ref_s[5]) and np.all(np.isfinite(np.delete(ref_s, 5)))
    settings = [(3, 0, 0), (1, 32, -1), (1, 32, 0), (1, 32, 5), (4, 32, -1),
                (8, 16, 2), (-1, -1, -1)]
    for chunks, order, wait in settings:
        ll.set_batching(chunks, order, wait)
        with np.errstate(all='ignore'):
            s, g = ll.evaluateS1_batch(X)
            f = ll.evaluate_batch(X)
        np.testing.assert_array_equal(s, ref_s)
        np.testing.assert_array_equal(g, ref_g)
        np.testing.assert_array_equal(f, ref_f)
    # the same under a shard of individuals (column blocks copied per chunk)
    ll.set_shard(11, 30)
    ll.set_batching(1, 0, 0)
    with np.errstate(all='ignore'):
        sh_s, sh_g = ll.evaluateS1_batch(X)
    ll.set_batching(4, 32, -1)
    with np.errstate(all='ignore'):
        s, g = ll.evaluateS1_batch(X)
    np.testing.assert_array_equal(s, sh_s)
    np.testing.assert_array_equal(g, sh_g)
    ll.set_shard(0, n_ids)
    ll.set_batching()
    ll.set_small_batch(-1)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['tgi_direct_daily', 'two_cmt_indirect_indefinite'])
def test_log_likelihood_nonlinear_and_indirect_models(name):
    """LogLikelihood.evaluateS1 of the nonlinear erlotinib TGI model (30 daily
    doses) and the 3-state indirectly dosed two-compartment model against the
    oracle (exact / DOP853 solution + numpy error model), through both
    observation paths of the streamed kernel."""
    model, omodel, psi, times = _build(name)
    model.set_tolerance(**TIGHT)
    times = times[times > 0]
    rng = np.random.default_rng(4)
    y_ref, _ = ode_exact.simulate(omodel, psi, times, _events(model))
    obs = np.abs(y_ref[0] * np.exp(0.1 * rng.normal(size=len(times)))) + 1e-6
    ll = chi_b200.LogLikelihood(
        model, chi_b200.LogNormalErrorModel(), obs, times)
    x = np.array(list(psi) + [0.15])
    _, f_s1 = _oracle_individual(
        omodel, 'lognormal', times, obs, _events(model), len(psi))
    score_ref, grad_ref = f_s1(x)
    for budget in (-1, 0):
        ll._get_device().set_snapshot_budget(budget)
        score, grad = ll.evaluateS1(x)
        assert score == pytest.approx(score_ref, rel=1e-8)
        np.testing.assert_allclose(
            grad, grad_ref, rtol=2e-6, atol=2e-6 * np.max(np.abs(grad_ref)))
