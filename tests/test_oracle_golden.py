"""The CPU oracle against the reference's golden vectors (no GPU).

``tests/golden/*.json`` are outputs of the unmodified reference
(tests/golden/make_golden.py); ``tests/kats.py`` restates the closed-form
known answers of the reference's unit tests.
"""
import json
import os

import numpy as np
import pytest

from oracle import ode_exact, stats
from tests import kats

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    with open(os.path.join(HERE, 'golden', name)) as f:
        return json.load(f)


def test_error_model_known_answers():
    for kind, sigma, ybar, sens, obs, score, grad in kats.error_model_kats():
        n_sig, f_ll, f_s1 = stats.ERROR_MODELS[kind]
        ll, g = f_s1(sigma, ybar, sens, obs)
        assert ll == pytest.approx(score, rel=1e-12, abs=1e-12)
        assert f_ll(sigma, ybar, obs) == pytest.approx(score, rel=1e-12)
        assert np.sum(stats.error_pointwise_ll(kind, sigma, ybar, obs)) == \
            pytest.approx(score, rel=1e-12)
        if grad is not None:
            np.testing.assert_allclose(g, grad, rtol=1e-12, atol=1e-12)


def test_error_model_conventions():
    """sigma <= 0 -> -inf / inf; lognormal ybar <= 0 -> -inf
    (chi/_error_models.py:238-240, 297-300, 590-592, 641-644, 934-936,
    993-996)."""
    s = np.ones((2, 2))
    for kind, (n_sig, f_ll, f_s1) in stats.ERROR_MODELS.items():
        ll, g = f_s1([-0.1] * n_sig, [1.0, 2.0], s, [1.0, 2.0])
        assert ll == -np.inf and np.all(g == np.inf)
        assert len(g) == 2 + n_sig
        assert f_ll([0.0] * n_sig, [1.0], [1.0]) == -np.inf
    ll, g = stats.lognormal_error_s1([0.2], [1.0, -1.0], s, [1.0, 2.0])
    assert ll == -np.inf and np.all(g == np.inf)


def test_error_models_match_reference_golden():
    for case in _load('stats_golden.json')['error_models']:
        n_sig, f_ll, f_s1 = stats.ERROR_MODELS[case['kind']]
        ll, g = f_s1(case['sigma'], case['ybar'], np.array(case['sens']),
                     case['obs'])
        assert ll == pytest.approx(case['ll'], rel=1e-13)
        np.testing.assert_allclose(g, case['grad'], rtol=1e-12, atol=1e-14)
        assert f_ll(case['sigma'], case['ybar'], case['obs']) == \
            pytest.approx(case['ll_call'], rel=1e-13)
        np.testing.assert_allclose(
            stats.error_pointwise_ll(
                case['kind'], case['sigma'], case['ybar'], case['obs']),
            case['pointwise'], rtol=1e-13)


def _layout_from_golden(case):
    blocks = []
    sel = iter(case['selections'])
    for b in case['spec']:
        if b[0] == 'covariate':
            pidx, didx = next(sel)
            blocks.append(('covariate', tuple(b[1]), b[2],
                           np.array(pidx, int), np.array(didx, int)))
        else:
            blocks.append(tuple(b))
    return stats.PopLayout(blocks, case['n_ids'])


def test_population_models_match_reference_golden():
    for case in _load('stats_golden.json')['population_models']:
        lay = _layout_from_golden(case)
        assert (lay.n_bottom, lay.n_top) == (case['n_bottom'], case['n_top'])
        cov = None if case['covariates'] is None else \
            np.array(case['covariates'])
        top, bottom = np.array(case['top']), np.array(case['bottom'])
        eta = stats.pop_individual_parameters(
            lay, top, bottom, cov, return_eta=True)
        np.testing.assert_allclose(eta, case['eta'], rtol=1e-14)
        psi = stats.pop_individual_parameters(lay, top, eta, cov)
        np.testing.assert_allclose(psi, case['psi'], rtol=1e-13)
        assert stats.pop_log_likelihood(lay, top, eta, cov) == \
            pytest.approx(case['score'], rel=1e-13)
        s, g = stats.pop_sensitivities_reduced(
            lay, top, eta, np.array(case['dlogp_dpsi']), cov)
        assert s == pytest.approx(case['score_s1'], rel=1e-13)
        np.testing.assert_allclose(g, case['grad'], rtol=1e-12, atol=1e-13)


def test_population_sigma_negative_and_nan_rules():
    """chi/_population_models.py:1639-1641, 2480-2482 (sigma < 0 -> -inf),
    1594-1595, 2438-2439 (psi NaN)."""
    lay = stats.PopLayout([('lognormal', 2, False), ('pooled', 1)], 3)
    top = np.array([0.1, 0.2, 0.3, -0.4, 1.0])
    bottom = np.linspace(0.1, 0.6, 6)
    eta = stats.pop_individual_parameters(lay, top, bottom, return_eta=True)
    psi = stats.pop_individual_parameters(lay, top, eta)
    assert np.all(np.isnan(psi[:, :2])) and np.all(psi[:, 2] == 1.0)
    s, _ = stats.pop_sensitivities_reduced(lay, top, eta, np.ones((3, 3)))
    assert s == -np.inf
    lay = stats.PopLayout([('gaussian', 1, True)], 2)
    assert stats.pop_log_likelihood(
        lay, np.array([0.0, -1.0]), np.array([[0.1], [0.2]])) == -np.inf


class _Prior(object):
    def __init__(self, kinds, locs, scales):
        self.kinds, self.locs, self.scales = kinds, locs, scales

    def evaluateS1(self, x):
        score, grad = 0.0, np.zeros(len(x))
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


def golden_layout(case):
    """Oracle layout of a hierarchical golden case; the covariate selection
    order is read back from the reference's parameter names."""
    names = case['ll_names']
    mid = ('lognormal', 5, False)
    n_sig = len(names) - 7
    if case['covariates'] is not None:
        pidx, didx = [], []
        cov_names = [n for n in case['top_names'] if 'Cov.' in n]
        for n in cov_names[::2]:
            pidx.append(0 if n.startswith('Log mean') else 1)
            base = n.replace('Log mean ', '').replace('Log std. ', '')
            base = base[:base.index(' Cov.')]
            didx.append(names.index(base) - 2)
        mid = ('covariate', mid, 2, np.array(pidx), np.array(didx))
    return stats.PopLayout(
        [('pooled', 2), mid, ('pooled', n_sig)], case['n_ids'])


@pytest.mark.parametrize('name', ['c2', 'c5', 'c2_120'])
def test_hierarchical_posterior_matches_reference_golden(name):
    """oracle (numpy restatement + exact ODE) == unmodified reference
    classes driving the same exact ODE."""
    case = _load('round2_golden.json')[name] if name == 'c2_120' \
        else _load('hier_golden.json')[name]
    lay = golden_layout(case)
    omodel = ode_exact.two_compartment(True)
    omodel.set_outputs([case['output']])
    cov = None if case['covariates'] is None else np.array(case['covariates'])

    def individual(i, psi):
        times = np.array(case['times'][i])
        ev = [tuple(e) for e in case['events'][i]]

        def simulate(p, with_sens):
            return ode_exact.simulate(omodel, p, times, ev, with_sens)
        return stats.loglikelihood_s1(
            simulate, 7, [case['error']], [np.array(case['obs'][i])],
            np.ones((1, len(times)), bool), psi, True)

    prior = _Prior(*case['prior'])
    x = np.array(case['x'])
    score, grad = stats.posterior_s1(lay, individual, prior, x, cov)
    assert score == pytest.approx(case['score'], rel=1e-12)
    np.testing.assert_allclose(
        grad, case['grad'], rtol=1e-10,
        atol=1e-12 * np.max(np.abs(case['grad'])))
    score0 = stats.posterior_call(
        lay, lambda i, p: individual(i, p)[0], prior, x, cov)
    assert score0 == pytest.approx(case['score_call'], rel=1e-12)


def test_c1_log_posterior_matches_reference_golden():
    case = _load('hier_golden.json')['c1']
    omodel = ode_exact.one_compartment(True)
    times = np.array(case['times'])
    ev = [tuple(e) for e in case['events']]

    def simulate(p, with_sens):
        return ode_exact.simulate(omodel, p, times, ev, with_sens)
    x = np.array(case['x'])
    ll, g = stats.loglikelihood_s1(
        simulate, 3, ['gaussian'], [np.array(case['obs'])],
        np.ones((1, len(times)), bool), x, True)
    p, pg = _Prior(*case['prior']).evaluateS1(x)
    assert ll + p == pytest.approx(case['score'], rel=1e-12)
    np.testing.assert_allclose(g + pg, case['grad'], rtol=1e-10)


def test_exact_ode_oracle_closed_form():
    """one-compartment bolus has an elementary closed form."""
    m = ode_exact.one_compartment(True)
    times = np.linspace(0.1, 10, 50)
    r, T, k, V = 1000.0, 0.01, 1.3, 2.0
    y, s = ode_exact.simulate(m, [0.5, V, k], times, [(r, 0.0, T, 0, 0)])
    a_T = 0.5 * np.exp(-k * T) + r / k * (1 - np.exp(-k * T))
    amount = a_T * np.exp(-k * (times - T))
    np.testing.assert_allclose(y[0], amount / V, rtol=1e-12)
    np.testing.assert_allclose(s[:, 0, 0], np.exp(-k * times) / V, rtol=1e-12)
    np.testing.assert_allclose(s[:, 0, 1], -amount / V ** 2, rtol=1e-12)
    da_T = -T * 0.5 * np.exp(-k * T) - r / k ** 2 * (1 - np.exp(-k * T)) \
        + r / k * T * np.exp(-k * T)
    dk = (da_T - (times - T) * a_T) * np.exp(-k * (times - T)) / V
    np.testing.assert_allclose(s[:, 0, 2], dk, rtol=1e-11)


def test_exact_ode_oracle_linear_vs_nonlinear_path():
    m = ode_exact.two_compartment(False)
    p = [0.1, 0.2, 0.3, 2, 10, 1, 5, 1, 10]
    reg = [(200.0, 0.0, 0.01, 1.0, 0)]
    ts = [0.5, 1, 1.5, 2, 2.5, 3]
    y1, s1 = ode_exact.simulate(m, p, ts, reg)
    m.linear = False
    y2, s2 = ode_exact.simulate(m, p, ts, reg)
    np.testing.assert_allclose(y1, y2, rtol=1e-11)
    np.testing.assert_allclose(s1, s2, rtol=1e-9, atol=1e-12)


def test_regimen_expansion():
    """myokit blocktrain semantics (chi/_mechanistic_models.py:834-855)."""
    edges, rates = ode_exact.expand_regimen([(5.0, 1.0, 0.5, 2.0, 3)], 10.0)
    np.testing.assert_allclose(
        edges, [0, 1, 1.5, 3, 3.5, 5, 5.5, 10])
    np.testing.assert_allclose(rates, [0, 5, 0, 5, 0, 5, 0])
    edges, rates = ode_exact.expand_regimen([(2.0, 0.0, 0.01, 0, 0)], 1.0)
    np.testing.assert_allclose(edges, [0, 0.01, 1.0])
    np.testing.assert_allclose(rates, [2.0, 0.0])
    # indefinite train stops at the horizon
    edges, rates = ode_exact.expand_regimen([(1.0, 0.0, 0.25, 1.0, 0)], 2.5)
    assert len(rates) == 6 and rates[-1] == 0.0


def _pred_layout(case):
    blocks = []
    for b in case['spec']:
        if b[0] == 'covariate':
            names = case['theta_names']
            cov_names = [n for n in names if 'Cov.' in n][::b[2]]
            pidx = [0 if n.startswith('Log mean') else 1 for n in cov_names]
            first = min(int(n.split('Dim. ')[1].split(' ')[0])
                        for n in cov_names)
            didx = [int(n.split('Dim. ')[1].split(' ')[0]) - first
                    for n in cov_names]
            blocks.append(('covariate', tuple(b[1]), b[2],
                           np.array(pidx), np.array(didx)))
        else:
            blocks.append(tuple(b))
    return stats.PopLayout(blocks, case['n_samples'])


@pytest.mark.parametrize('name', ['predictive_c3', 'predictive_c5'])
def test_population_predictive_sample_matches_reference_golden(name):
    """PopulationPredictiveModel.sample(return_df=False) of the unmodified
    reference (exact ODE plugged in) == oracle restatement, same seed."""
    case = _load('hier_golden.json')[name]
    if name == 'predictive_c3':
        omodel = ode_exact.tumour_growth_inhibition_pkpd(True)
    else:
        omodel = ode_exact.two_compartment(True)
    omodel.set_outputs([case['output']])
    ev = [tuple(e) for e in case['events']]
    times = np.array(case['times'])

    def simulate(psi):
        return ode_exact.simulate(omodel, psi, times, ev, False)
    out = stats.population_predictive_sample(
        _pred_layout(case), simulate, omodel.P, [case['error']],
        np.array(case['theta']), times, case['n_samples'], case['seed'],
        case['covariates'])
    np.testing.assert_allclose(out, case['samples'], rtol=1e-12)


def test_truncated_gaussian_matches_reference_golden():
    """TruncatedGaussianModel (chi/_population_models.py:3500-3939), stand-
    alone as the reference supports it: log-likelihood, reduced and flat
    sensitivities, -inf rules."""
    for case in _load('stats_golden_extra.json')['truncated_gaussian']:
        d, n_ids = case['n_dim'], case['n_ids']
        lay = stats.PopLayout([('truncgauss', d, True)], n_ids)
        theta = np.array(case['theta'])
        obs = np.array(case['obs'])
        dlp = np.array(case['dlogp_dpsi'])
        assert stats.pop_log_likelihood(lay, theta, obs) == pytest.approx(
            case['ll'], rel=1e-13)
        s, g = stats.pop_sensitivities_reduced(lay, theta, obs, dlp)
        assert s == pytest.approx(case['score_reduced'], rel=1e-13)
        np.testing.assert_allclose(
            g, case['grad_reduced'], rtol=1e-12, atol=1e-13)
        # flat form: d psi (n_ids, d) and d theta summed over individuals
        np.testing.assert_allclose(
            g[:n_ids * d].reshape(n_ids, d), case['dpsi'], rtol=1e-12)
        neg = obs.copy()
        neg[0, 0] = -0.1
        assert stats.pop_log_likelihood(lay, theta, neg) == -np.inf
        bad = theta.copy()
        bad[d] = 0.0
        assert stats.pop_log_likelihood(lay, bad, obs) == -np.inf
        assert case['ll_negative_obs'] == -np.inf
        assert case['ll_zero_sigma'] == -np.inf
