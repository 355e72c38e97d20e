"""Randomised cross-checks of the oracle against the unmodified reference
modules.  Only runs where /root/reference exists (the build container); the
committed golden vectors cover the same ground elsewhere."""
import warnings

import numpy as np
import pytest

from oracle import refload, stats

pytestmark = pytest.mark.skipif(
    not refload.available(), reason='reference checkout not present')


@pytest.fixture(scope='module')
def chi():
    warnings.simplefilter('ignore')
    return refload.load()


def test_error_models_random(chi):
    rng = np.random.default_rng(0)
    pairs = [('gaussian', chi.GaussianErrorModel()),
             ('lognormal', chi.LogNormalErrorModel()),
             ('cam', chi.ConstantAndMultiplicativeGaussianErrorModel()),
             ('multiplicative', chi.MultiplicativeGaussianErrorModel())]
    for kind, em in pairs:
        n_sig = em.n_parameters()
        for _ in range(10):
            n, p = rng.integers(1, 20), rng.integers(1, 8)
            yb, y = rng.uniform(0.5, 3, n), rng.uniform(0.5, 3, n)
            s = rng.normal(size=(n, p))
            sig = rng.uniform(0.1, 1, n_sig)
            a = em.compute_sensitivities(sig, yb, s, y)
            b = stats.ERROR_MODELS[kind][2](sig, yb, s, y)
            assert a[0] == pytest.approx(b[0], rel=1e-12)
            np.testing.assert_allclose(a[1], b[1], rtol=1e-11, atol=1e-13)


class _ToyModel(object):
    """y = y0 exp(g t) with exact sensitivities, the reference's own toy
    model (chi/tests/test_log_pdfs.py:18-113)."""

    @staticmethod
    def simulate(p, times, sens):
        y0, g = p
        times = np.asarray(times)
        y = (y0 * np.exp(g * times))[None, :]
        if not sens:
            return y
        s = np.empty((len(times), 1, 2))
        s[:, 0, 0] = np.exp(g * times)
        s[:, 0, 1] = times * y[0]
        return y, s


def test_hierarchical_random(chi):
    rng = np.random.default_rng(1)

    class Toy(chi.MechanisticModel):
        def __init__(self):
            super(Toy, self).__init__()
            self._s = False

        def copy(self):
            return Toy()

        def enable_sensitivities(self, enabled, parameter_names=None):
            self._s = bool(enabled)

        def has_sensitivities(self):
            return self._s

        def n_outputs(self):
            return 1

        def n_parameters(self):
            return 2

        def outputs(self):
            return ['Count']

        def parameters(self):
            return ['Initial count', 'Growth rate']

        def simulate(self, parameters, times):
            return _ToyModel.simulate(parameters, times, self._s)

    n_ids = 9
    times = [np.sort(rng.uniform(0, 2, 4)) for _ in range(n_ids)]
    obs = [rng.uniform(0.5, 3, 4) for _ in range(n_ids)]
    lls = [chi.LogLikelihood(Toy(), chi.LogNormalErrorModel(), obs[i],
                             times[i]) for i in range(n_ids)]
    pm = chi.ComposedPopulationModel([
        chi.LogNormalModel(n_dim=2, centered=False), chi.PooledModel(1)])
    hll = chi.HierarchicalLogLikelihood(lls, pm)
    lay = stats.PopLayout([('lognormal', 2, False), ('pooled', 1)], n_ids)

    def ind(i, psi):
        def sim(p, sens):
            return _ToyModel.simulate(p, times[i], sens)
        return stats.loglikelihood_s1(
            sim, 2, ['lognormal'], [obs[i]], np.ones((1, 4), bool), psi, True)

    for _ in range(3):
        x = np.concatenate([rng.normal(size=2 * n_ids) * 0.3,
                            [0.1, 0.2, 0.3, 0.4, 0.2]])
        a = hll.evaluateS1(x)
        b = stats.hierarchical_s1(lay, ind, x)
        assert a[0] == pytest.approx(b[0], rel=1e-12)
        np.testing.assert_allclose(a[1], b[1], rtol=1e-10, atol=1e-12)
        c = stats.hierarchical_call(lay, lambda i, p: ind(i, p)[0], x)
        assert hll(x) == pytest.approx(c, rel=1e-12)
