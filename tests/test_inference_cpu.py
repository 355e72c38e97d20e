"""Batched samplers on a plain numpy log-pdf (no GPU): they must recover the
mean and spread of a correlated Gaussian."""
import numpy as np
import pytest

from chi_b200 import _inference


class _Gaussian(object):
    def __init__(self):
        self.mean = np.array([1.0, -2.0, 0.5])
        a = np.array([[1.0, 0.3, 0.0], [0.0, 0.5, 0.2], [0.0, 0.0, 2.0]])
        self.cov = a @ a.T
        self.prec = np.linalg.inv(self.cov)
        self.calls = 0

    def n_parameters(self):
        return 3

    def evaluate_batch(self, X):
        self.calls += 1
        d = np.asarray(X) - self.mean
        return -0.5 * np.einsum('ci,ij,cj->c', d, self.prec, d)

    def evaluateS1_batch(self, X):
        d = np.asarray(X) - self.mean
        return self.evaluate_batch(X), -d @ self.prec


@pytest.mark.parametrize('sampler,n_iter', [
    (_inference.HaarioBardenetACMC, 4000), (_inference.HamiltonianMCMC, 600)])
def test_batched_sampler_recovers_gaussian(sampler, n_iter):
    target = _Gaussian()
    ctrl = _inference.BatchedSamplingController(target, seed=3)
    ctrl.set_n_runs(32)
    ctrl.set_sampler(sampler)
    ctrl.set_initial_parameters(np.zeros(3))
    out = ctrl.run(n_iterations=n_iter)
    assert out['samples'].shape == (32, n_iter, 3)
    tail = out['samples'][:, n_iter // 2:, :].reshape(-1, 3)
    np.testing.assert_allclose(tail.mean(0), target.mean, atol=0.12)
    np.testing.assert_allclose(
        np.sqrt(np.diag(np.cov(tail.T))), np.sqrt(np.diag(target.cov)),
        rtol=0.12)
    assert np.all(out['acceptance_rate'] > 0.05)
    assert np.all(np.isfinite(out['log_pdf']))
    # one batched evaluation per iteration (plus the initial point), not one
    # per chain
    if sampler is _inference.HaarioBardenetACMC:
        assert target.calls == n_iter + 1
        assert out['evaluations'] == 32 * (n_iter + 1)


def test_controller_input_checks():
    target = _Gaussian()
    ctrl = _inference.BatchedSamplingController(target)
    with pytest.raises(ValueError, match='at least 1'):
        ctrl.set_n_runs(0)
    with pytest.raises(ValueError, match='wrong shape'):
        ctrl.set_initial_parameters(np.zeros((2, 2)))
    with pytest.raises(ValueError, match='Sampler has to be'):
        ctrl.set_sampler(int)
    with pytest.raises(ValueError, match='have not been set'):
        ctrl.run(3)
    with pytest.raises(ValueError, match='has to provide'):
        _inference.BatchedSamplingController(object())
    ctrl.set_initial_parameters(np.full(3, np.inf))
    with pytest.raises(ValueError, match='finite logpdf'):
        ctrl.run(3)
