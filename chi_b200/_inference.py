"""Batched MCMC driver for the GPU log-posteriors.

chi's ``SamplingController`` (chi/_inference.py:605-770) hands the posterior
to ``pints.MCMCController``, which asks every chain for a proposal and then
evaluates the proposals one vector at a time (optionally in forked worker
processes, chi/_inference.py:740).  A GPU wants the opposite: all chains'
proposals in ONE call.  This module is that driver (SURVEY.md section 8f,
"next" row 1): ask/tell samplers vectorised over chains on the host (numpy)
feeding ``evaluate_batch`` / ``evaluateS1_batch``.

Samplers
--------
``HaarioBardenetACMC``  adaptive covariance Metropolis, chi's default sampler
                        (chi/_inference.py:630; algorithm of
                        pints.HaarioBardenetACMC: global scaling
                        ``exp(log_lambda)`` adapted towards a 23.4 %
                        acceptance rate, running mean/covariance with
                        ``gamma = (t+1)^-eta``, ``eta = 0.6``).  A diagonal
                        covariance is used above 256 parameters, where one
                        dense matrix per chain is not affordable.
``HamiltonianMCMC``     leapfrog HMC with a per-chain step size adapted by
                        dual averaging during warm-up; uses the gradients.

The controller accepts any object with ``n_parameters()``,
``evaluate_batch(X)`` (and ``evaluateS1_batch(X)`` for HMC) -- the
chi_b200 posteriors, or a plain numpy log-pdf in the CPU tests.
"""
import numpy as np


class HaarioBardenetACMC(object):
    """Adaptive covariance Metropolis, vectorised over chains."""

    needs_sensitivities = False

    def __init__(self, x0, sigma0=None, rng=None, full_covariance=None):
        x0 = np.array(x0, dtype=float)
        self._n_chains, self._n = x0.shape
        self._x = x0
        self._rng = rng or np.random.default_rng()
        if full_covariance is None:
            full_covariance = self._n <= 256
        self._full = bool(full_covariance)
        if sigma0 is None:
            # pints' default: 1 % of the starting point per dimension (a zero
            # entry gets 1 % of one)
            sigma0 = np.abs(x0) * 0.01
            sigma0 = np.where(sigma0 == 0, 0.01, sigma0)
        sigma0 = np.broadcast_to(
            np.asarray(sigma0, dtype=float), x0.shape)
        if self._full:
            self._sigma = np.zeros((self._n_chains, self._n, self._n))
            idx = np.arange(self._n)
            self._sigma[:, idx, idx] = sigma0 ** 2
        else:
            self._sigma = sigma0 ** 2
        self._mu = x0.copy()
        self._log_lambda = np.zeros(self._n_chains)
        self._fx = None
        self._eta = 0.6
        self._target = 0.234
        self._adaptations = 2
        self._adaptive = False
        self._proposed = None
        self.acceptance = np.zeros(self._n_chains)
        self._iterations = 0

    def set_adaptive(self, adaptive):
        self._adaptive = bool(adaptive)

    def current(self):
        return self._x

    def ask(self):
        if self._fx is None:
            self._proposed = self._x.copy()
            return self._proposed
        z = self._rng.standard_normal((self._n_chains, self._n))
        scale = np.exp(0.5 * self._log_lambda)[:, None]
        if self._full:
            jitter = 1e-12 * np.eye(self._n)[None]
            chol = np.linalg.cholesky(self._sigma + jitter)
            step = np.einsum('cij,cj->ci', chol, z)
        else:
            step = np.sqrt(self._sigma) * z
        self._proposed = self._x + scale * step
        return self._proposed

    def tell(self, fx):
        fx = np.asarray(fx, dtype=float)
        if self._fx is None:
            if not np.all(np.isfinite(fx)):
                raise ValueError(
                    'Initial point for MCMC must have finite logpdf.')
            self._fx = fx.copy()
            return self._x
        with np.errstate(invalid='ignore', over='ignore'):
            log_ratio = fx - self._fx
        u = np.log(self._rng.uniform(size=self._n_chains))
        accept = np.isfinite(fx) & (u < log_ratio)
        self._x = np.where(accept[:, None], self._proposed, self._x)
        self._fx = np.where(accept, fx, self._fx)
        self._iterations += 1
        self.acceptance += (accept - self.acceptance) / self._iterations
        if self._adaptive:
            gamma = self._adaptations ** (-self._eta)
            self._adaptations += 1
            self._mu = (1 - gamma) * self._mu + gamma * self._x
            d = self._x - self._mu
            if self._full:
                outer = d[:, :, None] * d[:, None, :]
                self._sigma = (1 - gamma) * self._sigma + gamma * outer
            else:
                self._sigma = (1 - gamma) * self._sigma + gamma * d * d
            self._log_lambda += gamma * (accept - self._target)
        return self._x

    def log_pdf(self):
        return self._fx


class HamiltonianMCMC(object):
    """Leapfrog HMC vectorised over chains (identity mass matrix scaled per
    parameter, step size by dual averaging during warm-up)."""

    needs_sensitivities = True

    def __init__(self, x0, sigma0=None, rng=None, n_leapfrog=10,
                 step_size=0.1):
        x0 = np.array(x0, dtype=float)
        self._n_chains, self._n = x0.shape
        self._x = x0
        self._rng = rng or np.random.default_rng()
        if sigma0 is None:
            sigma0 = np.abs(x0) * 0.05 + 1e-3
        self._scale = np.broadcast_to(
            np.asarray(sigma0, dtype=float), x0.shape).copy()
        self._n_leapfrog = int(n_leapfrog)
        self._eps = np.full(self._n_chains, float(step_size))
        self._log_eps_bar = np.log(self._eps)
        self._h_bar = np.zeros(self._n_chains)
        self._mu = np.log(10 * self._eps)
        self._adaptive = False
        self._t = 0
        self._fx = None
        self._gx = None
        self.acceptance = np.zeros(self._n_chains)
        self._iterations = 0

    def set_adaptive(self, adaptive):
        if self._adaptive and not adaptive:
            self._eps = np.exp(self._log_eps_bar)
        self._adaptive = bool(adaptive)

    def current(self):
        return self._x

    def step(self, evaluate_s1):
        """One HMC transition for every chain; ``evaluate_s1(X)`` returns
        (scores, gradients)."""
        if self._fx is None:
            self._fx, self._gx = [np.array(a, dtype=float)
                                  for a in evaluate_s1(self._x)]
            if not np.all(np.isfinite(self._fx)):
                raise ValueError(
                    'Initial point for MCMC must have finite logpdf.')
        s = self._scale
        p0 = self._rng.standard_normal((self._n_chains, self._n))
        x, p = self._x.copy(), p0.copy()
        g = self._gx.copy()
        eps = self._eps[:, None]
        ok = np.ones(self._n_chains, dtype=bool)
        fx = self._fx.copy()
        for _ in range(self._n_leapfrog):
            p = p + 0.5 * eps * s * g
            x = x + eps * s * p
            fx, g = [np.array(a, dtype=float) for a in evaluate_s1(x)]
            bad = ~np.isfinite(fx) | ~np.all(np.isfinite(g), axis=1)
            ok &= ~bad
            g = np.where(bad[:, None], 0.0, g)
            p = p + 0.5 * eps * s * g
        with np.errstate(invalid='ignore', over='ignore'):
            h0 = -self._fx + 0.5 * np.sum(p0 * p0, axis=1)
            h1 = -fx + 0.5 * np.sum(p * p, axis=1)
            log_ratio = np.where(ok, h0 - h1, -np.inf)
        alpha = np.exp(np.minimum(0.0, np.nan_to_num(
            log_ratio, nan=-np.inf)))
        accept = self._rng.uniform(size=self._n_chains) < alpha
        self._x = np.where(accept[:, None], x, self._x)
        self._fx = np.where(accept, fx, self._fx)
        self._gx = np.where(accept[:, None], g, self._gx)
        self._iterations += 1
        self.acceptance += (accept - self.acceptance) / self._iterations
        if self._adaptive:
            # Nesterov dual averaging (Hoffman & Gelman 2014), delta = 0.8
            self._t += 1
            t, t0, kappa, gam, delta = self._t, 10.0, 0.75, 0.05, 0.8
            self._h_bar = (1 - 1 / (t + t0)) * self._h_bar \
                + (delta - alpha) / (t + t0)
            log_eps = self._mu - np.sqrt(t) / gam * self._h_bar
            w = t ** (-kappa)
            self._log_eps_bar = w * log_eps + (1 - w) * self._log_eps_bar
            self._eps = np.exp(log_eps)
        return self._x

    def log_pdf(self):
        return self._fx


class BatchedSamplingController(object):
    """Runs many MCMC chains with one batched posterior evaluation per
    iteration (counterpart of ``chi.SamplingController.run``,
    chi/_inference.py:702-760).

    :param log_posterior: object with ``n_parameters()`` and
        ``evaluate_batch(X)`` (``evaluateS1_batch(X)`` for HMC).
    """

    def __init__(self, log_posterior, seed=None):
        for name in ('n_parameters', 'evaluate_batch'):
            if not hasattr(log_posterior, name):
                raise ValueError(
                    'The log-posterior has to provide ' + name + '.')
        self._log_posterior = log_posterior
        self._rng = np.random.default_rng(seed)
        self._seed = seed
        self._n_runs = 5          # chi/_inference.py:435
        self._sampler = HaarioBardenetACMC   # chi/_inference.py:630
        self._initial_params = None
        self._initial_phase = 200  # pints.MCMCController default
        self._n_warmup = None      # HMC: end of the step-size adaptation

    def set_n_runs(self, n_runs):
        n_runs = int(n_runs)
        if n_runs < 1:
            raise ValueError('Number of runs has to be at least 1.')
        self._n_runs = n_runs
        self._initial_params = None

    def set_sampler(self, sampler):
        if sampler not in (HaarioBardenetACMC, HamiltonianMCMC):
            raise ValueError(
                'Sampler has to be HaarioBardenetACMC or HamiltonianMCMC.')
        self._sampler = sampler

    def set_initial_parameters(self, x0):
        x0 = np.asarray(x0, dtype=float)
        if x0.ndim == 1:
            x0 = np.tile(x0, (self._n_runs, 1))
        if x0.shape != (self._n_runs, self._log_posterior.n_parameters()):
            raise ValueError('Initial parameters have the wrong shape.')
        self._initial_params = x0

    def set_initial_phase_iterations(self, n):
        self._initial_phase = int(n)

    def set_warmup_iterations(self, n):
        """Iteration at which a gradient-based sampler stops adapting its
        step size and freezes it at the dual-averaging mean (default: half of
        the run).  Samples before it belong to the warm-up."""
        n = int(n)
        if n < 0:
            raise ValueError('The number of warm-up iterations is negative.')
        self._n_warmup = n

    def _x0(self):
        if self._initial_params is not None:
            return self._initial_params
        if hasattr(self._log_posterior, 'sample_initial_parameters'):
            return np.asarray(
                self._log_posterior.sample_initial_parameters(
                    n_samples=self._n_runs, seed=self._seed), dtype=float)
        raise ValueError('Initial parameters have not been set.')

    def run(self, n_iterations=10000, hyperparameters=None):
        """Returns a dict with the chains ``samples`` of shape
        ``(n_runs, n_iterations, n_parameters)``, the scores ``log_pdf``
        ``(n_runs, n_iterations)``, the per-chain ``acceptance_rate`` and the
        number of posterior ``evaluations``."""
        post = self._log_posterior
        hyper = dict(hyperparameters or {})
        sampler = self._sampler(self._x0(), rng=self._rng, **hyper)
        n, d = self._n_runs, post.n_parameters()
        samples = np.empty((n, n_iterations, d))
        log_pdf = np.empty((n, n_iterations))
        evaluations = 0
        n_warmup = self._n_warmup if self._n_warmup is not None \
            else max(self._initial_phase, n_iterations // 2)
        for it in range(n_iterations):
            if it == self._initial_phase:
                sampler.set_adaptive(True)
            if sampler.needs_sensitivities and it == n_warmup and \
                    it > self._initial_phase:
                # freeze the step size at its averaged value: the chain is
                # a valid Markov chain from here on
                sampler.set_adaptive(False)
            if sampler.needs_sensitivities:
                def f(X):
                    return post.evaluateS1_batch(X)
                x = sampler.step(f)
                evaluations += n * (sampler._n_leapfrog + (it == 0))
            else:
                if it == 0:
                    sampler.tell(post.evaluate_batch(sampler.ask()))
                    evaluations += n
                x = sampler.tell(post.evaluate_batch(sampler.ask()))
                evaluations += n
            samples[:, it] = x
            log_pdf[:, it] = sampler.log_pdf()
        names = None
        if hasattr(post, 'get_parameter_names'):
            try:
                names = post.get_parameter_names(include_ids=True)
            except TypeError:
                names = post.get_parameter_names()
        return {
            'samples': samples, 'log_pdf': log_pdf,
            'acceptance_rate': sampler.acceptance.copy(),
            'evaluations': evaluations, 'parameter_names': names,
            'warmup_iterations': n_warmup if sampler.needs_sensitivities
            else self._initial_phase}
