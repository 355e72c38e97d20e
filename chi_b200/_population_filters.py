"""Population filters (interface of ``chi/_population_filters.py``).

A population filter scores simulated measurements ``y[s, o, t]`` of virtual
individuals against snapshot data ``obs[i, o, t]`` (missing values: NaN) by
summary statistics over the simulated population at each (observable, time)
point.  It is the adjacent consumer of the batched forward solve (SURVEY.md
section 8f, rank 4): the ODE solves of the virtual individuals -- with
sensitivities for ``evaluateS1`` -- are one launch of the integrator
(``SBMLModel.simulate_batch``); what is left for the filter are closed-form
reductions over arrays of a few thousand entries, done here with numpy on the
host.  The formulas follow the reference's definitions (file:line in each
class); the gradients are written out from them.

All filters return ``-inf`` (and an unspecified gradient) when every
observation is missing, as the reference does.
"""
import numpy as np

try:
    import chi as _chi
    _FilterBase = _chi.PopulationFilter
except Exception:  # pragma: no cover
    _chi = None
    _FilterBase = object

_LOG_2PI = np.log(2 * np.pi)


def _logsumexp(a, axis):
    m = np.max(a, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    out = np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True)) + m
    return np.squeeze(out, axis=axis)


def _softmax(a, axis):
    m = np.max(a, axis=axis, keepdims=True)
    e = np.exp(a - m)
    return e / np.sum(e, axis=axis, keepdims=True)


class PopulationFilter(_FilterBase):
    """Base class, chi/_population_filters.py:11-115.

    :param observations: measurements of shape
        ``(n_ids, n_observables, n_times)``, NaN where missing.
    """

    def __init__(self, observations):
        observations = np.asarray(observations, dtype=float)
        if observations.ndim != 3:
            raise ValueError(
                'The observations must be of shape '
                '(n_ids, n_observables, n_times).')
        self._set_data(observations)

    def _set_data(self, observations):
        self._data = observations
        self._valid = ~np.isnan(observations)
        _, self._n_observables, self._n_times = observations.shape

    def _obs(self):
        """Observations with zeros at the missing entries and the validity
        mask (products with the mask drop the missing entries)."""
        return np.where(self._valid, self._data, 0.0), \
            self._valid.astype(float)

    def compute_log_likelihood(self, simulated_obs):
        raise NotImplementedError

    def compute_sensitivities(self, simulated_obs):
        raise NotImplementedError

    def n_observables(self):
        return self._n_observables

    def n_times(self):
        return self._n_times

    def sort_times(self, order):
        order = np.asarray(order)
        if len(order) != self._n_times:
            raise ValueError('Order has to be of length n_times.')
        if len(order) != len(np.unique(order)):
            raise ValueError('Order has to contain n_times unique elements.')
        self._set_data(self._data[..., order])


class _MomentFilter(PopulationFilter):
    """Gaussian density with the simulated population's mean and variance at
    every (observable, time) point; ``LOG`` selects the log scale."""
    LOG = False

    def _moment_terms(self, z):
        n_sim = len(z)
        mu = np.mean(z, axis=0)
        var = np.var(z, axis=0, ddof=1)
        obs, valid = self._obs()
        if self.LOG:
            with np.errstate(all='ignore'):
                obs = np.where(self._valid, np.log(
                    np.where(self._valid, self._data, 1.0)), 0.0)
        resid = (obs - mu) * valid
        count = np.sum(valid, axis=0)
        score = -0.5 * np.sum(
            count * (_LOG_2PI + np.log(var)) + np.sum(resid ** 2, axis=0) / var)
        if self.LOG:
            # the reference adds log(obs) INSIDE the halved sum
            # (chi/_population_filters.py:715-726)
            score -= 0.5 * np.sum(obs * valid)
        return n_sim, mu, var, resid, count, score

    def _evaluate(self, simulated_obs, with_grad):
        y = np.asarray(simulated_obs, dtype=float)
        if not np.any(self._valid):
            return (-np.inf, np.empty(y.shape)) if with_grad else -np.inf
        with np.errstate(all='ignore'):
            z = np.log(y) if self.LOG else y
        n_sim, mu, var, resid, count, score = self._moment_terms(z)
        if not with_grad:
            return float(score)
        d_mu = np.sum(resid, axis=0) / var
        d_var = np.sum(resid ** 2, axis=0) / var ** 2 - count / var
        grad = d_mu / n_sim + d_var * (z - mu) / (n_sim - 1)
        if self.LOG:
            grad = grad / y
        return float(score), grad

    def compute_log_likelihood(self, simulated_obs):
        return self._evaluate(simulated_obs, False)

    def compute_sensitivities(self, simulated_obs):
        return self._evaluate(simulated_obs, True)


class GaussianFilter(_MomentFilter):
    """chi/_population_filters.py:270-378."""
    LOG = False


class LogNormalFilter(_MomentFilter):
    """chi/_population_filters.py:663-780."""
    LOG = True


class _KDEFilter(PopulationFilter):
    """Gaussian kernel density estimate over the simulated population with
    the rule-of-thumb bandwidth ``bw^2 = (4 / 3 n)^(2/5) var``."""
    LOG = False

    def _evaluate(self, simulated_obs, with_grad):
        y = np.asarray(simulated_obs, dtype=float)
        n_sim = len(y)
        if not np.any(self._valid):
            return (-np.inf, np.empty(y.shape)) if with_grad else -np.inf
        with np.errstate(all='ignore'):
            z = np.log(y) if self.LOG else y
            obs = np.where(self._valid, np.log(np.where(
                self._valid, self._data, 1.0)), 0.0) if self.LOG \
                else np.where(self._valid, self._data, 0.0)
        valid = self._valid.astype(float)
        mean = np.mean(z, axis=0)
        var = np.var(z, axis=0, ddof=1)
        bw2 = (4.0 / 3.0 / n_sim) ** 0.4 * var
        # q[s, i, o, t]
        q = -(z[:, None] - obs[None]) ** 2 / bw2 / 2
        per_obs = _logsumexp(q, axis=0) - np.log(n_sim) - _LOG_2PI / 2 \
            - np.log(bw2) / 2
        score = float(np.sum(per_obs * valid))
        if np.isnan(score):
            return (-np.inf, np.empty(y.shape)) if with_grad else -np.inf
        if not with_grad:
            return score
        w = _softmax(q, axis=0)
        # d log bw^2 / d z_s
        dlog_bw2 = 2 * (z - mean) / (n_sim - 1) / var
        grad = np.sum(
            (w * (obs[None] - z[:, None]) / bw2
             - np.sum(w * q, axis=0, keepdims=True) * dlog_bw2[:, None]
             - dlog_bw2[:, None] / 2) * valid[None], axis=1)
        if self.LOG:
            grad = grad / y
        return score, grad

    def compute_log_likelihood(self, simulated_obs):
        return self._evaluate(simulated_obs, False)

    def compute_sensitivities(self, simulated_obs):
        return self._evaluate(simulated_obs, True)


class GaussianKDEFilter(_KDEFilter):
    """chi/_population_filters.py:381-510."""
    LOG = False


class LogNormalKDEFilter(_KDEFilter):
    """chi/_population_filters.py:783-915."""
    LOG = True

    def __init__(self, observations, bandwidth=None):
        super(LogNormalKDEFilter, self).__init__(observations)


class GaussianMixtureFilter(PopulationFilter):
    """Equal-weight mixture of ``n_kernels`` Gaussians, each with the mean and
    variance of its share of the simulated population
    (chi/_population_filters.py:513-660)."""

    def __init__(self, observations, n_kernels=2):
        super(GaussianMixtureFilter, self).__init__(observations)
        n_kernels = int(n_kernels)
        if n_kernels < 2:
            raise ValueError(
                'Invalid number of kernels. A Gaussian mixture filter expects '
                'at least 2 kernels.')
        self._n_kernels = n_kernels

    def _evaluate(self, simulated_obs, with_grad):
        y = np.asarray(simulated_obs, dtype=float)
        if len(y) % self._n_kernels > 0:
            raise ValueError(
                'Invalid simulated_obs. The number of simulated observations '
                'needs to be a multiple of the number of kernels.')
        if not np.any(self._valid):
            return (-np.inf, np.empty(y.shape)) if with_grad else -np.inf
        n_sim, n_obs, n_times = y.shape
        K = self._n_kernels
        n_per = n_sim // K
        yk = y.reshape(K, n_per, n_obs, n_times)
        mu = np.mean(yk, axis=1)                       # (K, o, t)
        var = np.var(yk, axis=1, ddof=1)
        obs, valid = self._obs()
        # q[k, i, o, t]
        resid = obs[None] - mu[:, None]
        q = -resid ** 2 / var[:, None] / 2 - np.log(var[:, None]) / 2
        per_obs = _logsumexp(q, axis=0) - np.log(K) - _LOG_2PI / 2
        score = float(np.sum(per_obs * valid))
        if not with_grad:
            return score
        w = _softmax(q, axis=0) * valid[None]
        d_mu = np.sum(w * resid / var[:, None], axis=1)            # (K, o, t)
        d_var = np.sum(w * (resid ** 2 / var[:, None] ** 2
                            - 1 / var[:, None]), axis=1)
        grad = d_mu[:, None] / n_per + \
            d_var[:, None] * (yk - mu[:, None]) / (n_per - 1)
        return score, grad.reshape(n_sim, n_obs, n_times)

    def compute_log_likelihood(self, simulated_obs):
        return self._evaluate(simulated_obs, False)

    def compute_sensitivities(self, simulated_obs):
        return self._evaluate(simulated_obs, True)


class ComposedPopulationFilter(PopulationFilter):
    """Different filters for consecutive blocks of time points
    (chi/_population_filters.py:118-267)."""

    def __init__(self, population_filters):
        for population_filter in population_filters:
            if not isinstance(population_filter, PopulationFilter):
                raise TypeError(
                    'All filters have to be instances '
                    'of chi.PopulationFilter.')
        n_observables = population_filters[0].n_observables()
        for population_filter in population_filters:
            if population_filter.n_observables() != n_observables:
                raise ValueError(
                    'All filters need to model the same number of '
                    'observables.')
        self._filters = list(population_filters)
        self._n_observables = n_observables
        self._n_times = int(np.sum([f.n_times() for f in self._filters]))
        # position of every input time point in the filters' concatenated
        # time axis (None: identical)
        self._to_filter_order = None
        self._from_filter_order = None

    def _blocks(self):
        start = 0
        for f in self._filters:
            yield f, slice(start, start + f.n_times())
            start += f.n_times()

    def compute_log_likelihood(self, simulated_obs):
        y = np.asarray(simulated_obs, dtype=float)
        if self._to_filter_order is not None:
            y = y[:, :, self._to_filter_order]
        return sum(f.compute_log_likelihood(y[:, :, block])
                   for f, block in self._blocks())

    def compute_sensitivities(self, simulated_obs):
        y = np.asarray(simulated_obs, dtype=float)
        if self._to_filter_order is not None:
            y = y[:, :, self._to_filter_order]
        score = 0.0
        grad = np.zeros(y.shape)
        for f, block in self._blocks():
            s, g = f.compute_sensitivities(y[:, :, block])
            score += s
            grad[:, :, block] = g
        if self._from_filter_order is not None:
            grad = grad[:, :, self._from_filter_order]
        return score, grad

    def sort_times(self, order):
        """The sub-filters hold their own data, so the order is remembered
        and applied to the simulated measurements instead."""
        order = np.asarray(order)
        if len(order) != self._n_times:
            raise ValueError('Order has to be of length n_times.')
        if len(order) != len(np.unique(order)):
            raise ValueError('Order has to contain n_times unique elements.')
        if np.all(order == np.arange(self._n_times)):
            return
        self._from_filter_order = np.copy(order)
        self._to_filter_order = np.argsort(order)
