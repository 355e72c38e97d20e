"""Population-filter log-posterior on the batched integrator.

Interface of ``chi.PopulationFilterLogPosterior`` (chi/_log_pdfs.py:1325-2075;
SURVEY.md section 8f, rank 4).  One evaluation simulates ``n_samples``
virtual individuals and scores their noisy outputs against snapshot data with
a :class:`chi_b200.PopulationFilter`.  The reference runs one CVODES solve
per virtual individual in a Python loop (chi/_log_pdfs.py:1599-1602,
1825-1828); here

* the population kernels turn ``(theta, eta)`` into the individuals'
  parameters and give the population density and its adjoint
  (``chi_b200_population_eval``),
* ONE launch of the integrator (``SBMLModel.simulate_batch``) returns all
  virtual individuals' outputs -- and, for ``evaluateS1``, their
  sensitivities,
* the filter's closed-form reductions run on the host
  (chi_b200/_population_filters.py).

Parameter vector: ``[theta (population parameters, then one sigma per
observable unless fixed) | eta of the virtual individuals (hierarchical
dimensions only, individual-major) | epsilon (n_samples, n_observables,
n_times)]``.
"""
import copy
import warnings

import numpy as np

from chi_b200._log_pdfs import (
    HierarchicalLogPosterior, _check_prior, _chi_base, _LogPDFBase)
from chi_b200._mechanistic_models import MechanisticModel, SBMLModel
from chi_b200._population_filters import PopulationFilter
from chi_b200._population_models import PopulationModel


class PopulationFilterLogPosterior(
        _chi_base('PopulationFilterLogPosterior', HierarchicalLogPosterior)):
    """See the module docstring and chi/_log_pdfs.py:1325-1420 for the
    meaning of the arguments."""

    def __init__(
            self, population_filter, times, mechanistic_model,
            population_model, log_prior, sigma=None, error_on_log_scale=False,
            n_samples=100, covariates=None):
        if not isinstance(population_filter, PopulationFilter):
            raise TypeError(
                'The population filter has to be an instance of '
                'chi.PopulationFilter.')
        self._filter = copy.deepcopy(population_filter)
        self._n_times = population_filter.n_times()
        self._n_observables = population_filter.n_observables()
        times = np.asarray(times, dtype=float).reshape(-1)
        if len(times) != len(np.unique(times)):
            raise ValueError(
                'The measurement times in times have to be unique.')
        if len(times) != self._n_times:
            raise ValueError(
                'The length of times does not match the time dimension of '
                'observations.')
        self._filter.sort_times(np.argsort(times))
        self._times = np.sort(times)
        if not isinstance(mechanistic_model, MechanisticModel):
            raise TypeError(
                'The mechanistic model has to be an instance of '
                'chi.MechanisticModel.')
        if mechanistic_model.n_outputs() != self._n_observables:
            raise ValueError(
                'The number of mechanistic model outputs does not match the '
                'number of observables.')
        if not isinstance(mechanistic_model, SBMLModel):
            raise TypeError(
                'The device evaluation needs a chi_b200.SBMLModel / PKPDModel '
                'mechanistic model.')
        self._mechanistic_model = mechanistic_model.copy()
        if not isinstance(population_model, PopulationModel):
            raise TypeError(
                'The population model has to be an instance of '
                'chi.PopulationModel.')
        if population_model.n_dim() != self._mechanistic_model.n_parameters():
            raise ValueError(
                'The number of population model dimensions does not match the '
                'number of mechanistic model parameters.')
        self._population_model = copy.deepcopy(population_model)
        n_samples = int(n_samples)
        if n_samples <= 0:
            raise ValueError(
                'The number of samples of the population filter has to be '
                'greater than zero.')
        self._n_samples = n_samples
        self._covariates = None
        n_cov = self._population_model.n_covariates()
        if n_cov > 0:
            covariates = np.array(covariates, dtype=float)
            if covariates.ndim == 1:
                covariates = covariates[np.newaxis, :]
            if covariates.shape[1] != n_cov:
                raise ValueError(
                    'Invalid covariates. The provided covariates do not match '
                    'the number of covariates.')
            try:
                covariates = np.broadcast_to(covariates, (n_samples, n_cov))
            except ValueError:
                raise ValueError(
                    'Invalid covariates. The provided covariates cannot be '
                    'broadcasted to the shape (n_samples, n_cov).')
            self._covariates = np.ascontiguousarray(covariates)
        self._population_model.set_n_ids(n_samples)
        self._population_model.set_dim_names(mechanistic_model.parameters())
        self._hier_dims = self._population_model._hier_columns()
        self._n_hdim = len(self._hier_dims)
        self._n_pop = self._population_model.n_parameters()
        self._n_top = self._n_pop
        self._sigma = None
        if sigma is not None:
            try:
                sigma = list(sigma)
            except TypeError:
                sigma = [sigma]
            if len(sigma) != self._n_observables:
                raise ValueError(
                    'One sigma for each observable has to provided.')
            sigma = np.array([float(s) for s in sigma])
            if np.any(sigma < 0):
                raise ValueError(
                    'The elements of sigma have to be greater or equal '
                    'to zero.')
            self._sigma = sigma.reshape(1, self._n_observables, 1)
        names = self._population_model.get_parameter_names()
        if self._sigma is None:
            names = names + [
                'Sigma %s' % name for name in mechanistic_model.outputs()]
            self._n_top += self._n_observables
        _check_prior(log_prior)
        if log_prior.n_parameters() != self._n_top:
            raise ValueError(
                'The dimensionality of the log-prior does not match the '
                'number of population parameters. The population parameters '
                'are <' + str(names) + '>.')
        self._log_prior = log_prior
        self._error_on_log_scale = bool(error_on_log_scale)
        self._top_names = names
        self._end_bottom = self._n_top + n_samples * self._n_hdim
        self._n_parameters = self._end_bottom + \
            n_samples * self._n_times * self._n_observables
        self._n_bottom = self._n_parameters - self._n_top

    # -- pieces of an evaluation -------------------------------------------------
    def _split(self, parameters):
        parameters = np.asarray(parameters, dtype=float).reshape(-1)
        pop = parameters[:self._n_pop]
        sigma = self._sigma
        if sigma is None:
            sigma = parameters[self._n_pop:self._n_top].reshape(
                1, self._n_observables, 1)
        eta = np.zeros((self._n_samples, self._population_model.n_dim()))
        eta[:, self._hier_dims] = parameters[
            self._n_top:self._end_bottom].reshape(
                self._n_samples, self._n_hdim)
        epsilon = parameters[self._end_bottom:].reshape(
            self._n_samples, self._n_observables, self._n_times)
        return parameters, pop, sigma, eta, epsilon

    def _noise_constant(self, epsilon):
        # standard-normal density of the noise realisations
        # (chi/_log_pdfs.py:1577-1579: n_samples * n_observables terms)
        return -self._n_samples * self._n_observables * np.log(2 * np.pi) / 2 \
            - np.sum(epsilon ** 2) / 2

    def _failure(self, status):
        warnings.warn(
            'An error occured while solving the mechanistic model: \n'
            'integrator status %s.\n A score of -infinity is returned.'
            % str(status[status != 0][0]), RuntimeWarning)

    def __call__(self, parameters):
        """chi/_log_pdfs.py:1529-1622."""
        parameters, pop, sigma, eta, epsilon = self._split(parameters)
        score = self._log_prior(parameters[:self._n_top])
        if np.isinf(score):
            return score
        pm = self._population_model
        o_eta, psi, pop_score, _ = pm._device_eval(
            pop, eta, None, self._covariates, False)
        score += pop_score
        if np.isinf(score) or np.isnan(score):
            return -np.inf
        score += self._noise_constant(epsilon)
        y, status = self._mechanistic_model.simulate_batch(psi, self._times)
        if np.any(status != 0):
            self._failure(status)
            return -np.inf
        if self._error_on_log_scale:
            y = y * np.exp(sigma * epsilon)
        else:
            y = y + sigma * epsilon
        return score + self._filter.compute_log_likelihood(y)

    def evaluateS1(self, parameters):
        """chi/_log_pdfs.py:1766-1883."""
        parameters, pop, sigma, eta, epsilon = self._split(parameters)
        sens = np.zeros(self._n_parameters)
        score, sens[:self._n_top] = self._log_prior.evaluateS1(
            parameters[:self._n_top])
        if np.isinf(score):
            return score, sens
        score += self._noise_constant(epsilon)
        sens[self._end_bottom:] = -epsilon.reshape(-1)
        pm = self._population_model
        psi = pm._device_eval(pop, eta, None, self._covariates, False)[1]
        y, status, dy = self._mechanistic_model.simulate_batch(
            psi, self._times, sensitivities=True)
        if np.any(status != 0):
            self._failure(status)
            return -np.inf, sens
        if self._error_on_log_scale:
            scale = np.exp(sigma * epsilon)
            y = y * scale
        else:
            scale = 1.0
            y = y + sigma * epsilon
        s, ds_y = self._filter.compute_sensitivities(simulated_obs=y)
        score += s
        # dy: (n_samples, n_times, n_observables, n_parameters)
        ds_dpsi = np.einsum('sot,stop->sp', ds_y * scale, dy)
        if self._error_on_log_scale:
            sens[self._end_bottom:] += (ds_y * y * sigma).reshape(-1)
        else:
            sens[self._end_bottom:] += (
                ds_y * sigma * np.ones_like(epsilon)).reshape(-1)
        if self._sigma is None:
            d_sigma = ds_y * epsilon * (y if self._error_on_log_scale else 1.0)
            sens[self._n_pop:self._n_top] += np.sum(d_sigma, axis=(0, 2))
        # population density and the adjoint of (theta, eta) -> psi, in the
        # hierarchical layout [d eta (hierarchical dims) | d theta]
        _, _, pop_score, grad = pm._device_eval(
            pop, eta, ds_dpsi, self._covariates, True)
        score += pop_score
        n_b = self._n_samples * self._n_hdim
        sens[:self._n_pop] += grad[n_b:]
        if np.isinf(score) or np.isnan(score):
            return -np.inf, sens
        sens[self._n_top:self._end_bottom] = grad[:n_b]
        return score, sens

    # -- bookkeeping ---------------------------------------------------------------
    def get_log_likelihood(self):
        """The population filter takes the place of the log-likelihood."""
        return copy.deepcopy(self._filter)

    def get_log_prior(self):
        return self._log_prior

    def get_population_model(self):
        return self._population_model

    def get_id(self, unique=False):
        labels = ['Sim. %d' % (k + 1) for k in range(self._n_samples)]
        if unique:
            return labels
        ids = [None] * self._n_top
        for label in labels:
            ids += [label] * self._n_hdim
        for label in labels:
            ids += [label] * (self._n_observables * self._n_times)
        return ids

    def get_parameter_names(
            self, exclude_bottom_level=False, include_ids=False):
        names = list(self._top_names)
        if not exclude_bottom_level:
            mech = self._mechanistic_model.parameters()
            names += [mech[d] for d in self._hier_dims] * self._n_samples
            eps = []
            for output in self._mechanistic_model.outputs():
                eps += ['%s Epsilon time %d' % (output, t + 1)
                        for t in range(self._n_times)]
            names += eps * self._n_samples
        if include_ids:
            ids = self.get_id()
            names = [(ids[k] + ' ' + n) if ids[k] else n
                     for k, n in enumerate(names)]
        return names

    def n_parameters(self, exclude_bottom_level=False):
        return self._n_top if exclude_bottom_level else self._n_parameters

    def n_samples(self):
        return self._n_samples

    def sample_initial_parameters(self, n_samples=1, seed=None):
        """chi/_log_pdfs.py:2003-2075."""
        n_samples = int(n_samples)
        if n_samples <= 0:
            raise ValueError('The number of samples has to be at least 1.')
        out = np.empty((n_samples, self._n_parameters))
        np.random.seed(seed)
        out[:, :self._n_top] = self._log_prior.sample(n_samples)
        rng = np.random.default_rng(None if seed is None else seed + 1)
        for k in range(n_samples):
            kwargs = {}
            if self._covariates is not None:
                kwargs['covariates'] = self._covariates
            psi = self._population_model.sample(
                parameters=out[k, :self._n_pop], n_samples=self._n_samples,
                seed=rng, **kwargs)
            out[k, self._n_top:self._end_bottom] = \
                psi[:, self._hier_dims].reshape(-1)
        out[:, self._end_bottom:] = rng.normal(
            loc=0, scale=1, size=(
                n_samples,
                self._n_samples * self._n_times * self._n_observables))
        return out
