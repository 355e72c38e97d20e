"""Predictive models averaged over a parameter distribution.

Interface of ``chi.PosteriorPredictiveModel`` and ``chi.PriorPredictiveModel``
(chi/_predictive_models.py:17-344, 1121-1230).
The reference draws one parameter set per virtual measurement series and runs
one CVODES solve for each (chi/_predictive_models.py:302-331).  Here ALL
parameter sets are drawn first, then integrated in ONE batched forward launch
(``PredictiveModel._sample_patients``), then the measurement noise is added.

Randomness: the reference interleaves the draw of a parameter set with the
noise draws of the previous series on one generator; here the parameter sets
are drawn first and the noise afterwards (same generator, same
distributions).  Seeded results therefore agree with the reference in
distribution, not number by number.
"""
import numpy as np

from chi_b200._predictive_models import (
    PopulationPredictiveModel, PredictiveModel)

try:
    import xarray as _xr
    _Dataset = getattr(_xr, 'Dataset', None)
except Exception:  # pragma: no cover
    _xr = None
    _Dataset = None


class AveragedPredictiveModel(object):
    """chi/_predictive_models.py:17-127."""

    def __init__(self, predictive_model):
        if not isinstance(predictive_model, PredictiveModel):
            raise ValueError(
                'The provided predictive model has to be an instance of a '
                'chi.PredictiveModel.')
        self._predictive_model = predictive_model

    def get_dosing_regimen(self, final_time=None):
        return self._predictive_model.get_dosing_regimen(final_time)

    def get_n_outputs(self):
        return self._predictive_model.get_n_outputs()

    def get_output_names(self):
        return self._predictive_model.get_output_names()

    def get_predictive_model(self):
        return self._predictive_model

    def set_dosing_regimen(
            self, dose, start=0, duration=0.01, period=None, num=None):
        self._predictive_model.set_dosing_regimen(
            dose, start, duration, period, num)

    def sample(self, times, n_samples=None, seed=None,
               include_regimen=False):
        raise NotImplementedError

    # -- shared machinery ---------------------------------------------------------
    def _measure(self, parameter_sets, times, rng, covariates):
        """One measurement series per row of ``parameter_sets``; returns
        ``(n_outputs, n_times, n_sets)``."""
        model = self._predictive_model
        if isinstance(model, PopulationPredictiveModel):
            # one virtual individual per population-parameter set: the
            # individual's parameters come from the population model, the
            # solves of all of them share one launch
            pop = model._population_model
            patients = np.empty((len(parameter_sets), pop.n_dim()))
            n_ids = pop.n_ids()
            pop.set_n_ids(1)
            try:
                for k, theta in enumerate(parameter_sets):
                    kwargs = {}
                    cov = None
                    if pop.n_covariates() > 0:
                        cov = np.asarray(covariates, dtype=float).reshape(
                            -1, pop.n_covariates())
                        cov = cov[k % len(cov)][np.newaxis, :]
                        kwargs['covariates'] = cov
                    eta = pop.sample(
                        parameters=theta, n_samples=1, seed=rng, **kwargs)
                    patients[k] = pop.compute_individual_parameters(
                        parameters=theta, eta=eta, covariates=cov)[0]
            finally:
                pop.set_n_ids(n_ids)
            return model._predictive_model._sample_patients(
                patients, times, rng)
        return model._sample_patients(parameter_sets, times, rng)

    def _frame(self, measurements, times, include_regimen):
        import pandas as pd
        n_outputs, n_times, n_sets = measurements.shape
        names = self.get_output_names()
        ids = np.arange(1, n_sets + 1)
        frame = pd.DataFrame({
            'ID': np.tile(np.repeat(ids, n_times), n_outputs),
            'Time': np.tile(times, n_sets * n_outputs),
            'Observable': np.repeat(names, n_sets * n_times),
            'Value': measurements.transpose(0, 2, 1).reshape(-1)})
        if include_regimen:
            regimen = self.get_dosing_regimen(np.max(times))
            if regimen is not None:
                frame = pd.concat([frame, regimen])
        return frame


class PosteriorPredictiveModel(AveragedPredictiveModel):
    """chi/_predictive_models.py:130-344.

    :param posterior_samples: an ``xarray.Dataset`` with dimensions
        ``(chain, draw[, individual])`` as produced by chi's sampling
        controller, or -- xarray is an optional dependency here -- a mapping
        ``{parameter name: array of shape (n_chains, n_draws) or (n_chains,
        n_draws, n_ids)}`` with an optional entry ``'individual'`` listing
        the ids.
    """

    def __init__(self, predictive_model, posterior_samples, param_map=None):
        super(PosteriorPredictiveModel, self).__init__(predictive_model)
        self._ids = None
        if _Dataset is not None and isinstance(posterior_samples, _Dataset):
            dims = sorted(list(posterior_samples.dims))
            for dim in ['chain', 'draw']:
                if dim not in dims:
                    raise ValueError(
                        'The posterior samples must have the dimensions '
                        '(chain, draw, individual). The current dimensions '
                        'are <' + str(dims) + '>.')
            table = {str(k): np.asarray(v.values)
                     for k, v in posterior_samples.data_vars.items()}
            if 'individual' in dims:
                self._ids = [str(i) for i in
                             posterior_samples.individual.values]
        elif hasattr(posterior_samples, 'items'):
            table = {str(k): np.asarray(v, dtype=float)
                     for k, v in posterior_samples.items()
                     if k != 'individual'}
            if 'individual' in posterior_samples:
                self._ids = [str(i) for i in posterior_samples['individual']]
        else:
            raise TypeError(
                'The posterior samples have to be a xarray.Dataset.')
        try:
            param_map = dict(param_map or {})
        except (TypeError, ValueError):
            raise ValueError(
                'The parameter map has to be convertable to a python '
                'dictionary.')
        names = [param_map.get(n, n)
                 for n in self._predictive_model.get_parameter_names()]
        for name in names:
            if name not in table:
                raise ValueError(
                    'The parameter <' + str(name) + '> cannot be found '
                    'in the posterior.')
        self._parameter_names = names
        self._table = table

    def _posterior_matrix(self, individual):
        """Rows: pooled (chain, draw) samples; columns: model parameters.
        Draws with NaN (chains of unequal length) are dropped."""
        columns = []
        for name in self._parameter_names:
            values = self._table[name]
            if values.ndim == 3:
                values = values[:, :, self._ids.index(individual)]
            columns.append(values.reshape(-1))
        matrix = np.stack(columns, axis=1)
        return matrix[~np.any(np.isnan(matrix), axis=1)]

    def sample(
            self, times, n_samples=None, individual=None, seed=None,
            include_regimen=False, covariates=None, return_df=True):
        n_samples = int(n_samples) if n_samples is not None else 1
        if isinstance(self._predictive_model, PopulationPredictiveModel) \
                or self._ids is None:
            individual = None
        else:
            if individual is None:
                individual = self._ids[0]
            if str(individual) not in self._ids:
                raise ValueError(
                    'The individual <' + str(individual) + '> could not be '
                    'found in the ID column.')
            individual = str(individual)
        times = np.sort(np.asarray(times, dtype=float))
        rng = np.random.default_rng(seed=seed)
        posterior = self._posterior_matrix(individual)
        sets = posterior[rng.integers(0, len(posterior), size=n_samples)]
        measurements = self._measure(sets, times, rng, covariates)
        if not return_df:
            return measurements
        return self._frame(measurements, times, include_regimen)


class PriorPredictiveModel(AveragedPredictiveModel):
    """chi/_predictive_models.py:1121-1230: parameter sets from a
    ``pints.LogPrior``."""

    def __init__(self, predictive_model, log_prior):
        super(PriorPredictiveModel, self).__init__(predictive_model)
        if not all(hasattr(log_prior, a) for a in ('sample', 'n_parameters')):
            raise ValueError(
                'The provided log-prior has to be an instance of a '
                'pints.LogPrior.')
        if predictive_model.n_parameters() != log_prior.n_parameters():
            raise ValueError(
                'The dimension of the log-prior has to be the same as the '
                'number of parameters of the predictive model.')
        self._log_prior = log_prior

    def sample(self, times, n_samples=None, seed=None, include_regimen=False,
               covariates=None, return_df=True):
        n_samples = int(n_samples) if n_samples is not None else 1
        times = np.sort(np.asarray(times, dtype=float))
        np.random.seed(seed)
        sets = np.asarray(self._log_prior.sample(n_samples), dtype=float)
        rng = np.random.default_rng(seed=seed)
        measurements = self._measure(sets, times, rng, covariates)
        if not return_df:
            return measurements
        return self._frame(measurements, times, include_regimen)
