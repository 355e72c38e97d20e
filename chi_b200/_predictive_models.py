"""Predictive models: forward simulation + measurement noise.

Host-side mirror of ``chi.PredictiveModel`` and
``chi.PopulationPredictiveModel`` (chi/_predictive_models.py:347-809,
812-1118).  The reference loops over virtual
patients in Python and runs one CVODES solve per patient
(chi/_predictive_models.py:1030-1033); here all patients are integrated in one
launch of the forward-only kernel (``SBMLModel.simulate_batch``) and the
individual parameters come from the population kernels.  Measurement noise is
drawn on the host with numpy's generator in exactly the order the reference
consumes it (patient-major, then output, then time), so seeded results
reproduce the reference's stream: bit-identical for the Gaussian-type error
models, within one ulp for the log-normal model (numpy's vectorised ``exp``
versus the C library's).
"""
import copy

import numpy as np

from chi_b200._error_models import (
    ErrorModel, GaussianErrorModel, LogNormalErrorModel,
    ConstantAndMultiplicativeGaussianErrorModel,
    MultiplicativeGaussianErrorModel)
from chi_b200._mechanistic_models import MechanisticModel, SBMLModel
from chi_b200._population_models import PopulationModel


def _noise(error_model, sigma, outputs, rng):
    """Vectorised counterpart of ``error_model.sample(sigma_p, outputs_p,
    n_samples=1, seed=rng)`` for every patient p.  ``sigma``: (B, n_sigma),
    ``outputs``: (B, n_times).  Consumes ``rng`` as the per-patient calls
    would: one block of draws per patient."""
    n_b, n_t = outputs.shape
    if isinstance(error_model, GaussianErrorModel):
        z = rng.standard_normal((n_b, n_t))
        return outputs + sigma[:, 0:1] * z
    if isinstance(error_model, LogNormalErrorModel):
        z = rng.standard_normal((n_b, n_t))
        s = sigma[:, 0:1]
        return outputs * np.exp(-s ** 2 / 2 + s * z)
    if isinstance(error_model, ConstantAndMultiplicativeGaussianErrorModel):
        z = rng.standard_normal((n_b, 2, n_t))
        return outputs + sigma[:, 0:1] * z[:, 0] \
            + outputs * (sigma[:, 1:2] * z[:, 1])
    if isinstance(error_model, MultiplicativeGaussianErrorModel):
        z = rng.standard_normal((n_b, n_t))
        return outputs + outputs * (sigma[:, 0:1] * z)
    raise TypeError('Unknown error model.')


def name_parameters(mechanistic_model, error_models):
    """Parameter names of a model of the data-generating process:
    mechanistic parameters, then every output's error parameters (prefixed
    with the output name when there are several outputs;
    chi/_log_pdfs.py:885-930, chi/_predictive_models.py:422-457)."""
    outputs = mechanistic_model.outputs()
    names = list(mechanistic_model.parameters())
    for output, error_model in zip(outputs, error_models):
        error_model.set_parameter_names(None)
        if len(outputs) > 1:
            error_model.set_parameter_names(
                [output + ' ' + n for n in error_model.get_parameter_names()])
        names += error_model.get_parameter_names()
    return names


def long_frame(measurements, times, output_names, ids=None):
    """Measurements ``(n_outputs, n_times, n_series)`` as the long-format
    table chi returns: ID | Time | Observable | Value, output-major, then
    time, then series."""
    import pandas as pd
    n_outputs, n_times, n_series = measurements.shape
    if ids is None:
        ids = np.arange(1, n_series + 1)
    return pd.DataFrame({
        'ID': np.tile(ids, n_outputs * n_times),
        'Time': np.tile(np.repeat(times, n_series), n_outputs),
        'Observable': np.repeat(list(output_names), n_times * n_series),
        'Value': measurements.reshape(-1)})


def append_regimen(frame, regimen, ids):
    """Adds the dose table once per ID."""
    import pandas as pd
    if regimen is None:
        return frame
    parts = [frame]
    for _id in ids:
        part = regimen.copy()
        part['ID'] = _id
        parts.append(part)
    return pd.concat(parts)


class PredictiveModel(object):
    """A mechanistic model plus one error model per output
    (interface of chi/_predictive_models.py:347-809)."""

    def __init__(self, mechanistic_model, error_models, outputs=None):
        if not isinstance(mechanistic_model, MechanisticModel):
            raise TypeError(
                'The mechanistic model has to be an instance of a '
                'chi.MechanisticModel.')
        if not isinstance(error_models, list):
            error_models = [error_models]
        if not all(isinstance(em, ErrorModel) for em in error_models):
            raise TypeError(
                'All error models have to be instances of a '
                'chi.ErrorModel.')
        model = mechanistic_model.copy()
        if outputs is not None:
            model.set_outputs(outputs)
        if len(error_models) != model.n_outputs():
            raise ValueError(
                'Wrong number of error models. One error model has to be '
                'provided for each mechanistic error model.')
        self._mechanistic_model = model
        self._error_models = [copy.deepcopy(em) for em in error_models]
        self._parameter_names = name_parameters(model, self._error_models)
        self._n_parameters = len(self._parameter_names)

    def get_dosing_regimen(self, final_time=None):
        """Dose table (Time | Duration | Dose) of the model's regimen up to
        ``final_time``, or ``None`` (chi/_predictive_models.py:520-603): a
        periodic event is listed dose by dose, an indefinite one up to
        ``final_time`` (one dose when that is open too)."""
        import pandas as pd
        regimen = getattr(
            self._mechanistic_model, 'dosing_regimen', lambda: None)()
        if regimen is None:
            return None
        horizon = np.inf if final_time is None else final_time
        rows = []
        for event in regimen.events():
            duration = event.duration()
            amount = event.level() * duration
            first, every, count = \
                event.start(), event.period(), event.multiplier()
            if first > horizon:
                continue
            if every == 0:
                starts = [first]
            else:
                if count == 0:
                    count = int(abs(horizon) // every) \
                        if np.isfinite(horizon) else 1
                starts = [first + k * every for k in range(count)
                          if first + k * every <= horizon]
            rows += [(t, duration, amount) for t in starts]
        if not rows:
            return None
        return pd.DataFrame(rows, columns=['Time', 'Duration', 'Dose'])

    def get_n_outputs(self):
        return self._mechanistic_model.n_outputs()

    def get_output_names(self):
        return self._mechanistic_model.outputs()

    def get_parameter_names(self):
        return list(self._parameter_names)

    def get_submodels(self):
        return {'Mechanistic model': self._mechanistic_model,
                'Error models': list(self._error_models)}

    def n_parameters(self):
        return self._n_parameters

    def _split(self, parameters):
        n_mech = self._mechanistic_model.n_parameters()
        return parameters[..., :n_mech], parameters[..., n_mech:]

    def sample(
            self, parameters, times, n_samples=None, seed=None,
            return_df=True, include_regimen=False, *args, **kwargs):
        """``n_samples`` noisy measurement series of one individual: one
        solve, then every error model draws for its output
        (chi/_predictive_models.py:659-762)."""
        parameters = np.asarray(parameters, dtype=float)
        if len(parameters) != self._n_parameters:
            raise ValueError(
                'The length of parameters does not match n_parameters.')
        psi, sigma = self._split(parameters)
        times = np.sort(times)
        ybar = self._mechanistic_model.simulate(psi, times)
        if isinstance(ybar, tuple):
            ybar = ybar[0]
        n_series = 1 if n_samples is None else n_samples
        measured = np.empty((len(ybar), len(times), n_series))
        first = 0
        for k, error_model in enumerate(self._error_models):
            last = first + error_model.n_parameters()
            measured[k] = error_model.sample(
                parameters=sigma[first:last], model_output=ybar[k],
                n_samples=n_series, seed=seed)
            first = last
        if not return_df:
            return measured
        ids = np.arange(1, n_series + 1)
        frame = long_frame(measured, times, self.get_output_names(), ids)
        if include_regimen:
            frame = append_regimen(
                frame, self.get_dosing_regimen(np.max(times)), ids)
        return frame

    def set_dosing_regimen(
            self, dose, start=0, duration=0.01, period=None, num=None):
        setter = getattr(self._mechanistic_model, 'set_dosing_regimen', None)
        if setter is None:
            raise AttributeError(
                'The mechanistic model does not support to set dosing '
                'regimens. This may be because the underlying '
                'chi.MechanisticModel is a '
                'chi.PharmacodynamicModel.')
        setter(dose, start, duration, period, num)

    # -- batched forward path ----------------------------------------------
    def _sample_patients(self, patients, times, rng):
        """One virtual measurement series per patient (rows of `patients`):
        a single batched solve plus vectorised noise in the reference's
        stream order.  Returns (n_outputs, n_times, n_patients)."""
        model = self._mechanistic_model
        if not isinstance(model, SBMLModel):
            raise TypeError(
                'Batched sampling needs a chi_b200.SBMLModel / PKPDModel.')
        mech, err = self._split(np.asarray(patients, dtype=float))
        outputs, status = model.simulate_batch(mech, times)
        if np.any(status != 0):
            raise RuntimeError(
                'The integrator failed for %d virtual patients.'
                % int(np.sum(status != 0)))
        n_b, n_out, n_t = outputs.shape
        # per patient the reference draws output after output; build the
        # blocks per output and interleave them patient-major
        kinds = self._error_models
        draws = [2 if isinstance(
            em, ConstantAndMultiplicativeGaussianErrorModel) else 1
            for em in kinds]
        if len(kinds) == 1:
            sig = err[:, :kinds[0].n_parameters()]
            noisy = _noise(kinds[0], sig, outputs[:, 0, :], rng)[:, None, :]
        else:
            # draw the whole patient block at once, then slice per output
            z = rng.standard_normal((n_b, sum(draws) * n_t))
            noisy = np.empty_like(outputs)
            col = 0
            start = 0
            for o, em in enumerate(kinds):
                end = start + em.n_parameters()
                block = z[:, col:col + draws[o] * n_t]

                class _Replay(object):
                    def __init__(self, values):
                        self._v = values

                    def standard_normal(self, shape):
                        return self._v.reshape(shape)
                noisy[:, o, :] = _noise(
                    em, err[:, start:end], outputs[:, o, :], _Replay(block))
                col += draws[o] * n_t
                start = end
        return noisy.transpose(1, 2, 0).copy()


class PopulationPredictiveModel(PredictiveModel):
    """chi/_predictive_models.py:812-1118."""

    def __init__(self, predictive_model, population_model):
        if not isinstance(predictive_model, PredictiveModel):
            raise TypeError(
                'The predictive model has to be an instance of '
                'chi.PredictiveModel.')
        if not isinstance(population_model, PopulationModel):
            raise TypeError(
                'The population model has to be an instance of '
                'chi.PopulationModel.')
        n_parameters = predictive_model.n_parameters()
        if population_model.n_dim() != n_parameters:
            raise ValueError(
                'The dimension of the population model has to be the same as '
                'the number of predictive model parameters.')
        self._predictive_model = predictive_model
        self._population_model = population_model

    def get_dosing_regimen(self, final_time=None):
        return self._predictive_model.get_dosing_regimen(final_time)

    def get_n_outputs(self):
        return self._predictive_model.get_n_outputs()

    def get_output_names(self):
        return self._predictive_model.get_output_names()

    def get_parameter_names(self):
        return self._population_model.get_parameter_names()

    def n_parameters(self):
        return self._population_model.n_parameters()

    def sample(
            self, parameters, times, n_samples=None, seed=None,
            return_df=True, include_regimen=False, covariates=None):
        """chi/_predictive_models.py:951-1080."""
        if not n_samples:
            n_samples = 1
        n_samples = int(n_samples)
        parameters = np.asarray(parameters, dtype=float)
        if len(parameters) != self.n_parameters():
            raise ValueError(
                'The length of parameters does not match n_parameters.')
        pop = self._population_model
        if pop.n_covariates() > 0:
            covariates = np.asarray(covariates, dtype=float)
            if covariates.ndim == 1:
                covariates = covariates[np.newaxis, :]
            n_s, n_c = covariates.shape
            if n_c != pop.n_covariates():
                raise ValueError(
                    'Provided covariates do not match the number of '
                    'covariates.')
            if (n_s > 1) and (n_s != n_samples):
                raise ValueError(
                    'Provided covariates cannot be broadcasted to number of '
                    'samples.')
        rng = np.random.default_rng(seed)
        kwargs = {}
        if pop.n_covariates() > 0:
            kwargs['covariates'] = covariates
        n_ids = pop.n_ids()
        pop.set_n_ids(n_samples)
        try:
            patients = pop.sample(
                parameters=parameters, n_samples=n_samples, seed=rng,
                **kwargs)
            cov = None
            if pop.n_covariates() > 0:
                cov = np.broadcast_to(
                    covariates, (n_samples, pop.n_covariates()))
            patients = pop.compute_individual_parameters(
                parameters=parameters, eta=patients, covariates=cov)
        finally:
            pop.set_n_ids(n_ids)
        times = np.sort(times)
        measurements = self._predictive_model._sample_patients(
            patients, times, rng)
        if not return_df:
            return measurements
        import pandas as pd
        ids = np.arange(1, n_samples + 1)
        frame = long_frame(
            measurements, times, self._predictive_model.get_output_names(),
            ids)
        if pop.n_covariates() > 0:
            # one row per individual and covariate, without a time
            cov = np.broadcast_to(covariates, (n_samples, pop.n_covariates()))
            frame = pd.concat([frame] + [
                pd.DataFrame({'ID': ids, 'Time': np.nan, 'Observable': name,
                              'Value': cov[:, k]})
                for k, name in enumerate(pop.get_covariate_names())])
        if include_regimen:
            frame = append_regimen(
                frame, self.get_dosing_regimen(np.max(times)), ids)
        return frame

    def set_dosing_regimen(
            self, dose, start=0, duration=0.01, period=None, num=None):
        self._predictive_model.set_dosing_regimen(
            dose, start, duration, period, num)
