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


class PredictiveModel(object):
    """chi/_predictive_models.py:347-809."""

    def __init__(self, mechanistic_model, error_models, outputs=None):
        if not isinstance(mechanistic_model, MechanisticModel):
            raise TypeError(
                'The mechanistic model has to be an instance of a '
                'chi.MechanisticModel.')
        error_models = \
            error_models if isinstance(error_models, list) else [error_models]
        for error_model in error_models:
            if not isinstance(error_model, ErrorModel):
                raise TypeError(
                    'All error models have to be instances of a '
                    'chi.ErrorModel.')
        mechanistic_model = mechanistic_model.copy()
        if outputs is not None:
            mechanistic_model.set_outputs(outputs)
        n_outputs = mechanistic_model.n_outputs()
        if len(error_models) != n_outputs:
            raise ValueError(
                'Wrong number of error models. One error model has to be '
                'provided for each mechanistic error model.')
        error_models = [
            copy.deepcopy(error_model) for error_model in error_models]
        self._mechanistic_model = mechanistic_model
        self._error_models = error_models
        self._set_error_model_parameter_names()
        self._set_number_and_parameter_names()

    def _set_error_model_parameter_names(self):
        for error_model in self._error_models:
            error_model.set_parameter_names(None)
        n_outputs = self._mechanistic_model.n_outputs()
        if n_outputs > 1:
            outputs = self._mechanistic_model.outputs()
            for output_id, error_model in enumerate(self._error_models):
                names = error_model.get_parameter_names()
                output = outputs[output_id]
                names = [output + ' ' + name for name in names]
                error_model.set_parameter_names(names)

    def _set_number_and_parameter_names(self):
        parameter_names = self._mechanistic_model.parameters()
        for error_model in self._error_models:
            parameter_names += error_model.get_parameter_names()
        self._parameter_names = parameter_names
        self._n_parameters = len(self._parameter_names)

    def get_dosing_regimen(self, final_time=None):
        """chi/_predictive_models.py:520-603: DataFrame with Time, Duration
        and Dose columns, or None."""
        import pandas as pd
        try:
            regimen = self._mechanistic_model.dosing_regimen()
        except AttributeError:
            return None
        if regimen is None:
            return regimen
        if final_time is None:
            final_time = np.inf
        frames = []
        for dose_event in regimen.events():
            dose_duration = dose_event.duration()
            dose_amount = dose_event.level() * dose_duration
            start_time = dose_event.start()
            period = dose_event.period()
            n_doses = dose_event.multiplier()
            if start_time > final_time:
                continue
            if period == 0:
                frames.append(pd.DataFrame({
                    'Time': [start_time], 'Duration': [dose_duration],
                    'Dose': [dose_amount]}))
                continue
            if n_doses == 0:
                n_doses = 1
                if np.isfinite(final_time):
                    n_doses = int(abs(final_time) // period)
            dose_times = np.array(
                [start_time + n * period for n in range(n_doses)])
            dose_times = dose_times[dose_times <= final_time]
            frames.append(pd.DataFrame({
                'Time': dose_times, 'Duration': dose_duration,
                'Dose': dose_amount}))
        if not frames or all(f.empty for f in frames):
            return None
        return pd.concat(frames)

    def get_n_outputs(self):
        return self._mechanistic_model.n_outputs()

    def get_output_names(self):
        return self._mechanistic_model.outputs()

    def get_parameter_names(self):
        return copy.copy(self._parameter_names)

    def get_submodels(self):
        return dict({
            'Mechanistic model': self._mechanistic_model,
            'Error models': list(self._error_models)})

    def n_parameters(self):
        return self._n_parameters

    def _split(self, parameters):
        n_mech = self._mechanistic_model.n_parameters()
        return parameters[..., :n_mech], parameters[..., n_mech:]

    def sample(
            self, parameters, times, n_samples=None, seed=None,
            return_df=True, include_regimen=False, *args, **kwargs):
        """chi/_predictive_models.py:659-762."""
        parameters = np.asarray(parameters, dtype=float)
        if len(parameters) != self._n_parameters:
            raise ValueError(
                'The length of parameters does not match n_parameters.')
        mechanistic_params, error_params = self._split(parameters)
        times = np.sort(times)
        outputs = self._mechanistic_model.simulate(mechanistic_params, times)
        if isinstance(outputs, tuple):
            outputs = outputs[0]
        n_outputs = len(outputs)
        n_times = len(times)
        n_samples = n_samples if n_samples is not None else 1
        container = np.empty(shape=(n_outputs, n_times, n_samples))
        start_index = 0
        for output_id, error_model in enumerate(self._error_models):
            end_index = start_index + error_model.n_parameters()
            container[output_id, ...] = error_model.sample(
                parameters=error_params[start_index:end_index],
                model_output=outputs[output_id], n_samples=n_samples,
                seed=seed)
            start_index = end_index
        if return_df is False:
            return container
        import pandas as pd
        output_names = self._mechanistic_model.outputs()
        sample_ids = np.arange(start=1, stop=n_samples + 1)
        frames = []
        for output_id, name in enumerate(output_names):
            for time_id, time in enumerate(times):
                frames.append(pd.DataFrame({
                    'ID': sample_ids, 'Time': time, 'Observable': name,
                    'Value': container[output_id, time_id, :]}))
        samples = pd.concat(frames)
        final_time = np.max(times)
        regimen = self.get_dosing_regimen(final_time)
        if (regimen is not None) and (include_regimen is True):
            for _id in sample_ids:
                regimen = regimen.copy()
                regimen['ID'] = _id
                samples = pd.concat([samples, regimen])
        return samples

    def set_dosing_regimen(
            self, dose, start=0, duration=0.01, period=None, num=None):
        """chi/_predictive_models.py:764-809."""
        try:
            self._mechanistic_model.set_dosing_regimen(
                dose, start, duration, period, num)
        except AttributeError:
            raise AttributeError(
                'The mechanistic model does not support to set dosing '
                'regimens. This may be because the underlying '
                'chi.MechanisticModel is a '
                'chi.PharmacodynamicModel.')

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
        if return_df is False:
            return measurements
        import pandas as pd
        n_outputs, n_times = measurements.shape[:2]
        names = self._predictive_model.get_output_names()
        output_names = []
        for name in names:
            output_names += [name] * (n_times * n_samples)
        t = np.broadcast_to(
            times[np.newaxis, :, np.newaxis],
            shape=(n_outputs, n_times, n_samples)).flatten()
        sample_ids = np.arange(start=1, stop=n_samples + 1)
        ids = np.broadcast_to(
            sample_ids[np.newaxis, np.newaxis, :],
            shape=(n_outputs, n_times, n_samples)).flatten()
        df = pd.DataFrame({
            'ID': ids, 'Time': t, 'Observable': output_names,
            'Value': measurements.flatten()})
        if pop.n_covariates() > 0:
            cov = np.broadcast_to(covariates, (n_samples, pop.n_covariates()))
            for idc, covariate in enumerate(pop.get_covariate_names()):
                df = pd.concat([df, pd.DataFrame({
                    'ID': sample_ids, 'Time': np.nan,
                    'Observable': covariate, 'Value': cov[..., idc]})])
        if include_regimen:
            regimen = self.get_dosing_regimen(np.max(times))
            if regimen is not None:
                for _id in sample_ids:
                    regimen = regimen.copy()
                    regimen['ID'] = _id
                    df = pd.concat([df, regimen])
        return df

    def set_dosing_regimen(
            self, dose, start=0, duration=0.01, period=None, num=None):
        self._predictive_model.set_dosing_regimen(
            dose, start, duration, period, num)
