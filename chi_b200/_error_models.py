"""Error models evaluated on the GPU.

Host-side mirror of ``chi/_error_models.py`` (seam B2):
same classes, method names and conventions (sigma <= 0 -> ``-inf`` score and
``inf`` gradient, ...).  The arithmetic runs in ``chi_b200_error_model`` (one
warp per call; stand-alone parity path) or, inside a log-likelihood, fused
into the integrator kernel.  ``sample`` stays on the host because bit parity
with numpy's PCG64 stream cannot be had on the device (SURVEY.md section 2.1,
K4).
"""
import copy

import numpy as np

from chi_b200 import _lib

try:
    import chi as _chi
    _ErrorModelBase = _chi.ErrorModel
except Exception:  # pragma: no cover
    _chi = None
    _ErrorModelBase = object


class ErrorModel(_ErrorModelBase):
    """Base class, see chi/_error_models.py:13-179."""

    _kind = None
    _default_names = None
    _device = None

    def __init__(self):
        super(ErrorModel, self).__init__()
        self._parameter_names = list(self._default_names)
        self._n_parameters = len(self._default_names)

    # -- device evaluation ---------------------------------------------------
    def _evaluate(self, parameters, model_output, sens, observations,
                  with_grad, pointwise):
        lib = _lib.lib()
        _lib.require_gpu()
        parameters = np.asarray(parameters, dtype=float).reshape(-1)
        if len(parameters) != self._n_parameters:
            raise ValueError(
                'The number of provided parameters does not match the expected'
                ' number of model parameters.')
        obs = np.asarray(observations, dtype=float).reshape(-1)
        n_obs = len(obs)
        ybar = np.asarray(model_output, dtype=float).reshape(-1)
        if len(ybar) != n_obs:
            raise ValueError(
                'The number of model outputs must match the number of '
                'observations, otherwise they cannot be compared pair-wise.')
        n_par = 0
        p_s = None
        if with_grad:
            sens = np.asarray(sens, dtype=float)
            if len(sens) != n_obs:
                raise ValueError(
                    'The first dimension of the sensitivities must match the '
                    'number of observations.')
            sens = sens.reshape(n_obs, -1)
            n_par = sens.shape[1]
            a_s, p_s = _lib.f64(sens)
        a_sig, p_sig = _lib.f64(parameters)
        a_y, p_y = _lib.f64(ybar)
        a_o, p_o = _lib.f64(obs)
        out, p_out = _lib.out_f64(1 + n_par + self._n_parameters)
        pw, p_pw = (None, None)
        if pointwise:
            pw, p_pw = _lib.out_f64(max(n_obs, 1))
        device = self._device if self._device is not None \
            else _lib.default_device()
        _lib.check(lib.chi_b200_error_model(
            device, _lib.ERR_KINDS[self._kind], p_sig, n_obs, n_par,
            p_y, p_s, p_o, 1 if with_grad else 0, p_out, p_pw))
        if pointwise:
            return pw[:n_obs]
        if with_grad:
            return float(out[0]), out[1:].copy()
        return float(out[0])

    # -- reference API ---------------------------------------------------------
    def compute_log_likelihood(self, parameters, model_output, observations):
        """chi/_error_models.py:333-378, 666-708, 1023-1066, 1377-1420."""
        return self._evaluate(
            parameters, model_output, None, observations, False, False)

    def compute_pointwise_ll(self, parameters, model_output, observations):
        """chi/_error_models.py:380-425, 710-759, 1068-1116, 1422-1470."""
        return self._evaluate(
            parameters, model_output, None, observations, False, True)

    def compute_sensitivities(
            self, parameters, model_output, model_sensitivities, observations):
        """chi/_error_models.py:427-467, 761-801, 1118-1158, 1472-1512."""
        return self._evaluate(
            parameters, model_output, model_sensitivities, observations,
            True, False)

    def get_parameter_names(self):
        return copy.copy(self._parameter_names)

    def n_parameters(self):
        return self._n_parameters

    def set_parameter_names(self, names=None):
        if names is None:
            self._parameter_names = list(self._default_names)
            return None
        if len(names) != self._n_parameters:
            raise ValueError(
                'Length of names does not match n_parameters.')
        self._parameter_names = [str(label) for label in names]

    def _sample_shape(self, parameters, model_output, n_samples):
        if len(parameters) != self._n_parameters:
            raise ValueError(
                'The number of provided parameters does not match the expected'
                ' number of model parameters.')
        model_output = np.asarray(model_output)
        if n_samples is None:
            n_samples = 1
        return model_output, (len(model_output), int(n_samples))


class GaussianErrorModel(ErrorModel):
    """chi/_error_models.py:539-869."""
    _kind = 'gaussian'
    _default_names = ['Sigma']

    def sample(self, parameters, model_output, n_samples=None, seed=None):
        """chi/_error_models.py:803-847 (host RNG, same draw order)."""
        model_output, shape = self._sample_shape(
            parameters, model_output, n_samples)
        rng = np.random.default_rng(seed=seed)
        samples = rng.normal(loc=0, scale=parameters[0], size=shape)
        return np.expand_dims(model_output, axis=1) + samples


class LogNormalErrorModel(ErrorModel):
    """chi/_error_models.py:872-1228."""
    _kind = 'lognormal'
    _default_names = ['Sigma log']

    def sample(self, parameters, model_output, n_samples=None, seed=None):
        """chi/_error_models.py:1160-1206."""
        model_output, shape = self._sample_shape(
            parameters, model_output, n_samples)
        sigma_log = parameters[0]
        rng = np.random.default_rng(seed=seed)
        samples = rng.lognormal(
            mean=-sigma_log**2 / 2, sigma=sigma_log, size=shape)
        return np.expand_dims(model_output, axis=1) * samples


class ConstantAndMultiplicativeGaussianErrorModel(ErrorModel):
    """chi/_error_models.py:182-536."""
    _kind = 'cam'
    _default_names = ['Sigma base', 'Sigma rel.']

    def sample(self, parameters, model_output, n_samples=None, seed=None):
        """chi/_error_models.py:469-514."""
        model_output, shape = self._sample_shape(
            parameters, model_output, n_samples)
        sigma_base, sigma_rel = parameters
        rng = np.random.default_rng(seed=seed)
        base_samples = rng.normal(loc=0, scale=sigma_base, size=shape)
        rel_samples = rng.normal(loc=0, scale=sigma_rel, size=shape)
        model_output = np.expand_dims(model_output, axis=1)
        return model_output + base_samples + model_output * rel_samples


class MultiplicativeGaussianErrorModel(ErrorModel):
    """chi/_error_models.py:1231-1580."""
    _kind = 'multiplicative'
    _default_names = ['Sigma rel.']

    def sample(self, parameters, model_output, n_samples=None, seed=None):
        """chi/_error_models.py:1514-1558."""
        model_output, shape = self._sample_shape(
            parameters, model_output, n_samples)
        rng = np.random.default_rng(seed=seed)
        rel_samples = rng.normal(loc=0, scale=parameters[0], size=shape)
        model_output = np.expand_dims(model_output, axis=1)
        return model_output + model_output * rel_samples


class ReducedErrorModel(object):
    """Error model with some parameters held at fixed values
    (chi/_error_models.py:1583-1906).  The wrapped model does the arithmetic
    (on the device); this class scatters the free parameters into the full
    parameter vector and drops the fixed entries from the gradient
    (1729-1750)."""

    def __init__(self, error_model):
        if not isinstance(error_model, ErrorModel):
            raise ValueError(
                'The error model has to be an instance of a '
                'chi.ErrorModel')
        self._error_model = error_model
        self._n_parameters = error_model.n_parameters()
        self._parameter_names = error_model.get_parameter_names()
        self._fixed_mask = None
        self._fixed_values = None

    # -- parameter bookkeeping ------------------------------------------------
    def _full(self, parameters):
        if self._fixed_mask is None:
            return parameters
        self._fixed_values[~self._fixed_mask] = parameters
        return self._fixed_values

    def fix_parameters(self, name_value_dict):
        """chi/_error_models.py:1752-1796: ``{name: value}`` fixes, ``{name:
        None}`` frees a parameter."""
        try:
            name_value_dict = dict(name_value_dict)
        except (TypeError, ValueError):
            raise ValueError(
                'The name-value dictionary has to be convertable to a python '
                'dictionary.')
        mask = np.zeros(self._n_parameters, dtype=bool) \
            if self._fixed_mask is None else self._fixed_mask
        values = np.empty(self._n_parameters) \
            if self._fixed_values is None else self._fixed_values
        for index, name in enumerate(self._parameter_names):
            if name not in name_value_dict:
                continue
            value = name_value_dict[name]
            mask[index] = value is not None
            values[index] = np.nan if value is None else value
        if not np.any(mask):
            mask = values = None
        self._fixed_mask, self._fixed_values = mask, values

    def get_error_model(self):
        return self._error_model

    def get_parameter_names(self):
        names = list(self._parameter_names)
        if self._fixed_mask is not None:
            names = [n for n, f in zip(names, self._fixed_mask) if not f]
        return names

    def n_fixed_parameters(self):
        return 0 if self._fixed_mask is None else int(np.sum(self._fixed_mask))

    def n_parameters(self):
        return self._n_parameters - self.n_fixed_parameters()

    def set_parameter_names(self, names=None):
        """chi/_error_models.py:1872-1906."""
        if names is None:
            self._error_model.set_parameter_names(None)
            self._parameter_names = self._error_model.get_parameter_names()
            return None
        if len(names) != self.n_parameters():
            raise ValueError('Length of names does not match n_parameters.')
        for name in names:
            if len(name) > 50:
                raise ValueError(
                    'Parameter names cannot exceed 50 characters.')
        names = [str(label) for label in names]
        full = list(self._parameter_names)
        free = [i for i in range(self._n_parameters)
                if self._fixed_mask is None or not self._fixed_mask[i]]
        for i, name in zip(free, names):
            full[i] = name
        self._error_model.set_parameter_names(full)
        self._parameter_names = full

    # -- evaluation -----------------------------------------------------------
    def compute_log_likelihood(self, parameters, model_output, observations):
        return self._error_model.compute_log_likelihood(
            self._full(parameters), model_output, observations)

    def compute_pointwise_ll(self, parameters, model_output, observations):
        return self._error_model.compute_pointwise_ll(
            self._full(parameters), model_output, observations)

    def compute_sensitivities(
            self, parameters, model_output, model_sensitivities, observations):
        score, sens = self._error_model.compute_sensitivities(
            self._full(parameters), model_output, model_sensitivities,
            observations)
        if self._fixed_mask is None:
            return score, sens
        n_mechanistic = np.asarray(model_sensitivities).shape[1]
        keep = np.ones(n_mechanistic + self._n_parameters, dtype=bool)
        keep[n_mechanistic:] = ~self._fixed_mask
        return score, sens[keep]

    def sample(self, parameters, model_output, n_samples=None, seed=None):
        return self._error_model.sample(
            self._full(parameters), model_output, n_samples, seed)
