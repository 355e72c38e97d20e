"""Covariate models.

Host-side mirror of ``chi/_covariate_models.py``.  A covariate model here is a
*descriptor* (number of covariates, which population parameters are shifted,
names): the arithmetic ``vartheta_i = theta + X_i beta`` and its adjoint
(chi/_covariate_models.py:257-320) runs on the GPU inside the population
kernels (``pop_mu_sigma`` and the beta reductions of ``pop_backward_kernel``
in csrc/core.cu) whenever a :class:`chi_b200.CovariatePopulationModel` is
evaluated.
"""
import copy

import numpy as np

try:
    import chi as _chi
    _CovariateModelBase = _chi.CovariateModel
except Exception:  # pragma: no cover
    _chi = None
    _CovariateModelBase = object


class CovariateModel(_CovariateModelBase):
    """Base class, see chi/_covariate_models.py:13-222."""

    def __init__(self, n_cov=1, cov_names=None):
        n_cov = int(n_cov)
        if n_cov < 1:
            raise ValueError(
                'The number of covariates has to be greater or equal to 1.')
        self._n_cov = n_cov
        if cov_names:
            if len(cov_names) != self._n_cov:
                raise ValueError(
                    'The number of covariate names has to match the number of '
                    'covariates.')
            cov_names = [str(name) for name in cov_names]
        else:
            cov_names = [
                'Cov. %d' % (id_cov + 1) for id_cov in range(self._n_cov)]
        self._cov_names = cov_names
        self._pidx = np.array([0])
        self._didx = np.array([0])
        self._n_selected = 1
        self._n_parameters = self._n_selected * self._n_cov
        self._parameter_names = ['Param. 1'] * self._n_cov

    def compute_population_parameters(
            self, parameters, pop_parameters, covariates):
        raise NotImplementedError

    def compute_sensitivities(
            self, parameters, pop_parameters, covariates, dlogp_dvartheta):
        raise NotImplementedError

    def get_covariate_names(self):
        return self._cov_names

    def get_parameter_names(self, exclude_cov_names=False):
        param_names = copy.copy(self._parameter_names)
        if not exclude_cov_names:
            names = []
            for name_id, name in enumerate(param_names):
                cov_name = self._cov_names[name_id % self._n_cov]
                names += [name + ' ' + cov_name]
            param_names = names
        return param_names

    def get_set_population_parameters(self):
        return (self._pidx.copy(), self._didx.copy())

    def n_covariates(self):
        return self._n_cov

    def n_parameters(self):
        return self._n_parameters

    def set_covariate_names(self, names=None):
        if names is None:
            self._cov_names = [
                'Cov. %d' % (id_cov + 1) for id_cov in range(self._n_cov)]
            return None
        if len(names) != self._n_cov:
            raise ValueError(
                'Length of names does not match number of covariates.')
        self._cov_names = [str(label) for label in names]

    def set_parameter_names(self, names=None, mask_names=False):
        if names is None:
            names = []
            for id_p in range(self._n_selected):
                names += ['Param. %d' % (id_p + 1)] * self._n_cov
            self._parameter_names = names
            return None
        if len(names) != self._n_parameters:
            raise ValueError(
                'Length of names does not match number of parameters.')
        self._parameter_names = [str(label) for label in names]

    def set_population_parameters(self, indices):
        raise NotImplementedError


class LinearCovariateModel(CovariateModel):
    """chi/_covariate_models.py:225-358."""

    def __init__(self, n_cov=1, cov_names=None):
        super(LinearCovariateModel, self).__init__(n_cov, cov_names)
        self._n_parameters = self._n_cov

    def compute_population_parameters(
            self, parameters, pop_parameters, covariates):
        raise NotImplementedError(
            'chi_b200 evaluates the covariate shift on the GPU as part of '
            'CovariatePopulationModel; there is no stand-alone host path.')

    def compute_sensitivities(
            self, parameters, pop_parameters, covariates, dlogp_dvartheta):
        raise NotImplementedError(
            'chi_b200 evaluates the covariate adjoint on the GPU as part of '
            'CovariatePopulationModel; there is no stand-alone host path.')

    def set_population_parameters(self, indices):
        """chi/_covariate_models.py:322-358.  The reference orders the pairs
        with two successive non-stable argsorts; the sort kind is made
        explicit here ('quicksort', numpy's default) so the selection order
        matches the reference on the same platform, and everything that
        consumes the order reads it back through
        ``get_set_population_parameters``."""
        unique = []
        for idx in indices:
            pair = [int(idx[0]), int(idx[1])]
            if pair not in unique:
                unique.append(pair)
        indices = np.array(unique)
        indices = indices[np.argsort(indices[:, 1]), :]
        indices = indices[np.argsort(indices[:, 0]), :]
        self._pidx = indices[:, 0]
        self._didx = indices[:, 1]
        n_params = len(self._pidx)
        self._n_parameters = self._n_cov * n_params
        self._n_selected = n_params
        self.set_parameter_names()
