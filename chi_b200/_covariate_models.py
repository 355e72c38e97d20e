"""Covariate models (interface of ``chi/_covariate_models.py``).

A covariate model shifts selected parameters of a population model with an
individual's covariates, ``vartheta_i = theta + X_i beta``
(chi/_covariate_models.py:257-320).  Inside a
:class:`chi_b200.CovariatePopulationModel` that shift and its adjoint are part
of the population kernels (``pop_mu_sigma`` and the beta reductions of
``pop_backward_kernel``, csrc/core.cu); called stand-alone, the two compute
methods run the small kernels behind ``chi_b200_covariate_linear``.

The bookkeeping is one table: the selected ``(parameter, dimension)`` pairs
in the order the reference produces them, and one name per (pair, covariate)
entry; everything else is derived from it on request.
"""
import numpy as np

from chi_b200 import _lib

try:
    import chi as _chi
    _CovariateModelBase = _chi.CovariateModel
except Exception:  # pragma: no cover
    _chi = None
    _CovariateModelBase = object


def _default_cov_names(n_cov):
    return ['Cov. %d' % (c + 1) for c in range(n_cov)]


class CovariateModel(_CovariateModelBase):
    """Interface of chi/_covariate_models.py:13-222."""

    def __init__(self, n_cov=1, cov_names=None):
        n_cov = int(n_cov)
        if n_cov < 1:
            raise ValueError(
                'The number of covariates has to be greater or equal to 1.')
        self._n_cov = n_cov
        if cov_names and len(cov_names) != n_cov:
            raise ValueError(
                'The number of covariate names has to match the number of '
                'covariates.')
        self._cov_names = [str(c) for c in cov_names] if cov_names \
            else _default_cov_names(n_cov)
        # selected (parameter, dimension) pairs; one name per (pair,
        # covariate) entry, pair-major
        self._pairs = [(0, 0)]
        self._entry_names = ['Param. 1'] * n_cov

    # -- the table -------------------------------------------------------------
    @property
    def _pidx(self):
        return np.array([p for p, _ in self._pairs])

    @property
    def _didx(self):
        return np.array([d for _, d in self._pairs])

    @property
    def _n_selected(self):
        return len(self._pairs)

    def get_set_population_parameters(self):
        return self._pidx, self._didx

    def n_covariates(self):
        return self._n_cov

    def n_parameters(self):
        return len(self._pairs) * self._n_cov

    # -- names -------------------------------------------------------------------
    def get_covariate_names(self):
        return self._cov_names

    def set_covariate_names(self, names=None):
        if names is None:
            self._cov_names = _default_cov_names(self._n_cov)
            return
        if len(names) != self._n_cov:
            raise ValueError(
                'Length of names does not match number of covariates.')
        self._cov_names = [str(c) for c in names]

    def get_parameter_names(self, exclude_cov_names=False):
        """One name per (selected pair, covariate), pair-major."""
        if exclude_cov_names:
            return list(self._entry_names)
        return [name + ' ' + self._cov_names[k % self._n_cov]
                for k, name in enumerate(self._entry_names)]

    def set_parameter_names(self, names=None, mask_names=False):
        if names is None:
            self._entry_names = [
                'Param. %d' % (s + 1) for s in range(len(self._pairs))
                for _ in range(self._n_cov)]
            return
        if len(names) != self.n_parameters():
            raise ValueError(
                'Length of names does not match number of parameters.')
        self._entry_names = [str(name) for name in names]

    # -- arithmetic: supplied by the concrete model ------------------------------
    def compute_population_parameters(
            self, parameters, pop_parameters, covariates):
        raise NotImplementedError

    def compute_sensitivities(
            self, parameters, pop_parameters, covariates, dlogp_dvartheta):
        raise NotImplementedError

    def set_population_parameters(self, indices):
        raise NotImplementedError


class LinearCovariateModel(CovariateModel):
    """chi/_covariate_models.py:225-358."""

    def _device_call(self, parameters, pop_parameters, covariates,
                     dlogp_dvartheta, forward):
        pop = np.ascontiguousarray(pop_parameters, dtype=float)
        if pop.ndim != 2:
            raise ValueError(
                'pop_parameters has to be of shape (n_pop_params_per_dim, '
                'n_dim).')
        n_pop, n_dim = pop.shape
        X = np.ascontiguousarray(covariates, dtype=float).reshape(
            -1, self._n_cov)
        n_ids = len(X)
        n_sel = len(self._pairs)
        beta = np.ascontiguousarray(parameters, dtype=float).reshape(
            n_sel, self._n_cov)
        _lib.require_gpu()
        _, p_pi = _lib.i32(self._pidx)
        _, p_di = _lib.i32(self._didx)
        a_b, p_b = _lib.f64(beta)
        a_p, p_p = _lib.f64(pop)
        a_x, p_x = _lib.f64(X)
        vartheta = dpop = dbeta = None
        p_v = p_dl = p_dp = p_db = None
        if forward:
            vartheta, p_v = _lib.out_f64((n_ids, n_pop, n_dim))
        else:
            dl = np.ascontiguousarray(dlogp_dvartheta, dtype=float).reshape(
                n_ids, n_pop, n_dim)
            a_dl, p_dl = _lib.f64(dl)
            dpop, p_dp = _lib.out_f64(n_pop * n_dim)
            dbeta, p_db = _lib.out_f64(n_sel * self._n_cov)
        _lib.check(_lib.lib().chi_b200_covariate_linear(
            _lib.default_device(), n_ids, n_pop, n_dim, self._n_cov, n_sel,
            p_pi, p_di, p_b, p_p, p_x, p_dl, p_v, p_dp, p_db))
        return vartheta if forward else (dpop, dbeta)

    def compute_population_parameters(
            self, parameters, pop_parameters, covariates):
        """``vartheta`` of shape ``(n_ids, n_pop_params_per_dim, n_dim)``,
        chi/_covariate_models.py:257-284."""
        return self._device_call(
            parameters, pop_parameters, covariates, None, True)

    def compute_sensitivities(
            self, parameters, pop_parameters, covariates, dlogp_dvartheta):
        """``(d pop (n_pop_params,), d beta (n_parameters,))``,
        chi/_covariate_models.py:286-320."""
        return self._device_call(
            parameters, pop_parameters, covariates, dlogp_dvartheta, False)

    def set_population_parameters(self, indices):
        """chi/_covariate_models.py:322-358.  The reference orders the unique
        pairs with two successive NON-stable argsorts (dimension, then
        parameter); the same two numpy calls are made here so that the order
        -- which fixes the layout of beta -- matches the reference on the same
        platform.  Everything downstream reads the order back through
        :meth:`get_set_population_parameters`."""
        pairs = []
        for idx in indices:
            pair = (int(idx[0]), int(idx[1]))
            if pair not in pairs:
                pairs.append(pair)
        table = np.array(pairs)
        table = table[np.argsort(table[:, 1]), :]
        table = table[np.argsort(table[:, 0]), :]
        self._pairs = [(int(p), int(d)) for p, d in table]
        self.set_parameter_names()
