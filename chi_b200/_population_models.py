"""Population models evaluated on the GPU.

Host-side mirror of ``chi/_population_models.py`` for the models on the hot
path (Pooled, Heterogeneous, Gaussian, LogNormal, Composed, Covariate;
SURVEY.md section 8a rows a15-a19).  The classes keep chi's bookkeeping API
(names, dimensions, special dims, ``n_hierarchical_parameters`` ...); every
``compute_*`` method flattens the model into the block descriptor of the C
ABI (``chi_b200_population_set``) and evaluates it with the population
kernels.  ``sample`` draws on the host with numpy's generator in the
reference's order (distribution- and seed-level parity).
"""
import copy

import numpy as np

from chi_b200 import _lib
from chi_b200._covariate_models import CovariateModel

try:
    import chi as _chi
    _PopulationModelBase = _chi.PopulationModel
except Exception:  # pragma: no cover
    _chi = None
    _PopulationModelBase = object


class PopulationModel(_PopulationModelBase):
    """Base class, see chi/_population_models.py:18-326."""

    _kind = None

    def __init__(self, n_dim=1, dim_names=None):
        if n_dim < 1:
            raise ValueError(
                'The dimension of the population model has to be greater or '
                'equal to 1.')
        self._n_dim = int(n_dim)
        self._n_hierarchical_dim = self._n_dim
        self._special_dims = []
        self._n_pooled_dims = 0
        self._n_hetero_dims = 0
        self._n_covariates = 0
        self._n_ids = 1
        self._device = None
        self._dev_cache = None
        if dim_names:
            if len(dim_names) != self._n_dim:
                raise ValueError(
                    'The number of dimension names has to match the number of '
                    'dimensions of the population model.')
            dim_names = [str(name) for name in dim_names]
        else:
            dim_names = [
                'Dim. %d' % (id_dim + 1) for id_dim in range(self._n_dim)]
        self._dim_names = dim_names

    def __deepcopy__(self, memo):
        cls = self.__class__
        result = cls.__new__(cls)
        memo[id(self)] = result
        for k, v in self.__dict__.items():
            setattr(result, k, None if k == '_dev_cache'
                    else copy.deepcopy(v, memo))
        return result

    # -- device plumbing -----------------------------------------------------
    def _blocks(self):
        """Flat list of block descriptors
        (kind, n_dim, centered, n_cov, pidx, didx)."""
        raise NotImplementedError

    def _n_top(self, n_ids):
        return self.n_hierarchical_parameters(n_ids)[1]

    def _device_problem(self, n_ids, covariates):
        blocks = self._blocks()
        cov_bytes = None
        if self._n_covariates > 0:
            if covariates is None:
                raise ValueError(
                    'The population model needs covariates, but no covariates '
                    'have been provided.')
            covariates = np.ascontiguousarray(
                covariates, dtype=float).reshape(n_ids, self._n_covariates)
            cov_bytes = covariates.tobytes()
        sig = (n_ids, repr([(b[0], b[1], b[2], b[3],
                             None if b[4] is None else list(b[4]),
                             None if b[5] is None else list(b[5]))
                            for b in blocks]), cov_bytes)
        if self._dev_cache is not None and self._dev_cache[0] == sig:
            return self._dev_cache[1]
        problem = _lib.Problem(-1, self._device)
        _lib.check(_lib.lib().chi_b200_problem_set_n_ids(
            problem.handle, n_ids))
        set_population(problem, blocks, covariates)
        self._dev_cache = (sig, problem)
        return problem

    def _hier_columns(self):
        """Indices of the hierarchical (non pooled / heterogeneous) dims."""
        cols = []
        d = 0
        for b in self._blocks():
            if b[0] in ('gaussian', 'lognormal', 'truncated_gaussian'):
                cols += list(range(d, d + b[1]))
            d += b[1]
        return cols

    def _device_eval(self, parameters, eta, dlogp_dpsi, covariates,
                     with_grad):
        """Returns eta, psi (n_ids, n_dim), score and the reduced gradient
        [d bottom | d top]."""
        parameters = np.asarray(parameters, dtype=float).reshape(-1)
        eta = np.asarray(eta, dtype=float)
        if eta.ndim == 1:
            eta = eta.reshape(-1, 1) if self._n_dim == 1 else \
                eta.reshape(-1, self._n_dim)
        n_ids = eta.shape[0]
        self._check_n_ids(n_ids)
        n_top = self._n_top(n_ids)
        if len(parameters) != n_top:
            raise ValueError(
                'The number of provided parameters does not match the expected'
                ' number of top-level parameters.')
        problem = self._device_problem(n_ids, covariates)
        hcols = self._hier_columns()
        bottom = np.ascontiguousarray(eta[:, hcols])
        lib = _lib.lib()
        a_t, p_t = _lib.f64(parameters)
        a_b, p_b = _lib.f64(bottom.reshape(-1))
        p_d = None
        if dlogp_dpsi is not None:
            a_d, p_d = _lib.f64(np.asarray(
                dlogp_dpsi, dtype=float).reshape(n_ids, self._n_dim))
        o_eta, p_eta = _lib.out_f64((n_ids, self._n_dim))
        o_psi, p_psi = _lib.out_f64((n_ids, self._n_dim))
        o_s, p_s = _lib.out_f64(1)
        o_g, p_g = _lib.out_f64(max(bottom.size + n_top, 1))
        _lib.check(lib.chi_b200_population_eval(
            problem.handle, p_t, p_b, p_d, 1 if with_grad else 0, p_eta,
            p_psi, p_s, p_g))
        return o_eta, o_psi, float(o_s[0]), o_g[:bottom.size + n_top]

    def _check_n_ids(self, n_ids):
        pass

    # -- reference API ---------------------------------------------------------
    def compute_individual_parameters(
            self, parameters, eta, return_eta=False, covariates=None,
            *args, **kwargs):
        eta = np.asarray(eta, dtype=float)
        if eta.ndim == 1:
            eta = eta.reshape(self._n_ids, self._n_dim)
        o_eta, o_psi, _, _ = self._device_eval(
            parameters, eta, None, covariates, False)
        return o_eta if return_eta else o_psi

    def compute_log_likelihood(
            self, parameters, observations, covariates=None, *args, **kwargs):
        observations = np.asarray(observations, dtype=float)
        if observations.ndim == 1:
            observations = observations[:, np.newaxis]
        o_eta, _, score, _ = self._device_eval(
            parameters, observations, None, covariates, False)
        if not self._special_match(o_eta, observations):
            return -np.inf
        return score

    def _special_match(self, eta_dev, observations):
        """Pooled / heterogeneous dims are delta distributions: -inf unless
        the observations equal the parameters
        (chi/_population_models.py:1986-1988, 2866-2868)."""
        for info in self._special_dims:
            s, e = info[0], info[1]
            if np.any(eta_dev[:, s:e] != observations[:, s:e]):
                return False
        return True

    def compute_pointwise_ll(self, parameters, observations, *args, **kwargs):
        raise NotImplementedError

    def compute_sensitivities(
            self, parameters, observations, dlogp_dpsi=None, reduce=False,
            covariates=None, *args, **kwargs):
        """Returns ``(score, dpsi, dtheta)`` or, with ``reduce=True``,
        ``(score, [d bottom | d top])``
        (chi/_population_models.py:128-159, 56-74, 386-449)."""
        observations = np.asarray(observations, dtype=float)
        if observations.ndim == 1:
            observations = observations[:, np.newaxis]
        n_ids = observations.shape[0]
        o_eta, _, score, grad = self._device_eval(
            parameters, observations, dlogp_dpsi, covariates, True)
        if not self._special_match(o_eta, observations):
            score = -np.inf
        if reduce:
            return score, grad
        hcols = self._hier_columns()
        n_bottom = n_ids * len(hcols)
        dpsi = np.zeros((n_ids, self._n_dim))
        if dlogp_dpsi is not None:
            dpsi += np.asarray(dlogp_dpsi, dtype=float).reshape(
                n_ids, self._n_dim)
        if hcols:
            dpsi[:, hcols] = grad[:n_bottom].reshape(n_ids, len(hcols))
        dtop = grad[n_bottom:]
        dtheta = np.zeros(self.n_parameters())
        # non-reduced layout: pooled / heterogeneous entries are zero
        t_red = t_full = 0
        for b in self._blocks():
            nt = _block_n_top(b, n_ids)
            if b[0] in ('gaussian', 'lognormal', 'truncated_gaussian'):
                dtheta[t_full:t_full + nt] = dtop[t_red:t_red + nt]
            t_red += nt
            t_full += nt
        return score, dpsi, dtheta

    def get_covariate_names(self):
        return []

    def get_dim_names(self):
        return copy.copy(self._dim_names)

    def get_special_dims(self):
        return self._special_dims, self._n_pooled_dims, self._n_hetero_dims

    def get_parameter_names(self, exclude_dim_names=False):
        if exclude_dim_names:
            return copy.copy(self._parameter_names)
        names = []
        for name_id, name in enumerate(self._parameter_names):
            current_dim = name_id % self._n_dim
            names += [name + ' ' + self._dim_names[current_dim]]
        return names

    def n_covariates(self):
        return self._n_covariates

    def n_dim(self):
        return self._n_dim

    def n_hierarchical_dim(self):
        return self._n_hierarchical_dim

    def n_hierarchical_parameters(self, n_ids):
        raise NotImplementedError

    def n_ids(self):
        return self._n_ids

    def n_parameters(self):
        return self._n_parameters

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        raise NotImplementedError

    def set_covariate_names(self, names=None):
        return None

    def set_dim_names(self, names=None):
        if names is None:
            self._dim_names = [
                'Dim. %d' % (id_dim + 1) for id_dim in range(self._n_dim)]
            return None
        if len(names) != self._n_dim:
            raise ValueError(
                'Length of names does not match the number of dimensions.')
        self._dim_names = [str(label) for label in names]

    def set_n_ids(self, n_ids):
        n_ids = int(n_ids)
        if n_ids < 1:
            raise ValueError(
                'n_ids is invalid. The number of individuals has to be '
                'positive.')
        self._n_ids = n_ids

    def set_parameter_names(self, names=None):
        if names is None:
            self._parameter_names = list(self._default_parameter_names())
            return None
        if len(names) != self._n_parameters:
            raise ValueError(
                'Length of names does not match the number of parameters.')
        self._parameter_names = [str(label) for label in names]


def _block_n_top(block, n_ids):
    kind, n_dim, _, n_cov, pidx, _ = block
    if kind == 'pooled':
        return n_dim
    if kind == 'heterogeneous':
        return n_ids * n_dim
    n = 2 * n_dim
    if n_cov:
        n += len(pidx) * n_cov
    return n


def set_population(problem, blocks, covariates):
    """Sends a block list to ``chi_b200_population_set``."""
    lib = _lib.lib()
    kind = [_lib.POP_KINDS[b[0]] for b in blocks]
    n_dim = [b[1] for b in blocks]
    centered = [1 if b[2] else 0 for b in blocks]
    n_cov = [b[3] for b in blocks]
    n_sel = [0 if b[4] is None else len(b[4]) for b in blocks]
    pidx, didx = [], []
    for b in blocks:
        if b[4] is not None:
            pidx += [int(v) for v in b[4]]
            didx += [int(v) for v in b[5]]
    a1, p1 = _lib.i32(kind)
    a2, p2 = _lib.i32(n_dim)
    a3, p3 = _lib.i32(centered)
    a4, p4 = _lib.i32(n_cov)
    a5, p5 = _lib.i32(n_sel)
    a6, p6 = _lib.i32(pidx if pidx else [0])
    a7, p7 = _lib.i32(didx if didx else [0])
    p_cov = None
    if covariates is not None:
        a8, p_cov = _lib.f64(np.asarray(covariates, dtype=float).reshape(-1))
    _lib.check(lib.chi_b200_population_set(
        problem.handle, len(blocks), p1, p2, p3, p4, p5, p6, p7, p_cov))


class _DensityModel(PopulationModel):
    """Shared code of GaussianModel and LogNormalModel."""

    def __init__(self, n_dim=1, dim_names=None, centered=True):
        super(_DensityModel, self).__init__(n_dim, dim_names)
        self._n_parameters = 2 * self._n_dim
        self._parameter_names = list(self._default_parameter_names())
        self._centered = bool(centered)

    def _blocks(self):
        return [(self._kind, self._n_dim, self._centered, 0, None, None)]

    def n_hierarchical_parameters(self, n_ids):
        n_ids = int(n_ids)
        return (n_ids * self._n_dim, self._n_parameters)

    def _sample_parameters(self, parameters):
        parameters = np.asarray(parameters)
        if len(parameters.flatten()) != self._n_parameters:
            raise ValueError(
                'The number of provided parameters does not match the expected'
                ' number of top-level parameters.')
        if parameters.ndim != 2:
            n_parameters = len(parameters) // self._n_dim
            parameters = parameters.reshape(n_parameters, self._n_dim)
        return parameters


class GaussianModel(_DensityModel):
    """chi/_population_models.py:1385-1874."""
    _kind = 'gaussian'

    def _default_parameter_names(self):
        return ['Mean'] * self._n_dim + ['Std.'] * self._n_dim

    def get_mean_and_std(self, parameters):
        """chi/_population_models.py:1759-1777."""
        parameters = self._sample_parameters(parameters)
        return np.vstack([parameters[0], parameters[1]])

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        """chi/_population_models.py:1797-1846."""
        parameters = self._sample_parameters(parameters)
        if n_samples is None:
            n_samples = 1
        sample_shape = (int(n_samples), self._n_dim)
        mus = parameters[0]
        sigmas = parameters[1]
        if not self._centered:
            mus = np.zeros(mus.shape)
            sigmas = np.ones(sigmas.shape)
        if np.any(sigmas < 0):
            raise ValueError(
                'A Gaussian distribution only accepts strictly positive '
                'standard deviations.')
        rng = np.random.default_rng(seed=seed)
        return rng.normal(loc=mus, scale=sigmas, size=sample_shape)


class LogNormalModel(_DensityModel):
    """chi/_population_models.py:2212-2756."""
    _kind = 'lognormal'

    def _default_parameter_names(self):
        return ['Log mean'] * self._n_dim + ['Log std.'] * self._n_dim

    def get_mean_and_std(self, parameters):
        """chi/_population_models.py:2601-2634."""
        parameters = np.asarray(parameters)
        if parameters.ndim != 2:
            n_parameters = len(parameters) // self._n_dim
            parameters = parameters.reshape(n_parameters, self._n_dim)
        mus = parameters[0]
        sigmas = parameters[1]
        if np.any(sigmas < 0):
            raise ValueError('The standard deviation cannot be negative.')
        mean = np.exp(mus + sigmas**2 / 2)
        std = np.sqrt(
            np.exp(2 * mus + sigmas**2) * (np.exp(sigmas**2) - 1))
        return np.vstack([mean, std])

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        """chi/_population_models.py:2678-2732."""
        parameters = self._sample_parameters(parameters)
        if n_samples is None:
            n_samples = 1
        sample_shape = (int(n_samples), self._n_dim)
        rng = np.random.default_rng(seed=seed)
        mus = parameters[0]
        sigmas = parameters[1]
        if not self._centered:
            mus = np.zeros(mus.shape)
            sigmas = np.ones(sigmas.shape)
            return rng.normal(loc=mus, scale=sigmas, size=sample_shape)
        if np.any(sigmas <= 0):
            raise ValueError(
                'A log-normal distribution only accepts strictly positive '
                'standard deviations.')
        return rng.lognormal(mean=mus, sigma=sigmas, size=sample_shape)


class TruncatedGaussianModel(_DensityModel):
    """Gaussian population truncated at zero,
    chi/_population_models.py:3500-3939.  Always centered: the bottom-level
    parameters are the individual parameters themselves; sigma <= 0 or a
    negative individual parameter scores -inf (3557-3558, 3656-3658)."""
    _kind = 'truncated_gaussian'

    def __init__(self, n_dim=1, dim_names=None):
        super(TruncatedGaussianModel, self).__init__(
            n_dim, dim_names, centered=True)

    def _default_parameter_names(self):
        return ['Mu'] * self._n_dim + ['Sigma'] * self._n_dim

    def get_mean_and_std(self, parameters):
        """chi/_population_models.py:3771-3821."""
        from scipy.stats import norm
        parameters = np.asarray(parameters)
        if parameters.ndim != 2:
            n_parameters = len(parameters) // self._n_dim
            parameters = parameters.reshape(n_parameters, self._n_dim)
        mus = parameters[0]
        sigmas = parameters[1]
        if np.any(sigmas < 0):
            raise ValueError('The standard deviation cannot be negative.')
        f = norm.pdf(mus / sigmas) / (1 - norm.cdf(-mus / sigmas))
        # (the reference allocates (n_parameters, n_dim) and leaves all but
        # the first two rows uninitialised; the documented shape is returned)
        return np.vstack([
            mus + sigmas * f,
            np.sqrt(sigmas**2 * (1 - mus / sigmas * f - f**2))])

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        """chi/_population_models.py:3863-3919.  Like the reference this
        draws ``truncnorm.rvs(a=0, b=inf, loc=mu, scale=sigma)`` from numpy's
        global generator seeded with ``seed`` (scipy's ``a`` is in standard
        units, so the draws are truncated at ``mu``; reproduced as is for
        seed-level parity)."""
        from scipy.stats import truncnorm
        parameters = self._sample_parameters(parameters)
        if n_samples is None:
            n_samples = 1
        sample_shape = (int(n_samples), self._n_dim)
        mus = parameters[0]
        sigmas = parameters[1]
        if np.any(sigmas < 0):
            raise ValueError(
                'A truncated Gaussian distribution only accepts strictly '
                'positive standard deviations.')
        if isinstance(seed, np.random.Generator):
            seed = seed.integers(low=0, high=1E6)
        np.random.seed(seed)
        return truncnorm.rvs(
            a=0, b=np.inf, loc=mus, scale=sigmas, size=sample_shape)


class PooledModel(PopulationModel):
    """chi/_population_models.py:2759-3061."""
    _kind = 'pooled'

    def __init__(self, n_dim=1, dim_names=None):
        super(PooledModel, self).__init__(n_dim, dim_names)
        self._n_parameters = self._n_dim
        self._n_hierarchical_dim = 0
        self._special_dims = [[0, self._n_dim, 0, self._n_parameters, True]]
        self._n_pooled_dims = self._n_dim
        self._n_hetero_dims = 0
        self._parameter_names = ['Pooled'] * self._n_dim

    def _default_parameter_names(self):
        return ['Pooled'] * self._n_dim

    def _blocks(self):
        return [('pooled', self._n_dim, True, 0, None, None)]

    def n_hierarchical_parameters(self, n_ids):
        return (0, self._n_parameters)

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        """chi/_population_models.py:3005-3037."""
        parameters = np.asarray(parameters)
        if len(parameters.flatten()) != self._n_parameters:
            raise ValueError(
                'The number of provided parameters does not match the expected'
                ' number of top-level parameters.')
        if parameters.ndim != 2:
            n_parameters = len(parameters) // self._n_dim
            parameters = parameters.reshape(n_parameters, self._n_dim)
        if n_samples is None:
            return parameters
        return np.broadcast_to(
            parameters, shape=(int(n_samples), self._n_dim))


class HeterogeneousModel(PopulationModel):
    """chi/_population_models.py:1877-2209."""
    _kind = 'heterogeneous'

    def __init__(self, n_dim=1, dim_names=None, n_ids=1):
        super(HeterogeneousModel, self).__init__(n_dim, dim_names)
        self._n_hierarchical_dim = 0
        self.set_n_ids(n_ids, no_shortcut=True)
        self._n_pooled_dims = 0
        self._n_hetero_dims = self._n_dim

    def _default_parameter_names(self):
        names = []
        for _id in range(self._n_ids):
            names += ['ID %d' % (_id + 1)] * self._n_dim
        return names

    def _blocks(self):
        return [('heterogeneous', self._n_dim, True, 0, None, None)]

    def _check_n_ids(self, n_ids):
        if n_ids != self._n_ids:
            raise ValueError(
                'The number of individuals does not match n_ids of the '
                'heterogeneous model.')

    def n_hierarchical_parameters(self, n_ids):
        return (0, int(n_ids) * self._n_dim)

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        """chi/_population_models.py:2113-2152."""
        parameters = np.asarray(parameters)
        if parameters.ndim != 2:
            parameters = parameters.reshape(self._n_ids, self._n_dim)
        ids = np.arange(self._n_ids)
        rng = np.random.default_rng(seed=seed)
        n_samples = n_samples if n_samples else 1
        return parameters[rng.choice(ids, size=n_samples, replace=True)]

    def set_n_ids(self, n_ids, no_shortcut=False):
        """chi/_population_models.py:2154-2187."""
        n_ids = int(n_ids)
        if n_ids < 1:
            raise ValueError(
                'The number of modelled individuals needs to be greater or '
                'equal to 1.')
        if (n_ids == self._n_ids) and not no_shortcut:
            return None
        self._n_ids = n_ids
        self._n_parameters = self._n_ids * self._n_dim
        self._parameter_names = self._default_parameter_names()
        self._special_dims = [[0, self._n_dim, 0, self._n_parameters, False]]
        self._dev_cache = None


class CovariatePopulationModel(PopulationModel):
    """chi/_population_models.py:956-1382."""

    def __init__(self, population_model, covariate_model, dim_names=None):
        if not isinstance(population_model, PopulationModel):
            raise TypeError(
                'The population model has to be an instance of a '
                'chi.PopulationModel.')
        if not isinstance(covariate_model, CovariateModel):
            raise TypeError(
                'The covariate model has to be an instance of a '
                'chi.CovariateModel.')
        if isinstance(population_model, ComposedPopulationModel):
            raise TypeError(
                'The population model cannot be an instance of a '
                'chi.ComposedPopulationModel. Please compose multiple '
                'covariate models instead.')
        if not isinstance(population_model, _DensityModel) or isinstance(
                population_model, TruncatedGaussianModel):
            raise TypeError(
                'chi_b200 supports covariate models over GaussianModel and '
                'LogNormalModel only.')
        super(CovariatePopulationModel, self).__init__(
            population_model.n_dim())
        self._population_model = copy.deepcopy(population_model)
        self._covariate_model = copy.deepcopy(covariate_model)
        self._n_dim = self._population_model.n_dim()
        self._n_pop = self._population_model.n_parameters()
        self._n_hierarchical_dim = self._population_model.n_hierarchical_dim()
        self._n_covariates = self._covariate_model.n_covariates()
        self._special_dims, self._n_pooled_dims, self._n_hetero_dims = \
            self._population_model.get_special_dims()
        n_cov = self._covariate_model.n_covariates()
        self._population_model.set_dim_names(dim_names)
        indices = []
        for dim_id in range(self._n_dim):
            for param_id in range(self._n_pop // self._n_dim):
                indices.append([param_id, dim_id])
        self._covariate_model.set_population_parameters(indices)
        names = []
        for name in self._population_model.get_parameter_names():
            names += [name] * n_cov
        self._covariate_model.set_parameter_names(names)

    def _blocks(self):
        pidx, didx = self._covariate_model.get_set_population_parameters()
        inner = self._population_model
        return [(inner._kind, inner._n_dim, inner._centered,
                 self._n_covariates, list(pidx), list(didx))]

    def get_covariate_names(self):
        return self._covariate_model.get_covariate_names()

    def get_dim_names(self):
        return self._population_model.get_dim_names()

    def get_parameter_names(self, exclude_dim_names=False):
        names = self._population_model.get_parameter_names(exclude_dim_names)
        names += self._covariate_model.get_parameter_names()
        return names

    def n_hierarchical_parameters(self, n_ids):
        n_ids, _ = self._population_model.n_hierarchical_parameters(n_ids)
        return (n_ids, self.n_parameters())

    def n_parameters(self):
        return self._population_model.n_parameters() + \
            self._covariate_model.n_parameters()

    def sample(self, parameters, covariates, n_samples=None, seed=None):
        """chi/_population_models.py:1247-1299 (host RNG; the covariate
        shift of the handful of sampled sub-populations is applied with the
        same indexing the kernels use)."""
        covariates = np.array(covariates)
        if covariates.ndim == 1:
            covariates = covariates[np.newaxis, :]
        if covariates.shape[1] != self._n_covariates:
            raise ValueError(
                'Provided covariates do not match the number of covariates.')
        if n_samples is None:
            n_samples = 1
        n_samples = int(n_samples)
        covariates = np.broadcast_to(
            covariates, (n_samples, self._n_covariates))
        parameters = np.asarray(parameters)
        pop_params = parameters[:self._n_pop].reshape(
            self._n_pop // self._n_dim, self._n_dim)
        pidx, didx = self._covariate_model.get_set_population_parameters()
        beta = parameters[self._n_pop:].reshape(
            len(pidx), self._n_covariates)
        vartheta = np.zeros((n_samples,) + pop_params.shape)
        vartheta += pop_params[np.newaxis, ...]
        vartheta[:, pidx, didx] += covariates @ beta.T
        seed = np.random.default_rng(seed)
        psi = np.empty(shape=(n_samples, self._n_dim))
        for ids, params in enumerate(vartheta):
            psi[ids] = self._population_model.sample(
                params, n_samples=1, seed=seed)[0]
        return psi

    def set_covariate_names(self, names=None):
        self._covariate_model.set_covariate_names(names)

    def set_dim_names(self, names=None):
        self._population_model.set_dim_names(names)
        names = self._population_model.get_parameter_names()
        names = np.array(names).reshape(
            self._n_pop // self._n_dim, self._n_dim)
        pidx, didx = self._covariate_model.get_set_population_parameters()
        names = names[pidx, didx]
        n = []
        for name in names:
            n += [name] * self._n_covariates
        self._covariate_model.set_parameter_names(n)

    def set_parameter_names(self, names=None):
        if names is None:
            self._population_model.set_parameter_names()
            self._covariate_model.set_parameter_names()
            return None
        self._population_model.set_parameter_names(names[:self._n_pop])
        self._covariate_model.set_parameter_names(names[self._n_pop:])

    def set_population_parameters(self, indices):
        """chi/_population_models.py:1354-1382."""
        indices = np.array(indices)
        upper = np.max(indices, axis=0)
        n_pop = self._n_pop // self._n_dim
        out_of_bounds = \
            (upper[0] >= n_pop) or (upper[1] >= self._n_dim) or \
            (np.min(indices) < 0)
        if out_of_bounds:
            raise IndexError('The provided indices are out of bounds.')
        self._covariate_model.set_population_parameters(indices)
        names = np.array(self._population_model.get_parameter_names())
        names = names.reshape(n_pop, self._n_dim)[indices[:, 0], indices[:, 1]]
        n = []
        for name in names:
            n += [name] * self._n_covariates
        self._covariate_model.set_parameter_names(n)
        self._dev_cache = None


class ComposedPopulationModel(PopulationModel):
    """chi/_population_models.py:329-953."""

    def __init__(self, population_models):
        for pop_model in population_models:
            if not isinstance(pop_model, PopulationModel):
                raise TypeError(
                    'The population models have to be instances of '
                    'chi.PopulationModel.')
        n_ids = 1
        for pop_model in population_models:
            if (n_ids > 1) and (pop_model.n_ids() > 1) and (
                    n_ids != pop_model.n_ids()):
                raise ValueError(
                    'All population models must model the same number of '
                    'individuals.')
            n_ids = n_ids if n_ids > 1 else pop_model.n_ids()
        self._population_models = population_models
        self._n_ids = n_ids
        self._device = None
        self._dev_cache = None
        self._n_bottom, self._n_top_cached = self.n_hierarchical_parameters(
            self._n_ids)
        self._set_population_model_properties()
        names = self.get_parameter_names()
        if len(np.unique(names)) != len(names):
            dim_names = [
                'Dim. %d' % (dim_id + 1) for dim_id in range(self._n_dim)]
            self.set_dim_names(dim_names)

    def _set_population_model_properties(self):
        """chi/_population_models.py:451-492."""
        n_dim = 0
        n_parameters = 0
        n_covariates = 0
        n_hierarchical_dim = 0
        special_dim = []
        n_pooled = 0
        n_hetero = 0
        for pop_model in self._population_models:
            s, p, h = pop_model.get_special_dims()
            n_pooled += p
            n_hetero += h
            for entry in s:
                special_dim += [[
                    entry[0] + n_dim, entry[1] + n_dim,
                    entry[2] + n_parameters, entry[3] + n_parameters,
                    entry[4]]]
            n_covariates += pop_model.n_covariates()
            n_dim += pop_model.n_dim()
            n_hierarchical_dim += pop_model.n_hierarchical_dim()
            n_parameters += pop_model.n_parameters()
        self._n_dim = n_dim
        self._n_parameters = n_parameters
        self._n_covariates = n_covariates
        self._n_hierarchical_dim = n_hierarchical_dim
        self._special_dims = special_dim
        self._n_pooled_dims = n_pooled
        self._n_hetero_dims = n_hetero
        self._dev_cache = None

    def _blocks(self):
        blocks = []
        for pop_model in self._population_models:
            blocks += pop_model._blocks()
        return blocks

    def _check_n_ids(self, n_ids):
        for pop_model in self._population_models:
            pop_model._check_n_ids(n_ids)

    def _shape_eta(self, eta):
        """chi/_population_models.py:534-565."""
        n_dim = len(eta) // self._n_ids
        eta = eta.reshape(self._n_ids, n_dim)
        if n_dim == self._n_dim:
            return eta
        if (n_dim + self._n_pooled_dims + self._n_hetero_dims) != self._n_dim:
            raise ValueError(
                'eta is invalid. Eta cannot be reshaped into (n_ids, n_dim), '
                'even after taking pooled and heterogenuous dimensions into '
                'account.')
        eta_prime = np.zeros(shape=(self._n_ids, self._n_dim))
        eta_prime[:, self._hier_columns()] = eta
        return eta_prime

    def compute_individual_parameters(
            self, parameters, eta, covariates=None, return_eta=False):
        eta = np.asarray(eta, dtype=float)
        if eta.ndim == 1:
            eta = self._shape_eta(eta)
        o_eta, o_psi, _, _ = self._device_eval(
            parameters, eta, None, covariates, False)
        return o_eta if return_eta else o_psi

    def get_covariate_names(self):
        names = []
        for pop_model in self._population_models:
            names += pop_model.get_covariate_names()
        return names

    def get_dim_names(self):
        names = []
        for pop_model in self._population_models:
            names += pop_model.get_dim_names()
        return names

    def get_parameter_names(self, exclude_dim_names=False):
        names = []
        for pop_model in self._population_models:
            names += pop_model.get_parameter_names(exclude_dim_names)
        return names

    def get_population_models(self):
        return self._population_models

    def n_hierarchical_parameters(self, n_ids):
        n_bottom, n_top = 0, 0
        for pop_model in self._population_models:
            n_b, n_t = pop_model.n_hierarchical_parameters(n_ids)
            n_bottom += n_b
            n_top += n_t
        return n_bottom, n_top

    def sample(
            self, parameters, n_samples=None, seed=None, covariates=None,
            *args, **kwargs):
        """chi/_population_models.py:793-872."""
        parameters = np.asarray(parameters)
        if len(parameters) != self._n_parameters:
            raise ValueError(
                'The number of provided parameters does not match the expected'
                ' number of top-level parameters.')
        if self._n_covariates > 0:
            covariates = np.asarray(covariates)
            if covariates.ndim == 1:
                covariates = covariates[np.newaxis, :]
            if covariates.shape[1] != self._n_covariates:
                raise ValueError(
                    'Provided covariates do not match the number of '
                    'covariates.')
        if n_samples is None:
            n_samples = 1
        samples = np.empty(shape=(int(n_samples), self._n_dim))
        rng = np.random.default_rng(seed=seed)
        cov = None
        current_dim = 0
        current_param = 0
        current_cov = 0
        for pop_model in self._population_models:
            end_dim = current_dim + pop_model.n_dim()
            end_param = current_param + pop_model.n_parameters()
            if self._n_covariates > 0:
                end_cov = current_cov + pop_model.n_covariates()
                cov = covariates[:, current_cov:end_cov]
                current_cov = end_cov
            kwargs = {}
            if pop_model.n_covariates() > 0:
                kwargs['covariates'] = cov
            samples[:, current_dim:end_dim] = pop_model.sample(
                parameters=parameters[current_param:end_param],
                n_samples=n_samples, seed=rng, **kwargs)
            current_dim = end_dim
            current_param = end_param
        return samples

    def set_covariate_names(self, names=None):
        current = 0
        for pop_model in self._population_models:
            n = pop_model.n_covariates()
            pop_model.set_covariate_names(
                None if names is None else names[current:current + n])
            current += n

    def set_dim_names(self, names=None):
        if names is None:
            for pop_model in self._population_models:
                pop_model.set_dim_names()
            return None
        if len(names) != self._n_dim:
            raise ValueError(
                'Length of names does not match the number of dimensions.')
        names = [str(label) for label in names]
        current_dim = 0
        for pop_model in self._population_models:
            end_dim = current_dim + pop_model.n_dim()
            pop_model.set_dim_names(names[current_dim:end_dim])
            current_dim = end_dim

    def set_n_ids(self, n_ids):
        """chi/_population_models.py:902-927."""
        n_ids = int(n_ids)
        if n_ids == self._n_ids:
            return None
        for pop_model in self._population_models:
            pop_model.set_n_ids(n_ids)
        self._n_ids = n_ids
        self._set_population_model_properties()
        self._n_bottom, self._n_top_cached = self.n_hierarchical_parameters(
            self._n_ids)

    def set_parameter_names(self, names=None):
        if names is None:
            for pop_model in self._population_models:
                pop_model.set_parameter_names()
            return None
        if len(names) != self._n_parameters:
            raise ValueError(
                'Length of names does not match the number of parameters.')
        current = 0
        for pop_model in self._population_models:
            n = pop_model.n_parameters()
            pop_model.set_parameter_names(names[current:current + n])
            current += n
