"""Population models evaluated on the GPU (interface of
``chi/_population_models.py``).

Every ``compute_*`` method flattens the model into the block descriptor of
the C ABI (``chi_b200_population_set``) and evaluates it with the population
kernels (csrc/core.cu).  ``sample`` draws on the host with numpy's generator
in the reference's order (distribution- and seed-level parity).

Bookkeeping.  A model is a list of LEAVES (``_leaves()``: a Pooled,
Heterogeneous, Gaussian, LogNormal, TruncatedGaussian or Covariate model; a
composed model concatenates the leaves of its parts).  A leaf knows its kind,
its dimensions and how many top-level parameters it has for ``n_ids``
individuals; everything the reference caches and patches up on every change
(number of parameters, special dimensions, hierarchical dimensions, names) is
derived from that list on request.
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

# kinds with a density over the individuals (bottom-level parameters exist)
_DENSITY_KINDS = ('gaussian', 'lognormal', 'truncated_gaussian')


def _default_dim_names(n_dim):
    return ['Dim. %d' % (d + 1) for d in range(n_dim)]


def _as_rows(parameters, n_dim, n_expected=None):
    """Top-level parameters as an array of rows of length ``n_dim``."""
    parameters = np.asarray(parameters)
    if n_expected is not None and parameters.size != n_expected:
        raise ValueError(
            'The number of provided parameters does not match the expected'
            ' number of top-level parameters.')
    if parameters.ndim != 2:
        parameters = parameters.reshape(-1, n_dim)
    return parameters


class PopulationModel(_PopulationModelBase):
    """Base class (interface of chi/_population_models.py:18-326)."""

    KIND = None            # block kind of a leaf
    ROWS = ()              # default names of a leaf's parameter rows
    SPECIAL = None         # True: pooled, False: heterogeneous, None: density

    def __init__(self, n_dim=1, dim_names=None):
        if n_dim < 1:
            raise ValueError(
                'The dimension of the population model has to be greater or '
                'equal to 1.')
        self._n_dim = int(n_dim)
        self._n_ids = 1
        self._device = None
        self._dev_cache = None
        self._custom_names = None
        if dim_names and len(dim_names) != self._n_dim:
            raise ValueError(
                'The number of dimension names has to match the number of '
                'dimensions of the population model.')
        self._dim_names = [str(n) for n in dim_names] if dim_names \
            else _default_dim_names(self._n_dim)

    def __deepcopy__(self, memo):
        clone = self.__class__.__new__(self.__class__)
        memo[id(self)] = clone
        for key, value in self.__dict__.items():
            # (the device problem is rebuilt on demand)
            setattr(clone, key, None if key == '_dev_cache'
                    else copy.deepcopy(value, memo))
        return clone

    # -- the leaf table --------------------------------------------------------
    def _leaves(self):
        return [self]

    def _leaf_n_top(self, n_ids):
        """Top-level parameters of this leaf for ``n_ids`` individuals."""
        return len(self.ROWS) * self._n_dim

    def _leaf_block(self):
        """(kind, n_dim, centered, n_cov, pidx, didx) for the C ABI."""
        return (self.KIND, self._n_dim, getattr(self, '_centered', True), 0,
                None, None)

    def _blocks(self):
        return [leaf._leaf_block() for leaf in self._leaves()]

    def _default_parameter_names(self):
        return [row for row in self.ROWS for _ in range(self._n_dim)]

    @property
    def _parameter_names(self):
        if self._custom_names is not None:
            return self._custom_names
        return self._default_parameter_names()

    @property
    def _n_parameters(self):
        return self.n_parameters()

    # -- device plumbing -----------------------------------------------------
    def _n_top(self, n_ids):
        return self.n_hierarchical_parameters(n_ids)[1]

    def _device_problem(self, n_ids, covariates):
        blocks = self._blocks()
        cov_bytes = None
        if self.n_covariates() > 0:
            if covariates is None:
                raise ValueError(
                    'The population model needs covariates, but no covariates '
                    'have been provided.')
            covariates = np.ascontiguousarray(
                covariates, dtype=float).reshape(n_ids, self.n_covariates())
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
        for info in self.get_special_dims()[0]:
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

    # -- interface of the reference ----------------------------------------------
    def get_covariate_names(self):
        return []

    def get_dim_names(self):
        return list(self._dim_names)

    def get_special_dims(self):
        """``([[dim_start, dim_end, top_start, top_end, is_pooled], ...],
        n_pooled_dims, n_heterogeneous_dims)`` over the leaves
        (chi/_population_models.py:2784, 2187, 451-492)."""
        entries, n_pooled, n_hetero = [], 0, 0
        dim = top = 0
        for leaf in self._leaves():
            n_top = leaf._leaf_n_top(self._n_ids)
            if leaf.SPECIAL is not None:
                entries.append(
                    [dim, dim + leaf._n_dim, top, top + n_top, leaf.SPECIAL])
                if leaf.SPECIAL:
                    n_pooled += leaf._n_dim
                else:
                    n_hetero += leaf._n_dim
            dim += leaf._n_dim
            top += n_top
        return entries, n_pooled, n_hetero

    def get_parameter_names(self, exclude_dim_names=False):
        names = list(self._parameter_names)
        if exclude_dim_names:
            return names
        return [name + ' ' + self._dim_names[k % self._n_dim]
                for k, name in enumerate(names)]

    def n_covariates(self):
        return 0

    def n_dim(self):
        return self._n_dim

    def n_hierarchical_dim(self):
        return sum(leaf._n_dim for leaf in self._leaves()
                   if leaf.SPECIAL is None)

    def n_hierarchical_parameters(self, n_ids):
        n_ids = int(n_ids)
        n_bottom = sum(n_ids * leaf._n_dim for leaf in self._leaves()
                       if leaf.SPECIAL is None)
        return n_bottom, sum(
            leaf._leaf_n_top(n_ids) for leaf in self._leaves())

    def n_ids(self):
        return self._n_ids

    def n_parameters(self):
        return sum(leaf._leaf_n_top(self._n_ids) for leaf in self._leaves())

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        raise NotImplementedError

    def set_covariate_names(self, names=None):
        return None

    def set_dim_names(self, names=None):
        if names is None:
            self._dim_names = _default_dim_names(self._n_dim)
            return
        if len(names) != self._n_dim:
            raise ValueError(
                'Length of names does not match the number of dimensions.')
        self._dim_names = [str(n) for n in names]

    def set_n_ids(self, n_ids):
        n_ids = int(n_ids)
        if n_ids < 1:
            raise ValueError(
                'n_ids is invalid. The number of individuals has to be '
                'positive.')
        self._n_ids = n_ids

    def set_parameter_names(self, names=None):
        if names is None:
            self._custom_names = None
            return
        if len(names) != self.n_parameters():
            raise ValueError(
                'Length of names does not match the number of parameters.')
        self._custom_names = [str(n) for n in names]


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
    """Leaves with a density over the individuals: top-level parameters are
    a location row and a scale row."""

    def __init__(self, n_dim=1, dim_names=None, centered=True):
        super(_DensityModel, self).__init__(n_dim, dim_names)
        self._centered = bool(centered)

    def _location_and_scale(self, parameters):
        rows = _as_rows(parameters, self._n_dim, 2 * self._n_dim)
        return rows[0], rows[1]

    def _sample_parameters(self, parameters):
        return _as_rows(parameters, self._n_dim, 2 * self._n_dim)


class GaussianModel(_DensityModel):
    """chi/_population_models.py:1385-1874."""
    KIND = 'gaussian'
    ROWS = ('Mean', 'Std.')

    def get_mean_and_std(self, parameters):
        return np.vstack(self._location_and_scale(parameters))

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        """Draws individuals -- or, non-centered, their standard-normal etas
        (chi/_population_models.py:1797-1846)."""
        mu, sigma = self._location_and_scale(parameters)
        if not self._centered:
            mu, sigma = np.zeros(self._n_dim), np.ones(self._n_dim)
        if np.any(sigma < 0):
            raise ValueError(
                'A Gaussian distribution only accepts strictly positive '
                'standard deviations.')
        rng = np.random.default_rng(seed=seed)
        return rng.normal(
            loc=mu, scale=sigma, size=(int(n_samples or 1), self._n_dim))


class LogNormalModel(_DensityModel):
    """chi/_population_models.py:2212-2756."""
    KIND = 'lognormal'
    ROWS = ('Log mean', 'Log std.')

    def get_mean_and_std(self, parameters):
        """Mean and standard deviation on the natural scale
        (chi/_population_models.py:2601-2634)."""
        mu, sigma = _as_rows(parameters, self._n_dim)[:2]
        if np.any(sigma < 0):
            raise ValueError('The standard deviation cannot be negative.')
        var = sigma ** 2
        return np.vstack([
            np.exp(mu + var / 2),
            np.sqrt(np.exp(2 * mu + var) * (np.exp(var) - 1))])

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        """chi/_population_models.py:2678-2732."""
        mu, sigma = self._location_and_scale(parameters)
        shape = (int(n_samples or 1), self._n_dim)
        rng = np.random.default_rng(seed=seed)
        if not self._centered:
            return rng.normal(loc=np.zeros(self._n_dim),
                              scale=np.ones(self._n_dim), size=shape)
        if np.any(sigma <= 0):
            raise ValueError(
                'A log-normal distribution only accepts strictly positive '
                'standard deviations.')
        return rng.lognormal(mean=mu, sigma=sigma, size=shape)


class TruncatedGaussianModel(_DensityModel):
    """Gaussian population truncated at zero
    (chi/_population_models.py:3500-3939).  Always centered: the bottom-level
    parameters are the individual parameters themselves; sigma <= 0 or a
    negative individual parameter scores -inf (3557-3558, 3656-3658)."""
    KIND = 'truncated_gaussian'
    ROWS = ('Mu', 'Sigma')

    def __init__(self, n_dim=1, dim_names=None):
        super(TruncatedGaussianModel, self).__init__(
            n_dim, dim_names, centered=True)

    def get_mean_and_std(self, parameters):
        """chi/_population_models.py:3771-3821 (the reference allocates
        (n_parameters, n_dim) and fills two rows; the documented shape is
        returned)."""
        from scipy.stats import norm
        mu, sigma = _as_rows(parameters, self._n_dim)[:2]
        if np.any(sigma < 0):
            raise ValueError('The standard deviation cannot be negative.')
        hazard = norm.pdf(mu / sigma) / (1 - norm.cdf(-mu / sigma))
        return np.vstack([
            mu + sigma * hazard,
            np.sqrt(sigma ** 2 * (1 - mu / sigma * hazard - hazard ** 2))])

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        """chi/_population_models.py:3863-3919.  Like the reference this
        draws ``truncnorm.rvs(a=0, b=inf, loc=mu, scale=sigma)`` from numpy's
        global generator seeded with ``seed`` (scipy's ``a`` is in standard
        units, so the draws are truncated at ``mu``; reproduced as is for
        seed-level parity)."""
        from scipy.stats import truncnorm
        mu, sigma = self._location_and_scale(parameters)
        if np.any(sigma < 0):
            raise ValueError(
                'A truncated Gaussian distribution only accepts strictly '
                'positive standard deviations.')
        if isinstance(seed, np.random.Generator):
            seed = seed.integers(low=0, high=1E6)
        np.random.seed(seed)
        return truncnorm.rvs(
            a=0, b=np.inf, loc=mu, scale=sigma,
            size=(int(n_samples or 1), self._n_dim))


class PooledModel(PopulationModel):
    """Every individual has the same value
    (chi/_population_models.py:2759-3061)."""
    KIND = 'pooled'
    ROWS = ('Pooled',)
    SPECIAL = True

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        rows = _as_rows(parameters, self._n_dim, self._n_dim)
        if n_samples is None:
            return rows
        return np.broadcast_to(rows, shape=(int(n_samples), self._n_dim))


class HeterogeneousModel(PopulationModel):
    """Every individual has its own value, no density
    (chi/_population_models.py:1877-2209)."""
    KIND = 'heterogeneous'
    SPECIAL = False

    def __init__(self, n_dim=1, dim_names=None, n_ids=1):
        super(HeterogeneousModel, self).__init__(n_dim, dim_names)
        self.set_n_ids(n_ids)

    def _leaf_n_top(self, n_ids):
        return int(n_ids) * self._n_dim

    def _default_parameter_names(self):
        return ['ID %d' % (i + 1) for i in range(self._n_ids)
                for _ in range(self._n_dim)]

    def _check_n_ids(self, n_ids):
        if n_ids != self._n_ids:
            raise ValueError(
                'The number of individuals does not match n_ids of the '
                'heterogeneous model.')

    def sample(self, parameters, n_samples=None, seed=None, *args, **kwargs):
        """A bootstrap sample of the individuals
        (chi/_population_models.py:2113-2152)."""
        rows = _as_rows(parameters, self._n_dim)
        rng = np.random.default_rng(seed=seed)
        picks = rng.choice(
            np.arange(self._n_ids), size=n_samples if n_samples else 1,
            replace=True)
        return rows[picks]

    def set_n_ids(self, n_ids, no_shortcut=False):
        n_ids = int(n_ids)
        if n_ids < 1:
            raise ValueError(
                'The number of modelled individuals needs to be greater or '
                'equal to 1.')
        if n_ids != self._n_ids:
            # the parameters are per individual: names start over
            self._custom_names = None
            self._dev_cache = None
        self._n_ids = n_ids


class CovariatePopulationModel(PopulationModel):
    """A Gaussian or LogNormal population whose parameters are shifted by
    the individuals' covariates (chi/_population_models.py:956-1382)."""

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
        if population_model.KIND not in ('gaussian', 'lognormal'):
            raise TypeError(
                'chi_b200 supports covariate models over GaussianModel and '
                'LogNormalModel only.')
        super(CovariatePopulationModel, self).__init__(
            population_model.n_dim())
        self._population_model = copy.deepcopy(population_model)
        self._covariate_model = copy.deepcopy(covariate_model)
        self._population_model.set_dim_names(dim_names)
        # by default every (parameter row, dimension) pair is shifted
        n_rows = self._population_model.n_parameters() // self._n_dim
        self._covariate_model.set_population_parameters(
            [[row, dim] for dim in range(self._n_dim)
             for row in range(n_rows)])
        self._name_the_shifts()

    def _name_the_shifts(self):
        """A shift carries the name of the population parameter it acts on,
        once per covariate."""
        names = np.array(self._population_model.get_parameter_names())
        names = names.reshape(-1, self._n_dim)
        pidx, didx = self._covariate_model.get_set_population_parameters()
        n_cov = self._covariate_model.n_covariates()
        self._covariate_model.set_parameter_names(
            [name for name in names[pidx, didx] for _ in range(n_cov)])

    @property
    def _n_pop(self):
        return self._population_model.n_parameters()

    @property
    def KIND(self):
        return self._population_model.KIND

    def _leaf_n_top(self, n_ids):
        return self._n_pop + self._covariate_model.n_parameters()

    def _leaf_block(self):
        pidx, didx = self._covariate_model.get_set_population_parameters()
        inner = self._population_model
        return (inner.KIND, inner._n_dim, inner._centered,
                self._covariate_model.n_covariates(), list(pidx), list(didx))

    def get_covariate_names(self):
        return self._covariate_model.get_covariate_names()

    def get_dim_names(self):
        return self._population_model.get_dim_names()

    def get_parameter_names(self, exclude_dim_names=False):
        return self._population_model.get_parameter_names(
            exclude_dim_names) + self._covariate_model.get_parameter_names()

    def n_covariates(self):
        return self._covariate_model.n_covariates()

    def sample(self, parameters, covariates, n_samples=None, seed=None):
        """One draw per individual from the population shifted by that
        individual's covariates (chi/_population_models.py:1247-1299; host
        RNG, the handful of shifts is applied with the indexing the kernels
        use)."""
        n_cov = self.n_covariates()
        covariates = np.array(covariates)
        if covariates.ndim == 1:
            covariates = covariates[np.newaxis, :]
        if covariates.shape[1] != n_cov:
            raise ValueError(
                'Provided covariates do not match the number of covariates.')
        n_samples = int(n_samples or 1)
        covariates = np.broadcast_to(covariates, (n_samples, n_cov))
        parameters = np.asarray(parameters)
        base = parameters[:self._n_pop].reshape(-1, self._n_dim)
        pidx, didx = self._covariate_model.get_set_population_parameters()
        beta = parameters[self._n_pop:].reshape(len(pidx), n_cov)
        shifted = np.repeat(base[np.newaxis], n_samples, axis=0)
        shifted[:, pidx, didx] += covariates @ beta.T
        rng = np.random.default_rng(seed)
        return np.vstack([
            self._population_model.sample(rows, n_samples=1, seed=rng)
            for rows in shifted])

    def set_covariate_names(self, names=None):
        self._covariate_model.set_covariate_names(names)

    def set_dim_names(self, names=None):
        self._population_model.set_dim_names(names)
        self._name_the_shifts()

    def set_parameter_names(self, names=None):
        if names is None:
            self._population_model.set_parameter_names()
            self._covariate_model.set_parameter_names()
            return
        self._population_model.set_parameter_names(names[:self._n_pop])
        self._covariate_model.set_parameter_names(names[self._n_pop:])

    def set_population_parameters(self, indices):
        """Which (parameter row, dimension) pairs the covariates shift
        (chi/_population_models.py:1354-1382)."""
        indices = np.array(indices)
        n_rows = self._n_pop // self._n_dim
        if np.min(indices) < 0 or np.max(indices[:, 0]) >= n_rows or \
                np.max(indices[:, 1]) >= self._n_dim:
            raise IndexError('The provided indices are out of bounds.')
        self._covariate_model.set_population_parameters(indices)
        # (named in the order the indices were given, as the reference does)
        names = np.array(self._population_model.get_parameter_names())
        names = names.reshape(n_rows, self._n_dim)
        n_cov = self._covariate_model.n_covariates()
        self._covariate_model.set_parameter_names(
            [name for name in names[indices[:, 0], indices[:, 1]]
             for _ in range(n_cov)])
        self._dev_cache = None


class ComposedPopulationModel(PopulationModel):
    """Population models over consecutive dimensions
    (chi/_population_models.py:329-953)."""

    def __init__(self, population_models):
        for part in population_models:
            if not isinstance(part, PopulationModel):
                raise TypeError(
                    'The population models have to be instances of '
                    'chi.PopulationModel.')
        n_ids = 1
        for part in population_models:
            if n_ids > 1 and part.n_ids() > 1 and n_ids != part.n_ids():
                raise ValueError(
                    'All population models must model the same number of '
                    'individuals.')
            n_ids = n_ids if n_ids > 1 else part.n_ids()
        self._population_models = population_models
        self._n_ids = n_ids
        self._device = None
        self._dev_cache = None
        self._custom_names = None
        names = self.get_parameter_names()
        if len(set(names)) != len(names):
            # ambiguous names: number the dimensions across the parts
            self.set_dim_names(_default_dim_names(self._n_dim))

    @property
    def _n_dim(self):
        return sum(part.n_dim() for part in self._population_models)

    def _leaves(self):
        return [leaf for part in self._population_models
                for leaf in part._leaves()]

    def _parts(self, attribute):
        """(part, start, end) along ``n_dim`` / ``n_parameters`` /
        ``n_covariates``."""
        start = 0
        for part in self._population_models:
            end = start + getattr(part, attribute)()
            yield part, start, end
            start = end

    def _check_n_ids(self, n_ids):
        for part in self._population_models:
            part._check_n_ids(n_ids)

    def _shape_eta(self, eta):
        """A flat eta of the hierarchical dimensions only becomes
        ``(n_ids, n_dim)`` with zeros in the pooled / heterogeneous columns
        (chi/_population_models.py:534-565)."""
        eta = eta.reshape(self._n_ids, -1)
        n_dim = self._n_dim
        if eta.shape[1] == n_dim:
            return eta
        columns = self._hier_columns()
        if eta.shape[1] != len(columns):
            raise ValueError(
                'eta is invalid. Eta cannot be reshaped into (n_ids, n_dim), '
                'even after taking pooled and heterogenuous dimensions into '
                'account.')
        full = np.zeros((self._n_ids, n_dim))
        full[:, columns] = eta
        return full

    def compute_individual_parameters(
            self, parameters, eta, covariates=None, return_eta=False):
        eta = np.asarray(eta, dtype=float)
        if eta.ndim == 1:
            eta = self._shape_eta(eta)
        o_eta, o_psi, _, _ = self._device_eval(
            parameters, eta, None, covariates, False)
        return o_eta if return_eta else o_psi

    def get_covariate_names(self):
        return [n for part in self._population_models
                for n in part.get_covariate_names()]

    def get_dim_names(self):
        return [n for part in self._population_models
                for n in part.get_dim_names()]

    def get_parameter_names(self, exclude_dim_names=False):
        return [n for part in self._population_models
                for n in part.get_parameter_names(exclude_dim_names)]

    def get_population_models(self):
        return self._population_models

    def n_covariates(self):
        return sum(part.n_covariates() for part in self._population_models)

    def sample(
            self, parameters, n_samples=None, seed=None, covariates=None,
            *args, **kwargs):
        """Part by part on one generator
        (chi/_population_models.py:793-872)."""
        parameters = np.asarray(parameters)
        if len(parameters) != self.n_parameters():
            raise ValueError(
                'The number of provided parameters does not match the expected'
                ' number of top-level parameters.')
        n_cov = self.n_covariates()
        if n_cov > 0:
            covariates = np.asarray(covariates)
            if covariates.ndim == 1:
                covariates = covariates[np.newaxis, :]
            if covariates.shape[1] != n_cov:
                raise ValueError(
                    'Provided covariates do not match the number of '
                    'covariates.')
        n_samples = 1 if n_samples is None else n_samples
        samples = np.empty((int(n_samples), self._n_dim))
        rng = np.random.default_rng(seed=seed)
        dims = self._parts('n_dim')
        tops = self._parts('n_parameters')
        covs = self._parts('n_covariates')
        for (part, d0, d1), (_, t0, t1), (_, c0, c1) in zip(dims, tops, covs):
            kwargs = {'covariates': covariates[:, c0:c1]} if c1 > c0 else {}
            samples[:, d0:d1] = part.sample(
                parameters=parameters[t0:t1], n_samples=n_samples, seed=rng,
                **kwargs)
        return samples

    def set_covariate_names(self, names=None):
        for part, c0, c1 in self._parts('n_covariates'):
            part.set_covariate_names(None if names is None else names[c0:c1])

    def set_dim_names(self, names=None):
        if names is not None and len(names) != self._n_dim:
            raise ValueError(
                'Length of names does not match the number of dimensions.')
        for part, d0, d1 in self._parts('n_dim'):
            part.set_dim_names(
                None if names is None else [str(n) for n in names[d0:d1]])

    def set_n_ids(self, n_ids):
        n_ids = int(n_ids)
        if n_ids == self._n_ids:
            return
        for part in self._population_models:
            part.set_n_ids(n_ids)
        self._n_ids = n_ids
        self._dev_cache = None

    def set_parameter_names(self, names=None):
        if names is not None and len(names) != self.n_parameters():
            raise ValueError(
                'Length of names does not match the number of parameters.')
        for part, t0, t1 in self._parts('n_parameters'):
            part.set_parameter_names(None if names is None else names[t0:t1])
