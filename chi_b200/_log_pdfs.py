"""Log-likelihoods and log-posteriors evaluated on the GPU.

Host-side mirror of the hot-path classes of ``chi/_log_pdfs.py`` (seam B3,
SURVEY.md section 8b): ``LogLikelihood``, ``HierarchicalLogLikelihood``,
``LogPosterior`` and ``HierarchicalLogPosterior`` with the pints ``LogPDF``
protocol (``__call__``, ``evaluateS1``, ``n_parameters``) plus the batched
entry points ``evaluate_batch`` / ``evaluateS1_batch`` that evaluate many
parameter vectors (MCMC chains, optimiser populations) in one launch.

Where the reference loops over individuals in Python and crosses into a
compiled CVODES module once per individual (chi/_log_pdfs.py:289-292), a
hierarchical log-likelihood here owns ONE device-resident problem holding
every individual's observations and dosing regimen, and one evaluation is
K3a -> K1+K2 -> K3b -> K3c on the GPU (csrc/core.cu: run_hier).
"""
import copy
import warnings

import numpy as np

from chi_b200 import _lib
from chi_b200._error_models import ErrorModel, ReducedErrorModel
from chi_b200._mechanistic_models import (
    MechanisticModel, SBMLModel, ReducedMechanisticModel, _events_of)
from chi_b200._population_models import (
    PopulationModel, ComposedPopulationModel, set_population)

try:
    import pints as _pints
    _LogPDFBase = _pints.LogPDF
    _LogPriorBase = _pints.LogPrior
except Exception:  # pragma: no cover
    _pints = None
    _LogPDFBase = object
    _LogPriorBase = None

try:
    import chi as _chi
except Exception:  # pragma: no cover
    _chi = None


def _chi_base(name, fallback):
    """chi's own class of that name when chi is importable, so that
    ``isinstance(x, chi.<name>)`` holds for the classes below without any
    change to chi (chi/_inference.py:423-429 checks the log-posterior that
    way; chi's ``__init__`` is never called, every method is overridden)."""
    base = getattr(_chi, name, None) if _chi is not None else None
    return base if isinstance(base, type) else fallback


def _vector(x):
    """pints.vector: read-only 1-d float copy."""
    x = np.array(x, dtype=float, copy=True).reshape(-1)
    x.setflags(write=False)
    return x


def _failure_warning(status):
    warnings.warn(
        'An error occured while solving the mechanistic model: \n'
        'integrator status %s.\n A score of -infinity is returned.'
        % str(status), RuntimeWarning)


class _DeviceLikelihoods(object):
    """Device-resident data of one or many structurally identical
    individual log-likelihoods (one ``chi_b200_problem``)."""

    def __init__(self, mechanistic_model, error_kinds, times, observations,
                 regimens, device=None):
        """times / observations: per individual, per output arrays."""
        if isinstance(mechanistic_model, ReducedMechanisticModel):
            # (LogLikelihood passes the wrapped model and its fixed entries)
            mechanistic_model = mechanistic_model.mechanistic_model()
        if not isinstance(mechanistic_model, SBMLModel):
            raise TypeError(
                'Device log-likelihoods need a chi_b200.SBMLModel / '
                'PKPDModel mechanistic model.')
        self.model = mechanistic_model
        code = mechanistic_model.model_code()
        self.n_mech = mechanistic_model.n_parameters()
        self.error_kinds = list(error_kinds)
        self.n_ids = len(times)
        self.device = device
        out_ids = [code.output_names.index(o)
                   for o in mechanistic_model._output_names]
        kinds = [_lib.ERR_KINDS[k] for k in self.error_kinds]
        self.n_sigma = sum(2 if k == 'cam' else 1 for k in self.error_kinds)
        self.n_dim = self.n_mech + self.n_sigma

        lib = _lib.lib()
        self.problem = _lib.Problem(_lib.find_model(code), device)
        h = self.problem.handle
        _, p_out = _lib.i32(out_ids)
        _, p_kind = _lib.i32(kinds)
        _lib.check(lib.chi_b200_problem_set_outputs(
            h, len(out_ids), p_out, p_kind))
        atol, rtol, max_steps = mechanistic_model.tolerance()
        _lib.check(lib.chi_b200_problem_set_tolerances(
            h, rtol, atol, max_steps))
        if mechanistic_model.kernel_variant() is not None:
            _lib.check(lib.chi_b200_problem_set_variant(
                h, mechanistic_model.kernel_variant()))

        # records: per individual the (time, output, value) triples sorted by
        # time -- the union grid + masks of chi/_log_pdfs.py:845-883 in
        # ragged form
        rec_off = [0]
        rec_t, rec_obs, rec_out = [], [], []
        dose_off = [0]
        events = []
        self.n_obs = []
        for i in range(self.n_ids):
            ts = np.concatenate(
                [np.asarray(t, dtype=float) for t in times[i]])
            ys = np.concatenate(
                [np.asarray(y, dtype=float) for y in observations[i]])
            os_ = np.concatenate(
                [np.full(len(t), o, dtype=np.int32)
                 for o, t in enumerate(times[i])])
            order = np.lexsort((os_, ts))
            rec_t.append(ts[order])
            rec_obs.append(ys[order])
            rec_out.append(os_[order])
            rec_off.append(rec_off[-1] + len(ts))
            self.n_obs.append(len(ts))
            ev = _events_of(regimens[i] if regimens is not None else None)
            events.append(ev)
            dose_off.append(dose_off[-1] + len(ev))
        cat = np.concatenate
        self._rec_off = np.array(rec_off, dtype=np.int64)
        self._rec_out = rec_out      # per individual: output of each record
        a1, p1 = _lib.i64(self._rec_off)
        a2, p2 = _lib.f64(cat(rec_t) if rec_off[-1] else np.zeros(0))
        with np.errstate(all='ignore'):
            a3, p3 = _lib.f64(cat(rec_obs) if rec_off[-1] else np.zeros(0))
        a4, p4 = _lib.i32(cat(rec_out) if rec_off[-1] else np.zeros(0))
        a5, p5 = _lib.i64(dose_off)
        a6, p6 = _lib.f64(cat(events).reshape(-1) if dose_off[-1]
                          else np.zeros(0))
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            _lib.check(lib.chi_b200_problem_upload(
                h, self.n_ids, p1, p2, p3, p4, p5, p6))

    def set_sensitivity_columns(self, columns):
        _, p = _lib.i32(columns)
        _lib.check(_lib.lib().chi_b200_problem_set_sensitivity_columns(
            self.problem.handle, len(columns), p))

    def set_snapshot_budget(self, n_bytes):
        """Device memory the deferred observation may use
        (``chi_b200_problem_set_snapshot_budget``): 0 keeps the error model
        fused in the solve kernel, -1 restores the default."""
        _lib.check(_lib.lib().chi_b200_problem_set_snapshot_budget(
            self.problem.handle, int(n_bytes)))

    def set_fixed(self, mask, values):
        """Holds entries of the system vector [psi | sigma] at fixed values
        (LogLikelihood.fix_parameters).  Sensitivities are then only
        integrated for the free mechanistic parameters."""
        mask = np.asarray(mask, dtype=bool)
        idx = np.flatnonzero(mask).astype(np.int32)
        a_i, p_i = _lib.i32(idx)
        a_v, p_v = _lib.f64(np.asarray(values, dtype=float)[mask])
        _lib.check(_lib.lib().chi_b200_problem_set_fixed_parameters(
            self.problem.handle, len(idx), p_i, p_v))
        free_mech = [j for j in range(self.n_mech) if not mask[j]]
        if free_mech:
            self.set_sensitivity_columns(free_mech)
        self.fixed_mask = mask

    def simulate(self, psi, ids=None, with_sens=False):
        """Batched ``simulate``: psi (B, n_mech) -> model output at every
        record of every system (concatenated) [and sensitivities]."""
        psi = np.ascontiguousarray(psi, dtype=float).reshape(-1, self.n_mech)
        B = psi.shape[0]
        idv = np.arange(B) % self.n_ids if ids is None else np.asarray(ids)
        rows = int(np.sum(self._rec_off[idv + 1] - self._rec_off[idv]))
        lib = _lib.lib()
        a_x, p_x = _lib.f64(psi)
        p_ids = None
        if ids is not None:
            a_i, p_ids = _lib.i32(ids)
        y, p_y = _lib.out_f64(max(rows, 1))
        s, p_s = _lib.out_f64((max(rows, 1), self.n_mech))
        status, p_st = _lib.out_i32(B)
        _lib.check(lib.chi_b200_simulate(
            self.problem.handle, B, p_ids, p_x, 1 if with_sens else 0, p_y,
            p_s if with_sens else None, p_st))
        return y[:rows], s[:rows], status

    def loglik(self, x, ids=None, with_grad=False):
        """x: (B, n_dim) -> ll (B,), grad (B, n_dim), status (B,)."""
        x = np.ascontiguousarray(x, dtype=float).reshape(-1, self.n_dim)
        B = x.shape[0]
        lib = _lib.lib()
        a_x, p_x = _lib.f64(x)
        p_ids = None
        if ids is not None:
            a_i, p_ids = _lib.i32(ids)
        ll, p_ll = _lib.out_f64(B)
        grad, p_g = _lib.out_f64((B, self.n_dim))
        status, p_st = _lib.out_i32(B)
        _lib.check(lib.chi_b200_loglik(
            self.problem.handle, B, p_ids, p_x, 1 if with_grad else 0, p_ll,
            p_g if with_grad else None, p_st, None, None))
        return ll, grad, status


class LogLikelihood(_chi_base('LogLikelihood', _LogPDFBase)):
    """Log-likelihood of one individual, see chi/_log_pdfs.py:637-1149."""

    def __init__(
            self, mechanistic_model, error_model, observations, times,
            outputs=None):
        if not isinstance(mechanistic_model, MechanisticModel):
            raise TypeError(
                'The mechanistic model as to be an instance of a '
                'chi.MechanisticModel.')
        if not isinstance(error_model, list):
            error_model = [error_model]
        mechanistic_model = mechanistic_model.copy()
        if outputs is not None:
            mechanistic_model.set_outputs(outputs)
        n_outputs = mechanistic_model.n_outputs()
        if len(error_model) != n_outputs:
            raise ValueError(
                'One error model has to be provided for each mechanistic '
                'model output.')
        for em in error_model:
            if not isinstance(em, (ErrorModel, ReducedErrorModel)):
                raise TypeError(
                    'The error models have to instances of a '
                    'chi.ErrorModel.')
        if n_outputs == 1:
            if len(observations) != n_outputs:
                observations = [observations]
            if len(times) != n_outputs:
                times = [times]
        if len(observations) != n_outputs:
            raise ValueError(
                'The observations have the wrong length. For a '
                'multi-output problem the observations are expected to be '
                'a list of array-like objects with measurements for each '
                'of the mechanistic model outputs.')
        if len(times) != n_outputs:
            raise ValueError(
                'The times have the wrong length. For a multi-output problem '
                'the times are expected to be a list of array-like objects '
                'with the measurement time points for each of the mechanistic '
                'model outputs.')
        observations = [_vector(obs) for obs in observations]
        times = [_vector(ts) for ts in times]
        for ts in times:
            if np.any(ts < 0):
                raise ValueError('Times cannot be negative.')
            if np.any(ts[:-1] > ts[1:]):
                raise ValueError('Times must be increasing.')
        for output_id, output_times in enumerate(times):
            if observations[output_id].shape != output_times.shape:
                raise ValueError(
                    'The observations and times have to be of the same '
                    'dimension.')
            order = np.argsort(output_times, kind='stable')
            times[output_id] = output_times[order]
            observations[output_id] = observations[output_id][order]
        error_model = [copy.deepcopy(em) for em in error_model]

        self._mechanistic_model = mechanistic_model
        self._error_models = error_model
        self._observations = observations
        self._output_times = times
        self._n_obs = [len(obs) for obs in observations]
        self._set_number_and_parameter_names()
        self._unwrap_reduced_models()
        self._id = None
        self._device = None

    def _unwrap_reduced_models(self):
        """The device works on the FULL vector [all mechanistic parameters |
        all error parameters] of the wrapped models.  Entries that a
        ReducedMechanisticModel (chi/_mechanistic_models.py:1160-1186) or a
        ReducedErrorModel (chi/_error_models.py:1704-1750) holds fixed are
        locked into the fixed set here; they are not parameters of this
        log-likelihood and ``fix_parameters`` cannot release them."""
        model = self._mechanistic_model
        locked = {}
        self._base_model = model
        if isinstance(model, ReducedMechanisticModel):
            self._base_model = model.mechanistic_model()
            locked.update(model.fixed_parameters().values)
        full_names = list(self._base_model.parameters())
        n_mech = len(full_names)
        self._error_kinds = []
        for output_id, em in enumerate(self._error_models):
            inner = em
            if isinstance(em, ReducedErrorModel):
                inner = em.get_error_model()
                if em._fixed_mask is not None:
                    for k in np.flatnonzero(em._fixed_mask):
                        locked[len(full_names) + int(k)] = float(
                            em._fixed_values[k])
                names = list(em._parameter_names)
            else:
                names = inner.get_parameter_names()
            self._error_kinds.append(inner._kind)
            full_names += names
        self._full_parameter_names = full_names
        self._n_mechanistic_params = n_mech
        self._locked = dict(locked)
        self._fixed_mask = None
        self._fixed_values = None
        if locked:
            self._fixed_mask = np.zeros(len(full_names), dtype=bool)
            self._fixed_values = np.zeros(len(full_names))
            for position, value in locked.items():
                self._fixed_mask[position] = True
                self._fixed_values[position] = value

    def fix_parameters(self, name_value_dict):
        """Fixes model parameters at the given values and removes them from
        the parameter vector; ``None`` frees a parameter again
        (chi/_log_pdfs.py:1029-1073).  Sensitivities are computed for the
        free parameters only."""
        try:
            name_value_dict = dict(name_value_dict)
        except (TypeError, ValueError):
            raise ValueError(
                'The name-value dictionary has to be convertable to a python '
                'dictionary.')
        n_full = len(self._full_parameter_names)
        if self._fixed_mask is None:
            self._fixed_mask = np.zeros(n_full, dtype=bool)
            self._fixed_values = np.zeros(n_full)
        for index, name in enumerate(self._full_parameter_names):
            if name not in name_value_dict or index in self._locked:
                continue
            value = name_value_dict[name]
            if value is None:
                self._fixed_mask[index] = False
                continue
            self._fixed_mask[index] = True
            self._fixed_values[index] = float(value)
        if not np.any(self._fixed_mask):
            self._fixed_mask = None
            self._fixed_values = None
        free = np.ones(n_full, dtype=bool) if self._fixed_mask is None \
            else ~self._fixed_mask
        self._parameter_names = [
            n for n, f in zip(self._full_parameter_names, free) if f]
        self._n_parameters = len(self._parameter_names)
        self._device = None

    def n_fixed_parameters(self):
        return 0 if self._fixed_mask is None else int(
            np.sum(self._fixed_mask))

    def _expand(self, parameters):
        parameters = np.asarray(parameters, dtype=float).reshape(-1)
        if len(parameters) != self._n_parameters:
            raise ValueError(
                'The number of parameters does not match n_parameters.')
        if self._fixed_mask is None:
            return parameters
        full = self._fixed_values.copy()
        full[~self._fixed_mask] = parameters
        return full

    # -- device ----------------------------------------------------------------
    def _regimen(self):
        try:
            return self._mechanistic_model.dosing_regimen()
        except AttributeError:
            return None

    def _get_device(self):
        if self._device is None:
            self._device = _DeviceLikelihoods(
                self._base_model, self._error_kinds,
                [self._output_times], [self._observations],
                [self._regimen()])
            if self._fixed_mask is not None:
                n_mech = self._n_mechanistic_params
                free_mech = [j for j in range(n_mech)
                             if not self._fixed_mask[j]]
                if free_mech:
                    self._device.set_sensitivity_columns(free_mech)
        return self._device

    def __call__(self, parameters):
        """chi/_log_pdfs.py:802-843."""
        full = self._expand(parameters)
        ll, _, status = self._get_device().loglik(full[None, :])
        if status[0] != 0:
            _failure_warning(status[0])
            return -np.inf
        return float(ll[0])

    def evaluateS1(self, parameters):
        """chi/_log_pdfs.py:973-1027."""
        full = self._expand(parameters)
        ll, grad, status = self._get_device().loglik(
            full[None, :], with_grad=True)
        if status[0] != 0:
            _failure_warning(status[0])
            return -np.inf, np.full(
                shape=self._n_parameters, fill_value=np.inf)
        grad = grad[0]
        if self._fixed_mask is not None:
            grad = grad[~self._fixed_mask]
        return float(ll[0]), grad.copy()

    def compute_pointwise_ll(self, parameters):
        """Log-likelihood of every observation, output by output
        (chi/_log_pdfs.py:932-971): one forward launch of the integrator for
        the model outputs at the records, then the error model of each
        output on the device (``chi_b200_error_model`` with its pointwise
        result)."""
        full = self._expand(parameters)
        device = self._get_device()
        n_mech = self._n_mechanistic_params
        ybar, _, status = device.simulate(full[None, :n_mech])
        if status[0] != 0:
            _failure_warning(status[0])
            return np.full(int(np.sum(self._n_obs)), -np.inf)
        rec_out = device._rec_out[0]
        start = n_mech
        pointwise = []
        for output_id, error_model in enumerate(self._error_models):
            inner = error_model.get_error_model() if isinstance(
                error_model, ReducedErrorModel) else error_model
            end = start + inner.n_parameters()
            pointwise.append(inner.compute_pointwise_ll(
                parameters=full[start:end],
                model_output=ybar[rec_out == output_id],
                observations=self._observations[output_id]))
            start = end
        return np.hstack(pointwise)

    def _set_number_and_parameter_names(self):
        """Names of the mechanistic parameters, then of every output's error
        parameters (chi/_log_pdfs.py:885-930)."""
        from chi_b200._predictive_models import name_parameters
        self._parameter_names = name_parameters(
            self._mechanistic_model, self._error_models)
        self._n_parameters = len(self._parameter_names)
        self._n_mechanistic_params = self._mechanistic_model.n_parameters()
        self._n_error_params = [
            em.n_parameters() for em in self._error_models]

    def get_id(self, *args, **kwargs):
        return self._id

    def get_parameter_names(self, *args, **kwargs):
        return copy.copy(self._parameter_names)

    def get_submodels(self):
        return dict({
            'Mechanistic model': self._mechanistic_model.copy(),
            'Error models': copy.copy(self._error_models)})

    def n_parameters(self, *args, **kwargs):
        return self._n_parameters

    def n_observations(self):
        return self._n_obs

    def set_id(self, label):
        if isinstance(label, float):
            label = int(label)
        self._id = str(label)


def fill_owned_columns(grad, rows, value, shard, n_ids, n_bottom):
    """``grad[rows] = value`` restricted to the columns a shard of
    individuals owns: its block of the bottom level and the top level.  The
    page-locked result buffers are reused across evaluations and only the
    owned columns are rewritten by the library, so nothing may ever be
    written into the other individuals' columns (they have to stay zero)."""
    begin, end = shard
    if (begin, end) == (0, n_ids):
        grad[rows] = value
        return
    n_h = n_bottom // n_ids
    grad[rows, begin * n_h:end * n_h] = value
    grad[rows, n_bottom:] = value


def _same_fixed(a, b):
    if a._fixed_mask is None or b._fixed_mask is None:
        return a._fixed_mask is None and b._fixed_mask is None
    return np.array_equal(a._fixed_mask, b._fixed_mask) and np.array_equal(
        a._fixed_values[a._fixed_mask], b._fixed_values[b._fixed_mask])


def _check_structure(log_likelihoods):
    first = log_likelihoods[0]
    key = first._base_model.model_key()
    outs = first._base_model._output_names
    kinds = first._error_kinds
    for ll in log_likelihoods[1:]:
        if ll._base_model.model_key() != key or \
                ll._base_model._output_names != outs or \
                ll._error_kinds != kinds or \
                not _same_fixed(first, ll):
            raise ValueError(
                'The log-likelihoods of a hierarchical model have to be '
                'structurally identical (same mechanistic model, outputs and '
                'error models).')


class HierarchicalLogLikelihood(
        _chi_base('HierarchicalLogLikelihood', object)):
    """chi/_log_pdfs.py:18-403."""

    def __init__(
            self, log_likelihoods, population_model, covariates=None):
        for log_likelihood in log_likelihoods:
            if not isinstance(log_likelihood, LogLikelihood):
                raise ValueError(
                    'The log-likelihoods have to be instances of a '
                    'chi.LogLikelihood.')
        if not isinstance(population_model, PopulationModel):
            raise ValueError(
                'The population model has to be an instances of '
                'chi.PopulationModel.')
        n_parameters = population_model.n_dim()
        for log_likelihood in log_likelihoods:
            if log_likelihood.n_parameters() != n_parameters:
                raise ValueError(
                    'The dimension of the population model does not match the '
                    'the dimensions of the log-likelihood parameter space.')
        n_ids = len(log_likelihoods)
        if population_model.n_covariates() > 0:
            if covariates is None:
                raise ValueError(
                    'The population model needs covariates, but no covariates '
                    'have been provided.')
            n_cov = population_model.n_covariates()
            try:
                covariates = np.asarray(covariates).reshape(n_ids, n_cov)
            except ValueError:
                raise ValueError(
                    'The covariates do not have the correct format. '
                    'Covariates need to be of shape (n_ids, n_cov).')
        _check_structure(log_likelihoods)
        self._log_likelihoods = log_likelihoods
        self._label_log_likelihoods()
        first = log_likelihoods[0]
        device = _DeviceLikelihoods(
            first._base_model, first._error_kinds,
            [ll._output_times for ll in log_likelihoods],
            [ll._observations for ll in log_likelihoods],
            [ll._regimen() for ll in log_likelihoods])
        if first._fixed_mask is not None:
            device.set_fixed(first._fixed_mask, first._fixed_values)
        self._setup(device, population_model, covariates,
                    first.get_parameter_names(),
                    [ll.get_id() for ll in log_likelihoods])

    @classmethod
    def from_arrays(cls, mechanistic_model, error_models, times,
                    observations, population_model, regimens=None,
                    covariates=None, ids=None, fixed_parameters=None):
        """Builds the hierarchical log-likelihood straight from ragged
        arrays (one entry per individual; single output: arrays, multiple
        outputs: lists of arrays) without creating one Python object per
        individual -- the ingestion path for 10^4-10^5 individuals."""
        if not isinstance(error_models, list):
            error_models = [error_models]
        n_out = mechanistic_model.n_outputs()
        if len(error_models) != n_out:
            raise ValueError(
                'One error model has to be provided for each mechanistic '
                'model output.')
        n_ids = len(times)
        if n_out == 1:
            times = [[t] if np.ndim(t[0]) == 0 else t for t in times]
            observations = [[y] if np.ndim(y[0]) == 0 else y
                            for y in observations]
        self = cls.__new__(cls)
        self._log_likelihoods = None
        template = LogLikelihood(
            mechanistic_model, error_models, observations[0], times[0])
        device = _DeviceLikelihoods(
            template._base_model, template._error_kinds, times, observations,
            regimens)
        if fixed_parameters:
            template.fix_parameters(fixed_parameters)
        if template._fixed_mask is not None:
            device.set_fixed(template._fixed_mask, template._fixed_values)
        if population_model.n_covariates() > 0:
            if covariates is None:
                raise ValueError(
                    'The population model needs covariates, but no covariates '
                    'have been provided.')
            covariates = np.asarray(covariates).reshape(
                n_ids, population_model.n_covariates())
        if ids is None:
            ids = ['Log-likelihood %d' % (k + 1) for k in range(n_ids)]
        self._setup(device, population_model, covariates,
                    template.get_parameter_names(), [str(i) for i in ids])
        return self

    def _setup(self, device, population_model, covariates, ll_names, ids):
        self._device = device
        self._population_model = population_model
        self._covariates = covariates
        self._n_ids = device.n_ids
        self._n_dim = population_model.n_dim()
        self._ll_parameter_names = list(ll_names)
        self._ids = list(ids)
        n_free = device.n_dim - int(np.sum(
            getattr(device, 'fixed_mask', np.zeros(1, dtype=bool))))
        if self._n_dim != n_free:
            raise ValueError(
                'The dimension of the population model does not match the '
                'the dimensions of the log-likelihood parameter space.')
        population_model.set_n_ids(self._n_ids)
        self._n_parameters = int(np.sum(
            population_model.n_hierarchical_parameters(self._n_ids)))
        self._n_bottom = self._n_parameters - \
            population_model.n_hierarchical_parameters(self._n_ids)[1]
        self._n_obs = list(device.n_obs)
        set_population(device.problem, population_model._blocks(), covariates)
        sizes = np.zeros(4, dtype=np.int64)
        _lib.check(_lib.lib().chi_b200_population_sizes(
            device.problem.handle, sizes.ctypes.data_as(_lib.c_i64p)))
        assert sizes[0] == self._n_bottom, (sizes, self._n_bottom)
        assert sizes[0] + sizes[1] == self._n_parameters

    def _label_log_likelihoods(self):
        """chi/_log_pdfs.py:147-161."""
        ids = []
        for idl, log_likelihood in enumerate(self._log_likelihoods):
            _id = log_likelihood.get_id()
            if not _id:
                _id = 'Log-likelihood %d' % (idl + 1)
            if _id in ids:
                raise ValueError('Log-likelihood IDs need to be unique.')
            log_likelihood.set_id(_id)
            ids.append(_id)

    # -- evaluation ------------------------------------------------------------
    def _host_buffers(self, C, with_grad):
        """Page-locked result buffers, reused across calls."""
        cache = getattr(self, '_pinned', None)
        if cache is None or cache[0] < C or (with_grad and cache[3] is None):
            score = _lib.pinned_empty((C,))
            status = _lib.pinned_empty((C,), np.int32)
            grad = (cache[3] if cache and cache[0] >= C else None)
            if with_grad:
                # zero-filled: with a shard set only the shard's columns are
                # written, the other individuals' entries stay zero
                grad = _lib.pinned_empty((C, self._n_parameters))
                grad[...] = 0.0
            cache = (C, score, status, grad)
            self._pinned = cache
            # ctypes pointers of the buffers (creating them costs more than a
            # microsecond each: the single-vector path counts those)
            self._pinned_ptrs = (
                score.ctypes.data_as(_lib.c_f64p),
                status.ctypes.data_as(_lib.c_i32p),
                grad.ctypes.data_as(_lib.c_f64p) if grad is not None
                else None)
        _, score, status, grad = cache
        return score[:C], status[:C], (grad[:C] if grad is not None else None)

    def _evaluate(self, X, with_grad):
        X = np.ascontiguousarray(X, dtype=float).reshape(
            -1, self._n_parameters)
        C = X.shape[0]
        lib = _lib.lib()
        a_x, p_x = _lib.f64(X)
        score, status, grad = self._host_buffers(C, with_grad)
        p_s, p_st, p_g = self._pinned_ptrs
        _lib.check(lib.chi_b200_hier_logpdf(
            self._device.problem.handle, C, p_x, 1 if with_grad else 0, p_s,
            p_g if with_grad else None, p_st))
        if status.any():
            _failure_warning(status[status != 0][0])
            score[status != 0] = -np.inf
        return score, grad, status

    def _begin(self, X, with_grad):
        """Queues the evaluation of the rows of ``X`` and returns without
        waiting (``chi_b200_hier_logpdf_begin``); :meth:`_end` waits and
        returns what :meth:`_evaluate` returns.  The caller's host work in
        between -- the log-prior -- overlaps the GPU."""
        X = np.ascontiguousarray(X, dtype=float).reshape(
            -1, self._n_parameters)
        C = X.shape[0]
        a_x, p_x = _lib.f64(X)
        score, status, grad = self._host_buffers(C, with_grad)
        p_s, p_st, p_g = self._pinned_ptrs
        _lib.check(_lib.lib().chi_b200_hier_logpdf_begin(
            self._device.problem.handle, C, p_x, 1 if with_grad else 0, p_s,
            p_g if with_grad else None, p_st))
        return a_x, score, status, grad

    def _end(self, pending, warn=True):
        _, score, status, grad = pending
        _lib.check(_lib.lib().chi_b200_problem_synchronize(
            self._device.problem.handle))
        if status.any():
            if warn:
                _failure_warning(status[status != 0][0])
            score[status != 0] = -np.inf
        return score, grad, status

    def evaluate_batch(self, X, copy=True):
        """Scores of many parameter vectors, shape ``(n_vectors,)``.  With
        ``copy=False`` the returned array is the page-locked result buffer,
        valid until the next evaluation."""
        score = self._evaluate(X, False)[0]
        return score.copy() if copy else score

    def evaluateS1_batch(self, X, copy=True):
        """Scores and gradients of many parameter vectors."""
        score, grad, _ = self._evaluate(X, True)
        if copy:
            return score.copy(), grad.copy()
        return score, grad

    def compute_pointwise_ll(self, parameters, per_individual=True):
        """Not implemented by the reference either
        (chi/_log_pdfs.py:161-198 raises ``NotImplementedError``)."""
        raise NotImplementedError

    def __call__(self, parameters):
        """chi/_log_pdfs.py:104-145."""
        return float(self._evaluate(np.asarray(parameters), False)[0][0])

    def evaluateS1(self, parameters):
        """chi/_log_pdfs.py:255-300."""
        score, grad, _ = self._evaluate(np.asarray(parameters), True)
        return float(score[0]), grad[0].copy()

    def evaluate_device(self, d_x, n_vectors, with_grad, d_score, d_grad,
                        d_status, stream=0):
        """Evaluates ``n_vectors`` parameter vectors that are already
        resident on the device.  All arguments are raw device pointers
        (ints), ``stream`` a ``cudaStream_t``; nothing is copied and the call
        does not synchronise."""
        _lib.check(_lib.lib().chi_b200_hier_logpdf_dev(
            self._device.problem.handle, int(n_vectors), int(d_x),
            1 if with_grad else 0, int(d_score),
            int(d_grad) if with_grad else None, int(d_status),
            int(stream) if stream else None))

    def set_shard(self, id_begin, id_end):
        """Restricts evaluation to individuals ``[id_begin, id_end)``.
        Without a communicator the returned score and top-level gradient are
        partial sums; with one (:meth:`set_communicator`) they are the sums
        over all ranks, reduced on the device."""
        _lib.check(_lib.lib().chi_b200_problem_set_shard(
            self._device.problem.handle, int(id_begin), int(id_end)))
        self._shard = (int(id_begin), int(id_end))
        self._pinned = None   # result buffers are re-created zero-filled

    def shard(self):
        """Individuals ``[begin, end)`` this object evaluates."""
        return getattr(self, '_shard', (0, self._n_ids))

    @staticmethod
    def new_communicator_id():
        """Opaque bytes identifying a new communicator; rank 0 creates them
        and hands them to the other ranks."""
        import ctypes
        buf = (ctypes.c_uint8 * _lib.COMM_ID_BYTES)()
        _lib.check(_lib.lib().chi_b200_comm_unique_id(
            buf, _lib.COMM_ID_BYTES))
        return bytes(buf)

    def set_communicator(self, world, rank, comm_id):
        """Joins the all-reduce group of an evaluation whose individuals are
        sharded over ``world`` GPUs (collective: every rank calls it with the
        same ``comm_id``).  From then on every evaluation ends with one
        all-reduce of ``[score | d theta]`` on the device, inside the library
        (``chi_b200_problem_comm_init``)."""
        import ctypes
        comm_id = bytes(comm_id)
        if len(comm_id) != _lib.COMM_ID_BYTES:
            raise ValueError('Invalid communicator id.')
        buf = (ctypes.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(comm_id)
        _lib.check(_lib.lib().chi_b200_problem_comm_init(
            self._device.problem.handle, int(world), int(rank), buf))
        self._comm = (int(world), int(rank))

    def has_communicator(self):
        return getattr(self, '_comm', None) is not None

    def _fill_rows(self, grad, rows, value):
        """Writes ``value`` into the given rows of a gradient array -- into the
        columns this object owns only."""
        fill_owned_columns(grad, rows, value, self.shard(), self._n_ids,
                           self._n_bottom)

    def set_snapshot_budget(self, n_bytes):
        """See ``chi_b200_problem_set_snapshot_budget`` (0: observation
        fused in the solve kernel, -1: default)."""
        self._device.set_snapshot_budget(n_bytes)

    def set_batching(self, n_chunks=-1, order_min_vectors=-1, refill_wait=-1):
        """See ``chi_b200_problem_set_batching``: how a batch of parameter
        vectors is worked through (chunks of the host entry point, work order
        and refill policy of the solve kernel); -1 keeps the default.  The
        results do not depend on it."""
        _lib.check(_lib.lib().chi_b200_problem_set_batching(
            self._device.problem.handle, int(n_chunks),
            int(order_min_vectors), int(refill_wait)))

    def set_small_batch(self, max_systems=-1):
        """See ``chi_b200_problem_set_small_batch``: launches of at most
        ``max_systems`` systems run one warp per system (-1: default, 0:
        never)."""
        _lib.check(_lib.lib().chi_b200_problem_set_small_batch(
            self._device.problem.handle, int(max_systems)))

    def set_timing(self, enabled):
        _lib.check(_lib.lib().chi_b200_problem_set_timing(
            self._device.problem.handle, 1 if enabled else 0))

    def timings(self):
        """(sum of solve-kernel milliseconds, launches) since timing was
        enabled."""
        out = np.zeros(2)
        _lib.check(_lib.lib().chi_b200_problem_timings(
            self._device.problem.handle, out.ctypes.data_as(_lib.c_f64p)))
        return float(out[0]), int(out[1])

    def counters(self):
        """systems, accepted steps, rejected steps, kernel launches of the
        last evaluation."""
        out = np.zeros(4, dtype=np.int64)
        _lib.check(_lib.lib().chi_b200_problem_counters(
            self._device.problem.handle, out.ctypes.data_as(_lib.c_i64p)))
        return out

    # -- bookkeeping -------------------------------------------------------------
    def get_id(self, unique=False):
        """chi/_log_pdfs.py:302-325."""
        ids = list(self._ids)
        if unique:
            return ids
        n_copies = self._n_bottom // self._n_ids
        _ids = []
        for _id in ids:
            _ids += [_id] * n_copies
        ids = _ids
        ids += [None] * self._population_model.n_parameters()
        return ids

    def get_parameter_names(
            self, exclude_bottom_level=False, include_ids=False):
        """chi/_log_pdfs.py:327-367."""
        current_dim = 0
        names = list(self._ll_parameter_names)
        n = []
        special_dims, _, _ = self._population_model.get_special_dims()
        for info in special_dims:
            start_dim, end_dim, _, _, _ = info
            n += names[current_dim:start_dim]
            current_dim = end_dim
        n += names[current_dim:]
        names = n
        names = names * self._n_ids
        names += self._population_model.get_parameter_names()
        if include_ids:
            ids = self.get_id()
            n = []
            for idn, name in enumerate(names):
                if ids[idn]:
                    name = ids[idn] + ' ' + name
                n.append(name)
            names = n
        if exclude_bottom_level:
            names = names[self._n_bottom:]
        return names

    def get_population_model(self):
        return self._population_model

    def n_log_likelihoods(self):
        return self._n_ids

    def n_parameters(self, exclude_bottom_level=False):
        if exclude_bottom_level:
            return self._population_model.n_parameters()
        return self._n_parameters

    def n_observations(self):
        return self._n_obs


class HierarchicalLogPosterior(
        _chi_base('HierarchicalLogPosterior', _LogPDFBase)):
    """chi/_log_pdfs.py:406-634."""

    def __init__(self, log_likelihood, log_prior):
        if not isinstance(log_likelihood, HierarchicalLogLikelihood):
            raise TypeError(
                'The log-likelihood has to be an instance of a '
                'chi.HierarchicalLogLikelihood.')
        _check_prior(log_prior)
        n_top = log_likelihood.n_parameters(exclude_bottom_level=True)
        if log_prior.n_parameters() != n_top:
            raise ValueError(
                'The log-prior has to have as many parameters as population '
                'parameters in the log-likelihood. There are '
                '<' + str(n_top) + '> population parameters.')
        self._log_prior = log_prior
        self._log_likelihood = log_likelihood
        self._n_parameters = log_likelihood.n_parameters()
        self._n_bottom = self._n_parameters - n_top

    def __call__(self, parameters):
        """chi/_log_pdfs.py:463-472."""
        parameters = np.asarray(parameters)
        begin = getattr(self._log_likelihood, '_begin', None)
        if begin is None:
            score = self._log_prior(parameters[self._n_bottom:])
            if np.isinf(score):
                return score
            return score + self._log_likelihood(parameters)
        # likelihood queued first, prior on the host meanwhile (evaluateS1)
        pending = begin(parameters, False)
        try:
            score = self._log_prior(parameters[self._n_bottom:])
        finally:
            ll_s, _, _ = self._log_likelihood._end(pending, warn=False)
        if np.isinf(score):
            return score
        if not np.isfinite(ll_s[0]) and pending[2].any():
            _failure_warning(pending[2][pending[2] != 0][0])
        return score + float(ll_s[0])

    def evaluateS1(self, parameters):
        """chi/_log_pdfs.py:474-498."""
        parameters = np.asarray(parameters)
        begin = getattr(self._log_likelihood, '_begin', None)
        if begin is None:
            score, sens = self._log_prior.evaluateS1(
                parameters[self._n_bottom:])
            if np.isinf(score):
                return score, np.full(
                    shape=len(parameters), fill_value=np.inf)
            ll_score, sensitivities = self._log_likelihood.evaluateS1(
                parameters)
        else:
            # the likelihood is queued on the GPU first, the (host) prior is
            # evaluated while it runs; a vector outside the prior's support
            # returns as in the reference (which does not evaluate the
            # likelihood then)
            pending = begin(parameters, True)
            try:
                score, sens = self._log_prior.evaluateS1(
                    parameters[self._n_bottom:])
            finally:
                ll_s, ll_g, _ = self._log_likelihood._end(
                    pending, warn=False)
            if np.isinf(score):
                return score, np.full(
                    shape=len(parameters), fill_value=np.inf)
            if not np.isfinite(ll_s[0]) and pending[2].any():
                _failure_warning(pending[2][pending[2] != 0][0])
            ll_score, sensitivities = float(ll_s[0]), ll_g[0].copy()
        score += ll_score
        sensitivities[self._n_bottom:] += sens
        return score, sensitivities

    def _prior_batch(self, top, with_grad):
        """Prior (and its gradient) of every row of ``top`` on the host."""
        if hasattr(self._log_prior, 'evaluateS1_batch'):
            p, s = self._log_prior.evaluateS1_batch(top)
            return p, (s if with_grad else None)
        if not with_grad:
            return np.array([self._log_prior(x) for x in top]), None
        p = np.empty(len(top))
        s = np.empty(top.shape)
        for c, x in enumerate(top):
            p[c], s[c] = self._log_prior.evaluateS1(x)
        return p, s

    def _prior_async(self, X, with_grad):
        """Starts the host prior on a worker thread so that it runs while the
        calling thread waits for the GPU inside the (GIL-free) library call;
        returns a future."""
        pool = self.__dict__.get('_pool')
        if pool is None:
            # one worker per posterior object (not shared between objects:
            # two posteriors evaluated from two threads do not queue behind
            # each other)
            import concurrent.futures
            pool = concurrent.futures.ThreadPoolExecutor(
                1, thread_name_prefix='chi_b200-prior')
            self._pool = pool
        return pool.submit(self._prior_batch, X[:, self._n_bottom:], with_grad)

    def evaluate_batch(self, X, copy=True):
        """Batched ``__call__``: the prior is evaluated per vector on the
        host (top-level slice only, concurrently with the GPU pass), all
        likelihoods in one GPU pass."""
        X = np.asarray(X, dtype=float).reshape(-1, self._n_parameters)
        future = self._prior_async(X, False)
        try:
            score = self._log_likelihood.evaluate_batch(X, copy=False)
        finally:
            prior = future.result()[0]
        score = prior + score
        score[np.isinf(prior)] = prior[np.isinf(prior)]
        return score

    def evaluateS1_batch(self, X, copy=True):
        """Batched ``evaluateS1``.  With ``copy=False`` the results are views
        of the page-locked result buffers (valid until the next call)."""
        X = np.asarray(X, dtype=float).reshape(-1, self._n_parameters)
        future = self._prior_async(X, True)
        try:
            score, grad = self._log_likelihood.evaluateS1_batch(
                X, copy=False)
        finally:
            p, s = future.result()
        bad = np.isinf(p)
        score += p
        if np.any(bad):
            s = np.where(bad[:, None], 0.0, s)
        grad[:, self._n_bottom:] += s
        if np.any(bad):
            score[bad] = p[bad]
            fill = getattr(self._log_likelihood, '_fill_rows', None)
            if fill is not None:
                # a sharded likelihood owns only part of the columns
                fill(grad, bad, np.inf)
            else:
                grad[bad] = np.inf
        if copy:
            return score.copy(), grad.copy()
        return score, grad

    def get_log_likelihood(self):
        return self._log_likelihood

    def get_log_prior(self):
        return self._log_prior

    def get_id(self, unique=False):
        return self._log_likelihood.get_id(unique)

    def get_parameter_names(
            self, exclude_bottom_level=False, include_ids=False):
        return self._log_likelihood.get_parameter_names(
            exclude_bottom_level, include_ids)

    def get_population_model(self):
        return self._log_likelihood.get_population_model()

    def n_ids(self):
        return self._log_likelihood.n_log_likelihoods()

    def n_parameters(self, exclude_bottom_level=False):
        return self._log_likelihood.n_parameters(exclude_bottom_level)

    def sample_initial_parameters(self, n_samples=1, seed=None):
        """chi/_log_pdfs.py:552-634."""
        n_samples = int(n_samples)
        if n_samples <= 0:
            raise ValueError(
                'The number of samples has to be greater or equal to 1.')
        initial_params = np.empty(shape=(n_samples, self._n_parameters))
        np.random.seed(seed)
        n_top = self._log_prior.n_parameters()
        n_bottom = self._n_parameters - n_top
        initial_params[:, n_bottom:] = self._log_prior.sample(n_samples)
        if n_bottom == 0:
            return initial_params
        if seed is not None:
            seed += 1
        rng = np.random.default_rng(seed=seed)
        n_ids = self._log_likelihood.n_log_likelihoods()
        covariates = self._log_likelihood._covariates
        population_model = self._log_likelihood.get_population_model()
        bottom_parameters = []
        for sample_id in range(n_samples):
            kwargs = {}
            if population_model.n_covariates() > 0:
                kwargs['covariates'] = covariates
            bottom_parameters.append(population_model.sample(
                parameters=initial_params[sample_id, n_bottom:],
                n_samples=n_ids, seed=rng, **kwargs))
        dims = population_model._hier_columns()
        for idx, bottom_params in enumerate(bottom_parameters):
            bottom_parameters[idx] = bottom_params[:, dims].flatten()
        initial_params[:, :n_bottom] = np.vstack(bottom_parameters)
        return initial_params


def _check_prior(log_prior):
    ok = all(hasattr(log_prior, name) for name in (
        '__call__', 'evaluateS1', 'n_parameters', 'sample'))
    if _LogPriorBase is not None and isinstance(log_prior, _LogPriorBase):
        ok = True
    if not ok:
        raise TypeError(
            'The log-prior has to be an instance of a pints.LogPrior.')


class LogPosterior(_chi_base('LogPosterior', _LogPDFBase)):
    """chi/_log_pdfs.py:1152-1322."""

    def __init__(self, log_likelihood, log_prior):
        if not isinstance(log_likelihood, LogLikelihood):
            raise TypeError(
                'The log-likelihood has to extend chi.LogLikelihood.')
        _check_prior(log_prior)
        if log_prior.n_parameters() != log_likelihood.n_parameters():
            raise ValueError(
                'The log-prior and the log-likelihood must have the same '
                'dimension.')
        self._log_prior = log_prior
        self._log_likelihood = log_likelihood
        self._n_parameters = log_likelihood.n_parameters()

    def __call__(self, parameters):
        """chi/_log_pdfs.py:1245-1251."""
        score = self._log_prior(parameters)
        if np.isinf(score):
            return score
        return score + self._log_likelihood(parameters)

    def evaluateS1(self, parameters):
        """chi/_log_pdfs.py:1253-1273."""
        score, sens = self._log_prior.evaluateS1(parameters)
        if np.isinf(score):
            return score, np.full(shape=len(parameters), fill_value=np.inf)
        ll_score, sensitivities = self._log_likelihood.evaluateS1(parameters)
        score += ll_score
        sensitivities += sens
        return score, sensitivities

    def get_log_likelihood(self):
        return self._log_likelihood

    def get_log_prior(self):
        return self._log_prior

    def get_id(self, *args, **kwargs):
        return self._log_likelihood.get_id()

    def get_parameter_names(self, *args, **kwargs):
        return self._log_likelihood.get_parameter_names()

    def n_parameters(self, *args, **kwargs):
        return self._n_parameters

    def sample_initial_parameters(self, n_samples=1, seed=None):
        """chi/_log_pdfs.py:1302-1322."""
        n_samples = int(n_samples)
        if n_samples <= 0:
            raise ValueError(
                'The number of samples has to be greater or equal to 1.')
        np.random.seed(seed)
        return self._log_prior.sample(n_samples)
