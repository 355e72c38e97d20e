"""Mechanistic models backed by the CUDA integrator.

Host-side mirror of ``chi/_mechanistic_models.py``: same class names, method
names, argument meaning and error behaviour, so that these classes satisfy
the seams chi's log-pdfs and predictive models call
(``simulate(parameters, times)``, ``enable_sensitivities``, ``copy`` ...;
SURVEY.md section 8b, seam B1).  What differs is the engine: instead of
``myokit.Simulation`` (one CVODES module compiled per instance) a model is a
key into the registry of generated sm_100a kernels, and ``simulate`` is one
launch of the batched integrator through the C ABI.
"""
import copy

import numpy as np

from chi_b200 import _codegen, _lib, _sbml

try:  # drop-in: satisfy chi's isinstance checks when chi is importable
    import chi as _chi
    _MechanisticModelBase = _chi.MechanisticModel
except Exception:  # pragma: no cover - depends on the environment
    _chi = None
    _MechanisticModelBase = object


# Default integrator tolerances.  chi never sets tolerances, i.e. it runs at
# myokit's defaults (abs 1e-6, rel 1e-4), at which trajectories are only
# ~1e-4 accurate; the defaults here deliver the 1e-6 parity contract on
# trajectories and sensitivities ("Tolerances").
DEFAULT_RTOL = 1e-8
DEFAULT_ATOL = 1e-10
DEFAULT_MAX_STEPS = 1000000


class MechanisticModel(_MechanisticModelBase):
    """Base class, see chi/_mechanistic_models.py:15-119."""

    def __init__(self):
        super(MechanisticModel, self).__init__()

    def copy(self):
        return copy.deepcopy(self)

    def enable_sensitivities(self, enabled, parameter_names=None):
        raise NotImplementedError

    def has_sensitivities(self):
        raise NotImplementedError

    def n_outputs(self):
        raise NotImplementedError

    def n_parameters(self):
        raise NotImplementedError

    def outputs(self):
        raise NotImplementedError

    def parameters(self):
        raise NotImplementedError

    def simulate(self, parameters, times):
        raise NotImplementedError

    def supports_dosing(self):
        return False


class ProtocolEvent(object):
    """Stand-in for ``myokit.ProtocolEvent`` (fields used by chi:
    chi/tests/test_mechanistic_models.py:498-520)."""

    def __init__(self, level, start, duration, period=0, multiplier=0):
        self._level = float(level)
        self._start = float(start)
        self._duration = float(duration)
        self._period = float(period)
        self._multiplier = int(multiplier)

    def level(self):
        return self._level

    def start(self):
        return self._start

    def duration(self):
        return self._duration

    def period(self):
        return self._period

    def multiplier(self):
        return self._multiplier

    def as_tuple(self):
        return (self._level, self._start, self._duration, self._period,
                float(self._multiplier))


class DosingRegimen(object):
    """Stand-in for ``myokit.Protocol``: a list of :class:`ProtocolEvent`."""

    def __init__(self, events=None):
        self._events = list(events or [])

    def schedule(self, level, start, duration, period=0, multiplier=0):
        self._events.append(
            ProtocolEvent(level, start, duration, period, multiplier))

    def events(self):
        return list(self._events)

    def as_array(self):
        if not self._events:
            return np.zeros((0, 5))
        return np.array([e.as_tuple() for e in self._events], dtype=float)

    def clone(self):
        return DosingRegimen(self._events)


def blocktrain(period, duration, offset=0, level=1.0, limit=0):
    """``myokit.pacing.blocktrain`` (call site
    chi/_mechanistic_models.py:853-855)."""
    return DosingRegimen(
        [ProtocolEvent(level, offset, duration, period, limit)])


def _events_of(regimen):
    """Accepts a DosingRegimen, a myokit.Protocol-like object or None."""
    if regimen is None:
        return np.zeros((0, 5))
    if isinstance(regimen, DosingRegimen):
        return regimen.as_array()
    events = [(e.level(), e.start(), e.duration(), e.period(),
               float(e.multiplier())) for e in regimen.events()]
    return np.array(events, dtype=float).reshape(-1, 5)


class SBMLModel(MechanisticModel):
    """Mechanistic model from an SBML file, see
    chi/_mechanistic_models.py:122-539.

    :param sbml_file: path to an SBML file, or a
        :class:`chi_b200._sbml.SymbolicModel`.
    """

    def __init__(self, sbml_file):
        super(SBMLModel, self).__init__()
        if isinstance(sbml_file, _sbml.SymbolicModel):
            self._model = sbml_file.clone()
        else:
            self._model = _sbml.read_sbml(sbml_file)
        self._time_unit = self._model.time_unit
        self._has_sensitivities = False
        self._sens_columns = None
        self._rtol = DEFAULT_RTOL
        self._atol = DEFAULT_ATOL
        self._max_steps = DEFAULT_MAX_STEPS
        self._device = None
        self._reset_engine()
        self._set_number_and_names()

    # -- engine -------------------------------------------------------------
    def _reset_engine(self):
        self._code = _codegen.ModelCode(self._model)
        self._problem = None
        self._uploaded = None

    def _set_number_and_names(self):
        """chi/_mechanistic_models.py:177-217."""
        self._n_states = self._code.n
        self._state_names = list(self._code.state_names)
        self._const_names = list(self._code.const_names)
        self._n_parameters = self._n_states + len(self._const_names)
        self._parameter_names = self._state_names + self._const_names
        self._output_names = list(self._state_names)
        self._n_outputs = self._n_states
        self._parameter_name_map = dict(
            zip(self._parameter_names, self._parameter_names))
        self._output_name_map = dict(
            zip(self._output_names, self._output_names))

    def model_key(self):
        """Key of the generated kernels of this model."""
        return self._code.key

    def model_code(self):
        return self._code

    def set_tolerance(self, abs_tol=DEFAULT_ATOL, rel_tol=DEFAULT_RTOL,
                      max_steps=None):
        """Counterpart of ``model._simulator.set_tolerance`` (the simulator
        attribute chi leaves public, chi/_mechanistic_models.py:144-146)."""
        self._atol = float(abs_tol)
        self._rtol = float(rel_tol)
        if max_steps is not None:
            self._max_steps = int(max_steps)
        self._uploaded = None

    def tolerance(self):
        return self._atol, self._rtol, self._max_steps

    def set_kernel_variant(self, variant):
        """Selects the sensitivity-kernel variant (index into
        ``model_code().variants()``; ``None`` = default).  For testing and
        tuning."""
        self._variant = None if variant is None else int(variant)
        self._uploaded = None

    def kernel_variant(self):
        return getattr(self, '_variant', None)

    def set_device(self, device):
        self._device = int(device)
        self._problem = None
        self._uploaded = None

    def _regimen_events(self):
        return np.zeros((0, 5))

    def _get_problem(self):
        if self._problem is None:
            model_id = _lib.find_model(self._code)
            self._problem = _lib.Problem(model_id, self._device)
            self._uploaded = None
        return self._problem

    def _configure(self, times):
        """Uploads the time grid (one record per time and output)."""
        problem = self._get_problem()
        lib = _lib.lib()
        out_ids = [self._code.output_names.index(o)
                   for o in self._output_names]
        events = self._regimen_events()
        cols = self._sens_columns if self._has_sensitivities else None
        state = (times.tobytes(), tuple(out_ids), events.tobytes(),
                 None if cols is None else tuple(cols),
                 self._rtol, self._atol, self._max_steps,
                 self.kernel_variant())
        if self._uploaded == state:
            return problem
        n_out = len(out_ids)
        _, p_out = _lib.i32(out_ids)
        _lib.check(lib.chi_b200_problem_set_outputs(
            problem.handle, n_out, p_out, None))
        if cols is None:
            cols = list(range(self._n_parameters))
        _, p_cols = _lib.i32(cols)
        _lib.check(lib.chi_b200_problem_set_sensitivity_columns(
            problem.handle, len(cols), p_cols))
        _lib.check(lib.chi_b200_problem_set_tolerances(
            problem.handle, self._rtol, self._atol, self._max_steps))
        _lib.check(lib.chi_b200_problem_set_variant(
            problem.handle,
            -1 if self.kernel_variant() is None else self.kernel_variant()))
        n_t = len(times)
        rec_t = np.repeat(times, n_out)
        rec_out = np.tile(np.arange(n_out, dtype=np.int32), n_t)
        rec_off = np.array([0, n_t * n_out], dtype=np.int64)
        dose_off = np.array([0, len(events)], dtype=np.int64)
        a_t, p_t = _lib.f64(rec_t)
        a_o, p_o = _lib.i32(rec_out)
        a_ro, p_ro = _lib.i64(rec_off)
        a_do, p_do = _lib.i64(dose_off)
        a_ev, p_ev = _lib.f64(events.reshape(-1))
        _lib.check(lib.chi_b200_problem_upload(
            problem.handle, 1, p_ro, p_t, None, p_o, p_do, p_ev))
        self._uploaded = state
        return problem

    # -- reference API ------------------------------------------------------
    def copy(self):
        """chi/_mechanistic_models.py:219-242: copying resets the
        sensitivity settings."""
        problem, uploaded = self._problem, self._uploaded
        self._problem, self._uploaded = None, None
        model = copy.deepcopy(self)
        self._problem, self._uploaded = problem, uploaded
        model._has_sensitivities = False
        model._sens_columns = None
        return model

    def enable_sensitivities(self, enabled, parameter_names=None):
        """chi/_mechanistic_models.py:244-312."""
        enabled = bool(enabled)
        if not enabled:
            self._has_sensitivities = False
            self._sens_columns = None
            return None
        columns = list(range(self._n_parameters))
        if parameter_names is not None:
            columns = [
                index for index, public_name in enumerate(
                    self._parameter_name_map.values())
                if public_name in parameter_names]
        if not columns:
            raise ValueError(
                'None of the parameters could be identified. The valid '
                'parameter names are <' + str(self._parameter_names) + '>.')
        self._sens_columns = columns
        self._has_sensitivities = True

    def has_sensitivities(self):
        return self._has_sensitivities

    def n_outputs(self):
        return self._n_outputs

    def n_parameters(self):
        return self._n_parameters

    def outputs(self):
        return [self._output_name_map[name] for name in self._output_names]

    def parameters(self):
        return [
            self._parameter_name_map[name] for name in self._parameter_names]

    def set_outputs(self, outputs):
        """chi/_mechanistic_models.py:356-407."""
        outputs = list(outputs)
        for internal, public_name in self._output_name_map.items():
            if public_name in outputs:
                outputs[outputs.index(public_name)] = internal
        for output in outputs:
            if not self._model.has_variable(output):
                raise KeyError(
                    'The variable <' + str(output) + '> does not exist in the '
                    'model.')
            if output not in self._code.output_names:
                raise ValueError(
                    'Outputs have to be state or intermediary variables.')
        self._output_names = outputs
        self._n_outputs = len(outputs)
        output_name_map = {}
        for name in self._output_names:
            output_name_map[name] = self._output_name_map.get(name, name)
        self._output_name_map = output_name_map
        self.enable_sensitivities(False)

    def set_output_names(self, names):
        """chi/_mechanistic_models.py:409-446."""
        if not isinstance(names, dict):
            raise TypeError(
                'Names has to be a dictionary with the current output names'
                'as keys and the new output names as values.')
        new_names = list(names.values())
        if len(new_names) != len(set(new_names)):
            raise ValueError('The new output names have to be unique.')
        for new_name in new_names:
            if new_name in list(self._output_name_map.values()):
                raise ValueError(
                    'The output names cannot coincide with existing '
                    'output names. One output is already called '
                    '<' + str(new_name) + '>.')
        for internal in self._output_names:
            old_name = self._output_name_map[internal]
            if old_name in names:
                self._output_name_map[internal] = str(names[old_name])

    def set_parameter_names(self, names):
        """chi/_mechanistic_models.py:448-488."""
        if not isinstance(names, dict):
            raise TypeError(
                'Names has to be a dictionary with the current parameter names'
                'as keys and the new parameter names as values.')
        new_names = list(names.values())
        if len(new_names) != len(set(new_names)):
            raise ValueError('The new parameter names have to be unique.')
        for new_name in new_names:
            if new_name in list(self._parameter_name_map.values()):
                raise ValueError(
                    'The parameter names cannot coincide with existing '
                    'parameter names. One parameter is already called '
                    '<' + str(new_name) + '>.')
        for internal in self._parameter_names:
            old_name = self._parameter_name_map[internal]
            if old_name in names:
                self._parameter_name_map[internal] = str(names[old_name])

    def simulate(self, parameters, times):
        """chi/_mechanistic_models.py:490-533: returns outputs of shape
        ``(n_outputs, n_times)`` and, with sensitivities enabled, also
        ``(n_times, n_outputs, n_parameters)``."""
        parameters = np.asarray(parameters, dtype=float).reshape(-1)
        if len(parameters) != self._n_parameters:
            raise ValueError(
                'The number of parameters does not match the model.')
        times = np.ascontiguousarray(times, dtype=float).reshape(-1)
        if len(times) == 0:
            raise IndexError('times must not be empty.')
        n_t, n_out = len(times), self._n_outputs
        problem = self._configure(times)
        lib = _lib.lib()
        with_sens = self._has_sensitivities
        n_cols = len(self._sens_columns) if with_sens else 0
        a_psi, p_psi = _lib.f64(parameters)
        y, p_y = _lib.out_f64(n_t * n_out)
        s, p_s = _lib.out_f64(max(n_t * n_out * n_cols, 1))
        status, p_st = _lib.out_i32(1)
        _lib.check(lib.chi_b200_simulate(
            problem.handle, 1, None, p_psi, 1 if with_sens else 0, p_y,
            p_s if with_sens else None, p_st))
        if status[0] != 0:
            raise SimulationError(
                'The integrator failed with status %d.' % int(status[0]))
        output = y.reshape(n_t, n_out).T.copy()
        if not with_sens:
            return output
        sens = s[:n_t * n_out * n_cols].reshape(n_t, n_out, n_cols).copy()
        return output, sens

    def simulate_batch(self, parameters, times, sensitivities=False):
        """Simulates many parameter vectors on the same time grid in one
        launch (not part of chi's API; used by the predictive models and the
        population-filter posterior).

        :param parameters: array of shape ``(n_vectors, n_parameters)``.
        :param sensitivities: also return d output / d parameter (of the
            enabled sensitivity columns, all parameters if none were chosen).
        :returns: outputs of shape ``(n_vectors, n_outputs, n_times)``, the
            per-vector integrator status and, with ``sensitivities``, an
            array of shape ``(n_vectors, n_times, n_outputs, n_columns)``.
        """
        parameters = np.ascontiguousarray(
            parameters, dtype=float).reshape(-1, self._n_parameters)
        times = np.ascontiguousarray(times, dtype=float).reshape(-1)
        if len(times) == 0:
            raise IndexError('times must not be empty.')
        n_t, n_out, n_b = len(times), self._n_outputs, len(parameters)
        if sensitivities and not self._has_sensitivities:
            self.enable_sensitivities(True)
        problem = self._configure(times)
        lib = _lib.lib()
        a_psi, p_psi = _lib.f64(parameters)
        y, p_y = _lib.out_f64(n_b * n_t * n_out)
        status, p_st = _lib.out_i32(n_b)
        if not sensitivities:
            _lib.check(lib.chi_b200_simulate(
                problem.handle, n_b, None, p_psi, 0, p_y, None, p_st))
            return y.reshape(n_b, n_t, n_out).transpose(0, 2, 1).copy(), \
                status
        n_cols = len(self._sens_columns)
        sens, p_s = _lib.out_f64(max(n_b * n_t * n_out * n_cols, 1))
        _lib.check(lib.chi_b200_simulate(
            problem.handle, n_b, None, p_psi, 1, p_y, p_s, p_st))
        return (y.reshape(n_b, n_t, n_out).transpose(0, 2, 1).copy(), status,
                sens[:n_b * n_t * n_out * n_cols].reshape(
                    n_b, n_t, n_out, n_cols))

    def time_unit(self):
        return self._time_unit


class SimulationError(Exception):
    """Counterpart of ``myokit.SimulationError``
    (caught at chi/_log_pdfs.py:816, 990)."""


class PKPDModel(SBMLModel):
    """PKPD model with dose administration, see
    chi/_mechanistic_models.py:542-864."""

    def __init__(self, sbml_file):
        super(PKPDModel, self).__init__(sbml_file)
        self._administration = None
        self._dosing_regimen = None
        self._vanilla_model = self._model.clone()

    def _regimen_events(self):
        return _events_of(self._dosing_regimen)

    def administration(self):
        return self._administration

    def copy(self):
        return super(PKPDModel, self).copy()

    def dosing_regimen(self):
        return self._dosing_regimen

    def set_administration(
            self, compartment, amount_var='drug_amount', direct=True):
        """chi/_mechanistic_models.py:705-793."""
        model = self._vanilla_model.clone()
        if compartment not in model.components():
            raise ValueError(
                'The model does not have a compartment named <'
                + str(compartment) + '>.')
        drug_amount = compartment + '.' + amount_var
        if not model.has_variable(drug_amount):
            raise ValueError(
                'The drug amount variable <' + str(amount_var) + '> could not '
                'be found in the compartment.')
        if drug_amount not in model.states:
            raise ValueError(
                'The variable <' + str(drug_amount) + '> is not a state '
                'variable, and can therefore not be dosed.')
        original_outputs = self._output_names
        if not direct:
            drug_amount = model.add_dose_compartment(drug_amount)
        model.add_dose_rate(drug_amount)
        self._model = model
        self._reset_engine()
        # chi/_mechanistic_models.py:604-607 re-derives the names and restores
        # the outputs on the indirect route only, which leaves stale names
        # when switching back to a direct route; here both routes re-derive.
        names = dict(self._parameter_name_map)
        self._set_number_and_names()
        for name in self._parameter_names:
            if name in names:
                self._parameter_name_map[name] = names[name]
        self.set_outputs(original_outputs)
        self._has_sensitivities = False
        self._sens_columns = None
        self._administration = dict(
            {'compartment': compartment, 'direct': direct})

    def set_dosing_regimen(
            self, dose, start=0, duration=0.01, period=None, num=None):
        """chi/_mechanistic_models.py:795-857."""
        if self._administration is None:
            raise ValueError(
                'The route of administration of the dose has not been set.')
        if num is None:
            num = 0
        if period is None:
            period = 0
            num = 0
        if isinstance(dose, DosingRegimen) or hasattr(dose, 'events'):
            self._dosing_regimen = dose
            self._uploaded = None
            return None
        dose_rate = dose / duration
        self._dosing_regimen = blocktrain(
            period=period, duration=duration, offset=start, level=dose_rate,
            limit=num)
        self._uploaded = None

    def supports_dosing(self):
        return True


class FixedParameters(object):
    """Which entries of a named parameter vector are held at fixed values.

    Shared by the Reduced* wrappers (chi fixes parameters with
    ``fix_parameters({name: value})``, ``None`` releasing a parameter again:
    chi/_mechanistic_models.py:949-1019, chi/_error_models.py:1753-1821,
    chi/_log_pdfs.py:1029-1073).  ``expand`` always returns a fresh array
    (the reference scatters into a shared scratch buffer and returns it).
    """

    def __init__(self, names):
        self.names = list(names)
        self.values = {}            # position -> value

    def update(self, name_value_dict):
        try:
            name_value_dict = dict(name_value_dict)
        except (TypeError, ValueError):
            raise ValueError(
                'The name-value dictionary has to be convertable to a python '
                'dictionary.')
        for position, name in enumerate(self.names):
            if name not in name_value_dict:
                continue
            value = name_value_dict[name]
            if value is None:
                self.values.pop(position, None)
            else:
                self.values[position] = float(value)

    def rename(self, names):
        self.names = list(names)

    def n_fixed(self):
        return len(self.values)

    def mask(self):
        mask = np.zeros(len(self.names), dtype=bool)
        mask[list(self.values)] = True
        return mask

    def free_names(self):
        return [n for k, n in enumerate(self.names) if k not in self.values]

    def free_positions(self):
        return [k for k in range(len(self.names)) if k not in self.values]

    def full_values(self):
        full = np.zeros(len(self.names))
        for position, value in self.values.items():
            full[position] = value
        return full

    def expand(self, free_values):
        free_values = np.asarray(free_values, dtype=float).reshape(-1)
        if not self.values:
            return free_values.copy()
        full = self.full_values()
        full[self.free_positions()] = free_values
        return full


class ReducedMechanisticModel(MechanisticModel):
    """A mechanistic model with some parameters held at fixed values
    (interface of chi/_mechanistic_models.py:867-1206).  Everything but the
    parameter bookkeeping is forwarded to the wrapped model; inside a
    log-likelihood the fixed entries go to the device once
    (``chi_b200_problem_set_fixed_parameters``) and sensitivities are
    integrated for the free parameters only."""

    _FORWARDED = (
        'has_sensitivities', 'n_outputs', 'outputs', 'set_output_names',
        'supports_dosing', 'time_unit')

    def __init__(self, mechanistic_model):
        super(ReducedMechanisticModel, self).__init__()
        if not isinstance(mechanistic_model, MechanisticModel):
            raise TypeError(
                'The mechanistic model has to be an instance of a '
                'chi.MechanisticModel.')
        self._mechanistic_model = mechanistic_model
        self._fixed = FixedParameters(mechanistic_model.parameters())

    def __getattr__(self, name):
        # only reached for attributes that are not defined on the wrapper
        if name in ReducedMechanisticModel._FORWARDED:
            return getattr(self.__dict__['_mechanistic_model'], name)
        raise AttributeError(name)

    # (defined explicitly: the base class declares them)
    def has_sensitivities(self):
        return self._mechanistic_model.has_sensitivities()

    def n_outputs(self):
        return self._mechanistic_model.n_outputs()

    def outputs(self):
        return self._mechanistic_model.outputs()

    def supports_dosing(self):
        return self._mechanistic_model.supports_dosing()

    def copy(self):
        model = ReducedMechanisticModel(self._mechanistic_model.copy())
        model._fixed.values = dict(self._fixed.values)
        return model

    def mechanistic_model(self):
        return self._mechanistic_model

    def fixed_parameters(self):
        """The :class:`FixedParameters` table (positions in the wrapped
        model's parameter vector)."""
        return self._fixed

    def fix_parameters(self, name_value_dict):
        self._fixed.update(name_value_dict)
        self._mechanistic_model.enable_sensitivities(False)

    def n_fixed_parameters(self):
        return self._fixed.n_fixed()

    def n_parameters(self):
        return len(self._fixed.names) - self._fixed.n_fixed()

    def parameters(self):
        return self._fixed.free_names()

    def set_parameter_names(self, names):
        self._mechanistic_model.set_parameter_names(names)
        self._fixed.rename(self._mechanistic_model.parameters())

    def set_outputs(self, outputs):
        self._mechanistic_model.set_outputs(outputs)

    def enable_sensitivities(self, enabled):
        """Sensitivities with respect to the free parameters only
        (chi/_mechanistic_models.py:931-947)."""
        if enabled:
            self._mechanistic_model.enable_sensitivities(
                True, self.parameters())
        else:
            self._mechanistic_model.enable_sensitivities(False)

    def dosing_regimen(self):
        getter = getattr(self._mechanistic_model, 'dosing_regimen', None)
        return getter() if getter is not None else None

    def set_dosing_regimen(
            self, dose, start=0, duration=0.01, period=None, num=None):
        setter = getattr(self._mechanistic_model, 'set_dosing_regimen', None)
        if setter is None:
            raise AttributeError(
                'The mechanistic model does not support dosing regimens.')
        setter(dose, start, duration, period, num)

    def simulate(self, parameters, times):
        """chi/_mechanistic_models.py:1160-1186."""
        return self._mechanistic_model.simulate(
            self._fixed.expand(parameters), times)
