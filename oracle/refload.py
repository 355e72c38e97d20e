"""Load the UNMODIFIED reference statistical modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  This loader exists so that golden vectors can be
generated from (and the numpy restatement in ``oracle/stats.py`` validated
against) chi's own code in the build container.  ``/root/reference`` does not
exist on the GPU box, so nothing that runs there may import this module; the
only consumers are ``tests/golden/make_golden.py`` and the (skipped when the
reference is absent) cross-checks in ``tests/test_oracle_vs_reference.py``.

chi cannot be imported normally here because ``myokit``, ``pints``, ``arviz``,
``xarray`` and ``plotly`` are not installed.  The five
modules below only need numpy/scipy plus the *names* ``myokit.SimulationError``,
``myokit.formats.sbml``, ``pints.LogPDF``, ``pints.LogPrior`` and
``pints.vector``; tiny stand-ins for those are registered before loading.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('CHI_REFERENCE_ROOT', '/root/reference')

_MODULES = [
    '_covariate_models', '_error_models', '_population_models',
    '_population_filters', '_mechanistic_models', '_log_pdfs',
    '_predictive_models', '_problems']


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'chi', '_log_pdfs.py'))


def _stub_pints():
    pints = types.ModuleType('pints')

    class LogPDF(object):
        def __call__(self, x):
            raise NotImplementedError

        def evaluateS1(self, x):
            raise NotImplementedError

        def n_parameters(self):
            raise NotImplementedError

    class LogPrior(LogPDF):
        pass

    def vector(x):
        x = np.array(x, dtype=float, copy=True)
        if x.ndim != 1:
            x = x.reshape(-1)
        x.setflags(write=False)
        return x

    pints.LogPDF = LogPDF
    pints.LogPrior = LogPrior
    pints.vector = vector
    return pints


def _stub_myokit():
    myokit = types.ModuleType('myokit')

    class SimulationError(Exception):
        pass

    class ProtocolEvent(object):
        """(level, start, duration): what chi/_problems.py:360-372 builds."""
        def __init__(self, level, start, duration, period=0, multiplier=0):
            self.fields = (float(level), float(start), float(duration),
                           float(period), float(multiplier))

    class Protocol(object):
        def __init__(self):
            self.events = []

        def add(self, event):
            self.events.append(event.fields)

    myokit.SimulationError = SimulationError
    myokit.Protocol = Protocol
    myokit.ProtocolEvent = ProtocolEvent
    formats = types.ModuleType('myokit.formats')
    sbml = types.ModuleType('myokit.formats.sbml')
    formats.sbml = sbml
    myokit.formats = formats
    return myokit, formats, sbml


def load():
    """Returns a stub ``chi`` namespace holding the reference's classes."""
    if 'chi' in sys.modules and getattr(
            sys.modules['chi'], '_oracle_refload', False):
        return sys.modules['chi']
    if not available():
        raise RuntimeError('reference not present at ' + REFERENCE_ROOT)
    if 'chi' in sys.modules:
        raise RuntimeError('a real `chi` is already imported')

    if 'pints' not in sys.modules:
        sys.modules['pints'] = _stub_pints()
    if 'xarray' not in sys.modules:
        # only referenced by the posterior/prior predictive classes at call
        # time; a bare module is enough to import chi/_predictive_models.py
        sys.modules['xarray'] = types.ModuleType('xarray')
    if 'myokit' not in sys.modules:
        myokit, formats, sbml = _stub_myokit()
        sys.modules['myokit'] = myokit
        sys.modules['myokit.formats'] = formats
        sys.modules['myokit.formats.sbml'] = sbml

    chi = types.ModuleType('chi')
    chi._oracle_refload = True
    chi.__path__ = []
    sys.modules['chi'] = chi
    for name in _MODULES:
        path = os.path.join(REFERENCE_ROOT, 'chi', name + '.py')
        spec = importlib.util.spec_from_file_location('chi.' + name, path)
        module = importlib.util.module_from_spec(spec)
        sys.modules['chi.' + name] = module
        spec.loader.exec_module(module)
        for key, value in vars(module).items():
            if isinstance(value, type) and value.__module__ == 'chi.' + name:
                setattr(chi, key, value)
    return chi
