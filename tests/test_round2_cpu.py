"""CPU-side checks of the round-2 additions against golden vectors of the
UNMODIFIED reference (tests/golden/round2_golden.json): population filters,
DataFrame ingestion, covariate-model bookkeeping, reduced-model bookkeeping
and the drop-in class hierarchy under an importable ``chi``."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import chi_b200
from chi_b200 import _problems

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _golden():
    with open(os.path.join(HERE, 'golden', 'round2_golden.json')) as f:
        return json.load(f)


def _nan(a):
    return np.array([[[np.nan if v is None else v for v in r] for r in o]
                     for o in a], dtype=float)


def test_population_filters_match_reference():
    """chi/_population_filters.py: scores and sensitivities of every filter,
    with missing observations, and the composed filter with re-ordered
    times."""
    g = _golden()['filters']
    obs = _nan(g['observations'])
    y = np.array(g['simulated'])
    for case in g['cases']:
        if case['filter'] == 'composed':
            f = chi_b200.ComposedPopulationFilter([
                chi_b200.GaussianFilter(obs[:, :, :2]),
                chi_b200.LogNormalKDEFilter(obs[:, :, 2:])])
            f.sort_times(np.array(case['order'], dtype=int))
        else:
            f = getattr(chi_b200, case['filter'])(obs, **case['kwargs'])
            assert f.compute_log_likelihood(y) == pytest.approx(
                case['score_call'], rel=1e-13)
        score, grad = f.compute_sensitivities(y)
        assert score == pytest.approx(case['score'], rel=1e-13)
        np.testing.assert_allclose(grad, case['grad'], rtol=1e-11, atol=1e-13)
        assert f.n_observables() == 2 and f.n_times() == 4
    # everything missing: -inf, as in the reference
    empty = chi_b200.GaussianFilter(np.full((3, 1, 2), np.nan))
    assert empty.compute_log_likelihood(np.ones((4, 1, 2))) == -np.inf
    with pytest.raises(ValueError):
        chi_b200.GaussianFilter(np.ones((3, 2)))
    with pytest.raises(ValueError):
        chi_b200.GaussianMixtureFilter(obs, n_kernels=5) \
            .compute_log_likelihood(y)


def test_dataframe_ingestion_matches_the_reference_controller():
    """What chi.ProblemModellingController builds per individual from a
    long-format table (chi/_problems.py:246-376): measurement arrays per
    output, one dose event per dose row (bolus = 0.01), covariates."""
    import pandas as pd
    g = _golden()['ingestion']
    rows = []
    for r, is_str in zip(g['rows'], g['row_id_is_str']):
        r = [np.nan if v is None else v for v in r]
        r[0] = r[0] if is_str else int(r[0])
        rows.append(r)
    data = pd.DataFrame(rows, columns=g['columns'])
    outputs = list(g['output_observable_dict'])
    ids, times, values = _problems.extract_measurements(
        data, outputs, g['output_observable_dict'])
    assert [str(i) for i in ids] == [c['id'] for c in g['individuals']]
    regimens = _problems.extract_dosing_regimens(data, ids)
    for k, case in enumerate(g['individuals']):
        for o in range(len(outputs)):
            np.testing.assert_array_equal(times[k][o], case['times'][o])
            np.testing.assert_array_equal(
                values[k][o], case['observations'][o])
        events = regimens[k].as_array()
        ref = np.array(case['events'])
        np.testing.assert_allclose(events[:, :3], ref[:, :3], rtol=1e-15)
        assert np.all(events[:, 3:] == 0)
    cov = _problems.extract_covariates(
        data, ids, ['Age', 'Weight'])
    np.testing.assert_array_equal(cov, g['covariates'])


def test_covariate_model_bookkeeping_matches_reference():
    for case in _golden()['linear_covariate']:
        cm = chi_b200.LinearCovariateModel(n_cov=case['n_cov'])
        if case['selection'] is not None:
            cm.set_population_parameters(case['selection'])
        pidx, didx = cm.get_set_population_parameters()
        assert list(pidx) == [int(v) for v in case['pidx']]
        assert list(didx) == [int(v) for v in case['didx']]
        assert cm.get_parameter_names() == case['names']
        assert cm.n_parameters() == len(case['beta'])
    cm = chi_b200.LinearCovariateModel(n_cov=2, cov_names=['Age', 'Sex'])
    cm.set_population_parameters([[0, 0], [1, 0]])
    cm.set_parameter_names(['a', 'a', 'b', 'b'])
    assert cm.get_parameter_names() == ['a Age', 'a Sex', 'b Age', 'b Sex']
    assert cm.get_parameter_names(exclude_cov_names=True) == [
        'a', 'a', 'b', 'b']
    cm.set_covariate_names(None)
    assert cm.get_covariate_names() == ['Cov. 1', 'Cov. 2']
    with pytest.raises(ValueError):
        cm.set_parameter_names(['x'])
    with pytest.raises(ValueError):
        chi_b200.LinearCovariateModel(n_cov=0)


def test_fixed_parameter_table():
    from chi_b200._mechanistic_models import FixedParameters
    t = FixedParameters(['a', 'b', 'c', 'd'])
    t.update({'b': 2.0, 'd': 4.0, 'zzz': 1.0})
    assert t.n_fixed() == 2 and t.free_names() == ['a', 'c']
    np.testing.assert_array_equal(t.mask(), [False, True, False, True])
    full = t.expand([10.0, 30.0])
    np.testing.assert_array_equal(full, [10.0, 2.0, 30.0, 4.0])
    full[0] = -1.0           # a fresh array every time
    np.testing.assert_array_equal(t.expand([1, 3]), [1.0, 2.0, 3.0, 4.0])
    t.update({'b': None})
    assert t.free_names() == ['a', 'b', 'c']
    with pytest.raises(ValueError):
        t.update(3)


def test_classes_subclass_chi_when_chi_is_importable():
    """chi/_inference.py:423-429 accepts a log-posterior only if it is an
    instance of chi.LogPosterior / chi.HierarchicalLogPosterior: with a
    ``chi`` on the path the classes here derive from chi's own (no patch to
    chi).  Checked in a fresh interpreter against the reference modules
    (build container only: skipped where /root/reference is absent)."""
    from oracle import refload
    if not refload.available():
        pytest.skip('reference not present')
    code = (
        'import sys; sys.path.insert(0, %r)\n'
        'from oracle import refload\n'
        'chi = refload.load()\n'
        'import chi_b200\n'
        'for name in ("LogPosterior", "HierarchicalLogPosterior", '
        '"LogLikelihood", "HierarchicalLogLikelihood", '
        '"PopulationFilterLogPosterior", "MechanisticModel", "ErrorModel", '
        '"PopulationModel", "CovariateModel", "PopulationFilter"):\n'
        '    assert issubclass(getattr(chi_b200, name), getattr(chi, name)), '
        'name\n'
        'post = object.__new__(chi_b200.HierarchicalLogPosterior)\n'
        'assert isinstance(post, (chi.LogPosterior, '
        'chi.HierarchicalLogPosterior))\n'
        'print("ok")\n' % ROOT)
    out = subprocess.run([sys.executable, '-c', code], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stdout
