"""Ingestion of long-format measurement tables.

chi's ``ProblemModellingController`` (chi/_problems.py:23-830) turns a
DataFrame with ID / Time / Observable / Value / Dose / Duration columns into
one ``chi.LogLikelihood`` (with its own copy of the compiled simulator and its
own ``myokit.Protocol``) per individual (chi/_problems.py:246-376).  Here the
same table goes straight into the ragged device tensors of ONE problem
(`HierarchicalLogLikelihood.from_arrays`): per individual the sorted
(time, value) pairs of every output and one dose event per dose row
(``rate = dose / duration``, bolus = 0.01 time units).  SURVEY.md section 8f,
"next" row 2.
"""
import numpy as np

from chi_b200._log_pdfs import HierarchicalLogLikelihood
from chi_b200._mechanistic_models import DosingRegimen


def extract_measurements(data, outputs, output_observable_dict=None,
                         id_key='ID', time_key='Time', obs_key='Observable',
                         value_key='Value'):
    """Per individual and output the (times, values) arrays
    (chi/_problems.py:270-310)."""
    if output_observable_dict is None:
        output_observable_dict = {o: o for o in outputs}
    ids = list(data[id_key].dropna().unique())
    times, values = [], []
    for individual in ids:
        sub = data[data[id_key] == individual]
        t_i, v_i = [], []
        for output in outputs:
            rows = sub[sub[obs_key] == output_observable_dict[output]]
            rows = rows[rows[value_key].notnull() & rows[time_key].notnull()]
            order = np.argsort(rows[time_key].to_numpy(), kind='stable')
            t_i.append(rows[time_key].to_numpy(dtype=float)[order])
            v_i.append(rows[value_key].to_numpy(dtype=float)[order])
        times.append(t_i)
        values.append(v_i)
    return ids, times, values


def extract_dosing_regimens(data, ids, dose_key='Dose',
                            duration_key='Duration', id_key='ID',
                            time_key='Time'):
    """One dose event per dose row (chi/_problems.py:331-376)."""
    regimens = []
    for individual in ids:
        sub = data[data[id_key] == individual]
        sub = sub[sub[dose_key].notnull() & sub[time_key].notnull()]
        regimen = DosingRegimen()
        for _, row in sub.iterrows():
            duration = 0.01
            if duration_key is not None and duration_key in sub.columns:
                duration = row[duration_key]
                if np.isnan(duration):
                    duration = 0.01
            regimen.schedule(row[dose_key] / duration, row[time_key],
                             duration)
        regimens.append(regimen)
    return regimens


def extract_covariates(data, ids, covariate_names, covariate_dict=None,
                       id_key='ID', obs_key='Observable', value_key='Value'):
    """(n_ids, n_cov) covariate matrix (chi/_problems.py:312-329)."""
    if covariate_dict is None:
        covariate_dict = {c: c for c in covariate_names}
    out = np.empty((len(ids), len(covariate_names)))
    for idc, name in enumerate(covariate_names):
        rows = data[data[obs_key] == covariate_dict[name]]
        for idn, _id in enumerate(ids):
            vals = rows.loc[rows[id_key] == _id, value_key].dropna().values
            out[idn, idc] = vals[0]
    return out


def hierarchical_log_likelihood_from_dataframe(
        data, mechanistic_model, error_models, population_model,
        output_observable_dict=None, covariate_dict=None, dose_key='Dose',
        duration_key='Duration', id_key='ID', time_key='Time',
        obs_key='Observable', value_key='Value'):
    """Builds a :class:`chi_b200.HierarchicalLogLikelihood` from a
    long-format DataFrame (the schema of chi/library/data_library/*.csv)."""
    outputs = mechanistic_model.outputs()
    ids, times, values = extract_measurements(
        data, outputs, output_observable_dict, id_key, time_key, obs_key,
        value_key)
    regimens = None
    if dose_key is not None and dose_key in data.columns and \
            mechanistic_model.supports_dosing():
        regimens = extract_dosing_regimens(
            data, ids, dose_key, duration_key, id_key, time_key)
    covariates = None
    if population_model.n_covariates() > 0:
        covariates = extract_covariates(
            data, ids, population_model.get_covariate_names(),
            covariate_dict, id_key, obs_key, value_key)
    if len(outputs) == 1:
        times = [t[0] for t in times]
        values = [v[0] for v in values]
    return HierarchicalLogLikelihood.from_arrays(
        mechanistic_model, error_models, times, values, population_model,
        regimens=regimens, covariates=covariates, ids=ids)
