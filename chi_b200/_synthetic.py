"""Synthetic populations for tests and benchmarks.

The workloads mirror BASELINE.json's configs: a two-compartment PK model
with per-individual infusion dosing, log-normally distributed individual
parameters and a LogNormal or ConstantAndMultiplicativeGaussian error model.
Observations are synthesised with the product's own batched ``simulate``
(GPU) plus host-side numpy noise, so no reference or oracle code is involved.
"""
import numpy as np

import chi_b200
from chi_b200._log_pdfs import _DeviceLikelihoods

# docs/source/getting_started/code/3_fitting_models_1.py:159-169
# (V_c, k_e, k_cp, k_pc, V_p) = (2, 1, 5, 1, 10)
TWO_CMT_TYPICAL = np.array([2.0, 1.0, 5.0, 1.0, 10.0])

# Integrator tolerances of bench.py.  tests/test_gpu_bench_tolerance.py checks
# THIS setting against the exact solution under the contract (summed log-pdf
# 1e-8 relative, gradient 1e-6 relative).
BENCH_RTOL = 1e-8
BENCH_ATOL = 1e-10


def two_compartment_model(direct=True):
    model = chi_b200.library.ModelLibrary().two_compartment_pk_model()
    model.set_administration('central', direct=direct)
    model.set_outputs(['central.drug_concentration'])
    return model


def two_compartment_design(n_ids, seed=0, n_obs=10, n_doses=1):
    """Per-individual observation times and infusion regimens (config C2):
    dose ~ U(5, 15) infused over 1 time unit (optionally repeated every 8),
    n_obs observation times ~ sorted U(0.25, 24)."""
    rng = np.random.default_rng(seed)
    times = [np.sort(rng.uniform(0.25, 24.0, n_obs)) for _ in range(n_ids)]
    doses = rng.uniform(5.0, 15.0, n_ids)
    regimens = []
    for i in range(n_ids):
        r = chi_b200.DosingRegimen()
        if n_doses > 1:
            r.schedule(doses[i] / 1.0, 0.0, 1.0, 8.0, n_doses)
        else:
            r.schedule(doses[i] / 1.0, 0.0, 1.0, 0, 0)
        regimens.append(r)
    return times, regimens


def two_compartment_problem(n_ids, seed=0, n_obs=10, n_doses=1,
                            error='lognormal', covariates=False,
                            rtol=None, atol=None):
    """Builds the hierarchical log-posterior of configs C2 / C5.

    Returns a dict with the posterior, a parameter vector near the truth
    (``x0``) and the pieces it was assembled from.
    """
    rng = np.random.default_rng(seed + 12345)
    model = two_compartment_model(True)
    if rtol is not None:
        model.set_tolerance(abs_tol=atol, rel_tol=rtol)
    times, regimens = two_compartment_design(n_ids, seed, n_obs, n_doses)

    mu_log = np.log(TWO_CMT_TYPICAL)
    sigma_log = np.full(5, 0.3)
    n_cov = 2 if covariates else 0
    X = rng.normal(size=(n_ids, n_cov)) if covariates else None
    eta = rng.normal(size=(n_ids, 5))
    if covariates:
        inner = chi_b200.LogNormalModel(n_dim=5, centered=False)
        cpm = chi_b200.CovariatePopulationModel(
            inner, chi_b200.LinearCovariateModel(n_cov=n_cov))
        pidx, didx = cpm._covariate_model.get_set_population_parameters()
        beta = np.zeros((len(pidx), n_cov))
        # modest covariate effects on the log means only
        beta[np.asarray(pidx) == 0] = rng.normal(
            scale=0.05, size=(int(np.sum(np.asarray(pidx) == 0)), n_cov))
        mu_i = mu_log[None, :] + np.zeros((n_ids, 5))
        for s, (p, d) in enumerate(zip(pidx, didx)):
            if p == 0:
                mu_i[:, d] += X @ beta[s]
        middle = cpm
        theta_mid = np.concatenate([mu_log, sigma_log, beta.reshape(-1)])
    else:
        mu_i = mu_log[None, :] + np.zeros((n_ids, 5))
        middle = chi_b200.LogNormalModel(n_dim=5, centered=False)
        theta_mid = np.concatenate([mu_log, sigma_log])
    psi_mid = np.exp(mu_i + sigma_log[None, :] * eta)

    if error == 'lognormal':
        error_model = chi_b200.LogNormalErrorModel()
        sig = np.array([0.1])
    elif error == 'cam':
        error_model = chi_b200.ConstantAndMultiplicativeGaussianErrorModel()
        sig = np.array([0.05, 0.1])
    elif error == 'gaussian':
        error_model = chi_b200.GaussianErrorModel()
        sig = np.array([0.05])
    else:
        raise ValueError(error)

    population_model = chi_b200.ComposedPopulationModel([
        chi_b200.PooledModel(n_dim=2), middle,
        chi_b200.PooledModel(n_dim=len(sig))])
    population_model.set_dim_names(
        model.parameters() + error_model.get_parameter_names())

    # synthesise observations with the product's batched simulate
    psi = np.hstack([np.zeros((n_ids, 2)), psi_mid])
    dummy = [np.ones(len(t)) for t in times]
    dev = _DeviceLikelihoods(
        model, [error_model._kind], [[t] for t in times],
        [[y] for y in dummy], regimens)
    ybar, _, status = dev.simulate(psi)
    if np.any(status != 0):
        raise RuntimeError('synthetic data generation failed')
    if error == 'lognormal':
        noise = rng.lognormal(-sig[0]**2 / 2, sig[0], size=len(ybar))
        yobs = ybar * noise
    elif error == 'cam':
        yobs = ybar + (sig[0] + sig[1] * ybar) * rng.normal(size=len(ybar))
        yobs = np.abs(yobs) + 1e-9
    else:
        yobs = ybar + sig[0] * rng.normal(size=len(ybar))
    off = np.concatenate([[0], np.cumsum([len(t) for t in times])])
    observations = [yobs[off[i]:off[i + 1]] for i in range(n_ids)]
    dev.problem.close()

    log_likelihood = chi_b200.HierarchicalLogLikelihood.from_arrays(
        model, [error_model], times, observations, population_model,
        regimens=regimens, covariates=X)
    n_top = log_likelihood.n_parameters(exclude_bottom_level=True)
    priors = [chi_b200.GaussianLogPrior(0.0, 1.0)] * 2
    for k in range(len(theta_mid)):
        if k < 5:
            priors.append(chi_b200.GaussianLogPrior(mu_log[k], 1.0))
        elif k < 10:
            priors.append(chi_b200.LogNormalLogPrior(np.log(0.3), 0.5))
        else:
            priors.append(chi_b200.GaussianLogPrior(0.0, 0.5))
    for s in sig:
        priors.append(chi_b200.LogNormalLogPrior(np.log(s), 0.5))
    log_prior = chi_b200.ComposedLogPrior(*priors)
    assert log_prior.n_parameters() == n_top
    log_posterior = chi_b200.HierarchicalLogPosterior(
        log_likelihood, log_prior)
    x0 = np.concatenate([eta.reshape(-1), [0.0, 0.0], theta_mid, sig])
    return {
        'log_posterior': log_posterior, 'log_likelihood': log_likelihood,
        'log_prior': log_prior, 'x0': x0, 'model': model,
        'error_model': error_model, 'times': times,
        'observations': observations, 'regimens': regimens,
        'population_model': population_model, 'covariates': X,
        'theta_mid': theta_mid, 'sigma': sig}


def spread_parameter_vectors(prob, n_vectors, spread, seed=0):
    """Parameter vectors spread around the data-generating vector ``x0`` of a
    problem built by :func:`two_compartment_problem`: additive N(0, spread)
    on the etas and the log-means, multiplicative (1 + spread N(0,1)) on the
    (positive) log-standard deviations and error-model sigmas, |N(0, 0.1
    spread)| initial amounts, additive 0.05 spread N(0,1) on covariate
    effects (the individuals' log-standard deviations stay positive)."""
    rng = np.random.default_rng(seed)
    x0 = np.asarray(prob['x0'], dtype=float)
    n_ids = len(prob['times'])
    n_bottom = 5 * n_ids
    n_sig = len(prob['sigma'])
    n_params = len(x0)
    X = np.tile(x0, (n_vectors, 1))
    X[:, :n_bottom] += spread * rng.normal(size=(n_vectors, n_bottom))
    top = n_bottom
    X[:, top:top + 2] = np.abs(0.1 * spread * rng.normal(size=(n_vectors, 2)))
    X[:, top + 2:top + 7] += spread * rng.normal(size=(n_vectors, 5))
    X[:, top + 7:top + 12] *= np.abs(
        1.0 + spread * rng.normal(size=(n_vectors, 5)))
    n_beta = n_params - n_sig - (top + 12)
    if n_beta > 0:
        X[:, top + 12:top + 12 + n_beta] += 0.05 * spread * rng.normal(
            size=(n_vectors, n_beta))
    X[:, n_params - n_sig:] *= np.abs(
        1.0 + spread * rng.normal(size=(n_vectors, n_sig)))
    return X
