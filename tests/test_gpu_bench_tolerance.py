"""Parity of the CUDA path AT THE BENCH'S INTEGRATOR TOLERANCE.

The rest of the GPU suite asserts the contract (summed log-pdf 1e-8
relative, gradient 1e-6 relative; BASELINE.json north_star) with the
integrator at rtol 1e-11.  bench.py runs at ``_synthetic.BENCH_RTOL`` /
``BENCH_ATOL``; these tests check THAT setting against the exact ODE solution
(oracle/ode_exact.py) + the numpy restatement of chi's algebra
(oracle/stats.py), at sizes well above the golden vectors (>= 200
individuals), with widely spread parameter vectors (the individual-major work
order of the solve kernel then has lanes of very different step counts in one
warp), for configs C2 (evaluateS1), C4 (forward-only batch) and C5.
"""
import numpy as np
import pytest

from chi_b200 import _synthetic
from oracle import ode_exact, stats

pytestmark = pytest.mark.gpu

SCORE_REL = 1e-8      # contract: summed log-pdf
GRAD_REL = 1e-6       # contract: gradient (relative to its max-norm)


def spread_vectors(prob, n_vectors, spread, seed):
    """Parameter vectors spread widely around the data-generating values:
    additive on eta and the log-means, multiplicative on the (positive)
    standard deviations and sigmas, small positive initial amounts."""
    return _synthetic.spread_parameter_vectors(prob, n_vectors, spread, seed)


def oracle_scores(prob, X, error, with_grad, covariates=None):
    n_ids = len(prob['times'])
    if prob['covariates'] is not None:
        cpm = prob['population_model'].get_population_models()[1]
        pidx, didx = cpm._covariate_model.get_set_population_parameters()
        blocks = [('pooled', 2),
                  ('covariate', ('lognormal', 5, False), 2, pidx, didx),
                  ('pooled', 2 if error == 'cam' else 1)]
    else:
        blocks = [('pooled', 2), ('lognormal', 5, False),
                  ('pooled', 2 if error == 'cam' else 1)]
    lay = stats.PopLayout(blocks, n_ids)
    omodel = ode_exact.two_compartment(True)
    omodel.set_outputs(['central.drug_concentration'])
    events = [[tuple(r) for r in reg.as_array()] for reg in prob['regimens']]

    def individual(i, psi, grad):
        times = prob['times'][i]

        def simulate(p, with_sens):
            return ode_exact.simulate(
                omodel, p, times, events[i], sensitivities=with_sens)
        return stats.loglikelihood_s1(
            simulate, 7, [error], [prob['observations'][i]],
            np.ones((1, len(times)), bool), psi, grad)

    scores, grads = [], []
    for x in X:
        if with_grad:
            s, g = stats.posterior_s1(
                lay, lambda i, p: individual(i, p, True), prob['log_prior'],
                x, prob['covariates'])
            grads.append(g)
        else:
            s = stats.posterior_call(
                lay, lambda i, p: individual(i, p, False), prob['log_prior'],
                x, prob['covariates'])
        scores.append(s)
    return np.array(scores), (np.array(grads) if with_grad else None)


def test_c2_evaluateS1_at_bench_tolerance_200_individuals_64_vectors():
    prob = _synthetic.two_compartment_problem(
        200, seed=11, n_doses=3, rtol=_synthetic.BENCH_RTOL,
        atol=_synthetic.BENCH_ATOL)
    post = prob['log_posterior']
    X = spread_vectors(prob, 64, 0.3, seed=1)
    score, grad = post.evaluateS1_batch(X)
    ref, gref = oracle_scores(prob, X, 'lognormal', True)
    rel = np.abs(score - ref) / np.abs(ref)
    assert np.max(rel) <= SCORE_REL, (np.max(rel), np.median(rel))
    gerr = np.max(np.abs(grad - gref), axis=1) / np.max(np.abs(gref), axis=1)
    assert np.max(gerr) <= GRAD_REL, np.max(gerr)
    # the single-vector entry point runs one warp per system, every
    # sensitivity column with the state under its own step-size control
    # (chi_b200_problem_set_small_batch): the same contract against the
    # oracle, and agreement with the batch far inside it
    s0, g0 = post.evaluateS1(X[3])
    assert abs(s0 - ref[3]) / abs(ref[3]) <= SCORE_REL
    assert np.max(np.abs(g0 - gref[3])) / np.max(np.abs(gref[3])) <= GRAD_REL
    assert s0 == pytest.approx(score[3], rel=3e-9)
    np.testing.assert_allclose(g0, grad[3], rtol=1e-6,
                               atol=1e-7 * np.max(np.abs(grad[3])))
    # ... and with that form switched off: the thread-per-system kernel in
    # system order, bit for bit the arithmetic of the batch
    prob['log_likelihood'].set_small_batch(0)
    s1, g1 = post.evaluateS1(X[3])
    prob['log_likelihood'].set_small_batch(-1)
    assert s1 == pytest.approx(score[3], rel=1e-12)
    np.testing.assert_allclose(g1, grad[3], rtol=1e-9,
                               atol=1e-12 * np.max(np.abs(grad[3])))


def test_c4_forward_batch_at_bench_tolerance_500_individuals():
    """C4: HierarchicalLogPosterior.__call__ semantics batched over chains
    (forward-only kernel), 500 individuals x 32 vectors."""
    prob = _synthetic.two_compartment_problem(
        500, seed=12, n_doses=3, rtol=_synthetic.BENCH_RTOL,
        atol=_synthetic.BENCH_ATOL)
    post = prob['log_posterior']
    X = spread_vectors(prob, 32, 0.3, seed=2)
    score = post.evaluate_batch(X)
    ref, _ = oracle_scores(prob, X[:8], 'lognormal', False)
    rel = np.abs(score[:8] - ref) / np.abs(ref)
    assert np.max(rel) <= SCORE_REL, (np.max(rel), np.median(rel))
    # forward-only and sensitivity kernels take different steps: they agree
    # with one another to the contract as well
    s1, _ = post.evaluateS1_batch(X)
    assert np.max(np.abs(s1 - score) / np.abs(score)) <= 2 * SCORE_REL


def test_c5_covariates_cam_at_bench_tolerance_300_individuals():
    prob = _synthetic.two_compartment_problem(
        300, seed=13, n_doses=3, error='cam', covariates=True,
        rtol=_synthetic.BENCH_RTOL, atol=_synthetic.BENCH_ATOL)
    post = prob['log_posterior']
    X = spread_vectors(prob, 8, 0.1, seed=3)
    score, grad = post.evaluateS1_batch(X)
    ref, gref = oracle_scores(prob, X, 'cam', True)
    rel = np.abs(score - ref) / np.abs(ref)
    assert np.max(rel) <= SCORE_REL, (np.max(rel), np.median(rel))
    gerr = np.max(np.abs(grad - gref), axis=1) / np.max(np.abs(gref), axis=1)
    assert np.max(gerr) <= GRAD_REL, np.max(gerr)


def test_device_all_reduce_inside_the_library_single_rank():
    """The product path of the individuals-sharded mode on one GPU: a
    one-rank communicator (chi_b200_problem_comm_init) makes every evaluation
    go through pack -> ncclAllReduce -> unpack inside the library; the result
    must equal the plain evaluation, through both entry points."""
    import torch
    prob = _synthetic.two_compartment_problem(
        37, seed=4, n_doses=2, rtol=1e-9, atol=1e-11)
    post = prob['log_posterior']
    ll = prob['log_likelihood']
    X = spread_vectors(prob, 5, 0.1, seed=5)
    s_ref, g_ref = post.evaluateS1_batch(X)
    f_ref = post.evaluate_batch(X)
    comm_id = ll.new_communicator_id()
    assert len(comm_id) == 128
    ll.set_shard(0, 37)
    ll.set_communicator(1, 0, comm_id)
    assert ll.has_communicator()
    s, g = post.evaluateS1_batch(X)
    np.testing.assert_allclose(s, s_ref, rtol=1e-14)
    np.testing.assert_allclose(g, g_ref, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(post.evaluate_batch(X), f_ref, rtol=1e-14)
    # device entry point
    dev = torch.device('cuda', 0)
    d_x = torch.from_numpy(X).to(dev)
    d_s = torch.empty(len(X), dtype=torch.float64, device=dev)
    d_g = torch.zeros(X.shape, dtype=torch.float64, device=dev)
    d_st = torch.zeros(len(X), dtype=torch.int32, device=dev)
    ll.evaluate_device(d_x.data_ptr(), len(X), True, d_s.data_ptr(),
                       d_g.data_ptr(), d_st.data_ptr(),
                       torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize(dev)
    lp = np.array([prob['log_prior'](x[ll._n_bottom:]) for x in X])
    np.testing.assert_allclose(d_s.cpu().numpy() + lp, s_ref, rtol=1e-13)
    # an invalid population (negative sigma) and a failed solve keep their
    # conventions through the packed flags
    Xb = X.copy()
    Xb[1, ll._n_bottom + 7] = -0.3          # a log-std of the LogNormal block
    with np.errstate(all='ignore'):
        sb, _ = post.evaluateS1_batch(Xb)
    assert sb[1] == -np.inf and np.isfinite(sb[0])
