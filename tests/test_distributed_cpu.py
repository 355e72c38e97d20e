"""World-size-2 `gloo` test of the individual-sharded evaluation (CPU).

The per-rank partial sums come from the CPU oracle (the CUDA path needs a
GPU); what is under test is the host-side logic of chi_b200._distributed:
shard bounds, packing, the all-reduce and the prior handling.
"""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from chi_b200 import _distributed
from oracle import stats


def test_shard_bounds_cover_everything():
    for n, w in [(10, 3), (8, 8), (5, 2), (1000, 7), (3, 4)]:
        bounds = [_distributed.shard_bounds(n, r, w) for r in range(w)]
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        for a, b in zip(bounds[:-1], bounds[1:]):
            assert a[1] == b[0]
        sizes = [e - b for b, e in bounds]
        assert max(sizes) - min(sizes) <= 1


def _toy_problem():
    rng = np.random.default_rng(4)
    n_ids = 7
    times = [np.sort(rng.uniform(0, 2, 4)) for _ in range(n_ids)]
    obs = [rng.uniform(0.5, 3, 4) for _ in range(n_ids)]
    blocks = [('pooled', 1), ('lognormal', 1, False), ('pooled', 1)]

    def individual(i, psi):
        def sim(p, sens):
            y0, g = p
            y = (y0 * np.exp(g * times[i]))[None, :]
            if not sens:
                return y
            s = np.empty((4, 1, 2))
            s[:, 0, 0] = np.exp(g * times[i])
            s[:, 0, 1] = times[i] * y[0]
            return y, s
        return stats.loglikelihood_s1(
            sim, 2, ['lognormal'], [obs[i]], np.ones((1, 4), bool), psi, True)
    return n_ids, blocks, individual


class _Prior(object):
    def n_parameters(self):
        return 4

    def evaluateS1(self, x):
        x = np.asarray(x)
        return -0.5 * float(np.sum(x ** 2)), -x

    def __call__(self, x):
        return self.evaluateS1(x)[0]


class _FakePosterior(object):
    """Just the bookkeeping _distributed needs."""

    def __init__(self, n_ids, n_params, n_top):
        self._n = (n_ids, n_params, n_top)

    def get_log_likelihood(self):
        return self

    def get_log_prior(self):
        return _Prior()

    def n_ids(self):
        return self._n[0]

    def n_parameters(self, exclude_bottom_level=False):
        return self._n[2] if exclude_bottom_level else self._n[1]


def _worker(rank, world, port, queue):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n_ids, blocks, individual = _toy_problem()
    n_hdim, n_top = 1, 4
    n_params = n_ids * n_hdim + n_top
    n_bottom = n_ids * n_hdim

    def partial_fn(X, begin, end):
        scores, grads = [], []
        for x in X:
            lay = stats.PopLayout(blocks, end - begin)
            xs = np.concatenate([x[begin * n_hdim:end * n_hdim],
                                 x[n_bottom:]])
            s, g = stats.hierarchical_s1(
                lay, lambda i, p: individual(begin + i, p), xs)
            full = np.zeros(n_params)
            full[begin * n_hdim:end * n_hdim] = g[:lay.n_bottom]
            full[n_bottom:] = g[lay.n_bottom:]
            scores.append(s)
            grads.append(full)
        return np.array(scores), np.array(grads)

    post = _FakePosterior(n_ids, n_params, n_top)
    rng = np.random.default_rng(0)
    X = np.hstack([0.3 * rng.normal(size=(3, n_bottom)),
                   np.tile([1.0, 0.2, 0.4, 0.3], (3, 1))])
    out = {}
    for gather in (True, False):
        sharded = _distributed.ShardedHierarchicalLogPosterior(
            post, partial_fn=partial_fn, gather_bottom=gather)
        out[gather] = sharded.evaluateS1_batch(X)
        out[('shard', gather)] = sharded.shard()
    queue.put((rank, out, X))
    dist.barrier()
    dist.destroy_process_group()


def test_individual_sharding_two_ranks_gloo():
    sock = socket.socket()
    sock.bind(('127.0.0.1', 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue))
             for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_ids, blocks, individual = _toy_problem()
    lay = stats.PopLayout(blocks, n_ids)
    prior = _Prior()
    X = results[0][2]
    ref = [stats.posterior_s1(lay, individual, prior, x) for x in X]
    for rank, out, _ in results:
        begin, end = out[('shard', True)]
        assert (begin, end) == _distributed.shard_bounds(n_ids, rank, 2)
        score, grad = out[True]
        for c in range(len(X)):
            assert score[c] == pytest.approx(ref[c][0], rel=1e-12)
            np.testing.assert_allclose(grad[c], ref[c][1], rtol=1e-10,
                                       atol=1e-12)
        # without the gather only the rank's own bottom block is filled
        score, grad = out[False]
        for c in range(len(X)):
            assert score[c] == pytest.approx(ref[c][0], rel=1e-12)
            np.testing.assert_allclose(
                grad[c, lay.n_bottom:], ref[c][1][lay.n_bottom:], rtol=1e-10)
            np.testing.assert_allclose(
                grad[c, begin:end], ref[c][1][begin:end], rtol=1e-10)
