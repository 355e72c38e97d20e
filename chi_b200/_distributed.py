"""Sharding of a hierarchical log-posterior over the GPUs of one node.

Individuals are independent given (theta, eta_i) (chi/_log_pdfs.py:289-292),
so a population shards by individual: rank r evaluates the contiguous block
``shard_bounds(n_ids, r, world)`` and produces

* a partial score            sum_i [log p(D_i|psi_i) + log p(eta_i|theta)],
* a partial top-level gradient (pooled sums, d mu, d sigma, d beta) and
* its own block of the bottom-level gradient (d eta_i, disjoint per rank).

One all-reduce (sum) of ``[score | d theta]`` (8 B ... a few KB per parameter
vector) completes the evaluation; the bottom blocks are disjoint and are
all-gathered only if the caller wants the full gradient on every rank.  The
prior acts on theta only and is added on the host after the reduction.

The collective goes through ``torch.distributed`` (NCCL over NVLink on the
GPU box, gloo in the CPU tests); the per-rank evaluation is the CUDA path
(`HierarchicalLogLikelihood.set_shard`), injectable for tests.
"""
import numpy as np


def shard_bounds(n_ids, rank, world):
    """Contiguous, balanced split of ``n_ids`` individuals."""
    base, extra = divmod(int(n_ids), int(world))
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin, end


def pack_partials(score, grad, n_bottom):
    """[score | d theta] of every parameter vector, one contiguous buffer."""
    score = np.asarray(score, dtype=float).reshape(-1, 1)
    grad = np.asarray(grad, dtype=float).reshape(len(score), -1)
    return np.ascontiguousarray(np.hstack([score, grad[:, n_bottom:]]))


def unpack_partials(buf, score, grad, n_bottom):
    score[:] = buf[:, 0]
    grad[:, n_bottom:] = buf[:, 1:]


class ShardedHierarchicalLogPosterior(object):
    """Evaluates a :class:`chi_b200.HierarchicalLogPosterior` whose
    individuals are sharded over the ranks of a ``torch.distributed`` group.

    :param log_posterior: the posterior (every rank holds the static data of
        all individuals; only its shard is evaluated).
    :param partial_fn: optional ``f(X, begin, end) -> (score, grad)``
        returning the shard's partial sums (tests inject the CPU oracle).
    """

    def __init__(self, log_posterior, group=None, partial_fn=None,
                 gather_bottom=True):
        import torch.distributed as dist
        self._dist = dist
        self._group = group
        self._post = log_posterior
        self._ll = log_posterior.get_log_likelihood()
        self._prior = log_posterior.get_log_prior()
        self._rank = dist.get_rank(group)
        self._world = dist.get_world_size(group)
        self._n_ids = log_posterior.n_ids()
        self._n_params = log_posterior.n_parameters()
        n_top = log_posterior.n_parameters(exclude_bottom_level=True)
        self._n_bottom = self._n_params - n_top
        self._begin, self._end = shard_bounds(
            self._n_ids, self._rank, self._world)
        self._gather_bottom = gather_bottom
        self._partial_fn = partial_fn
        if partial_fn is None:
            self._ll.set_shard(self._begin, self._end)

    def shard(self):
        return self._begin, self._end

    def _partials(self, X, with_grad):
        if self._partial_fn is not None:
            return self._partial_fn(X, self._begin, self._end)
        if with_grad:
            # page-locked result buffers; entries of other shards are zero
            return self._ll.evaluateS1_batch(X, copy=False)
        score = self._ll.evaluate_batch(X, copy=False)
        return score, None

    def _all_reduce(self, array):
        import torch
        t = torch.from_numpy(array)
        backend = self._dist.get_backend(self._group)
        if backend == 'nccl':
            t = t.cuda()
        self._dist.all_reduce(t, group=self._group)
        return t.cpu().numpy() if backend == 'nccl' else t.numpy()

    def _add_prior(self, X, score, grad):
        top = X[:, self._n_bottom:]
        if hasattr(self._prior, 'evaluateS1_batch'):
            p, s = self._prior.evaluateS1_batch(top)
        else:
            ps = [self._prior.evaluateS1(x) for x in top]
            p = np.array([q[0] for q in ps], dtype=float)
            s = np.array([q[1] for q in ps], dtype=float)
        bad = np.isinf(p)
        score += np.where(bad, 0.0, p)
        score[bad] = p[bad]
        if grad is not None:
            grad[:, self._n_bottom:] += np.where(bad[:, None], 0.0, s)
            grad[bad] = np.inf

    def evaluateS1_batch(self, X, with_grad=True, copy=True):
        """Scores and gradients of the parameter vectors ``X``.  With
        ``gather_bottom=False`` every rank's gradient holds the reduced
        top-level entries and its own individuals' bottom-level entries (the
        others are zero); with ``copy=False`` the arrays are the page-locked
        result buffers of the likelihood, valid until the next evaluation."""
        X = np.asarray(X, dtype=float).reshape(-1, self._n_params)
        score, grad = self._partials(X, with_grad)
        if self._partial_fn is not None or copy:
            score = np.array(score, dtype=float)
            grad = None if grad is None else np.array(grad, dtype=float)
        if grad is None:
            score[:] = self._all_reduce(
                np.ascontiguousarray(score.reshape(-1, 1)))[:, 0]
            grad_out = None
        elif self._gather_bottom:
            # bottom blocks are disjoint: a sum is a concatenation
            buf = self._all_reduce(
                np.ascontiguousarray(np.hstack([score[:, None], grad])))
            score[:] = buf[:, 0]
            grad[:] = buf[:, 1:]
            grad_out = grad
        else:
            buf = self._all_reduce(pack_partials(score, grad, self._n_bottom))
            unpack_partials(buf, score, grad, self._n_bottom)
            grad_out = grad
        self._add_prior(X, score, grad if grad is not None else None)
        return score, grad_out

    def evaluate_batch(self, X):
        return self.evaluateS1_batch(X, with_grad=False)[0]

    def evaluateS1(self, parameters):
        score, grad = self.evaluateS1_batch(np.asarray(parameters)[None, :])
        return float(score[0]), grad[0]

    def __call__(self, parameters):
        return float(self.evaluate_batch(np.asarray(parameters)[None, :])[0])

    def n_parameters(self):
        return self._n_params
