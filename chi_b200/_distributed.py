"""Sharding of a hierarchical log-posterior over the GPUs of one node.

Individuals are independent given (theta, eta_i) (chi/_log_pdfs.py:289-292),
so a population shards by individual: rank r evaluates the contiguous block
``shard_bounds(n_ids, r, world)`` and produces

* a partial score            sum_i [log p(D_i|psi_i) + log p(eta_i|theta)],
* a partial top-level gradient (pooled sums, d mu, d sigma, d beta) and
* its own block of the bottom-level gradient (d eta_i, disjoint per rank).

One all-reduce (sum) of ``[score | d theta]`` (8 B ... a few KB per parameter
vector) completes the evaluation.  On the GPU path that all-reduce happens
INSIDE the library, on the compute stream and on the device buffers, before
anything is copied to the host (``chi_b200_problem_comm_init``: a persistent
NCCL communicator per problem; csrc/core.cu: run_hier): this module only sets
the shard, hands the communicator id round and adds the (host) prior.  The
bottom blocks are disjoint and are gathered only if the caller asks for the
full gradient on every rank (``gather_bottom=True``).

``torch.distributed`` supplies the process group (rendezvous, the broadcast
of the communicator id; NCCL on the GPU box, gloo in the CPU tests).  For the
CPU tests the per-rank evaluation is injectable (``partial_fn``); the partial
sums are then reduced through ``torch.distributed`` itself.
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
        returning the shard's partial sums (tests inject the CPU oracle); the
        reduction then goes through ``torch.distributed`` on the host.
    :param gather_bottom: also return the other ranks' blocks of the
        bottom-level gradient (default: every rank gets the complete score
        and top-level gradient plus its own individuals' bottom-level block,
        the other individuals' entries are zero).
    """

    def __init__(self, log_posterior, group=None, partial_fn=None,
                 gather_bottom=False):
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
            # the library's own communicator: rank 0 creates the id, the
            # process group carries it to the others
            box = [self._ll.new_communicator_id() if self._rank == 0
                   else None]
            src = 0 if group is None else dist.get_global_rank(group, 0)
            dist.broadcast_object_list(box, src=src, group=group)
            self._ll.set_shard(self._begin, self._end)
            self._ll.set_communicator(self._world, self._rank, box[0])

    def shard(self):
        return self._begin, self._end

    def _host_all_reduce(self, array):
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

    def _gather(self, grad):
        """The bottom-level blocks are disjoint (zero elsewhere): a sum over
        the ranks is their concatenation.  Works on a copy."""
        bottom = np.ascontiguousarray(grad[:, :self._n_bottom])
        # an invalid prior fills a rank's own columns with inf on every rank
        # alike; take those rows out of the sum and restore them
        bad = np.isinf(grad[:, self._n_bottom:]).all(axis=1) \
            if grad.shape[1] > self._n_bottom else np.zeros(len(grad), bool)
        bottom[bad] = 0.0
        bottom = self._host_all_reduce(bottom)
        bottom[bad] = np.inf
        out = np.array(grad, dtype=float)
        out[:, :self._n_bottom] = bottom
        return out

    def _evaluate_injected(self, X, with_grad):
        score, grad = self._partial_fn(X, self._begin, self._end)
        score = np.array(score, dtype=float)
        grad = None if (grad is None or not with_grad) else np.array(
            grad, dtype=float)
        if grad is None:
            score[:] = self._host_all_reduce(
                np.ascontiguousarray(score.reshape(-1, 1)))[:, 0]
        else:
            buf = self._host_all_reduce(
                pack_partials(score, grad, self._n_bottom))
            unpack_partials(buf, score, grad, self._n_bottom)
            if self._gather_bottom:
                grad = self._gather(grad)
        self._add_prior(X, score, grad)
        return score, grad

    def evaluateS1_batch(self, X, with_grad=True, copy=True):
        """Scores and gradients of the parameter vectors ``X`` (the same on
        every rank, except for the other ranks' bottom-level blocks when
        ``gather_bottom`` is off).  With ``copy=False`` the arrays are the
        page-locked result buffers of the likelihood, valid until the next
        evaluation."""
        X = np.asarray(X, dtype=float).reshape(-1, self._n_params)
        if self._partial_fn is not None:
            return self._evaluate_injected(X, with_grad)
        # the likelihood's evaluation ends with the all-reduce on the device;
        # what is left is the host prior on the top-level slice
        if not with_grad:
            return self._post.evaluate_batch(X), None
        score, grad = self._post.evaluateS1_batch(X, copy=copy)
        if self._gather_bottom:
            grad = self._gather(grad)
        return score, grad

    def evaluate_batch(self, X):
        return self.evaluateS1_batch(X, with_grad=False)[0]

    def evaluateS1(self, parameters):
        score, grad = self.evaluateS1_batch(np.asarray(parameters)[None, :])
        return float(score[0]), grad[0]

    def __call__(self, parameters):
        return float(self.evaluate_batch(np.asarray(parameters)[None, :])[0])

    def n_parameters(self):
        return self._n_params
