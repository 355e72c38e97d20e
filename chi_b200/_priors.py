"""Minimal host-side log-priors with the ``pints.LogPrior`` protocol.

chi takes its priors from pints (``pints.LogPrior``; call sites
chi/_log_pdfs.py:444, 468, 486, 1229, 1247, 1262).  pints is a caller of the
hot path, not part of it, and is not installed in this image; these classes
are duck-typed stand-ins so that posteriors can be assembled in tests and
benchmarks.  A real ``pints.LogPrior`` works in their place.  Priors act on
the handful of top-level parameters only and stay on the host, as in chi.
"""
import numpy as np

try:
    import pints as _pints
    _Base = _pints.LogPrior
except Exception:  # pragma: no cover
    _Base = object


class GaussianLogPrior(_Base):
    """``pints.GaussianLogPrior(mean, sd)``."""

    def __init__(self, mean, sd):
        self._mean = float(mean)
        self._sd = float(sd)
        if self._sd <= 0:
            raise ValueError('sd must be positive.')

    def __call__(self, x):
        z = (float(np.asarray(x).reshape(-1)[0]) - self._mean) / self._sd
        return -0.5 * np.log(2 * np.pi) - np.log(self._sd) - 0.5 * z * z

    def evaluateS1(self, x):
        z = (float(np.asarray(x).reshape(-1)[0]) - self._mean) / self._sd
        return self(x), np.array([-z / self._sd])

    def n_parameters(self):
        return 1

    def sample(self, n=1):
        return np.random.normal(self._mean, self._sd, size=(n, 1))


class LogNormalLogPrior(_Base):
    """``pints.LogNormalLogPrior(log_mean, scale)``."""

    def __init__(self, log_mean, scale):
        self._mu = float(log_mean)
        self._s = float(scale)
        if self._s <= 0:
            raise ValueError('scale must be positive.')

    def __call__(self, x):
        x = float(np.asarray(x).reshape(-1)[0])
        if x <= 0:
            return -np.inf
        z = (np.log(x) - self._mu) / self._s
        return -0.5 * np.log(2 * np.pi) - np.log(self._s) - np.log(x) \
            - 0.5 * z * z

    def evaluateS1(self, x):
        xv = float(np.asarray(x).reshape(-1)[0])
        if xv <= 0:
            return -np.inf, np.array([np.inf])
        z = (np.log(xv) - self._mu) / self._s
        return self(x), np.array([-(1 + z / self._s) / xv])

    def n_parameters(self):
        return 1

    def sample(self, n=1):
        return np.random.lognormal(self._mu, self._s, size=(n, 1))


class UniformLogPrior(_Base):
    """``pints.UniformLogPrior(lower, upper)`` (one dimension)."""

    def __init__(self, lower, upper):
        self._a = float(lower)
        self._b = float(upper)
        if not self._b > self._a:
            raise ValueError('upper must exceed lower.')

    def __call__(self, x):
        x = float(np.asarray(x).reshape(-1)[0])
        if x < self._a or x >= self._b:
            return -np.inf
        return -np.log(self._b - self._a)

    def evaluateS1(self, x):
        return self(x), np.zeros(1)

    def n_parameters(self):
        return 1

    def sample(self, n=1):
        return np.random.uniform(self._a, self._b, size=(n, 1))


def _batch_1d(prior, x):
    """Vectorised (score, d score / dx) of a one-dimensional prior over an
    array of values; used by the batched posterior entry points."""
    x = np.asarray(x, dtype=float)
    if isinstance(prior, GaussianLogPrior):
        z = (x - prior._mean) / prior._sd
        return (-0.5 * np.log(2 * np.pi) - np.log(prior._sd) - 0.5 * z * z,
                -z / prior._sd)
    if isinstance(prior, LogNormalLogPrior):
        with np.errstate(all='ignore'):
            z = (np.log(x) - prior._mu) / prior._s
            score = -0.5 * np.log(2 * np.pi) - np.log(prior._s) - np.log(x) \
                - 0.5 * z * z
            grad = -(1 + z / prior._s) / x
        bad = x <= 0
        score = np.where(bad, -np.inf, score)
        grad = np.where(bad, np.inf, grad)
        return score, grad
    if isinstance(prior, UniformLogPrior):
        inside = (x >= prior._a) & (x < prior._b)
        return (np.where(inside, -np.log(prior._b - prior._a), -np.inf),
                np.zeros_like(x))
    out = [prior.evaluateS1(np.array([v])) for v in x]
    return (np.array([o[0] for o in out]),
            np.array([np.asarray(o[1]).reshape(-1)[0] for o in out]))


class ComposedLogPrior(_Base):
    """``pints.ComposedLogPrior(*priors)``: independent product.  When every
    factor is one of the one-dimensional priors above the whole product is
    evaluated with a handful of array expressions over the parameters (the
    single-vector calls of a sampler spend tens of microseconds otherwise)."""

    def __init__(self, *priors):
        if not priors:
            raise ValueError('Must have at least one sub-prior.')
        self._priors = list(priors)
        self._n = sum(p.n_parameters() for p in self._priors)
        self._table = None
        known = (GaussianLogPrior, LogNormalLogPrior, UniformLogPrior)
        if all(type(p) in known for p in self._priors):
            kind = np.array([known.index(type(p)) for p in self._priors])
            a = np.array([p._mean if k == 0 else (p._mu if k == 1 else p._a)
                          for p, k in zip(self._priors, kind)])
            b = np.array([p._sd if k == 0 else (p._s if k == 1 else p._b)
                          for p, k in zip(self._priors, kind)])
            const = np.where(
                kind == 2, -np.log(np.where(kind == 2, b - a, 1.0)),
                -0.5 * np.log(2 * np.pi) - np.log(np.where(kind == 2, 1.0, b)))
            self._table = (kind == 0, kind == 1, kind == 2, a, b, const,
                           bool(np.any(kind == 1)), bool(np.any(kind == 2)))

    def _fast(self, X):
        """(scores, gradients) of the rows of X through the table."""
        gauss, logn, unif, a, b, const, any_logn, any_unif = self._table
        v = X
        bad = None
        if any_logn:
            pos = X > 0
            bad = np.any(logn & ~pos, axis=-1)
            with np.errstate(all='ignore'):
                v = np.where(logn, np.log(np.where(logn & pos, X, 1.0)), X)
        z = (v - a) / b
        term = const - 0.5 * z * z
        grad = -z / b
        if any_logn:
            term = np.where(logn, term - v, term)
            with np.errstate(all='ignore'):
                grad = np.where(logn, (grad - 1.0) / X, grad)
        if any_unif:
            inside = (X >= a) & (X < b)
            out = np.any(unif & ~inside, axis=-1)
            bad = out if bad is None else (bad | out)
            term = np.where(unif, const, term)
            grad = np.where(unif, 0.0, grad)
        score = np.sum(term, axis=-1)
        if bad is not None and np.any(bad):
            score = np.where(bad, -np.inf, score)
            if any_logn:
                grad = np.where(logn & ~(X > 0), np.inf, grad)
        return score, grad

    def __call__(self, x):
        x = np.asarray(x, dtype=float).reshape(-1)
        if self._table is not None:
            return float(self._fast(x)[0])
        out = 0.0
        lo = 0
        for p in self._priors:
            hi = lo + p.n_parameters()
            out += p(x[lo:hi])
            lo = hi
        return out

    def evaluateS1(self, x):
        x = np.asarray(x, dtype=float).reshape(-1)
        if self._table is not None:
            score, grad = self._fast(x)
            return float(score), grad
        out = 0.0
        grads = []
        lo = 0
        for p in self._priors:
            hi = lo + p.n_parameters()
            v, g = p.evaluateS1(x[lo:hi])
            out += v
            grads.append(np.asarray(g, dtype=float).reshape(-1))
            lo = hi
        return out, np.concatenate(grads)

    def evaluateS1_batch(self, X):
        """Scores (n,) and gradients (n, n_parameters) of many vectors."""
        X = np.asarray(X, dtype=float).reshape(-1, self._n)
        if self._table is not None:
            return self._fast(X)
        score = np.zeros(len(X))
        grad = np.empty(X.shape)
        lo = 0
        for p in self._priors:
            if p.n_parameters() != 1:
                out = [p.evaluateS1(x[lo:lo + p.n_parameters()]) for x in X]
                score += np.array([o[0] for o in out])
                grad[:, lo:lo + p.n_parameters()] = np.array(
                    [o[1] for o in out])
            else:
                s, g = _batch_1d(p, X[:, lo])
                score += s
                grad[:, lo] = g
            lo += p.n_parameters()
        return score, grad

    def n_parameters(self):
        return self._n

    def sample(self, n=1):
        return np.hstack([p.sample(n) for p in self._priors])
