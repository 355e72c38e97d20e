"""numpy restatement of chi's statistical layers (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  The restatement is deliberately *functional* (plain
arrays and small spec tuples, no class hierarchy) so that it shares no
structure with ``chi_b200`` and can be validated against the unmodified
reference classes through ``oracle/refload.py``.

Pinned by: ``tests/test_oracle_stats.py`` (reference known-answer tests) and
``tests/golden/stats_golden.json`` (outputs of the unmodified reference).
"""
import numpy as np

LOG_2PI = np.log(2 * np.pi)

# --------------------------------------------------------------------------
# Error models.  All take  sigma-parameters, model output ybar (n_obs,),
# sensitivities S (n_obs, P) and observations y (n_obs,).
# --------------------------------------------------------------------------


def gaussian_error_ll(params, ybar, y):
    """chi/_error_models.py:581-600."""
    sigma = params[0]
    if sigma <= 0:
        return -np.inf
    ybar = np.asarray(ybar, float)
    y = np.asarray(y, float)
    n = len(ybar)
    return -n * (LOG_2PI / 2 + np.log(sigma)) \
        - np.sum((ybar - y) ** 2) / sigma ** 2 / 2


def gaussian_error_s1(params, ybar, S, y):
    """chi/_error_models.py:625-664."""
    sigma = params[0]
    S = np.asarray(S, float)
    if sigma <= 0:
        return -np.inf, np.full(S.shape[1] + 1, np.inf)
    ybar = np.asarray(ybar, float)
    y = np.asarray(y, float)
    n = len(ybar)
    e = y - ybar
    sse = np.sum(e ** 2)
    ll = -n * (LOG_2PI / 2 + np.log(sigma)) - sse / sigma ** 2 / 2
    dpsi = np.sum(e[:, None] * S, axis=0) / sigma ** 2
    dsigma = sse / sigma ** 3 - n / sigma
    return ll, np.concatenate((dpsi, [dsigma]))


def lognormal_error_ll(params, ybar, y):
    """chi/_error_models.py:925-948."""
    sigma = params[0]
    ybar = np.asarray(ybar, float)
    y = np.asarray(y, float)
    if sigma <= 0 or np.any(ybar <= 0):
        return -np.inf
    n = len(ybar)
    return -n * (LOG_2PI / 2 + np.log(sigma)) - np.sum(np.log(y)) \
        - np.sum((np.log(ybar) - sigma ** 2 / 2 - np.log(y)) ** 2) \
        / sigma ** 2 / 2


def lognormal_error_s1(params, ybar, S, y):
    """chi/_error_models.py:977-1021."""
    sigma = params[0]
    ybar = np.asarray(ybar, float)
    y = np.asarray(y, float)
    S = np.asarray(S, float)
    if sigma <= 0 or np.any(ybar <= 0):
        return -np.inf, np.full(S.shape[1] + 1, np.inf)
    n = len(ybar)
    e = np.log(y) - np.log(ybar) + sigma ** 2 / 2
    sse = np.sum(e ** 2)
    ll = -n * (LOG_2PI / 2 + np.log(sigma)) - np.sum(np.log(y)) \
        - sse / sigma ** 2 / 2
    dpsi = np.sum((e / ybar)[:, None] * S, axis=0) / sigma ** 2
    dsigma = -np.sum(e) / sigma + sse / sigma ** 3 - n / sigma
    return ll, np.concatenate((dpsi, [dsigma]))


def cam_error_ll(params, ybar, y):
    """chi/_error_models.py:229-252 (ConstantAndMultiplicativeGaussian)."""
    sb, sr = params
    if sb <= 0 or sr <= 0:
        return -np.inf
    ybar = np.asarray(ybar, float)
    y = np.asarray(y, float)
    st = sb + sr * ybar
    n = len(ybar)
    return -n * LOG_2PI / 2 - np.sum(np.log(st)) \
        - np.sum((ybar - y) ** 2 / st ** 2) / 2


def cam_error_s1(params, ybar, S, y):
    """chi/_error_models.py:281-331."""
    sb, sr = params
    S = np.asarray(S, float)
    if sb <= 0 or sr <= 0:
        return -np.inf, np.full(S.shape[1] + 2, np.inf)
    ybar = np.asarray(ybar, float)
    y = np.asarray(y, float)
    st = sb + sr * ybar
    e = y - ybar
    e2 = e ** 2
    n = len(ybar)
    ll = -n * LOG_2PI / 2 - np.sum(np.log(st)) - np.sum(e2 / st ** 2) / 2
    w = e / st ** 2 - sr / st + sr * e2 / st ** 3
    dpsi = np.sum(w[:, None] * S, axis=0)
    dsb = np.sum(e2 / st ** 3) - np.sum(1 / st)
    dsr = np.sum(e2 / st ** 3 * ybar) - np.sum(ybar / st)
    return ll, np.concatenate((dpsi, [dsb, dsr]))


def multiplicative_error_ll(params, ybar, y):
    """chi/_error_models.py:1276-1299 (sigma_tot = sigma_rel * ybar)."""
    sr = params[0]
    if sr <= 0:
        return -np.inf
    ybar = np.asarray(ybar, float)
    y = np.asarray(y, float)
    st = sr * ybar
    n = len(ybar)
    return -n * LOG_2PI / 2 - np.sum(np.log(st)) \
        - np.sum((ybar - y) ** 2 / st ** 2) / 2


def multiplicative_error_s1(params, ybar, S, y):
    """chi/_error_models.py:1328-1375."""
    sr = params[0]
    S = np.asarray(S, float)
    if sr <= 0:
        return -np.inf, np.full(S.shape[1] + 1, np.inf)
    ybar = np.asarray(ybar, float)
    y = np.asarray(y, float)
    st = sr * ybar
    e = y - ybar
    e2 = e ** 2
    n = len(ybar)
    ll = -n * LOG_2PI / 2 - np.sum(np.log(st)) - np.sum(e2 / st ** 2) / 2
    w = e / st ** 2 - sr / st + sr * e2 / st ** 3
    dpsi = np.sum(w[:, None] * S, axis=0)
    dsr = np.sum(e2 / st ** 3 * ybar) - np.sum(ybar / st)
    return ll, np.concatenate((dpsi, [dsr]))


ERROR_MODELS = {
    # kind: (n_sigma, ll, s1)
    'gaussian': (1, gaussian_error_ll, gaussian_error_s1),
    'lognormal': (1, lognormal_error_ll, lognormal_error_s1),
    'cam': (2, cam_error_ll, cam_error_s1),
    'multiplicative': (1, multiplicative_error_ll, multiplicative_error_s1),
}


def error_pointwise_ll(kind, params, ybar, y):
    """compute_pointwise_ll of the four error models
    (chi/_error_models.py:254-279, 602-623, 950-975, 1301-1326)."""
    ybar = np.asarray(ybar, float)
    y = np.asarray(y, float)
    n = len(ybar)
    if kind == 'gaussian':
        s = params[0]
        if s <= 0:
            return np.full(n, -np.inf)
        return -(LOG_2PI / 2 + np.log(s)) - (ybar - y) ** 2 / s ** 2 / 2
    if kind == 'lognormal':
        s = params[0]
        if s <= 0 or np.any(ybar <= 0):
            return np.full(n, -np.inf)
        return -(LOG_2PI / 2 + np.log(s)) - np.log(y) \
            - (np.log(ybar) - s ** 2 / 2 - np.log(y)) ** 2 / s ** 2 / 2
    if kind == 'cam':
        sb, sr = params
        if sb <= 0 or sr <= 0:
            return np.full(n, -np.inf)
        st = sb + sr * ybar
    else:
        sr = params[0]
        if sr <= 0:
            return np.full(n, -np.inf)
        st = sr * ybar
    return -LOG_2PI / 2 - np.log(st) - (ybar - y) ** 2 / st ** 2 / 2


# --------------------------------------------------------------------------
# LogLikelihood  (one individual)
# --------------------------------------------------------------------------

def union_times(times_per_output):
    """Union grid + masks, chi/_log_pdfs.py:845-883."""
    union = sorted(set(float(t) for ts in times_per_output for t in ts))
    union = np.array(union, float)
    masks = np.zeros((len(times_per_output), len(union)), bool)
    for o, ts in enumerate(times_per_output):
        for t in ts:
            masks[o, union == t] = True
    return union, masks


def loglikelihood_s1(simulate, n_mech, error_kinds, obs, masks, x,
                     with_grad=True):
    """LogLikelihood.__call__ / evaluateS1, chi/_log_pdfs.py:802-843, 973-1027.

    ``simulate(psi, with_sens)`` returns ybar (n_out, n_t) [and S (n_t, n_out,
    n_mech)] on the union time grid; ``obs[o]`` are the observations of output
    ``o`` (sorted by time), ``masks`` the (n_out, n_t) boolean masks.
    """
    x = np.asarray(x, float)
    psi = x[:n_mech]
    sig = x[n_mech:]
    if with_grad:
        ybar, S = simulate(psi, True)
    else:
        ybar = simulate(psi, False)
    score = 0.0
    grad = np.zeros(len(x))
    start = 0
    for o, kind in enumerate(error_kinds):
        n_sig, f_ll, f_s1 = ERROR_MODELS[kind]
        end = start + n_sig
        yo = ybar[o, masks[o]]
        if with_grad:
            So = S[masks[o], o, :]
            ll, s = f_s1(sig[start:end], yo, So, obs[o])
            grad[:n_mech] += s[:n_mech]
            grad[n_mech + start:n_mech + end] += s[n_mech:]
        else:
            ll = f_ll(sig[start:end], yo, obs[o])
        score += ll
        start = end
    if with_grad:
        return score, grad
    return score


# --------------------------------------------------------------------------
# Population models.  A population model is described by a list of blocks:
#   ('pooled', n_dim)
#   ('hetero', n_dim)
#   ('gaussian', n_dim, centered)
#   ('lognormal', n_dim, centered)
#   ('truncgauss', n_dim, True)      TruncatedGaussianModel, always centered
#   ('covariate', inner_block, n_cov, pidx, didx)   inner: gaussian/lognormal
# theta layout per block follows the reference's get_parameter_names():
#   pooled: [n_dim]; hetero: [n_ids*n_dim] id-major; gaussian/lognormal:
#   [mu_1..mu_d | sigma_1..sigma_d]; covariate: [inner theta | beta
#   (n_selected, n_cov) row-major, pair order (pidx, didx)].
# --------------------------------------------------------------------------

def block_n_dim(b):
    return block_n_dim(b[1]) if b[0] == 'covariate' else b[1]


def block_n_cov(b):
    return b[2] if b[0] == 'covariate' else 0


def block_is_special(b):
    return b[0] in ('pooled', 'hetero')


def block_n_top(b, n_ids):
    """n_hierarchical_parameters()[1], chi/_population_models.py:1222-1237,
    1779-1791, 2095-2107, 2656-2670, 2987-2997."""
    kind = b[0]
    if kind == 'pooled':
        return b[1]
    if kind == 'hetero':
        return n_ids * b[1]
    if kind in ('gaussian', 'lognormal', 'truncgauss'):
        return 2 * b[1]
    if kind == 'covariate':
        return block_n_top(b[1], n_ids) + len(b[3]) * b[2]
    raise ValueError(kind)


def default_covariate_selection(n_dim, n_param_per_dim=2):
    """Index pairs chosen by CovariatePopulationModel.__init__
    (chi/_population_models.py:1024-1029) as ordered by
    LinearCovariateModel.set_population_parameters
    (chi/_covariate_models.py:337-352).  The reference sorts with two
    *non-stable* argsorts, so the order is platform dependent (SURVEY.md
    section 8a); callers that want the reference's order on this host must
    read it from the reference object instead of calling this."""
    pairs = [[p, d] for d in range(n_dim) for p in range(n_param_per_dim)]
    idx = np.array(pairs)
    idx = idx[np.argsort(idx[:, 1]), :]
    idx = idx[np.argsort(idx[:, 0]), :]
    return idx[:, 0].copy(), idx[:, 1].copy()


class PopLayout(object):
    """Bookkeeping of a composed population model for ``n_ids`` individuals
    (chi/_population_models.py:451-492 `_set_population_model_properties`,
    767-785 `n_hierarchical_parameters`)."""

    def __init__(self, blocks, n_ids):
        self.blocks = list(blocks)
        self.n_ids = int(n_ids)
        self.n_dim = sum(block_n_dim(b) for b in blocks)
        self.n_cov = sum(block_n_cov(b) for b in blocks)
        self.n_hdim = sum(
            0 if block_is_special(b) else block_n_dim(b) for b in blocks)
        self.n_top = sum(block_n_top(b, n_ids) for b in blocks)
        self.n_bottom = self.n_ids * self.n_hdim
        self.n_parameters = self.n_bottom + self.n_top


def _covariate_vartheta(b, theta, cov):
    """LinearCovariateModel.compute_population_parameters,
    chi/_covariate_models.py:257-284 -> (n_ids, 2, n_dim)."""
    inner, n_cov, pidx, didx = b[1], b[2], np.asarray(b[3]), np.asarray(b[4])
    d = inner[1]
    pop = np.asarray(theta[:2 * d], float).reshape(2, d)
    beta = np.asarray(theta[2 * d:], float).reshape(len(pidx), n_cov)
    n_ids = len(cov)
    vt = np.zeros((n_ids, 2, d)) + pop[None]
    vt[:, pidx, didx] += cov @ beta.T
    return vt


def _split(layout, theta, cov):
    """Yields (block, theta slice, dim slice, cov slice) following the
    sub-model loop of chi/_population_models.py:596-623."""
    cd = ct = cc = 0
    for b in layout.blocks:
        nd = block_n_dim(b)
        nt = block_n_top(b, layout.n_ids)
        nc = block_n_cov(b)
        c = None if cov is None or nc == 0 else cov[:, cc:cc + nc]
        yield b, theta[ct:ct + nt], slice(cd, cd + nd), c
        cd += nd
        ct += nt
        cc += nc


def pop_shape_eta(layout, bottom):
    """ComposedPopulationModel._shape_eta, chi/_population_models.py:534-565:
    (n_ids*n_hdim,) -> (n_ids, n_dim) with NaN placeholders in pooled /
    heterogeneous columns (the reference leaves them uninitialised; they are
    always overwritten by the special models)."""
    eta_h = np.asarray(bottom, float).reshape(layout.n_ids, layout.n_hdim)
    eta = np.full((layout.n_ids, layout.n_dim), np.nan)
    cd = ch = 0
    for b in layout.blocks:
        nd = block_n_dim(b)
        if not block_is_special(b):
            eta[:, cd:cd + nd] = eta_h[:, ch:ch + nd]
            ch += nd
        cd += nd
    return eta


def pop_individual_parameters(layout, theta, eta, cov=None, return_eta=False):
    """compute_individual_parameters: Composed 567-625, Pooled 2832-2837,
    Heterogeneous 1948-1951, Gaussian 1575-1599, LogNormal 2394-2443,
    Covariate 1034-1075 (all chi/_population_models.py)."""
    theta = np.asarray(theta, float)
    eta = np.asarray(eta, float)
    if eta.ndim == 1:
        eta = pop_shape_eta(layout, eta)
    psi = np.empty(eta.shape)
    for b, th, ds, c in _split(layout, theta, cov):
        kind = b[0]
        e = eta[:, ds]
        if kind == 'pooled':
            psi[:, ds] = np.broadcast_to(th.reshape(1, -1), e.shape)
        elif kind == 'hetero':
            psi[:, ds] = th.reshape(layout.n_ids, -1)
        else:
            if kind == 'covariate':
                inner = b[1]
                vt = _covariate_vartheta(b, th, c)
                mu, sigma = vt[:, 0], vt[:, 1]
            else:
                inner = b
                d = b[1]
                mu, sigma = th[None, :d], th[None, d:]
            if inner[2] or return_eta:
                psi[:, ds] = e
            elif np.any(sigma < 0):
                psi[:, ds] = np.nan
            elif inner[0] == 'gaussian':
                psi[:, ds] = mu + sigma * e
            else:
                psi[:, ds] = np.exp(mu + sigma * e)
    return psi


def _density_block(inner, mu, sigma, obs, dlp):
    """Score and unreduced sensitivities of one Gaussian / LogNormal block.

    Gaussian: chi/_population_models.py:1454-1472 (ll), 1474-1493
    (non-centered), 1513-1549 (centered grads), 1687-1757 (dispatch).
    LogNormal: 2285-2303, 2306-2320, 2322-2339, 2362-2392, 2529-2599.
    Returns (score, d_bottom (n_ids, d), d_theta (n_ids, 2, d)).
    """
    kind, d, centered = inner
    n_ids = len(obs)
    if kind == 'truncgauss':
        return _truncated_gaussian_block(d, mu, sigma, obs, dlp)
    if np.any(sigma < 0):
        return -np.inf, np.full((n_ids, d), np.nan), \
            np.full((n_ids, 2, d), np.nan)
    dth = np.empty((n_ids, 2, d))
    with np.errstate(divide='ignore', invalid='ignore'):
        if not centered:
            score = -np.sum(LOG_2PI / 2 + obs ** 2 / 2)
            if np.isnan(score):
                score = -np.inf
            z = np.zeros((1, d)) if dlp is None else dlp
            if kind == 'gaussian':
                dbot = z * sigma - obs
                dth[:, 0] = z * np.ones((n_ids, d))
                dth[:, 1] = z * obs
            else:
                psi = np.exp(mu + sigma * obs)
                dbot = z * sigma * psi - obs
                dth[:, 0] = z * psi
                dth[:, 1] = z * obs * psi
            return score, dbot, dth
        var = sigma ** 2
        if kind == 'gaussian':
            score = -np.sum(
                np.log(2 * np.pi * var) / 2 + (obs - mu) ** 2 / (2 * var))
            dbot = (mu - obs) / var * np.ones((n_ids, d))
            dth[:, 0] = (obs - mu) / var
            dth[:, 1] = (-1 + (obs - mu) ** 2 / var) / np.sqrt(var)
        else:
            lo = np.log(obs)
            score = -np.sum(
                np.log(2 * np.pi * var) / 2 + lo + (lo - mu) ** 2 / 2 / var)
            dbot = -((lo - mu) / var + 1) / obs
            dth[:, 0] = (lo - mu) / var
            dth[:, 1] = (-1 + (lo - mu) ** 2 / var) / np.sqrt(var)
        if np.isnan(score):
            score = -np.inf
        if dlp is not None:
            dbot = dbot + dlp
    return score, dbot, dth


def _norm_cdf(x):
    """Phi(x) = (1 + erf(x / sqrt 2)) / 2, chi/_population_models.py (module
    level helper used by TruncatedGaussianModel)."""
    from scipy.special import erf
    return (1 + erf(np.asarray(x, float) / np.sqrt(2))) / 2


def _norm_pdf(x):
    return np.exp(-np.asarray(x, float) ** 2 / 2) / np.sqrt(2 * np.pi)


def _truncated_gaussian_block(d, mu, sigma, obs, dlp):
    """TruncatedGaussianModel: compute_log_likelihood 3629-3660 with
    _compute_log_likelihood 3542-3571, compute_sensitivities 3703-3769 with
    _compute_sensitivities 3592-3627 (chi/_population_models.py)."""
    n_ids = len(obs)
    nan_b = np.full((n_ids, d), np.nan)
    nan_t = np.full((n_ids, 2, d), np.nan)
    if np.any(sigma <= 0) or np.any(obs < 0):
        return -np.inf, nan_b, nan_t
    with np.errstate(divide='ignore', invalid='ignore'):
        tail = 1 - _norm_cdf(-mu / sigma)
        score = -np.sum(
            np.log(2 * np.pi * sigma ** 2) / 2
            + (obs - mu) ** 2 / (2 * sigma ** 2)
            + np.log(tail) * np.ones((n_ids, d)))
        if np.isnan(score) or np.isinf(score):
            return -np.inf, nan_b, nan_t
        dbot = (mu - obs) / sigma ** 2 * np.ones((n_ids, d))
        dth = np.empty((n_ids, 2, d))
        dth[:, 0] = ((obs - mu) / sigma - _norm_pdf(mu / sigma) / tail) / sigma
        dth[:, 1] = (-1 + (obs - mu) ** 2 / sigma ** 2
                     + _norm_pdf(mu / sigma) * mu / sigma / tail) / sigma
        if dlp is not None:
            dbot = dbot + dlp
    return score, dbot, dth


def pop_log_likelihood(layout, theta, obs, cov=None):
    """compute_log_likelihood: Composed 627-666, Pooled 2866-2868,
    Heterogeneous 1986-1988, Gaussian 1601-1643, LogNormal 2445-2484,
    Covariate 1077-1102."""
    theta = np.asarray(theta, float)
    obs = np.asarray(obs, float)
    score = 0.0
    for b, th, ds, c in _split(layout, theta, cov):
        kind = b[0]
        o = obs[:, ds]
        if kind == 'pooled':
            if np.any(o != th.reshape(1, -1)):
                score += -np.inf
        elif kind == 'hetero':
            if np.any(o != th.reshape(layout.n_ids, -1)):
                score += -np.inf
        else:
            if kind == 'covariate':
                inner = b[1]
                vt = _covariate_vartheta(b, th, c)
                mu, sigma = vt[:, 0], vt[:, 1]
            else:
                inner = b
                d = b[1]
                mu, sigma = th[None, :d], th[None, d:]
            if not inner[2]:
                # non-centered: score before the sigma check
                # (chi/_population_models.py:1622-1626, 2465-2466)
                s = -np.sum(LOG_2PI / 2 + o ** 2 / 2)
                s = -np.inf if np.isnan(s) else s
            else:
                s, _, _ = _density_block(inner, mu, sigma, o, None)
            score += s
    return score


def pop_sensitivities_reduced(layout, theta, eta, dlogp_dpsi, cov=None):
    """ComposedPopulationModel._compute_reduced_sensitivities,
    chi/_population_models.py:386-449, with the per-model `_shape(reduce=
    True)` rules (56-74, 1907-1926, 2792-2810) and the covariate adjoint
    (1134-1194; chi/_covariate_models.py:286-320).

    Returns (score, gradient of length n_bottom + n_top) in the layout
    [d eta (id-major over hierarchical dims) | d theta].
    """
    theta = np.asarray(theta, float)
    eta = np.asarray(eta, float)
    n_ids = layout.n_ids
    score = 0.0
    dtop = []
    dbot = []
    for b, th, ds, c in _split(layout, theta, cov):
        kind = b[0]
        o = eta[:, ds]
        dlp = None if dlogp_dpsi is None else dlogp_dpsi[:, ds]
        if kind == 'pooled':
            if np.any(o != th.reshape(1, -1)):
                score += -np.inf
                dtop.append(np.full(b[1], np.nan))
            else:
                z = np.zeros(o.shape) if dlp is None else dlp
                dtop.append(np.sum(z, axis=0))
        elif kind == 'hetero':
            if np.any(o != th.reshape(n_ids, -1)):
                score += -np.inf
                dtop.append(np.full(n_ids * b[1], np.nan))
            else:
                z = np.zeros(o.shape) if dlp is None else dlp
                dtop.append(np.asarray(z).flatten())
        elif kind == 'covariate':
            inner = b[1]
            pidx, didx = np.asarray(b[3]), np.asarray(b[4])
            vt = _covariate_vartheta(b, th, c)
            s, db, dvt = _density_block(inner, vt[:, 0], vt[:, 1], o, dlp)
            score += s
            dbot.append(db)
            dpop = np.sum(dvt, axis=0).flatten()
            dbeta = np.sum(
                dvt[:, pidx, didx, None] * c[:, None, :], axis=0).flatten()
            dtop.append(np.concatenate((dpop, dbeta)))
        else:
            d = b[1]
            s, db, dth = _density_block(b, th[None, :d], th[None, d:], o, dlp)
            score += s
            dbot.append(db)
            dtop.append(np.sum(dth, axis=0).flatten())
    if dbot:
        dbot = np.hstack(dbot).flatten()
    else:
        dbot = np.zeros(0)
    return score, np.concatenate([dbot] + dtop)


# --------------------------------------------------------------------------
# Hierarchical log-likelihood / log-posterior
# --------------------------------------------------------------------------

def hierarchical_call(layout, individual_ll, x, cov=None):
    """HierarchicalLogLikelihood.__call__, chi/_log_pdfs.py:104-145.

    ``individual_ll(i, psi_i)`` -> float is LogLikelihood i.
    """
    x = np.asarray(x, float)
    bottom, top = x[:layout.n_bottom], x[layout.n_bottom:]
    eta = pop_individual_parameters(layout, top, bottom, cov, return_eta=True)
    score = pop_log_likelihood(layout, top, eta, cov)
    if np.isinf(score):
        return score
    psi = pop_individual_parameters(layout, top, eta, cov)
    for i in range(layout.n_ids):
        score += individual_ll(i, psi[i])
    return score


def hierarchical_s1(layout, individual_s1, x, cov=None):
    """HierarchicalLogLikelihood.evaluateS1, chi/_log_pdfs.py:255-300.

    ``individual_s1(i, psi_i)`` -> (float, grad (n_dim,)).
    """
    x = np.asarray(x, float)
    bottom, top = x[:layout.n_bottom], x[layout.n_bottom:]
    eta = pop_individual_parameters(layout, top, bottom, cov, return_eta=True)
    psi = pop_individual_parameters(layout, top, eta, cov)
    score = 0.0
    dlp = np.empty((layout.n_ids, layout.n_dim))
    for i in range(layout.n_ids):
        ll, g = individual_s1(i, psi[i])
        score += ll
        dlp[i] = g
    s, ds = pop_sensitivities_reduced(layout, top, eta, dlp, cov)
    return score + s, ds


def posterior_call(layout, individual_ll, prior, x, cov=None):
    """HierarchicalLogPosterior.__call__, chi/_log_pdfs.py:463-472.
    ``prior`` is an object with __call__/evaluateS1 on the top-level slice."""
    x = np.asarray(x, float)
    score = prior(x[layout.n_bottom:])
    if np.isinf(score):
        return score
    return score + hierarchical_call(layout, individual_ll, x, cov)


def posterior_s1(layout, individual_s1, prior, x, cov=None):
    """HierarchicalLogPosterior.evaluateS1, chi/_log_pdfs.py:474-498."""
    x = np.asarray(x, float)
    score, sens = prior.evaluateS1(x[layout.n_bottom:])
    if np.isinf(score):
        return score, np.full(len(x), np.inf)
    ll, g = hierarchical_s1(layout, individual_s1, x, cov)
    g = np.array(g, float)
    g[layout.n_bottom:] += sens
    return score + ll, g


# --------------------------------------------------------------------------
# Forward sampling (predictive models)
# --------------------------------------------------------------------------

def pop_sample(layout, theta, n_samples, rng, cov=None):
    """PopulationModel.sample: Composed 793-872, Pooled 3005-3037,
    Gaussian 1797-1846, LogNormal 2678-2732, Covariate 1247-1299
    (chi/_population_models.py).  Returns (n_samples, n_dim) draws of eta (or
    of psi for centered models), consuming ``rng`` in the reference's order.
    Heterogeneous blocks are not supported here."""
    theta = np.asarray(theta, float)
    out = np.empty((n_samples, layout.n_dim))
    cd = ct = cc = 0
    for b in layout.blocks:
        nd = block_n_dim(b)
        nt = block_n_top(b, layout.n_ids)
        nc = block_n_cov(b)
        th = theta[ct:ct + nt]
        kind = b[0]
        if kind == 'pooled':
            out[:, cd:cd + nd] = th.reshape(1, -1)
        elif kind in ('gaussian', 'lognormal'):
            mu, sigma = th[:nd], th[nd:]
            if not b[2]:
                out[:, cd:cd + nd] = rng.normal(
                    loc=np.zeros(nd), scale=np.ones(nd), size=(n_samples, nd))
            elif kind == 'gaussian':
                out[:, cd:cd + nd] = rng.normal(
                    loc=mu, scale=sigma, size=(n_samples, nd))
            else:
                out[:, cd:cd + nd] = rng.lognormal(
                    mean=mu, sigma=sigma, size=(n_samples, nd))
        elif kind == 'covariate':
            inner = b[1]
            c = np.broadcast_to(cov[:, cc:cc + nc], (n_samples, nc))
            vt = _covariate_vartheta(b, th, c)
            for i in range(n_samples):
                mu, sigma = vt[i, 0], vt[i, 1]
                if not inner[2]:
                    out[i, cd:cd + nd] = rng.normal(
                        loc=np.zeros(nd), scale=np.ones(nd), size=(1, nd))[0]
                elif inner[0] == 'gaussian':
                    out[i, cd:cd + nd] = rng.normal(
                        loc=mu, scale=sigma, size=(1, nd))[0]
                else:
                    out[i, cd:cd + nd] = rng.lognormal(
                        mean=mu, sigma=sigma, size=(1, nd))[0]
        else:
            raise NotImplementedError(kind)
        cd += nd
        ct += nt
        cc += nc
    return out


def error_sample(kind, sigma, ybar, rng):
    """ErrorModel.sample with n_samples = 1 (chi/_error_models.py:469-514,
    803-847, 1160-1206, 1514-1558) -> (n_times,)."""
    ybar = np.asarray(ybar, float)
    shape = (len(ybar), 1)
    if kind == 'gaussian':
        return (ybar[:, None] + rng.normal(0, sigma[0], size=shape))[:, 0]
    if kind == 'lognormal':
        s = sigma[0]
        return (ybar[:, None] * rng.lognormal(
            mean=-s ** 2 / 2, sigma=s, size=shape))[:, 0]
    if kind == 'cam':
        base = rng.normal(0, sigma[0], size=shape)
        rel = rng.normal(0, sigma[1], size=shape)
        return (ybar[:, None] + base + ybar[:, None] * rel)[:, 0]
    rel = rng.normal(0, sigma[0], size=shape)
    return (ybar[:, None] + ybar[:, None] * rel)[:, 0]


def population_predictive_sample(layout, simulate, n_mech, error_kinds, theta,
                                 times, n_samples, seed, cov=None):
    """PopulationPredictiveModel.sample(return_df=False),
    chi/_predictive_models.py:951-1037: sample patients, transform to psi,
    then per patient one forward solve and one noise draw per output.
    ``simulate(psi)`` -> (n_outputs, n_times).  Returns
    (n_outputs, n_times, n_samples)."""
    rng = np.random.default_rng(seed)
    lay = PopLayout(layout.blocks, n_samples)
    c = None if cov is None else np.broadcast_to(
        np.asarray(cov, float).reshape(-1, lay.n_cov), (n_samples, lay.n_cov))
    patients = pop_sample(lay, theta, n_samples, rng, c)
    patients = pop_individual_parameters(lay, theta, patients, c)
    times = np.sort(times)
    out = np.empty((len(error_kinds), len(times), n_samples))
    for p, patient in enumerate(patients):
        ybar = simulate(patient[:n_mech])
        start = 0
        for o, kind in enumerate(error_kinds):
            n_sig = ERROR_MODELS[kind][0]
            out[o, :, p] = error_sample(
                kind, patient[n_mech + start:n_mech + start + n_sig],
                ybar[o], rng)
            start += n_sig
    return out
