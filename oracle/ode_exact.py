"""Converged solutions of chi's mechanistic models (TEST INFRASTRUCTURE ONLY).

What this restates
------------------
``chi.SBMLModel.simulate`` (chi/_mechanistic_models.py:490-533) hands the
model to ``myokit.Simulation.run`` (myokit >= 1.34, setup.py:33; generated C
linked against SUNDIALS CVODES, apt.txt:1).  Neither is vendored or installed
and the reference's tests assert no trajectory or sensitivity *values*
(SURVEY.md section 4), so this file is **PARITY UNPINNED** with respect to the
reference's numbers; it is pinned against mathematics instead: it returns the
*exact* solution of the ODE + forward-sensitivity system that CVODES
approximates, which is what any correct integrator converges to.

* Models are typed in by hand below from the SBML files (cited per model) plus
  chi's structural edits for dosing (chi/_mechanistic_models.py:563-636) --
  on purpose NOT through ``chi_b200``'s SBML reader, so the two derivations
  check one another.
* Linear models (all PK models): the state + sensitivity system is linear time
  invariant between dose edges, so it is advanced with one matrix exponential
  of the augmented matrix per interval (scipy.linalg.expm, ~1e-15).
* Nonlinear models (tumour growth inhibition): scipy DOP853 at rtol 1e-13 on
  the augmented system, restarted at every dose edge.
* Dosing: ``set_dosing_regimen`` semantics (chi/_mechanistic_models.py:
  795-857): rate = dose/duration, pulse train (start, duration, period, num),
  num=0 -> indefinitely, period=0 -> single pulse; the "bolus" is a 0.01 wide
  infusion.

Parameter order is chi's: sorted(state names) + sorted(constant names)
(chi/_mechanistic_models.py:186-205); sensitivities are w.r.t. the initial
values for the first n entries (``init(state)``, :280-288).
"""
import numpy as np
import scipy.integrate
import scipy.linalg
import sympy as sp


class OdeModel(object):
    """Symbolic model: states, constants, rhs, outputs (all hand typed)."""

    def __init__(self, states, consts, rhs, outputs, dosed_state):
        # names
        self.state_names = list(states)
        self.const_names = list(consts)
        self.n = len(states)
        self.parameter_names = sorted(states) + sorted(consts)
        self.P = len(self.parameter_names)
        self.output_exprs = dict(outputs)
        self.dosed_state = dosed_state

        self._y = [sp.Symbol('y%d' % i) for i in range(self.n)]
        self._c = {c: sp.Symbol('c_' + c.replace('.', '__')) for c in consts}
        self._u = sp.Symbol('u')
        ns = dict(zip(states, self._y))
        ns.update(self._c)
        ns['u'] = self._u
        self._ns = ns
        f = [sp.sympify(rhs[s], locals=self._local(ns)) for s in states]
        f[self.state_names.index(dosed_state)] += self._u
        self._f = sp.Matrix(f)
        self._J = self._f.jacobian(self._y)
        cs = [self._c[c] for c in sorted(consts)]
        self._cs = cs
        self._fp = self._f.jacobian(cs)
        self.linear = not any(self._J.has(y) for y in self._y)
        args = self._y + cs + [self._u]
        self._f_num = sp.lambdify(args, self._f, 'numpy')
        self._J_num = sp.lambdify(args, self._J, 'numpy')
        self._fp_num = sp.lambdify(args, self._fp, 'numpy')
        # d(J v + fp_k)/dy pieces for nonlinear models
        if not self.linear:
            v = [sp.Symbol('v%d' % i) for i in range(self.n)]
            self._v = v
            Jv = self._J * sp.Matrix(v)
            self._dJv_num = sp.lambdify(
                args + v, Jv.jacobian(self._y), 'numpy')
            self._dfp_num = [
                sp.lambdify(args, self._fp[:, k].jacobian(self._y), 'numpy')
                for k in range(len(cs))]
        # position of the sorted states inside the declared state order
        self._state_perm = [
            self.state_names.index(s) for s in sorted(states)]

    @staticmethod
    def _local(ns):
        return {k.replace('.', '__'): v for k, v in ns.items()}

    def set_outputs(self, names):
        self.outputs = list(names)
        args = self._y + self._cs
        g = sp.Matrix([
            sp.sympify(self.output_exprs[o], locals=self._local(self._ns))
            for o in names])
        self._g_num = sp.lambdify(args, g, 'numpy')
        self._gy_num = sp.lambdify(args, g.jacobian(self._y), 'numpy')
        self._gp_num = sp.lambdify(args, g.jacobian(self._cs), 'numpy')


def _expr(s):
    return s.replace('.', '__')


def one_compartment(direct=True):
    """chi/library/model_library/pk_one_comp.xml:24-52 (+ dosing edits)."""
    states = ['central.drug_amount']
    consts = ['central.size', 'global.elimination_rate']
    # kinetic law: central * elimination_rate * drug, drug = amount / size
    rhs = {'central.drug_amount': _expr(
        '-central.size * global.elimination_rate '
        '* (central.drug_amount / central.size)')}
    outputs = {
        'central.drug_amount': _expr('central.drug_amount'),
        'central.drug_concentration': _expr(
            'central.drug_amount / central.size')}
    dosed = 'central.drug_amount'
    if not direct:
        states.append('dose.drug_amount')
        consts.append('dose.absorption_rate')
        rhs['central.drug_amount'] += _expr(
            ' + dose.absorption_rate * dose.drug_amount')
        rhs['dose.drug_amount'] = _expr(
            '-dose.absorption_rate * dose.drug_amount')
        outputs['dose.drug_amount'] = _expr('dose.drug_amount')
        dosed = 'dose.drug_amount'
    m = OdeModel(states, consts, rhs, outputs, dosed)
    m.set_outputs(['central.drug_concentration'])
    return m


def two_compartment(direct=True):
    """docs/source/getting_started/code/two_compartment_pk_model.xml:24-93."""
    states = ['central.drug_amount', 'peripheral.drug_amount']
    consts = ['central.size', 'peripheral.size', 'global.elimination_rate',
              'global.transition_rate_c2p', 'global.transition_rate_p2c']
    cc = '(central.drug_amount / central.size)'
    cp = '(peripheral.drug_amount / peripheral.size)'
    r_el = 'central.size * global.elimination_rate * ' + cc
    r_cp = 'central.size * global.transition_rate_c2p * ' + cc
    r_pc = 'peripheral.size * global.transition_rate_p2c * ' + cp
    rhs = {
        'central.drug_amount': _expr('-%s - %s + %s' % (r_el, r_cp, r_pc)),
        'peripheral.drug_amount': _expr('%s - %s' % (r_cp, r_pc))}
    outputs = {
        'central.drug_amount': _expr('central.drug_amount'),
        'peripheral.drug_amount': _expr('peripheral.drug_amount'),
        'central.drug_concentration': _expr(cc),
        'peripheral.drug_concentration': _expr(cp)}
    dosed = 'central.drug_amount'
    if not direct:
        states.append('dose.drug_amount')
        consts.append('dose.absorption_rate')
        rhs['central.drug_amount'] += _expr(
            ' + dose.absorption_rate * dose.drug_amount')
        rhs['dose.drug_amount'] = _expr(
            '-dose.absorption_rate * dose.drug_amount')
        outputs['dose.drug_amount'] = _expr('dose.drug_amount')
        dosed = 'dose.drug_amount'
    m = OdeModel(states, consts, rhs, outputs, dosed)
    m.set_outputs(['central.drug_concentration'])
    return m


def tumour_growth_inhibition_pkpd(direct=True):
    """chi/library/model_library/temporary_full_pkpd_model.xml:41-106
    (`ModelLibrary.erlotinib_tumour_growth_inhibition_model`)."""
    states = ['central.drug_amount', 'global.tumour_volume']
    consts = ['central.size', 'global.critical_volume',
              'global.elimination_rate', 'global.kappa', 'global.lambda']
    cc = '(central.drug_amount / central.size)'
    rhs = {
        'central.drug_amount': _expr(
            '-central.size * global.elimination_rate * ' + cc),
        'global.tumour_volume': _expr(
            'global.lambda * global.critical_volume * global.tumour_volume'
            ' / (global.tumour_volume + global.critical_volume)'
            ' - global.kappa * %s * global.tumour_volume' % cc)}
    outputs = {
        'central.drug_amount': _expr('central.drug_amount'),
        'central.drug_concentration': _expr(cc),
        'global.tumour_volume': _expr('global.tumour_volume')}
    dosed = 'central.drug_amount'
    if not direct:
        states.append('dose.drug_amount')
        consts.append('dose.absorption_rate')
        rhs['central.drug_amount'] += _expr(
            ' + dose.absorption_rate * dose.drug_amount')
        rhs['dose.drug_amount'] = _expr(
            '-dose.absorption_rate * dose.drug_amount')
        outputs['dose.drug_amount'] = _expr('dose.drug_amount')
        dosed = 'dose.drug_amount'
    m = OdeModel(states, consts, rhs, outputs, dosed)
    m.set_outputs(['global.tumour_volume'])
    return m


def tumour_growth_koch(reparametrised=False):
    """chi/library/model_library/tgi_Koch_2009.xml:40-86 and
    tgi_Koch_2009_reparametrised.xml (no dosing: plain SBMLModel)."""
    states = ['global.tumour_volume']
    if reparametrised:
        consts = ['global.critical_volume', 'global.drug_concentration',
                  'global.kappa', 'global.lambda']
        growth = ('global.lambda * global.critical_volume * '
                  'global.tumour_volume / (global.tumour_volume + '
                  'global.critical_volume)')
    else:
        consts = ['global.drug_concentration', 'global.kappa',
                  'global.lambda_0', 'global.lambda_1']
        growth = ('2 * global.lambda_0 * global.lambda_1 * '
                  'global.tumour_volume / (2 * global.lambda_0 * '
                  'global.tumour_volume + global.lambda_1)')
    rhs = {'global.tumour_volume': _expr(
        growth + ' - global.kappa * global.drug_concentration * '
        'global.tumour_volume')}
    outputs = {'global.tumour_volume': _expr('global.tumour_volume')}
    m = OdeModel(states, consts, rhs, outputs, 'global.tumour_volume')
    m.set_outputs(['global.tumour_volume'])
    return m


# --------------------------------------------------------------------------
# Dosing
# --------------------------------------------------------------------------

def expand_regimen(events, t_end):
    """Piecewise-constant dose rate on [0, t_end].

    ``events`` is a list of (level, start, duration, period, multiplier) as in
    ``myokit.pacing.blocktrain`` / ``ProtocolEvent`` (call sites
    chi/_mechanistic_models.py:853-855, chi/_problems.py:360-372).  Returns
    (edges, rates): rate ``rates[k]`` applies on [edges[k], edges[k+1]).
    """
    pulses = []
    for level, start, duration, period, mult in events:
        if period == 0:
            pulses.append((start, start + duration, level))
            continue
        k = 0
        while True:
            t_on = start + k * period
            if t_on >= t_end or (mult > 0 and k >= mult):
                break
            pulses.append((t_on, t_on + duration, level))
            k += 1
    edges = {0.0, float(t_end)}
    for a, b, _ in pulses:
        if a < t_end:
            edges.add(float(max(a, 0.0)))
        if b < t_end:
            edges.add(float(max(b, 0.0)))
    edges = sorted(edges)
    rates = []
    for k in range(len(edges) - 1):
        mid = 0.5 * (edges[k] + edges[k + 1])
        rates.append(sum(lv for a, b, lv in pulses if a <= mid < b))
    return np.array(edges), np.array(rates)


# --------------------------------------------------------------------------
# Solve
# --------------------------------------------------------------------------

def simulate(model, parameters, times, regimen=None, sensitivities=True):
    """Exact counterpart of SBMLModel.simulate
    (chi/_mechanistic_models.py:490-533).

    Returns ybar (n_out, n_t) and, if requested, S (n_t, n_out, P).
    """
    parameters = np.asarray(parameters, float)
    times = np.asarray(times, float)
    n, P = model.n, model.P
    n_t = len(times)
    n_out = len(model.outputs)
    if n_t == 0:
        if sensitivities:
            return np.zeros((n_out, 0)), np.zeros((0, n_out, P))
        return np.zeros((n_out, 0))
    t_end = float(times[-1])
    edges, rates = expand_regimen(regimen or [], max(t_end, 0.0))

    # initial values arrive in sorted-state order
    y0 = np.empty(n)
    for k, pos in enumerate(model._state_perm):
        y0[pos] = parameters[k]
    c = parameters[n:]
    S0 = np.zeros((n, P))
    for k, pos in enumerate(model._state_perm):
        S0[pos, k] = 1.0

    Y = np.empty((n_t, n))
    SS = np.empty((n_t, n, P))
    y, S = y0.copy(), S0.copy()
    t = 0.0
    it = 0
    while it < n_t and times[it] <= 0.0:
        Y[it], SS[it] = y, S
        it += 1
    for k in range(len(edges) - 1):
        a, b, u = edges[k], edges[k + 1], rates[k]
        last = k == len(edges) - 2
        # output times inside (a, b]
        ts = []
        while it + len(ts) < n_t and (
                times[it + len(ts)] <= b):
            ts.append(times[it + len(ts)])
        stops = list(ts)
        if not stops or stops[-1] < b:
            stops.append(b)
        if model.linear:
            sol = _advance_linear(model, y, S, c, u, a, stops)
        else:
            sol = _advance_nonlinear(model, y, S, c, u, a, stops)
        for j in range(len(ts)):
            Y[it + j], SS[it + j] = sol[j]
        it += len(ts)
        y, S = sol[-1]
        t = b
        if it >= n_t and last:
            break
    assert it == n_t, (it, n_t)

    # outputs g(y, c) and chain rule
    out = np.empty((n_out, n_t))
    sens = np.empty((n_t, n_out, P))
    for j in range(n_t):
        args = list(Y[j]) + list(c)
        out[:, j] = np.asarray(model._g_num(*args), float).reshape(n_out)
        if sensitivities:
            gy = np.asarray(model._gy_num(*args), float).reshape(n_out, n)
            gp = np.asarray(model._gp_num(*args), float).reshape(n_out, P - n)
            sens[j] = gy @ SS[j]
            sens[j][:, n:] += gp
    if sensitivities:
        return out, sens
    return out


def _advance_linear(model, y, S, c, u, t0, stops):
    n, P = model.n, model.P
    zero = [0.0] * n
    A = np.asarray(model._J_num(*(zero + list(c) + [u])), float).reshape(n, n)
    b = np.asarray(model._f_num(*(zero + list(c) + [u])), float).reshape(n)
    # dA/dc_k via the symbolic Jacobian (entries are polynomial/rational in c)
    N = n * (P + 1) + 1
    M = np.zeros((N, N))
    M[:n, :n] = A
    M[:n, -1] = b
    for k in range(P):
        r = n * (k + 1)
        M[r:r + n, r:r + n] = A
    dA = _dA_dc(model, c, u)
    for kc in range(P - n):
        r = n * (n + kc + 1)
        M[r:r + n, :n] = dA[kc]
    z = np.concatenate([y, S.T.reshape(-1), [1.0]])
    out = []
    t = t0
    for ts in stops:
        z = scipy.linalg.expm(M * (ts - t)) @ z
        t = ts
        out.append((z[:n].copy(), z[n:-1].reshape(P, n).T.copy()))
    return out


def _dA_dc(model, c, u):
    if not hasattr(model, '_dA_num'):
        model._dA_num = [
            sp.lambdify(model._cs + [model._u],
                        model._J.diff(ck), 'numpy') for ck in model._cs]
    n = model.n
    return [np.asarray(f(*(list(c) + [u])), float).reshape(n, n)
            for f in model._dA_num]


def _advance_nonlinear(model, y, S, c, u, t0, stops):
    n, P = model.n, model.P
    cl = list(c)

    def rhs(t, z):
        yy = list(z[:n])
        SS = z[n:].reshape(P, n).T
        args = yy + cl + [u]
        f = np.asarray(model._f_num(*args), float).reshape(n)
        J = np.asarray(model._J_num(*args), float).reshape(n, n)
        fp = np.asarray(model._fp_num(*args), float).reshape(n, P - n)
        dS = J @ SS
        dS[:, n:] += fp
        return np.concatenate([f, dS.T.reshape(-1)])

    z0 = np.concatenate([y, S.T.reshape(-1)])
    sol = scipy.integrate.solve_ivp(
        rhs, (t0, stops[-1]), z0, method='DOP853', t_eval=stops,
        rtol=1e-13, atol=1e-15)
    if not sol.success:
        raise RuntimeError(sol.message)
    out = []
    for j in range(len(stops)):
        z = sol.y[:, j]
        out.append((z[:n].copy(), z[n:].reshape(P, n).T.copy()))
    return out
