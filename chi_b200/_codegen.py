"""Symbolic model -> CUDA device functions (RHS, Jacobian, sensitivity terms).

This is the B200 counterpart of what myokit does when chi constructs a
``myokit.Simulation(model, protocol, sensitivities)``
(chi/_mechanistic_models.py:146, 307-308): myokit generates C for the RHS and
lets CVODES difference-quotient the sensitivity right-hand sides; here sympy
derives the exact Jacobian ``J = df/dy``, the parameter Jacobian ``df/dc`` and
the second-order term ``d/dy [J s + (df/dc) e] v`` that a Rosenbrock method
needs for the forward-sensitivity system, and emits them as ``__device__``
functions for the integrator template in ``csrc/integrator.cuh``.

State order is chi's parameter order: ``sorted(states)`` then
``sorted(constants)`` (chi/_mechanistic_models.py:186-205).
"""
import hashlib

import sympy as sp
from sympy.printing.c import C99CodePrinter


class _Printer(C99CodePrinter):
    def _print_Pow(self, expr):
        base, exp = expr.as_base_exp()
        if exp.is_Integer and 1 < abs(int(exp)) <= 4:
            b = self.parenthesize(base, 1000)
            prod = '*'.join([b] * abs(int(exp)))
            if int(exp) < 0:
                return '(1.0/(%s))' % prod
            return '(%s)' % prod
        if exp == -1:
            return '(1.0/%s)' % self.parenthesize(base, 1000)
        return super(_Printer, self)._print_Pow(expr)

    def _print_rcp(self, expr):
        return '(1.0/%s)' % self.parenthesize(expr.args[0], 1000)

    def _print_Rational(self, expr):
        return '(%d.0/%d.0)' % (expr.p, expr.q)

    def _print_Integer(self, expr):
        return '%d.0' % int(expr)


_printer = _Printer()


class rcp(sp.Function):
    """Reciprocal kept as an opaque function so that it can be shared."""
    nargs = 1


_rcp = rcp


def _share_reciprocals(expr):
    """x**-k -> rcp(x)**k, so that the common-subexpression pass emits one
    division per distinct denominator and multiplications elsewhere."""
    expr = sp.sympify(expr)
    return expr.replace(
        lambda e: e.is_Pow and e.exp.is_Integer and e.exp.is_negative,
        lambda e: _rcp(e.base) ** (-e.exp))


def _count_flops(exprs):
    n = 0
    for e in exprs:
        n += int(sp.count_ops(e, visual=False))
    return n


def _emit_block(assignments, indent='    '):
    """assignments: list of (lhs string, sympy expr); applies CSE."""
    exprs = [_share_reciprocals(e) for _, e in assignments]
    repl, reduced = sp.cse(exprs, symbols=sp.numbered_symbols('t_'),
                           optimizations='basic')
    lines = []
    for sym, e in repl:
        lines.append('%sconst double %s = %s;' % (
            indent, sym, _printer.doprint(e)))
    for (lhs, _), e in zip(assignments, reduced):
        lines.append('%s%s = %s;' % (indent, lhs, _printer.doprint(e)))
    flops = _count_flops([e for _, e in repl] + list(reduced))
    return '\n'.join(lines), flops


class ModelCode(object):
    """Everything the host needs to know about a generated model."""

    def __init__(self, model):
        self.model = model
        self.state_names = sorted(model.states)
        self.const_names = sorted(model.constants)
        self.n = len(self.state_names)
        self.nc = len(self.const_names)
        self.output_names = list(self.state_names) + sorted(
            model.intermediates)
        self._derive()

    def _derive(self):
        m = self.model
        y = [m.symbol(s) for s in self.state_names]
        c = [m.symbol(s) for s in self.const_names]
        u = m.dose_rate
        self._y, self._c, self._u = y, c, u
        f = sp.Matrix([m.expanded(m.rhs[s]) for s in self.state_names])
        f = f.applyfunc(sp.cancel)
        self.f = f
        self.J = f.jacobian(y) if self.n else sp.zeros(0, 0)
        self.J = self.J.applyfunc(sp.cancel)
        self.fc = f.jacobian(c) if self.nc else sp.zeros(self.n, 0)
        self.fc = self.fc.applyfunc(sp.cancel)
        self.linear = not any(self.J.has(s) for s in y)
        # constants that do not enter the rhs at all (e.g. compartment sizes
        # of mass-action PK models): dy/dc_k == 0 identically, only the
        # output map depends on them
        self.inert = [all(self.fc[i, k] == 0 for i in range(self.n))
                      for k in range(self.nc)]
        allowed = set(y) | set(c) | ({u} if u is not None else set())
        stray = f.free_symbols - allowed
        if stray:
            raise ValueError(
                'The model rhs depends on undefined symbols: %s' % stray)
        self.g = [sp.cancel(m.expression(o)) for o in self.output_names]

        canon = repr((
            self.state_names, self.const_names, self.output_names,
            [sp.srepr(e) for e in f], [sp.srepr(e) for e in self.g],
            None if u is None else str(u), m.dosed_state))
        self.key = hashlib.sha1(canon.encode()).hexdigest()[:12]

    # ------------------------------------------------------------------
    def header(self):
        """Returns the text of the generated model header."""
        n, nc = self.n, self.nc
        y, c, u = self._y, self._c, self._u
        ysym = {s: sp.Symbol('y[%d]' % i) for i, s in enumerate(y)}
        csym = {s: sp.Symbol('c[%d]' % i) for i, s in enumerate(c)}
        sub = dict(ysym)
        sub.update(csym)
        if u is not None:
            sub[u] = sp.Symbol('u')
        s_vec = [sp.Symbol('s[%d]' % i) for i in range(n)]
        e_vec = [sp.Symbol('e[%d]' % i) for i in range(nc)]
        v_vec = [sp.Symbol('v[%d]' % i) for i in range(n)]

        f = self.f.xreplace(sub)
        J = self.J.xreplace(sub)
        fc = self.fc.xreplace(sub)
        # the Jacobian must not depend on the dose rate (additive input)
        if u is not None and J.has(sp.Symbol('u')):
            raise NotImplementedError('Dose rate enters non-additively.')

        rhs_code, fl_rhs = _emit_block(
            [('f[%d]' % i, f[i]) for i in range(n)])
        jac_code, fl_jac = _emit_block(
            [('J[%d]' % (i * n + j), J[i, j])
             for i in range(n) for j in range(n)])
        fce = [sum((fc[i, k] * e_vec[k] for k in range(nc)), sp.Integer(0))
               for i in range(n)]
        dfdc_code, fl_dfdc = _emit_block(
            [('r[%d]' % i, fce[i]) for i in range(n)])
        # sens rhs direction: J s + fc e ; its y-derivative applied to v
        sens = [sum((J[i, j] * s_vec[j] for j in range(n)), sp.Integer(0))
                + fce[i] for i in range(n)]
        ylist = [ysym[s] for s in y]
        hess = [sum((sp.diff(sens[i], ylist[j]) * v_vec[j]
                     for j in range(n)), sp.Integer(0)) for i in range(n)]
        hess = [sp.expand(h) if self.linear else h for h in hess]
        hess_code, fl_hess = _emit_block(
            [('r[%d]' % i, hess[i]) for i in range(n)])

        # column-specialised sensitivity terms (uniform switch on the constant
        # index): r += (df/dc_k)(y)   and   r += d/dy[J(y) s + (df/dc_k)(y)] v
        common = [sum((sp.diff(sum((J[i, j] * s_vec[j] for j in range(n)),
                                   sp.Integer(0)), ylist[l]) * v_vec[l]
                       for l in range(n)), sp.Integer(0)) for i in range(n)]
        common = [sp.expand(h) if self.linear else h for h in common]

        def _acc_block(exprs, indent):
            items = [('r[%d]' % i, sp.Symbol('r[%d]' % i) + e)
                     for i, e in enumerate(exprs) if e != 0]
            if not items:
                return '', 0
            code, fl = _emit_block(items, indent=indent)
            return code, fl

        dfdc_cases, hess_cases = [], []
        self.flops_col = []       # per constant: ops of its extra terms
        for k in range(nc):
            e1 = [fc[i, k] for i in range(n)]
            e2 = [sum((sp.diff(fc[i, k], ylist[l]) * v_vec[l]
                       for l in range(n)), sp.Integer(0)) for i in range(n)]
            c1, f1 = _acc_block(e1, '        ')
            c2, f2 = _acc_block(e2, '        ')
            if c1:
                dfdc_cases.append(
                    '      case %d: {\n%s\n        break; }' % (k, c1))
            if c2:
                hess_cases.append(
                    '      case %d: {\n%s\n        break; }' % (k, c2))
            self.flops_col.append(f1 + f2)
        common_code, self.flops_hess_common = _acc_block(common, '    ')

        # linear models: (df/dc_k)(y) = D_k y + e_k with D_k, e_k free of y,
        # so the column terms of a stage are one small matrix-vector product
        # with data fetched once per column (no switch inside the stages)
        lin_cases = []
        # ... and when every D_k has at most one non-zero column q_k (a
        # constant that multiplies one state, the mass-action case),
        # (df/dc_k)(y) = d_k y[q_k] + e_k: a scalar times a vector per stage
        r1_cases = []
        col_rank1 = bool(self.linear)
        if self.linear:
            for k in range(nc):
                col = sp.Matrix([fc[i, k] for i in range(n)])
                Dk = col.jacobian(ylist) if n else sp.zeros(0, 0)
                ek = (col - Dk * sp.Matrix(ylist)).applyfunc(sp.expand)
                if any(Dk.has(v) for v in ylist) or any(
                        ek.has(v) for v in ylist):
                    raise AssertionError('model is not linear in the states')
                items = [('D[%d]' % (i * n + j), Dk[i, j])
                         for i in range(n) for j in range(n) if Dk[i, j] != 0]
                items += [('e[%d]' % i, ek[i]) for i in range(n)
                          if ek[i] != 0]
                if items:
                    code, _ = _emit_block(items, indent='        ')
                    lin_cases.append(
                        '      case %d: {\n%s\n        break; }' % (k, code))
                nzcols = [j for j in range(n)
                          if any(Dk[i, j] != 0 for i in range(n))]
                if len(nzcols) > 1:
                    col_rank1 = False
                elif col_rank1:
                    q = nzcols[0] if nzcols else 0
                    items = [('d[%d]' % i, Dk[i, q]) for i in range(n)
                             if nzcols and Dk[i, q] != 0]
                    items += [('e[%d]' % i, ek[i]) for i in range(n)
                              if ek[i] != 0]
                    if items:
                        code, _ = _emit_block(items, indent='        ')
                        r1_cases.append(
                            '      case %d: {\n%s\n        q = %d;\n'
                            '        break; }' % (k, code, q))
        if not col_rank1:
            r1_cases = []

        out_cases = []
        dout_cases = []
        grad_cases = []
        for k, g in enumerate(self.g):
            gk = g.xreplace(sub)
            items = [('r', gk)]
            items += [('gy[%d]' % j, sp.diff(gk, ylist[j]))
                      for j in range(n)]
            items += [('gc[%d]' % j, sp.diff(gk, csym[c[j]]))
                      for j in range(nc)]
            code, _ = _emit_block(items, indent='        ')
            grad_cases.append('      case %d: {\n%s\n        break; }' % (
                k, code))
            code, _ = _emit_block([('r', gk)], indent='        ')
            out_cases.append('      case %d: {\n%s\n        break; }' % (
                k, code))
            dg = sum((sp.diff(gk, ylist[j]) * s_vec[j] for j in range(n)),
                     sp.Integer(0)) + \
                sum((sp.diff(gk, csym[c[j]]) * e_vec[j] for j in range(nc)),
                    sp.Integer(0))
            code, _ = _emit_block([('r', dg)], indent='        ')
            dout_cases.append('      case %d: {\n%s\n        break; }' % (
                k, code))

        ncs = max(nc, 1)
        name = 'Model_' + self.key
        lines = []
        a = lines.append
        a('// GENERATED by chi_b200/_codegen.py -- do not edit.')
        a('// model key: %s' % self.key)
        a('// states   : %s' % ', '.join(self.state_names))
        a('// constants: %s' % ', '.join(self.const_names))
        a('// outputs  : %s' % ', '.join(self.output_names))
        a('#pragma once')
        a('#include "model_api.cuh"')
        a('struct %s {' % name)
        a('  static constexpr int N = %d;' % n)
        a('  static constexpr int NC = %d;' % nc)
        a('  static constexpr int NCS = %d;' % ncs)
        a('  static constexpr int P = N + NC;')
        a('  static constexpr int NOUT = %d;' % len(self.g))
        a('  static constexpr bool LINEAR = %s;' % (
            'true' if self.linear else 'false'))
        a('  // bit k: constant k does not enter the rhs (dy/dc_k == 0)')
        a('  static constexpr unsigned INERT_MASK = 0x%xu;' % sum(
            (1 << k) for k, flag in enumerate(self.inert) if flag))
        a('  // resident blocks per SM the streamed kernel is compiled for')
        a('  static constexpr int STREAM_MIN_BLOCKS = %d;' % (
            int(__import__('os').environ.get('CHI_B200_MINB', '0')) or (
                4 if (self.linear and n <= 2) else 2)))
        a('  static constexpr int DOSED_STATE = %d;' % (
            self.state_names.index(self.model.dosed_state)
            if self.model.dosed_state else -1))
        a('  static constexpr int FLOPS_RHS = %d;' % fl_rhs)
        a('  static constexpr int FLOPS_JAC = %d;' % fl_jac)
        a('  static constexpr int FLOPS_DFDC = %d;' % fl_dfdc)
        a('  static constexpr int FLOPS_HESS = %d;' % fl_hess)
        a('  // ops of the column-specialised terms, summed over the constants'
          ' that enter the rhs')
        a('  static constexpr int FLOPS_COL_SUM = %d;' % sum(
            f for f, flag in zip(self.flops_col, self.inert) if not flag))
        a('  static constexpr int FLOPS_HESS_COMMON = %d;'
          % self.flops_hess_common)
        a('  static const char* key() { return "%s"; }' % self.key)
        a('  CHI_HD static void rhs(const double* y, const double* c, '
          'double u, double* f) {')
        a('    (void)y; (void)c; (void)u;')
        a(rhs_code)
        a('  }')
        a('  CHI_HD static void jac(const double* y, const double* c, '
          'double* J) {')
        a('    (void)y; (void)c;')
        a(jac_code)
        a('  }')
        a('  // r = (df/dc) e')
        a('  CHI_HD static void dfdc(const double* y, const double* c, '
          'const double* e, double* r) {')
        a('    (void)y; (void)c; (void)e;')
        a(dfdc_code)
        a('  }')
        a('  // r = d/dy [ J(y) s + (df/dc)(y) e ] v')
        a('  CHI_HD static void hess(const double* y, const double* c, '
          'const double* s, const double* e, const double* v, double* r) {')
        a('    (void)y; (void)c; (void)s; (void)e; (void)v;')
        a(hess_code)
        a('  }')
        a('  // Column-specific sensitivity terms; k = index of the constant '
          '(uniform')
        a('  // over the grid), k < 0: initial-value column.  '
          'r += (df/dc_k)(y)')
        a('  CHI_HD static void dfdc_col_add(int k, const double* y, '
          'const double* c, double* r) {')
        a('    (void)y; (void)c; (void)r;')
        a('    switch (k) {')
        lines.extend(dfdc_cases)
        a('      default: break;')
        a('    }')
        a('  }')
        a('  // r += d/dy [ J(y) s + (df/dc_k)(y) ] v')
        a('  CHI_HD static void hess_col_add(int k, const double* y, '
          'const double* c, const double* s, const double* v, double* r) {')
        a('    (void)y; (void)c; (void)s; (void)v; (void)r;')
        a(common_code)
        a('    switch (k) {')
        lines.extend(hess_cases)
        a('      default: break;')
        a('    }')
        a('  }')
        a('  // linear models: (df/dc_k)(y) = D y + e   (k < 0: D = 0, e = 0)')
        a('  CHI_HD static void col_linear(int k, const double* c, double u, '
          'double* D, double* e) {')
        a('    (void)c; (void)u;')
        a('    for (int j = 0; j < N * N; ++j) D[j] = 0.0;')
        a('    for (int j = 0; j < N; ++j) e[j] = 0.0;')
        a('    switch (k) {')
        lines.extend(lin_cases)
        a('      default: break;')
        a('    }')
        a('  }')
        a('  // linear models whose D_k all have at most one non-zero column:')
        a('  // (df/dc_k)(y) = d y[q] + e, returns q')
        a('  static constexpr bool COL_RANK1 = %s;' % (
            'true' if col_rank1 else 'false'))
        a('  CHI_HD static int col_rank1(int k, const double* c, double u, '
          'double* d, double* e) {')
        a('    (void)c; (void)u;')
        a('    for (int j = 0; j < N; ++j) d[j] = 0.0;')
        a('    for (int j = 0; j < N; ++j) e[j] = 0.0;')
        a('    int q = 0;')
        a('    switch (k) {')
        lines.extend(r1_cases)
        a('      default: break;')
        a('    }')
        a('    return q;')
        a('  }')
        a('  CHI_HD static double out(int id, const double* y, '
          'const double* c) {')
        a('    (void)y; (void)c;')
        a('    double r = 0.0;')
        a('    switch (id) {')
        lines.extend(out_cases)
        a('      default: break;')
        a('    }')
        a('    return r;')
        a('  }')
        a('  // r = (dg/dy) s + (dg/dc) e')
        a('  CHI_HD static double dout(int id, const double* y, '
          'const double* c, const double* s, const double* e) {')
        a('    (void)y; (void)c; (void)s; (void)e;')
        a('    double r = 0.0;')
        a('    switch (id) {')
        lines.extend(dout_cases)
        a('      default: break;')
        a('    }')
        a('    return r;')
        a('  }')
        a('  // returns g(y, c); gy = dg/dy, gc = dg/dc')
        a('  CHI_HD static double out_grad(int id, const double* y, '
          'const double* c, double* gy, double* gc) {')
        a('    (void)y; (void)c;')
        a('    double r = 0.0;')
        a('    for (int j = 0; j < N; ++j) gy[j] = 0.0;')
        a('    for (int j = 0; j < NCS; ++j) gc[j] = 0.0;')
        a('    switch (id) {')
        lines.extend(grad_cases)
        a('      default: break;')
        a('    }')
        a('    return r;')
        a('  }')
        a('};')
        return '\n'.join(lines) + '\n'

    def variants(self):
        """(lanes per system, method code) kernel variants, most preferred
        first.  lanes == 0: the streamed-columns kernels
        (integrator_stream.cuh); the second entry is then the method: 1 =
        RODAS4, 5 = RODAS5, 9 / 11 = ROSL(9) / ROSL(11), the Rosenbrock
        methods for linear models.  (The lane-parallel RODAS4 kernels of
        round 1 -- several lanes per system, columns split over them -- were
        slower than the streamed kernels at every batch size and are no longer
        built; small batches run the warp-per-system form of the ROSL
        kernels.)"""
        return [(0, 11), (0, 9), (0, 5), (0, 1)] if self.linear \
            else [(0, 5), (0, 1)]

    def translation_unit(self):
        """Text of the .cu file that instantiates the kernels for this
        model and registers them with the core library."""
        name = 'Model_' + self.key
        vs = ', '.join('chi_b200::Variant<%d, %d>' % v
                       for v in self.variants())
        return (
            '// GENERATED by chi_b200/_codegen.py -- do not edit.\n'
            '#include <initializer_list>\n'
            '#include "model_%s.cuh"\n'
            '#include "integrator.cuh"\n'
            'CHI_B200_REGISTER_MODEL(%s, %s)\n' % (self.key, name, vs))
