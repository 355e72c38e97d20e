"""ctypes wrapper of the host port of the solve (TEST INFRASTRUCTURE ONLY).

See oracle/csolver/port.cpp.  Used by bench.py's CPU baseline and by the
CPU-side tests of the integrator's step control.
"""
import ctypes
import json
import os

import numpy as np

from oracle import ode_exact

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_OUT = 8
MAX_COLS = 32

ERR_KINDS = {'gaussian': 0, 'lognormal': 1, 'cam': 2, 'multiplicative': 3}
N_SIGMA = {0: 1, 1: 1, 2: 2, 3: 1}

_vp = ctypes.c_void_p


class SolveArgs(ctypes.Structure):
    """Mirror of chi_b200::SolveArgs (chi_b200/csrc/model_api.cuh)."""
    _fields_ = [
        ('B', ctypes.c_int64), ('n_ids', ctypes.c_int32), ('ids', _vp),
        ('x', _vp), ('x_stride', ctypes.c_int32),
        ('rec_off', _vp), ('rec_t', _vp), ('rec_obs', _vp),
        ('rec_logobs', _vp), ('rec_out', _vp),
        ('seg_off', _vp), ('seg_t', _vp), ('seg_rate', _vp),
        ('n_out', ctypes.c_int32),
        ('out_id', ctypes.c_int32 * MAX_OUT),
        ('err_kind', ctypes.c_int32 * MAX_OUT),
        ('err_off', ctypes.c_int32 * MAX_OUT),
        ('n_sigma', ctypes.c_int32),
        ('n_cols', ctypes.c_int32),
        ('col_param', ctypes.c_int32 * MAX_COLS),
        ('param_col', ctypes.c_int32 * MAX_COLS),
        ('n_int', ctypes.c_int32),
        ('int_col', ctypes.c_int32 * MAX_COLS),
        ('int_param', ctypes.c_int32 * MAX_COLS),
        ('mode', ctypes.c_int32), ('with_sens', ctypes.c_int32),
        ('rtol', ctypes.c_double), ('atol', ctypes.c_double),
        ('max_steps', ctypes.c_int32),
        ('traj_off', _vp), ('traj_y', _vp), ('traj_s', _vp),
        ('ll', _vp), ('grad', _vp), ('grad_stride', ctypes.c_int32),
        ('status', _vp), ('n_steps', _vp), ('n_rej', _vp),
        ('snap', _vp), ('snap_rows', ctypes.c_int32),
        ('snap_width', ctypes.c_int32),
        ('order_inner', ctypes.c_int32), ('order_stride', ctypes.c_int32),
        ('refill_wait', ctypes.c_int32), ('work_counter', _vp),
        ('split', ctypes.c_int32),
        ('perm', _vp), ('order_keys', _vp),
    ]


_lib = None


def _configure(port):
    port.chi_port_solve.argtypes = [
        ctypes.c_char_p, ctypes.POINTER(SolveArgs), ctypes.c_int]
    port.chi_port_fill_integrated.argtypes = [
        ctypes.c_char_p, ctypes.POINTER(SolveArgs)]
    port.chi_port_snap_width.argtypes = [
        ctypes.c_char_p, ctypes.POINTER(SolveArgs)]
    port.chi_port_solve2.argtypes = [
        ctypes.c_char_p, ctypes.POINTER(SolveArgs), ctypes.c_int,
        ctypes.c_int]
    assert port.chi_port_sizeof_args() == ctypes.sizeof(SolveArgs), (
        port.chi_port_sizeof_args(), ctypes.sizeof(SolveArgs))
    return port


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, 'csolver', 'libchi_port.so')
        if not os.path.exists(path):
            from oracle.csolver import build
            build.build()
        _lib = _configure(ctypes.CDLL(path))
    return _lib


def open_port(path):
    """A port library built by oracle.csolver.build.build_models (models that
    are not built in); pass it to Population(..., port=...)."""
    return _configure(ctypes.CDLL(path))


def max_threads():
    return lib().chi_port_max_threads()


def model_index():
    path = os.path.join(
        os.path.dirname(HERE), 'chi_b200', 'csrc', 'gen', 'index.json')
    with open(path) as f:
        return json.load(f)


def find_key(model, compartment=None, direct=True):
    for key, info in model_index().items():
        adm = info['administration']
        if info['model'] != model:
            continue
        if compartment is None and adm is None:
            return key
        if adm is not None and compartment == adm['compartment'] and \
                bool(direct) == adm['direct']:
            return key
    raise KeyError((model, compartment, direct))


def _ptr(a):
    return None if a is None else a.ctypes.data


class Population(object):
    """Static data of a population in the layout the solve expects."""

    def __init__(self, key, outputs, error_kinds, times, observations,
                 out_index, regimens, info=None, port=None, columns=None):
        """times/observations/out_index: per-individual arrays of records
        (sorted by time); regimens: per-individual list of events.  `info`
        (states, constants, outputs) and `port` (open_port) for a model that
        is not built in.  `columns`: mechanistic parameters whose
        sensitivities are computed (default: all, in order)."""
        info = info if info is not None else model_index()[key]
        self.columns = None if columns is None else [int(c) for c in columns]
        self._port = port
        self.key = key
        self.info = info
        self.P = len(info['states']) + len(info['constants'])
        self.out_id = [info['outputs'].index(o) for o in outputs]
        self.err_kind = [ERR_KINDS[k] for k in (error_kinds or [])]
        self.n_ids = len(times)
        off = [0]
        for t in times:
            off.append(off[-1] + len(t))
        self.rec_off = np.array(off, np.int64)
        cat = lambda xs, dt: (np.concatenate([np.asarray(x, dt) for x in xs])
                              if len(xs) and off[-1] > 0
                              else np.zeros(0, dt))
        self.rec_t = np.ascontiguousarray(cat(times, float))
        if observations is None:
            self.rec_obs = np.zeros(off[-1])
        else:
            self.rec_obs = np.ascontiguousarray(cat(observations, float))
        with np.errstate(all='ignore'):
            self.rec_logobs = np.log(self.rec_obs)
        if out_index is None:
            self.rec_out = np.zeros(off[-1], np.int32)
        else:
            self.rec_out = np.ascontiguousarray(cat(out_index, np.int32))
        seg_off = [0]
        seg_t, seg_rate = [], []
        for i in range(self.n_ids):
            t_end = float(times[i][-1]) if len(times[i]) else 0.0
            ev = regimens[i] if regimens is not None else []
            edges, rates = ode_exact.expand_regimen(ev, max(t_end, 0.0))
            k = max(len(rates), 1)
            seg_t += list(edges[:k])
            seg_rate += list(rates) if len(rates) else [0.0]
            seg_off.append(seg_off[-1] + k)
        self.seg_off = np.array(seg_off, np.int64)
        self.seg_t = np.array(seg_t, float)
        self.seg_rate = np.array(seg_rate, float)

    def args(self):
        a = SolveArgs()
        a.n_ids = self.n_ids
        a.rec_off = _ptr(self.rec_off)
        a.rec_t = _ptr(self.rec_t)
        a.rec_obs = _ptr(self.rec_obs)
        a.rec_logobs = _ptr(self.rec_logobs)
        a.rec_out = _ptr(self.rec_out)
        a.seg_off = _ptr(self.seg_off)
        a.seg_t = _ptr(self.seg_t)
        a.seg_rate = _ptr(self.seg_rate)
        a.n_out = len(self.out_id)
        off = 0
        for o, oid in enumerate(self.out_id):
            a.out_id[o] = oid
            if self.err_kind:
                a.err_kind[o] = self.err_kind[o]
                a.err_off[o] = off
                off += N_SIGMA[self.err_kind[o]]
        a.n_sigma = off
        cols = self.columns if self.columns is not None else range(self.P)
        a.n_cols = len(cols)
        for k in range(MAX_COLS):
            a.param_col[k] = -1
        for k, pk in enumerate(cols):
            a.col_param[k] = pk
            a.param_col[pk] = k
        (self._port or lib()).chi_port_fill_integrated(self.key.encode(), ctypes.byref(a))
        return a

    def loglik(self, x, ids=None, with_grad=True, rtol=1e-8, atol=1e-10,
               max_steps=100000, n_threads=1, stream_cb=0, deferred=None,
               buffered=False, split=False):
        """``deferred``: evaluate the observations from snapshots after the
        solve (SolveArgs::snap); False keeps them fused in the solve.
        ``buffered`` (fused only): records go through the one-record buffer,
        as in a GPU launch whose warps share an individual.  ``split``: one
        work item per (system, sensitivity column), as in a small GPU
        launch."""
        x = np.ascontiguousarray(x, float)
        if x.ndim == 1:
            x = x[None, :]
        B = x.shape[0]
        a = self.args()
        nd = self.P + a.n_sigma
        assert x.shape[1] == nd, (x.shape, nd)
        ids_a = None if ids is None else np.ascontiguousarray(ids, np.int32)
        ll = np.zeros(B)
        grad = np.zeros((B, nd))
        status = np.zeros(B, np.int32)
        nst = np.zeros(B, np.int32)
        nrj = np.zeros(B, np.int32)
        a.B = B
        a.ids = _ptr(ids_a)
        a.x = _ptr(x)
        a.x_stride = nd
        a.mode = 1
        a.with_sens = 1 if with_grad else 0
        a.rtol, a.atol, a.max_steps = rtol, atol, max_steps
        a.ll = _ptr(ll)
        a.grad = _ptr(grad)
        a.grad_stride = nd
        a.status = _ptr(status)
        a.n_steps = _ptr(nst)
        a.n_rej = _ptr(nrj)
        if deferred is None:
            deferred = stream_cb != 0 and not buffered
        if buffered:
            a.refill_wait = 2 ** 31 - 1
        if split:
            # one work item per (system, integrated column), SolveArgs::split
            a.split = a.n_int
        snap = None
        if deferred and stream_cb != 0:
            width = (self._port or lib()).chi_port_snap_width(
                self.key.encode(), ctypes.byref(a))
            rows = int(np.max(np.diff(self.rec_off))) if self.n_ids else 0
            snap = np.full(B * max(rows, 1) * width, np.nan)
            a.snap = _ptr(snap)
            a.snap_rows = rows
            a.snap_width = width
        rc = (self._port or lib()).chi_port_solve2(self.key.encode(), ctypes.byref(a),
                                   n_threads, stream_cb)
        assert rc == 0, rc
        return ll, grad, status, nst, nrj

    def simulate(self, psi, ids=None, with_sens=True, rtol=1e-8, atol=1e-10,
                 max_steps=100000, n_threads=1, stream_cb=0):
        psi = np.ascontiguousarray(psi, float)
        if psi.ndim == 1:
            psi = psi[None, :]
        B = psi.shape[0]
        a = self.args()
        ids_a = None if ids is None else np.ascontiguousarray(ids, np.int32)
        idv = np.arange(B) % self.n_ids if ids is None else np.asarray(ids)
        counts = self.rec_off[idv + 1] - self.rec_off[idv]
        toff = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        rows = int(toff[-1])
        y = np.zeros(rows)
        s = np.zeros((rows, a.n_cols))
        status = np.zeros(B, np.int32)
        nst = np.zeros(B, np.int32)
        nrj = np.zeros(B, np.int32)
        a.B = B
        a.ids = _ptr(ids_a)
        a.x = _ptr(psi)
        a.x_stride = self.P
        a.mode = 0
        a.with_sens = 1 if with_sens else 0
        a.rtol, a.atol, a.max_steps = rtol, atol, max_steps
        a.traj_off = _ptr(toff)
        a.traj_y = _ptr(y)
        a.traj_s = _ptr(s)
        a.status = _ptr(status)
        a.n_steps = _ptr(nst)
        a.n_rej = _ptr(nrj)
        rc = (self._port or lib()).chi_port_solve2(self.key.encode(), ctypes.byref(a),
                                   n_threads, stream_cb)
        assert rc == 0, rc
        return y, s, status, nst, nrj
