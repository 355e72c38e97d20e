#!/usr/bin/env python
"""Benchmark of the likelihood hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # product (CUDA)
    python bench.py --impl reference --gpus N ...          # CPU reference arm
    python bench.py --config c1|c2|c2_1000|c3|c4|c5        # another BASELINE config

Metric: log-posterior+gradient evaluations per second (one evaluation = one
parameter vector over ALL individuals).  Default workload: the north star's
target configuration -- the hierarchical two-compartment PK model of
BASELINE.json configs[1] at 10 000 individuals; one *step* evaluates a batch
of `--chains` parameter vectors (`HierarchicalLogPosterior.evaluateS1` for
each).

Multi-GPU (`torchrun`, one rank per GPU):
  individuals sharded (default for c2 / c5): the population is FIXED
      (strong scaling); rank r holds individuals [r n/N, (r+1) n/N), every rank
      evaluates all chains on its shard, and the library finishes each
      evaluation with ONE NCCL all-reduce of [score | d theta] on the device
      (chi_b200_problem_comm_init; chi/_log_pdfs.py:289-292 partitions over
      individuals the same way).
  chains sharded (default for c4 / c2_1000): every rank evaluates its own
      batch of chains on all individuals, no data-path collective (weak).

The JSON line carries `value` (inputs resident in HBM, device-timed),
`e2e` (public API with host buffers, host<->device copies and the host-side
prior inside the timed region), `roofline` (FP64: algorithmic flops of the
solve kernel / its CUDA-event time / measured DFMA peak), `cpu_baseline` and
`parity` (N = 1: the oracle's CPU restatement of the reference path on the
host cores, and the kernel's deviation from the exact ODE solution at the
bench tolerance on sampled individuals -- both outside the timed region),
`latency_ms` (single-vector evaluateS1, what pints calls) and `configs`
(the other BASELINE configs, each with its own roofline, same run).
"""
import argparse
import json
import multiprocessing
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'log-posterior+gradient evals/sec (ODE+sens)'
UNIT = 'evals/s'
# integrator tolerances of the bench (= chi_b200._synthetic.BENCH_RTOL/ATOL,
# the setting tests/test_gpu_bench_tolerance.py checks against the exact
# solution under the 1e-8 / 1e-6 contract; tests/test_host.py keeps the two
# in sync)
BENCH_RTOL = 1e-8
BENCH_ATOL = 1e-10

# hierarchical configs of BASELINE.json
HIER_CONFIGS = {
    'c2': dict(n_ids=10000, chains=1024, with_grad=True, error='lognormal',
               covariates=False, shard='individuals',
               name='C2 at the north-star size'),
    'c2_1000': dict(n_ids=1000, chains=1024, with_grad=True,
                    error='lognormal', covariates=False, shard='chains',
                    name='C2'),
    'c4': dict(n_ids=500, chains=1024, with_grad=False, error='lognormal',
               covariates=False, shard='chains', name='C4'),
    'c5': dict(n_ids=100000, chains=32, with_grad=True, error='cam',
               covariates=True, shard='individuals', name='C5'),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='chi_b200',
                    choices=['chi_b200', 'reference'])
    ap.add_argument('--config', default='c2',
                    choices=['c1', 'c2', 'c2_1000', 'c3', 'c4', 'c5'])
    ap.add_argument('--n-ids', type=int, default=None,
                    help='individuals of the whole population')
    ap.add_argument('--chains', type=int, default=None,
                    help='parameter vectors per step (and per GPU when the '
                         'chains are sharded)')
    ap.add_argument('--n-doses', type=int, default=3)
    ap.add_argument('--n-obs', type=int, default=10)
    ap.add_argument('--rtol', type=float, default=BENCH_RTOL)
    ap.add_argument('--atol', type=float, default=BENCH_ATOL)
    ap.add_argument('--shard', default=None,
                    choices=['chains', 'individuals'])
    ap.add_argument('--variant', type=int, default=-1)
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true',
                    help='skip the latency and the other-config sub-records')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--spread', type=float, default=0.1,
                    help='spread of the parameter vectors of one step around '
                         'the data-generating values '
                         '(_synthetic.spread_parameter_vectors)')
    return ap.parse_args()


def hier_config(args, name=None):
    """Settings of a hierarchical config with the command line's overrides
    (the overrides apply to the config the line is about)."""
    name = name or args.config
    cfg = dict(HIER_CONFIGS[name])
    cfg['key'] = name
    if name == args.config:
        if args.n_ids is not None:
            cfg['n_ids'] = args.n_ids
        if args.chains is not None:
            cfg['chains'] = args.chains
        if args.shard is not None:
            cfg['shard'] = args.shard
    return cfg


def workload_config(args, cfg, n_gpus):
    call = ('HierarchicalLogPosterior.evaluateS1' if cfg['with_grad'] else
            'HierarchicalLogPosterior.__call__')
    err = {'lognormal': 'LogNormalErrorModel',
           'cam': 'ConstantAndMultiplicativeGaussianErrorModel'}[cfg['error']]
    mid = ('CovariatePopulationModel(LogNormal(5, non-centered), '
           'LinearCovariateModel(n_cov=2))' if cfg['covariates']
           else 'LogNormal(5, non-centered)')
    sharded = n_gpus > 1
    return {
        'workload': (
            '%s: two-compartment PK (direct infusion into central, %d doses '
            'of U(5,15) over 1.0, period 8), %d individuals x %d obs at '
            'sorted U(0.25,24), %s, population = Composed[Pooled(2), %s, '
            'Pooled(%d)], %s batched over %d parameter vectors per step%s'
            % (cfg['name'], args.n_doses, cfg['n_ids'], args.n_obs, err, mid,
               2 if cfg['error'] == 'cam' else 1, call, cfg['chains'],
               ' and GPU' if (sharded and cfg['shard'] == 'chains') else '')),
        'config': cfg['key'], 'n_ids': cfg['n_ids'],
        'chains_per_step': cfg['chains'],
        'n_obs': args.n_obs, 'n_doses': args.n_doses,
        'rtol': args.rtol, 'atol': args.atol,
        'parameter_spread': args.spread,
        'parallelism': ('%s-sharded x%d' % (cfg['shard'], n_gpus)
                        if sharded else 'single GPU'),
        'collective': ('one NCCL all-reduce of [score | d theta] per step, '
                       'inside the library, on the device'
                       if (sharded and cfg['shard'] == 'individuals')
                       else 'none'),
        'l2': 'flushed between timed steps (256 MiB write)',
    }


# --------------------------------------------------------------------------
# flop model ("Algorithmic flops")
# --------------------------------------------------------------------------

def flops_per_step_rosl(n, n_cols, info, stages, n_init=0):
    """Algorithmic flops of one ROSL(s) step attempt of a linear model
    (integrator_stream.cuh: step_rosl): state column + n_cols integrated
    sensitivity columns, n_init of them initial-value columns (no column
    terms).  FMA = 2, every other FP64 operation 1; error norm and step-size
    controller are FP32 and not counted.  Follows the kernel's current form:
    T = (I - g J)^-1 with one reciprocal, E = T - I, the error estimate formed
    in the last stage (tests/test_host.py pins the count of the C2 model)."""
    rank1 = bool(info[3] & 2)
    f_rhs = int(info[4])
    s = stages
    inv = {1: 2, 2: 12, 3: 51}.get(n, (2 * n ** 3) // 3 + 2 * n * n)
    mv = n * (2 * n - 1)        # N x N product
    mvf = 2 * n * n             # N x N product accumulated onto a vector
    # once per attempt: g = gamma h, g J, I - g J, its inverse, E = T - I
    common = 1 + n * n + n + inv + n
    sig = n if n_cols else 0    # w_{k-1} + w_k of the state, for the columns
    # accumulation z += c_k w_k in stages 1 .. s-3; the last stage forms
    # e = c_{s-2} w_{s-2} + c_{s-1} w_{s-1} (3 N) and adds it (N)
    acc = (s - 3) * 2 * n + 4 * n
    # state column: w_0 = h f + E (h f), z = y + c_0 w_0, then w_k = E w_{k-1}
    state = f_rhs + n + mvf + 2 * n + (s - 1) * (mv + sig) + acc
    # initial-value column: w_0 = (E S) / gamma, z = S + c_0 w_0, w_k = E w
    plain = mv + n + 2 * n + (s - 1) * mv + acc
    if rank1:
        # column of a constant, D_k = d e_q^T: G = g T d (counted N),
        # w_0 = (E S + G y_q) / gamma + G w_0^y[q], z; then
        # w_k = E w_{k-1} + G (w_{k-1}^y + w_k^y)[q]
        const = mv + n + 2 * n + n + 2 * n + 2 * n + \
            (s - 1) * (n + mvf) + acc
    else:
        # dense D_k: G = g T D_k once, G y and G w_0^y in stage 0, one more
        # N x N product per later stage
        const = mv + (n * mv + n * n) + 2 * mvf + 4 * n + \
            (s - 1) * (mv + mvf) + acc
    n_const = max(n_cols - n_init, 0)
    return common + state + min(n_init, n_cols) * plain + n_const * const


def flops_per_step(n, n_cols, info, stages=6, n_init=0, method=None):
    if method is not None and method.startswith('ROSL'):
        return flops_per_step_rosl(n, n_cols, info, stages, n_init)
    """Algorithmic flops of one Rosenbrock step attempt of one system (RODAS4:
    6 stages, RODAS5: 8): the state column once + the n_cols integrated
    sensitivity columns (FMA = 2, every other FP64 operation 1)."""
    linear = bool(info[3] & 1)
    rank1 = bool(info[3] & 2)
    f_rhs, f_jac, f_col_sum, f_hess_common = [int(v) for v in info[4:8]]
    s = stages
    pairs = s * (s - 1) // 2              # (i, j<i) pairs of the stage table
    inv = {1: 2, 2: 12, 3: 51}.get(n, (2 * n ** 3) // 3 + 2 * n * n)
    comb_c = 2 * n * pairs                # sum_j c_ij U_j over all stages
    # stage values sum_j a_ij U_j: the last rows of the stiffly accurate
    # methods repeat the row before them plus one unit entry, i.e. they are
    # the previous stage value plus the previous increment (n additions)
    incr = 2 if s == 8 else 1
    pairs_a = (s - incr) * (s - incr - 1) // 2
    comb_a = 2 * n * pairs_a + incr * n
    solve = s * n * (2 * n - 1)           # s mat-vecs with W^-1
    if linear:
        # k-stage form (integrator_stream.cuh, K constants): per stage ONE
        # combination arg = z + sum_j K_ij kappa_j and
        # kappa = E arg + W^-1 g, E = W^-1 J formed once per step.  RODAS5:
        # K72 = 0 and row 8 = row 7 + two entries -> 22 of the 28 pairs
        pairs_k = 22 if s == 8 else pairs
        comb_k = 2 * n * pairs_k
        mv = n * (2 * n - 1)              # one N x N product
        # state column: kappa = E arg + W^-1 f(0) (FMAs throughout),
        # V = arg + kappa per stage; f(0), W^-1 f(0) and E once per step
        state = comb_k + s * 2 * n * n + s * n + f_rhs + mv + n * mv
        # initial-value columns: kappa = E arg; the last two V's
        plain = comb_k + s * mv + 2 * n
        # columns of constants: + W^-1 (D_k V + e_k) per stage, the sparse
        # D_k V + e_k itself is in `extra` below (the kernel multiplies
        # W^-1 D_k out once per step and executes a dense product instead:
        # slightly more than is counted here)
        sens = plain + s * (mv + n)
        if rank1:
            # every D_k has one non-zero column q: D_k V + e_k = d_k V[q] +
            # e_k, so a stage adds (W^-1 d_k) V[q] + W^-1 e_k (n FMAs), with
            # W^-1 d_k and W^-1 e_k formed once per step; nothing of the
            # column terms is left for `extra`
            sens = plain + s * 2 * n + 2 * mv
        state -= min(n_init, n_cols) * (sens - plain)   # term-free columns
    else:
        state = comb_a + comb_c + s * 2 * n + s * f_rhs + solve
        # per column: stage combinations, J Z_s, + hinv * combination, solves
        sens = comb_a + comb_c + s * 2 * n + s * n * (2 * n - 1) + solve
    # column-specific terms (df/dc_k and its y-derivative) of all columns and
    # the column-independent second-order term, per stage
    extra = 0 if (linear and rank1) else \
        s * (f_col_sum + n_cols * f_hess_common)
    jac = (0 if linear else (s + 1) * f_jac) if n_cols else (
        0 if linear else f_jac)
    # new solution = last stage value + last increment (nonlinear) / error
    # estimate = difference of the last two V's (linear); the error norm and
    # the step-size controller run in FP32 and are not counted
    new = (n_cols + 1) * n
    return inv + state + n_cols * sens + extra + jac + new


def flops_per_record(n, n_cols):
    """Output map, its directional derivatives and the error model."""
    return 3 + n_cols * (2 * n + 3) + 25 + 2 * n_cols


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------

class ClockSampler(object):
    """Samples SM clock and throttle reasons of one GPU while the timed
    region runs: NVML in a thread (a sample every few ms, so even a 100 ms
    region is covered), `nvidia-smi -lms` only when NVML is unavailable.
    `start()` is called before the warm-up so that start-up cost stays out
    of the timed region; samples count from `mark()` on."""
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,'
              'clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')
    PERIOD_S = 0.004

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.nvml = None
        self.handle = None
        self.samples = []          # (t, sm_mhz, reasons bitmask)
        self.lines = []            # (t, nvidia-smi csv line)
        self.t_mark = None
        self.smax = None
        self.running = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        handle = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.device).uuid)
            if not uuid.startswith('GPU-'):
                uuid = 'GPU-' + uuid
            handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            handle = None
        if handle is None:
            index = self.device
            visible = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            ids = [v for v in visible.split(',') if v.strip()]
            if ids and ids[min(index, len(ids) - 1)].strip().isdigit():
                index = int(ids[index])
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        return pynvml, handle

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.smax = float(self.nvml.nvmlDeviceGetMaxClockInfo(
                self.handle, self.nvml.NVML_CLOCK_SM))
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device),
                 '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                 '-lms', '20'], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        """The timed region starts now."""
        self.t_mark = time.perf_counter()

    def _poll(self):
        nv = self.nvml
        while self.running:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(
                        self.handle)
                except Exception:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(
                        self.handle)
                self.samples.append((time.perf_counter(), float(sm),
                                     int(why)))
            except Exception:
                pass
            time.sleep(self.PERIOD_S)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        t_mark = self.t_mark if self.t_mark is not None else 0.0
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=1)
            nv = self.nvml
            flags = (
                ('hw_slowdown', nv.nvmlClocksThrottleReasonHwSlowdown),
                ('hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ('sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                ('sw_power_cap', nv.nvmlClocksThrottleReasonSwPowerCap))
            inside = [s for s in self.samples if s[0] >= t_mark]
            reasons = set()
            for _, _, why in inside:
                for name, bit in flags:
                    if why & bit:
                        reasons.add(name)
            sm = [s[1] for s in inside]
            return {
                'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': self.smax, 'samples': len(sm),
                'reasons': sorted(reasons), 'source': 'nvml'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'samples': 0,
                    'reasons': [], 'source': None}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for t, line in self.lines:
            if t < t_mark:
                continue
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {
            'sm_mhz': float(np.median(sm)) if sm else None,
            'sm_max_mhz': float(np.max(smax)) if smax else None,
            'samples': len(sm), 'reasons': sorted(reasons),
            'source': 'nvidia-smi'}


# --------------------------------------------------------------------------
# CPU baseline (oracle): the reference's cost structure on the host cores
# --------------------------------------------------------------------------

class ReferencePathCPU(object):
    """chi's evaluation architecture restated on the CPU
    (oracle/stats.py + oracle/csolver): a Python loop over individuals, one
    compiled adaptive stiff solve with forward sensitivities per individual
    (-> ybar, S), numpy error model per individual, numpy population model
    (chi/_log_pdfs.py:255-300, 973-1027)."""

    def __init__(self, data, rtol, atol):
        from oracle import cport, stats
        self.stats = stats
        self.cport = cport
        self.rtol, self.atol = rtol, atol
        self.n_ids = len(data['times'])
        key = cport.find_key('two_compartment_pk_model', 'central', True)
        self.pops = []
        for i in range(self.n_ids):
            self.pops.append(cport.Population(
                key, ['central.drug_concentration'], None,
                [data['times'][i]], None, None, [data['events'][i]]))
        self.obs = data['observations']
        self.layout = stats.PopLayout(
            [('pooled', 2), ('lognormal', 5, False), ('pooled', 1)],
            self.n_ids)
        self.prior = data['prior']
        self.whole = cport.Population(
            key, ['central.drug_concentration'], ['lognormal'],
            data['times'], data['observations'], None, data['events'])

    def individual_s1(self, i, x):
        pop = self.pops[i]
        y, s, status, _, _ = pop.simulate(
            x[:7], rtol=self.rtol, atol=self.atol, stream_cb=5)
        return self.stats.lognormal_error_s1(x[7:], y, s, self.obs[i])

    def evaluateS1(self, x):
        return self.stats.posterior_s1(
            self.layout, self.individual_s1, self.prior, x)

    def openmp_systems_per_second(self, x, seconds, threads):
        """Best-effort CPU: C++ loop over individuals with OpenMP, fused
        error model, no Python in the loop."""
        lay = self.layout
        eta = self.stats.pop_individual_parameters(
            lay, x[lay.n_bottom:], x[:lay.n_bottom], None, return_eta=True)
        psi = self.stats.pop_individual_parameters(
            lay, x[lay.n_bottom:], eta, None)
        reps = 1
        t_used, n_sys = 0.0, 0
        while t_used < seconds:
            X = np.tile(psi, (reps, 1))
            ids = np.tile(np.arange(self.n_ids, dtype=np.int32), reps)
            t0 = time.perf_counter()
            self.whole.loglik(X, ids=ids, with_grad=True, rtol=self.rtol,
                              atol=self.atol, n_threads=threads, stream_cb=5)
            dt = time.perf_counter() - t0
            t_used += dt
            n_sys += len(X)
            if dt < 1.0:
                reps *= 2
        return n_sys / t_used


def _reference_worker(payload):
    data, rtol, atol, x, seconds = payload
    os.environ['OMP_NUM_THREADS'] = '1'
    ref = ReferencePathCPU(data, rtol, atol)
    for _ in range(max(1, data.get('warmup', 1))):
        ref.evaluateS1(x)  # warm-up (library load, caches)
    n = 0
    t0 = time.perf_counter()
    while True:
        ref.evaluateS1(x)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n, dt


def reference_evals_per_second(data, rtol, atol, x, seconds, procs):
    """One process per core, each evaluating the full log-posterior +
    gradient (what pints.ParallelEvaluator gives chi over chains,
    chi/_inference.py:740)."""
    ctx = multiprocessing.get_context('spawn')
    with ctx.Pool(procs) as pool:
        res = pool.map(
            _reference_worker, [(data, rtol, atol, x, seconds)] * procs)
    return sum(n / dt for n, dt in res)


TWO_CMT_TYPICAL = np.array([2.0, 1.0, 5.0, 1.0, 10.0])


def host_design(n_ids, seed, n_obs, n_doses):
    """Synthetic design in plain numpy -- the same draws as
    chi_b200._synthetic.two_compartment_design (tests/test_host_bench.py
    checks that), so the reference arm needs nothing from the product."""
    rng = np.random.default_rng(seed)
    times = [np.sort(rng.uniform(0.25, 24.0, n_obs)) for _ in range(n_ids)]
    doses = rng.uniform(5.0, 15.0, n_ids)
    events = []
    for i in range(n_ids):
        if n_doses > 1:
            events.append([(doses[i] / 1.0, 0.0, 1.0, 8.0, float(n_doses))])
        else:
            events.append([(doses[i] / 1.0, 0.0, 1.0, 0.0, 0.0)])
    return times, events


class HostPrior(object):
    """Independent Gaussian / log-normal priors on the top-level parameters
    (numpy; the counterpart of a pints.ComposedLogPrior)."""

    def __init__(self, kinds, locs, scales):
        self.kinds = np.asarray(kinds)
        self.locs = np.asarray(locs, float)
        self.scales = np.asarray(scales, float)

    def n_parameters(self):
        return len(self.kinds)

    def evaluateS1(self, x):
        x = np.asarray(x, float)
        score = 0.0
        grad = np.zeros(len(x))
        for k, (kind, m, s) in enumerate(
                zip(self.kinds, self.locs, self.scales)):
            if kind == 0:
                z = (x[k] - m) / s
                score += -0.5 * np.log(2 * np.pi) - np.log(s) - 0.5 * z * z
                grad[k] = -z / s
            else:
                if x[k] <= 0:
                    return -np.inf, np.full(len(x), np.inf)
                z = (np.log(x[k]) - m) / s
                score += -0.5 * np.log(2 * np.pi) - np.log(s) \
                    - np.log(x[k]) - 0.5 * z * z
                grad[k] = -(1 + z / s) / x[k]
        return score, grad

    def __call__(self, x):
        return self.evaluateS1(x)[0]


def c2_prior():
    mu_log = np.log(TWO_CMT_TYPICAL)
    kinds = [0, 0] + [0] * 5 + [1] * 5 + [1]
    locs = [0.0, 0.0] + list(mu_log) + [np.log(0.3)] * 5 + [np.log(0.1)]
    scales = [1.0, 1.0] + [1.0] * 5 + [0.5] * 5 + [0.5]
    return HostPrior(kinds, locs, scales)


# --------------------------------------------------------------------------
# arms
# --------------------------------------------------------------------------

def run_reference(args, rank, world):
    """CPU arm: the oracle's restatement of the reference path, all host
    cores, bounded sample (rank 0 only)."""
    if rank != 0:
        return
    if args.config not in ('c2', 'c2_1000'):
        print(json.dumps({
            'impl': 'reference',
            'unavailable': 'the CPU arm restates configs c2 / c2_1000 only'}))
        return
    from oracle import ode_exact
    cfg = hier_config(args)
    n_ids = cfg['n_ids']
    times, events = host_design(n_ids, args.seed, args.n_obs, args.n_doses)
    rng = np.random.default_rng(args.seed + 12345)
    mu_log = np.log(TWO_CMT_TYPICAL)
    sigma_log = np.full(5, 0.3)
    eta = rng.normal(size=(n_ids, 5))
    psi_mid = np.exp(mu_log[None, :] + sigma_log[None, :] * eta)
    # observations from the exact solution (the GPU is not used in this arm)
    omodel = ode_exact.two_compartment(True)
    omodel.set_outputs(['central.drug_concentration'])
    observations = []
    n_gen = min(n_ids, 64)
    for i in range(n_ids):
        j = i % n_gen
        if i < n_gen:
            psi = np.concatenate([[0.0, 0.0], psi_mid[i]])
            y = ode_exact.simulate(
                omodel, psi, times[i], events[i], sensitivities=False)[0]
            observations.append(
                y * rng.lognormal(-0.005, 0.1, size=len(y)))
        else:
            # reuse designs beyond the first 64 individuals (data values do
            # not change the cost of an evaluation)
            times[i] = times[j]
            events[i] = events[j]
            psi_mid[i] = psi_mid[j]
            eta[i] = eta[j]
            observations.append(observations[j])
    data = {'times': times, 'observations': observations, 'events': events,
            'prior': c2_prior(), 'warmup': min(args.warmup, 1)}
    x = np.concatenate(
        [eta.reshape(-1), [0.0, 0.0], mu_log, sigma_log, [0.1]])
    cores = os.cpu_count() or 1
    # a "step" of this arm is one full evaluateS1 in one process; every
    # process does an untimed evaluation and then evaluates for a bounded
    # time
    per_step = max(args.cpu_seconds, 2.0)
    value = reference_evals_per_second(
        data, args.rtol, args.atol, x, per_step, cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 / value if value > 0 else None,
        'higher_is_better': True,
        'scaling': 'strong' if cfg['shard'] == 'individuals' else 'weak',
        'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, cfg, world),
        'cpu_baseline': {
            'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': (
                'full evaluateS1 of the %d-individual model repeated for '
                '%.1f s on %d processes (one per core): Python loop '
                'over individuals, one compiled RODAS5+sensitivities solve '
                'per individual (a CPU port of the same routine the GPU '
                'kernel runs, not CVODES), numpy error and population models '
                '(oracle/stats.py, oracle/csolver)' % (
                    n_ids, per_step, cores))},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


class Context(object):
    """Per-process state shared by the measurements."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        self.args = args
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.torch = torch
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group(
                'nccl', device_id=torch.device('cuda', local_rank))
            self.dist = dist
        torch.cuda.set_device(local_rank)
        self.dev = torch.device('cuda', local_rank)
        import chi_b200
        chi_b200.set_default_device(local_rank)
        # several ranks on one host: keep each rank's page-locked buffers in
        # the memory of its GPU's NUMA node
        self.numa_node = None
        if world > 1:
            from chi_b200 import _lib
            self.numa_node = _lib.bind_to_device_numa_node(local_rank)
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32,
                                 device=self.dev)
        self.stream = torch.cuda.current_stream(self.dev)
        self._peak = None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, value):
        if self.dist is None:
            return float(value)
        t = self.torch.tensor([value], dtype=self.torch.float64,
                              device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def fp64_peak(self):
        """Measured DFMA peak (16 unrolled chains per thread, 16 warps per
        scheduler) with the SM clock sampled under it, and the nominal peak
        SMs x 64 FMA/clk x 2 x max SM clock."""
        if self._peak is not None:
            return self._peak
        from chi_b200 import _lib
        lib = _lib.lib()
        peak = np.zeros(2)
        sampler = ClockSampler(self.local_rank)
        sampler.start()
        time.sleep(0.02)
        sampler.mark()
        _lib.check(lib.chi_b200_fp64_peak(
            self.local_rank, peak[0:1].ctypes.data_as(_lib.c_f64p),
            peak[1:2].ctypes.data_as(_lib.c_f64p)))
        clocks = sampler.stop()
        sms = self.torch.cuda.get_device_properties(
            self.dev).multi_processor_count
        smax = clocks.get('sm_max_mhz') or float(peak[1])
        self._peak = {
            'peak': float(peak[0]),
            'peak_nominal': sms * 64 * 2 * smax * 1e6 * 1e-12,
            'peak_sm_mhz': clocks.get('sm_mhz'),
            'sm_max_mhz': smax, 'sms': sms}
        return self._peak


def solve_roofline(ctx, prob_model, ll_timings, with_grad, n_sys, steps_acc,
                   steps_rej, n_rec_total, dev_ms, variant, traffic=None,
                   vectors=0):
    """FP64 roofline record of the solve kernel of one workload."""
    from chi_b200 import _lib
    import ctypes
    lib = _lib.lib()
    info = np.zeros(8, dtype=np.int32)
    model_id = lib.chi_b200_model_find(prob_model.model_key().encode())
    _lib.check(lib.chi_b200_model_info(
        model_id, info.ctypes.data_as(_lib.c_i32p)))
    n_states = int(info[0])
    n_cols = n_states + int(info[1])
    inert = ctypes.c_uint32(0)
    _lib.check(lib.chi_b200_model_inert_mask(model_id, ctypes.byref(inert)))
    n_inert = bin(inert.value).count('1')
    variants = np.zeros((3, 8), dtype=np.int32)
    nv = np.zeros(1, dtype=np.int32)
    lib.chi_b200_model_variants(
        model_id, nv.ctypes.data_as(_lib.c_i32p),
        variants[0].ctypes.data_as(_lib.c_i32p),
        variants[1].ctypes.data_as(_lib.c_i32p), 8)
    vsel = variant if variant >= 0 else 0
    streamed = int(variants[0][vsel]) == 0
    code = int(variants[1][vsel]) if streamed else 1
    method = {5: 'RODAS5', 9: 'ROSL9', 11: 'ROSL11'}.get(code, 'RODAS4')
    stages = {5: 8, 9: 9, 11: 11}.get(code, 6)
    # sensitivity columns of constants that do not enter the rhs are
    # identically zero and are not integrated: they are not counted either
    n_int = (n_cols - n_inert) if with_grad else 0
    f_step = flops_per_step(n_states, n_int, info, stages,
                            n_init=n_states if with_grad else 0,
                            method=method)
    flops_launch = (steps_acc + steps_rej) * f_step + \
        n_rec_total * flops_per_record(n_states, n_cols if with_grad else 0)
    solve_ms, solve_launches = ll_timings
    solve_avg_ms = solve_ms / max(solve_launches, 1)
    achieved = flops_launch / (solve_avg_ms * 1e-3) * 1e-12
    n_out_cols = (n_cols + 1) if with_grad else 0
    alg_bytes = n_sys * 8 * (n_cols + 1 + n_out_cols + 2) + n_rec_total * 28
    desc = np.zeros(4, dtype=np.int32)
    lib.chi_b200_model_describe(
        model_id, ctx.local_rank, vsel, desc.ctypes.data_as(_lib.c_i32p))
    pk = ctx.fp64_peak()
    return {
        'bound': 'fp64', 'achieved': achieved, 'peak': pk['peak'],
        'unit': 'TFLOP/s', 'frac': achieved / pk['peak'],
        'peak_nominal': pk['peak_nominal'],
        'frac_of_nominal': achieved / pk['peak_nominal'],
        'peak_sm_mhz': pk['peak_sm_mhz'],
        # DRAM bytes per launch (read + write) of the solve kernel from the
        # committed `ncu --set full` capture of this workload, if present
        'traffic': traffic, 'algorithmic_bytes': int(alg_bytes),
        'peak_source': (
            'peak: DFMA micro-kernel measured in this run '
            '(chi_b200_fp64_peak, 16 unrolled chains per thread, SM clock '
            'under it in peak_sm_mhz; MEASURED_PEAKS.json has no FP64 '
            'figure); peak_nominal: %d SMs x 64 FMA/clk x 2 x %.0f MHz'
            % (pk['sms'], pk['sm_max_mhz'])),
        'kernel': (
            ('solve_warp_kernel<%s, %s> (persistent, thread per system, the '
             'lanes of a warp share an individual and walk through its '
             'stops together, all sensitivity columns advance together in '
             'registers, observation fused)' % (
                 prob_model.model_key(), method)
             if (method.startswith('ROSL') and vectors >= 32
                 and vectors % 32 == 0) else
             'solve_stream_kernel<%s, %s> (persistent, thread per system, '
             'per-lane event handling, observation deferred to '
             'observe_kernel)' % (prob_model.model_key(), method))
            if streamed else
            'solve_kernel<%s, G=%d, MC=%d>' % (
                prob_model.model_key(), int(variants[0][vsel]),
                int(variants[1][vsel]))),
        'kernel_ms': solve_avg_ms,
        'kernel_share_of_step': (solve_ms / dev_ms if dev_ms > 0 else None),
        'method': method, 'flops_per_step_attempt': f_step,
        'flops_per_launch': float(flops_launch),
        'sensitivity_columns_integrated': n_int,
        'steps_per_system': (steps_acc + steps_rej) / max(n_sys, 1),
        'rejected_fraction': steps_rej / max(steps_acc + steps_rej, 1),
        'registers': int(desc[0]), 'blocks_per_sm': int(desc[2]),
        'hbm_gbs_algorithmic': alg_bytes / (solve_avg_ms * 1e-3) * 1e-9,
    }


def _ncu_traffic(cfg, args):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the solve
    kernel, from profiles/solve_kernel_traffic.json (written from an
    `ncu --set full` capture of the default workload); None for any other
    workload or when the file is missing."""
    path = os.path.join(ROOT, 'profiles', 'solve_kernel_traffic.json')
    try:
        with open(path) as f:
            rec = json.load(f)
    except (OSError, ValueError):
        return None
    mine = {'n_ids': cfg['n_ids'], 'chains': cfg['chains'],
            'n_obs': args.n_obs, 'n_doses': args.n_doses,
            'rtol': args.rtol, 'atol': args.atol}
    for entry in rec.get('records', []) if isinstance(rec, dict) else []:
        if all(entry.get(k) == v for k, v in mine.items()):
            return entry.get('dram_bytes_per_launch')
    return None


def measure_hier(ctx, cfg, steps, warmup, keep=False):
    """Device-resident and end-to-end evaluations per second of one
    hierarchical config; returns the record (rank 0 fills everything)."""
    import torch
    from chi_b200 import _lib, _synthetic
    args = ctx.args
    lib = _lib.lib()
    dist, rank, world, dev = ctx.dist, ctx.rank, ctx.world, ctx.dev
    with_grad = cfg['with_grad']
    n_total = cfg['n_ids']
    shard = cfg['shard'] if world > 1 else None
    prob = _synthetic.two_compartment_problem(
        n_total, seed=args.seed, n_obs=args.n_obs, n_doses=args.n_doses,
        error=cfg['error'], covariates=cfg['covariates'], rtol=args.rtol,
        atol=args.atol)
    post = prob['log_posterior']
    ll = prob['log_likelihood']
    n_params = post.n_parameters()
    n_top = ll.n_parameters(exclude_bottom_level=True)
    n_bottom = n_params - n_top
    handle = ll._device.problem.handle
    if args.variant >= 0:
        _lib.check(lib.chi_b200_problem_set_variant(handle, args.variant))
    api = post
    n_shard = n_total
    if shard == 'individuals':
        # sets the shard and joins the library's all-reduce group
        from chi_b200._distributed import ShardedHierarchicalLogPosterior
        api = ShardedHierarchicalLogPosterior(post)
        begin, end = api.shard()
        n_shard = end - begin

    C = cfg['chains']
    X_host = _lib.pinned_empty((C, n_params))
    X_host[:] = _synthetic.spread_parameter_vectors(
        prob, C, args.spread, seed=args.seed + 100 + (
            rank if shard == 'chains' else 0))
    d_x = torch.from_numpy(np.array(X_host)).to(dev)
    d_score = torch.empty(C, dtype=torch.float64, device=dev)
    d_grad = torch.zeros((C, n_params) if with_grad else (1, 1),
                         dtype=torch.float64, device=dev)
    d_status = torch.zeros(C, dtype=torch.int32, device=dev)
    stream = ctx.stream

    def step_device():
        # with individuals sharded the call ends with the all-reduce of
        # [score | d theta] on this stream, inside the library
        ll.evaluate_device(d_x.data_ptr(), C, with_grad, d_score.data_ptr(),
                           d_grad.data_ptr(), d_status.data_ptr(),
                           stream.cuda_stream)

    # ---- device-resident timing ------------------------------------------
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    for _ in range(warmup):
        step_device()
    ctx.barrier()
    ll.set_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches = 0
    counters = np.zeros(4, dtype=np.int64)
    ctx.barrier()
    sampler.mark()
    t_wall0 = time.perf_counter()
    for k in range(steps):
        ctx.flush.zero_()                  # L2 flush, outside the event pair
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        _lib.check(lib.chi_b200_problem_counters(
            handle, counters.ctypes.data_as(_lib.c_i64p)))
        launches += int(counters[3])
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    timings = ll.timings()
    ll.set_timing(False)
    dev_ms = ctx.max_over_ranks(dev_ms)
    assert int(d_status.max().item()) == 0, 'integrator failures in bench'
    assert bool(torch.isfinite(d_score).all().item())

    evals = steps * C * (world if shard == 'chains' else 1)
    value = evals / (dev_ms * 1e-3)

    # step counters of one step (the host API call leaves them behind)
    if with_grad:
        score_h, grad_h = api.evaluateS1_batch(X_host, copy=False)
    else:
        score_h = api.evaluate_batch(X_host)
    score_h0 = float(score_h[0])
    cnt = ll.counters()
    n_sys = int(cnt[0])
    steps_acc, steps_rej = int(cnt[1]), int(cnt[2])
    assert n_sys == C * n_shard, (n_sys, C, n_shard)

    # ---- end to end through the public API ---------------------------------
    def step_api():
        if with_grad:
            s, _ = api.evaluateS1_batch(X_host, copy=False)
        else:
            s = api.evaluate_batch(X_host)
        return float(s[0])

    for _ in range(max(1, warmup)):
        step_api()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_api()
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e_value = evals / e2e_s

    copied_cols = n_params
    if shard == 'individuals':
        copied_cols = n_top + (n_bottom // n_total) * n_shard
    rec = {
        'value': value, 'ms_per_step': dev_ms / steps,
        'e2e': {
            'value': e2e_value, 'unit': UNIT,
            # with a shard set only the shard's bottom-level block and the
            # top-level entries travel (chi_b200_hier_logpdf), per rank
            'h2d_bytes_per_step': int(C * copied_cols * 8),
            'd2h_bytes_per_step': int(
                C * ((copied_cols if with_grad else 0) + 1) * 8 + C * 4),
            'ms_per_step': e2e_s * 1e3 / steps},
        'gpu_launches': launches,
        'system_solves_per_s': value * n_total,
        'wall_ms_per_step_incl_flush': t_wall * 1e3 / steps,
        'clocks': clocks,
    }
    if rank == 0:
        rec['roofline'] = solve_roofline(
            ctx, prob['model'], timings, with_grad, n_sys, steps_acc,
            steps_rej, n_sys * args.n_obs, dev_ms, args.variant,
            traffic=_ncu_traffic(cfg, args), vectors=cfg['chains'])
    if shard == 'individuals' and rank == 0:
        # the sharded, all-reduced score of vector 0 against an unsharded
        # evaluation of the same vector on this rank alone (every rank holds
        # the static data of all individuals); outside the timed region
        ref_prob = _synthetic.two_compartment_problem(
            n_total, seed=args.seed, n_obs=args.n_obs, n_doses=args.n_doses,
            error=cfg['error'], covariates=cfg['covariates'], rtol=args.rtol,
            atol=args.atol)
        full = float(ref_prob['log_posterior'].evaluate_batch(
            np.array(X_host[:1]))[0]) if not with_grad else float(
            ref_prob['log_posterior'].evaluateS1_batch(
                np.array(X_host[:1]))[0][0])
        rec['sharded_vs_unsharded_score_rel_diff'] = abs(
            score_h0 - full) / abs(full)
    if keep:
        rec['_prob'] = prob
        rec['_X'] = X_host
        rec['_score0'] = score_h0
    return rec


def parity_vs_exact(prob, X, n_ids_sample=48, n_vectors=2, seed=0):
    """Deviation of the kernel AT THE BENCH TOLERANCE from the exact ODE
    solution + numpy restatement of chi's algebra (oracle/), on sampled
    (vector, individual) pairs: relative error of the summed log-likelihood
    and of its gradient w.r.t. the individual parameters (max-norm).  The
    oracle is the checker here; outside the timed region."""
    from oracle import ode_exact, stats
    ll = prob['log_likelihood']
    n_ids = ll.n_log_likelihoods()
    rng = np.random.default_rng(seed)
    ids = np.sort(rng.choice(n_ids, size=min(n_ids_sample, n_ids),
                             replace=False))
    lay = stats.PopLayout(
        [('pooled', 2), ('lognormal', 5, False), ('pooled', 1)], n_ids)
    omodel = ode_exact.two_compartment(True)
    omodel.set_outputs(['central.drug_concentration'])
    xs, sys_ids = [], []
    for c in range(n_vectors):
        x = np.asarray(X[c])
        eta = stats.pop_individual_parameters(
            lay, x[lay.n_bottom:], x[:lay.n_bottom], None, return_eta=True)
        psi = stats.pop_individual_parameters(
            lay, x[lay.n_bottom:], eta, None)
        xs.append(psi[ids])
        sys_ids.append(ids)
    xs = np.concatenate(xs)
    sys_ids = np.concatenate(sys_ids).astype(np.int32)
    # the arithmetic of the batched launches (all columns of a system in one
    # thread: the thread-per-system kernels, bit-identical per system to
    # solve_warp_kernel), and the warp-per-system form that a launch of this
    # size would take by default
    ll.set_small_batch(0)
    g_ll, g_grad, status = ll._device.loglik(xs, ids=sys_ids, with_grad=True)
    ll.set_small_batch(-1)
    assert not np.any(status)
    w_ll, w_grad, status = ll._device.loglik(xs, ids=sys_ids, with_grad=True)
    assert not np.any(status)
    o_ll = np.empty(len(xs))
    o_grad = np.empty((len(xs), 8))
    for k, (i, psi) in enumerate(zip(sys_ids, xs)):
        ev = [tuple(r) for r in prob['regimens'][i].as_array()]
        times = prob['times'][i]

        def simulate(p, with_sens):
            return ode_exact.simulate(
                omodel, p, times, ev, sensitivities=with_sens)
        o_ll[k], o_grad[k] = stats.loglikelihood_s1(
            simulate, 7, ['lognormal'], [prob['observations'][i]],
            np.ones((1, len(times)), bool), psi, True)
    return {
        'score_rel_err_vs_exact': float(
            abs(np.sum(g_ll) - np.sum(o_ll)) / abs(np.sum(o_ll))),
        'grad_rel_err_vs_exact': float(
            np.max(np.abs(g_grad - o_grad)) / np.max(np.abs(o_grad))),
        'per_individual_ll_max_abs_err': float(np.max(np.abs(g_ll - o_ll))),
        'warp_per_system_form': {
            'score_rel_err_vs_exact': float(
                abs(np.sum(w_ll) - np.sum(o_ll)) / abs(np.sum(o_ll))),
            'grad_rel_err_vs_exact': float(
                np.max(np.abs(w_grad - o_grad)) / np.max(np.abs(o_grad)))},
        'sample': '%d individuals x %d parameter vectors at rtol %g / atol '
                  '%g against oracle/ode_exact.py + oracle/stats.py' % (
                      len(ids), n_vectors, prob['model'].tolerance()[1],
                      prob['model'].tolerance()[0]),
        'contract': 'score 1e-8 relative, gradient 1e-6 relative',
    }


def cpu_baseline(args, cfg, prob, X_host, score_h0):
    """The oracle's restatement of the reference path on the host cores
    (rank 0, N = 1), on the same model."""
    n_total = cfg['n_ids']
    n_params = X_host.shape[1]
    n_bottom = n_total * 5
    data = {
        'times': prob['times'], 'observations': prob['observations'],
        'events': [[tuple(r) for r in reg.as_array()]
                   for reg in prob['regimens']],
        'prior': prob['log_prior']}
    cores = os.cpu_count() or 1
    x = np.array(X_host[0])
    v = reference_evals_per_second(
        data, args.rtol, args.atol, x, args.cpu_seconds, cores)
    ref = ReferencePathCPU(data, args.rtol, args.atol)
    omp = ref.openmp_systems_per_second(
        x, min(5.0, args.cpu_seconds), cores) / n_total
    sc, gc = ref.evaluateS1(x)
    return {
        'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
        'sample': (
            'evaluateS1 of the same %d-individual model for %.0f s on '
            '%d processes (one per core): Python loop over individuals, '
            'one compiled solve with sensitivities per individual (a CPU '
            'port of the same RODAS5 routine the GPU kernel runs, not '
            'CVODES), numpy error/population models (oracle/stats.py + '
            'oracle/csolver)' % (n_total, args.cpu_seconds, cores)),
        # what one chain of the reference gets (one process on one core)
        'per_core_value': v / cores,
        'openmp_no_python_value': omp,
        # determinism check against the CPU build of the same routine (NOT
        # parity: see the `parity` record for the error against the exact
        # solution)
        # (score_h0 is the log-posterior of vector 0: likelihood + prior)
        'gpu_vs_cpu_port_score_rel_diff': float(abs(sc - score_h0) / abs(sc)),
    }


def measure_latency(ctx, n_ids=1000, reps=200):
    """Single-vector evaluateS1 / __call__ latency through the public API
    (what pints calls, chi/_inference.py:729-747): C2 model."""
    from chi_b200 import _synthetic
    args = ctx.args
    prob = _synthetic.two_compartment_problem(
        n_ids, seed=args.seed, n_obs=args.n_obs, n_doses=args.n_doses,
        rtol=args.rtol, atol=args.atol)
    post = prob['log_posterior']
    x = prob['x0']
    for _ in range(10):
        post.evaluateS1(x)
        post(x)
    ctx.torch.cuda.synchronize(ctx.dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        post.evaluateS1(x)
    t1 = time.perf_counter()
    for _ in range(reps):
        post(x)
    t2 = time.perf_counter()
    return {'evaluateS1': (t1 - t0) * 1e3 / reps,
            'call': (t2 - t1) * 1e3 / reps}


def measure_c1(ctx, reps=200):
    """C1: one-compartment PK, single bolus, 1 individual, Gaussian error,
    100 times: LogPosterior.evaluateS1 latency."""
    import chi_b200
    args = ctx.args
    m1 = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    m1.set_administration('central')
    m1.set_dosing_regimen(10, 0, 0.01)
    m1.set_tolerance(abs_tol=args.atol, rel_tol=args.rtol)
    t1 = np.linspace(0.1, 10, 100)
    y1 = np.abs(5 * np.exp(-t1) + 0.1 * np.random.default_rng(1).normal(
        size=100))
    ll = chi_b200.LogLikelihood(m1, chi_b200.GaussianErrorModel(), y1, t1)
    lp = chi_b200.LogPosterior(ll, chi_b200.ComposedLogPrior(
        *[chi_b200.LogNormalLogPrior(0, 1)] * 4))
    x = [0.3, 2.0, 1.0, 0.1]
    for _ in range(10):
        lp.evaluateS1(x)
    ctx.torch.cuda.synchronize(ctx.dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        s, g = lp.evaluateS1(x)
    dt = (time.perf_counter() - t0) / reps
    return {
        'workload': 'C1: one-compartment PK, bolus (0.01-wide infusion), 1 '
                    'individual, GaussianErrorModel, 100 times; '
                    'LogPosterior.evaluateS1 through the public API',
        'metric': 'latency', 'value': dt * 1e3, 'unit': 'ms',
        'evals_per_s': 1.0 / dt, 'finite': bool(np.isfinite(s))}


def measure_c3(ctx, n_samples=10000, steps=5, warmup=2):
    """C3: erlotinib TGI PKPD model, LogNormalErrorModel,
    PopulationPredictiveModel.sample for 10 k virtual patients x 31 daily
    times (chi/_predictive_models.py:951-1037), forward-only solve."""
    import chi_b200
    from chi_b200 import _lib
    args = ctx.args
    lib = _lib.lib()
    model = chi_b200.library.ModelLibrary() \
        .erlotinib_tumour_growth_inhibition_model()
    model.set_administration('central', direct=True)
    model.set_outputs(['global.tumour_volume'])
    model.set_tolerance(abs_tol=args.atol, rel_tol=args.rtol)
    pred = chi_b200.PredictiveModel(model, [chi_b200.LogNormalErrorModel()])
    pred.set_dosing_regimen(dose=5, start=0, duration=0.01, period=1, num=30)
    pm = chi_b200.ComposedPopulationModel([
        chi_b200.PooledModel(1), chi_b200.LogNormalModel(1),
        chi_b200.PooledModel(2),
        chi_b200.LogNormalModel(3, centered=False), chi_b200.PooledModel(1)])
    theta = [0.0, np.log(0.2), 0.2, 2.0, 1.0, np.log(0.5), np.log(0.3),
             np.log(0.1), 0.2, 0.3, 0.1, 0.1]
    ppm = chi_b200.PopulationPredictiveModel(pred, pm)
    times = np.arange(0.0, 31.0)

    def step():
        return ppm.sample(theta, times, n_samples=n_samples, seed=1,
                          return_df=False)
    for _ in range(warmup):
        out = step()
    mech = ppm._predictive_model._mechanistic_model
    handle = mech._get_problem().handle
    if handle is not None:
        _lib.check(lib.chi_b200_problem_set_timing(handle, 1))
    ctx.torch.cuda.synchronize(ctx.dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = step()
    dt = (time.perf_counter() - t0) / steps
    rec = {
        'workload': 'C3: erlotinib tumour-growth-inhibition PKPD model '
                    '(nonlinear), 30 daily 0.01-wide doses, '
                    'LogNormalErrorModel, PopulationPredictiveModel.sample '
                    'of %d virtual patients x 31 times (return_df=False), '
                    'end to end incl. host noise' % n_samples,
        'metric': 'virtual patients/s', 'value': n_samples / dt,
        'unit': 'patients/s', 'ms_per_step': dt * 1e3,
        'finite': bool(np.isfinite(out).all())}
    if handle is not None:
        tm = np.zeros(2)
        _lib.check(lib.chi_b200_problem_timings(
            handle, tm.ctypes.data_as(_lib.c_f64p)))
        cnt = np.zeros(4, dtype=np.int64)
        _lib.check(lib.chi_b200_problem_counters(
            handle, cnt.ctypes.data_as(_lib.c_i64p)))
        _lib.check(lib.chi_b200_problem_set_timing(handle, 0))
        rec['roofline'] = solve_roofline(
            ctx, mech, (float(tm[0]), int(tm[1])), False, int(cnt[0]),
            int(cnt[1]), int(cnt[2]), int(cnt[0]) * len(times), dt * 1e3 *
            steps, -1)
    return rec


def sub_record(ctx, name, steps, warmup):
    """One of the other hierarchical configs, measured in the same run."""
    args = ctx.args
    cfg = hier_config(args, name)
    rec = measure_hier(ctx, cfg, steps, warmup)
    out = {
        'workload': workload_config(args, cfg, ctx.world)['workload'],
        'parallelism': workload_config(args, cfg, ctx.world)['parallelism'],
        'metric': METRIC if cfg['with_grad'] else
        'log-posterior evals/sec (ODE, forward only)',
        'value': rec['value'], 'unit': UNIT,
        'ms_per_step': rec['ms_per_step'], 'e2e': rec['e2e'],
        'system_solves_per_s': rec['system_solves_per_s'],
        'scaling': 'strong' if (cfg['shard'] == 'individuals') else 'weak',
    }
    if 'roofline' in rec:
        out['roofline'] = rec['roofline']
    return out


def run_product(args, rank, world, local_rank):
    ctx = Context(args, rank, world, local_rank)
    line = None
    if args.config in HIER_CONFIGS:
        cfg = hier_config(args)
        want_cpu = (rank == 0 and world == 1 and not args.no_cpu_baseline
                    and cfg['key'] in ('c2', 'c2_1000'))
        rec = measure_hier(ctx, cfg, args.steps, args.warmup, keep=True)
        prob, X_host = rec.pop('_prob'), rec.pop('_X')
        score_h0 = rec.pop('_score0')
        if rank == 0:
            line = {
                'metric': METRIC if cfg['with_grad'] else
                'log-posterior evals/sec (ODE, forward only)',
                'value': rec['value'], 'unit': UNIT, 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': rec['ms_per_step'], 'higher_is_better': True,
                'scaling': ('strong' if cfg['shard'] == 'individuals'
                            else 'weak'),
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': workload_config(args, cfg, world)}
            for k in ('e2e', 'gpu_launches', 'system_solves_per_s',
                      'wall_ms_per_step_incl_flush', 'clocks', 'roofline',
                      'sharded_vs_unsharded_score_rel_diff'):
                if k in rec:
                    line[k] = rec[k]
        if want_cpu:
            line['parity'] = parity_vs_exact(prob, X_host)
            line['cpu_baseline'] = cpu_baseline(
                args, cfg, prob, X_host, score_h0)
        if not args.no_extras:
            # the other BASELINE configs and the single-vector latency, each
            # measured in this run (short: 3 warm-up + 5 timed steps)
            configs = {}
            if cfg['key'] != 'c2_1000':
                # the round-1 headline: C2 at 1000 individuals, chains sharded
                sub = sub_record(ctx, 'c2_1000', 5, 3)
                if rank == 0:
                    configs['c2_1000'] = sub
            if world == 1:
                for name in ('c4', 'c5'):
                    if name != cfg['key']:
                        configs[name] = sub_record(ctx, name, 5, 3)
                configs['c1'] = measure_c1(ctx)
                configs['c3'] = measure_c3(ctx)
                lat = measure_latency(ctx)
                line['latency_ms'] = {
                    'c2_1000_individuals_evaluateS1': lat['evaluateS1'],
                    'c2_1000_individuals_call': lat['call'],
                    'c1_evaluateS1': configs['c1']['value'],
                    'what': 'one parameter vector per call through the '
                            'public API (host buffers), mean of 200 calls; '
                            'launches of this size run one warp per system '
                            '(solve_warp_kernel, SolveArgs::split)'}
            if rank == 0:
                line['configs'] = configs
    elif args.config == 'c1':
        rec = measure_c1(ctx)
        line = {'metric': 'LogPosterior.evaluateS1 latency',
                'value': rec['value'], 'unit': 'ms', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': rec['value'], 'higher_is_better': False,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic', 'config': {'workload': rec['workload']}}
    elif args.config == 'c3':
        rec = measure_c3(ctx, steps=args.steps, warmup=args.warmup)
        line = {'metric': rec['metric'], 'value': rec['value'],
                'unit': rec['unit'], 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': rec['ms_per_step'],
                'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': rec['workload']},
                'roofline': rec.get('roofline')}
    if rank == 0 and line is not None:
        if world > 1 and isinstance(line.get('config'), dict):
            # NUMA node rank 0 was bound to (None: topology not exposed)
            line['config']['host_numa_node_rank0'] = ctx.numa_node
        print(json.dumps(line))
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def _only_json_on_stdout():
    """Everything that libraries print on file descriptor 1 (e.g. NCCL's
    version banner under NCCL_DEBUG) goes to stderr; `print` writes to a
    duplicate of the real stdout: the JSON line is the only line there."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, 'w', buffering=1)


def main():
    args = parse_args()
    _only_json_on_stdout()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    run_product(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
