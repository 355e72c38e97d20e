#!/usr/bin/env python
"""Benchmark of the likelihood hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # product (CUDA)
    python bench.py --impl reference --gpus N ...          # CPU reference arm

Metric: log-posterior+gradient evaluations per second of the hierarchical
two-compartment PK model (BASELINE.json configs[1]).  One *step* evaluates a
batch of `--chains` parameter vectors (`HierarchicalLogPosterior.evaluateS1`
for each) over all `--n-ids` individuals; one evaluation = one parameter
vector over all individuals.

Multi-GPU (`torchrun`, one rank per GPU):
  --shard chains        every rank evaluates its own batch of chains on all
                        individuals: no data-path collective (default)
  --shard individuals   every rank holds `--n-ids` individuals of an
                        N x n_ids population and evaluates all chains on its
                        shard; one NCCL all-reduce of [score | d theta] per
                        step (the large-population mode of the north star)

The JSON line carries `value` (inputs resident in HBM, device-timed),
`e2e` (public API with host buffers, host<->device copies and the host-side
prior inside the timed region), `roofline` (FP64: algorithmic flops of the
solve kernel / its CUDA-event time / measured DFMA peak) and `cpu_baseline`
(the oracle's CPU restatement of the reference path on the host cores).
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='chi_b200',
                    choices=['chi_b200', 'reference'])
    ap.add_argument('--n-ids', type=int, default=1000)
    ap.add_argument('--chains', type=int, default=1024,
                    help='parameter vectors per step and GPU')
    ap.add_argument('--n-doses', type=int, default=3)
    ap.add_argument('--n-obs', type=int, default=10)
    ap.add_argument('--rtol', type=float, default=1e-6)
    ap.add_argument('--atol', type=float, default=1e-8)
    ap.add_argument('--shard', default='chains',
                    choices=['chains', 'individuals'])
    ap.add_argument('--variant', type=int, default=-1)
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--spread', type=float, default=0.01,
                    help='relative spread of the parameter vectors of one '
                         'step around the data-generating values')
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        'workload': (
            'C2: two-compartment PK (direct infusion into central, %d doses '
            'of U(5,15) over 1.0, period 8), %d individuals x %d obs at '
            'sorted U(0.25,24), LogNormalErrorModel, population = '
            'Composed[Pooled(2), LogNormal(5, non-centered), Pooled(1)], '
            'HierarchicalLogPosterior.evaluateS1 batched over %d parameter '
            'vectors per step and GPU' % (
                args.n_doses, args.n_ids, args.n_obs, args.chains)),
        'n_ids': args.n_ids, 'chains_per_step_per_gpu': args.chains,
        'n_obs': args.n_obs, 'n_doses': args.n_doses,
        'rtol': args.rtol, 'atol': args.atol,
        'parameter_spread': args.spread,
        'parallelism': '%s-sharded x%d' % (args.shard, n_gpus),
        'l2': 'flushed between timed steps (256 MiB write)',
    }


# --------------------------------------------------------------------------
# flop model ("Algorithmic flops")
# --------------------------------------------------------------------------

def flops_per_step(n, n_cols, info, stages=6, n_init=0):
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


def _ncu_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the solve
    kernel, taken from profiles/solve_kernel_traffic.json (written from an
    `ncu --set full` capture of the default workload); None for any other
    workload or when the file is missing."""
    path = os.path.join(ROOT, 'profiles', 'solve_kernel_traffic.json')
    try:
        with open(path) as f:
            rec = json.load(f)
    except (OSError, ValueError):
        return None
    same = all(rec.get(k) == getattr(args, k) for k in (
        'n_ids', 'chains', 'n_obs', 'n_doses', 'rtol', 'atol'))
    return rec.get('dram_bytes_per_launch') if same else None


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
    from oracle import ode_exact
    times, events = host_design(
        args.n_ids, args.seed, args.n_obs, args.n_doses)
    rng = np.random.default_rng(args.seed + 12345)
    n_ids = args.n_ids
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
            'prior': c2_prior(), 'warmup': args.warmup}
    x = np.concatenate(
        [eta.reshape(-1), [0.0, 0.0], mu_log, sigma_log, [0.1]])
    cores = os.cpu_count() or 1
    # a "step" of this arm is one full evaluateS1 in one process; every
    # process does `warmup` untimed evaluations and then evaluates for a
    # bounded time (at least `steps` evaluations in total across processes)
    per_step = max(args.cpu_seconds, 2.0)
    value = reference_evals_per_second(
        data, args.rtol, args.atol, x, per_step, cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 / value if value > 0 else None,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, world),
        'cpu_baseline': {
            'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': (
                'full evaluateS1 of the %d-individual model repeated for '
                '%.1f s on %d processes (one per core): Python loop '
                'over individuals, one compiled RODAS5+sensitivities solve '
                'per individual, numpy error and population models '
                '(oracle/stats.py, oracle/csolver)' % (
                    n_ids, per_step, cores))},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def run_product(args, rank, world, local_rank):
    import torch
    import chi_b200
    from chi_b200 import _lib, _synthetic

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(
            'nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    chi_b200.set_default_device(local_rank)

    n_total = args.n_ids * (world if args.shard == 'individuals' else 1)
    prob = _synthetic.two_compartment_problem(
        n_total, seed=args.seed, n_obs=args.n_obs, n_doses=args.n_doses,
        error='lognormal', rtol=args.rtol, atol=args.atol)
    post = prob['log_posterior']
    ll = prob['log_likelihood']
    n_params = post.n_parameters()
    n_bottom = n_params - ll.n_parameters(exclude_bottom_level=True)
    n_top = n_params - n_bottom
    handle = ll._device.problem.handle
    lib = _lib.lib()
    if args.variant >= 0:
        _lib.check(lib.chi_b200_problem_set_variant(handle, args.variant))
    if args.shard == 'individuals' and world > 1:
        ll.set_shard(rank * args.n_ids, (rank + 1) * args.n_ids)

    C = args.chains
    rng = np.random.default_rng(args.seed + 100 + (
        rank if args.shard == 'chains' else 0))
    X_host = _lib.pinned_empty((C, n_params))
    X_host[:] = prob['x0'][None, :] * (
        1.0 + args.spread * rng.normal(size=(C, n_params)))
    d_x = torch.from_numpy(np.array(X_host)).to(dev)
    d_score = torch.empty(C, dtype=torch.float64, device=dev)
    d_grad = torch.zeros(C, n_params, dtype=torch.float64, device=dev)
    d_status = torch.zeros(C, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32,
                        device=dev)
    stream = torch.cuda.current_stream(dev)

    def step_device():
        ll.evaluate_device(d_x.data_ptr(), C, True, d_score.data_ptr(),
                           d_grad.data_ptr(), d_status.data_ptr(),
                           stream.cuda_stream)
        if dist is not None and args.shard == 'individuals':
            packed = torch.cat([d_score[:, None], d_grad[:, n_bottom:]], 1)
            dist.all_reduce(packed)
            d_score.copy_(packed[:, 0])
            d_grad[:, n_bottom:] = packed[:, 1:]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing ------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    ll.set_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    steps_acc = steps_rej = 0
    counters = np.zeros(4, dtype=np.int64)
    barrier()
    sampler.mark()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush, outside the event pair
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        _lib.check(lib.chi_b200_problem_counters(
            handle, counters.ctypes.data_as(_lib.c_i64p)))
        launches += int(counters[3])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    solve_ms, solve_launches = ll.timings()
    ll.set_timing(False)
    if dist is not None:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    assert int(d_status.max().item()) == 0, 'integrator failures in bench'
    assert bool(torch.isfinite(d_score).all().item())

    evals = args.steps * C * (world if args.shard == 'chains' else 1)
    value = evals / (dev_ms * 1e-3)
    copied_cols = n_params
    if args.shard == 'individuals' and world > 1:
        copied_cols = n_top + (n_bottom // n_total) * args.n_ids

    # step counters of the last step (host API call gives them)
    score_h, _ = ll.evaluateS1_batch(np.array(X_host), copy=False)
    score_h0 = float(score_h[0])
    cnt = ll.counters()
    n_sys = int(cnt[0])
    steps_acc, steps_rej = int(cnt[1]), int(cnt[2])

    # ---- end to end through the public API ---------------------------------
    api = post
    if dist is not None and args.shard == 'individuals':
        # shard evaluation + the all-reduce of [score | d theta] + host prior
        from chi_b200._distributed import ShardedHierarchicalLogPosterior
        api = ShardedHierarchicalLogPosterior(post, gather_bottom=False)
    for _ in range(max(1, args.warmup)):
        api.evaluateS1_batch(X_host, copy=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s, g = api.evaluateS1_batch(X_host, copy=False)
        _ = float(s[0])
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = evals / e2e_s

    # ---- roofline of the dominant kernel (solve) -----------------------------
    info = np.zeros(8, dtype=np.int32)
    model_id = lib.chi_b200_model_find(prob['model'].model_key().encode())
    _lib.check(lib.chi_b200_model_info(
        model_id, info.ctypes.data_as(_lib.c_i32p)))
    n_states = int(info[0])
    n_cols = n_states + int(info[1])
    import ctypes
    inert = ctypes.c_uint32(0)
    _lib.check(lib.chi_b200_model_inert_mask(model_id, ctypes.byref(inert)))
    n_inert = bin(inert.value).count('1')
    # sensitivity columns of constants that do not enter the rhs are
    # identically zero and are not integrated: they are not counted either
    variants = np.zeros((3, 8), dtype=np.int32)
    nv = np.zeros(1, dtype=np.int32)
    lib.chi_b200_model_variants(
        model_id, nv.ctypes.data_as(_lib.c_i32p),
        variants[0].ctypes.data_as(_lib.c_i32p),
        variants[1].ctypes.data_as(_lib.c_i32p), 8)
    vsel = args.variant if args.variant >= 0 else 0
    streamed = int(variants[0][vsel]) == 0
    method = 'RODAS5' if (streamed and int(variants[1][vsel]) == 5) \
        else 'RODAS4'
    stages = 8 if method == 'RODAS5' else 6
    f_step = flops_per_step(n_states, n_cols - n_inert, info, stages,
                            n_init=n_states)
    n_rec = args.n_obs
    flops_launch = (steps_acc + steps_rej) * f_step + \
        n_sys * n_rec * flops_per_record(n_states, n_cols)
    peak = np.zeros(2)
    _lib.check(lib.chi_b200_fp64_peak(
        local_rank, peak[0:1].ctypes.data_as(_lib.c_f64p),
        peak[1:2].ctypes.data_as(_lib.c_f64p)))
    solve_avg_ms = solve_ms / max(solve_launches, 1)
    achieved = flops_launch / (solve_avg_ms * 1e-3) * 1e-12
    alg_bytes = n_sys * 8 * (2 * (n_cols + 1) + 2) + n_sys * n_rec * 28
    desc = np.zeros(4, dtype=np.int32)
    lib.chi_b200_model_describe(
        model_id, local_rank, vsel, desc.ctypes.data_as(_lib.c_i32p))
    roofline = {
        'bound': 'fp64', 'achieved': achieved, 'peak': float(peak[0]),
        'unit': 'TFLOP/s', 'frac': achieved / float(peak[0]),
        # DRAM bytes per launch (read + write) of the solve kernel from the
        # committed `ncu --set full` capture of this workload, if present
        'traffic': _ncu_traffic(args),
        'peak_source': ('measured in this run: DFMA micro-kernel '
                        '(chi_b200_fp64_peak); MEASURED_PEAKS.json has no '
                        'FP64 figure'),
        'kernel': (
            'solve_stream_kernel<two-compartment, %s> (persistent, thread '
            'per system, sensitivity columns streamed)' % method
            if streamed else
            'solve_kernel<two-compartment, G=%d, MC=%d>' % (
                int(variants[0][vsel]), int(variants[1][vsel]))),
        'kernel_ms': solve_avg_ms, 'kernel_share_of_step':
            solve_ms / dev_ms if dev_ms > 0 else None,
        'method': method, 'flops_per_step_attempt': f_step,
        'sensitivity_columns_integrated': n_cols - n_inert,
        'steps_per_system': (steps_acc + steps_rej) / max(n_sys, 1),
        'rejected_fraction': steps_rej / max(steps_acc + steps_rej, 1),
        'registers': int(desc[0]), 'blocks_per_sm': int(desc[2]),
        'hbm_gbs_algorithmic': alg_bytes / (solve_avg_ms * 1e-3) * 1e-9,
    }

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
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
            x, min(5.0, args.cpu_seconds), cores) / args.n_ids
        sc, gc = ref.evaluateS1(x)
        cpu = {
            'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': (
                'evaluateS1 of the same %d-individual model for %.0f s on '
                '%d processes (one per core): Python loop over individuals, '
                'one compiled solve with sensitivities per individual, numpy '
                'error/population models (oracle/stats.py + oracle/csolver)'
                % (args.n_ids, args.cpu_seconds, cores)),
            # what one chain of the reference gets (one process on one core)
            'per_core_value': v / cores,
            'openmp_no_python_value': omp,
            'gpu_vs_cpu_score_rel_diff': float(
                abs(sc - (score_h0 + prob['log_prior'](
                    x[n_bottom:]))) / abs(sc)),
        }

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(args, world),
            'e2e': {
                'value': e2e_value, 'unit': UNIT,
                # with a shard set only the shard's bottom-level block and
                # the top-level entries travel (chi_b200_hier_logpdf)
                'h2d_bytes_per_step': int(C * copied_cols * 8),
                'd2h_bytes_per_step': int(C * (copied_cols + 1) * 8 + C * 4),
                'ms_per_step': e2e_s * 1e3 / args.steps},
            'gpu_launches': launches,
            'system_solves_per_s': value * n_total if args.shard == 'chains'
            else value * n_total,
            'wall_ms_per_step_incl_flush': t_wall * 1e3 / args.steps,
            'clocks': clocks, 'roofline': roofline,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    run_product(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
