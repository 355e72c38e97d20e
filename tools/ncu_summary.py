"""Writes profiles/<name>_ncu_summary.md and profiles/solve_kernel_traffic.json
from an `ncu --set full` report and an `ncu --metrics gpu__time_duration.sum`
launch list of `bench.py --steps 2 --warmup 1 --no-cpu-baseline` (dev helper).

usage: python tools/ncu_summary.py <report.ncu-rep> <launches.csv> <name> "<title>" "<bench args>" <plain.json>
"""
import collections, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, launches, name, title = sys.argv[1:5]
bench_args = sys.argv[5] if len(sys.argv) > 5 else ''
plain = json.load(open(sys.argv[6])) if len(sys.argv) > 6 else None
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, u = rows[0], rows[1]
KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
    'sm__cycles_elapsed.max', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
    'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores']
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out, notes, traffic = [], [], None
for v in rows[2:]:
    d, un = dict(zip(h, v)), dict(zip(h, u))
    out.append('### `%s`\n\n| metric | unit | value |\n|---|---|---|' % d['Kernel Name'][:90])
    out += ['| %s | %s | %s |' % (k, un[k], d[k]) for k in KEYS if k in d]
    out.append('')
    rd = float(d['dram__bytes_read.sum']) * SCALE[un['dram__bytes_read.sum']]
    wr = float(d['dram__bytes_write.sum']) * SCALE[un['dram__bytes_write.sum']]
    ms = float(d['gpu__time_duration.sum']) * {'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[un['gpu__time_duration.sum']]
    if 'solve_stream' in d['Kernel Name'] or 'solve_warp' in d['Kernel Name']:
        kname = 'solve_warp_kernel' if 'solve_warp' in d['Kernel Name'] else 'solve_stream_kernel'
        traffic = rd + wr
        cyc = float(d['sm__cycles_elapsed.max'])
        fl = (2 * float(d['smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed'])
              + float(d['smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed'])
              + float(d['smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed'])) * cyc
        notes.append('* `%s`: executed FP64 flops (2 DFMA + DADD + DMUL, predicated-on thread '
                     'instructions) %.3e -> %.2f TFLOP/s issued at %.2f ms; FP64 pipe active %s %%; DRAM %.0f MB '
                     'read + %.0f MB written = %.0f GB/s.' % (
                         kname, fl, fl / ms / 1e9, ms,
                         d['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'][:5],
                         rd / 1e6, wr / 1e6, (rd + wr) / ms / 1e6))
    else:
        notes.append('* `%s`: %.2f ms, DRAM %.0f MB read + %.0f MB written = %.2f TB/s.' % (
            d['Kernel Name'].split('<')[0].split()[-1], ms, rd / 1e6, wr / 1e6, (rd + wr) / ms / 1e9))
lst = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(lst) if r and r[0] == 'ID'][0]
hdr = lst[hi]
body = [r for r in lst[hi + 1:] if len(r) == len(hdr)]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in body:
    f = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[mu], 1e-6)
    a = agg.setdefault(r[kn], [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(',', '')) * f
tot = sum(a[1] for a in agg.values())
tab = ['| kernel | launches | total ms | share |', '|---|---|---|---|']
tab += ['| `%s` | %d | %.3f | %.1f %% |' % (k[:90], a[0], a[1], 100 * a[1] / tot)
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]) if a[1] / tot > 2e-4]
dst_csv = os.path.join(ROOT, 'profiles', name + '_launches_ncu.csv')
if os.path.abspath(launches) != dst_csv:
    open(dst_csv, 'w').write(open(launches).read())
cfg = plain['config'] if plain else {}
roof = plain['roofline'] if plain else {}
md = '''# %s

Command: `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras %s` (one B200; %s);
the plain run exited 0 first (bench line of that run: value %s evals/s, `roofline.frac` %s, `kernel_share_of_step` %s).

## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, profiles/%s_launches_ncu.csv)

%s

`fp64_peak_kernel` is the DFMA micro-benchmark that measures the roofline denominator (outside the timed region);
the `at::` kernels are bench.py's own L2 flush and result checks.  Inside a step the solve kernel
(`solve_warp_kernel` / `solve_stream_kernel`: integration with the observation fused) is what `bench.py` times as "the solve" (`roofline.kernel_ms`, CUDA
events); the population kernels are about 1 %% of a step.  Launch times under ncu are cold-cache and
serialised: compare SHARES.

## `ncu --set full --clock-control none --import-source on -k regex:"solve_stream|observe_kernel" -s 2 -c 2`

%s
Reading:
%s
''' % (title, bench_args, cfg.get('workload', '')[:160], 
       ('%.0f' % plain['value']) if plain else '-', ('%.3f' % roof.get('frac', 0)) if plain else '-',
       ('%.3f' % roof.get('kernel_share_of_step', 0)) if plain else '-',
       name, '\n'.join(tab), '\n'.join(out), '\n'.join(notes))
open(os.path.join(ROOT, 'profiles', name + '_ncu_summary.md'), 'w').write(md)
path = os.path.join(ROOT, 'profiles', 'solve_kernel_traffic.json')
try:
    records = json.load(open(path)).get('records', [])
except (OSError, ValueError, AttributeError):
    records = []
if plain and traffic is not None:
    rec = {'n_ids': cfg['n_ids'], 'chains': cfg['chains_per_step'], 'n_obs': cfg['n_obs'],
           'n_doses': cfg['n_doses'], 'rtol': cfg['rtol'], 'atol': cfg['atol'],
           'dram_bytes_per_launch': traffic,
           'source': 'ncu --set full capture of the solve kernel, profiles/%s_ncu_summary.md '
                     '(dram__bytes_read.sum + dram__bytes_write.sum)' % name}
    records = [r for r in records if not all(r.get(k) == rec[k] for k in (
        'n_ids', 'chains', 'n_obs', 'n_doses', 'rtol', 'atol'))] + [rec]
    json.dump({'records': records}, open(path, 'w'), indent=1)
print('\n'.join(notes))
