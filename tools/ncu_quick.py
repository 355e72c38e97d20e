"""Quick look at an ncu report (dev helper): headline metrics, stall mix and
the opcode histogram of the kernel with the most samples.
usage: python tools/ncu_quick.py report.ncu-rep"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
def page(args):
    txt = subprocess.run(['ncu', '-i', rep] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))
rows = page(['--page', 'raw', '--csv'])
h, u = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores']
for v in rows[2:]:
    d = dict(zip(h, v)); un = dict(zip(h, u))
    print(d['Kernel Name'][:90])
    for k in keys:
        if k in d: print('   %-70s %-16s %s' % (k, un[k], d[k]))
    for k in h:
        if 'issue_stalled' in k and 'per_issue_active' in k:
            try:
                if float(d[k]) > 0.08: print('   stall %-30s %s' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), d[k]))
            except ValueError: pass
rows = page(['--page', 'source', '--csv', '--print-source', 'sass'])
st = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
for a, b in zip(st[:-1], st[1:]):
    hdr = rows[a + 1]; data = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    iS = hdr.index('Source'); iI = hdr.index('Instructions Executed'); iN = hdr.index('# Samples')
    tot = sum(int(r[iI]) for r in data); tots = sum(int(r[iN]) for r in data)
    ops = collections.Counter(); sm = collections.Counter()
    for r in data:
        s = r[iS].strip()
        if s.startswith('@'): s = s.split(' ', 1)[1].strip()
        op = s.split(' ')[0].split('.')[0]
        ops[op] += int(r[iI]); sm[op] += int(r[iN])
    print('total warp instr', tot, 'samples', tots, 'static', len(data))
    print('  '.join('%s %.1f%%' % (op, 100 * c / tot) for op, c in ops.most_common(24)))
