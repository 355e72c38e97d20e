"""Basic-block level summary of an ncu report's SASS page (dev helper).

usage: python tools/ncu_blocks.py report.ncu-rep [min_percent]
"""
import csv, subprocess, sys, io
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# a report may hold several kernels: take the one with the most samples
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
best = None
for a, b in zip(starts[:-1], starts[1:]):
    hdr = rows[a + 1]
    seg = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    n = sum(int(r[hdr.index('# Samples')]) for r in seg)
    if best is None or n > best[0]:
        best = (n, rows[a][1], hdr, seg)
print(best[1][:100])
hdr, data = best[2], best[3]
iA = hdr.index('Source'); iI = hdr.index('Instructions Executed')
iT = hdr.index('Avg. Threads Executed'); iS = hdr.index('# Samples')
tots = sum(int(r[iS]) for r in data); tot = sum(int(r[iI]) for r in data)
print('instructions', tot, 'samples', tots, 'static', len(data))
for lo, hi in [(0, 4), (4, 12), (12, 24), (24, 29), (29, 33)]:
    ii = sum(int(r[iI]) for r in data if lo <= float(r[iT]) < hi)
    ss = sum(int(r[iS]) for r in data if lo <= float(r[iT]) < hi)
    print('threads [%d,%d): inst %.1f%% samples %.1f%%' % (lo, hi, 100 * ii / tot, 100 * ss / tots))
blk = []; cur = None
for k, r in enumerate(data):
    ie = int(r[iI])
    if cur is None or cur['ie'] != ie:
        cur = {'ie': ie, 'k0': k, 'k1': k, 's': 0, 't': float(r[iT]), 'ops': []}; blk.append(cur)
    cur['k1'] = k; cur['s'] += int(r[iS])
    w = r[iA].split()
    cur['ops'].append(w[1] if w[0].startswith('@') else w[0])
for b in blk:
    n = b['k1'] - b['k0'] + 1
    sh = 100 * b['s'] / tots; ih = 100 * b['ie'] * n / tot
    if sh > thr or ih > thr:
        print('%4d-%4d n=%3d exec=%9d thr=%4.1f inst%%=%5.2f samp%%=%5.2f  %s' % (
            b['k0'], b['k1'], n, b['ie'], b['t'], ih, sh, ' '.join(b['ops'][:8])))
