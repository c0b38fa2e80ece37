#!/usr/bin/env python3
"""Join an ncu SASS source page (csv) with nvdisasm line info to get per-source-line
instruction counts and stall samples.  usage: ncu_lines.py <rep> <cubin> <kernel-substring>"""
import csv, re, subprocess, sys, collections
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
SORTK = 1 if (len(sys.argv) > 5 and sys.argv[5] == "inst") else 0
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = rows[1] if rows[0][0] == 'Kernel Name' else rows[0]
ia, isrc, ismp, iinst, ithr = (hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'),
                               hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'))
inst = [(r[isrc], float(r[ismp] or 0), float(r[iinst] or 0), float(r[ithr] or 0)) for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith('0x')]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
# find the function section
lines = []; cur = None; infn = False
for l in dis:
    if l.startswith('.text.') :
        infn = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
    if infn and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
print('sass instrs ncu', len(inst), 'nvdisasm', len(lines))
n = min(len(inst), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0.0])
for (src, smp, ni, thr), ln in zip(inst[:n], lines[:n]):
    a = agg[ln]; a[0] += smp; a[1] += ni; a[2] += ni * thr
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print('total samples', ts, 'total warp-inst', ti)
srcs = {}
def src(ln):
    if ln is None: return ''
    f, n_ = ln
    if f not in srcs:
        try: srcs[f] = open('/root/repo/magical_b200/csrc/' + f).read().splitlines()
        except Exception: srcs[f] = []
    return srcs[f][n_ - 1].strip()[:80] if 0 < n_ <= len(srcs[f]) else ''
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][SORTK])[:top]:
    print(f'{a[0]/ts*100:5.1f}% smp {a[1]/ti*100:5.1f}% inst thr={a[2]/max(a[1],1):4.1f}  {ln}  {src(ln)}')
