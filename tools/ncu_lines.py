#!/usr/bin/env python3
"""Join an ncu SASS source page (csv) with nvdisasm line info to get per-source-line
instruction counts and stall samples (with the dominant stall reasons).
usage: ncu_lines.py <rep> <cubin> <kernel-substring-in-mangled-name> [top] [inst]
The report may hold several kernels; the one whose demangled name contains the
part of <kernel-substring> before the first 'I' (template marker) is used."""
import collections
import csv
import re
import subprocess
import sys

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
SORTK = 1 if (len(sys.argv) > 5 and sys.argv[5] == "inst") else 0
plain = kname.split('I')[0]
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
# split per kernel
blocks = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        blocks.append([r[1], None, []])
    elif r and r[0] == 'Address' and blocks:
        blocks[-1][1] = r
    elif r and r[0].startswith('0x') and blocks:
        blocks[-1][2].append(r)
blk = next(b for b in blocks if plain in b[0])
hdr = blk[1]
ia, isrc, ismp, iinst, ithr = (hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'),
                               hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'))
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
inst = []
for r in blk[2]:
    st = {n: float(r[i] or 0) for i, n in stall_cols}
    inst.append((r[isrc], float(r[ismp] or 0), float(r[iinst] or 0), float(r[ithr] or 0), st))
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
lines = []
cur = None
infn = False
for l in dis:
    if l.startswith('.text.'):
        infn = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
    if infn and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
print(blk[0][:70], '| sass instrs ncu', len(inst), 'nvdisasm', len(lines))
n = min(len(inst), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0.0, collections.Counter()])
tot_stall = collections.Counter()
for (src, smp, ni, thr, st), ln in zip(inst[:n], lines[:n]):
    a = agg[ln]
    a[0] += smp
    a[1] += ni
    a[2] += ni * thr
    a[3].update(st)
    tot_stall.update(st)
ts = sum(a[0] for a in agg.values())
ti = sum(a[1] for a in agg.values())
print('total samples', ts, 'total warp-inst', ti)
print('stalls overall:', ' '.join(f'{k}={v/ts*100:.1f}%' for k, v in tot_stall.most_common(8)))
srcs = {}


def src(ln):
    if ln is None:
        return ''
    f, n_ = ln
    if f not in srcs:
        try:
            srcs[f] = open('/root/repo/magical_b200/csrc/' + f).read().splitlines()
        except Exception:
            srcs[f] = []
    return srcs[f][n_ - 1].strip()[:70] if 0 < n_ <= len(srcs[f]) else ''


for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][SORTK])[:top]:
    why = ','.join(f'{k}:{v/max(a[0],1)*100:.0f}' for k, v in a[3].most_common(2))
    print(f'{a[0]/ts*100:5.1f}% smp {a[1]/ti*100:5.1f}% inst thr={a[2]/max(a[1],1):4.1f} [{why}] {ln}  {src(ln)}')
