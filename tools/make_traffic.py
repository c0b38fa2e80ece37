#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv`
log of the bench workload: DRAM bytes (read + write) per launch of each kernel.

The median of the FULL launches is reported, not the mean: the log also holds the reset renders (k_raster
with only_fresh = 1 touches 1/240 of the batch in the prelude) and cold first launches, which must not dilute
the steady-state figure (the round-1 mean came out at 0.87x algorithmic where the
steady-state launches are 1.146x).  A launch counts as full when it moves at least half of what the kernel's
largest launch moves."""
import collections
import csv
import json
import statistics
import sys


def compute(csv_path, batch=65536, env_id='ClusterColour-Demo-LoRes4E-v0', source=None):
    """DRAM bytes per (steady-state) launch of every kernel in an ncu --csv metric log."""
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
    hdr = rows[0]
    ii, ki, mi, ui, vi = (hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'),
                          hdr.index('Metric Unit'), hdr.index('Metric Value'))
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    per_launch = collections.defaultdict(lambda: collections.defaultdict(lambda: {'r': 0.0, 'w': 0.0}))
    for r in rows[1:]:
        if not r[mi].startswith('dram__bytes_'):
            continue
        name = r[ki].split('(')[0].split('<')[0].replace('void ', '').strip()
        per_launch[name][r[ii]]['r' if 'read' in r[mi] else 'w'] += float(r[vi].replace(',', '')) * scale[r[ui]]
    out, detail = {}, {}
    for name, launches in per_launch.items():
        tot = sorted(v['r'] + v['w'] for v in launches.values())
        steady = [v for v in launches.values() if v['r'] + v['w'] >= 0.5 * tot[-1]]
        med = statistics.median(v['r'] + v['w'] for v in steady)
        out[name] = med
        detail[name] = {'launches': len(tot), 'full_launches': len(steady), 'median_total': med,
                        'median_read': statistics.median(v['r'] for v in steady),
                        'median_write': statistics.median(v['w'] for v in steady),
                        'min_total': tot[0], 'max_total': tot[-1]}
    out['_detail'] = detail
    out['_batch'] = batch
    out['_env_id'] = env_id
    out['_source'] = source or csv_path
    return out


if __name__ == '__main__':
    # usage: make_traffic.py <ncu csv> [source note] [batch] [env id]  -> writes profiles/traffic.json
    res = compute(sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else 65536,
                  sys.argv[4] if len(sys.argv) > 4 else 'ClusterColour-Demo-LoRes4E-v0',
                  sys.argv[2] if len(sys.argv) > 2 else None)
    json.dump(res, open('profiles/traffic.json', 'w'), indent=1)
    print(json.dumps(res, indent=1))
