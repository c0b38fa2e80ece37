#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv`
log of the bench workload: DRAM bytes (read + write) per launch of each kernel, averaged."""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, ui, vi = (hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'),
                  hdr.index('Metric Value'))
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
acc = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    if not r[mi].startswith('dram__bytes_'):
        continue
    name = r[ki].split('(')[0].split('<')[0].replace('void ', '').strip()
    acc[name][0] += float(r[vi].replace(',', '')) * scale[r[ui]]
    acc[name][1] += 1
out = {k: v[0] / (v[1] / 2) for k, v in acc.items()}  # two metrics per launch
out['_batch'] = 65536
out['_env_id'] = 'ClusterColour-Demo-LoRes4E-v0'
out['_source'] = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
json.dump(out, open('profiles/traffic.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
