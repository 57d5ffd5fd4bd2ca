"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel into a markdown table."""
import collections
import csv
import re
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ''
lines = [l for l in open(src) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(row['Metric Unit'], 1e-6)
    name = row['Kernel Name']
    short = re.sub(r'\(.*', '', re.sub(r'<.*', '', name))[:80]
    if short.startswith('void at::') or short.startswith('at::'):
        short = 'torch: ' + short.replace('void ', '')
    agg[short][0] += 1
    agg[short][1] += v
    tot += v
with open(dst, 'w') as f:
    f.write(f'# {title}\n\n')
    f.write('Source: `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off` around ONE '
            'training step (tools/profile_step.py; B=4, 800x1344, after 3 warm-up steps).  Per-launch times under ncu '
            'are cold-cache and serialised: compare SHARES, not absolutes.\n\n')
    f.write(f'Total kernel time {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches.\n\n')
    f.write('| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n')
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        f.write(f'| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |\n')
    ours = sum(ms for k, (n, ms) in agg.items() if k.startswith('lsn::') or 'lsn::' in k)
    f.write(f'\nKernels of liblsnet_sm100.so (`lsn::*`): {ours:.2f} ms = {100 * ours / tot:.1f}% of the step.\n')
print(open(dst).read()[:1500])
