"""Aggregates an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` pass over ONE
training step by kernel class: launches, device time, DRAM bytes per launch.  Writes profiles/<name>.json (read by
bench.py for `roofline.traffic`) and profiles/<name>.md.
usage: python tools/step_traffic.py gpurun_out/step_traffic.csv profiles/r02_step_traffic"""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
UNIT_T = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}
UNIT_B = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
per = collections.defaultdict(lambda: dict(ms=0.0, rd=0.0, wr=0.0))       # per launch id
names = {}
for row in csv.DictReader(lines):
    i = row['ID']
    names[i] = row['Kernel Name']
    v = float(row['Metric Value'].replace(',', '') or 0)
    m, u = row['Metric Name'], row['Metric Unit']
    if m == 'gpu__time_duration.sum':
        per[i]['ms'] += v * UNIT_T.get(u, 1e-6)
    elif m == 'dram__bytes_read.sum':
        per[i]['rd'] += v * UNIT_B.get(u, 1.0)
    elif m == 'dram__bytes_write.sum':
        per[i]['wr'] += v * UNIT_B.get(u, 1.0)


def cls(name):
    short = re.sub(r'\(.*', '', re.sub(r'<.*', '', name)).replace('void ', '')
    if 'lsn::' in short:
        short = short.split('lsn::')[-1]
        return re.sub(r'_kernel$', '', short)
    if any(k in short for k in ('sgd_momentum', 'pred_reg_', 'add_softplus')):      # library kernels outside namespace lsn
        return re.sub(r'_kernel$', '', short)
    if short.startswith('at::'):
        return 'torch elementwise / reductions'
    if 'nccl' in short.lower():
        return 'nccl'
    return 'library (cuDNN / cuBLAS): ' + short[:40]


agg = collections.defaultdict(lambda: dict(launches=0, ms=0.0, rd=0.0, wr=0.0))
for i, d in per.items():
    a = agg[cls(names[i])]
    a['launches'] += 1
    a['ms'] += d['ms']; a['rd'] += d['rd']; a['wr'] += d['wr']
tot_ms = sum(a['ms'] for a in agg.values())
out = dict(source=src, note='per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes',
           total_kernel_ms=tot_ms, launches=sum(a['launches'] for a in agg.values()), classes={})
# names as bench.py's kernel classes use them
alias = {'gemm_kmajor': 'gemm_kmajor', 'gemm_mnmajor': 'gemm_mnmajor', 'dcn_im2col': 'dcn_im2col',
         'dcn_col2im_binned': 'dcn_col2im', 'dcn_col2im': 'dcn_col2im', 'dcn_adjoint': 'dcn_col2im',
         'dcn_fused_fwd': 'dcn_fused_fwd', 'dcn_fused_wgrad': 'dcn_fused_wgrad'}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
    e = dict(launches=a['launches'], ms=a['ms'], share=a['ms'] / tot_ms, dram_read_bytes=a['rd'], dram_write_bytes=a['wr'],
             dram_bytes_per_launch=(a['rd'] + a['wr']) / a['launches'],
             dram_gbs=(a['rd'] + a['wr']) / (a['ms'] * 1e-3) / 1e9 if a['ms'] > 0 else 0.0)
    out['classes'][k] = e
    if k in alias and alias[k] != k:
        t = out['classes'].setdefault(alias[k], dict(e))
json.dump(out, open(dst + '.json', 'w'), indent=1)
with open(dst + '.md', 'w') as f:
    f.write('# DRAM traffic and device time per kernel class over ONE training step (B=4, 800x1344, CUDA-graph step)\n\n')
    f.write('`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none '
            '--profile-from-start off python tools/profile_step.py`; per-launch times are cold-cache and serialised '
            '(compare shares).\n\n')
    f.write(f'{out["launches"]} launches, {tot_ms:.2f} ms of kernel time.\n\n')
    f.write('| kernel class | launches | ms | share | DRAM read MB | DRAM write MB | MB / launch | GB/s |\n|---|---:|---:|---:|---:|---:|---:|---:|\n')
    for k, e in list(out['classes'].items())[:40]:
        f.write(f"| `{k}` | {e['launches']} | {e['ms']:.3f} | {100 * e['share']:.1f}% | {e['dram_read_bytes'] / 1e6:.1f} | "
                f"{e['dram_write_bytes'] / 1e6:.1f} | {e['dram_bytes_per_launch'] / 1e6:.2f} | {e['dram_gbs']:.0f} |\n")
print(open(dst + '.md').read()[:3000])
