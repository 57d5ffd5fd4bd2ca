"""Kernel timeline of the CUDA-graph training step via torch.profiler (CUPTI; no nsys in the image): per-kernel totals,
per-stream busy time, the union busy time of the GPU and the idle gaps inside the step -> gpurun_out/trace_summary.md."""
import collections
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from lsnet_b200.data import MODEL_CFG, synthetic_batch
from lsnet_b200.train import GraphTrainer

host = [synthetic_batch(s, 0, 4, (800, 1333), pin=True) for s in range(2)]
tr = GraphTrainer(MODEL_CFG['bbox_r50'], host[0], device='cuda:0')
for w in range(3):
    tr.step(host[w % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(host[1])
    torch.cuda.synchronize()
os.makedirs('gpurun_out', exist_ok=True)
prof.export_chrome_trace('gpurun_out/trace_step.json')
ev = [e for e in json.load(open('gpurun_out/trace_step.json'))['traceEvents']
      if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
ev.sort(key=lambda e: e['ts'])
t0, t1 = ev[0]['ts'], max(e['ts'] + e['dur'] for e in ev)
by_name = collections.defaultdict(lambda: [0, 0.0])
by_stream = collections.defaultdict(float)
for e in ev:
    n = e['name']
    for pre in ('void lsn::', 'lsn::', 'void at::native::', 'at::native::', 'void '):
        if n.startswith(pre):
            n = n[len(pre):]
    n = n.split('<')[0].split('(')[0][:60]
    by_name[n][0] += 1
    by_name[n][1] += e['dur']
    by_stream[e['args'].get('stream', -1)] += e['dur']
# union busy time
busy, cur_end = 0.0, t0
gaps = []
for e in ev:
    s, d = e['ts'], e['ts'] + e['dur']
    if s > cur_end:
        gaps.append((s - cur_end, e['name'][:50]))
        busy += d - s
        cur_end = d
    elif d > cur_end:
        busy += d - cur_end
        cur_end = d
with open('gpurun_out/trace_summary.md', 'w') as f:
    f.write(f'# step timeline (torch.profiler / CUPTI), {len(ev)} GPU activities\n\n')
    f.write(f'span {1e-3 * (t1 - t0):.2f} ms, GPU busy (union) {1e-3 * busy:.2f} ms, idle {1e-3 * (t1 - t0 - busy):.2f} ms, '
            f'sum of kernel durations {1e-3 * sum(e["dur"] for e in ev):.2f} ms\n\n')
    f.write('| stream | busy ms |\n|---|---:|\n')
    for s, d in sorted(by_stream.items(), key=lambda x: -x[1]):
        f.write(f'| {s} | {1e-3 * d:.2f} |\n')
    f.write('\n| kernel | launches | total ms |\n|---|---:|---:|\n')
    for n, (c, d) in sorted(by_name.items(), key=lambda x: -x[1][1])[:45]:
        f.write(f'| `{n}` | {c} | {1e-3 * d:.3f} |\n')
    # per-launch view of the two GEMM kernels: (grid size, template) -> launches, total, mean
    f.write('\n| GEMM launch group (kernel, grid) | launches | total ms | mean us |\n|---|---:|---:|---:|\n')
    grp = collections.defaultdict(list)
    for e in ev:
        if 'gemm_kmajor' in e['name'] or 'gemm_mnmajor' in e['name']:
            tmpl = e['name'].split('<')[1].split('>')[0] if '<' in e['name'] else ''
            kind = 'kmajor' if 'kmajor' in e['name'] else 'mnmajor'
            grp[(kind, tmpl, str(e['args'].get('grid')))].append(e['dur'])
    for k, v in sorted(grp.items(), key=lambda kv: -sum(kv[1])):
        f.write(f'| {k[0]}<{k[1]}> grid {k[2]} | {len(v)} | {1e-3 * sum(v):.3f} | {sum(v) / len(v):.1f} |\n')
    gaps.sort(reverse=True)
    f.write('\nlargest idle gaps (us, next kernel): ' + ', '.join(f'{g:.0f} ({n})' for g, n in gaps[:12]) + '\n')
print(open('gpurun_out/trace_summary.md').read())
# compact per-activity timeline for offline analysis: start us (relative), duration us, stream, grid, name
with open('gpurun_out/trace_timeline.csv', 'w') as f:
    f.write('start_us,dur_us,stream,grid,block,name\n')
    for e in ev:
        a = e['args']
        g = a.get('grid'); b = a.get('block')
        g = 'x'.join(str(v) for v in g) if isinstance(g, (list, tuple)) else str(g)
        b = 'x'.join(str(v) for v in b) if isinstance(b, (list, tuple)) else str(b)
        f.write(f"{e['ts'] - t0:.1f},{e['dur']:.1f},{a.get('stream', -1)},{g},{b},\"{e['name'][:90]}\"\n")
os.remove('gpurun_out/trace_step.json')
