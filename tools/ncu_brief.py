"""Condensed view of one `ncu --set full` report: headline metrics + stall reasons + the hottest SASS lines.
usage: python tools/ncu_brief.py gpurun_out/x.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe', 'sm__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__shared_mem_config_size', 'launch__grid_size',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'sm__inst_executed_pipe_lsu', 'smsp__inst_executed.avg.per_cycle_active',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__pipe_tensor_cycles_active_realtime', 'gpc__cycles_elapsed.avg.per_second']
print('kernel:', vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) or w in h for w in want) and 'peak_sustained' not in h.replace('pct_of_peak_sustained', ''):
        print(f'  {h:95s} {u:12s} {v}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
data = rows[2:]
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
agg = {k: sum(int(r[ix[k]] or 0) for r in data) for k in stalls}
print('samples', tot, ' '.join(f'{k[6:]}={100 * v / tot:.1f}%' for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:ntop]:
    st = sorted(((k[6:], int(r[ix[k]] or 0)) for k in stalls), key=lambda x: -x[1])[:2]
    print(f"  {r[ix['Address']][-5:]} {r[ix['# Samples']]:>6} ex={r[ix['Instructions Executed']]:>9} {r[ix['Source']][:64]:64s} {st}")
