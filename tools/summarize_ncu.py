"""Extracts the headline metrics of `ncu --set full` captures (gpurun_out/full_*.ncu-rep) into profiles/*.md."""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def main(rep, out, title, algo_bytes=None, algo_flops=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write(f'# {title}\n\nSource: `{rep}` (`ncu --set full --clock-control none --import-source on`, one launch, '
                'tools/bench_kernels.py shapes).  Times under ncu are replayed/cold; the live CUDA-event numbers are in '
                'the bench JSON.\n\n')
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write(f"## `{d.get('Kernel Name', '?')[:110]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f'| {k} | {d[k]} | {units[hdr.index(k)]} |\n')
            try:
                t = float(d['gpu__time_duration.sum'].replace(',', ''))
                tu = units[hdr.index('gpu__time_duration.sum')]
                t_s = t * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'msecond': 1e-3}.get(tu, 1e-6)
                rd = float(d['dram__bytes_read.sum'].replace(',', '')); wr = float(d['dram__bytes_write.sum'].replace(',', ''))
                sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
                rd *= sc.get(units[hdr.index('dram__bytes_read.sum')], 1); wr *= sc.get(units[hdr.index('dram__bytes_write.sum')], 1)
                f.write(f'\nDRAM traffic (read+write) = {(rd + wr) / 1e6:.1f} MB in {t_s * 1e6:.1f} us = {(rd + wr) / t_s / 1e9:.0f} GB/s under ncu.\n')
                if algo_bytes:
                    f.write(f'Algorithmic bytes = {float(algo_bytes) / 1e6:.1f} MB -> traffic/algorithmic = {(rd + wr) / float(algo_bytes):.2f}.\n')
                if algo_flops:
                    f.write(f'Algorithmic FLOPs = {float(algo_flops) / 1e9:.1f} GF -> {float(algo_flops) / t_s / 1e12:.0f} TFLOP/s under ncu.\n')
            except Exception as e:
                f.write(f'\n(derived numbers unavailable: {e})\n')
            f.write('\n')
    print(open(out).read()[:2500])


if __name__ == '__main__':
    main(*sys.argv[1:])
