#!/usr/bin/env python3
"""Turn ncu outputs brought back from the GPU box into the markdown summaries under profiles/.

    profile_summary.py launches <launches.csv> <out.md> "<command line that produced it>"
    profile_summary.py kernels  <report.ncu-rep> <out.md> "<command line that produced it>"
"""
import collections
import re
import csv
import io
import subprocess
import sys


def launches(path, out, cmd):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v
        k = r[ki].split('(')[0].replace('void ', '')
        # every kernel of the library runs behind the generic chain-batched wrapper: name the body
        m = re.search(r'batched_kernel(?:_nb)?<&\(?(?:void )?([\w:]+(?:<[^>]*>)?)', r[ki])
        if m:
            k = m.group(1)
        agg[k][0] += 1
        agg[k][1] += v
        agg[k][2] = max(agg[k][2], v)
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write(f'# ncu launch list\n\nCommand (B200, via gpurun): `{cmd}`\n\n'
                'Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n'
                '| kernel | launches | total us | avg us | max us | share |\n|---|---|---|---|---|---|\n')
        for k, (n, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| {k[:70]} | {n} | {t:.1f} | {t / n:.1f} | {m:.1f} | {t / tot:.1%} |\n')
        f.write(f'\ntotal {tot:.0f} us over {sum(v[0] for v in agg.values())} launches\n')


WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__waves_per_multiprocessor', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
    'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
]


def kernels(path, out, cmd):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    ki = hdr.index('Kernel Name')
    with open(out, 'w') as f:
        f.write(f'# ncu --set full captures\n\nCommand (B200, via gpurun): `{cmd}`\n\n'
                'Values are per launch; durations under ncu are cold-cache (not bench numbers).\n')
        for r in rows[2:]:
            f.write(f'\n## {r[ki].split("(")[0].replace("void ", "")}\n\n| metric | value | unit |\n|---|---|---|\n')
            for w, i in idx:
                f.write(f'| {w} | {r[i]} | {units[i]} |\n')
            try:
                rd = float(r[hdr.index('dram__bytes_read.sum')])
                wr = float(r[hdr.index('dram__bytes_write.sum')])
                ur, uw = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_write.sum')]
                f.write(f'| traffic (read+write) | {rd:.3f} {ur} + {wr:.3f} {uw} | |\n')
            except (ValueError, IndexError):
                pass


if __name__ == '__main__':
    {'launches': launches, 'kernels': kernels}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
