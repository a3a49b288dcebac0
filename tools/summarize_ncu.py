"""Turn ncu captures from gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python tools/summarize_ncu.py kernel   gpurun_out/r1_attn.ncu-rep  profiles/r1_attn64.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__cycles_elapsed.avg.per_second', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
    'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    h = rows[0]
    ki, vi, ii = h.index('Kernel Name'), h.index('Metric Value'), h.index('ID')
    per = []
    for r in rows[1:]:
        try:
            per.append((int(r[ii]), r[ki], float(r[vi].replace(',', ''))))
        except ValueError:
            pass
    # one forward = from one positions_kernel launch to the next
    starts = [i for i, p in enumerate(per) if 'positions_kernel' in p[1]]
    seg = per[starts[-2]:starts[-1]] if len(starts) >= 2 else per
    tot = sum(p[2] for p in seg)
    agg, cnt = collections.defaultdict(float), collections.Counter()
    for _, name, ns in seg:
        short = name.split('(')[0].replace('void ', '').replace('esmk::', '').replace('<unnamed>::', '')
        agg[short] += ns
        cnt[short] += 1
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list ({src}): one forward of `bench.py` (ESM2-650M, config 2)\n\n')
        f.write('`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and '
                'serialised: compare SHARES with the bench `kernels` block, not absolutes.\n\n')
        f.write(f'launches in this forward: {len(seg)}; sum of kernel durations: {tot / 1e6:.2f} ms\n\n')
        f.write('| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n')
        for k, v in sorted(agg.items(), key=lambda x: -x[1]):
            f.write(f'| `{k}` | {cnt[k]} | {v / 1e6:.3f} | {100 * v / tot:.1f} % |\n')
        f.write('\n## first layer, launch by launch (ns)\n\n```\n')
        for i, name, ns in seg[:14]:
            f.write(f'{i:6d} {ns:12.0f}  {name[:110]}\n')
        f.write('```\n')


def kernel(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full summary: {src}\n\n')
        for data in rows[2:]:
            d = {n: (data[i], u[i]) for i, n in enumerate(h)}
            f.write(f'## {d["Kernel Name"][0][:150]}\n\n| metric | value | unit |\n|---|---:|---|\n')
            for k in KEYS:
                if k in d and d[k][0] != '':
                    f.write(f'| {k} | {d[k][0]} | {d[k][1]} |\n')
            try:
                rd = float(d['dram__bytes_read.sum'][0]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[d['dram__bytes_read.sum'][1]]
                wr = float(d['dram__bytes_write.sum'][0]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[d['dram__bytes_write.sum'][1]]
                f.write(f'| DRAM traffic (read + write) | {(rd + wr) / 1e6:.1f} | MB per launch |\n')
            except Exception:
                pass
            f.write('\n')


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
