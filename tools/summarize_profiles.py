"""Turns gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and *.ncu-rep captures into the
small text summaries committed under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches.csv profiles/r01_launches_<tag>.txt
    python tools/summarize_profiles.py rep gpurun_out/prof_conv.ncu-rep profiles/r01_ncu_<tag>.txt
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct']

EXTRA = re.compile(r'issue_stalled.*(ratio|pct)$|smsp__issue_active\.avg\.pct|smsp__inst_executed\.sum$|sm__inst_executed_pipe_[a-z0-9_]+\.sum$|'
                   r'smsp__warps_eligible\.avg\.per_cycle_active|smsp__thread_inst_executed_per_inst_executed')


def short(name):
    name = re.sub(r'\(.*', '', name)
    return name.replace('afcm::', '')


def launches(src, dst):
    """Launch list with device time (and, when the capture has them, DRAM bytes) per launch; also writes
    <dst>.traffic.json = {kernel: average DRAM read+write bytes per launch}, which bench.py reports as roofline.traffic."""
    import json
    per_id = OrderedDict()
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        e = per_id.setdefault(r['ID'], dict(k=short(r['Kernel Name']), g=r['Grid Size'], b=r['Block Size'], ns=0.0, dram=None))
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', '')
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            e['ns'] = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
        elif r.get('Metric Name') in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
            e['dram'] = (e['dram'] or 0.0) + v * mult
    rows = list(per_id.values())
    agg = OrderedDict()
    for e in rows:
        a = agg.setdefault(e['k'], [0, 0.0, 0.0, False])
        a[0] += 1; a[1] += e['ns']
        if e['dram'] is not None:
            a[2] += e['dram']; a[3] = True
    total = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write(f'# per-kernel totals over {len(rows)} launches (ncu gpu__time_duration.sum [+ dram__bytes_read/write.sum], --clock-control none;\n')
        f.write('# cold-cache, serialised: compare SHARES, not absolutes)\n')
        f.write(f'{"kernel":60s} {"launches":>8s} {"total_ms":>10s} {"avg_us":>9s} {"share":>7s} {"dram_MB/launch":>15s}\n')
        for k, (n, ns, dram, has) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{k:60s} {n:8d} {ns / 1e6:10.3f} {ns / n / 1e3:9.2f} {ns / total:7.3f} {(dram / n / 1e6 if has else float("nan")):15.2f}\n')
        f.write(f'{"TOTAL":60s} {len(rows):8d} {total / 1e6:10.3f}\n\n# launch list (id, kernel, us, grid, block, dram MB)\n')
        for i, e in enumerate(rows):
            f.write(f'{i:5d} {e["k"]:60s} {e["ns"] / 1e3:10.2f} {e["g"]:>16s} {e["b"]:>14s} {(e["dram"] or 0) / 1e6:10.2f}\n')
    traffic = {k: dict(launches=a[0], dram_bytes_per_launch=a[2] / a[0]) for k, a in agg.items() if a[3]}
    if traffic:
        with open(dst + '.traffic.json', 'w') as f:
            json.dump(traffic, f, indent=1, sort_keys=True)
    print(open(dst).read().split('# launch list')[0])


def rep(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    rd = list(csv.reader(lines))
    hdr, units, data = rd[0], rd[1], rd[2:]
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full --clock-control none capture: {src}\n')
        for row in data:
            d = dict(zip(hdr, row))
            f.write(f'\n== {short(d["Kernel Name"])}  grid {d.get("Grid Size")} block {d.get("Block Size")}\n')
            for k in KEYS:
                if k in d:
                    f.write(f'  {k:80s} {d[k]:>18s} {units[hdr.index(k)]}\n')
            # warp-state breakdown (why the issue slots idle) and per-pipe instruction counts, whatever this ncu calls them
            for k in hdr:
                if k not in KEYS and EXTRA.search(k) and d.get(k) not in (None, '', 'n/a'):
                    f.write(f'  {k:80s} {d[k]:>18s} {units[hdr.index(k)]}\n')
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'rep': rep}[sys.argv[1]](sys.argv[2], sys.argv[3])
