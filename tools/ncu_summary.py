#!/usr/bin/env python
"""
Turn ncu captures into the tracked summaries under profiles/:

    python tools/ncu_summary.py full  <report.ncu-rep> <title> > profiles/rNN_ncu_<name>.md
    python tools/ncu_summary.py list  <launches.csv from --metrics gpu__time_duration.sum --csv> > profiles/rNN_launches.csv
    python tools/ncu_summary.py traffic <report.ncu-rep> <kernel regex> <reaches> <rows>   # JSON for profiles/traffic.json

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).
"""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sectors.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
]
STALL = re.compile(r'smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio')


def raw(report):
    out = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(value, unit):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    return float(value) * scale.get(unit, 1)


def full(report, title):
    hdr, units, rows = raw(report)
    col = {h: i for i, h in enumerate(hdr)}
    print(f'# {title}\n')
    print('Cold-cache, serialised launches under the profiler: compare shares and ratios, not absolute times; the timed '
          'numbers are the CUDA-event ones of bench.py / tools/configs_report.py.\n')
    seen = {}
    for r in rows:
        name = r[col['Kernel Name']]
        key = (name, r[col['launch__grid_size']] if 'launch__grid_size' in col else '')
        if key in seen:
            continue
        seen[key] = True
        print(f'## {name[:110]}\n')
        print('| metric | value | unit |\n|---|---|---|')
        for m in METRICS:
            if m in col:
                print(f'| {m} | {r[col[m]]} | {units[col[m]]} |')
        stalls = []
        for h, i in col.items():
            mm = STALL.match(h)
            if mm and 'per_issue_active' in h:
                try:
                    stalls.append((float(r[i]), mm.group(1)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            print('| top warp stalls (per issue) | ' + ', '.join(f'{n} {v:.2f}' for v, n in stalls[:6]) + ' | |')
        if 'dram__bytes_read.sum' in col:
            rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
            wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
            ms = float(r[col['gpu__time_duration.sum']]) * {'ms': 1, 'us': 1e-3, 's': 1e3, 'ns': 1e-6}.get(units[col['gpu__time_duration.sum']], 1)
            print(f'\nDRAM {rd / 1e9:.3f} GB read + {wr / 1e9:.3f} GB written in {ms:.3f} ms = {(rd + wr) / ms / 1e9:.2f} TB/s.\n')


def launch_list(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    print('id,kernel,grid,block,duration_us')
    for r in rows[1:]:
        if r[col['Metric Name']] != 'gpu__time_duration.sum':
            continue
        v = float(r[col['Metric Value']]) * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}.get(r[col['Metric Unit']], 1)
        print(f'{r[col["ID"]]},{r[col["Kernel Name"]][:70]!r},"{r[col["Grid Size"]]}","{r[col["Block Size"]]}",{v:.1f}')


def traffic(report, pattern, reaches, rows_):
    hdr, units, rows = raw(report)
    col = {h: i for i, h in enumerate(hdr)}
    best = None
    for r in rows:
        if re.search(pattern, r[col['Kernel Name']]):
            ms = float(r[col['gpu__time_duration.sum']])
            if best is None or ms > best[0]:
                best = (ms, r)
    r = best[1]
    rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
    wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
    sect = float(r[col['lts__t_sectors.sum']]) * 32
    per = float(reaches) * float(rows_)
    print(json.dumps({'dram_bytes_per_reach_step': (rd + wr) / per, 'dram_read_per_reach_step': rd / per,
                      'dram_write_per_reach_step': wr / per, 'l2_sector_bytes_per_reach_step': sect / per,
                      'source': f'ncu --set full, {report}: {r[col["Kernel Name"]][:40]}, {reaches} reaches x {rows_} steps'}, indent=1))


def traffic_step(report, reaches, rows_, variant, out_path):
    """DRAM bytes per reach-timestep of every kernel class of one bench step (largest launch of each class) ->
    profiles/traffic.json entry `variants[variant]` (what bench.py's roofline.traffic / frac_dram read)."""
    hdr, units, rows = raw(report)
    col = {h: i for i, h in enumerate(hdr)}
    per = float(reaches) * float(rows_)
    classes = {'route': 'rr_wavefront|rr_direct', 'permute_to_working': 'permute_to_working|stage_in', 'permute_to_user': 'permute_to_user|stage_out'}
    entry = {'reaches': int(reaches), 'rows': int(rows_), 'source': f'ncu --set full --clock-control none, {report}'}
    for cls, pat in classes.items():
        best = None
        for r in rows:
            if re.search(pat, r[col['Kernel Name']]):
                ms = float(r[col['gpu__time_duration.sum']])
                if best is None or ms > best[0]:
                    best = (ms, r)
        if best is None:
            continue
        r = best[1]
        rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
        wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
        entry[cls] = {'kernel': r[col['Kernel Name']][:48], 'dram_read_per_reach_step': rd / per, 'dram_write_per_reach_step': wr / per,
                      'dram_bytes_per_reach_step': (rd + wr) / per, 'ncu_ms': best[0],
                      'l2_sector_bytes_per_reach_step': float(r[col['lts__t_sectors.sum']]) * 32 / per,
                      'l2_hit_pct': float(r[col['lts__t_sector_hit_rate.pct']])}
    try:
        doc = json.load(open(out_path))
    except Exception:
        doc = {}
    doc.setdefault('variants', {})[variant] = entry
    json.dump(doc, open(out_path, 'w'), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == '__main__':
    {'full': full, 'list': launch_list, 'traffic': traffic, 'traffic_step': traffic_step}[sys.argv[1]](*sys.argv[2:])
