"""Condense ncu CSV exports into the small tables kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/<tag>_launches.csv          -> per-kernel count / avg us / share
  python tools/ncu_summary.py raw gpurun_out/<tag>_s64.raw.csv                -> per-kernel key metrics of a --set full capture
"""
from __future__ import annotations

import csv
import sys
from collections import OrderedDict, defaultdict

KEY_METRICS = [
    ('gpu__time_duration.sum', 'dur'),
    ('smsp__inst_executed.sum', 'inst'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__occupancy_limit_registers', 'occ_lim_regs'),
    ('dram__bytes_read.sum', 'dram_rd'),
    ('dram__bytes_write.sum', 'dram_wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('lts__t_sectors.sum', 'l2_sectors'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('l1tex__t_sector_hit_rate.pct', 'l1_hit_pct'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1_pipe_pct'),
]
STALLS = ['barrier', 'branch_resolving', 'dispatch_stall', 'drain', 'lg_throttle', 'long_scoreboard', 'math_pipe_throttle',
          'membar', 'mio_throttle', 'misc', 'no_instruction', 'not_selected', 'selected', 'short_scoreboard', 'sleeping',
          'tex_throttle', 'wait']
UNIT_SCALE = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0,
              'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 'second': 1.0, 'sector': 1.0}


def short(name):
    return name.split('(')[0].replace('void ', '').strip()


def launches(path):
    rows = defaultdict(list)
    with open(path, newline='') as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(unit, 1.0)
        rows[short(r['Kernel Name']) + ' grid ' + r['Grid Size'].replace(' ', '')].append(v)
    tot = sum(sum(v) for v in rows.values())
    print('| kernel (grid) | launches | avg us | total us | share |\n|---|---|---|---|---|')
    for k, v in sorted(rows.items(), key=lambda kv: -sum(kv[1])):
        print(f'| {k} | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.0f} | {100 * sum(v) / tot:.1f} % |')
    print(f'\ntotal {tot:.0f} us over {sum(len(v) for v in rows.values())} launches')


def raw(path):
    """Per kernel (averaged over its launches in the capture): the key metrics of a `--set full` export, the derived
    DRAM and L2 GB/s (bytes moved / kernel duration; L2 = lts__t_sectors x 32 B) and the three largest warp-stall reasons
    (warps stalled per issue-active cycle)."""
    with open(path, newline='') as f:
        rd = csv.reader(f)
        hdr = next(rd)
        units = next(rd)
        col = {h: i for i, h in enumerate(hdr)}
        seen = OrderedDict()
        for r in rd:
            if len(r) != len(hdr):
                continue
            name = short(r[col['Kernel Name']]) + ' grid ' + r[col['Grid Size']].replace(' ', '')
            seen.setdefault(name, []).append(r)

    def val(rs, m, scaled=False):
        if m not in col:
            return None
        vals = []
        for r in rs:
            try:
                v = float(r[col[m]].replace(',', ''))
            except ValueError:
                continue
            vals.append(v * (UNIT_SCALE.get(units[col[m]].lower(), 1.0) if scaled else 1.0))
        return sum(vals) / len(vals) if vals else None

    names = [n for _, n in KEY_METRICS] + ['dram_GBs', 'l2_GBs', 'top stalls (warps / issue-active cycle)']
    print('| kernel (grid) | n | ' + ' | '.join(names) + ' |\n|---|---|' + '---|' * len(names))
    for k, rs in seen.items():
        cells = []
        for m, _ in KEY_METRICS:
            v = val(rs, m)
            cells.append('-' if v is None else f'{v:.4g} {units[col[m]]}'.strip())
        dur = val(rs, 'gpu__time_duration.sum', True)
        rdb, wrb = val(rs, 'dram__bytes_read.sum', True), val(rs, 'dram__bytes_write.sum', True)
        sect = val(rs, 'lts__t_sectors.sum', True)
        cells.append('-' if not dur or rdb is None else f'{(rdb + wrb) / dur / 1e9:.1f}')
        cells.append('-' if not dur or sect is None else f'{sect * 32 / dur / 1e9:.1f}')
        st = [(val(rs, f'smsp__average_warps_issue_stalled_{n}_per_issue_active.ratio') or 0.0, n) for n in STALLS]
        st = sorted((x for x in st if x[1] != 'selected'), reverse=True)[:3]
        cells.append(', '.join(f'{n} {v:.2f}' for v, n in st))
        print(f'| {k} | {len(rs)} | ' + ' | '.join(cells) + ' |')


def traffic(path):
    """DRAM bytes (read + write) of ONE steady-state frame chain from a --set full capture: every kernel name once
    (k_pyr_down / k_stereo_candidates may launch more than once per frame: their launches inside one frame are summed
    by dividing the capture's total by the number of frames it covers = launches of k_finish)."""
    import json
    with open(path, newline='') as f:
        rd = csv.reader(f)
        hdr = next(rd)
        units = next(rd)
        col = {h: i for i, h in enumerate(hdr)}
        per = defaultdict(lambda: [0, 0.0])
        for r in rd:
            if len(r) != len(hdr):
                continue
            name = short(r[col['Kernel Name']])
            tot = 0.0
            for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                v = float(r[col[m]].replace(',', ''))
                u = units[col[m]].lower()
                tot += v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
            per[name][0] += 1
            per[name][1] += tot
    frames = max(per.get('k_finish', [1])[0], 1)
    out = {'frames_in_capture': frames, 'source': path,
           'per_kernel_bytes_per_frame': {k: v[1] / frames for k, v in per.items() if not k.startswith('k_stereo_buckets')},
           'note': 'ncu --set full replays every kernel with flushed caches: cold-cache DRAM traffic, an upper bound on the '
                   'traffic of the warm chain'}
    out['bytes_per_frame'] = sum(out['per_kernel_bytes_per_frame'].values())
    print(json.dumps(out, indent=1))


def traffic_warm(path, cold_json=None):
    """DRAM bytes per steady-state frame chain from a `--cache-control none --metrics dram__bytes_read.sum,
    dram__bytes_write.sum` pass (one pass per kernel, caches NOT flushed between kernels: the traffic of the chain as it
    runs, previous-frame pyramids and the just-copied input block still in L2).  With `cold_json` (output of `traffic`)
    the result is merged into it."""
    import json
    per = defaultdict(lambda: [0, 0.0])
    with open(path, newline='') as f:
        lines = [l for l in f if l.startswith('"')]
    ids = defaultdict(set)
    for r in csv.DictReader(lines):
        m = r.get('Metric Name')
        if m not in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            continue
        name = short(r['Kernel Name'])
        v = float(r['Metric Value'].replace(',', ''))
        v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(r['Metric Unit'].lower(), 1)
        per[name][1] += v
        ids[name].add(r['ID'])
    frames = max(len(ids.get('k_finish', [1])), 1)
    # the capture window need not hold whole frames: average bytes per launch x that kernel's launches per frame
    warm = {k: v[1] / len(ids[k]) * max(1, round(len(ids[k]) / frames)) for k, v in per.items()
            if not k.startswith('k_stereo_buckets')}
    out = json.load(open(cold_json)) if cold_json else {}
    out['warm'] = {'frames_in_capture': frames, 'source': path, 'per_kernel_bytes_per_frame': warm,
                   'bytes_per_frame': sum(warm.values()),
                   'note': 'ncu --cache-control none, DRAM counters only (single pass, no replay): caches keep what the '
                           'preceding copies and kernels left in them, as in the running chain'}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    {'launches': launches, 'raw': raw, 'traffic': traffic, 'traffic_warm': traffic_warm}[sys.argv[1]](*sys.argv[2:])
