"""Summarise `ncu --page raw --csv` exports into profiles/: per kernel launch the duration, DRAM bytes read + written,
DRAM throughput %, tensor-pipe activity %, registers, and (given the algorithmic bytes of the launch) the ratio of DRAM
traffic to algorithmic bytes.

  python scripts/summarise_ncu.py <raw.csv> <out.md> [--json key=kernel_regex ...] [--alg kernel_regex=bytes ...]

--json entries are merged into profiles/ncu_traffic.json as {key: {kernel, dram_bytes_per_launch, ...}} (mean over the
matching launches) — bench.py reads `roofline.traffic` from there."""
import csv
import json
import os
import re
import sys

WANT = {
    'gpu__time_duration.sum': 'duration_ns',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct',
    'sm__inst_executed_pipe_tensor.sum': 'tensor_inst',
    'launch__registers_per_thread': 'regs',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct',
}
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6, 'ns': 1, 'us': 1e3,
        'ms': 1e6}


def load(path):
    rows = list(csv.reader(open(path, newline='')))
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        rec = {'kernel': r[names.index('Kernel Name')], 'grid': r[names.index('Grid Size')] if 'Grid Size' in names else ''}
        for col, key in WANT.items():
            if col in names:
                i = names.index(col)
                try:
                    rec[key] = float(r[i].replace(',', '')) * UNIT.get(units[i], 1)
                except ValueError:
                    pass
        out.append(rec)
    return out


def main():
    raw, out_md = sys.argv[1], sys.argv[2]
    json_keys, alg = {}, {}
    args = sys.argv[3:]
    mode = None
    for a in args:
        if a in ('--json', '--alg'):
            mode = a
        elif mode == '--json':
            k, rx = a.split('=', 1)
            json_keys[k] = rx
        elif mode == '--alg':
            rx, b = a.rsplit('=', 1)
            alg[rx] = float(b)
    recs = load(raw)
    lines = ['| kernel | grid | duration (us) | DRAM read (MB) | DRAM write (MB) | DRAM % of peak | tensor pipe % | regs | '
             'traffic / algorithmic bytes |', '|---|---|---|---|---|---|---|---|---|']
    for r in recs:
        ratio = ''
        for rx, b in alg.items():
            if re.search(rx, r['kernel']):
                ratio = '{:.2f}'.format((r.get('dram_read', 0) + r.get('dram_write', 0)) / b)
        lines.append('| `{}` | {} | {:.1f} | {:.2f} | {:.2f} | {:.1f} | {:.1f} | {:.0f} | {} |'.format(
            r['kernel'][:70], r['grid'], r.get('duration_ns', 0) / 1e3, r.get('dram_read', 0) / 1e6,
            r.get('dram_write', 0) / 1e6, r.get('dram_pct', 0), r.get('tensor_pct', 0), r.get('regs', 0), ratio))
    os.makedirs(os.path.dirname(out_md) or '.', exist_ok=True)
    with open(out_md, 'w') as f:
        f.write('Source: `{}` (ncu --set full --clock-control none; per launch)\n\n'.format(os.path.basename(raw)))
        f.write('\n'.join(lines) + '\n')
    if json_keys:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json')
        data = json.load(open(path)) if os.path.exists(path) else {}
        for k, rx in json_keys.items():
            sel = [r for r in recs if re.search(rx, r['kernel'])]
            if sel:
                n = len(sel)
                data[k] = {'kernel_regex': rx, 'launches': n,
                           'dram_bytes_per_launch': sum(r.get('dram_read', 0) + r.get('dram_write', 0) for r in sel) / n,
                           'duration_us': sum(r.get('duration_ns', 0) for r in sel) / n / 1e3,
                           'tensor_pipe_pct': sum(r.get('tensor_pct', 0) for r in sel) / n,
                           'dram_pct_of_peak': sum(r.get('dram_pct', 0) for r in sel) / n,
                           'source': os.path.basename(raw)}
        json.dump(data, open(path, 'w'), indent=1, sort_keys=True)
    print('\n'.join(lines[:40]))


if __name__ == '__main__':
    main()
