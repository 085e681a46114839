"""profiles/*.raw.csv (`ncu -i X.ncu-rep --page raw --csv`) -> profiles/r2_traffic.json + a short summary.

bench.py's `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch)
comes from HERE, i.e. from an ncu capture of the shipped build, not from a literal in bench.py.  Usage:

    python tools/ncu_traffic.py LABEL FRAMES profiles/r2_xxx.raw.csv [LABEL FRAMES CSV ...]

LABEL  = the kernel-profile label bench.py uses (e.g. "sepconv 32x32x576->32x32x576 k5x5"),
FRAMES = frames in the captured launch (tools/prof_conv.py N); the JSON stores bytes per frame so that bench.py
can scale to its own launch size.  The LAST row of the CSV (the last captured launch: warm) is used.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
KEYS = ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread')


def read_last(path):
    rows = list(csv.reader(open(path)))
    hdr, units, last = rows[0], rows[1], rows[-1]
    out = {'kernel': last[hdr.index('Kernel Name')], 'grid': last[hdr.index('Grid Size')], 'block': last[hdr.index('Block Size')]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            try:
                out[k] = (float(last[i].replace(',', '')), units[i])
            except ValueError:
                pass
    return out


def main(argv):
    table, lines = {}, []
    path_json = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if os.path.exists(path_json):
        table = json.load(open(path_json))
    for label, frames, path in zip(argv[0::3], argv[1::3], argv[2::3]):
        r = read_last(path)
        rd = r['dram__bytes_read.sum'][0] * UNIT[r['dram__bytes_read.sum'][1]]
        wr = r['dram__bytes_write.sum'][0] * UNIT[r['dram__bytes_write.sum'][1]]
        frames = int(frames)
        table[label] = {'bytes_per_frame': (rd + wr) / frames,
                        'source': '%s (%d-frame launch of %s: %.1f MB read + %.1f MB written, %.1f us under ncu), '
                                  'scaled per frame' % (os.path.relpath(path, ROOT), frames, r['kernel'].split('(')[0],
                                                        rd / 1e6, wr / 1e6, r.get('gpu__time_duration.sum', (0, ''))[0])}
        lines.append('%s | %s grid %s block %s' % (label, r['kernel'].split('(')[0], r['grid'], r['block']))
        for k in KEYS:
            if k in r:
                lines.append('    %-70s %14.3f %s' % (k, r[k][0], r[k][1]))
    with open(path_json, 'w') as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print('\n'.join(lines))
    print('wrote', path_json)


if __name__ == '__main__':
    main(sys.argv[1:])
