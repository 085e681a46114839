"""Warp-stall sampling summary from `ncu -i X.ncu-rep --page source --csv` of a warp-specialised kernel:
samples by code region (the regions are recognised by their marker opcodes: FFMA2 = depthwise producers,
LDTM = epilogue, UTMALDG / UTCHMMA = control warps, the out-of-line mbarrier spin loops at the end) with the
stall reasons of each, and the hottest instructions.   usage: python tools/ncu_stalls.py source.csv > summary.txt"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    print('#', rows[0][1] if len(rows[0]) > 1 else '')
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    ins = []
    for r in data:
        try:
            ins.append((int(r[ix['Address']], 16), r, int(r[ix['# Samples']])))
        except (ValueError, IndexError):
            continue
    ins.sort()
    base = ins[0][0]
    total = sum(s for _, _, s in ins)
    # 2 KB buckets labelled by the marker opcodes they contain
    buckets = collections.OrderedDict()
    for a, r, s in ins:
        b = buckets.setdefault((a - base) // 0x800, {'s': 0, 'marks': set(), 'stall': collections.Counter()})
        b['s'] += s
        src = r[ix['Source']]
        for key in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'FFMA2', 'STG.E.128', 'LDG.E.128', 'BAR.SYNC', 'UCGABAR', 'SYNCS.PHASECHK'):
            if key in src:
                b['marks'].add(key)
        for h in stalls:
            b['stall'][h[6:]] += int(r[ix[h]] or 0)
    print('total samples %d over %d instructions' % (total, len(ins)))
    print('%-8s %8s %6s  %-46s %s' % ('offset', 'samples', '%', 'marker opcodes', 'top stall reasons'))
    for k, b in buckets.items():
        if b['s'] == 0:
            continue
        print('%06x   %8d %5.1f%%  %-46s %s' % (k * 0x800, b['s'], 100.0 * b['s'] / total, ' '.join(sorted(b['marks'])),
                                              ' '.join('%s=%d' % kv for kv in b['stall'].most_common(4))))
    print('\nhottest instructions')
    for a, r, s in sorted(ins, key=lambda t: -t[2])[:40]:
        st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
        print('%6d %5.1f%%  +%06x  %-72s [%s]' % (s, 100.0 * s / total, a - base, r[ix['Source']][:72],
                                                ' '.join('%s=%d' % (h, v) for v, h in st)))


if __name__ == '__main__':
    main(sys.argv[1])
