"""SASS opcode evidence for the in-tree library (no GPU needed): per-kernel counts of the tcgen05 / TMEM / TMA /
mbarrier opcodes + the top opcodes overall.  Writes profiles/r2_sass_histogram.txt.

    python tools/sass_histogram.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'deephar_b200', 'libdeephar_b200.so')
KEYS = ['UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'UTCATOMSWS', 'FFMA2', 'FFMA', 'STL', 'LDL']
LEGEND = ('UTCHMMA = tcgen05.mma kind::f16 | LDTM = tcgen05.ld | UTMALDG = cp.async.bulk.tensor (TMA load) | '
          'UBLKCP = cp.async.bulk (DSMEM push, 1-D bulk) | UTCBAR = tcgen05.commit | SYNCS = mbarrier ops | '
          'UTCATOMSWS = TMEM alloc/dealloc | STL/LDL = register spills')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', SO], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = {}
    per, total, allops = collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
    fn = None
    ins = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)')
    for line in sass.splitlines():
        if 'Function :' in line:
            fn = line.split('Function :')[1].strip()
            continue
        m = ins.match(line)
        if m and fn:
            op = m.group(1)
            per[fn][op] += 1
            total[fn] += 1
            allops[op] += 1
    names = list(total)
    try:
        out = subprocess.run(['cu++filt'] + names, stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    except Exception:
        pass
    ver = subprocess.run(['nvcc', '--version'], stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()[-1]
    lines = ['# cuobjdump -sass deephar_b200/libdeephar_b200.so   (%s)' % ver, '# ' + LEGEND, '',
             '%-74s %7s ' % ('kernel', 'instrs') + ' '.join('%7s' % k[:7] for k in KEYS)]
    for f in sorted(names, key=lambda f: -per[f]['UTCHMMA']):
        nm = demangle.get(f, f)
        nm = (nm[:nm.rfind('(')] if '(' in nm else nm).replace('void ', '').replace('(bool)', '')
        lines.append('%-74s %7d ' % (nm[:74], total[f]) + ' '.join('%7d' % per[f][k] for k in KEYS))
    lines += ['', '# all opcodes, whole library (top 40)']
    lines += ['%9d %s' % (c, op) for op, c in allops.most_common(40)]
    path = os.path.join(ROOT, 'profiles', 'r2_sass_histogram.txt')
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\n')
    print('\n'.join(lines[:34]))
    print('wrote', path)


if __name__ == '__main__':
    main()
