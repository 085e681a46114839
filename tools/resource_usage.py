"""Per-kernel resource usage of the in-tree library (no GPU needed): registers, static shared memory, local-memory
stack, constant banks -- `cuobjdump -res-usage`, demangled.  Dynamic shared memory is requested at launch (see
DESIGN.md section 4 for the budgets).  Writes profiles/r2_resource_usage.txt.

    python tools/resource_usage.py
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'deephar_b200', 'libdeephar_b200.so')


def main():
    text = subprocess.run(['cuobjdump', '-res-usage', SO], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows, fn = [], None
    for line in text.splitlines():
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            fn = m.group(1)
            continue
        if fn and 'REG:' in line:
            f = dict(kv.split(':') for kv in line.split())
            rows.append((fn, int(f['REG']), int(f.get('SHARED', 0)), int(f.get('STACK', 0)), int(f.get('LOCAL', 0)),
                         sum(int(v) for k, v in f.items() if k.startswith('CONSTANT'))))
            fn = None
    names = [r[0] for r in rows]
    try:
        dem = subprocess.run(['cu++filt'] + names, stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
    except Exception:
        dem = names
    def strip_params(d):
        d = d.replace('void ', '').replace('(anonymous namespace)::', '')
        return d[:d.rindex('>(') + 1] if '>(' in d else d.split('(')[0]
    short = [strip_params(d) for d in dem]
    out = ['# cuobjdump -res-usage deephar_b200/libdeephar_b200.so (sm_100a); bytes except REG (32-bit registers per thread)',
           '# 65 536 registers per SM: a 512-thread CTA fits at <= 128 registers per thread on average (the warp-specialised',
           '# kernels re-balance with setmaxnreg); STACK > 0 = spill slots, see the STL/LDL columns of r2_sass_histogram.txt',
           '', '%-78s %5s %8s %6s %6s %9s' % ('kernel', 'REG', 'SHARED', 'STACK', 'LOCAL', 'CONSTANT')]
    for (fn, reg, sh, st, lo, co), s in sorted(zip(rows, short), key=lambda r: r[1]):
        out.append('%-78s %5d %8d %6d %6d %9d' % (s[:78], reg, sh, st, lo, co))
    path = os.path.join(ROOT, 'profiles', 'r2_resource_usage.txt')
    with open(path, 'w') as f:
        f.write('\n'.join(out) + '\n')
    print('\n'.join(out[:12]))
    print('... %d kernels -> %s' % (len(rows), path))


if __name__ == '__main__':
    main()
