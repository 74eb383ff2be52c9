"""Opcode histogram per kernel of libavb.so (cuobjdump -sass): the evidence for what the kernels are made of -- TMA
(UTMALDG / UTMASTG), byte dot products (IDP), warp reductions (REDUX), three-input min/max (VIMNMX3), FP64 -- and how large
they are.  Runs without a GPU.

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.md
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'uav-airvision_b200', 'lib', 'libavb.so')
MARK = ['UTMALDG', 'UTMASTG', 'SYNCS', 'IDP', 'REDUX', 'VIMNMX3', 'VIMNMX', 'SHFL', 'PRMT', 'LOP3', 'IMAD', 'LDG', 'STG', 'LDS', 'STS',
        'ATOMS', 'ATOMG', 'RED', 'BAR', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'I2F', 'F2I', 'VOTE']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    demangle = subprocess.run(['c++filt'] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print('# SASS opcode histogram of `uav-airvision_b200/lib/libavb.so` (sm_100a, `cuobjdump -sass`)\n')
    print('Static instruction counts per kernel (not executed counts).  `UTMALDG` / `UTMASTG` = TMA tensor load / store '
          '(`cp.async.bulk.tensor`), `SYNCS` = mbarrier, `IDP` = dp4a / dp2a, `REDUX` = warp reduce, `VIMNMX3` = three-input '
          'integer min/max.\n')
    print('| kernel | total | ' + ' | '.join(MARK) + ' | other top opcodes |')
    print('|---|---|' + '---|' * (len(MARK) + 1))
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r'\(.*', '', dn).replace('void ', '')
        rest = [(k, v) for k, v in c.most_common() if k not in MARK][:4]
        print(f'| `{short}` | {sum(c.values())} | ' + ' | '.join(str(c.get(k, 0)) for k in MARK) + ' | ' +
              ', '.join(f'{k} {v}' for k, v in rest) + ' |')


if __name__ == '__main__':
    main()
