"""SASS opcode summary of liblinkb200.so (cuobjdump -sass; no GPU needed): per kernel, the instructions
that prove the Blackwell paths are real -- UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit), STTM / LDTM
(tcgen05.st / ld), UBLKCP (cp.async.bulk), LDGSTS (cp.async), SYNCS (mbarrier), REDG / ATOMG (global
reductions / atomics, incl. the 128-bit CAS of the hash table), ACQBULK/PDL (griddepcontrol).
    python scripts/sass_summary.py > profiles/rNN_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'link_b200', 'liblinkb200.so')
KEYS = ['UTCHMMA', 'UTCBAR', 'STTM', 'LDTM', 'UTCATOMSWS', 'UBLKCP', 'LDGSTS', 'SYNCS', 'REDG', 'ATOMG', 'ATOM.',
        'ACQBULK', 'PREEXIT', 'BAR.SYNC', 'HMMA', 'FFMA', 'MUFU', 'SHFL', 'LDG', 'STG', 'LDS', 'STS']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
        if m and cur:
            op = m.group(1)
            per[cur]['_total'] += 1
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k] += 1
            if op.startswith('ATOMG') and '.128' in op or (op.startswith('ATOMG') and 'CAS' in op and '128' in line):
                per[cur]['ATOMG.CAS.128'] += 1
    demangle = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.splitlines()
    print(f'# SASS opcode summary of link_b200/liblinkb200.so (sm_100a; `cuobjdump -sass`), {len(per)} kernels\n')
    cols = ['UTCHMMA', 'UTCBAR', 'STTM', 'LDTM', 'UBLKCP', 'LDGSTS', 'SYNCS', 'REDG', 'ATOMG', 'ACQBULK', 'FFMA', 'MUFU', '_total']
    print('| kernel | ' + ' | '.join(c.strip('_') for c in cols) + ' |')
    print('|---|' + '---|' * len(cols))
    for (name, cnt), dm in zip(per.items(), demangle):
        short = re.sub(r'\(.*', '', dm).replace('void ', '')
        if len(short) > 70:
            short = short[:67] + '...'
        print(f'| `{short}` | ' + ' | '.join(str(cnt.get(c, 0)) for c in cols) + ' |')


if __name__ == '__main__':
    main()
