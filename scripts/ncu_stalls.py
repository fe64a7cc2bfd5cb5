"""Aggregate warp-stall samples of one kernel by reason and by opcode.
    python scripts/ncu_stalls.py report.ncu-rep <kernel regex>"""
import csv, io, subprocess, sys, collections

def main(path, regex):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', f'regex:{regex}',
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
    hdr = rows[hi]
    i_src, i_s = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
    stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
    by_op, by_reason = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= i_s:
            continue
        try:
            s = int(r[i_s] or 0)
        except ValueError:
            continue
        toks = r[i_src].strip().split()
        op = (toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?'))
        by_op[op] += s
        for i, c in stall_cols:
            try:
                by_reason[c] += int(r[i] or 0)
            except (ValueError, IndexError):
                pass
    tot = sum(by_op.values()) or 1
    print('by reason:', [(k, f'{100 * v / tot:.0f}%') for k, v in by_reason.most_common(8)])
    print('by opcode:', [(k, f'{100 * v / tot:.0f}%') for k, v in by_op.most_common(12)])

if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
