"""Top stalled SASS instructions of one kernel from an Nsight Compute report (source page).
    python scripts/ncu_hot.py report.ncu-rep <kernel regex> [top]"""
import csv, io, subprocess, sys

def main(path, regex, top=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', f'regex:{regex}',
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
    hdr = rows[hi]
    i_src, i_s, i_ex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    data = []
    for idx, r in enumerate(rows[hi + 1:]):
        if len(r) <= i_ex:
            continue
        try:
            data.append((int(r[i_s] or 0), idx, r[i_src].strip(), int(r[i_ex] or 0)))
        except ValueError:
            pass
    tot = sum(d[0] for d in data) or 1
    print(rows[0][1][:100] if rows and len(rows[0]) > 1 else '')
    print('samples', tot, '| SASS instructions', len(data), '| executed', sum(1 for d in data if d[3] > 0),
          '| warp-instr executed', sum(d[3] for d in data))
    for s, idx, src, ex in sorted(data, reverse=True)[:top]:
        print(f'{s:6d} {100 * s / tot:5.1f}%  #{idx:5d} ex={ex:8d}  {src[:100]}')

if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
