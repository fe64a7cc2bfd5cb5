"""Summarise an Nsight Compute report into a small markdown table (run where `ncu` is installed;
no GPU needed):   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram rd'),
    ('dram__bytes_write.sum', 'dram wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('smsp__inst_executed.sum', 'warp inst'),
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f'# ncu summary of `{path}` (`--set full --clock-control none`, cold-cache serialised replays)\n')
    print('| kernel | ' + ' | '.join(n for _, n in METRICS) + ' |')
    print('|---|' + '---|' * len(METRICS))
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
        cells = []
        for m, _ in METRICS:
            if m in idx:
                v, u = r[idx[m]], units[idx[m]]
                try:
                    f = float(v)
                    v = f'{f:.1f}' if f < 1000 else f'{f:.0f}'
                except ValueError:
                    pass
                cells.append(f'{v} {u}'.strip())
            else:
                cells.append('-')
        print(f'| `{name}` | ' + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    main(sys.argv[1])
