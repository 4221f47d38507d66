"""Summarise an `ncu --set full` report (.ncu-rep) per captured launch: the figures DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ('duration_ms', 'gpu__time_duration.sum'),
    ('grid', 'launch__grid_size'), ('block', 'launch__block_size'), ('regs/thread', 'launch__registers_per_thread'),
    ('dram_read_MB', 'dram__bytes_read.sum'), ('dram_write_MB', 'dram__bytes_write.sum'),
    ('dram_throughput_%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
    ('l2_hit_%', 'lts__t_sector_hit_rate.pct'), ('l2_throughput_%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
    ('sm_throughput_%', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
    ('issue_active_%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
    ('warps_active_%', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
    ('tensor_pipe_active_%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
    ('tensor_pipe_subpipe_%', 'sm__pipe_tensor_subpipe_cycles_active.avg.pct_of_peak_sustained_active'),
    ('xu_pipe_%', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
    ('smem_wavefronts_%', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
    ('warp_insts', 'smsp__inst_executed.sum'),
]


def main(path, out=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f'# {path}: ncu --set full --clock-control none (one replayed launch per entry; durations are under the profiler)']
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        lines.append(name[:140])
        for label, m in METRICS:
            cands = [i for i, h in enumerate(hdr) if h == m]
            if not cands:
                continue
            i = cands[0]
            v, u = r[i], units[i]
            try:
                f = float(v.replace(',', ''))
                if label.endswith('_MB'):
                    f *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1)
                    u = 'MB'
                if label == 'duration_ms':
                    f *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}.get(u, 1)
                    u = 'ms'
                v = f'{f:,.3f}'.rstrip('0').rstrip('.')
            except ValueError:
                pass
            lines.append(f'    {label:24s} {v} {u if u not in ("%",) else ""}'.rstrip())
    text = '\n'.join(lines)
    print(text)
    if out:
        open(out, 'w').write(text + '\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
