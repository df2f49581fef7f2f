"""Summarise one kernel of an .ncu-rep into a small CSV of the counters DESIGN.md / profiles/README.md cite."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
        'sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_global_loads',
        'sass__inst_executed_global_stores', 'smsp__warps_eligible.avg.per_cycle_active']


def main(rep, out=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(('kernel', d['Kernel Name'], ''))
        lines.append(('block', d['Block Size'], '')); lines.append(('grid', d['Grid Size'], ''))
        for k in KEYS:
            if k in d:
                lines.append((k, d[k], units[hdr.index(k)]))
        stalls = []
        for h in hdr:
            if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
                try:
                    stalls.append((float(d[h]), h))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True)[:6]:
            lines.append((h, '%.3f' % v, 'warps/issue'))
    w = csv.writer(open(out, 'w') if out else sys.stdout)
    w.writerow(['metric', 'value', 'unit'])
    w.writerows(lines)


if __name__ == '__main__':
    main(*sys.argv[1:])
