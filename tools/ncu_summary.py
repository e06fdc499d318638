"""Summarise an ncu --set full report (raw page CSV) into the handful of metrics we track.
    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv"""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum',
        'lts__t_requests_srcunit_tex_op_red.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '') for r in data]
    print('| metric | ' + ' | '.join(names) + ' | unit |')
    print('|---|' + '---|' * (len(names) + 1))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'| {k} | ' + ' | '.join(r[i] for r in data) + f' | {units[i]} |')
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    for k in stall:
        i = hdr.index(k)
        vals = [float(r[i] or 0) for r in data]
        if max(vals) >= 0.15:
            print(f"| stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | "
                  + ' | '.join(f'{v:.2f}' for v in vals) + ' | warps/issue |')


if __name__ == '__main__':
    main(sys.argv[1])
