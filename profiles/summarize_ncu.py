"""Condense an .ncu-rep (ncu --set full) into the few lines the DESIGN / bench numbers cite.

    python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep > profiles/rN_ncu_x.txt
"""
import csv
import io
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_red.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__cycles_active.avg',
    'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_uniform.sum.pct_of_peak_sustained_active',
]
PREFIX = ('smsp__average_warps_issue_stalled_', 'sm__pipe_tensor', 'sm__inst_executed_pipe_tensor')


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('%-100s %s' % ('Kernel Name', d.get('Kernel Name', '?')))
        for name, unit in zip(hdr, units):
            if name in KEEP or (name.startswith(PREFIX) and ('per_issue_active' in name or 'pct_of_peak_sustained_active' in name)):
                print('%-100s %-12s %s' % (name, unit, d[name]))
        print()


if __name__ == '__main__':
    main(sys.argv[1])
