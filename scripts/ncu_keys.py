"""Key metrics of every launch in an .ncu-rep (read here with `ncu -i ... --page raw --csv`): duration, SM clock, tensor pipe,
DRAM bytes, L1 data-pipe shares, issue utilisation.  usage: ncu_keys.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = [('gpu__time_duration.sum', 'ms'), ('sm__cycles_elapsed.avg.per_second', 'GHz'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe % of elapsed'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % of active'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'L1 data pipe: LSU wavefronts %'),
        ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'L1 data pipe: tensor-core operand wavefronts %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('launch__registers_per_thread', 'registers'), ('launch__occupancy_limit_shared_mem', 'CTAs/SM (smem limit)')]
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print(r[ki][:110])
    for k, label in want:
        if k in hdr:
            i = hdr.index(k)
            print('    %-50s %s %s' % (label, r[i], rows[1][i]))
