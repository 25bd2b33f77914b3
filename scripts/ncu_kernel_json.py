"""One kernel of an .ncu-rep -> the few numbers bench.py quotes (python scripts/ncu_kernel_json.py rep.ncu-rep regex out.json).
Per launch: dram bytes read + written, duration, SM clock, tensor-pipe / issue / LSU utilisation."""
import csv
import json
import subprocess
import sys

rep, regex, out = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '-k', 'regex:' + regex], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, units, r = rows[0], rows[1], rows[-1]


def get(name, scale=1.0):
    if name not in h:
        return None
    i = h.index(name)
    try:
        v = float(r[i].replace(',', ''))
    except ValueError:
        return None
    u = units[i]
    mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0,
            'Ghz': 1e9, 'Mhz': 1e6}.get(u, 1.0)
    return v * mult * scale


d = dict(kernel=r[h.index('Kernel Name')][:80], ncu_file=rep.split('/')[-1],
         dram_bytes_read=get('dram__bytes_read.sum'), dram_bytes_write=get('dram__bytes_write.sum'),
         duration_ms=get('gpu__time_duration.sum', 1e3), sm_mhz=get('smsp__cycles_elapsed.avg.per_second', 1e-6),
         tensor_pipe_active_pct_of_elapsed=get('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed')
         or get('sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_elapsed')
         or get('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
         tensor_pipe_cycles_active_realtime_pct=get('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
         issue_active_pct=get('smsp__issue_active.avg.pct_of_peak_sustained_elapsed'),
         lsu_data_pipe_pct=get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
         l2_hit_pct=get('lts__t_sector_hit_rate.pct'))
if d['dram_bytes_read'] is not None and d['dram_bytes_write'] is not None:
    d['dram_bytes_per_launch'] = d['dram_bytes_read'] + d['dram_bytes_write']
tens = [n for n in h if 'tensor' in n and 'pct' in n]
d['tensor_metrics'] = {n: r[h.index(n)] for n in tens[:12]}
json.dump(d, open(out, 'w'), indent=1)
print(json.dumps(d, indent=1))
