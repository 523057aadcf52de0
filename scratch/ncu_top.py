import csv, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; vals = rows[2]
for k in ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed']:
    if k in hdr: print(k, '=', vals[hdr.index(k)], rows[1][hdr.index(k)])
src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]
iS = hdr.index('Source'); iN = hdr.index('# Samples'); iE = hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for idx, r in enumerate(rows[2:]):
    try: n = int(r[iN])
    except: continue
    data.append((n, idx, r))
tot = sum(n for n, _, _ in data); print('total samples', tot)
for n, idx, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = {hdr[i][6:]: int(r[i]) for i in stall if r[i] not in ('', '0') and int(r[i]) > n * 0.1}
    print(f"{n:7d} line{idx:5d} exec {r[iE]:>9s}  {r[iS][:70]:70s} {st}")
