"""Top stalled SASS instructions of an .ncu-rep (source page): python tools/ncu_src_top.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
tot = sum(int(r[idx['# Samples']]) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('total samples', tot, 'warp instr', sum(int(r[idx['Instructions Executed']]) for r in data))
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:n]:
    s = int(r[idx['# Samples']])
    dom = sorted(((int(r[idx[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{s:6d} {100*s/tot:5.1f}% exec={r[idx['Instructions Executed']]:>8} {r[idx['Source']].strip()[:64]:64s} {dom}")
print({h[6:]: sum(int(r[idx[h]]) for r in data) for h in stalls})
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
keys = ('gpu__time_duration.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active')
for nme, u, v in zip(rows[0], rows[1], rows[2]):
    if nme in keys:
        print(nme, u, v)
