"""Summarise an ncu report for profiles/ (run here, no GPU needed):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title line" > profiles/<name>.txt
"""
import csv
import io
import re
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keep = re.compile(
    r"^(dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed|"
    r"l1tex__m_xbar2l1tex_read_bytes\.sum|launch__(block_size|grid_size|occupancy_limit_registers|occupancy_limit_shared_mem|registers_per_thread|"
    r"shared_mem_per_block_dynamic|waves_per_multiprocessor)|lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__cycles_active\.(avg|max|min)|sm__cycles_elapsed\.avg(\.per_second)?|sm__pipe_(alu|fma)_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__warps_(active|eligible)\.avg\.per_cycle_active|sm__icc_requests\.sum|"
    r"sm__icc_request_hit_rate\.pct|smsp__sass_inst_executed_op_local_(ld|st)\.sum)$")
print(f"# {title}")
print(f"# kernel: {m.get('Kernel Name', ('?',))[0][:140]}")
for k in sorted(m):
    if keep.match(k):
        print(f"{k} = {m[k][0]} {m[k][1]}")
try:
    rd = float(m["dram__bytes_read.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_read.sum"][1]]
    wr = float(m["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_write.sum"][1]]
    print(f"# DRAM traffic per launch = {rd + wr:.0f} B")
except Exception as e:  # noqa: BLE001
    print("# traffic: n/a", e)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    data = [r for r in rows[2:] if len(r) == len(h)]
    tot_s = sum(int(r[ix["# Samples"]] or 0) for r in data)
    tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
    print(f"# source page: {tot_i} warp instructions executed, {tot_s} stall samples; top 12 instructions by samples:")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
        print(f"#   {r[ix['# Samples']]:>6} samples  x {r[ix['Instructions Executed']]:>8}  {r[ix['Source']].strip()[:90]}")
