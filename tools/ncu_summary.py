#!/usr/bin/env python3
"""Print the metrics we track from an .ncu-rep (raw page) -- used to write profiles/*.md."""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
    r'^(dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum|gpu__dram_throughput\.avg\.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|'
    r'l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|launch__(registers_per_thread$|grid_size|block_size|shared_mem_per_block_dynamic)|'
    r'sm__inst_executed_pipe_(alu|fma|lsu|fmaheavy|fmalite|xu|uniform|adu|cbu)\.sum\.pct_of_peak_sustained_active|sm__pipe_fma(heavy|lite)?_cycles_active\.avg\.pct_of_peak_sustained_elapsed|'
    r'sm__throughput\.avg\.pct|sm__warps_active\.avg\.pct|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum$|'
    r'smsp__issue_active\.avg\.pct|sm__cycles_elapsed\.max|lts__t_bytes\.sum$|lts__throughput\.avg\.pct|smsp__sass_thread_inst_executed_op_(fadd|fmul|ffma)_pred_on\.sum\.per_cycle_elapsed|l1tex__throughput\.avg\.pct)')
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for v in rows[2:]:
    print("## kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
    print("| metric | unit | value |\n|---|---|---|")
    for a, b, c in zip(h, u, v):
        if pat.search(a):
            print(f"| {a} | {b} | {c} |")
