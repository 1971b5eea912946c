#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics, opcode mix and instructions per source line.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [warp_periods_per_launch] > profiles/<name>_summary.txt"""
import csv, io, re, subprocess, sys
from collections import Counter

rep = sys.argv[1]
warp_periods = float(sys.argv[2]) if len(sys.argv) > 2 else None
KEEP = re.compile(r"dram__bytes_(read|write)\.sum$|gpu__dram_throughput.avg.pct|gpu__time_duration.sum|launch__registers_per_thread|"
                  r"launch__occupancy_limit|launch__waves|sm__warps_active.avg.per_cycle_active|smsp__issue_active.avg.pct|"
                  r"smsp__inst_executed.sum$|sm__inst_executed.avg.per_cycle_elapsed|smsp__average_warps_issue_stalled_.*per_issue_active|"
                  r"sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu).avg.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
                  r"smsp__warps_eligible.avg.per_cycle_active|smsp__thread_inst_executed_per_inst_executed.ratio|sm__cycles_elapsed.avg$|"
                  r"launch__shared_mem_per_block|launch__block_size|launch__grid_size|smsp__inst_executed_op_shared|lts__t_bytes.sum$|"
                  r"local_(load|store)|smsp__inst_executed_op_local")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
print("kernel:", vals[hdr.index("Kernel Name")])
for h, u, v in zip(hdr, units, vals):
    if KEEP.search(h):
        print("%-95s %s %s" % (h, v, u))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(k for k, r in enumerate(rows) if "Source" in r)          # row 0 is the kernel name
h = rows[hi]
rows = rows[hi:]
i_src, i_ex = h.index("Source"), h.index("Instructions Executed")
i_smp = h.index("# Samples") if "# Samples" in h else None
ops, smp, tot = Counter(), Counter(), 0
for r in rows[1:]:
    try:
        ex = int(r[i_ex])
    except Exception:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_src])
    op = m.group(2) if m else "?"
    ops[op] += ex; tot += ex
    if i_smp is not None:
        try: smp[op] += int(r[i_smp])
        except Exception: pass
print("\nSASS rows: %d; warp-instructions executed: %d%s" % (len(rows) - 1, tot,
      ("; per warp-period: %.1f" % (tot / warp_periods)) if warp_periods else ""))
ts = sum(smp.values()) or 1
print("opcode      %inst  %stall-samples" + ("  per-warp-period" if warp_periods else ""))
for op, c in ops.most_common(24):
    print("%-10s %6.2f %6.2f" % (op, 100.0 * c / tot, 100.0 * smp[op] / ts) + (("  %8.1f" % (c / warp_periods)) if warp_periods else ""))
