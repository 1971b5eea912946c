#!/usr/bin/env python
"""Instructions executed per CUDA source line straight from an .ncu-rep captured with --import-source on (works for the
run-time instantiated kernels too, no object file needed): ncu's own CUDA<->SASS correlation.
    python tools/ncu_cuda_lines.py rep.ncu-rep warp_periods [top]"""
import csv, io, os, subprocess, sys
rep, wp = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
per, cur_file, hdr, kname = [], None, None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if len(r) == 2 and r[0] == "Function Name":
        kname = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; i_ex, i_smp = r.index("Instructions Executed"), r.index("# Samples"); continue
    if hdr and r and r[0].isdigit():
        try:
            per.append((int(r[i_ex]), int(r[i_smp]), cur_file, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(p[0] for p in per); ts = sum(p[1] for p in per) or 1
print("kernel %s\nwarp-instructions per warp-period by CUDA source line (total %.1f; stall-sample share in the third column)" % ((kname or "?")[:110], tot / wp))
for ex, smp, f, ln, text in sorted(per, reverse=True)[:top]:
    print("%7.1f %5.1f%% %5.1f%%  %s:%4d  %s" % (ex / wp, 100.0 * ex / tot, 100.0 * smp / ts, f, ln, text[:100]))
