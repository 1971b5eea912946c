#!/usr/bin/env python
"""Instructions executed per CUDA source line for one kernel of an .ncu-rep: joins the SASS page of the report with
`nvdisasm -g` line info of the same build (object file given).   python tools/ncu_lines.py rep.ncu-rep obj.o warp_periods"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter
rep, obj, wp = sys.argv[1], sys.argv[2], float(sys.argv[3])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kname = rows[0][1]
hi = next(k for k, r in enumerate(rows) if "Source" in r)
h = rows[hi]; rows = rows[hi + 1:]
ia, ie, iss = h.index("Address"), h.index("Instructions Executed"), h.index("Source")
base = int(rows[0][ia], 16)
ex = {int(r[ia], 16) - base: (int(r[ie]), r[iss]) for r in rows if r[ie].isdigit()}
# mangled-name match: take the function whose instruction count equals the report's
tmp = tempfile.mkdtemp()
if obj.endswith(".cubin"):                      # a run-time instantiated kernel: the cubin of the JIT cache
    cubin = os.path.abspath(obj)
else:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
funcs, cur, line = {}, None, None
for l in dis.splitlines():
    m = re.match(r"\.text\.(\S+):", l)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m and cur: funcs[cur].append((int(m.group(1), 16), line, m.group(2)))
cands = [k for k, v in funcs.items() if len(v) == len(ex)]
want = re.sub(r"[^A-Za-z0-9]", "", kname.split("<")[0].split("::")[-1])
cands = [k for k in cands if want in k] or cands
if not cands:                                   # a cubin with one kernel (JIT cache): take it if the sizes are close
    cands = [k for k, v in funcs.items() if want in re.sub(r"[^A-Za-z0-9]", "", k) and abs(len(v) - len(ex)) <= 0.05 * len(ex)]
    if cands:
        print("WARNING: %d SASS rows in the report, %d in the cubin: line attribution is approximate" % (len(ex), len(funcs[cands[0]])))
f = funcs[cands[0]]
per = Counter()
for off, ln, txt in f:
    per[ln] += ex.get(off, (0, ""))[0]
tot = sum(per.values())
print("kernel %s\ninstructions per warp-period by CUDA source line (total %.1f)" % (cands[0][:80], tot / wp))
srccache = {}
for ln, c in per.most_common(60):
    if ln is None: continue
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym_supplychain_b200", "csrc", ln[0])
    if path not in srccache:
        try: srccache[path] = open(path).read().splitlines()
        except Exception: srccache[path] = []
    text = srccache[path][ln[1] - 1].strip()[:100] if 0 < ln[1] <= len(srccache[path]) else ""
    print("%7.1f %5.1f%%  %s:%4d  %s" % (c / wp, 100.0 * c / tot, ln[0], ln[1], text))
