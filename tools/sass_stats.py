#!/usr/bin/env python
"""Static SASS statistics of one kernel of a cubin (no GPU): instruction count, opcode histogram, and the instructions that
tell how the bulk (TMA) copies are addressed -- R2UR / BRA.U.ANY are the register-to-uniform loops the compiler puts around a
bulk copy whose address it cannot prove warp-uniform.

    python tools/sass_stats.py <file.cubin | lib.so> <kernel-name regex> [top]
"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter


def cubins(path):
    if path.endswith(".cubin"):
        return [path]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=tmp, capture_output=True)
    return sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin"))


def functions(cubin):
    dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout
    cur, f = None, {}
    for l in dis.splitlines():
        m = re.match(r"\.text\.(\S+):", l)
        if m:
            cur = m.group(1); f[cur] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m and cur:
            f[cur].append(m.group(2))
    return f


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    for cb in cubins(path):
        for name, ins in functions(cb).items():
            if not re.search(pat, name):
                continue
            ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in ins)
            print("%s\n  %d instructions (%.1f KB)  R2UR %d  BRA.U.ANY %d  UBLKCP %d  UBLKPF %d  BSSY %d  BRA %d" % (
                name[:110], len(ins), len(ins) * 16 / 1024.0, ops["R2UR"], sum(1 for t in ins if "BRA.U.ANY" in t), ops["UBLKCP"],
                ops["UBLKPF"], ops["BSSY"], ops["BRA"]))
            print("  " + "  ".join("%s %d" % kv for kv in ops.most_common(top)))


if __name__ == "__main__":
    main()
