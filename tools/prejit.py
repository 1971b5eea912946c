#!/usr/bin/env python
"""Warm the run-time-instantiation cache (gym_supplychain_b200/_jit) on the CPU: every supply-chain golden case whose chain
has no pre-built compile-time kernel gets its kernels compiled with NVRTC now, so a fresh GPU box does not spend the first
minutes of the GPU tests compiling.  Needs no GPU.   python tools/prejit.py [--all-modes]"""
import concurrent.futures, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(job):
    name, mode, rand = job
    from golden_util import load
    from gym_supplychain_b200 import FlatSpec
    meta, _ = load("sc", name)
    kw = {k: v for k, v in meta["env_kwargs"].items() if k not in ("seed", "build_info")}
    t0 = time.time()
    ok = FlatSpec(meta["nodes_info"], **kw).precompile(mode, in_kernel_actions=rand)
    return name, mode, rand, ok, time.time() - t0


def main():
    from golden_util import sc_cases
    jobs = [(n, m, False) for n in sc_cases() for m in ("fast_f32", "mixed", "fast_f64")]
    with concurrent.futures.ProcessPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) - 1)) as ex:
        for name, mode, rand, ok, dt in ex.map(one, jobs):
            print("%-28s %-9s %s %5.1fs" % (name, mode, "compile-time kernel ready" if ok else "runtime-topology kernel only", dt), flush=True)


if __name__ == "__main__":
    main()
