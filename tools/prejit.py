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


def extra(job):
    """Shapes the GPU tests / tools instantiate that are not golden cases."""
    kw, mode, rand = job
    from gym_supplychain_b200 import FlatSpec, presets
    t0 = time.time()
    if kw == "supplychain-v0/maxL3":
        from gym_supplychain_b200.envs import SupplyChainMainDemo  # noqa: F401  (its default nodes_info)
        info, k = presets.main_demo(max_leadtime=3) if hasattr(presets, "main_demo") else (None, None)
        if info is None:
            return str(kw), mode, rand, False, 0.0
    else:
        info, k = presets.n_per_stage(**kw)
    ok = FlatSpec(info, **k).precompile(mode, in_kernel_actions=rand)
    return str(kw), mode, rand, ok, time.time() - t0


EXTRA = [dict(nodes_per_echelon=[2, 3, 2, 4], num_products=3, avg_leadtime=3, max_leadtime=3, total_time_steps=30)]
RANDOM_CHAINS = [(s, False) for s in range(8)] + [(0, True), (23, False), (27, False)]     # tests/test_gpu_c_oracle.py


def random_chain(job):
    """The random layered chains of tests/test_gpu_c_oracle.py (same generator, same seeds)."""
    (seed, tight), mode = job
    import numpy as np
    from gym_supplychain_b200 import FlatSpec
    from test_gpu_c_oracle import _random_layered_chain
    t0 = time.time()
    info, kw, cnt = _random_layered_chain(np.random.RandomState(3000 + seed), tight_factory_ship=tight)
    okw = {k: v for k, v in kw.items() if k not in ("seed", "build_info")}
    ok = FlatSpec(info, **okw).precompile(mode)
    return "random chain %d %s" % (seed, cnt), mode, False, ok, time.time() - t0


def split_variant(job):
    """Large chains (node-parallel split kernel) as tests/test_gpu_split.py instantiates them: default lane count, the other
    lane counts, and the one-thread-per-env kernel they are compared with (SCG_SPLIT_G=0)."""
    shape, mode, rand, g = job
    from gym_supplychain_b200 import FlatSpec, presets
    if g is None:
        os.environ.pop("SCG_SPLIT_G", None)
    else:
        os.environ["SCG_SPLIT_G"] = str(g)
    t0 = time.time()
    extra_kw = {}
    if isinstance(shape, tuple):                        # (nodes per echelon, products, lead time)
        shape, P, L = shape
        extra_kw = dict(num_products=P, avg_leadtime=L, max_leadtime=L)
    info, k = presets.n_per_stage(nodes_per_echelon=shape, **extra_kw)
    ok = FlatSpec(info, **k).precompile(mode, in_kernel_actions=rand)
    os.environ.pop("SCG_SPLIT_G", None)
    return "%s G=%s" % (shape, g), mode, rand, ok, time.time() - t0


SPLIT = [([5, 4, 7, 10], m, r, None) for m in ("mixed", "fast_f32") for r in (False, True)] + \
        [(10, m, False, None) for m in ("mixed", "fast_f32")] + \
        [([5, 4, 7, 10], "mixed", False, g) for g in (2, 8)] + [([5, 4, 7, 10], m, False, 0) for m in ("mixed", "fast_f32")] + \
        [(sh, m, False, None) for sh in (([6, 5, 9, 7], 2, 2), ([4, 6, 5, 8], 3, 2), ([7, 3, 6, 9], 2, 3)) for m in ("mixed", "fast_f32")]


def main():
    from golden_util import sc_cases
    jobs = [(n, m, False) for n in sc_cases() for m in ("fast_f32", "mixed", "fast_f64")]
    xjobs = [(kw, m, r) for kw in EXTRA for m in ("fast_f32", "mixed", "fast_f64") for r in (False, True)]
    with concurrent.futures.ProcessPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) - 1)) as ex:
        for name, mode, rand, ok, dt in ex.map(one, jobs):
            print("%-28s %-9s %s %5.1fs" % (name, mode, "compile-time kernel ready" if ok else "runtime-topology kernel only", dt), flush=True)
        for name, mode, rand, ok, dt in ex.map(random_chain, [(c, m) for c in RANDOM_CHAINS for m in ("mixed", "fast_f64")]):
            print("%-28s %-9s %s %5.1fs" % (name[:28], mode, "compile-time kernel ready" if ok else "runtime-topology kernel only", dt), flush=True)
        for name, mode, rand, ok, dt in ex.map(split_variant, SPLIT):
            print("%-28s %-9s %s %5.1fs" % (name[:28], mode, "compile-time kernel ready" if ok else "runtime-topology kernel only", dt), flush=True)
        for name, mode, rand, ok, dt in ex.map(extra, xjobs):
            print("%-28s %-9s %s %5.1fs" % (name[:28], mode, "compile-time kernel ready" if ok else "runtime-topology kernel only", dt), flush=True)


if __name__ == "__main__":
    main()
