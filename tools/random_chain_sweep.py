"""Random layered chains (tests/test_gpu_c_oracle.py generator) through the run-time instantiated kernels against the C
oracle, whole batches: python tools/random_chain_sweep.py <first seed> <last seed + 1>  (GPU box)."""
import os, sys, traceback
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gym_supplychain_b200 as gsc
from oracle.c_oracle import CSupplyChainOracle
from test_gpu_c_oracle import _random_layered_chain, _inputs, _oracle_run, _gpu_run
lo, hi = int(sys.argv[1]), int(sys.argv[2])
tight = len(sys.argv) > 3 and sys.argv[3] == "tight"      # factories that CAN be ship-clipped (negative capacities, F4a)
for seed in range(lo, hi):
    rs = np.random.RandomState(3000 + seed)
    info, kw, cnt = _random_layered_chain(rs, tight_factory_ship=tight)
    okw = {k: v for k, v in kw.items() if k not in ("seed", "build_info")}
    cc = CSupplyChainOracle(info, **okw)
    n, T = 96, kw["total_time_steps"]
    dem, lts, a = _inputs(rs, cc, n, T, kw["demand_range"][0], kw["demand_range"][1], wide=(seed % 2 == 1))
    if lts is not None:
        lts = rs.randint(1, kw["max_leadtime"] + 1, size=lts.shape).astype(np.int32)
    ref = _oracle_run(cc, dem, lts, a, canonical=True)
    res = []
    for mode in ("mixed", "fast_f32"):
        env = gsc.SupplyChainEnv(info, num_envs=n, mode=mode, auto_reset=False, **okw)
        got = _gpu_run(env, dem, lts, a)
        rel = np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1.0)
        res.append("%s v%d rel %.2e obs %.2e done %s bad env-steps %d/%d" % (mode, env.scg_info.kernel_variant, rel.max(), np.abs(got[0] - ref[0]).max(),
                                                                          np.array_equal(got[2], ref[2]), int((rel > 1e-4).sum()), rel.size))
    print("seed", seed, cnt, "P", kw["num_products"], "L", kw["avg_leadtime"], kw["max_leadtime"], "stoch" if kw["stochastic_leadtimes"] else "const", "|", " | ".join(res), flush=True)
