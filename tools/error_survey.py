"""Measured reward / obs error of the fp32-I/O modes against the C oracle, whole batches, whole episodes (GPU box)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gym_supplychain_b200 as gsc
from gym_supplychain_b200 import presets
from oracle.c_oracle import CSupplyChainOracle
from test_gpu_c_oracle import _inputs, _oracle_run, _gpu_run

cases = [("supplychain-ntom-v0", presets.n_per_stage(nodes_per_echelon=3), {}, (0, 400), 2048),
         ("supplychain-2perstage-v0", presets.two_per_stage(), {}, (10, 20), 2048),
         ("supplychain-linear-v0", presets.n_per_stage(nodes_per_echelon=1, num_products=1), {}, (0, 400), 4096)]
for env_id, (info, okw), kw, (lo, hi), n in cases:
    cc = CSupplyChainOracle(info, **okw)
    dem, lts, a = _inputs(np.random.RandomState(7), cc, n, 360, lo, hi)
    ref = _oracle_run(cc, dem, lts, a, canonical=True)
    for mode in ("mixed", "fast_f32"):
        env = gsc.make(env_id, num_envs=n, mode=mode, auto_reset=False, **kw)
        got = _gpu_run(env, dem, lts, a)
        rel = np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1.0)
        ret = np.abs(got[1].sum(0) - ref[1].sum(0)) / np.abs(ref[1].sum(0))
        print("%-26s %-9s envs %5d  reward rel err: max %.2e  p99.9 %.2e  mean %.2e | episode return rel err max %.2e | obs abs err max %.2e"
              % (env_id, mode, n, rel.max(), np.quantile(rel, 0.999), rel.mean(), ret.max(), np.abs(got[0] - ref[0]).max()), flush=True)
