#!/usr/bin/env python
"""Mid-size layered chains (run-time instantiated one-thread-per-env kernel): fused-rollout timing, CUDA events, inputs in HBM.
    python tools/midsize_chains.py [--precompile]     (--precompile: NVRTC only, no GPU; SCG_JIT_WARPS selects the block size)"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHAINS = [dict(nodes_per_echelon=[2, 3, 2, 4], num_products=3, avg_leadtime=3, max_leadtime=3),
          dict(nodes_per_echelon=[3, 3, 3, 3], num_products=3),
          dict(nodes_per_echelon=[4, 4, 4, 4], num_products=2),
          dict(nodes_per_echelon=[2, 2, 2, 2], num_products=2),
          dict(nodes_per_echelon=[2, 3, 2, 4], num_products=1, avg_leadtime=3, max_leadtime=3),
          dict(nodes_per_echelon=[3, 3, 3, 3], num_products=3, stochastic_leadtimes=True, avg_leadtime=2, max_leadtime=4)]

ap = argparse.ArgumentParser()
ap.add_argument("--precompile", action="store_true")
ap.add_argument("--envs", type=int, default=1 << 19)
ap.add_argument("--T", type=int, default=64)
ap.add_argument("--modes", default="mixed,fast_f32")
args = ap.parse_args()

if args.precompile:
    from gym_supplychain_b200 import FlatSpec, presets
    for kw in CHAINS:
        info, k = presets.n_per_stage(**kw)
        for mode in args.modes.split(","):
            print(kw, mode, FlatSpec(info, **k).precompile(mode), flush=True)
    sys.exit(0)

import torch
import gym_supplychain_b200 as gsc
PEAK = 6549.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for kw in CHAINS:
    for mode in args.modes.split(","):
        n, T = args.envs, args.T
        env = gsc.make("supplychain-ntom-v0", num_envs=n, mode=mode, seed=0, auto_reset=True, **kw)
        env.reset()
        A, O = env.n_actions, env.n_obs
        acts = torch.empty((T, n, A), device="cuda").uniform_(-1, 1)
        obs = torch.empty((T, n, O), device="cuda"); rew = torch.empty((T, n), device="cuda")
        done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        for _ in range(2):
            env.rollout(T, acts, out=(obs, rew, done))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            env.rollout(T, acts, out=(obs, rew, done))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        B = 4 * A + 4 * O + 5
        v = n * T / (ms * 1e-3)
        print("%s %s variant %d %.3f ms %.3e steps/s frac %.3f rew_sum %.9e" % (kw, mode, env.scg_info.kernel_variant, ms, v, v * B / 1e9 / PEAK,
                                                                             rew.double().sum().item()), flush=True)
        del env, acts, obs, rew, done
        torch.cuda.empty_cache()
