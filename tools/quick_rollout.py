#!/usr/bin/env python
"""Kernel A/B timing of the fused rollout (CUDA events, inputs resident in HBM), for experiment builds:
    SCG_LIB=gym_supplychain_b200/libscg_<tag>.so python tools/quick_rollout.py [--policy random|sendall]
Prints one line per (env id, mode): ms per launch, env-steps/s, algorithmic GB/s and the fraction of the measured HBM peak,
plus a checksum of the rewards so that two builds can be compared on identical inputs."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc

ap = argparse.ArgumentParser()
ap.add_argument("--policy", default="random")
ap.add_argument("--cases", default="ntom:mixed,ntom:fast_f32,2perstage:mixed,2perstage:fast_f32,linear:mixed,linear:fast_f32")
ap.add_argument("--iters", type=int, default=6)
args = ap.parse_args()
PEAK = 6549.1
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
SIZES = {"ntom": (1 << 20, 128), "2perstage": (1 << 20, 128), "linear": (1 << 22, 128), "oneonen": (1 << 21, 128), "v0": (1 << 20, 128), "multiproduct": (1 << 20, 128)}
IDS = {"v0": "supplychain-v0", "multiproduct": "sc-2perstage-multiproduct-v0"}
for case in args.cases.split(","):
    name, mode = case.split(":")
    n, T = SIZES[name]
    env = gsc.make(IDS.get(name, "supplychain-%s-v0" % name), num_envs=n, mode=mode, seed=0, auto_reset=True)
    env.reset()
    A, O = env.n_actions, env.n_obs
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    acts = torch.empty((T, n, A), device="cuda").uniform_(-1, 1, generator=g)
    if args.policy == "sendall":          # adversarial for the fast paths: everything shipped to one destination, full supply
        acts.fill_(1.0)
    obs = torch.empty((T, n, O), device="cuda"); rew = torch.empty((T, n), device="cuda")
    done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        env.rollout(T, acts, out=(obs, rew, done))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        env.rollout(T, acts, out=(obs, rew, done))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.iters
    B = 4 * A + 4 * O + 5
    v = n * T / (ms * 1e-3)
    print("%-10s %-9s %-8s %8.3f ms  %.3e steps/s  %7.1f GB/s  frac %.3f  rew_sum %.9e obs_sum %.9e" %
          (name, mode, args.policy, ms, v, v * B / 1e9, v * B / 1e9 / PEAK, rew.double().sum().item(), obs.double().sum().item()), flush=True)
    del env, acts, obs, rew, done
    torch.cuda.empty_cache()
