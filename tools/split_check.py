#!/usr/bin/env python
"""Node-parallel (split) kernel: parity against the C oracle and timing, for several (lanes per env, min blocks) shapes.
    python tools/split_check.py [--quick]            (GPU box; SCG_SPLIT_G / SCG_SPLIT_MINB select the instantiation)"""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gym_supplychain_b200 as gsc
from gym_supplychain_b200 import presets
from oracle.c_oracle import CSupplyChainOracle
from test_gpu_c_oracle import _inputs, _oracle_run, _gpu_run

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="4:2,4:3,8:3,2:2,0:0")
ap.add_argument("--shapes", default="[5,4,7,10];10")
ap.add_argument("--modes", default="fast_f32,mixed")
ap.add_argument("--envs", type=int, default=1 << 18)
ap.add_argument("--T", type=int, default=32)
args = ap.parse_args()
PEAK = 6549.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for shape_s in args.shapes.split(";"):
    shape = eval(shape_s)
    info, okw = presets.n_per_stage(nodes_per_echelon=shape)
    cc = CSupplyChainOracle(info, **okw)
    n_par, T_par = 96, 120
    dem, lts, a = _inputs(np.random.RandomState(6), cc, n_par, T_par, 0, 400)
    cc2 = CSupplyChainOracle(info, **dict(okw, total_time_steps=T_par))
    ref = _oracle_run(cc2, dem[:, :T_par + 1], lts, a, canonical=True)
    for mode in args.modes.split(","):
        for cfg in args.configs.split(","):
            g, mb = cfg.split(":")
            os.environ["SCG_SPLIT_G"] = g
            if g == "d":                                   # "d:0" = the planned lane count and launch shape
                os.environ.pop("SCG_SPLIT_G")
            os.environ["SCG_SPLIT_MINB"] = mb
            if mb == "0":
                os.environ.pop("SCG_SPLIT_MINB")
            env = gsc.make("supplychain-ntom-v0", num_envs=n_par, mode=mode, auto_reset=False, nodes_per_echelon=shape, total_time_steps=T_par)
            got = _gpu_run(env, dem[:, :T_par + 1], lts, a)
            rel = np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1.0)
            oerr = np.abs(got[0] - ref[0]).max()
            kv = env.scg_info.kernel_variant
            del env
            n, T = args.envs, args.T
            env = gsc.make("supplychain-ntom-v0", num_envs=n, mode=mode, seed=0, auto_reset=True, nodes_per_echelon=shape)
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
            for _ in range(3):
                env.rollout(T, acts, out=(obs, rew, done))
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            B = 4 * A + 4 * O + 5
            v = n * T / (ms * 1e-3)
            print("shape %-12s %-8s G=%s minb=%s variant %3d | reward rel err max %.2e done_ok %s obs err %.2e | %8.3f ms %.3e steps/s frac %.3f rew_sum %.6e"
                  % (shape_s, mode, g, mb, kv, rel.max(), np.array_equal(got[2], ref[2]), oerr, ms, v, v * B / 1e9 / PEAK,
                     rew.double().sum().item()), flush=True)
            del env, acts, obs, rew, done
            torch.cuda.empty_cache()
