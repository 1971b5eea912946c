"""Kernel A/B timing of the beer-game rollout (8M envs, one 35-week episode per launch): SCG_LIB=... python tools/quick_bg.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc
n, T = 8 << 20, 35
for env_id, kw in (("beergame-v2", {}), ("beergame-v0", {}), ("beergame-v2", {"customer_demand": (0, 16), "shipment_delays": (0, 5)})):
    env = gsc.make(env_id, num_envs=n, seed=0, auto_reset=True, **kw)
    env.reset()
    acts = torch.randint(0, 30, (T, n, 4), dtype=torch.int32, device="cuda")
    obs = torch.empty((T, n, 4), dtype=torch.int32, device="cuda"); rew = torch.empty((T, n), dtype=torch.int32, device="cuda")
    done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        env.rollout(T, acts, out=(obs, rew, done))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        env.rollout(T, acts, out=(obs, rew, done))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    v = n * T / (ms * 1e-3)
    print("%-12s %-40s %7.3f ms  %.3e steps/s  %7.1f GB/s  frac %.3f  rew_sum %d" % (env_id, str(kw)[:40], ms, v, v * 37 / 1e9, v * 37 / 1e9 / 6549.1, int(rew.sum())), flush=True)
    del env, acts, obs, rew, done
    torch.cuda.empty_cache()
