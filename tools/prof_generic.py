"""Tiny ncu driver for the runtime-topology (generic) kernels: python tools/prof_generic.py <mode> [envs] [T] [nodes_per_echelon...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc
mode = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
kw = {}
if len(sys.argv) > 4:
    kw["nodes_per_echelon"] = [int(x) for x in sys.argv[4:8]]
env = gsc.make("supplychain-ntom-v0", num_envs=n, mode=mode, seed=0, auto_reset=True, force_generic=True, **kw)
env.reset()
dt = env._io_dtype
acts = torch.empty((T, n, env.n_actions), device="cuda", dtype=dt).uniform_(-1, 1)
obs = torch.empty((T, n, env.n_obs), device="cuda", dtype=dt); rew = torch.empty((T, n), device="cuda", dtype=dt)
done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(2):
    env.rollout(T, acts, out=(obs, rew, done))
ev[0].record()
for _ in range(3):
    env.rollout(T, acts, out=(obs, rew, done))
ev[1].record(); torch.cuda.synchronize()
print("ok %.3f ms per launch, %.3e env-steps/s" % (ev[0].elapsed_time(ev[1]) / 3, n * T * 3 / (ev[0].elapsed_time(ev[1]) * 1e-3)))
