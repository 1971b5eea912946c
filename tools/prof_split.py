"""Tiny driver for ncu: a few fused-rollout launches of a large layered chain (node-parallel split kernel).
    python tools/prof_split.py "[5,4,7,10]" fast_f32 [envs] [T]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc
shape, mode = eval(sys.argv[1]), sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 18
T = int(sys.argv[4]) if len(sys.argv) > 4 else 32
env = gsc.make("supplychain-ntom-v0", num_envs=n, mode=mode, seed=0, auto_reset=True, nodes_per_echelon=shape)
env.reset()
acts = torch.empty((T, n, env.n_actions), device="cuda").uniform_(-1, 1)
obs = torch.empty((T, n, env.n_obs), device="cuda"); rew = torch.empty((T, n), device="cuda")
done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
for _ in range(4):
    env.rollout(T, acts, out=(obs, rew, done))
torch.cuda.synchronize()
print("ok", env.scg_info.kernel_variant)
