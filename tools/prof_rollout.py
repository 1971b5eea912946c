"""Tiny driver for ncu: a few fused-rollout launches of one (env id, mode).  python tools/prof_rollout.py ntom mixed [envs] [T]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc
name, mode = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
T = int(sys.argv[4]) if len(sys.argv) > 4 else 128
env = gsc.make("supplychain-v0" if name == "v0" else "supplychain-%s-v0" % name, num_envs=n, mode=mode, seed=0, auto_reset=True)
env.reset()
acts = torch.empty((T, n, env.n_actions), device="cuda").uniform_(-1, 1)
obs = torch.empty((T, n, env.n_obs), device="cuda"); rew = torch.empty((T, n), device="cuda")
done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
for _ in range(4):
    env.rollout(T, acts, out=(obs, rew, done))
torch.cuda.synchronize()
print("ok")
