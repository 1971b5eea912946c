"""Tiny driver for profiling the beer-game rollout kernel (ncu): 8M envs, one 35-week episode per launch.
    python tools/prof_bg.py [stochastic]      (stochastic: customer_demand=(0,16), shipment_delays=(0,5) from Philox)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc
n, T = 8 << 20, 35
kw = {"customer_demand": (0, 16), "shipment_delays": (0, 5)} if len(sys.argv) > 1 and sys.argv[1] == "stochastic" else {}
env = gsc.make("beergame-v2", num_envs=n, seed=0, auto_reset=True, **kw)
env.reset()
acts = torch.randint(0, 30, (T, n, 4), dtype=torch.int32, device="cuda")
obs = torch.empty((T, n, 4), dtype=torch.int32, device="cuda"); rew = torch.empty((T, n), dtype=torch.int32, device="cuda")
done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
for _ in range(5):
    env.rollout(T, acts, out=(obs, rew, done))
torch.cuda.synchronize()
print("ok")
