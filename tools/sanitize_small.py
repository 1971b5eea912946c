"""Small driver that touches every kernel family once with awkward sizes (ragged tails, unaligned tiles), for
compute-sanitizer (one tool per gpurun call):   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc

torch.manual_seed(0)
for env_id in ("supplychain-linear-v0", "supplychain-2perstage-v0", "supplychain-ntom-v0", "supplychain-oneonen-v0",
               "sc-2perstage-multiproduct-v0", "supplychain-v0", "sc-2perstage-seasonal-v0"):
    for mode in ("fast_f32", "mixed", "fast_f64", "exact_f64"):
        for n in (1, 37, 200):                                  # 1 and 37: partial warps; 37*A*4 not 16-byte multiples
            env = gsc.make(env_id, num_envs=n, mode=mode, seed=1, auto_reset=True, total_time_steps=6)
            env.reset()
            a = (torch.rand(9, n, env.n_actions, device="cuda", dtype=env._io_dtype) * 2 - 1)
            o, r, d = env.rollout(9, a)
            o2, r2, d2 = env.rollout(3)                         # in-kernel actions
            env.step(a[0])
            env.reset(torch.ones(n, dtype=torch.uint8, device="cuda"))
            assert torch.isfinite(r).all() and int(d.sum()) == n
            del env
for mode in ("fast_f32", "mixed"):                              # node-parallel kernel (two action tiles with fp64 state), ragged batches
    for n in (3, 45):
        env = gsc.make("supplychain-ntom-v0", num_envs=n, mode=mode, seed=1, auto_reset=True, total_time_steps=5, nodes_per_echelon=[5, 4, 7, 10])
        env.reset()
        a = (torch.rand(7, n, env.n_actions, device="cuda") * 2 - 1)
        o, r, d = env.rollout(7, a)
        env.rollout(4, a[:4].contiguous(), write_obs=False)
        assert torch.isfinite(r).all() and env.scg_info.kernel_variant == 101
        del env
for n in (40, 64):                                               # two-tile pipeline without the observation output
    env = gsc.make("supplychain-ntom-v0", num_envs=n, mode="mixed", seed=1, auto_reset=True, total_time_steps=6)
    env.reset()
    a = (torch.rand(5, n, env.n_actions, device="cuda") * 2 - 1)
    for T in (1, 2, 5):
        env.rollout(T, a[:T].contiguous(), write_obs=False)
    del env
gen = gsc.make("supplychain-ntom-v0", num_envs=65, mode="fast_f32", force_generic=True, build_info=True)
gen.reset(); gen.rollout(4)
for n in (1, 33, 300):
    for kw in ({}, {"customer_demand": (0, 16), "shipment_delays": (0, 5)}, {"levels": 3, "initial_inventory": [12, 12, 12]}):
        bg = gsc.BeerGameEnv2(num_envs=n, seed=2, auto_reset=True, **kw)
        bg.reset()
        acts = torch.randint(0, 30, (37, n, bg.levels), dtype=torch.int32, device="cuda")
        o, r, d = bg.rollout(37, acts)
        bg.rollout(5)
        assert int(d.sum()) == n
host = gsc.make("supplychain-2perstage-v0", num_envs=100, seed=2)
host.reset()
acts = torch.rand(5, 100, 14).pin_memory(); obs = torch.empty(5, 100, 27).pin_memory()
rew = torch.empty(5, 100).pin_memory(); done = torch.empty(5, 100, dtype=torch.uint8).pin_memory()
host.rollout_host(5, acts, obs, rew, done)
torch.cuda.synchronize()
print("sanitize driver ok")
