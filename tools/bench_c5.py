#!/usr/bin/env python
"""BASELINE.json configs[4] on one GPU: the 50/50 sweep of beergame-v2 (stochastic demand (0,16), delays (0,5)) and
supplychain-v0 (the `__main__` 2-2-2-2 chain, stochastic lead times avg 2 / max 4, T=360) with auto-reset -- the per-GPU share
of the 16M-env job at 8 GPUs is 1M + 1M envs -- as TWO handles on one GPU, each on its own stream, >= 2 full episodes
of the longer env.  Prints one JSON line: env-steps/s of the two handles run back to back and run concurrently.

    python tools/bench_c5.py [--envs 1048576] [--mode mixed]
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--mode", default="mixed")
args = ap.parse_args()
n = args.envs
sc = gsc.make("supplychain-v0", num_envs=n, mode=args.mode, seed=0, auto_reset=True)
bg = gsc.make("beergame-v2", num_envs=n, seed=0, auto_reset=True, customer_demand=(0, 16), shipment_delays=(0, 5))
sc.reset(); bg.reset()
T_sc, T_bg = 72, 35                      # periods per launch: 10 launches = 2 supply-chain episodes, 20 beer-game episodes
launches = 10
a_sc = torch.empty((T_sc, n, sc.n_actions), device="cuda").uniform_(-1, 1)
o_sc = torch.empty((T_sc, n, sc.n_obs), device="cuda"); r_sc = torch.empty((T_sc, n), device="cuda")
d_sc = torch.empty((T_sc, n), dtype=torch.uint8, device="cuda")
a_bg = torch.randint(0, 30, (T_bg, n, 4), dtype=torch.int32, device="cuda")
o_bg = torch.empty((T_bg, n, 4), dtype=torch.int32, device="cuda"); r_bg = torch.empty((T_bg, n), dtype=torch.int32, device="cuda")
d_bg = torch.empty((T_bg, n), dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(concurrent):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if concurrent:
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for _ in range(launches):
            with torch.cuda.stream(s1):
                sc.rollout(T_sc, a_sc, out=(o_sc, r_sc, d_sc))
            with torch.cuda.stream(s2):
                bg.rollout(T_bg, a_bg, out=(o_bg, r_bg, d_bg)); bg.rollout(T_bg, a_bg, out=(o_bg, r_bg, d_bg))
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    else:
        for _ in range(launches):
            sc.rollout(T_sc, a_sc, out=(o_sc, r_sc, d_sc))
            bg.rollout(T_bg, a_bg, out=(o_bg, r_bg, d_bg)); bg.rollout(T_bg, a_bg, out=(o_bg, r_bg, d_bg))
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


run(False); run(True)
ms_seq, ms_con = run(False), run(True)
steps = launches * n * (T_sc + 2 * T_bg)
B_sc = 4 * sc.n_actions + 4 * sc.n_obs + 5
bytes_total = launches * n * (T_sc * B_sc + 2 * T_bg * 37)
st_sc, st_bg = sc.return_stats(), bg.return_stats()
print(json.dumps({"case": "C5 sweep: beergame-v2 (stochastic) + supplychain-v0 (stochastic lead times), two handles on one GPU, auto-reset",
                  "envs_per_handle": n, "mode": args.mode, "supplychain_kernel_variant": sc.scg_info.kernel_variant,
                  "env_steps": steps, "ms_back_to_back": ms_seq, "ms_two_streams": ms_con,
                  "env_steps_per_s_back_to_back": steps / (ms_seq * 1e-3), "env_steps_per_s_two_streams": steps / (ms_con * 1e-3),
                  "algorithmic_GBps_two_streams": bytes_total / (ms_con * 1e-3) / 1e9,
                  "episodes_finished": {"supplychain-v0": float(st_sc[0]), "beergame-v2": float(st_bg[0])}}), flush=True)
