#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (one GPU).  One JSON line per case; the headline
number is bench.py's.  Algorithmic bytes per env-step B = 4A + 4O + 4 + 1 (B' = 4O + 5 when actions are drawn in-kernel).

    python tools/bench_configs.py [--quick]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc  # noqa: E402

PEAK = 6549.1
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ONLY = os.environ.get("SCG_BENCH_ONLY", "")


def emit(case, n, T, ms, bytes_per_step, **extra):
    v = n * T / (ms * 1e-3)
    gbs = v * bytes_per_step / 1e9
    d = {"case": case, "envs": n, "periods_per_call": T, "ms_per_call": ms, "env_steps_per_s": v,
         "bytes_per_env_step": bytes_per_step, "achieved_GBps": gbs, "frac_of_measured_hbm_peak": gbs / PEAK}
    d.update(extra)
    print(json.dumps(d), flush=True)


def rollout_case(case, env_id, n, T, mode, iters, in_kernel_actions=False, **kw):
    if ONLY and ONLY not in case:
        return
    if "beergame" not in env_id:
        kw = dict(kw, mode=mode)
    env = gsc.make(env_id, num_envs=n, seed=0, auto_reset=True, **kw)
    env.reset()
    A, O = env.n_actions, env.n_obs
    dt = torch.int32 if "beergame" in env_id else env._io_dtype      # exact_f64 / fast_f64 handles move float64
    es = 8 if dt == torch.float64 else 4
    acts = None
    if not in_kernel_actions:
        if dt == torch.int32:
            acts = torch.randint(0, 30, (T, n, A), dtype=torch.int32, device="cuda")
        else:
            acts = torch.empty((T, n, A), dtype=dt, device="cuda").uniform_(-1, 1)
    obs = torch.empty((T, n, O), dtype=dt, device="cuda")
    rew = torch.empty((T, n), dtype=dt, device="cuda")
    done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: env.rollout(T, acts, out=(obs, rew, done)), iters)
    B = (0 if in_kernel_actions else es * A) + es * O + es + 1
    emit(case, n, T, ms, B, mode=mode, kernel_variant=env.scg_info.kernel_variant,
         actions="in-kernel Philox" if in_kernel_actions else "read from HBM",
         working_set_GB=n * T * B / 1e9)
    del env, acts, obs, rew, done
    torch.cuda.empty_cache()


def step_api_case(n, iters):
    """Config 2: supplychain-linear-v0, 65,536 envs, the per-step API (launch-bound: 5 MB per step, lives in L2)."""
    for mode in ("mixed", "fast_f32"):
        env = gsc.make("supplychain-linear-v0", num_envs=n, mode=mode, seed=0, auto_reset=True)
        env.reset()
        acts = torch.empty((n, env.n_actions), device="cuda").uniform_(-1, 1)
        obs = torch.empty((1, n, env.n_obs), device="cuda"); rew = torch.empty((1, n), device="cuda")
        done = torch.empty((1, n), dtype=torch.uint8, device="cuda")
        a1 = acts.unsqueeze(0)
        ms = timed(lambda: env.rollout(1, a1, out=(obs, rew, done)), iters * 20)
        emit("C2 linear step() API, one launch per period", n, 1, ms, 77, mode=mode, note="launch-latency bound; data set fits L2")
        # the same loop captured in a CUDA graph (360 periods per replay)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                env.rollout(1, a1, out=(obs, rew, done))
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(360):
                    env.rollout(1, a1, out=(obs, rew, done))
        ms = timed(g.replay, max(2, iters // 2))
        emit("C2 linear step() API, 360 periods in one CUDA graph", n, 360, ms, 77, mode=mode, note="data set fits L2")
        del g
        # fused rollout of a whole episode
        T = 360
        acts_T = torch.empty((T, n, env.n_actions), device="cuda").uniform_(-1, 1)
        obs_T = torch.empty((T, n, env.n_obs), device="cuda"); rew_T = torch.empty((T, n), device="cuda")
        done_T = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        ms = timed(lambda: env.rollout(T, acts_T, out=(obs_T, rew_T, done_T)), iters)
        emit("C2 linear fused 360-period rollout", n, T, ms, 77, mode=mode, working_set_GB=n * T * 77 / 1e9)
        del env


def big_step_api_case(env_id, n, iters, label):
    """Policy-in-the-loop use: one launch per period on a large batch (`env.step(actions)`): the state (fp64 in `mixed`) is
    read and written every period on top of the actions / obs / reward / done."""
    if ONLY and ONLY not in label:
        return
    for mode in ("mixed", "fast_f32"):
        env = gsc.make(env_id, num_envs=n, mode=mode, seed=0, auto_reset=True)
        env.reset()
        A, O = env.n_actions, env.n_obs
        acts = torch.empty((n, A), device="cuda").uniform_(-1, 1)
        ms = timed(lambda: env.step(acts), iters)
        B = 4 * A + 4 * O + 5
        S = 2 * env.scg_info.state_bytes / n
        emit(label, n, 1, ms, B, mode=mode, state_bytes_per_env_step=S, achieved_GBps_incl_state=n * (B + S) / (ms * 1e-3) / 1e9,
             frac_incl_state=n * (B + S) / (ms * 1e-3) / 1e9 / PEAK, kernel_variant=env.scg_info.kernel_variant)
        del env
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--skip-sc", action="store_true")
    args = ap.parse_args()
    it = 3 if args.quick else 10
    torch.cuda.set_device(0)
    M = 1 << 20
    if not args.skip_sc:
        step_api_case(65536, it)
    for mode in (() if args.skip_sc else ("mixed", "fast_f32")):
        rollout_case("C3 2perstage 1M envs, fused 128-period rollout", "supplychain-2perstage-v0", M, 128, mode, it)
        rollout_case("C3 2perstage 1M envs, fused 128-period rollout", "supplychain-2perstage-v0", M, 128, mode, it,
                     in_kernel_actions=True)
        rollout_case("linear 4M envs, fused 128-period rollout", "supplychain-linear-v0", 4 * M, 128, mode, it)
        rollout_case("C4 ntom 1M envs, in-kernel actions", "supplychain-ntom-v0", M, 128, mode, it, in_kernel_actions=True)
    rollout_case("beergame-v2 8M envs, fused 35-week episode", "beergame-v2", 8 * M, 35, "fast_f32", it)
    rollout_case("beergame-v2 stochastic demand(0,16)/delays(0,5), 8M envs", "beergame-v2", 8 * M, 35, "fast_f32", it,
                 customer_demand=(0, 16), shipment_delays=(0, 5))
    rollout_case("beergame-v0 8M envs, fused 35-week episode", "beergame-v0", 8 * M, 35, "fast_f32", it)
    # config 5 (per-GPU share of the 16M-env sweep at 8 GPUs: 1M + 1M), generic kernel for supplychain-v0
    for mode in ("mixed", "fast_f32"):
        rollout_case("C5 supplychain-v0 demo chain (stochastic lead times avg 2 / max 4), 1M envs", "supplychain-v0", M, 32,
                     mode, max(2, it // 3))
    rollout_case("C5 supplychain-v0 demo chain, exact_f64 heapq-exact kernel, 128k envs", "supplychain-v0", 1 << 17, 16,
                 "exact_f64", 2)
    big_step_api_case("supplychain-ntom-v0", 1 << 20, it * 4, "C4 ntom 1M envs, step() API (one launch per period, state in HBM)")
    big_step_api_case("supplychain-2perstage-v0", 1 << 20, it * 4, "C3 2perstage 1M envs, step() API (one launch per period, state in HBM)")
    # runtime-topology (generic) kernels: what an arbitrary nodes_info gets
    rollout_case("generic kernel: ntom 256k envs (force_generic)", "supplychain-ntom-v0", 1 << 18, 32, "fast_f32", 3, force_generic=True)
    rollout_case("generic kernel: ntom 256k envs (force_generic)", "supplychain-ntom-v0", 1 << 18, 32, "mixed", 3, force_generic=True)
    rollout_case("generic kernel: ntom exact_f64 heapq-exact, 128k envs", "supplychain-ntom-v0", 1 << 17, 16, "exact_f64", 2)
    # chains without a pre-built kernel: the layered template instantiated at run time (kernel_variant 100)
    for mode in ("fast_f32", "mixed"):
        rollout_case("run-time instantiated: n-to-m 5-4-7-10 (26 nodes, A=246, O=177), 256k envs", "sc-Nperstage-multiproduct-v0", 1 << 18, 32,
                     mode, 3, nodes_per_echelon=[5, 4, 7, 10])
        rollout_case("run-time instantiated: 3-3-3-3, 3 products, stochastic lead times avg 2 / max 4, 512k envs", "supplychain-ntom-v0",
                     1 << 19, 64, mode, 3, num_products=3, stochastic_leadtimes=True, avg_leadtime=2, max_leadtime=4)
        rollout_case("run-time instantiated: 2-3-2-4, 1 product, lead time 3, 1M envs", "supplychain-ntom-v0",
                     1 << 20, 128, mode, 3, nodes_per_echelon=[2, 3, 2, 4], num_products=1, avg_leadtime=3, max_leadtime=3)
    rollout_case("run-time instantiated: 10-per-stage (40 nodes, A=620, O=..), 64k envs", "sc-Nperstage-multiproduct-v0", 1 << 16, 16,
                 "fast_f32", 2, nodes_per_echelon=10)
    rollout_case("generic kernel: n-to-m 5-4-7-10 (26 nodes, A=246, O=177), 128k envs (force_generic)", "sc-Nperstage-multiproduct-v0",
                 1 << 17, 16, "fast_f32", 2, nodes_per_echelon=[5, 4, 7, 10], force_generic=True)
    rollout_case("generic kernel: 10-per-stage (40 nodes, A=620), 64k envs (force_generic)", "sc-Nperstage-multiproduct-v0", 1 << 16, 8,
                 "fast_f32", 2, nodes_per_echelon=10, force_generic=True)


if __name__ == "__main__":
    main()
