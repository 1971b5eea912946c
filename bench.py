#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched gym_supplychain hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's Python loop on host cores

Workload (config.workload): supplychain-ntom-v0 = SupplyChainNPerStage(3) defaults (12 nodes, 2 products,
A=60, O=79), 1,048,576 envs per GPU (weak scaling: 8 GPUs = BASELINE.json configs[3]'s 8M envs), one bench
"step" = one fused 128-period rollout of every env with fp32 random actions READ FROM HBM and obs / reward /
done WRITTEN TO HBM (561 algorithmic bytes per env-step), followed by the episode-return statistics
all-reduce.  Inputs + outputs per step are ~75 GB per GPU, far larger than the 126 MB L2, so no L2 flush
is needed between timed iterations.

One JSON line on stdout (rank 0).  `value` = whole-job env-steps/s with inputs resident in HBM; `e2e` = the
same metric through the host-buffer C-ABI call (pinned host actions in, obs/reward/done back to host,
copies inside the timed region); `roofline` = achieved algorithmic HBM GB/s of the rollout kernel (CUDA
events around each launch) over the measured peak; `cpu_baseline` = the CPU oracle port (a faithful Python
restatement of the reference's per-env loop) on the box's host cores.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_ID = "supplychain-ntom-v0"
A, O = 60, 79
BYTES_PER_ENV_STEP = 4 * A + 4 * O + 4 + 1          # fp32 actions in; fp32 obs + fp32 reward + u8 done out


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop, one process per host core
# ----------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n_envs, n_steps = args
    import numpy as np
    from gym_supplychain_b200 import presets
    from oracle.sc_oracle import SupplyChainOracle
    info, kw = presets.n_per_stage(nodes_per_echelon=3)
    env = SupplyChainOracle(info, **kw)
    env.seed(seed)
    rs = np.random.RandomState(seed)
    acts = rs.uniform(-1, 1, (n_steps, env.n_actions)).astype(np.float32)   # pre-sampled random policy
    t0 = time.perf_counter()
    done_steps = 0
    for _ in range(n_envs):
        env.reset()                                   # draws the whole-episode demand table like the reference
        for t in range(n_steps):
            env.step(acts[t])
        done_steps += n_steps
    return done_steps, time.perf_counter() - t0


def _cpu_worker_c(args):
    """The plain-C oracle (oracle/sc_oracle.c) on the same workload: one episode = reset + n_steps periods with obs."""
    seed, n_envs, n_steps = args
    import numpy as np
    from gym_supplychain_b200 import presets
    from oracle.c_oracle import CSupplyChainOracle
    info, kw = presets.n_per_stage(nodes_per_echelon=3)
    env = CSupplyChainOracle(info, **kw)
    rs = np.random.RandomState(seed)
    acts = rs.uniform(-1, 1, (n_steps, env.n_actions)).astype(np.float32).astype(np.float64)
    dem = rs.randint(0, 401, size=(env.T + 1, env.n_retailers, env.P)).astype(np.int32)
    t0 = time.perf_counter()
    for _ in range(n_envs):
        env.reset(dem)
        env.rollout(acts)
    return n_envs * n_steps, time.perf_counter() - t0


def cpu_sample(pool, cores, envs_per_worker, n_steps, seed0=0):
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, [(seed0 + i, envs_per_worker, n_steps) for i in range(cores)])
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


def run_reference(args):
    """--impl reference: rank 0 times the CPU port on all host cores; other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    T = args.rollout_T
    envs_per_worker = 16
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            cpu_sample(pool, cores, 1, 32, seed0=1000 * w)
        t0 = time.perf_counter()
        total = 0
        for k in range(args.steps):
            n, _ = cpu_sample(pool, cores, envs_per_worker, T, seed0=17 * k)
            total += n
        wall = time.perf_counter() - t0
    v = total / wall
    sample = "%d steps x %d workers x %d envs x %d periods (incl. reset)" % (args.steps, cores, envs_per_worker, T)
    line = {"impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ENV_ID + " (SupplyChainNPerStage(3): 12 nodes, 2 products, A=60, O=79), random policy",
                       "arm": "CPU oracle port of the reference's Python per-env loop (the reference itself needs `gym`, "
                              "which the box does not have); one process per host core"},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, skip=0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[skip:]:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(gpu_index):
    """Several ranks stream 90 GB/s each through host memory in the e2e leg: keep each rank's threads (and so its
    first-touched pinned buffers) on the NUMA node its GPU hangs off.  Best effort; returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(gpu_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                                     # 0000:1b:00.0
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:  # noqa: BLE001
        pass
    return None


def measured_traffic(mode):
    """DRAM bytes per launch of the rollout kernel from the committed ncu capture (profiles/r1_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            return float(json.load(f)[mode]["bytes_per_launch"])
    except Exception:  # noqa: BLE001
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import gym_supplychain_b200 as gsc
    from gym_supplychain_b200.dist import allreduce_return_stats, shard_range, summarize_stats

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout = the one JSON line
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None      # pinned e2e buffers next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_local, T = args.envs_per_gpu, args.rollout_T
    n_total = n_local * world
    first, n_mine = shard_range(n_total, rank, world)
    assert n_mine == n_local

    acts = torch.empty((T, n_local, A), dtype=torch.float32, device=dev)
    acts.uniform_(-1.0, 1.0)                                        # synthetic random policy, resident in HBM
    obs = torch.empty((T, n_local, O), dtype=torch.float32, device=dev)
    rew = torch.empty((T, n_local), dtype=torch.float32, device=dev)
    done = torch.empty((T, n_local), dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def measure(mode, steps, warmup, sample_clocks):
        """Time `steps` bench steps of `mode`; returns (env, dict).  A step = fused rollout + stats all-reduce."""
        env = gsc.make(ENV_ID, num_envs=n_local, device=dev, seed=args.seed, first_env=first, auto_reset=True, mode=mode)
        assert env.n_actions == A and env.n_obs == O and env.scg_info.kernel_variant == 3
        env.reset()
        stats_total = torch.zeros(5, dtype=torch.float64, device=dev)

        def one_step(ev_pair):
            ev_pair[0].record()
            env.rollout(T, acts, out=(obs, rew, done))
            ev_pair[1].record()
            s = allreduce_return_stats(env.return_stats(clear=True))
            stats_total[0:3] += s[0:3]
            stats_total[3] = torch.minimum(stats_total[3], s[3])
            stats_total[4] = torch.maximum(stats_total[4], s[4])

        sampler = ClockSampler(local)
        if rank == 0 and sample_clocks:
            sampler.start()                  # started before warm-up: NVML start-up must not perturb the timed region
        warm_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for _ in range(warmup):              # identical code path to the timed loop (no lazy module loads left for it)
            one_step(warm_ev)
        barrier()
        n_warm_samples = len(sampler.lines)
        stats_total.zero_()
        stats_total[3], stats_total[4] = 1e300, -1e300
        launches0 = env.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_start.record()
        for k in range(steps):
            one_step(ev[k])
        t_end.record()
        barrier()
        launches = env.launch_count() - launches0
        clocks = sampler.stop(skip=n_warm_samples) if (rank == 0 and sample_clocks) else None
        el = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
        km = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
            dist.all_reduce(km, op=dist.ReduceOp.MAX)
        el, km = float(el), float(km)
        achieved = n_local * T * BYTES_PER_ENV_STEP / (km * 1e-3) / 1e9
        return env, {"value": n_total * T * steps / (el * 1e-3), "elapsed_ms": el, "kernel_ms": km, "launches": launches,
                     "clocks": clocks, "achieved_gbs": achieved, "stats": stats_total.clone()}

    env, m = measure(args.mode, args.steps, args.warmup, True)
    value, elapsed_ms, kern_ms, launches, clocks, stats_total = (m["value"], m["elapsed_ms"], m["kernel_ms"], m["launches"],
                                                                  m["clocks"], m["stats"])

    # ---- e2e: host buffers through the C ABI (H2D actions, kernel, D2H obs/reward/done inside the timed region) ----
    # pinned host buffers: 0.6 GB per period per rank at 1M envs; with many ranks on one host keep the total modest and
    # fall back to fewer periods if the host refuses to pin that much
    Te = args.e2e_T if world <= 2 else max(4, min(args.e2e_T, 64 // world))
    while True:
        try:
            h_act = torch.empty((Te, n_local, A), dtype=torch.float32).pin_memory()
            h_obs = torch.empty((Te, n_local, O), dtype=torch.float32).pin_memory()
            h_rew = torch.empty((Te, n_local), dtype=torch.float32).pin_memory()
            h_done = torch.empty((Te, n_local), dtype=torch.uint8).pin_memory()
            break
        except RuntimeError:
            if Te <= 1:
                raise
            h_act = h_obs = h_rew = h_done = None
            Te = max(1, Te // 2)
    if world > 1:                                                   # every rank must run the same number of periods
        te = torch.tensor([Te], dtype=torch.int64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MIN)
        if int(te) < Te:
            Te = int(te)
            h_act, h_obs, h_rew, h_done = h_act[:Te], h_obs[:Te], h_rew[:Te], h_done[:Te]
    h_act.uniform_(-1.0, 1.0)
    for _ in range(2):
        env.rollout_host(Te, h_act, h_obs, h_rew, h_done)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        env.rollout_host(Te, h_act, h_obs, h_rew, h_done)        # returns when the results are in host memory
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_total * Te * args.e2e_steps / float(e2e_s)
    assert torch.isfinite(h_rew).all() and float(h_obs.abs().max()) <= 1.0

    # ---- the other fp32-I/O mode, same buffers, shorter run (reported beside the headline, never as `value`) ----
    alt = None
    if not args.no_alt:
        other = "fast_f32" if args.mode == "mixed" else "mixed"
        del env
        torch.cuda.empty_cache()
        env, ma = measure(other, max(3, min(args.steps, 5)), 3, False)
        peak_alt, _ = measured_peak()
        alt = {other: {"value": ma["value"], "unit": "env-steps/s", "kernel_ms": ma["kernel_ms"],
                       "roofline_frac": ma["achieved_gbs"] / peak_alt,
                       "note": ("all-fp32 arithmetic: per-step reward error vs the reference up to 2.2e-6 relative on this workload "
                                "(profiles/r1_precision_survey.txt), above the 1e-6 north-star bar, hence not the headline") if other == "fast_f32"
                               else "fp32 I/O, fp64 state and arithmetic (1e-6 bar met)"}}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            cpu_sample(pool, cores, 1, 16)
            n_steps_cpu, wall = cpu_sample(pool, cores, args.cpu_envs_per_worker, T)
            cpu = {"value": n_steps_cpu / wall, "unit": "env-steps/s", "cores": cores, "kind": "port",
                   "sample": "%d workers x %d envs x %d periods of %s, random policy, incl. reset()" %
                             (cores, args.cpu_envs_per_worker, T, ENV_ID),
                   "what": "oracle/sc_oracle.py: the same per-env Python loop as the reference (the reference itself needs `gym`)"}
            # the plain-C oracle on the same cores, for context (the checker, not the reference's own speed)
            try:
                from oracle import c_oracle
                c_oracle.build()
                pool.map(_cpu_worker_c, [(i, 2, T) for i in range(cores)])
                t0 = time.perf_counter()
                res = pool.map(_cpu_worker_c, [(100 + i, 40 * args.cpu_envs_per_worker, T) for i in range(cores)])
                wall_c = time.perf_counter() - t0
                cpu["c_port"] = {"value": sum(r[0] for r in res) / wall_c, "unit": "env-steps/s", "cores": cores,
                                 "what": "oracle/sc_oracle.c (plain C restatement, gcc -O2), one process per core"}
            except Exception as e:  # noqa: BLE001
                cpu["c_port"] = {"unavailable": str(e)[:200]}

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = m["achieved_gbs"]
        cnt, mean, std, mn, mx = summarize_stats(stats_total)
        line = {"metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.mode == "fast_f32" else "f32 I/O, f64 state+arithmetic",
                "data": "synthetic",
                "config": {"workload": ENV_ID + " (SupplyChainNPerStage(3): 12 nodes, 2 products, A=60, O=79), "
                                                "fused %d-period rollout, fp32 random actions read from HBM" % T,
                           "envs_per_gpu": n_local, "envs_total": n_total, "rollout_T": T, "mode": args.mode,
                           "bytes_per_env_step": BYTES_PER_ENV_STEP, "auto_reset": True,
                           "l2": "inputs+outputs per step are %.1f GB per GPU (> 126 MB L2): no flush needed"
                                 % (n_local * T * BYTES_PER_ENV_STEP / 1e9),
                           "collective": "one all-gather of the 5 episode-return statistics (40 B per rank) per step (NCCL), reduced locally" if world > 1 else "none (1 GPU)"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": n_local * Te * A * 4,
                        "d2h_bytes_per_step": n_local * Te * (O * 4 + 4 + 1), "rollout_T": Te, "steps": args.e2e_steps,
                        "api": "scg_rollout_host (pinned host buffers)", "numa_node": numa},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": measured_traffic(args.mode) if (n_local == 1 << 20 and T == 128) else None,
                             "traffic_source": "ncu --set full capture of the same command (profiles/r1_traffic.json)", "kernel": "sc_layered_rollout<%s,float,3,3,3,3,2,2>" % ("float" if args.mode == "fast_f32" else "double"), "kernel_ms": kern_ms,
                             "peak_source": peak_src, "bytes_per_launch": n_local * T * BYTES_PER_ENV_STEP},
                "cpu_baseline": cpu, "alt_modes": alt,
                "episodes_finished": cnt, "mean_episode_return": mean}
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit_line(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner goes to
    fd 1) has been redirected to stderr by main()."""
    (_JSON_OUT or sys.stdout).write(json.dumps(line) + "\n")
    (_JSON_OUT or sys.stdout).flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                                   # fd 1 -> stderr for the rest of the process (and its children)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--rollout-T", type=int, default=128)
    ap.add_argument("--e2e-T", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-envs-per-worker", type=int, default=96)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--mode", default="mixed", choices=["fast_f32", "mixed"])
    ap.add_argument("--no-alt", action="store_true", help="skip the secondary measurement of the other fp32-I/O mode")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
