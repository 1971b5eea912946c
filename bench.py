#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched gym_supplychain hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                   # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the UNMODIFIED reference on the host cores
    python bench.py --config {c1,c2,c3,c4,c4s,c5} ...               # another BASELINE.json config as the main line

Default workload (`--config c4`, BASELINE.json configs[3], the one the metric is quoted on): supplychain-ntom-v0 =
SupplyChainNPerStage(3) defaults (12 nodes, 2 products, A=60, O=79), 1,048,576 envs per GPU (weak scaling: 8 GPUs = 8M
envs); one bench "step" = one fused 128-period rollout of every env with fp32 random actions READ FROM HBM and
obs / reward / done WRITTEN TO HBM (561 algorithmic bytes per env-step), plus the episode-return statistics exchange
(one 40-byte all-gather per step, consumed one step late so that it overlaps the next rollout).  Inputs + outputs per
step are ~75 GB per GPU, far larger than the 126 MB L2, so no L2 flush is needed between timed iterations.

ONE JSON line on stdout (rank 0):
  value .......... whole-job env-steps/s with inputs resident in HBM
  e2e ............ the same metric through the host-buffer C-ABI call (pinned host actions in, obs/reward/done back to
                   host, copies inside the timed region), with the duplex PCIe ceiling measured in the same run
  roofline ....... achieved algorithmic HBM GB/s of the rollout kernel (CUDA events around each launch) / measured peak
  cpu_baseline ... the reference itself (baseline/_ref through oracle/gym_shim) on the box's host cores, bounded sample
  other_configs .. every other BASELINE.json config, measured in the same run: c1 (beergame-v0 single env), c2 (linear
                   65,536 envs: step API / CUDA graph / fused), c3 (2perstage 1M envs x 128), c4s (ntom STRONG scaling:
                   8,388,608 envs in total), c5 (beergame-v2 + supplychain-v0 16,777,216-env sweep, two handles per GPU)
                   -- c4s and c5 at every N, the single-GPU configs at N=1 only; plus `large_chain_5_4_7_10`, the reference's
                   26-node golden-return chain on the node-parallel split kernel (not a BASELINE config).
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
BG_BYTES = 37                                        # beer game: int32 actions[4] + obs[4] + reward, u8 done


def bytes_per_env_step(n_actions, n_obs):
    return 4 * n_actions + 4 * n_obs + 4 + 1


# ----------------------------------------------------------------------------------------------
# CPU arm.  kind "reference": the unmodified reference package installed under baseline/_ref, imported through the
# `gym` stand-in (oracle/gym_shim) and driven as its own tests drive it (tests/utils.py:13-22: reset(), then step()
# until the sample is over) with a pre-sampled random policy; one process per host core.
# ----------------------------------------------------------------------------------------------
REF_CASES = {
    # name: (gym id, constructor kwargs factory, action sampler)
    "ntom": "gym_supplychain:sc-Nperstage-multiproduct-v0",          # SupplyChainNPerStage() defaults = 3-3-3-3, P=2
    "linear": "gym_supplychain:sc-Nperstage-multiproduct-v0",        # nodes_per_echelon=1, num_products=1
    "2perstage": "gym_supplychain:sc-2perstage-v0",
    "supplychain-v0": "gym_supplychain:supplychain-v0",              # the __main__ demo chain, stochastic lead times
    "beergame-v0": "gym_supplychain:beergame-v0",
    "beergame-v2": "gym_supplychain:beergame-v2",
    "beergame-v2-stochastic": "gym_supplychain:beergame-v2",
}


def _make_reference_env(case):
    from baseline.install_ref import import_reference
    gym = import_reference()
    kw = {}
    if case == "linear":
        kw = dict(nodes_per_echelon=1, num_products=1)
    elif case == "supplychain-v0":
        from gym_supplychain_b200 import presets
        info, kw = presets.main_demo_chain()
        kw = dict(kw, nodes_info=info)
    elif case == "beergame-v2-stochastic":
        kw = dict(customer_demand=[0, 16], shipment_delays=[0, 5])
    return gym.make(REF_CASES[case], **kw)


def _ref_worker(args):
    """One host core: `n_eps` x (reset + `n_steps` periods) of the reference env `case`.  Returns (env-steps, seconds)."""
    case, seed, n_eps, n_steps = args
    import numpy as np
    env = _make_reference_env(case)
    rs = np.random.RandomState(seed)
    if case.startswith("beergame"):
        lo, hi = (0, 5) if case == "beergame-v0" else (0, 30)
        acts = rs.randint(lo, hi, (n_steps, 4))
    else:
        if hasattr(env, "seed"):
            env.seed(seed)
        acts = rs.uniform(-1, 1, (n_steps, env.action_space.shape[0])).astype(np.float32)
    t0 = time.perf_counter()
    done_steps = 0
    for _ in range(n_eps):
        env.reset()
        for t in range(n_steps):
            env.step(acts[t])
        done_steps += n_steps
    return done_steps, time.perf_counter() - t0


def _port_worker(args):
    """The Python restatement (oracle/sc_oracle.py) on the same ntom sample -- reported beside the reference."""
    seed, n_eps, n_steps = args
    import numpy as np
    from gym_supplychain_b200 import presets
    from oracle.sc_oracle import SupplyChainOracle
    info, kw = presets.n_per_stage(nodes_per_echelon=3)
    env = SupplyChainOracle(info, **kw)
    env.seed(seed)
    acts = np.random.RandomState(seed).uniform(-1, 1, (n_steps, env.n_actions)).astype(np.float32)
    t0 = time.perf_counter()
    for _ in range(n_eps):
        env.reset()
        for t in range(n_steps):
            env.step(acts[t])
    return n_eps * n_steps, time.perf_counter() - t0


def _c_port_worker(args):
    """The plain-C restatement (oracle/sc_oracle.c): one episode = reset + n_steps periods with obs."""
    seed, n_eps, n_steps = args
    import numpy as np
    from gym_supplychain_b200 import presets
    from oracle.c_oracle import CSupplyChainOracle
    info, kw = presets.n_per_stage(nodes_per_echelon=3)
    env = CSupplyChainOracle(info, **kw)
    rs = np.random.RandomState(seed)
    acts = rs.uniform(-1, 1, (n_steps, env.n_actions)).astype(np.float32).astype(np.float64)
    dem = rs.randint(0, 401, size=(env.T + 1, env.n_retailers, env.P)).astype(np.int32)
    t0 = time.perf_counter()
    for _ in range(n_eps):
        env.reset(dem)
        env.rollout(acts)
    return n_eps * n_steps, time.perf_counter() - t0


def reference_available():
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gym_supplychain"))


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except Exception:  # noqa: BLE001
        pass
    return "unknown"


def cpu_rate(pool, cores, case, n_eps, n_steps, seed0=0):
    """Aggregate env-steps/s of `case` on `cores` processes (wall clock around the whole map, reset() included)."""
    t0 = time.perf_counter()
    res = pool.map(_ref_worker, [(case, seed0 + i, n_eps, n_steps) for i in range(cores)])
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res) / wall, wall


REF_SAMPLE = {  # config -> [(reference case, periods per reset, weight = env-steps of that family in the GPU job's mix)]
    "c1": [("beergame-v0", 35, 1.0)],
    "c2": [("linear", 128, 1.0)],
    "c3": [("2perstage", 128, 1.0)],
    "c4": [("ntom", 128, 1.0)],
    "c4s": [("ntom", 128, 1.0)],
    "c5": [("supplychain-v0", 360, 720.0), ("beergame-v2-stochastic", 35, 700.0)],
}


def mixed_rate(rates_weights):
    """env-steps/s of a job that runs w_i env-steps of family i at rate r_i one after the other on the same cores."""
    tot = sum(w for _, w in rates_weights)
    return tot / sum(w / r for r, w in rates_weights)


def run_reference(args):
    """--impl reference: rank 0 times the reference on all host cores; other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    cfg = args.config
    if not reference_available():
        emit_line({"impl": "reference", "unavailable": "baseline/_ref missing (python baseline/install_ref.py in the build container)"})
        return
    cases = REF_SAMPLE[cfg]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            for case, T, _ in cases:
                cpu_rate(pool, cores, case, 1, min(T, 32), seed0=1000 * w)
        t0 = time.perf_counter()
        rates = []
        for k in range(args.steps):
            per = []
            for case, T, wgt in cases:
                eps = {"ntom": 16, "linear": 96, "2perstage": 48, "supplychain-v0": 4, "beergame-v0": 800,
                       "beergame-v2-stochastic": 320}[case]
                r, _ = cpu_rate(pool, cores, case, eps, T, seed0=17 * k)
                per.append((r, wgt))
            rates.append(mixed_rate(per))
        wall = time.perf_counter() - t0
    v = statistics.mean(rates)
    sample = "%d steps x %d workers x %s (reset() included)" % (
        args.steps, cores, " + ".join("%s: episodes of %d periods" % (c, T) for c, T, _ in cases))
    line = {"impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": full_config(cfg, args, int(os.environ.get("WORLD_SIZE", "1"))),
            "arm": "the unmodified reference (baseline/_ref, pip-installed from /root/reference) through the gym stand-in, its own "
                   "reset()/step() loop with a pre-sampled random policy, one process per host core",
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "cpu": cpu_model(), "kind": "reference", "sample": sample},
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
    """Several ranks stream tens of GB/s each through host memory in the e2e leg: keep each rank's threads (and so its
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


def offline_traffic(mode):
    """DRAM bytes per launch of the rollout kernel from the committed ncu capture of this command (measured OFFLINE,
    under the profiler -- never in this run)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)[mode]["bytes_per_launch"]), "profiles/" + name
        except Exception:  # noqa: BLE001
            continue
    return None, None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(cfg, n_local, n_total, T, mode):
    """`config` of the JSON line: names the workload (shared by both arms so that their configs agree)."""
    names = {
        "c1": "beergame-v0, single env, 4 stages, random policy RandomState(0).randint(0,5,4), one 35-week episode",
        "c2": "supplychain-linear-v0 (1-1-1-1, P=1, A=4, O=14), 65,536 envs, random actions",
        "c3": "supplychain-2perstage-v0 (A=14, O=27), 1,048,576 envs, fused 128-period rollout, fp32 random actions read from HBM",
        "c4": ENV_ID + " (SupplyChainNPerStage(3): 12 nodes, 2 products, A=60, O=79), fused 128-period rollout, "
                       "fp32 random actions read from HBM, 1,048,576 envs per GPU",
        "c4s": ENV_ID + " (A=60, O=79), STRONG scaling: 8,388,608 envs in total split over the GPUs, 128 periods per step "
                        "in fused rollouts of 128*N/8 periods",
        "c5": "50/50 sweep of beergame-v2 (stochastic demand (0,16), delays (0,5)) and supplychain-v0 (demo 2-2-2-2 chain, "
              "stochastic lead times avg 2 / max 4, T=360), 16,777,216 envs in total, auto-reset, two handles per GPU, "
              "2 supply-chain episodes (720 periods) + 20 beer-game episodes (700 weeks) per env",
    }
    c = {"workload": names[cfg], "id": cfg}
    if n_local is not None:
        c.update({"envs_per_gpu": n_local, "envs_total": n_total, "rollout_T": T, "mode": mode,
                  "obs_mode": "canonical (true per-arrival-time sums; equals the reference obs whenever its heap list is "
                              "time-sorted, SURVEY F3 -- only mode exact_f64 reproduces the heap-layout quirk)",
                  "auto_reset": True})
    return c


CONFIG_SHAPES = {"c3": (14, 27), "c4": (60, 79), "c4s": (60, 79)}     # (A, O) of the rollout configs


def full_config(cfg, args, world):
    """The `config` object of the JSON line, computed from the command line alone so that BOTH arms print the same one."""
    mode = args.mode
    if cfg in CONFIG_SHAPES:
        nA, nO = CONFIG_SHAPES[cfg]
        if cfg == "c4s":
            n_local, T_launch, lps = c4s_shape(world)
        else:
            n_local, T_launch, lps = args.envs_per_gpu, args.rollout_T, 1
        B = bytes_per_env_step(nA, nO)
        return dict(workload_config(cfg, n_local, n_local * world, T_launch * lps, mode), bytes_per_env_step=B, periods_per_launch=T_launch,
                    l2="inputs+outputs per launch are %.1f GB per GPU (> 126 MB L2): no flush needed" % (n_local * T_launch * B / 1e9),
                    collective=("one all-gather of the 5 episode-return statistics (40 B per rank) per step (NCCL), issued "
                                "async after the rollout and consumed one step late: overlapped with the next rollout")
                    if world > 1 else "none (1 GPU)")
    if cfg == "c2":
        return dict(workload_config("c2", 65536, 65536, 360, mode),
                    l2="fused rollout: 1.8 GB per call (> L2); the step API's 5 MB per step lives in L2 (reported as such)")
    if cfg == "c5":
        per_family = (1 << 24) // 2 // world
        return dict(workload_config("c5", per_family, 1 << 24, None, mode),
                    l2="every launch streams more than 2 GB per GPU (> 126 MB L2): no flush needed")
    return workload_config("c1", 1, 1, 35, None)


# ----------------------------------------------------------------------------------------------
# our arm: shared pieces
# ----------------------------------------------------------------------------------------------
class Ctx(object):
    """Per-process distributed context."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout = the one JSON line
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus)
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def min_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t)


class StatsExchange(object):
    """The episode-return statistics of a bench step, exchanged WITHOUT stalling the next rollout: the 40-byte
    all-gather of step k is issued right after rollout k (async, on NCCL's stream) and consumed after rollout k+1 has been
    enqueued, from a double-buffered record."""

    def __init__(self, ctx):
        self.ctx, torch = ctx, ctx.torch
        self.total = torch.zeros(5, dtype=torch.float64, device=ctx.dev)
        self.reset_total()
        self.buf = [torch.zeros(5, dtype=torch.float64, device=ctx.dev) for _ in range(2)]
        self.parts = [[torch.empty(5, dtype=torch.float64, device=ctx.dev) for _ in range(ctx.world)] for _ in range(2)]
        self.pending = None
        self.k = 0

    def reset_total(self):
        self.total.zero_()
        self.total[3], self.total[4] = 1e300, -1e300

    def _fold(self, allr):
        self.total[0:3] += allr[:, 0:3].sum(0)
        self.total[3] = self.ctx.torch.minimum(self.total[3], allr[:, 3].min())
        self.total[4] = self.ctx.torch.maximum(self.total[4], allr[:, 4].max())

    def submit(self, stats):
        """Call right after a rollout was enqueued: starts this step's exchange, then consumes the previous one."""
        b = self.k & 1
        self.k += 1
        self.buf[b].copy_(stats)
        prev = self.pending
        if self.ctx.world > 1:
            work = self.ctx.dist.all_gather(self.parts[b], self.buf[b], async_op=True)
            self.pending = (work, b)
        else:
            self.pending = (None, b)
        return prev

    def consume(self, item):
        if item is None:
            return
        work, b = item
        if work is not None:
            work.wait()
            self._fold(self.ctx.torch.stack(self.parts[b]))
        else:
            self._fold(self.buf[b].unsqueeze(0))

    def drain(self):
        self.consume(self.pending)
        self.pending = None


def time_steps(ctx, one_step, steps, warmup, after_enqueue=None, finish=None):
    """W warm-up steps, then exactly `steps` timed steps between barrier + synchronize, CUDA events on the launching
    stream; returns (elapsed_ms max over ranks, mean kernel ms per step max over ranks)."""
    torch = ctx.torch
    warm_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for _ in range(warmup):
        one_step(warm_ev)
    if finish:
        finish()
    ctx.barrier()
    if after_enqueue:
        after_enqueue()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    t0.record()
    for k in range(steps):
        one_step(ev[k])
    if finish:
        finish()
    t1.record()
    ctx.barrier()
    el = ctx.max_over_ranks(t0.elapsed_time(t1))
    km = ctx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev) / steps)
    return el, km


# ----------------------------------------------------------------------------------------------
# c3 / c4 / c4s: fused supply-chain rollouts with actions read from HBM
# ----------------------------------------------------------------------------------------------
def run_rollout_config(ctx, args, cfg, mode, steps, warmup, sample_clocks, env_id, n_local, T_launch, launches_per_step):
    """One bench step = `launches_per_step` fused rollouts of T_launch periods of n_local envs (+ the stats exchange)."""
    import gym_supplychain_b200 as gsc
    from gym_supplychain_b200.dist import shard_range, summarize_stats
    torch = ctx.torch
    n_total = n_local * ctx.world
    first, n_mine = shard_range(n_total, ctx.rank, ctx.world)
    assert n_mine == n_local
    env = gsc.make(env_id, num_envs=n_local, device=ctx.dev, seed=args.seed, first_env=first, auto_reset=True, mode=mode)
    assert env.scg_info.kernel_variant > 0
    nA, nO = env.n_actions, env.n_obs
    env.reset()
    acts = torch.empty((T_launch, n_local, nA), dtype=torch.float32, device=ctx.dev)
    acts.uniform_(-1.0, 1.0)                                        # synthetic random policy, resident in HBM
    obs = torch.empty((T_launch, n_local, nO), dtype=torch.float32, device=ctx.dev)
    rew = torch.empty((T_launch, n_local), dtype=torch.float32, device=ctx.dev)
    done = torch.empty((T_launch, n_local), dtype=torch.uint8, device=ctx.dev)
    xch = StatsExchange(ctx)

    def one_step(ev_pair):
        ev_pair[0].record()
        for _ in range(launches_per_step):
            env.rollout(T_launch, acts, out=(obs, rew, done))
        ev_pair[1].record()
        xch.consume(xch.submit(env.return_stats(clear=True)))       # start this step's exchange, fold the previous one

    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0 and sample_clocks:
        sampler.start()                  # started before warm-up: NVML start-up must not perturb the timed region
    marks = {}

    def after_warm():
        marks["n_warm"] = len(sampler.lines)
        xch.reset_total()
        marks["l0"] = env.launch_count()

    el, km = time_steps(ctx, one_step, steps, warmup, after_enqueue=after_warm, finish=xch.drain)
    launches = env.launch_count() - marks["l0"]
    clocks = sampler.stop(skip=marks["n_warm"]) if (ctx.rank == 0 and sample_clocks) else None
    T_step = T_launch * launches_per_step
    B = bytes_per_env_step(nA, nO)
    achieved = n_local * T_step * B / (km * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    cnt, mean, _, _, _ = summarize_stats(xch.total)
    res = {"value": n_total * T_step * steps / (el * 1e-3), "elapsed_ms": el, "kernel_ms": km, "launches": launches,
           "clocks": clocks, "achieved_gbs": achieved, "frac": achieved / peak, "peak": peak, "peak_src": peak_src,
           "bytes_per_env_step": B, "n_local": n_local, "n_total": n_total, "T_step": T_step, "T_launch": T_launch,
           "episodes_finished": cnt, "mean_episode_return": mean, "kernel_variant": env.scg_info.kernel_variant,
           "n_actions": nA, "n_obs": nO}
    return env, (acts, obs, rew, done), res


def kernel_name(env_id, mode):
    shape = {"supplychain-ntom-v0": "3,3,3,3,2,2", "supplychain-2perstage-v0": "2,2,2,2,1,2", "supplychain-linear-v0": "1,1,1,1,1,2",
             "supplychain-v0": "2,2,2,2,1,2,4,stoch"}[env_id]
    return "sc_layered_rollout<%s,float,%s>" % ("float" if mode == "fast_f32" else "double", shape)


def link_ceiling(ctx, h_act, h_obs, h_rew, h_done, n_local, Te, reps=3):
    """The PCIe ceiling of the e2e leg, measured in this run: the SAME pinned buffers copied H2D (actions) and D2H
    (obs / reward / done) concurrently on two streams with plain cudaMemcpyAsync, no kernel.  All ranks at once."""
    torch = ctx.torch
    d_act = torch.empty(h_act.shape, dtype=h_act.dtype, device=ctx.dev)
    d_obs = torch.empty(h_obs.shape, dtype=h_obs.dtype, device=ctx.dev)
    d_rew = torch.empty(h_rew.shape, dtype=h_rew.dtype, device=ctx.dev)
    d_done = torch.empty(h_done.shape, dtype=h_done.dtype, device=ctx.dev)
    s_in, s_out = torch.cuda.Stream(ctx.dev), torch.cuda.Stream(ctx.dev)

    def once():
        with torch.cuda.stream(s_in):
            d_act.copy_(h_act, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_obs.copy_(d_obs, non_blocking=True); h_rew.copy_(d_rew, non_blocking=True); h_done.copy_(d_done, non_blocking=True)
        s_in.synchronize(); s_out.synchronize()
    once()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    ctx.barrier()
    dt = ctx.max_over_ranks(time.perf_counter() - t0) / reps
    h2d, d2h = h_act.numel() * 4, h_obs.numel() * 4 + h_rew.numel() * 4 + h_done.numel()
    del d_act, d_obs, d_rew, d_done
    return {"value": n_local * ctx.world * Te / dt, "unit": "env-steps/s", "h2d_GBps_per_gpu": h2d / dt / 1e9,
            "d2h_GBps_per_gpu": d2h / dt / 1e9, "what": "duplex cudaMemcpyAsync of one e2e step's pinned buffers, no kernel, all ranks at once"}


def run_e2e(ctx, args, env, n_local, nA, nO):
    """Host buffers through the C ABI (H2D actions, kernel, D2H obs/reward/done inside the timed region)."""
    torch = ctx.torch
    per_period = n_local * (nA * 4 + nO * 4 + 5)
    Te = args.e2e_T
    try:                                 # keep >= 16 periods per rank when the host can pin that much
        import psutil
        avail = psutil.virtual_memory().available
        Te = max(4, min(Te, int(0.4 * avail / ctx.world / per_period)))
    except Exception:  # noqa: BLE001
        Te = Te if ctx.world <= 2 else max(4, min(Te, 64 // ctx.world))
    while True:
        try:
            h_act = torch.empty((Te, n_local, nA), dtype=torch.float32).pin_memory()
            h_obs = torch.empty((Te, n_local, nO), dtype=torch.float32).pin_memory()
            h_rew = torch.empty((Te, n_local), dtype=torch.float32).pin_memory()
            h_done = torch.empty((Te, n_local), dtype=torch.uint8).pin_memory()
            break
        except RuntimeError:
            if Te <= 1:
                raise
            h_act = h_obs = h_rew = h_done = None
            Te = max(1, Te // 2)
    te = int(ctx.min_over_ranks(Te))                                # every rank must run the same number of periods
    if te < Te:
        Te = te
        h_act, h_obs, h_rew, h_done = h_act[:Te], h_obs[:Te], h_rew[:Te], h_done[:Te]
    h_act.uniform_(-1.0, 1.0)
    for _ in range(2):
        env.rollout_host(Te, h_act, h_obs, h_rew, h_done)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        env.rollout_host(Te, h_act, h_obs, h_rew, h_done)        # returns when the results are in host memory
    ctx.barrier()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    value = n_local * ctx.world * Te * args.e2e_steps / dt
    assert torch.isfinite(h_rew).all() and float(h_obs.abs().max()) <= 1.0
    ceil = link_ceiling(ctx, h_act, h_obs, h_rew, h_done, n_local, Te)
    h2d, d2h = n_local * Te * nA * 4, n_local * Te * (nO * 4 + 4 + 1)
    per_step_s = dt / args.e2e_steps
    return {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "rollout_T": Te,
            "steps": args.e2e_steps, "api": "scg_rollout_host (pinned host buffers; H2D | kernel | D2H pipelined on three streams)",
            "numa_node": ctx.numa, "h2d_GBps_per_gpu": h2d / per_step_s / 1e9, "d2h_GBps_per_gpu": d2h / per_step_s / 1e9,
            "link_ceiling": ceil, "frac_of_link_ceiling": value / ceil["value"],
            "note": "PCIe-bound: 561 B per env-step cross the link (240 in, 321 out); with several ranks on one host the ceiling is "
                    "the host's memory / root-complex bandwidth shared by all ranks (see link_ceiling, measured with all ranks at once)"}


# ----------------------------------------------------------------------------------------------
# c2: linear, 65,536 envs -- step API, CUDA graph, fused
# ----------------------------------------------------------------------------------------------
def run_c2(ctx, args, mode, iters=6):
    import gym_supplychain_b200 as gsc
    torch = ctx.torch
    n = 65536
    env = gsc.make("supplychain-linear-v0", num_envs=n, mode=mode, seed=args.seed, auto_reset=True, device=ctx.dev)
    env.reset()
    nA, nO = env.n_actions, env.n_obs
    B = bytes_per_env_step(nA, nO)
    peak, _ = measured_peak()

    def timed(fn, it, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / it
    a1 = torch.empty((1, n, nA), device=ctx.dev).uniform_(-1, 1)
    o1 = torch.empty((1, n, nO), device=ctx.dev); r1 = torch.empty((1, n), device=ctx.dev)
    d1 = torch.empty((1, n), dtype=torch.uint8, device=ctx.dev)
    l0 = env.launch_count()
    ms_step = timed(lambda: env.rollout(1, a1, out=(o1, r1, d1)), 200 * iters)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(ctx.dev)
    with torch.cuda.stream(s):
        for _ in range(3):
            env.rollout(1, a1, out=(o1, r1, d1))
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(360):
                env.rollout(1, a1, out=(o1, r1, d1))
    ms_graph = timed(g.replay, iters)
    del g
    T = 360
    aT = torch.empty((T, n, nA), device=ctx.dev).uniform_(-1, 1)
    oT = torch.empty((T, n, nO), device=ctx.dev); rT = torch.empty((T, n), device=ctx.dev)
    dT = torch.empty((T, n), dtype=torch.uint8, device=ctx.dev)
    ms_fused = timed(lambda: env.rollout(T, aT, out=(oT, rT, dT)), iters)
    launches = env.launch_count() - l0

    def rate(ms, periods):
        v = n * periods / (ms * 1e-3)
        return {"value": v, "unit": "env-steps/s", "ms_per_call": ms, "roofline_frac": v * B / 1e9 / peak}
    out = {"workload": workload_config("c2", None, None, None, None)["workload"], "mode": mode, "envs": n, "bytes_per_env_step": B,
           "step_api_one_launch_per_period": dict(rate(ms_step, 1), note="launch-latency bound; 5 MB per step lives in L2"),
           "step_api_360_periods_in_one_cuda_graph": dict(rate(ms_graph, 360), note="data set lives in L2"),
           "fused_360_period_rollout": dict(rate(ms_fused, T), note="1.8 GB per call (> L2): HBM roofline applies"),
           "gpu_launches": launches}
    out["value"] = out["fused_360_period_rollout"]["value"]
    out["roofline_frac"] = out["fused_360_period_rollout"]["roofline_frac"]
    del env
    return out


# ----------------------------------------------------------------------------------------------
# c5: beergame-v2 + supplychain-v0 sweep, two handles per GPU
# ----------------------------------------------------------------------------------------------
def run_c5(ctx, args, mode, total_envs=1 << 24, reps=1):
    import gym_supplychain_b200 as gsc
    from gym_supplychain_b200.dist import shard_range, summarize_stats
    torch = ctx.torch
    per_family = total_envs // 2
    first, n = shard_range(per_family, ctx.rank, ctx.world)
    sc = gsc.make("supplychain-v0", num_envs=n, mode=mode, seed=args.seed, first_env=first, auto_reset=True, device=ctx.dev)
    bg = gsc.make("beergame-v2", num_envs=n, seed=args.seed, first_env=first, auto_reset=True, device=ctx.dev,
                  customer_demand=(0, 16), shipment_delays=(0, 5))
    sc.reset(); bg.reset()
    # 720 supply-chain periods (2 episodes) and 700 beer-game weeks (20 episodes) per env, in launches sized for HBM
    T_sc = 72 if n <= (1 << 21) else (36 if n <= (1 << 22) else 18)
    L_sc = 720 // T_sc
    T_bg, L_bg = 35, 20
    a_sc = torch.empty((T_sc, n, sc.n_actions), device=ctx.dev).uniform_(-1, 1)
    o_sc = torch.empty((T_sc, n, sc.n_obs), device=ctx.dev); r_sc = torch.empty((T_sc, n), device=ctx.dev)
    d_sc = torch.empty((T_sc, n), dtype=torch.uint8, device=ctx.dev)
    a_bg = torch.randint(0, 30, (T_bg, n, 4), dtype=torch.int32, device=ctx.dev)
    o_bg = torch.empty((T_bg, n, 4), dtype=torch.int32, device=ctx.dev); r_bg = torch.empty((T_bg, n), dtype=torch.int32, device=ctx.dev)
    d_bg = torch.empty((T_bg, n), dtype=torch.uint8, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(ctx.dev), torch.cuda.Stream(ctx.dev)
    xch = StatsExchange(ctx)

    def sweep(ev_pair):
        ev_pair[0].record()
        cur = torch.cuda.current_stream(ctx.dev)
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            for _ in range(L_sc):
                sc.rollout(T_sc, a_sc, out=(o_sc, r_sc, d_sc))
        with torch.cuda.stream(s2):
            for _ in range(L_bg):
                bg.rollout(T_bg, a_bg, out=(o_bg, r_bg, d_bg))
        cur.wait_stream(s1); cur.wait_stream(s2)
        ev_pair[1].record()
        xch.consume(xch.submit(sc.return_stats(clear=True)))
    l0 = sc.launch_count() + bg.launch_count()
    el, km = time_steps(ctx, sweep, reps, 1, finish=xch.drain, after_enqueue=xch.reset_total)
    launches = sc.launch_count() + bg.launch_count() - l0
    steps_per_env_pair = T_sc * L_sc + T_bg * L_bg
    total_steps = n * ctx.world * steps_per_env_pair * reps
    B_sc = bytes_per_env_step(sc.n_actions, sc.n_obs)
    bytes_local = n * (T_sc * L_sc * B_sc + T_bg * L_bg * BG_BYTES)
    peak, _ = measured_peak()
    cnt, mean, _, _, _ = summarize_stats(xch.total)
    st_bg = bg.return_stats()
    out = {"workload": workload_config("c5", None, None, None, None)["workload"], "mode": mode, "envs_total": 2 * n * ctx.world,
           "envs_per_gpu_per_family": n, "scaling": "strong", "value": total_steps / (el * 1e-3), "unit": "env-steps/s",
           "ms_per_sweep": el / reps, "roofline_frac": bytes_local / (km * 1e-3) / 1e9 / peak,
           "bytes_per_env_step": {"supplychain-v0": B_sc, "beergame-v2": BG_BYTES}, "gpu_launches": launches,
           "supplychain_kernel_variant": sc.scg_info.kernel_variant,
           "episodes_finished": {"supplychain-v0 (all ranks)": cnt, "beergame-v2 (this rank)": float(st_bg[0])},
           "periods_per_launch": {"supplychain-v0": T_sc, "beergame-v2": T_bg}}
    del sc, bg
    return out


# ----------------------------------------------------------------------------------------------
# c1: beergame-v0 single env
# ----------------------------------------------------------------------------------------------
def run_c1(ctx, args):
    """One beergame-v0 env, one 35-week episode under RandomState(0).randint(0, 5, 4): the drop-in single-env API (numpy in /
    numpy out, one launch + one synchronisation per step) beside the reference's own loop on one core."""
    import numpy as np

    import gym_supplychain_b200 as gsc
    env = gsc.make("beergame-v0", device=ctx.dev)

    def episode():
        rs = np.random.RandomState(0)
        env.reset()
        ret, done, n = 0, False, 0
        while not done:
            o, r, done, _ = env.step(rs.randint(0, 5, 4))
            ret += r; n += 1
        return ret, n
    episode()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        ret, n = episode()
    dt = (time.perf_counter() - t0) / reps
    out = {"workload": workload_config("c1", None, None, None, None)["workload"], "value": n / dt, "unit": "env-steps/s",
           "us_per_step": 1e6 * dt / n, "episode_return": int(ret), "gpu_launches": reps * (n + 1),
           "note": "latency-bound by construction (one env): one kernel launch + one synchronisation per step through scg_step_host"}
    if reference_available():
        ref = _make_reference_env("beergame-v0")

        def ref_episode():
            rs = np.random.RandomState(0)
            ref.reset()
            ret, done, k = 0, False, 0
            while not done:
                o, r, done, _ = ref.step(rs.randint(0, 5, 4))
                ret += r; k += 1
            return ret, k
        ref_episode()
        t0 = time.perf_counter()
        for _ in range(reps):
            rret, rn = ref_episode()
        rdt = (time.perf_counter() - t0) / reps
        out["reference"] = {"value": rn / rdt, "unit": "env-steps/s", "us_per_step": 1e6 * rdt / rn, "episode_return": int(rret),
                            "cores": 1, "kind": "reference"}
        out["episode_return_matches_reference"] = bool(int(rret) == int(ret))
    del env
    return out


def run_large_chain(ctx, args, mode, shape=(5, 4, 7, 10), n=1 << 18, T=32, iters=3):
    """Not a BASELINE.json config: the reference's large golden-return chain (test_Nperstage.py:23-53) on the node-parallel
    split kernel (several lanes per env, warp-shuffle sums; run-time instantiated), so that its number is in the line too."""
    import gym_supplychain_b200 as gsc
    torch = ctx.torch
    env = gsc.make(ENV_ID, num_envs=n, mode=mode, seed=args.seed, auto_reset=True, device=ctx.dev, nodes_per_echelon=list(shape))
    env.reset()
    nA, nO = env.n_actions, env.n_obs
    acts = torch.empty((T, n, nA), device=ctx.dev).uniform_(-1, 1)
    obs = torch.empty((T, n, nO), device=ctx.dev); rew = torch.empty((T, n), device=ctx.dev)
    done = torch.empty((T, n), dtype=torch.uint8, device=ctx.dev)
    for _ in range(2):
        env.rollout(T, acts, out=(obs, rew, done))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        env.rollout(T, acts, out=(obs, rew, done))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    B = bytes_per_env_step(nA, nO)
    peak, _ = measured_peak()
    v = n * T / (ms * 1e-3)
    out = {"workload": "SupplyChainNPerStage(%s), %d products (A=%d, O=%d), %d envs, fused %d-period rollout" % (list(shape), env.flat.P, nA, nO, n, T),
           "mode": mode, "value": v, "unit": "env-steps/s", "ms_per_call": ms, "bytes_per_env_step": B, "roofline_frac": v * B / 1e9 / peak,
           "kernel_variant": env.scg_info.kernel_variant, "kernel": "sc_split_rollout (node-parallel)" if env.scg_info.kernel_variant == 101 else "other",
           "gpu_launches": iters + 2}
    del env, acts, obs, rew, done
    return out


def compact(res, cfg, mode):
    return {"workload": workload_config(cfg, None, None, None, None)["workload"], "mode": mode, "value": res["value"],
            "unit": "env-steps/s", "envs_total": res["n_total"], "envs_per_gpu": res["n_local"], "periods_per_step": res["T_step"],
            "periods_per_launch": res["T_launch"], "kernel_ms_per_step": res["kernel_ms"], "roofline_frac": res["frac"],
            "achieved_GBps": res["achieved_gbs"], "bytes_per_env_step": res["bytes_per_env_step"], "gpu_launches": res["launches"],
            "kernel_variant": res["kernel_variant"]}


def c4s_shape(world):
    """STRONG scaling of BASELINE.json configs[3]: 8,388,608 envs in total; a step = 128 periods of every env, in fused
    rollouts short enough for one GPU's HBM (16 periods at N=1 ... 128 at N=8: 75 GB of I/O buffers per GPU either way)."""
    n_local = (1 << 23) // world
    T_launch = max(16, 128 * world // 8)
    return n_local, T_launch, 128 // T_launch


def cpu_reference_block(cfg, cores, budget_s):
    """Bounded sample of the reference for config `cfg` on all host cores (rank 0, N=1): after a warm-up map (imports, first
    reset) each env family runs whole episodes, map after map, until `budget_s` seconds of wall time are spent on it."""
    eps_per_map = {"ntom": 8, "linear": 48, "2perstage": 24, "supplychain-v0": 2, "beergame-v0": 400, "beergame-v2-stochastic": 160}
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        per = []
        for case, T, wgt in REF_SAMPLE[cfg]:
            cpu_rate(pool, cores, case, 1, min(T, 16))
            steps, t0, maps = 0, time.perf_counter(), 0
            while time.perf_counter() - t0 < budget_s / len(REF_SAMPLE[cfg]):
                res = pool.map(_ref_worker, [(case, 1000 * maps + i, eps_per_map[case], T) for i in range(cores)])
                steps += sum(r[0] for r in res); maps += 1
            wall = time.perf_counter() - t0
            per.append((steps / wall, wgt, case, wall, maps * eps_per_map[case], T))
    return {"value": mixed_rate([(r, w) for r, w, _, _, _, _ in per]), "unit": "env-steps/s", "cores": cores, "cpu": cpu_model(),
            "kind": "reference",
            "sample": "; ".join("%s: %d workers x %d episodes x %d periods (reset() included) in %.1f s = %.3g env-steps/s"
                                % (c, cores, eps, T, wall, r) for r, _, c, wall, eps, T in per)}


# ----------------------------------------------------------------------------------------------
# our arm: main
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    ctx = Ctx(args)
    torch, world, rank = ctx.torch, ctx.world, ctx.rank
    cfg, mode = args.config, args.mode
    other = {}
    line = None
    cores = os.cpu_count() or 1
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline

    if cfg in ("c3", "c4", "c4s"):
        env_id = "supplychain-2perstage-v0" if cfg == "c3" else ENV_ID
        if cfg == "c4s":
            n_local, T_launch, lps = c4s_shape(world)
        else:
            n_local, T_launch, lps = args.envs_per_gpu, args.rollout_T, 1
        env, bufs, m = run_rollout_config(ctx, args, cfg, mode, args.steps, args.warmup, True, env_id, n_local, T_launch, lps)
        del bufs
        torch.cuda.empty_cache()
        e2e = run_e2e(ctx, args, env, n_local, m["n_actions"], m["n_obs"])
        alt = None
        if not args.no_alt:
            other_mode = "fast_f32" if mode == "mixed" else "mixed"
            del env
            torch.cuda.empty_cache()
            env, bufs, ma = run_rollout_config(ctx, args, cfg, other_mode, max(3, min(args.steps, 5)), 3, False, env_id, n_local,
                                               T_launch, lps)
            del bufs
            alt = {other_mode: {"value": ma["value"], "unit": "env-steps/s", "kernel_ms": ma["kernel_ms"], "roofline_frac": ma["frac"],
                                "note": ("all-fp32 arithmetic: per-step reward error vs the reference up to 2.2e-6 relative on ntom "
                                         "(profiles/r1_precision_survey.txt), above the 1e-6 north-star bar, hence not the headline")
                                if other_mode == "fast_f32" else "fp32 I/O, fp64 stocks and arithmetic (1e-6 bar met)"}}
        del env
        torch.cuda.empty_cache()
        traffic, traffic_src = offline_traffic(mode) if (cfg == "c4" and n_local == 1 << 20 and T_launch == 128) else (None, None)
        line = {"metric": "env_steps_per_sec", "value": m["value"], "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": m["elapsed_ms"] / args.steps, "higher_is_better": True,
                "scaling": "strong" if cfg == "c4s" else "weak", "vs_baseline": None,
                "dtype": "f32" if mode == "fast_f32" else "f32 I/O; f64 stocks and arithmetic, in-transit amounts as 32-bit fixed point",
                "data": "synthetic",
                "config": full_config(cfg, args, world),
                "clocks": m["clocks"], "e2e": e2e, "gpu_launches": m["launches"],
                "roofline": {"bound": "hbm", "achieved": m["achieved_gbs"], "peak": m["peak"], "unit": "GB/s", "frac": m["frac"],
                             "traffic": traffic, "traffic_source": ("measured_offline: ncu --set full capture of this command, " + traffic_src)
                             if traffic_src else None,
                             "kernel": kernel_name(env_id, mode), "kernel_ms": m["kernel_ms"] / lps, "peak_source": m["peak_src"],
                             "bytes_per_launch": n_local * T_launch * m["bytes_per_env_step"]},
                "alt_modes": alt, "episodes_finished": m["episodes_finished"], "mean_episode_return": m["mean_episode_return"]}
    elif cfg == "c2":
        r = run_c2(ctx, args, mode)
        peak, peak_src = measured_peak()
        f = r["fused_360_period_rollout"]
        line = {"metric": "env_steps_per_sec", "value": r["value"], "unit": "env-steps/s", "n_gpus": world, "steps": 6, "warmup": 3,
                "ms_per_step": f["ms_per_call"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 I/O; f64 stocks", "data": "synthetic", "config": full_config("c2", args, world),
                "clocks": None, "gpu_launches": r["gpu_launches"], "variants": r,
                "roofline": {"bound": "hbm", "achieved": f["roofline_frac"] * peak, "peak": peak, "unit": "GB/s", "frac": f["roofline_frac"],
                             "traffic": None, "kernel": kernel_name("supplychain-linear-v0", mode), "kernel_ms": f["ms_per_call"], "peak_source": peak_src},
                "e2e": None}
    elif cfg == "c5":
        r = run_c5(ctx, args, mode)
        line = {"metric": "env_steps_per_sec", "value": r["value"], "unit": "env-steps/s", "n_gpus": world, "steps": 1, "warmup": 1,
                "ms_per_step": r["ms_per_sweep"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "int32 (beer game) / f32 I/O, f64 state (supply chain)", "data": "synthetic",
                "config": full_config("c5", args, world), "clocks": None,
                "gpu_launches": r["gpu_launches"], "detail": r,
                "roofline": {"bound": "hbm", "frac": r["roofline_frac"], "unit": "GB/s", "traffic": None,
                             "kernel": "bg_kernel<4,4,1> + " + kernel_name("supplychain-v0", mode)}, "e2e": None}
    elif cfg == "c1":
        r = run_c1(ctx, args)
        line = {"metric": "env_steps_per_sec", "value": r["value"], "unit": "env-steps/s", "n_gpus": 1, "steps": 20, "warmup": 1,
                "ms_per_step": r["us_per_step"] * 35e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic", "config": full_config("c1", args, 1), "clocks": None, "gpu_launches": r["gpu_launches"],
                "detail": r, "roofline": None, "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 16,
                                                       "d2h_bytes_per_step": 21, "api": "scg_step_host (numpy in / numpy out)"}}

    # ---- every other BASELINE.json config, measured in the same run ----
    if not args.no_other:
        def guarded(name, fn):
            try:
                other[name] = fn()
            except Exception as e:  # noqa: BLE001
                other[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            torch.cuda.empty_cache()
        if cfg != "c4s":
            def do_c4s():
                n_local, T_launch, lps = c4s_shape(world)
                env, bufs, m4 = run_rollout_config(ctx, args, "c4s", mode, 3, 3, False, ENV_ID, n_local, T_launch, lps)
                del env, bufs
                return dict(compact(m4, "c4s", mode), scaling="strong")
            guarded("c4s", do_c4s)
        if cfg != "c5":
            guarded("c5", lambda: run_c5(ctx, args, mode))
        if world == 1:
            if cfg != "c3":
                def do_c3():
                    env, bufs, m3 = run_rollout_config(ctx, args, "c3", mode, 5, 3, False, "supplychain-2perstage-v0", 1 << 20, 128, 1)
                    del env, bufs
                    return compact(m3, "c3", mode)
                guarded("c3", do_c3)
            if cfg != "c2":
                guarded("c2", lambda: run_c2(ctx, args, mode, iters=4))
            if cfg != "c1":
                guarded("c1", lambda: run_c1(ctx, args))
            guarded("large_chain_5_4_7_10", lambda: {m: run_large_chain(ctx, args, m) for m in ("fast_f32", "mixed")})

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload, the reference itself, on the host cores ----
    cpu = None
    if want_cpu:
        if reference_available():
            cpu = cpu_reference_block(cfg, cores, args.cpu_seconds)
            if not args.no_other:
                for oc in ("c2", "c3", "c5"):
                    if oc in other and isinstance(other[oc], dict) and "error" not in other[oc]:
                        try:
                            other[oc]["cpu_baseline"] = cpu_reference_block(oc, cores, max(2.0, args.cpu_seconds / 4))
                        except Exception as e:  # noqa: BLE001
                            other[oc]["cpu_baseline"] = {"unavailable": str(e)[:200]}
        else:
            cpu = {"unavailable": "baseline/_ref missing", "kind": "reference"}
        if cfg in ("c4", "c4s"):
            # the two restatements on the same ntom sample, for context (the checkers, not the baseline)
            ctxp = mp.get_context("fork")
            with ctxp.Pool(cores) as pool:
                try:
                    pool.map(_port_worker, [(i, 1, 16) for i in range(cores)])
                    t0 = time.perf_counter()
                    res = pool.map(_port_worker, [(i, 4, 128) for i in range(cores)])
                    cpu["port"] = {"value": sum(r[0] for r in res) / (time.perf_counter() - t0), "unit": "env-steps/s", "cores": cores,
                                   "what": "oracle/sc_oracle.py (Python restatement of the reference loop)"}
                    from oracle import c_oracle
                    c_oracle.build()
                    pool.map(_c_port_worker, [(i, 2, 128) for i in range(cores)])
                    t0 = time.perf_counter()
                    res = pool.map(_c_port_worker, [(100 + i, 1500, 128) for i in range(cores)])
                    cpu["c_port"] = {"value": sum(r[0] for r in res) / (time.perf_counter() - t0), "unit": "env-steps/s", "cores": cores,
                                     "what": "oracle/sc_oracle.c (plain C restatement, gcc -O2), one process per core"}
                except Exception as e:  # noqa: BLE001
                    cpu["ports_unavailable"] = str(e)[:200]

    if rank == 0:
        line["cpu_baseline"] = cpu
        line["other_configs"] = other if other else None
        emit_line(line)
    if world > 1:
        ctx.dist.destroy_process_group()


_JSON_OUT = None


def emit_line(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner goes to
    fd 1) has been redirected to stderr by main()."""
    (_JSON_OUT or sys.stdout).write(json.dumps(line) + "\n")
    (_JSON_OUT or sys.stdout).flush()


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=["c1", "c2", "c3", "c4", "c4s", "c5"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--rollout-T", type=int, default=128)
    ap.add_argument("--e2e-T", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="wall-time budget of the CPU (reference) baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the other BASELINE.json configs (other_configs block)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--mode", default="mixed", choices=["fast_f32", "mixed"])
    ap.add_argument("--no-alt", action="store_true", help="skip the secondary measurement of the other fp32-I/O mode")
    args = ap.parse_args(argv)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    return args


def main():
    global _JSON_OUT
    args = parse_args()
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                                   # fd 1 -> stderr for the rest of the process (and its children)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
