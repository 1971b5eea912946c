"""Histograms of the reference's OWN stochastic inputs -- TEST INFRASTRUCTURE (build container only).

    python oracle/make_demand_fixtures.py        # rewrites tests/golden/generator_hist.npz

The kernels draw demands and lead times from counter-based Philox streams, the reference from one MT19937 stream
(demands_generator.py:3-89, supplychain_env.py:664-672), so parity of the generators is DISTRIBUTIONAL: the GPU tests
(tests/test_gpu_generators.py) compare histograms of ~1e6 in-kernel draws with the histograms recorded here from
samples of the reference's `generate_demand` (all four kinds: uniform, clipped normal, sinusoid + normal, sinusoid +
uniform perturbation) and of the lead-time tables a reference env draws at reset().  The cases are chosen so that the
clip boundaries carry visible mass.  Per case the file holds, for each recorded row (period), the counts over the
integer support [lo, hi]; plus the counts of the FIRST call alone, which tests/test_generator_fixtures.py regenerates
live to pin this file to the reference.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SCG_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "gym_shim"))
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "generator_hist.npz")

T = 360
# name -> kwargs of SupplyChain2perStageEnv that select the generator (supplychain_env.py:641-661), recorded rows
CASES = {
    "uniform": dict(kw=dict(demand_range=(0, 400)), rows=None, calls=8, width=400),
    "normal": dict(kw=dict(demand_range=(10, 60), demand_std=12), rows=None, calls=8, width=400),
    "sin_normal": dict(kw=dict(demand_range=(0, 400), demand_std=40, demand_sen_peaks=4, avg_demand_range=(50, 350),
                               demand_perturb_norm=True), rows=[0, 22, 45, 67, 200, 360], calls=10, width=20000),
    "sin_uniform": dict(kw=dict(demand_range=(0, 400), demand_std=10, demand_sen_peaks=4, avg_demand_range=(50, 380),
                                demand_perturb_norm=False), rows=[0, 22, 45, 67, 200, 360], calls=10, width=20000),
}
LEADTIME_CASES = {"lt_avg2_max4": (2, 4), "lt_avg3_max6": (3, 6)}


def demand_call(rs, case):
    """One call of the reference generator, argument for argument as SupplyChainEnv.reset makes it (:641-648)."""
    from gym_supplychain.envs.demands_generator import generate_demand
    kw, width = case["kw"], case["width"]
    lo, hi = kw["demand_range"]
    avg = kw.get("avg_demand_range") or (None, None)
    return generate_demand(rs, (T + 1, width), T, lo, hi, std=kw.get("demand_std"), sen_peaks=kw.get("demand_sen_peaks"),
                           minavg=avg[0], maxavg=avg[1], perturb_norm=kw.get("demand_perturb_norm", True))


def hist_rows(data, rows, lo, hi):
    if rows is None:
        return np.bincount(data.ravel() - lo, minlength=hi - lo + 1)[None]
    return np.stack([np.bincount(data[r] - lo, minlength=hi - lo + 1) for r in rows])


def demand_hist(name, calls=None):
    case = CASES[name]
    lo, hi = case["kw"]["demand_range"]
    rs = np.random.RandomState(1234)
    total, first = None, None
    for k in range(calls or case["calls"]):
        h = hist_rows(demand_call(rs, case), case["rows"], lo, hi)
        first = h if first is None else first
        total = h if total is None else total + h
    return total, first


def leadtime_hist(avg, mx, resets=40):
    """Lead-time tables as a reference env draws them at reset() (:664-672)."""
    import gym_supplychain.envs as E
    env = E.SupplyChain2perStageEnv(stochastic_leadtimes=True, avg_leadtime=avg, max_leadtime=mx, seed=99)
    tot = np.zeros(mx + 1, dtype=np.int64)
    for _ in range(resets):
        env.reset()
        tot += np.bincount(np.asarray(env.leadtimes).ravel().astype(int), minlength=mx + 1)
    return tot


def main():
    out = {}
    meta = {"T": T, "cases": {k: dict(kw=v["kw"], rows=v["rows"]) for k, v in CASES.items()},
            "leadtimes": {k: list(v) for k, v in LEADTIME_CASES.items()}}
    for name in CASES:
        tot, first = demand_hist(name)
        out[name], out[name + "__first_call"] = tot, first
        print(name, tot.shape, int(tot.sum()))
    for name, (avg, mx) in LEADTIME_CASES.items():
        out[name] = leadtime_hist(avg, mx)
        print(name, out[name])
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
