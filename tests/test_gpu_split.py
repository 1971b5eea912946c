"""The node-parallel ("split") rollout kernel for layered chains whose state overflows one thread's registers
(csrc/scg_sc_split.cuh): G lanes per env, shipments per destination summed with warp shuffles.

Shapes: the reference's own large golden-return chains, SupplyChainNPerStage([5,4,7,10]) and 10 per stage
(gym_supplychain/envs/tests/test_Nperstage.py:23-53, built at supplychain_Nperstage_env.py:37-66).  Checked against the
plain-C oracle on every env and period (mixed <= 1e-6 relative on rewards = the north-star bar, fast_f32 <= 2e-5), against
the one-thread-per-env kernel it replaces (SCG_SPLIT_G=0), and for the properties every rollout kernel must have: ragged
batches, chunked rollouts bit-identical to one rollout, step() == rollout(), auto-reset and return statistics."""
import os

import numpy as np
import pytest
import torch

import gym_supplychain_b200 as gsc
from gym_supplychain_b200 import presets
from oracle.c_oracle import CSupplyChainOracle
from test_gpu_c_oracle import _gpu_run, _inputs, _oracle_run

pytestmark = pytest.mark.gpu
BIG = [5, 4, 7, 10]


@pytest.fixture(autouse=True)
def _default_split(monkeypatch):
    monkeypatch.delenv("SCG_SPLIT_G", raising=False)
    monkeypatch.delenv("SCG_SPLIT_MINB", raising=False)


@pytest.mark.parametrize("mode,rtol,oatol", [("mixed", 1e-6, 5e-7), ("fast_f32", 2e-5, 5e-6)])
@pytest.mark.parametrize("shape,n", [(BIG, 101), (10, 67)])
def test_split_kernel_whole_batch_vs_c_oracle(shape, n, mode, rtol, oatol):
    """Every env (an env count that is not a multiple of the envs per warp), 150 periods, out-of-range actions included."""
    T = 150
    info, okw = presets.n_per_stage(nodes_per_echelon=shape, total_time_steps=T)
    cc = CSupplyChainOracle(info, **okw)
    dem, lts, a = _inputs(np.random.RandomState(16), cc, n, T, 0, 400, wide=True)
    ref = _oracle_run(cc, dem, lts, a, canonical=True)
    env = gsc.make("supplychain-ntom-v0", num_envs=n, mode=mode, auto_reset=False, nodes_per_echelon=shape, total_time_steps=T)
    assert env.scg_info.kernel_variant == 101                    # run-time instantiated, node-parallel
    got = _gpu_run(env, dem, lts, a)
    assert env.scg_info.kernel_variant == 101, "the run-time instantiation failed and fell back"
    assert np.array_equal(got[2], ref[2]) and ref[2][-1].all()
    rel = np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1.0)
    assert rel.max() <= rtol, (mode, rel.max(), np.unravel_index(rel.argmax(), rel.shape))
    assert np.abs(got[0] - ref[0]).max() <= oatol
    assert np.allclose(got[1].sum(0), ref[1].sum(0), rtol=1e-6)


@pytest.mark.parametrize("g", ["2", "8"])
def test_other_lane_counts_agree_with_the_oracle(g, monkeypatch):
    monkeypatch.setenv("SCG_SPLIT_G", g)
    T, n = 60, 37
    info, okw = presets.n_per_stage(nodes_per_echelon=BIG, total_time_steps=T)
    cc = CSupplyChainOracle(info, **okw)
    dem, lts, a = _inputs(np.random.RandomState(3), cc, n, T, 0, 400)
    ref = _oracle_run(cc, dem, lts, a, canonical=True)
    env = gsc.make("supplychain-ntom-v0", num_envs=n, mode="mixed", auto_reset=False, nodes_per_echelon=BIG, total_time_steps=T)
    got = _gpu_run(env, dem, lts, a)
    assert env.scg_info.kernel_variant == 101
    rel = np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1.0)
    assert rel.max() <= 1e-6 and np.abs(got[0] - ref[0]).max() <= 5e-7 and np.array_equal(got[2], ref[2])


@pytest.mark.parametrize("mode", ["mixed", "fast_f32"])
def test_split_kernel_agrees_with_the_one_thread_per_env_kernel(mode, monkeypatch):
    """Same seed, Philox demand generated in-kernel, auto-reset inside the rollout: the split kernel and the kernel it
    replaces draw the same demands (integers) and compute the same trajectory up to summation order."""
    n, T = 500, 40
    a = gsc.make("supplychain-ntom-v0", num_envs=n, seed=7, mode=mode, auto_reset=True, total_time_steps=25, nodes_per_echelon=BIG)
    monkeypatch.setenv("SCG_SPLIT_G", "0")
    b = gsc.make("supplychain-ntom-v0", num_envs=n, seed=7, mode=mode, auto_reset=True, total_time_steps=25, nodes_per_echelon=BIG)
    assert a.scg_info.kernel_variant == 101 and b.scg_info.kernel_variant == 100
    assert torch.equal(a.reset(), b.reset())
    acts = torch.rand(T, n, a.n_actions, device="cuda") * 2 - 1
    o2, r2, d2 = b.rollout(T, acts)
    monkeypatch.delenv("SCG_SPLIT_G")
    o1, r1, d1 = a.rollout(T, acts)
    assert torch.equal(d1, d2) and int(d1.sum()) == n
    assert torch.allclose(r1, r2, rtol=2e-5 if mode == "fast_f32" else 1e-6, atol=1e-2)
    assert torch.allclose(o1, o2, rtol=0, atol=2e-5 if mode == "fast_f32" else 1e-6)
    nd = a.flat.n_retailers * a.flat.P
    assert torch.equal(((o1[..., :nd].double() + 1) * 200).round(), ((o2[..., :nd].double() + 1) * 200).round())
    sa, sb = a.return_stats(), b.return_stats()
    assert sa[0] == sb[0] == n and torch.allclose(sa[1:], sb[1:], rtol=1e-5)


@pytest.mark.parametrize("mode", ["mixed", "fast_f32"])
def test_split_kernel_chunked_rollouts_steps_and_random_policy(mode):
    """60 periods in launches of 37 + 1 + 22 (auto-reset inside, an odd env count) are bit-identical to one launch of 60 and
    to 60 step() calls; the state blob round-trips; the in-kernel random policy runs and is deterministic."""
    n, T = 203, 60
    mk = lambda: gsc.make("supplychain-ntom-v0", num_envs=n, seed=3, mode=mode, auto_reset=True, total_time_steps=25, nodes_per_echelon=BIG)
    a, b, c = mk(), mk(), mk()
    oa, ob, oc = a.reset(), b.reset(), c.reset()
    assert torch.equal(oa, ob) and torch.equal(oa, oc)
    acts = torch.rand(T, n, a.n_actions, device="cuda") * 2.4 - 1.2
    o1, r1, d1 = a.rollout(T, acts)
    parts = [b.rollout(37, acts[:37]), b.rollout(1, acts[37:38]), b.rollout(22, acts[38:])]
    o2, r2, d2 = (torch.cat([p[i] for p in parts]) for i in range(3))
    assert torch.equal(d1, d2) and torch.equal(r1, r2) and torch.equal(o1, o2)
    assert torch.equal(a.get_state(), b.get_state())
    for t in range(T):
        o3, r3, d3, _ = c.step(acts[t])
        assert torch.equal(o1[t], o3) and torch.equal(r1[t], r3) and torch.equal(d1[t].bool(), d3)
    assert a.scg_info.kernel_variant == 101
    ra, rb = a.rollout(30), b.rollout(30)                           # in-kernel uniform random actions
    assert all(torch.equal(x, y) for x, y in zip(ra, rb)) and torch.isfinite(ra[1]).all() and float(ra[0].abs().max()) <= 1.0


def _randomise_constants(rs, info, P, L):
    """Per-NODE and per-EDGE random constants (the presets give every node of an echelon the same numbers, which would hide
    a wrong node / edge index in the lanes' constant lookups).  Factories keep ship capacities they cannot exceed (the F4a
    regime only `exact_f64` can follow, see test_gpu_c_oracle._random_layered_chain)."""
    ri = lambda lo, hi, n: [int(x) for x in rs.randint(lo, hi, n)]              # noqa: E731
    for nd in info.values():
        nd["stock_capacity"] = ri(400, 3000, P)
        nd["initial_stock"] = [int(rs.randint(0, c)) for c in nd["stock_capacity"]]
        nd["stock_cost"] = ri(1, 4, P)
        if "supply_capacity" in nd:
            nd["supply_capacity"], nd["supply_cost"] = ri(100, 900, P), ri(1, 9, P)
            nd["initial_supply"] = [ri(0, 500, L) for _ in range(P)]
        else:
            nd["initial_shipments"] = [ri(0, 500, L) for _ in range(P)]
        if "processing_capacity" in nd:
            nd["processing_capacity"], nd["processing_cost"] = int(rs.randint(300, 2500)), ri(1, 15, P)
        if "destinations" in nd:
            D = len(nd["destinations"])
            nd["dest_costs"], nd["ship_capacity"] = [ri(1, 6, D) for _ in range(P)], ri(150, 1500, D)
            if "processing_capacity" in nd:
                nd["ship_capacity"] = [c + 2 * max(nd["stock_capacity"]) for c in nd["ship_capacity"]]


@pytest.mark.parametrize("shape,P,L,seed", [(BIG, 2, 2, 1), ([6, 5, 9, 7], 2, 2, 2), ([4, 6, 5, 8], 3, 2, 3), ([7, 3, 6, 9], 2, 3, 4)])
@pytest.mark.parametrize("mode,rtol,oatol", [("mixed", 1e-6, 5e-7), ("fast_f32", 2e-5, 5e-6)])
def test_split_kernel_with_random_per_node_constants_vs_c_oracle(shape, P, L, seed, mode, rtol, oatol):
    """Large layered chains with every capacity, cost, initial stock and initial shipment drawn per node / per edge, a
    per-product processing ratio, three products, lead time 3: the lanes' constant lookups (node, product, edge) and the
    ownership of echelons that are not a multiple of the lane count are all exercised."""
    rs = np.random.RandomState(7000 + seed)
    T, n = 40, 45
    info, kw = presets.n_per_stage(nodes_per_echelon=shape, num_products=P, avg_leadtime=L, max_leadtime=L, total_time_steps=T,
                                   demand_range=(0, int(rs.randint(50, 400))), unmet_demand_cost=int(rs.randint(20, 300)),
                                   exceeded_stock_capacity_cost=int(rs.randint(2, 20)), exceeded_process_capacity_cost=int(rs.randint(2, 20)),
                                   exceeded_ship_capacity_cost=int(rs.randint(2, 20)), processing_ratio=[int(x) for x in rs.randint(1, 4, P)])
    _randomise_constants(rs, info, P, L)
    okw = {k: v for k, v in kw.items() if k not in ("seed", "build_info")}
    cc = CSupplyChainOracle(info, **okw)
    dem, lts, a = _inputs(rs, cc, n, T, kw["demand_range"][0], kw["demand_range"][1], wide=(seed % 2 == 0))
    ref = _oracle_run(cc, dem, lts, a, canonical=True)
    env = gsc.SupplyChainEnv(info, num_envs=n, mode=mode, auto_reset=False, **okw)
    assert env.scg_info.kernel_variant == 101, shape
    got = _gpu_run(env, dem, lts, a)
    assert env.scg_info.kernel_variant == 101, "the run-time instantiation failed and fell back"
    assert np.array_equal(got[2], ref[2])
    rel = np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1.0)
    assert rel.max() <= rtol, (shape, mode, rel.max(), np.unravel_index(rel.argmax(), rel.shape))
    assert np.abs(got[0] - ref[0]).max() <= oatol, (shape, mode)


def test_split_kernel_is_shard_invariant_and_sane_at_size():
    """65,536 envs of 5-4-7-10 with Philox demand and auto-reset: the same global envs simulated by ONE handle or by two shards
    (first_env = n/2 on the second) give bit-identical observations, rewards and done flags -- the property the multi-GPU
    sharding rests on -- and the trajectories are sane (obs in [-1, 1], costs only, one episode end per env)."""
    n, T = 1 << 16, 30
    kw = dict(seed=11, mode="fast_f32", auto_reset=True, total_time_steps=25, nodes_per_echelon=BIG)
    whole = gsc.make("supplychain-ntom-v0", num_envs=n, **kw)
    lo = gsc.make("supplychain-ntom-v0", num_envs=n // 2, first_env=0, **kw)
    hi = gsc.make("supplychain-ntom-v0", num_envs=n // 2, first_env=n // 2, **kw)
    assert whole.scg_info.kernel_variant == 101
    o0 = whole.reset()
    assert torch.equal(o0, torch.cat([lo.reset(), hi.reset()]))
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    acts = torch.rand((T, n, whole.n_actions), device="cuda", generator=g) * 2 - 1
    o, r, d = whole.rollout(T, acts)
    o1, r1, d1 = lo.rollout(T, acts[:, :n // 2].contiguous())
    o2, r2, d2 = hi.rollout(T, acts[:, n // 2:].contiguous())
    assert torch.equal(o, torch.cat([o1, o2], dim=1)) and torch.equal(r, torch.cat([r1, r2], dim=1)) and torch.equal(d, torch.cat([d1, d2], dim=1))
    assert float(o.abs().max()) <= 1.0 and bool((r <= 0).all()) and torch.isfinite(r).all()
    assert int(d.sum()) == n and bool(d[24].all())                       # every env finishes its 25-period episode once
    st = whole.return_stats()
    assert int(st[0]) == n and float(st[4]) < 0                           # n finished episodes, all returns negative
