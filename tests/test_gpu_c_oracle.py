"""GPU parity at scale against the plain-C oracle (oracle/sc_oracle.c, itself pinned bit-for-bit on the reference goldens):
EVERY env of a batch with its own demand table, lead times and actions, whole episodes -- not sampled replays.

  exact_f64 ........ obs / reward / done bit-exact (layered presets AND random non-layered DAGs, stochastic lead times)
  mixed ............ rewards within 1e-6 relative (north-star bar), canonical obs within 5e-7
  fast_f32 ......... rewards within 2e-5 relative (fp32 state: tail of the per-step error, profiles/r1_precision_survey.txt)
  run-time instantiated kernels (5-4-7-10) held to the same bars."""
import numpy as np
import pytest
import torch

import gym_supplychain_b200 as gsc
from gym_supplychain_b200 import presets
from oracle.c_oracle import CSupplyChainOracle
from test_c_oracle import _random_chain

pytestmark = pytest.mark.gpu


def _inputs(rs, cc, n, T, lo, hi, wide=False):
    dem = rs.randint(lo, hi + 1, size=(n, T + 1, cc.n_retailers, cc.P)).astype(np.int32)
    lts = rs.randint(1, 5, size=(n, T, cc.n_leadtimes)).astype(np.int32) if cc.stoch_lt else None
    a = rs.uniform(-1.2 if wide else -1.0, 1.2 if wide else 1.0, size=(T, n, cc.n_actions)).astype(np.float32)
    u = rs.uniform(size=a.shape)
    a[u < 0.04] = 1.0; a[(u >= 0.04) & (u < 0.08)] = -1.0; a[(u >= 0.08) & (u < 0.10)] = 0.0
    return dem, lts, a


def _oracle_run(cc, dem, lts, a, canonical):
    T, n = a.shape[0], a.shape[1]
    obs = np.empty((T + 1, n, cc.n_obs)); rew = np.empty((T, n)); done = np.empty((T, n), dtype=bool)
    for e in range(n):
        cc.reset(dem[e], None if lts is None else lts[e])
        obs[0, e] = cc.observation(canonical=canonical)
        o, r, d = cc.rollout(a[:, e].astype(np.float64), canonical=canonical)
        obs[1:, e], rew[:, e], done[:, e] = o, r, d.astype(bool)
    return obs, rew, done


def _gpu_run(env, dem, lts, a):
    env.inject(demand=dem, leadtimes=lts)
    o0 = env.reset()
    acts = torch.from_numpy(a).to(device="cuda", dtype=env._io_dtype)
    o, r, d = env.rollout(a.shape[0], acts)
    return (np.concatenate([o0.double().cpu().numpy()[None], o.double().cpu().numpy()]), r.double().cpu().numpy(),
            d.cpu().numpy().astype(bool))


def test_exact_f64_whole_batch_bit_exact_vs_c_oracle_ntom_full_episode():
    info, kw = presets.n_per_stage(nodes_per_echelon=3)
    cc = CSupplyChainOracle(info, **kw)
    n, T = 768, 360
    dem, lts, a = _inputs(np.random.RandomState(5), cc, n, T, 0, 400)
    ref = _oracle_run(cc, dem, lts, a, canonical=False)
    env = gsc.make("supplychain-ntom-v0", num_envs=n, mode="exact_f64", auto_reset=False)
    got = _gpu_run(env, dem, lts, a)
    assert np.array_equal(got[2], ref[2]) and ref[2][-1].all() and not ref[2][:-1].any()
    assert np.array_equal(got[1], ref[1])
    assert np.array_equal(got[0], ref[0])


@pytest.mark.parametrize("mode,rtol,oatol", [("mixed", 1e-6, 5e-7), ("fast_f32", 2e-5, 5e-6)])
@pytest.mark.parametrize("env_id,kw", [("supplychain-ntom-v0", {}), ("supplychain-2perstage-v0", {}),
                                       ("supplychain-ntom-v0", {"nodes_per_echelon": [5, 4, 7, 10]})])
def test_fp32_io_modes_whole_batch_vs_c_oracle_full_episode(env_id, kw, mode, rtol, oatol):
    """Rewards of every env and period within the mode's bar (relative to max(|reference reward|, 1)); canonical obs."""
    if "2perstage" in env_id:
        info, okw = presets.two_per_stage()
        lo, hi = 10, 20
    else:
        info, okw = presets.n_per_stage(**kw)
        lo, hi = 0, 400
    cc = CSupplyChainOracle(info, **okw)
    n, T = (96 if kw else 512), 360
    dem, lts, a = _inputs(np.random.RandomState(6), cc, n, T, lo, hi)
    ref = _oracle_run(cc, dem, lts, a, canonical=True)
    env = gsc.make(env_id, num_envs=n, mode=mode, auto_reset=False, **kw)
    assert env.scg_info.kernel_variant > 0
    got = _gpu_run(env, dem, lts, a)
    assert np.array_equal(got[2], ref[2])
    rel = np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1.0)
    assert rel.max() <= rtol, (mode, rel.max(), np.unravel_index(rel.argmax(), rel.shape))
    assert np.abs(got[0] - ref[0]).max() <= oatol
    assert np.allclose(got[1].sum(0), ref[1].sum(0), rtol=1e-6)          # episode returns


@pytest.mark.parametrize("seed", range(6))
def test_exact_f64_bit_exact_on_random_non_layered_dags(seed):
    """Random DAGs (not layered, per-product constants, one processing node, stochastic lead times on half of them,
    out-of-range actions): the runtime-topology kernel in exact mode reproduces the C oracle bit for bit, every env."""
    rs = np.random.RandomState(2000 + seed)
    info, kw = _random_chain(rs)
    cc = CSupplyChainOracle(info, **kw)
    n, T = 160, kw["total_time_steps"]
    dem, lts, a = _inputs(rs, cc, n, T, kw["demand_range"][0], kw["demand_range"][1], wide=True)
    ref = _oracle_run(cc, dem, lts, a, canonical=False)
    env = gsc.SupplyChainEnv(info, num_envs=n, mode="exact_f64", auto_reset=False, **kw)
    assert env.scg_info.kernel_variant == 0
    got = _gpu_run(env, dem, lts, a)
    assert np.array_equal(got[2], ref[2])
    assert np.array_equal(got[1], ref[1]), seed
    assert np.array_equal(got[0], ref[0]), seed
    # the fp64 ring kernel on the same DAG: same dynamics up to the summation order inside a slot
    fast = gsc.SupplyChainEnv(info, num_envs=n, mode="fast_f64", auto_reset=False, **kw)
    g2 = _gpu_run(fast, dem, lts, a)
    assert np.allclose(g2[1], ref[1], rtol=1e-12, atol=1e-9)
    # and with fp32 I/O (mixed): the north-star bar on every reward, canonical observation
    canon = _oracle_run(cc, dem, lts, a, canonical=True)
    g3 = _gpu_run(gsc.SupplyChainEnv(info, num_envs=n, mode="mixed", auto_reset=False, **kw), dem, lts, a)
    rel = np.abs(g3[1] - canon[1]) / np.maximum(np.abs(canon[1]), 1.0)
    assert rel.max() <= 1e-6 and np.abs(g3[0] - canon[0]).max() <= 5e-7 and np.array_equal(g3[2], canon[2])


def _random_layered_chain(rs, tight_factory_ship=False):
    """A random layered full-bipartite chain (1-4 nodes per echelon, 1-3 products, lead time 1-3, constant or stochastic)
    with PER-NODE random constants: the shapes the run-time instantiation has to get right.

    Factories get ship capacities they cannot exceed unless `tight_factory_ship`: a ship clip at a factory leaves a
    NEGATIVE capacity for the next product (the reference subtracts raw-material units), after which any
    positive amount -- including a 1e-14 rounding residue of a just-emptied stock -- is "clipped" to that negative
    capacity and hundreds of units appear.  The reference's own trajectory is then decided by the last bit of its
    floating-point sums, which only the bit-exact mode can follow."""
    cnt = [int(rs.randint(1, 5)) for _ in range(4)]
    P, L = int(rs.randint(1, 4)), int(rs.randint(1, 4))
    stoch = bool(rs.randint(0, 2))
    maxL = L + int(rs.randint(1, 3)) if stoch else L
    info, kw = presets.n_per_stage(nodes_per_echelon=cnt, num_products=P, avg_leadtime=L, max_leadtime=maxL,
                                   stochastic_leadtimes=stoch, total_time_steps=50, demand_range=(0, int(rs.randint(50, 400))),
                                   unmet_demand_cost=int(rs.randint(20, 300)), exceeded_stock_capacity_cost=int(rs.randint(2, 20)),
                                   exceeded_process_capacity_cost=int(rs.randint(2, 20)), exceeded_ship_capacity_cost=int(rs.randint(2, 20)),
                                   processing_ratio=int(rs.randint(1, 4)))
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
            if "processing_capacity" in nd and not tight_factory_ship:
                nd["ship_capacity"] = [c + 2 * max(nd["stock_capacity"]) for c in nd["ship_capacity"]]
    return info, kw, cnt


@pytest.mark.parametrize("seed,tight", [(s, False) for s in range(8)] + [(0, True), (23, False), (27, False)])
def test_runtime_instantiated_kernels_on_random_layered_chains_vs_c_oracle(seed, tight):
    """(0, tight) is a regression: product 0 ship-clipped at a factory (negative capacity left), product 1 with the
    processing budget exactly exhausted -- its amount must be exactly 0, not a residue, or it is clipped as well."""
    rs = np.random.RandomState(3000 + seed)
    info, kw, cnt = _random_layered_chain(rs, tight_factory_ship=tight)
    okw = {k: v for k, v in kw.items() if k not in ("seed", "build_info")}
    cc = CSupplyChainOracle(info, **okw)
    n, T = 96, kw["total_time_steps"]
    dem, lts, a = _inputs(rs, cc, n, T, kw["demand_range"][0], kw["demand_range"][1], wide=(seed % 2 == 1))
    if lts is not None:
        lts = rs.randint(1, kw["max_leadtime"] + 1, size=lts.shape).astype(np.int32)
    ref = _oracle_run(cc, dem, lts, a, canonical=True)
    for mode, rtol, oatol in (("mixed", 1e-6, 5e-7), ("fast_f64", 1e-11, 1e-11)):
        env = gsc.SupplyChainEnv(info, num_envs=n, mode=mode, auto_reset=False, **okw)
        assert env.scg_info.kernel_variant > 0, (cnt, kw["num_products"])          # pre-built or instantiated at run time
        got = _gpu_run(env, dem, lts, a)
        assert env.scg_info.kernel_variant > 0, "the run-time instantiation failed and fell back"
        assert np.array_equal(got[2], ref[2])
        rel = np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1.0)
        assert rel.max() <= rtol, (seed, mode, cnt, rel.max())
        assert np.abs(got[0] - ref[0]).max() <= oatol, (seed, mode)


AOT_SHAPES = [dict(nodes_per_echelon=[1, 1, 1, 1], num_products=1), dict(nodes_per_echelon=[2, 2, 2, 2], num_products=1),
              dict(nodes_per_echelon=[2, 2, 2, 2], num_products=2), dict(nodes_per_echelon=[3, 3, 3, 3], num_products=2),
              dict(nodes_per_echelon=[1, 1, 1, 3], num_products=1),
              dict(nodes_per_echelon=[2, 2, 2, 2], num_products=1, stochastic_leadtimes=True, avg_leadtime=2, max_leadtime=4)]


@pytest.mark.parametrize("k", range(len(AOT_SHAPES)))
@pytest.mark.parametrize("mode,rtol,oatol", [("mixed", 1e-6, 5e-7), ("fast_f32", 5e-5, 5e-6)])
def test_prebuilt_kernels_with_random_per_node_constants_vs_c_oracle(k, mode, rtol, oatol):
    """The six pre-built chain shapes with per-node RANDOM constants (capacities, costs, initial pipeline), lead time 2,
    out-of-range actions on the odd ones: the kernels only specialise on the shape, every constant is a parameter."""
    rs = np.random.RandomState(4000 + k)
    shape = AOT_SHAPES[k]
    info, kw = presets.n_per_stage(total_time_steps=60, demand_range=(0, int(rs.randint(50, 400))),
                                   unmet_demand_cost=int(rs.randint(20, 300)), processing_ratio=int(rs.randint(1, 4)), **shape)
    P, L = kw["num_products"], kw["avg_leadtime"]
    ri = lambda lo, hi, n: [int(x) for x in rs.randint(lo, hi, n)]              # noqa: E731
    for nd in info.values():
        nd["stock_capacity"] = ri(400, 3000, P)
        nd["initial_stock"] = [int(rs.randint(0, c)) for c in nd["stock_capacity"]]
        nd["stock_cost"] = ri(1, 4, P)
        if "supply_capacity" in nd:
            nd["supply_capacity"], nd["supply_cost"], nd["initial_supply"] = ri(100, 900, P), ri(1, 9, P), [ri(0, 500, L) for _ in range(P)]
        else:
            nd["initial_shipments"] = [ri(0, 500, L) for _ in range(P)]
        if "destinations" in nd:
            D = len(nd["destinations"])
            nd["dest_costs"], nd["ship_capacity"] = [ri(1, 6, D) for _ in range(P)], ri(150, 1500, D)
        if "processing_capacity" in nd:          # factories: tight processing, ship capacity they cannot exceed (see above)
            nd["processing_capacity"], nd["processing_cost"] = int(rs.randint(300, 2500)), ri(1, 15, P)
            nd["ship_capacity"] = [c + 2 * max(nd["stock_capacity"]) for c in nd["ship_capacity"]]
    okw = {x: v for x, v in kw.items() if x not in ("seed", "build_info")}
    cc = CSupplyChainOracle(info, **okw)
    n, T = 160, kw["total_time_steps"]
    dem, lts, a = _inputs(rs, cc, n, T, kw["demand_range"][0], kw["demand_range"][1], wide=(k % 2 == 1))
    ref = _oracle_run(cc, dem, lts, a, canonical=True)
    env = gsc.SupplyChainEnv(info, num_envs=n, mode=mode, auto_reset=False, **okw)
    assert 1 <= env.scg_info.kernel_variant <= 6, env.scg_info.kernel_variant
    got = _gpu_run(env, dem, lts, a)
    assert np.array_equal(got[2], ref[2])
    rel = np.abs(got[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1.0)
    assert rel.max() <= rtol, (k, mode, rel.max())
    assert np.abs(got[0] - ref[0]).max() <= oatol


@pytest.mark.parametrize("seed", range(8))
def test_beergame_random_configs_whole_batch_bit_exact_vs_c_oracle(seed):
    """Random beer games (2-6 levels, v0 and v2, per-env demand and delay tables 0..4, random costs / inventories):
    every env, every week, bit for bit against the C oracle."""
    from oracle.c_oracle import CBeerGameOracle
    rs = np.random.RandomState(5000 + seed)
    version = 0 if seed % 2 == 0 else 2
    L, T, n = int(rs.randint(2, 7)), int(rs.randint(10, 40)), 300
    inv0 = [int(x) for x in rs.randint(0, 20, L)]
    kw = dict(inv_cost=int(rs.randint(1, 4)), backlog_cost=int(rs.randint(1, 5)))
    ship, orders = int(rs.randint(0, 8)), int(rs.randint(0, 8))
    max_delay = int(rs.choice([2, 4]))
    dem = rs.randint(0, 17, size=(n, T)).astype(np.int32)
    dl = rs.randint(0, max_delay + 1, size=(n, T)).astype(np.int32)
    if version == 0:
        env = gsc.BeerGameEnv({"levels": L, "customer_demand": [4] * T, "shipment_delays": [max_delay] * T, "initial_inventory": inv0,
                               "initial_shipment_value": ship, "initial_orders_value": orders, **kw}, num_envs=n, auto_reset=False)
        acts = rs.randint(-3, 6, size=(T, n, L)).astype(np.int32)
        ora = CBeerGameOracle(version=0, levels=L, weeks=T, initial_inventory=inv0, initial_shipment=ship, initial_orders=orders, **kw)
    else:
        ms, pen = int(rs.randint(20, 120)), int(rs.randint(1, 200))
        env = gsc.BeerGameEnv2(max_stock=ms, weeks=T, levels=L, customer_demand=[4] * T, initial_inventory=inv0,
                               exceeded_capacity_penalty=pen, shipment_delays=[max_delay] * T, initial_shipment=ship,
                               initial_orders=orders, num_envs=n, auto_reset=False, **kw)
        acts = rs.randint(0, 30, size=(T, n, L)).astype(np.int32)
        ora = CBeerGameOracle(version=2, levels=L, weeks=T, initial_inventory=inv0, initial_shipment=ship, initial_orders=orders,
                              max_stock=ms, exceeded_capacity_penalty=pen, **kw)
    env.inject(demand=dem, delays=dl)
    o0 = env.reset()
    o, r, d = env.rollout(T, torch.from_numpy(acts).cuda())
    o0, o, r, d = o0.cpu().numpy(), o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy().astype(bool)
    for e in range(n):
        obs = ora.reset(dem[e], dl[e])
        assert np.array_equal(obs, o0[e])
        for t in range(T):
            obs, rew, done, _ = ora.step(acts[t, e])
            assert np.array_equal(obs, o[t, e]) and rew == r[t, e] and done == d[t, e], (seed, e, t)


@pytest.mark.parametrize("seed", range(0, 40, 1))
def test_exact_f64_bit_exact_on_random_layered_chains_with_tight_factories(seed):
    """exact_f64 follows the reference bit for bit on every random chain, including the ones where a ship clip at a factory
    leaves a negative capacity and the trajectory hinges on a 1e-14 residue (seeds 23 and 27: the cases no tolerance
    mode can follow)."""
    rs = np.random.RandomState(3000 + seed)
    info, kw, cnt = _random_layered_chain(rs, tight_factory_ship=True)
    okw = {k: v for k, v in kw.items() if k not in ("seed", "build_info")}
    cc = CSupplyChainOracle(info, **okw)
    n, T = 96, kw["total_time_steps"]
    dem, lts, a = _inputs(rs, cc, n, T, kw["demand_range"][0], kw["demand_range"][1], wide=(seed % 2 == 1))
    if lts is not None:
        lts = rs.randint(1, kw["max_leadtime"] + 1, size=lts.shape).astype(np.int32)
    ref = _oracle_run(cc, dem, lts, a, canonical=False)
    got = _gpu_run(gsc.SupplyChainEnv(info, num_envs=n, mode="exact_f64", auto_reset=False, **okw), dem, lts, a)
    assert np.array_equal(got[2], ref[2]) and np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])


def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11) on numpy uint64 arrays holding 32-bit words: the generator of
    csrc/scg_common.cuh restated on the host."""
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & MASK for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & MASK, np.uint64(k1) & MASK
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def test_beergame_in_kernel_generators_are_the_documented_philox_streams():
    """beergame-v2 with randint demand / delays (beergame2_env.py:86-92, hi exclusive) generated IN the kernel: ONE Philox
    block keyed (seed, global env, episode, q, stream 4) feeds the four weeks 4q+1 .. 4q+4 of an env -- week w takes word
    (w-1) & 3, demand from its low 16 bits and delay from its high 16 bits, mapped by multiply-shift.  The tables rebuilt on
    the host with that rule drive the C oracle to the same trajectories, over two episodes with auto-reset and a non-zero
    first_env."""
    from oracle.c_oracle import CBeerGameOracle
    n, T, seed, first = 257, 35, 0x1234567890ABCDEF, 1000
    env = gsc.BeerGameEnv2(customer_demand=(0, 16), shipment_delays=(0, 5), num_envs=n, seed=seed, auto_reset=True, first_env=first)
    env.reset()
    acts = np.random.RandomState(1).randint(0, 30, size=(2 * T, n, 4)).astype(np.int32)
    o, r, d = env.rollout(2 * T, torch.from_numpy(acts).cuda())
    o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy().astype(bool)
    genv = np.arange(first, first + n, dtype=np.uint64)
    ora = CBeerGameOracle(version=2, weeks=T)
    for ep in range(2):
        dem, dl = np.empty((n, T), np.int32), np.empty((n, T), np.int32)
        for w in range(1, T + 1):
            c1 = ((genv >> np.uint64(32)) & np.uint64(0xFFFF)) | np.uint64(4 << 28)          # block 0, stream STREAM_BG_DEMAND = 4
            # episode counter: creating the handle resets once (episode 0), env.reset() above started episode 1
            words = _philox4x32_10(genv, c1, np.full(n, ep + 1, np.uint64), np.full(n, (w - 1) >> 2, np.uint64), seed & 0xFFFFFFFF, seed >> 32)
            word = words[(w - 1) & 3]
            dem[:, w - 1] = ((word & np.uint64(0xFFFF)) * np.uint64(16)) >> np.uint64(16)      # lo + (field * (hi - lo)) >> 16
            dl[:, w - 1] = ((word >> np.uint64(16)) * np.uint64(5)) >> np.uint64(16)
        assert dem.min() == 0 and dem.max() == 15 and dl.min() == 0 and dl.max() == 4
        for vals, k in ((dem, 16), (dl, 5)):                   # randint(lo, hi): every value equally likely (5 sigma)
            cnt = np.bincount(vals.ravel(), minlength=k)
            assert np.abs(cnt - vals.size / k).max() < 5 * np.sqrt(vals.size / k)
        for e in range(0, n, 5):
            ora.reset(dem[e], dl[e])
            for t in range(T):
                obs, rew, done, _ = ora.step(acts[ep * T + t, e])
                k = ep * T + t
                assert rew == r[k, e] and done == d[k, e], (ep, e, t)
                if not done:                                   # with auto-reset the obs returned with done is the next episode's first
                    assert np.array_equal(obs, o[k, e])


def test_supplychain_in_kernel_uniform_demand_is_the_documented_philox_stream():
    """ntom, uniform demand U{0..400} generated in the kernel: demand (row t, retailer-product rp) of (env, episode) is
    word rp & 3 of the Philox block keyed (seed, global env, episode, row t, stream 0, block rp >> 2), mapped by
    multiply-shift onto lo .. hi inclusive.  Read back from the first R*P observation entries of a fast_f64 rollout."""
    n, T, seed, first = 193, 40, 77, 5000
    env = gsc.make("supplychain-ntom-v0", num_envs=n, mode="fast_f64", seed=seed, first_env=first, auto_reset=True, total_time_steps=25)
    o0 = env.reset()
    o, r, d = env.rollout(T)
    RP, lo, span = 6, 0, 400
    seen = np.concatenate([o0[None].cpu().numpy(), o.cpu().numpy()])[..., :RP]        # obs after period t shows demand row t
    seen = np.rint((seen + 1) / 2 * span + lo).astype(np.int64)                       # [T+1, n, RP]
    genv = np.arange(first, first + n, dtype=np.uint64)
    done = d.cpu().numpy().astype(bool)
    episode, t = np.full(n, 1, np.uint64), np.zeros(n, np.uint64)                     # create() reset once (episode 0), reset() -> 1
    for k in range(T + 1):
        for b in range(2):
            c1 = ((genv >> np.uint64(32)) & np.uint64(0xFFFF)) | (np.uint64(b) << np.uint64(16))          # stream 0
            words = _philox4x32_10(genv, c1, episode, t, seed, 0)
            for j in range(4):
                rp = 4 * b + j
                if rp < RP:
                    want = (words[j] * np.uint64(span + 1)) >> np.uint64(32)
                    assert np.array_equal(want.astype(np.int64) + lo, seen[k, :, rp]), (k, rp)
        if k < T:
            t = t + np.uint64(1)
            fin = done[k]
            episode = np.where(fin, episode + np.uint64(1), episode)
            t = np.where(fin, np.uint64(0), t)


def test_supplychain_v0_in_kernel_demand_and_leadtimes_drive_the_oracle_to_the_same_trajectory():
    """supplychain-v0 (the reference's demo chain: stochastic lead times avg 2 / max 4, K = 14 values per period) with BOTH
    generators in the kernel.  Host restatement: demand as above (stream 0); lead time k of period t = 1 + #{j < 3 : word >=
    ceil(CDF_Poisson(1)(j) * 2^32)} with word (k & 3) of block (k >> 2), stream 1, row t (clip(1 + Poisson(avg-1), 1, max),
    supplychain_env.py:670-672).  The C oracle fed those tables reproduces every reward of the fp64 ring kernel."""
    import math
    n, seed, first = 130, 2024, 300
    info, kw = presets.main_demo_chain()
    okw = {k: v for k, v in kw.items() if k not in ("seed", "build_info")}
    T = okw["total_time_steps"]
    cc = CSupplyChainOracle(info, **okw)
    env = gsc.make("supplychain-v0", num_envs=n, mode="fast_f64", seed=seed, first_env=first, auto_reset=False)
    K, RP = cc.n_leadtimes, cc.n_retailers * cc.P
    lo, hi = okw["demand_range"]
    assert K == 14 and env.scg_info.n_leadtimes == K
    genv = np.arange(first, first + n, dtype=np.uint64)
    ep = np.full(n, 1, np.uint64)                                   # create() -> episode 0, reset() below -> 1

    def words(stream, row, count):
        out = np.empty((n, count), np.uint64)
        for b in range((count + 3) // 4):
            c1 = ((genv >> np.uint64(32)) & np.uint64(0xFFFF)) | (np.uint64(b) << np.uint64(16)) | (np.uint64(stream) << np.uint64(28))
            w4 = _philox4x32_10(genv, c1, ep, np.full(n, row, np.uint64), seed, 0)
            for j in range(4):
                if 4 * b + j < count:
                    out[:, 4 * b + j] = w4[j]
        return out
    dem = np.stack([(words(0, t, RP) * np.uint64(hi - lo + 1)) >> np.uint64(32) for t in range(T + 1)], axis=1).astype(np.int32) + lo
    cdf, p, thr = 0.0, math.exp(-1.0), []
    for j in range(3):
        cdf += p; thr.append(math.ceil(cdf * 4294967296.0)); p *= 1.0 / (j + 1)
    lts = np.stack([1 + sum((words(1, t, K) >= np.uint64(th)).astype(np.int32) for th in thr) for t in range(1, T + 1)], axis=1).astype(np.int32)
    assert lts.min() == 1 and lts.max() == 4 and abs(lts.mean() - 2.0) < 0.05
    a = np.random.RandomState(3).uniform(-1, 1, size=(T, n, cc.n_actions)).astype(np.float32)
    ref = _oracle_run(cc, dem.reshape(n, T + 1, cc.n_retailers, cc.P), lts, a, canonical=True)
    o0 = env.reset()
    o, r, d = env.rollout(T, torch.from_numpy(a).double().cuda())
    assert np.allclose(r.cpu().numpy(), ref[1], rtol=1e-11, atol=1e-9)
    assert np.abs(np.concatenate([o0[None].cpu().numpy(), o.cpu().numpy()]) - ref[0]).max() <= 1e-11
    assert np.array_equal(d.cpu().numpy().astype(bool), ref[2])
