"""Generate tests/golden/*.npz from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference, which does not travel to the
GPU box):

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

The reference is imported through oracle/gym_shim (the image has no `gym`).  Each case
drives one reference env with float64 actions whose values are exactly representable in
float32 (every reference op is then IEEE double under any numpy), and
records, per step: the action, the reference observation, reward, done, every node's
stock and the raw heap arrays of in-transit material; per episode: the demand and
lead-time tables the reference drew (which the tests inject into the oracle / kernels).
The constructor arguments the reference class passed to SupplyChainEnv.__init__
(`nodes_info` + kwargs) are captured into the file as JSON, which pins the topology
presets of the product package.
"""
import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SCG_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "gym_shim"))
sys.path.insert(0, REF)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _fp32_exact(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def make_actions(rs, A, style):
    """Random policy with the corner cases the SHIP scaling cares about (ties, +-1, zeros)."""
    a = rs.uniform(-1.0, 1.0, A)
    if style == "wide":           # out-of-range actions: the reference does not clip (F: :697-698)
        a = rs.uniform(-1.3, 1.3, A)
    u = rs.uniform(size=A)
    a[u < 0.06] = -1.0
    a[(u >= 0.06) & (u < 0.12)] = 1.0
    a[(u >= 0.12) & (u < 0.15)] = 0.0
    for i in range(1, A):         # ties between neighbouring entries
        if rs.uniform() < 0.08:
            a[i] = a[i - 1]
    if style == "sendall":
        a[rs.uniform(size=A) < 0.5] = 1.0
    return _fp32_exact(a)


def run_supplychain_case(name, cls_name, kwargs, episodes=1, seed=0, style="mixed", heap_cap=None):
    import gym_supplychain.envs as E
    from gym_supplychain.envs.supplychain_env import SupplyChainEnv

    captured = {}
    orig_init = SupplyChainEnv.__init__

    def spy(self, nodes_info, **kw):
        captured["nodes_info"] = copy.deepcopy(nodes_info)
        captured["kwargs"] = copy.deepcopy(kw)
        orig_init(self, nodes_info, **kw)

    SupplyChainEnv.__init__ = spy
    try:
        if cls_name == "SupplyChainEnv":
            env = SupplyChainEnv(kwargs["nodes_info"], build_info=True, **{k: v for k, v in kwargs.items() if k != "nodes_info"})
        else:
            env = getattr(E, cls_name)(build_info=True, **kwargs)     # info['sc_episode'] is recorded; dynamics unchanged
    finally:
        SupplyChainEnv.__init__ = orig_init

    env.seed(seed)
    N, P, T = len(env.nodes), env.num_products, env.total_time_steps
    A, O = env.action_space.shape[0], env.observation_space.shape[0]
    R = len(env.last_level_nodes)
    rs = np.random.RandomState(10007 + seed)
    if heap_cap is None:
        heap_cap = 4
        for nd in env.nodes:
            n_src = sum(1 for o in env.nodes if o.dests and nd in o.dests) + (1 if nd.num_supply_actions else 0)
            n_init = max((len(x) for x in (nd.initial_supply or [[]])), default=0) + \
                max((len(x) for x in (nd.initial_shipments or [[]])), default=0)
            heap_cap = max(heap_cap, n_src * env.max_leadtime + n_init)

    rec = {k: [] for k in ("demand", "leadtimes", "actions", "obs", "reward", "done", "stock",
                           "heap_len", "heap_time", "heap_amount", "info_costs", "info_units", "info_rewards")}
    KEYS = ("stock", "stock_pen", "supply", "process", "process_pen", "ship", "ship_pen", "unmet_dem")

    def snapshot():
        st = np.zeros((N, P))
        hl = np.zeros((N, P), dtype=np.int32)
        ht = np.zeros((N, P, heap_cap), dtype=np.int32)
        ha = np.zeros((N, P, heap_cap))
        for n, nd in enumerate(env.nodes):
            for p in range(P):
                st[n, p] = nd.stock[p]
                q = nd.shipments_by_prod[p]
                assert len(q) <= heap_cap, (len(q), heap_cap)
                hl[n, p] = len(q)
                for k, (w, amt) in enumerate(q):
                    ht[n, p, k] = w
                    ha[n, p, k] = amt
        return st, hl, ht, ha

    for ep in range(episodes):
        obs0 = env.reset()
        if env.demand_config_by_product:
            dem = np.stack([np.asarray(env.customer_demands[p]) for p in range(P)], axis=-1)
        else:
            dem = np.asarray(env.customer_demands)
        assert dem.shape == (T + 1, R, P), dem.shape
        lts = np.asarray(env.leadtimes) if env.stochastic_leadtimes else np.zeros((T, 0), dtype=np.int64)
        e = {k: [] for k in rec}
        e["demand"], e["leadtimes"] = dem.astype(np.int32), lts.astype(np.int32)
        st, hl, ht, ha = snapshot()
        e["obs"].append(np.asarray(obs0, dtype=np.float64))
        e["stock"].append(st); e["heap_len"].append(hl); e["heap_time"].append(ht); e["heap_amount"].append(ha)
        done = False
        while not done:
            a = make_actions(rs, A, style)
            obs, r, done, info = env.step(a)
            ep = info["sc_episode"]
            e["info_costs"].append(np.array([[float(ep["costs"][k][p]) for p in range(P)] for k in KEYS]))
            e["info_units"].append(np.array([[float(ep["units"][k][p]) for p in range(P)] for k in KEYS]))
            e["info_rewards"].append(float(ep["rewards"]))
            st, hl, ht, ha = snapshot()
            e["actions"].append(a)
            e["obs"].append(np.asarray(obs, dtype=np.float64))
            e["reward"].append(float(r)); e["done"].append(bool(done))
            e["stock"].append(st); e["heap_len"].append(hl); e["heap_time"].append(ht); e["heap_amount"].append(ha)
        for k in rec:
            rec[k].append(np.asarray(e[k]))

    captured["kwargs"].pop("build_info", None)
    meta = {"name": name, "cls": cls_name, "ctor_kwargs": _jsonable(kwargs), "nodes_info": captured["nodes_info"],
            "env_kwargs": _jsonable(captured["kwargs"]), "N": N, "P": P, "T": T, "A": A, "O": O, "R": R,
            "K": int(getattr(env, "count_leadtimes_per_timestep", 0)), "heap_cap": int(heap_cap),
            "episodes": episodes, "seed": seed, "numpy": np.__version__}
    out = {k: np.stack(v) for k, v in rec.items()}
    out["meta"] = np.array(json.dumps(meta))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "sc_%s.npz" % name), **out)
    print("sc_%-28s N=%2d P=%d A=%3d O=%3d T=%3d return[0]=%.6f" % (name, N, P, A, O, T, out["reward"][0].sum()))


def _jsonable(x):
    if isinstance(x, dict):
        return {k: _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    return x


def main_chain_nodes_info():
    """The 2-2-2-2 chain of the reference's __main__ demo (supplychain_env.py:816-868), re-typed."""
    stock_cost, cap, ship = 1, 300, 300
    dc = [[2 * stock_cost] * 2]
    info = {}
    for i, s0 in ((1, 10), (2, 0)):
        info["Supplier %d" % i] = {"initial_stock": s0, "stock_capacity": cap, "stock_cost": stock_cost,
                                   "supply_capacity": 50, "supply_cost": 5, "destinations": ["Factory  1", "Factory  2"],
                                   "dest_costs": dc, "ship_capacity": [ship] * 2}
    for i in (1, 2):
        info["Factory  %d" % i] = {"initial_stock": 0, "stock_capacity": cap, "stock_cost": stock_cost,
                                   "processing_capacity": 50, "processing_cost": 10,
                                   "destinations": ["Wholesal 1", "Wholesal 2"], "dest_costs": dc, "ship_capacity": [ship] * 2}
    for i, s0 in ((1, 10), (2, 15)):
        info["Wholesal %d" % i] = {"initial_stock": s0, "stock_capacity": cap, "stock_cost": stock_cost,
                                   "destinations": ["Retailer 1", "Retailer 2"], "dest_costs": dc, "ship_capacity": [ship] * 2}
    for i, s0 in ((1, 10), (2, 20)):
        info["Retailer %d" % i] = {"initial_stock": s0, "stock_capacity": cap, "stock_cost": stock_cost, "last_level": True}
    kw = dict(num_products=1, demand_range=(10, 20), unmet_demand_cost=54, exceeded_stock_capacity_cost=10,
              exceeded_process_capacity_cost=10, exceeded_ship_capacity_cost=10, processing_ratio=3,
              stochastic_leadtimes=True, avg_leadtime=2, max_leadtime=4, total_time_steps=120)
    return info, kw


def simple_chain_nodes_info():
    """A heterogeneous 1-1-1-1 chain with a tight factory and tight links (exercises every penalty)."""
    info = {
        "S": {"initial_stock": 10, "stock_capacity": 120, "stock_cost": 1, "supply_capacity": 90, "supply_cost": 5,
              "initial_supply": [[7, 0, 3]], "destinations": ["F"], "dest_costs": [[2]], "ship_capacity": [40]},
        "F": {"initial_stock": 30, "stock_capacity": 100, "stock_cost": 2, "processing_capacity": 25,
              "processing_cost": 10, "initial_shipments": [[5, 5]], "destinations": ["W"], "dest_costs": [[3]],
              "ship_capacity": [9]},
        "W": {"initial_stock": 12, "stock_capacity": 30, "stock_cost": 1, "initial_shipments": [[25, 25]],
              "destinations": ["R"], "dest_costs": [[1]], "ship_capacity": [15]},
        "R": {"initial_stock": 5, "stock_capacity": 20, "stock_cost": 3, "initial_shipments": [[18, 2]], "last_level": True},
    }
    kw = dict(num_products=1, demand_range=(0, 12), unmet_demand_cost=50, exceeded_stock_capacity_cost=7,
              exceeded_process_capacity_cost=11, exceeded_ship_capacity_cost=13, processing_ratio=2,
              stochastic_leadtimes=True, avg_leadtime=3, max_leadtime=5, total_time_steps=80)
    return info, kw


def run_beergame_case(name, version, kwargs, episodes, seed, action_lo, action_hi, inject=None):
    from gym_supplychain.envs import BeerGameEnv, BeerGameEnv2
    env = BeerGameEnv(kwargs) if version == 0 else BeerGameEnv2(**kwargs)
    if version == 2 and hasattr(env, "rand_generator"):
        env.seed(seed)
    rs = np.random.RandomState(777 + seed)
    rec = {k: [] for k in ("demand", "delays", "actions", "obs", "reward", "done", "inventory", "backlog")}
    for ep in range(episodes):
        e = {k: [] for k in rec}
        obs = env.reset()
        T = env.max_weeks
        e["demand"] = np.asarray(env.customer_demand[:T], dtype=np.int32)
        e["delays"] = np.asarray(env.shipment_delays[1:T + 1], dtype=np.int32)
        e["obs"].append(np.asarray(obs, dtype=np.int64))
        e["inventory"].append(env.inventory.copy()); e["backlog"].append(env.backlog.copy())
        done = False
        while not done:
            a = rs.randint(action_lo, action_hi + 1, env.levels)
            obs, r, done, _ = env.step(a)
            e["actions"].append(a.astype(np.int32))
            e["obs"].append(np.asarray(obs, dtype=np.int64))
            e["reward"].append(int(r)); e["done"].append(bool(done))
            e["inventory"].append(env.inventory.copy()); e["backlog"].append(env.backlog.copy())
        for k in rec:
            rec[k].append(np.asarray(e[k]))
    out = {k: np.stack(v) for k, v in rec.items()}
    meta = {"name": name, "version": version, "kwargs": _jsonable(kwargs), "episodes": episodes, "seed": seed,
            "T": int(env.max_weeks), "L": int(env.levels)}
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "beergame_%s.npz" % name), **out)
    print("beergame_%-22s T=%d return[0]=%d" % (name, env.max_weeks, out["reward"][0].sum()))


def main():
    # ---- SupplyChainEnv family ----
    run_supplychain_case("linear", "SupplyChainNPerStage", dict(nodes_per_echelon=1, num_products=1), episodes=2)
    run_supplychain_case("linear_stoch", "SupplyChainNPerStage",
                         dict(nodes_per_echelon=1, num_products=1, stochastic_leadtimes=True, avg_leadtime=2,
                              max_leadtime=4, total_time_steps=120), seed=3)
    info, kw = simple_chain_nodes_info()
    run_supplychain_case("simple_tight", "SupplyChainEnv", dict(nodes_info=info, **kw), seed=5, style="sendall")
    run_supplychain_case("2perstage", "SupplyChain2perStageEnv", dict(), episodes=2)
    run_supplychain_case("2perstage_stoch", "SupplyChain2perStageEnv",
                         dict(stochastic_leadtimes=True, avg_leadtime=2, max_leadtime=4, total_time_steps=180), seed=1)
    run_supplychain_case("2perstage_seasonal", "SupplyChain2perStageSeasonalEnv", dict(total_time_steps=180), seed=2)
    run_supplychain_case("2perstage_wide", "SupplyChain2perStageEnv", dict(total_time_steps=60), seed=4, style="wide")
    run_supplychain_case("ntom", "SupplyChainNPerStage", dict(nodes_per_echelon=3), episodes=1)
    run_supplychain_case("ntom_sendall", "SupplyChainNPerStage", dict(nodes_per_echelon=3, total_time_steps=90),
                         seed=7, style="sendall")
    run_supplychain_case("ntom_p3_stoch", "SupplyChainNPerStage",
                         dict(nodes_per_echelon=3, num_products=3, stochastic_leadtimes=True, avg_leadtime=3,
                              max_leadtime=5, total_time_steps=60), seed=6)
    run_supplychain_case("chain_5_4_7_10", "SupplyChainNPerStage",
                         dict(nodes_per_echelon=[5, 4, 7, 10], total_time_steps=40), seed=8)
    run_supplychain_case("chain_3_2_3_5_p4", "SupplyChainNPerStage",
                         dict(nodes_per_echelon=[3, 2, 3, 5], num_products=4, total_time_steps=40), seed=9)
    run_supplychain_case("oneonen", "SupplyChainNPerStage",
                         dict(nodes_per_echelon=[1, 1, 1, 3], num_products=1, total_time_steps=120), seed=10)
    run_supplychain_case("multiproduct", "SupplyChainMultiProduct", dict(total_time_steps=100), seed=11)
    run_supplychain_case("multiproduct_inccosts", "SupplyChainMultiProduct_IncreasingCosts",
                         dict(num_products=3, total_time_steps=60), seed=12)
    run_supplychain_case("multiproduct_bydem", "SupplyChainMultiProduct_DemConfigByProd",
                         dict(num_products=3, demand_std=20, demand_perturb_norm=True, stochastic_leadtimes=True,
                              avg_leadtime=2, max_leadtime=3, total_time_steps=60), seed=13)
    info, kw = main_chain_nodes_info()
    run_supplychain_case("supplychain_v0_main", "SupplyChainEnv", dict(nodes_info=info, **kw), seed=14)

    # ---- beer game ----
    run_beergame_case("v0_default", 0, {}, 2, 0, -3, 5)
    rs = np.random.RandomState(99)
    run_beergame_case("v0_delays", 0, {"customer_demand": [int(x) for x in rs.randint(0, 12, 50)],
                                        "shipment_delays": [int(x) for x in rs.randint(0, 4, 50)],
                                        "initial_inventory": [12, 8, 20, 3], "inv_cost": 2, "backlog_cost": 5,
                                        "initial_shipment_value": 3, "initial_orders_value": 6}, 1, 1, -3, 5)
    run_beergame_case("v2_default", 2, {}, 2, 2, 0, 29)
    run_beergame_case("v2_stochastic", 2, {"customer_demand": (0, 16), "shipment_delays": (0, 5)}, 3, 3, 0, 29)
    run_beergame_case("v2_small_cap", 2, {"max_stock": 20, "exceeded_capacity_penalty": 7, "shipment_delays": 1,
                                           "initial_inventory": [5, 30, 2, 9]}, 1, 4, 0, 29)


if __name__ == "__main__":
    main()
