"""The plain-C oracle (oracle/sc_oracle.c) against (a) every golden vector recorded from the unmodified reference, bit for
bit -- observations, rewards, done flags, stocks and the raw heapq arrays -- and (b) the Python oracle on random chains."""
import numpy as np
import pytest

from golden_util import beergame_cases, load, sc_cases
from oracle.c_oracle import CBeerGameOracle, CSupplyChainOracle
from oracle.sc_oracle import SupplyChainOracle


@pytest.mark.parametrize("name", sc_cases())
def test_c_oracle_matches_reference_golden(name):
    meta, z = load("sc", name)
    kw = {k: v for k, v in meta["env_kwargs"].items() if k != "build_info"}
    env = CSupplyChainOracle(meta["nodes_info"], **kw)
    assert env.n_actions == meta["A"] and env.n_obs == meta["O"]
    assert (env.n_leadtimes if env.stoch_lt else 0) == meta["K"]
    cap = meta["heap_cap"]
    for ep in range(meta["episodes"]):
        lts = z["leadtimes"][ep] if env.stoch_lt else None
        obs = env.reset(z["demand"][ep], lts)
        assert np.array_equal(obs, z["obs"][ep, 0])
        for t in range(meta["T"]):
            obs, r, done, _ = env.step(z["actions"][ep, t])
            assert r == z["reward"][ep, t], (name, ep, t)
            assert done == bool(z["done"][ep, t])
            assert np.array_equal(env.stocks(), z["stock"][ep, t + 1]), (name, ep, t)
            ln, tm, am = env.heap_arrays(cap)
            assert np.array_equal(ln, z["heap_len"][ep, t + 1])
            assert np.array_equal(tm, z["heap_time"][ep, t + 1])
            assert np.array_equal(am, z["heap_amount"][ep, t + 1])
            assert np.array_equal(obs, z["obs"][ep, t + 1]), (name, ep, t)
        assert done
        # the one-call rollout gives the same trajectory
        env.reset(z["demand"][ep], lts)
        o2, r2, d2 = env.rollout(z["actions"][ep])
        assert np.array_equal(o2, z["obs"][ep, 1:]) and np.array_equal(r2, z["reward"][ep]) and np.array_equal(d2.astype(bool), z["done"][ep])


def _random_chain(rs):
    """A random DAG that is NOT layered: random fan-out into later nodes (every node reachable), per-product constants,
    one processing node, constant or stochastic lead times."""
    P = int(rs.randint(1, 4)); n_mid = int(rs.randint(2, 5)); n_ret = int(rs.randint(1, 4))
    names = ["S0", "S1"] + ["M%d" % i for i in range(n_mid)] + ["R%d" % i for i in range(n_ret)]
    dests = {}
    for i, nm in enumerate(names):
        if not nm.startswith("R"):
            later = [m for m in names[i + 1:] if not m.startswith("S")]
            k = int(rs.randint(1, min(3, len(later)) + 1))
            dests[nm] = [str(x) for x in rs.choice(later, size=k, replace=False)]
    for i, nm in enumerate(names):                                   # every non-supplier node needs a source
        if not nm.startswith("S") and not any(nm in d for d in dests.values()):
            src = [m for m in names[:i] if m in dests]
            dests[str(rs.choice(src))].append(nm)
    info = {}
    for nm in names:
        d = {"initial_stock": [int(x) for x in rs.randint(0, 300, P)], "stock_capacity": [int(x) for x in rs.randint(200, 900, P)],
             "stock_cost": [int(x) for x in rs.randint(1, 4, P)]}
        if nm.startswith("S"):
            d.update(supply_capacity=[int(x) for x in rs.randint(50, 400, P)], supply_cost=[int(x) for x in rs.randint(1, 9, P)],
                     initial_supply=[[int(x) for x in rs.randint(0, 200, 2)] for _ in range(P)])
        else:
            d["initial_shipments"] = [[int(x) for x in rs.randint(0, 200, 2)] for _ in range(P)]
        if nm.startswith("R"):
            d["last_level"] = True
        else:
            k = len(dests[nm])
            d.update(destinations=dests[nm], dest_costs=[[int(x) for x in rs.randint(1, 5, k)] for _ in range(P)],
                     ship_capacity=[int(x) for x in rs.randint(60, 500, k)])
        if nm == "M0":
            d.update(processing_capacity=int(rs.randint(100, 600)), processing_cost=[int(x) for x in rs.randint(1, 6, P)])
        info[nm] = d
    stoch = bool(rs.randint(0, 2))
    kw = dict(num_products=P, demand_range=(0, int(rs.randint(20, 300))), stochastic_leadtimes=stoch, avg_leadtime=2,
              max_leadtime=4 if stoch else 2, total_time_steps=40, unmet_demand_cost=int(rs.randint(10, 300)),
              exceeded_stock_capacity_cost=7, exceeded_process_capacity_cost=11, exceeded_ship_capacity_cost=13)
    return info, kw


@pytest.mark.parametrize("seed", range(12))
def test_c_oracle_equals_python_oracle_on_random_dags(seed):
    rs = np.random.RandomState(1000 + seed)
    info, kw = _random_chain(rs)
    py, cc = SupplyChainOracle(info, seed=seed, **kw), CSupplyChainOracle(info, **kw)
    for ep in range(2):
        demand, lts = py.draw_inputs()
        o1, o2 = py.reset(demand=demand, leadtimes=lts), cc.reset(demand, lts)
        assert np.array_equal(o1, o2)
        for t in range(kw["total_time_steps"]):
            a = rs.uniform(-1.25, 1.25, py.n_actions).astype(np.float32).astype(np.float64)
            a[rs.uniform(size=a.size) < 0.1] = 1.0
            o1, r1, d1, _ = py.step(a)
            o2, r2, d2, _ = cc.step(a)
            assert r1 == r2 and d1 == d2, (seed, ep, t)
            assert np.array_equal(o1, o2) and np.array_equal(py.stocks(), cc.stocks())
            assert np.array_equal(py.canonical_observation(), cc.observation(canonical=True))


@pytest.mark.parametrize("name", beergame_cases())
def test_c_beergame_oracle_matches_reference_golden(name):
    meta, z = load("beergame", name)
    kw = dict(meta["kwargs"])
    if meta["version"] == 0:
        env = CBeerGameOracle(version=0, levels=kw.get("levels", 4), weeks=meta["T"], initial_inventory=kw.get("initial_inventory"),
                              initial_shipment=kw.get("initial_shipment_value", 4), initial_orders=kw.get("initial_orders_value", 4),
                              inv_cost=kw.get("inv_cost", 1), backlog_cost=kw.get("backlog_cost", 2))
    else:
        env = CBeerGameOracle(version=2, levels=kw.get("levels", 4), weeks=meta["T"], initial_inventory=kw.get("initial_inventory"),
                              **{k: kw[k] for k in ("initial_shipment", "initial_orders", "inv_cost", "backlog_cost", "max_stock",
                                                    "exceeded_capacity_penalty") if k in kw})
    for ep in range(meta["episodes"]):
        obs = env.reset(z["demand"][ep], z["delays"][ep])
        assert np.array_equal(obs, z["obs"][ep, 0])
        for t in range(meta["T"]):
            obs, r, done, _ = env.step(z["actions"][ep, t])
            assert np.array_equal(obs, z["obs"][ep, t + 1]), (name, ep, t)
            assert r == z["reward"][ep, t] and done == bool(z["done"][ep, t])


@pytest.mark.parametrize("seed", range(10))
def test_c_beergame_oracle_equals_python_oracle_on_random_configs(seed):
    from oracle.beergame_oracle import BeerGameOracle
    rs = np.random.RandomState(7000 + seed)
    version = 0 if seed % 2 == 0 else 2
    L, T = int(rs.randint(2, 7)), int(rs.randint(8, 45))
    inv0 = [int(x) for x in rs.randint(0, 20, L)]
    ship, orders, ic, bc = int(rs.randint(0, 8)), int(rs.randint(0, 8)), int(rs.randint(1, 4)), int(rs.randint(1, 5))
    ms, pen = int(rs.randint(20, 120)), int(rs.randint(1, 200))
    py = BeerGameOracle(version=version, levels=L, customer_demand=[4] * T, initial_inventory=inv0, shipment_delays=[2] * T,
                        initial_shipment=ship, initial_orders=orders, inv_cost=ic, backlog_cost=bc, max_stock=ms,
                        exceeded_capacity_penalty=pen, weeks=T)
    cc = CBeerGameOracle(version=version, levels=L, weeks=T, initial_inventory=inv0, initial_shipment=ship, initial_orders=orders,
                         inv_cost=ic, backlog_cost=bc, max_stock=ms, exceeded_capacity_penalty=pen)
    for ep in range(3):
        dem, dl = rs.randint(0, 17, T), rs.randint(0, 5, T)
        assert np.array_equal(py.reset(demand=dem, delays=dl), cc.reset(dem, dl))
        for t in range(T):
            a = rs.randint(-3, 6, L) if version == 0 else rs.randint(0, 30, L)
            o1, r1, d1, _ = py.step(a)
            o2, r2, d2, _ = cc.step(a)
            assert np.array_equal(o1, o2) and r1 == r2 and d1 == d2, (seed, ep, t)
