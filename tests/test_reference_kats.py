"""Known-answer tests transcribed from the reference's own test-suite (expected numbers only), run against the
CPU oracle (CPU suite) and against the CUDA exact_f64 path through the C ABI (GPU suite).

  KAT-A  gym_supplychain/envs/tests/test_supplychain_2perstage_env.py:29-170  (test_chain_dynamics)
  KAT-B  gym_supplychain/envs/tests/test_supplychain_env.py:60-127            (test_simpleenv)

The demand tables are the seed-0 RandomState draws the reference tests pin (:41 and :67); they are injected.
"""
import numpy as np
import pytest

from gym_supplychain_b200 import presets
from oracle.sc_oracle import SupplyChainOracle

# ---------------- KAT-A ----------------
A_DEMAND = np.array([15, 10, 13, 13, 17, 19, 13, 15, 12, 14, 17, 16]).reshape(6, 2, 1)
A_OBS0 = [0., -1., -1., 0., 0., -1., -0.2, -0.2, -1., -0.76, -0.76, -1., -0.76, -0.76, -1., -0.92, -0.92, -1.,
          -0.92, -0.92, -1., -0.92, -0.92, -1., -0.92, -0.92, 1.]
A_SUPPLY = 2 * np.array([1] + [0] * 13, dtype=np.float64) - 1
A_SEND_ALL = 2 * np.array([0, 1, 1] * 2 + [1] * 2 * 4, dtype=np.float64) - 1
A_SEND_HALF = 2 * np.array([0, 0.5, 1] * 2 + [0.5, 1] * 4, dtype=np.float64) - 1
A_STEPS = [
    (A_SUPPLY, -1015.0,
     [-0.4, -0.4, -0.4, 0., 1., -0.6, -0.2, -1., -0.4, -0.76, -1., -0.6, -0.76, -1., -0.8, -0.92, -1., -0.86666667, -0.92,
      -1., -0.95, -0.92, -1., -0.93333333, -0.92, -1., 0.6]),
    (A_SEND_ALL, -3469.0,
     [0.4, 0.8, -1., 1., -1., -1., -1., -1., -1., -1., -0.04, -1., -1., -1., -1., -1., -0.68, -1., -1., -1., -0.88, -1.,
      -0.68, -0.88666667, -1., -1., 0.2]),
    (A_SEND_HALF, -1752.0,
     [-0.4, 0., -1., -1., -1., -1., -1., -1., -1., -0.04, -0.76, -1., -1., -0.76, -1., -0.68, -1., -1., -1., -1., -1.,
      -0.68, -1., -1., -1., -1., -0.2]),
    (A_SEND_HALF, -6400.333,
     [-0.6, -0.2, -1., -1., -1., -1., -1., -1., -1., -0.76, -1., -1., -0.76, -1., -1., -1., -0.86666667, -1., -1.,
      -0.86666667, -0.33, -1., -0.84, -1., -1., -0.84, -0.6]),
    (A_SEND_HALF, -4479.0,
     [0.4, 0.2, -1., -1., -1., -1., -1., -1., -1., -1., -1., -1., -1., -1., -1., -0.86666667, -0.92, -1., -0.86666667,
      -0.92, -0.45, -0.84, -1., -1., -0.84, -1., -1.]),
]
A_FINAL_STOCK_RETAILERS = (55.0, 0.0)


def _kat_a_spec():
    return presets.two_per_stage(total_time_steps=5, ship_capacity=250)


def _check_kat_a(reset, step, final_stocks):
    obs = reset()
    assert np.allclose(obs, A_OBS0)
    done = False
    for act, rew, want in A_STEPS:
        obs, r, done = step(act)
        assert np.allclose(obs, want), np.abs(np.asarray(obs) - want).max()
        assert np.round(r, 3) == rew
    assert done
    st = final_stocks()
    assert np.allclose(st[:6].sum(), 0) and (st[6].sum(), st[7].sum()) == A_FINAL_STOCK_RETAILERS


def test_kat_a_oracle():
    info, kw = _kat_a_spec()
    env = SupplyChainOracle(info, **kw)
    _check_kat_a(lambda: env.reset(demand=A_DEMAND), lambda a: env.step(a)[:3], env.stocks)


# ---------------- KAT-B ----------------
B_DEMAND = np.array([4, 5, 0, 3, 3, 3]).reshape(6, 1, 1)


def _kat_b_spec():
    def node(**extra):
        d = {"initial_stock": 10, "stock_capacity": 100, "stock_cost": 1}
        d.update(extra)
        return d
    ship = dict(dest_costs=[[2] * 2], ship_capacity=[100] * 2)          # the reference passes 2-long lists for 1 dest
    info = {"Supplier": node(supply_capacity=50, supply_cost=5, destinations=["Factory"], **ship),
            "Factory": node(processing_capacity=100, processing_cost=10, destinations=["Wholesal"], **ship),
            "Wholesal": node(destinations=["Retailer"], **ship),
            "Retailer": node(last_level=True)}
    kw = dict(num_products=1, unmet_demand_cost=1000, exceeded_stock_capacity_cost=1000,
              exceeded_process_capacity_cost=1000, exceeded_ship_capacity_cost=1000, demand_range=(0, 5),
              processing_ratio=2, stochastic_leadtimes=False, avg_leadtime=2, max_leadtime=2, total_time_steps=5)
    return info, kw


B_SUPPLY = 2 * np.array([1, 0, 0, 0, 0, 0], dtype=np.float64) - 1     # 6 values for a 4-action env: extras ignored
B_SEND_ALL = np.ones(6)
# heaps[(node)] after each step (arrival period, amount), and stocks
B_EXPECT = [
    (B_SUPPLY, [[(3, 50)], [], [], []], [10, 10, 10, 10 - 4]),
    (B_SEND_ALL, [[(3, 50), (4, 50)], [(4, 10)], [(4, 5)], [(4, 10)]], [0, 0, 0, max(0, 10 - 9)]),
    (B_SEND_ALL, [[(4, 50), (5, 50)], [(4, 10), (5, 50)], [(4, 5)], [(4, 10)]], [0, 0, 0, max(0, 10 - 9)]),
    (B_SEND_ALL, [[(5, 50), (6, 50)], [(5, 50), (6, 50)], [(6, 5)], [(6, 5)]], [0, 0, 0, max(0, 20 - 12)]),
    (B_SEND_ALL, [[(6, 50), (7, 50)], [(6, 50), (7, 50)], [(6, 5), (7, 25)], [(6, 5)]], [0, 0, 0, max(0, 20 - 15)]),
]


def test_kat_b_oracle():
    info, kw = _kat_b_spec()
    env = SupplyChainOracle(info, **kw)
    assert env.n_actions == 4
    env.reset(demand=B_DEMAND)
    assert all(nd.q == [[]] for nd in env.nodes)
    for act, heaps, stocks in B_EXPECT:
        env.step(act[:4])
        for n, nd in enumerate(env.nodes):
            assert sorted(nd.q[0]) == heaps[n], (n, nd.q[0])
            assert nd.stock[0] == stocks[n]


# ---------------- the same KATs on the GPU (exact_f64 through the C ABI) ----------------
@pytest.mark.gpu
def test_kat_a_gpu_exact_and_mixed():
    import gym_supplychain_b200 as gsc
    for mode, single in (("exact_f64", True), ("mixed", False), ("fast_f32", False)):
        env = gsc.SupplyChain2perStageEnv(total_time_steps=5, ship_capacity=250, mode=mode,
                                          num_envs=None if single else 4, auto_reset=False)
        env.inject(demand=A_DEMAND)
        if single:
            _check_kat_a(env.reset, lambda a: env.step(a)[:3],
                         lambda: env.state_dict()["stock"][:, :, 0].cpu().numpy())
        else:
            import torch
            _check_kat_a(lambda: env.reset()[3].double().cpu().numpy(),
                         lambda a: (lambda o, r, d, _: (o[3].double().cpu().numpy(), float(r[3]), bool(d[3])))(
                             *env.step(torch.from_numpy(a).expand(4, -1))),
                         lambda: env.state_dict()["stock"][:, :, 3].double().cpu().numpy().round(3))


@pytest.mark.gpu
def test_kat_b_gpu_exact():
    import gym_supplychain_b200 as gsc
    info, kw = _kat_b_spec()
    env = gsc.SupplyChainEnv(info, **kw)                       # single env, exact_f64
    assert env.action_space.shape == (4,) and env.observation_space.shape == (14,)
    env.inject(demand=B_DEMAND)
    env.reset()
    for act, heaps, stocks in B_EXPECT:
        env.step(act)                                          # 6 values: trailing extras are ignored like in the reference
        st = env.state_dict()
        for n in range(4):
            ln = int(st["heap_len"][n, 0, 0])
            got = sorted((int(st["heap_time"][n, 0, k, 0]), float(st["heap_amount"][n, 0, k, 0])) for k in range(ln))
            assert got == [(t, float(a)) for t, a in heaps[n]], (n, got)
            assert float(st["stock"][n, 0, 0]) == stocks[n]
