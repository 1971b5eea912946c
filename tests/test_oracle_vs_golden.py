"""The CPU oracle must reproduce the golden vectors recorded from the unmodified reference, bit for bit."""
import numpy as np
import pytest

from golden_util import beergame_cases, load, sc_cases
from oracle.beergame_oracle import BeerGameOracle
from oracle.sc_oracle import INFO_KEYS, SupplyChainOracle


@pytest.mark.parametrize("name", sc_cases())
def test_supplychain_oracle_matches_reference_golden(name):
    meta, z = load("sc", name)
    env = SupplyChainOracle(meta["nodes_info"], build_info=True, **meta["env_kwargs"])
    assert env.n_actions == meta["A"] and env.n_obs == meta["O"]
    assert (env.n_leadtimes if env.stoch_lt else 0) == meta["K"]
    cap = meta["heap_cap"]
    for ep in range(meta["episodes"]):
        lts = z["leadtimes"][ep] if env.stoch_lt else None
        obs = env.reset(demand=z["demand"][ep], leadtimes=lts)
        assert np.array_equal(obs, z["obs"][ep, 0])
        for t in range(meta["T"]):
            obs, r, done, info = env.step(z["actions"][ep, t])
            assert r == z["reward"][ep, t], (name, ep, t)
            assert done == bool(z["done"][ep, t])
            epi = info["sc_episode"]                   # build_info statistics (supplychain_env.py:750-760, :793-796)
            assert epi["rewards"] == z["info_rewards"][ep, t]
            for ki, key in enumerate(INFO_KEYS):
                assert [float(x) for x in epi["costs"][key]] == list(z["info_costs"][ep, t, ki]), (name, t, key)
                assert [float(x) for x in epi["units"][key]] == list(z["info_units"][ep, t, ki]), (name, t, key)
            assert np.array_equal(env.stocks(), z["stock"][ep, t + 1]), (name, ep, t)
            ln, tm, am = env.heap_arrays(cap)
            assert np.array_equal(ln, z["heap_len"][ep, t + 1])
            assert np.array_equal(tm, z["heap_time"][ep, t + 1])
            assert np.array_equal(am, z["heap_amount"][ep, t + 1])
            assert np.array_equal(obs, z["obs"][ep, t + 1]), (name, ep, t)
        assert done


@pytest.mark.parametrize("name", ["linear", "simple_tight"])
def test_canonical_obs_equals_reference_obs_on_single_source_const_or_sorted(name):
    """On single-source chains the heap array stays time-sorted whenever lead times are constant."""
    meta, z = load("sc", name)
    env = SupplyChainOracle(meta["nodes_info"], **meta["env_kwargs"])
    lts = z["leadtimes"][0] if env.stoch_lt else None
    env.reset(demand=z["demand"][0], leadtimes=lts)
    assert np.array_equal(env.canonical_observation(), z["obs"][0, 0])
    if not env.stoch_lt:
        for t in range(meta["T"]):
            env.step(z["actions"][0, t])
            assert np.allclose(env.canonical_observation(), z["obs"][0, t + 1], rtol=0, atol=3e-16)


@pytest.mark.parametrize("name", beergame_cases())
def test_beergame_oracle_matches_reference_golden(name):
    meta, z = load("beergame", name)
    kw = dict(meta["kwargs"])
    if meta["version"] == 0:
        env = BeerGameOracle(version=0, levels=kw.get("levels", 4), customer_demand=kw.get("customer_demand"),
                             initial_inventory=kw.get("initial_inventory"), shipment_delays=kw.get("shipment_delays"),
                             initial_shipment=kw.get("initial_shipment_value", 4),
                             initial_orders=kw.get("initial_orders_value", 4),
                             inv_cost=kw.get("inv_cost", 1), backlog_cost=kw.get("backlog_cost", 2))
    else:
        kw2 = dict(kw)
        for k in ("customer_demand", "shipment_delays"):
            if k in kw2 and isinstance(kw2[k], list) and len(kw2[k]) == 2:
                kw2[k] = tuple(kw2[k])
        env = BeerGameOracle(version=2, **kw2)
    for ep in range(meta["episodes"]):
        obs = env.reset(demand=z["demand"][ep], delays=z["delays"][ep])
        assert np.array_equal(obs, z["obs"][ep, 0])
        for t in range(meta["T"]):
            obs, r, done, _ = env.step(z["actions"][ep, t])
            assert np.array_equal(obs, z["obs"][ep, t + 1]), (name, ep, t)
            assert r == z["reward"][ep, t]
            assert done == bool(z["done"][ep, t])
            assert np.array_equal(env.inventory, z["inventory"][ep, t + 1])
            assert np.array_equal(env.backlog, z["backlog"][ep, t + 1])


def test_beergame_survey_known_answers():
    """SURVEY.md Appendix E derived KATs (reference + shim)."""
    env = BeerGameOracle(version=0)
    env.reset()
    total, first = 0, []
    for t in range(35):
        obs, r, done, _ = env.step([0, 0, 0, 0])
        total += r
        if t < 6:
            first.append((list(obs), r))
    assert total == -312 and list(obs) == [0, 0, 0, 0] and done
    assert first[:4] == [([12, 12, 12, 12], -48)] * 4
    assert first[4] == ([8, 12, 12, 12], -44) and first[5] == ([4, 8, 12, 12], -36)
    env.reset()
    total = sum(env.step([1, 1, 1, 1])[1] for _ in range(35))
    assert total == -1848 and list(env._obs()) == [32, 29, 26, 23]
    for order, ret, final in ((4, -6968, [-12, 112, 112, 112]), (8, -1004, [116, 104, 104, 104]),
                              (29, -622423, [674, -14, 24, 62])):
        env = BeerGameOracle(version=2)
        env.reset()
        total = sum(env.step([order] * 4)[1] for _ in range(35))
        assert total == ret and list(env._obs()) == final
