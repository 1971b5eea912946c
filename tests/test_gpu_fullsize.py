"""Parity at the BASELINE.json sizes (GPU suite).  The oracle cannot run a million envs, so each test runs the
full-size batch on the GPU and then REPLAYS a handful of sampled envs on the CPU oracle with the same actions:
the Philox demands an env saw are recovered exactly from its observation stream (the first R*P entries of every
observation are the next period's integer demands), lead times are constant, so the oracle can be driven with
identical inputs.  Plus the size-independent invariants (one `done` per env per episode, return statistics equal
the sum of rewards, obs in [-1, 1])."""
import numpy as np
import pytest
import torch

import gym_supplychain_b200 as gsc
from gym_supplychain_b200 import presets
from oracle.beergame_oracle import BeerGameOracle
from oracle.sc_oracle import SupplyChainOracle

pytestmark = pytest.mark.gpu


def _replay_sc(env, preset, T_total, obs0, obs, rew, acts, idx, rtol):
    """Replay env `idx` of the batch on the oracle (first episode only) and compare rewards / canonical obs."""
    info, kw = preset
    f = env.flat
    RP = f.n_retailers * f.P
    lo, span = f.demand_lo[0], f.demand_hi[0] - f.demand_lo[0]
    T = min(T_total, f.T)
    dem = np.zeros((f.T + 1, f.n_retailers, f.P), dtype=np.int64)
    seq = np.concatenate([obs0[idx:idx + 1, :RP].double().cpu().numpy(), obs[:T, idx, :RP].double().cpu().numpy()])
    d = (seq + 1) / 2 * span + lo
    assert np.abs(d - np.rint(d)).max() < 1e-3
    dem[:T + 1] = np.rint(d).reshape(T + 1, f.n_retailers, f.P)
    ora = SupplyChainOracle(info, **kw)
    ora.reset(demand=dem)
    a = acts[:T, idx].double().cpu().numpy()
    r_gpu = rew[:T, idx].double().cpu().numpy()
    o_gpu = obs[:T, idx].double().cpu().numpy()
    for t in range(T):
        _, r, done, _ = ora.step(a[t])
        assert abs(r_gpu[t] - r) <= rtol * max(abs(r), 1.0), (idx, t, r_gpu[t], r)
        if not (done and env.auto_reset):
            assert np.abs(o_gpu[t] - ora.canonical_observation()).max() <= 5e-6, (idx, t)


@pytest.mark.parametrize("mode,rtol", [("mixed", 1e-6), ("fast_f32", 5e-6)])
def test_config4_ntom_one_million_envs(mode, rtol):
    n, T = 1 << 20, 64
    env = gsc.make("supplychain-ntom-v0", num_envs=n, seed=21, mode=mode)
    obs0 = env.reset()
    acts = torch.rand(T, n, env.n_actions, device="cuda") * 2 - 1
    obs, rew, done = env.rollout(T, acts)
    assert obs.abs().max() <= 1 and torch.isfinite(rew).all() and not done.any()
    for idx in (0, 31, 32, 524_287, n - 1):
        _replay_sc(env, presets.n_per_stage(nodes_per_echelon=3), T, obs0, obs, rew, acts, idx, rtol)


def test_config3_2perstage_one_million_envs_128_period_rollout():
    n, T = 1 << 20, 128
    env = gsc.make("supplychain-2perstage-v0", num_envs=n, seed=4, mode="mixed", total_time_steps=100)
    obs0 = env.reset()
    acts = torch.rand(T, n, env.n_actions, device="cuda") * 2 - 1
    obs, rew, done = env.rollout(T, acts)
    d = done.bool()
    assert d[99].all() and int(d.sum()) == n                       # one episode end (t = 100) inside 128 periods
    assert obs.abs().max() <= 1 and (rew <= 0).all()
    stats = env.return_stats().cpu().numpy()
    ret = rew[:100].double().sum(0)
    # (the statistics accumulate the fp64 rewards, the rollout returns them rounded to fp32: 1e-7 relative)
    assert stats[0] == n and np.isclose(stats[1], float(ret.sum()), rtol=1e-7)
    assert np.isclose(stats[3], float(ret.min()), rtol=1e-7) and np.isclose(stats[4], float(ret.max()), rtol=1e-7)
    for idx in (5, 777_777, n - 2):
        _replay_sc(env, presets.two_per_stage(total_time_steps=100), T, obs0, obs, rew, acts, idx, 1e-6)


def test_config2_linear_65536_envs_step_api_equals_rollout():
    n, T = 65_536, 40
    a = gsc.make("supplychain-linear-v0", num_envs=n, seed=8, mode="mixed")
    b = gsc.make("supplychain-linear-v0", num_envs=n, seed=8, mode="mixed")
    obs0 = a.reset(); b.reset()
    acts = torch.rand(T, n, 4, device="cuda") * 2 - 1
    obs, rew, done = a.rollout(T, acts)
    for t in range(T):
        o, r, d, _ = b.step(acts[t])
        assert torch.equal(o, obs[t]) and torch.equal(r, rew[t])
    for idx in (1, 40_000):
        _replay_sc(a, presets.n_per_stage(nodes_per_echelon=1, num_products=1), T, obs0, obs, rew, acts, idx, 1e-6)


def test_config5_mixed_sweep_beergame_v2_and_supplychain_v0_with_auto_reset():
    """Per-GPU share of the 16M-env sweep at 8 GPUs: 1M beergame-v2 + 1M supplychain-v0 envs, two handles on one GPU,
    auto-reset on, >= 2 full episodes of each."""
    n = 1 << 20
    bg = gsc.BeerGameEnv2(customer_demand=(0, 16), shipment_delays=(0, 5), num_envs=n, seed=1, auto_reset=True)
    sc = gsc.make("supplychain-v0", num_envs=n, seed=1, mode="mixed", auto_reset=True, total_time_steps=90)
    assert sc.scg_info.kernel_variant == 6
    bg.reset(); sc.reset()
    o, r, d = bg.rollout(70)                                       # 2 episodes of 35 weeks, in-kernel random orders
    assert int(d.sum()) == 2 * n and d[34].all() and d[69].all() and torch.isfinite(r.double()).all()
    st = bg.return_stats().cpu().numpy()
    assert st[0] == 2 * n and np.isclose(st[1], float(r.double().sum()), rtol=1e-9)
    total_done = 0
    ret_sum = 0.0
    for chunk in range(6):                                         # 180 periods = 2 episodes of 90, 30 at a time
        o, r, d = sc.rollout(30)
        total_done += int(d.sum()); ret_sum += float(r.double().sum())
        assert o.abs().max() <= 1 and (r <= 0).all()
    st = sc.return_stats().cpu().numpy()
    assert total_done == 2 * n and st[0] == 2 * n and np.isclose(st[1], ret_sum, rtol=1e-6)


def test_beergame_eight_million_envs_sampled_replay_is_bit_exact():
    n, T = 8 << 20, 35
    env = gsc.make("beergame-v2", num_envs=n, auto_reset=True)
    env.reset()
    acts = torch.randint(0, 30, (T, n, 4), dtype=torch.int32, device="cuda")
    obs, rew, done = env.rollout(T, acts)
    assert done[-1].all() and not done[:-1].any()
    for idx in (0, 255, 256, 4_194_303, n - 1):
        ora = BeerGameOracle(version=2)
        ora.reset()
        a = acts[:, idx].cpu().numpy()
        for t in range(T):
            o, r, dd, _ = ora.step(a[t])
            if t < T - 1:                                          # the last obs is the auto-reset obs of the next episode
                assert np.array_equal(o, obs[t, idx].cpu().numpy())
            assert r == int(rew[t, idx]) and dd == bool(done[t, idx])
    assert np.array_equal(obs[-1, 12345].cpu().numpy(), [112, 112, 112, 112])   # reset obs: 100 + 12 - 0
