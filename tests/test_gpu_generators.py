"""In-kernel stochastic inputs vs samples of the reference's own generators.

The kernels draw from counter-based Philox streams keyed (seed, env, episode, period); the reference draws whole-episode
tables from one MT19937 stream, so parity is distributional: ~1e6 in-kernel draws per case against the histograms that
oracle/make_demand_fixtures.py recorded from `generate_demand` (demands_generator.py:33-89: uniform :33-36, clipped
normal :38-49, sinusoid + normal / + uniform perturbation :51-89) and from the lead-time tables of a reference env
(supplychain_env.py:664-672).  Checked per recorded period: Kolmogorov-Smirnov distance of the two empirical
distributions (bar: the 0.1 % critical value for these sample sizes), the largest per-value probability difference,
the mass ON the clip boundaries lo / hi, mean and standard deviation."""
import json
import os

import numpy as np
import pytest
import torch

import gym_supplychain_b200 as gsc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fixture():
    z = np.load(os.path.join(ROOT, "tests", "golden", "generator_hist.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def _compare(gpu_counts, ref_counts, lo, what):
    n1, n2 = gpu_counts.sum(), ref_counts.sum()
    p1, p2 = gpu_counts / n1, ref_counts / n2
    ks = np.abs(np.cumsum(p1) - np.cumsum(p2)).max()
    crit = 1.95 * np.sqrt((n1 + n2) / (n1 * n2))                          # two-sample KS, alpha = 0.001
    assert ks < crit, (what, "KS", ks, crit)
    # every value's probability within 5 sigma of the pooled binomial error
    pooled = (gpu_counts + ref_counts) / (n1 + n2)
    sig = np.sqrt(pooled * (1 - pooled) * (1 / n1 + 1 / n2))
    assert (np.abs(p1 - p2) <= 5 * sig + 1e-9).all(), (what, "per-value", np.abs(p1 - p2).max())
    for edge in (0, -1):                                                   # clip-boundary mass (:44, :86)
        assert abs(p1[edge] - p2[edge]) <= 5 * sig[edge] + 1e-9, (what, "boundary", edge, p1[edge], p2[edge])
    v = np.arange(len(p1)) + lo
    m1, m2 = (p1 * v).sum(), (p2 * v).sum()
    s1, s2 = np.sqrt((p1 * (v - m1) ** 2).sum()), np.sqrt((p2 * (v - m2) ** 2).sum())
    assert abs(m1 - m2) < 5 * s2 * np.sqrt(1 / n1 + 1 / n2) + 1e-9, (what, "mean", m1, m2)
    assert abs(s1 - s2) < 0.01 * s2 + 1e-9, (what, "std", s1, s2)


def _demand_rows(env, rows, lo, hi):
    """Demand rows `rows` of every (env, retailer): row t is the 'next demand' shown in the obs after period t."""
    R = env.flat.n_retailers * env.flat.P
    out = {}

    def take(row, obs_rp):
        d = ((obs_rp.double() + 1) / 2 * (hi - lo) + lo).round().long().flatten()
        assert int(d.min()) >= lo and int(d.max()) <= hi
        out[row] = torch.bincount(d - lo, minlength=hi - lo + 1).cpu().numpy()
    o0 = env.reset()
    if 0 in rows:
        take(0, o0[:, :R])
    t, last = 0, max(rows)
    while t < last:
        c = min(30, last - t)
        obs, _, _ = env.rollout(c)                                         # in-kernel random policy
        for r in rows:
            if t < r <= t + c:
                take(r, obs[r - t - 1, :, :R])
        t += c
    return out


@pytest.mark.parametrize("name", ["uniform", "normal", "sin_normal", "sin_uniform"])
@pytest.mark.parametrize("mode,generic", [("mixed", False), ("fast_f32", True)])
def test_in_kernel_demand_matches_the_reference_generator(fixture, name, mode, generic):
    """All four demand kinds, on the compile-time kernel and on the runtime-topology kernel."""
    z, meta = fixture
    case = meta["cases"][name]
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in case["kw"].items()}
    lo, hi = kw["demand_range"]
    n = 1 << 18
    env = gsc.make("supplychain-2perstage-v0", num_envs=n, seed=11, mode=mode, auto_reset=False, force_generic=generic, **kw)
    assert (env.scg_info.kernel_variant == 0) == generic
    if case["rows"] is None:                                               # the same distribution in every period
        got = _demand_rows(env, [0, 1], lo, hi)
        _compare(got[0] + got[1], z[name][0].astype(np.int64), lo, name)
    else:
        rows = case["rows"] if not generic else case["rows"][:3]           # the slow kernel: the first three recorded rows
        got = _demand_rows(env, rows, lo, hi)
        for i, r in enumerate(rows):
            _compare(got[r], z[name][i].astype(np.int64), lo, "%s row %d" % (name, r))


@pytest.mark.parametrize("name", ["lt_avg2_max4", "lt_avg3_max6"])
def test_in_kernel_leadtimes_match_the_reference_tables(fixture, name):
    """1 + Poisson(avg - 1) clipped to [1, max]: the arrival time of the first supply order of every env."""
    z, meta = fixture
    avg, mx = meta["leadtimes"][name]
    n = 1 << 18
    env = gsc.make("supplychain-v0", num_envs=n, seed=2, mode="exact_f64", auto_reset=False, avg_leadtime=avg, max_leadtime=mx)
    env.reset()
    env.rollout(1, torch.ones(1, n, env.n_actions, dtype=torch.float64, device="cuda"))
    st = env.state_dict()
    lts = torch.cat([st["heap_time"][s, 0, 0, :] - 1 for s in (0, 1)]).long()      # both suppliers: one entry each
    assert bool((st["heap_len"][:2, 0] == 1).all())
    got = torch.bincount(lts, minlength=mx + 1).cpu().numpy()
    assert got[0] == 0 and got.sum() == 2 * n
    _compare(got[1:], z[name][1:].astype(np.int64), 1, name)
