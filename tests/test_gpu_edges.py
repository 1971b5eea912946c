"""Parity EDGES on the GPU: what the kernels do with inputs the reference would never produce (or would crash on).

  * injected lead times outside [1, max_leadtime] -- the reference draws clip(1 + Poisson(avg-1), 1, max)
    (supplychain_env.py:670-672): the host layer rejects them (ValueError, like the beer game's delays), and a raw
    scg_inject is clamped on read and flagged (scg_status), never an out-of-bounds ring / queue access;
  * an in-transit queue that overflows (only reachable through a hand-made state blob) is flagged, not silently dropped;
  * seed() is ordered against the handle's own streams (single-env host path);
  * a (node, product) nothing can be shipped to (max_ship == 0) observes zeros, not 0/0."""
import ctypes as C

import numpy as np
import pytest
import torch

import gym_supplychain_b200 as gsc
from gym_supplychain_b200 import _lib, presets

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["exact_f64", "fast_f64", "mixed", "fast_f32"])
def test_injected_leadtime_outside_the_ring_is_rejected_on_the_host(mode):
    env = gsc.make("supplychain-v0", num_envs=8, mode=mode)            # demo chain: stochastic lead times avg 2 / max 4
    assert env.scg_info.ring_depth == 4
    T, K = env.flat.T, env.flat.n_leadtimes
    lts = np.full((T, K), 2, dtype=np.int32)
    env.inject(leadtimes=lts)                                          # fine
    bad = lts.copy(); bad[3, 1] = 9
    with pytest.raises(ValueError, match="lead times must lie in"):
        env.inject(leadtimes=bad)
    bad[3, 1] = 0
    with pytest.raises(ValueError, match="lead times must lie in"):
        env.inject(leadtimes=bad)
    with pytest.raises(ValueError, match="must have shape"):
        env.inject(leadtimes=lts[:, :-1])
    with pytest.raises(ValueError, match="must have shape"):
        env.inject(demand=np.zeros((3, 3), dtype=np.int32))


@pytest.mark.parametrize("mode,generic", [("exact_f64", False), ("fast_f64", True), ("mixed", False), ("fast_f32", False),
                                          ("fast_f32", True)])
def test_raw_inject_of_a_bad_leadtime_is_clamped_and_flagged_not_out_of_bounds(mode, generic):
    """Through the raw C ABI (no host validation): lead time 9 on a ring of 4 behaves exactly like lead time 4 -- same
    rewards, same state -- and scg_status reports it; nothing outside the ring is written."""
    n, T = 64, 12

    def run(value):
        env = gsc.make("supplychain-v0", num_envs=n, mode=mode, auto_reset=False, force_generic=generic, seed=5)
        K = env.flat.n_leadtimes
        lts = torch.full((env.flat.T, K, n), 2, dtype=torch.int32, device="cuda")
        lts[:, 1::2, :] = value
        torch.cuda.synchronize()
        _lib.check(env._lib.scg_inject(env._h, None, C.c_void_p(lts.data_ptr())))
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        a = (torch.rand((T, n, env.n_actions), device="cuda", generator=g) * 2 - 1).to(env._io_dtype)
        o, r, d = env.rollout(T, a)
        torch.cuda.synchronize()
        return env, o.clone(), r.clone(), env.get_state().clone(), lts

    env9, o9, r9, s9, keep9 = run(9)
    with pytest.raises(_lib.ScgError, match="lead time"):
        env9.check()
    env9.check()                                                       # cleared by the first check
    env4, o4, r4, s4, keep4 = run(4)
    env4.check()                                                       # in range: no flag
    assert torch.equal(r9, r4) and torch.equal(o9, o4) and torch.equal(s9, s4)
    env0, o0, r0, s0, keep0 = run(0)                                   # below the range: clamps to 1
    with pytest.raises(_lib.ScgError, match="lead time"):
        env0.check()
    env1, o1, r1, s1, keep1 = run(1)
    assert torch.equal(r0, r1) and torch.equal(s0, s1)


def test_host_calls_report_the_flag_through_their_return_code():
    env = gsc.make("supplychain-v0", num_envs=4, mode="exact_f64", auto_reset=False)
    lts = torch.full((env.flat.T, env.flat.n_leadtimes, 4), 7, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    _lib.check(env._lib.scg_inject(env._h, None, C.c_void_p(lts.data_ptr())))
    A, O = env.n_actions, env.n_obs
    act = np.full((4, A), 0.5); obs = np.zeros((4, O)); rew = np.zeros(4); done = np.zeros(4, dtype=np.uint8)
    rc = env._lib.scg_step_host(env._h, act.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data)
    assert rc == -4 and b"lead time" in env._lib.scg_last_error()      # SCG_ESTATE
    assert np.isfinite(obs).all() and np.isfinite(rew).all()
    assert env._lib.scg_status(env._h, 1) == -4 and env._lib.scg_status(env._h, 0) == 0


def test_queue_overflow_from_a_hand_made_state_is_flagged():
    env = gsc.make("supplychain-2perstage-v0", num_envs=16, mode="exact_f64", auto_reset=False)
    env.reset()
    cap, lay = env.scg_info.heap_capacity, env._layout
    blob = env.get_state()
    n, NP = 16, env.flat.N * env.flat.P
    hlen = blob[lay[6]:lay[6] + NP * n * 4].view(torch.int32)
    htime = blob[lay[7]:lay[7] + NP * cap * n * 4].view(torch.int32)
    hlen.fill_(cap)                                                    # every queue full of entries that never arrive
    htime.fill_(1000)
    env.set_state(blob)
    env.step(torch.ones((16, env.n_actions), dtype=torch.float64, device="cuda"))
    with pytest.raises(_lib.ScgError, match="queue overflowed"):
        env.check()


def test_seed_is_ordered_against_the_single_env_host_path():
    """env.seed(s); env.reset() on the drop-in single env (library-owned streams): always episode 0 of key s, whatever
    was in flight -- the same trajectory on another handle and on every repetition (the demand drawn in-kernel shows in
    the observation and drives the rewards)."""
    other = gsc.make("supplychain-2perstage-v0", seed=5)
    other.seed(77)
    want = [other.reset().copy()]
    for _ in range(3):
        o, r, d, _ = other.step(np.full(other.n_actions, 0.25))
        want.append(np.append(o, r))
    env = gsc.make("supplychain-2perstage-v0", seed=1)
    for _ in range(20):
        env.reset(); env.step(np.zeros(env.n_actions))
        env.seed(77)
        got = [env.reset().copy()]
        for _ in range(3):
            o, r, d, _ = env.step(np.full(env.n_actions, 0.25))
            got.append(np.append(o, r))
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
    env.seed(78)
    assert not np.array_equal(env.reset(), want[0])                    # another key, another demand


def test_set_state_on_the_caller_stream_is_seen_by_the_host_calls():
    env = gsc.make("supplychain-2perstage-v0")                         # single env: host path on library streams
    env.reset()
    snap = env.get_state().clone()
    a = np.linspace(-0.5, 0.9, env.n_actions)
    first = env.step(a)
    for _ in range(10):
        env.step(-a)
        env.set_state(snap)                                            # async on the torch stream ...
        again = env.step(a)                                            # ... the host call must wait for it
        assert np.array_equal(first[0], again[0]) and first[1] == again[1]


@pytest.mark.parametrize("mode", ["fast_f64", "mixed", "fast_f32"])
def test_node_nothing_ships_to_observes_zeros_not_nan(mode):
    """A wholesaler whose inbound edges all have ship capacity 0 has max_ship == 0: its in-transit buckets read -1
    (empty queue, supplychain_env.py:438-439) in every tolerance mode, on both kernels.  (exact_f64 reproduces the
    reference bit for bit, including its own 0/0 = nan while the zero-amount INITIAL shipments are still queued.)"""
    info, kw = presets.two_per_stage(ship_capacity=300)
    for name in ("Factory1", "Factory2"):
        info[name] = dict(info[name], ship_capacity=[0, 300])          # nothing can reach WholeSaler1
    info["WholeSaler1"] = dict(info["WholeSaler1"], initial_shipments=[[0, 0]])
    for generic in (False, True):
        env = gsc.SupplyChainEnv(info, num_envs=32, mode=mode, auto_reset=False, force_generic=generic, **kw)
        o0 = env.reset()
        a = torch.rand((20, 32, env.n_actions), device="cuda").to(env._io_dtype) * 2 - 1
        o, r, d = env.rollout(20, a)
        assert torch.isfinite(o0).all() and torch.isfinite(o).all() and torch.isfinite(r).all()
        # obs layout: R*P demands, then per node [stock, bucket(t+1), bucket(rest)]; WholeSaler1 is node 4
        base = env.flat.n_retailers * env.flat.P + 4 * 3
        assert bool((o[:, :, base + 1:base + 3] == -1).all())


@pytest.mark.parametrize("env_id,kw", [("supplychain-ntom-v0", {}), ("supplychain-linear-v0", {}), ("supplychain-v0", {}),
                                       ("supplychain-ntom-v0", dict(nodes_per_echelon=[5, 4, 7, 10]))])
@pytest.mark.parametrize("mode", ["mixed", "fast_f32"])
@pytest.mark.parametrize("n", [32 * 5 + 7, 32 * 5 + 8])
def test_bulk_copy_pipeline_edges(env_id, kw, mode, n):
    """The action-tile pipeline of the rollout kernels (one or two tiles per warp, requests issued by an elected lane): a rollout of
    1, 2, 3 or 7 periods, with and without the observation output, and a batch whose last warp is ragged, all agree with the
    same periods taken through step()."""
    # the last warp is ragged; with 167 envs some shapes lose the 16-byte alignment the bulk copies need (plain warp copies instead)
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    ref = gsc.make(env_id, num_envs=n, mode=mode, seed=4, auto_reset=True, **kw)
    ref.reset()
    acts = torch.empty((13, n, ref.n_actions), device="cuda").uniform_(-1, 1, generator=g)
    want_o, want_r = [], []
    for t in range(13):
        o, r, d, _ = ref.step(acts[t])
        want_o.append(o.clone()); want_r.append(r.clone())
    want_o, want_r = torch.stack(want_o), torch.stack(want_r)
    for write_obs in (True, False):
        env = gsc.make(env_id, num_envs=n, mode=mode, seed=4, auto_reset=True, **kw)
        env.reset()
        t0 = 0
        for T in (1, 2, 3, 7):
            o, r, d = env.rollout(T, acts[t0:t0 + T].contiguous(), write_obs=write_obs)
            assert torch.equal(r, want_r[t0:t0 + T]), (write_obs, T)
            if write_obs:
                assert torch.equal(o, want_o[t0:t0 + T]), T
            else:
                assert o is None
            t0 += T
        env.check()                                             # no device-side flag was raised
