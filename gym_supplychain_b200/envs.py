"""Env classes with the reference's names, kwargs and (obs, reward, done, info) contract, running on
the CUDA extension (libscg.so) through the C ABI of include/scg.h.

Two ways to use every class:

* drop-in single env (default, `num_envs=None`): numpy in / numpy out, exactly the old-gym API of
  the reference -- `reset() -> obs[O]`, `step(a) -> (obs[O], reward, done, info)`, `seed(s)`,
  `action_space`, `observation_space`.  Runs one env on the GPU in `exact_f64` mode, whose
  dynamics AND observation are bit-identical to the reference for identical demand/lead-time
  sequences (inject them with `inject()`; the built-in generators are counter-based Philox streams,
  not the reference's MT19937 stream).
* vector env (`num_envs=N`): torch CUDA tensors in / out, `reset() -> obs[N,O]`,
  `step(actions[N,A]) -> (obs[N,O], reward[N], done[N], infos)`, plus the fused
  `rollout(T, actions=None) -> (obs[T,N,O], reward[T,N], done[T,N])`.  Default mode `fast_f32`.

Reference classes mirrored (paths relative to the reference repo):
  SupplyChainEnv ............................. gym_supplychain/envs/supplychain_env.py:478-813
  SupplyChain2perStageEnv / ...SeasonalEnv ... gym_supplychain/envs/supplychain_2perstage_env.py
  SupplyChainNPerStage ....................... gym_supplychain/envs/supplychain_Nperstage_env.py
  SupplyChainMultiProduct* ................... gym_supplychain/envs/supplychain_multiproduct_env.py
  BeerGameEnv / BeerGameEnv2 ................. gym_supplychain/envs/beergame_env.py / beergame2_env.py
"""
import ctypes as C

import numpy as np

from . import _lib, presets
from .spaces import Box, MultiDiscrete
from .spec import FlatSpec


def _torch():
    import torch
    return torch


class _ScgEnvBase(object):
    """Shared plumbing: handle lifetime, tensors, reset/step/rollout through the C ABI."""
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None

    def _init_handle(self, create_fn, c_spec, num_envs, device, mode, auto_reset, seed, first_env, force_generic=False,
                     build_info=False):
        torch = _torch()
        self._single = num_envs is None
        self.num_envs = 1 if num_envs is None else int(num_envs)
        lib = _lib.load()                                       # raises loudly if the extension is missing
        if not torch.cuda.is_available():
            raise _lib.ScgError("no CUDA device: gym_supplychain_b200 has no CPU fallback")
        self.device = torch.device("cuda:0" if device is None else device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._mode_name = mode
        opt = _lib.ScgOptions()
        opt.n_envs, opt.first_env = self.num_envs, int(first_env)
        opt.seed = int(seed if seed is not None else 0) & (2 ** 64 - 1)
        opt.device = self.device.index
        opt.mode = _lib.MODES[mode]
        opt.auto_reset = 1 if auto_reset else 0
        opt.flags = (_lib.FLAG_FORCE_GENERIC if force_generic else 0) | (_lib.FLAG_BUILD_INFO if build_info else 0)
        self.build_info = bool(build_info)
        self.auto_reset = bool(auto_reset)
        self._lib = lib
        self._h = C.c_void_p()
        _lib.check(create_fn(C.byref(c_spec), C.byref(opt), C.byref(self._h)))
        info = _lib.ScgInfo()
        _lib.check(lib.scg_info(self._h, C.byref(info)))
        self.scg_info = info
        self.n_actions, self.n_obs = info.n_actions, info.n_obs
        self._inj = (None, None)
        self._stats_buf = torch.zeros(5, dtype=torch.float64, device=self.device)
        lay = (C.c_int64 * 8)()
        _lib.check(lib.scg_state_layout(self._h, C.byref(lay)))
        self._layout = list(lay)
        if self._single:
            # the drop-in single env talks numpy: preallocated host buffers and the *_host entry points (one call,
            # one synchronisation per step -- no torch tensors on this path)
            npdt = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32}[self._io_dtype]
            self._np_act = np.zeros(self.n_actions, dtype=npdt)
            self._np_obs = np.zeros(self.n_obs, dtype=npdt)
            self._np_rew = np.zeros(1, dtype=npdt)
            self._np_done = np.zeros(1, dtype=np.uint8)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.scg_destroy(h)
            self._h = C.c_void_p()

    close = __del__

    # ---- helpers ----
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t):
        return C.c_void_p(0 if t is None else t.data_ptr())

    def _as_actions(self, actions, lead_shape):
        torch = _torch()
        a = torch.as_tensor(actions)
        a = a.to(device=self.device, dtype=self._io_dtype).contiguous()
        want = tuple(lead_shape) + (self.n_actions,)
        if a.shape[-1] > self.n_actions and a.shape[:-1] == tuple(lead_shape):
            a = a[..., :self.n_actions].contiguous()            # trailing extras are ignored by the reference (:716)
        if tuple(a.shape) != want:
            raise ValueError("actions must have shape %s, got %s" % (want, tuple(a.shape)))
        return a

    # ---- vector API ----
    def _reset_tensors(self, mask=None):
        torch = _torch()
        obs = torch.empty((self.num_envs, self.n_obs), dtype=self._io_dtype, device=self.device)
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.scg_reset(self._h, self._ptr(m), self._ptr(obs), self._stream()))
        return obs

    def _step_tensors(self, actions):
        torch = _torch()
        a = self._as_actions(actions, (self.num_envs,))
        obs = torch.empty((self.num_envs, self.n_obs), dtype=self._io_dtype, device=self.device)
        rew = torch.empty((self.num_envs,), dtype=self._io_dtype, device=self.device)
        done = torch.empty((self.num_envs,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.scg_step(self._h, self._ptr(a), self._ptr(obs), self._ptr(rew), self._ptr(done), self._stream()))
        return obs, rew, done.bool()

    def rollout(self, T, actions=None, out=None, write_obs=True):
        """Fused T-period rollout.  actions [T,N,A] or None (in-kernel uniform random policy).
        `out=(obs, reward, done)` reuses caller tensors.  Returns (obs[T,N,O] | None, reward[T,N], done[T,N] uint8)."""
        torch = _torch()
        T = int(T)
        a = None if actions is None else self._as_actions(actions, (T, self.num_envs))
        if out is not None:
            obs, rew, done = out
            # the kernels write T*N*O (T*N) elements of the handle's I/O dtype through raw pointers: a buffer of the
            # wrong dtype / shape / device would be an out-of-bounds device write, so check before launching
            for name, x, shape, dt in (("obs", obs, (T, self.num_envs, self.n_obs), self._io_dtype),
                                       ("reward", rew, (T, self.num_envs), self._io_dtype),
                                       ("done", done, (T, self.num_envs), torch.uint8)):
                if x is None and name == "obs":
                    continue
                if (not torch.is_tensor(x) or x.dtype != dt or tuple(x.shape) != shape or not x.is_contiguous()
                        or x.device != torch.device(self.device)):
                    raise ValueError("rollout(out=...): %s must be a contiguous %s tensor of shape %s on %s"
                                     % (name, dt, shape, self.device))
        else:
            obs = torch.empty((T, self.num_envs, self.n_obs), dtype=self._io_dtype, device=self.device) if write_obs else None
            rew = torch.empty((T, self.num_envs), dtype=self._io_dtype, device=self.device)
            done = torch.empty((T, self.num_envs), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.scg_rollout(self._h, T, self._ptr(a), self._ptr(obs), self._ptr(rew), self._ptr(done), self._stream()))
        return obs, rew, done

    def rollout_host(self, T, actions, obs, reward, done):
        """Same, with HOST (ideally pinned) torch/numpy buffers: H2D, kernel and D2H all happen inside."""
        def hp(x):
            if x is None:
                return C.c_void_p(0)
            return C.c_void_p(x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data)
        _lib.check(self._lib.scg_rollout_host(self._h, int(T), hp(actions), hp(obs), hp(reward), hp(done)))

    def return_stats(self, clear=True):
        """float64[5] on the device: count, sum, sum of squares, min, max of finished-episode returns."""
        with _torch().cuda.device(self.device):
            _lib.check(self._lib.scg_return_stats(self._h, self._ptr(self._stats_buf), 1 if clear else 0, self._stream()))
        return self._stats_buf.clone()

    def get_state(self):
        torch = _torch()
        blob = torch.empty(self.scg_info.state_bytes, dtype=torch.uint8, device=self.device)
        _lib.check(self._lib.scg_get_state(self._h, self._ptr(blob), self._stream()))
        return blob

    def set_state(self, blob):
        torch = _torch()
        if (not torch.is_tensor(blob) or not blob.is_cuda or blob.dtype != torch.uint8
                or blob.numel() != self.scg_info.state_bytes):
            raise ValueError("set_state: need a CUDA uint8 tensor of %d bytes (get_state())" % self.scg_info.state_bytes)
        _lib.check(self._lib.scg_set_state(self._h, self._ptr(blob.contiguous()), self._stream()))

    def check(self, clear=True):
        """Synchronise and raise ScgError if a kernel flagged a device-side check (scg_status): an injected lead time /
        delay out of range (clamped, never out of bounds) or an in-transit queue overflow."""
        _lib.check(self._lib.scg_status(self._h, 1 if clear else 0))

    def launch_count(self):
        return int(self._lib.scg_launch_count(self._h))

    def seed(self, seed=None):
        """New generator key (reference: supplychain_env.py:811-813 also seeds the action space with 0)."""
        _lib.check(self._lib.scg_seed(self._h, int(seed if seed is not None else 0) & (2 ** 64 - 1)))
        if self.action_space is not None:
            self.action_space.seed(0)

    # ---- gym API ----
    def reset(self, mask=None):
        if self._single:
            _lib.check(self._lib.scg_reset_host(self._h, None, self._np_obs.ctypes.data))
            return self._np_to_obs()
        return self._reset_tensors(mask)

    def step(self, action):
        if self._single:
            a = np.asarray(action).reshape(-1)
            if a.size < self.n_actions:
                raise ValueError("actions must have shape (%d,), got %s" % (self.n_actions, tuple(np.shape(action))))
            self._np_act[:] = a[:self.n_actions]                 # trailing extras are ignored by the reference (:716)
            _lib.check(self._lib.scg_step_host(self._h, self._np_act.ctypes.data, self._np_obs.ctypes.data,
                                               self._np_rew.ctypes.data, self._np_done.ctypes.data))
            return self._np_to_obs(), self._np_to_reward(), bool(self._np_done[0]), self._info()
        obs, rew, done = self._step_tensors(action)
        return obs, rew, done, self._info()

    def _np_to_obs(self):
        return self._np_obs.astype(np.float64)                   # a fresh float64 array each call, like the reference

    def _np_to_reward(self):
        return float(self._np_rew[0])

    def _info(self):
        return {}

    def render(self, mode="human"):
        raise NotImplementedError("render() is out of scope of the B200 hot path (SURVEY.md C13)")


class SupplyChainEnv(_ScgEnvBase):
    """Generic multi-echelon, multi-product, n-to-m supply chain (supplychain_env.py:478-813)."""

    def __init__(self, nodes_info, num_products=1, unmet_demand_cost=1000, exceeded_stock_capacity_cost=1000,
                 exceeded_process_capacity_cost=1000, exceeded_ship_capacity_cost=1000, demand_config_by_product=False,
                 demand_range=(10, 20), demand_std=None, demand_sen_peaks=None, avg_demand_range=None,
                 processing_ratio=3, stochastic_leadtimes=False, avg_leadtime=2, max_leadtime=2, total_time_steps=360,
                 seed=None, build_info=False, demand_perturb_norm=False,
                 num_envs=None, device=None, mode=None, auto_reset=None, first_env=0, force_generic=False):
        self.flat = FlatSpec(nodes_info, num_products=num_products, unmet_demand_cost=unmet_demand_cost,
                             exceeded_stock_capacity_cost=exceeded_stock_capacity_cost,
                             exceeded_process_capacity_cost=exceeded_process_capacity_cost,
                             exceeded_ship_capacity_cost=exceeded_ship_capacity_cost,
                             demand_config_by_product=demand_config_by_product, demand_range=demand_range,
                             demand_std=demand_std, demand_sen_peaks=demand_sen_peaks, avg_demand_range=avg_demand_range,
                             processing_ratio=processing_ratio, stochastic_leadtimes=stochastic_leadtimes,
                             avg_leadtime=avg_leadtime, max_leadtime=max_leadtime, total_time_steps=total_time_steps,
                             demand_perturb_norm=demand_perturb_norm)
        self.num_products = num_products
        self.total_time_steps = total_time_steps
        self.stochastic_leadtimes, self.avg_leadtime, self.max_leadtime = stochastic_leadtimes, avg_leadtime, max_leadtime
        self.action_space = Box(low=-1.0, high=1.0, shape=(self.flat.n_actions,))          # :625
        self.observation_space = Box(low=-1.0, high=1.0, shape=(self.flat.n_obs,))         # :626
        if mode is None:
            mode = "exact_f64" if num_envs is None else "fast_f32"
        if auto_reset is None:
            auto_reset = num_envs is not None
        if mode != "exact_f64":
            hits = self.flat.negative_ship_capacity_reachable()
            if hits:
                import warnings
                warnings.warn("mode %r on a chain where a saturated factory edge can reach a negative ship capacity (reference "
                              "quirk, supplychain_env.py:316-328; e.g. %s product %d -> %s): from there on only mode 'exact_f64' "
                              "follows the reference step for step" % ((mode,) + hits[0]), RuntimeWarning, stacklevel=3)
        torch = _torch()
        self._io_dtype = torch.float32 if mode in ("fast_f32", "mixed") else torch.float64
        self._init_handle(_lib.load().scg_sc_create, self.flat.c_struct(), num_envs, device, mode, auto_reset, seed, first_env,
                          force_generic=force_generic, build_info=build_info)
        assert self.n_actions == self.flat.n_actions and self.n_obs == self.flat.n_obs

    def inject(self, demand=None, leadtimes=None):
        """Replace the on-device generators by explicit tables for the following episodes.
        demand: int [T+1, R, P] (same for every env) or [N, T+1, R, P]; leadtimes: int [T, K] or [N, T, K]."""
        torch = _torch()
        T, R, P = self.flat.T, self.flat.n_retailers, self.flat.P

        def prep(x, shape, what):
            if x is None:
                return None
            x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(torch.int32)
            if tuple(x.shape) not in (tuple(shape), (self.num_envs,) + tuple(shape)):
                raise ValueError("inject: %s must have shape %s or %s, got %s"
                                 % (what, tuple(shape), (self.num_envs,) + tuple(shape), tuple(x.shape)))
            if x.dim() == len(shape):
                x = x.unsqueeze(0).expand((self.num_envs,) + tuple(shape))
            return x
        d = prep(demand, (T + 1, R, P), "demand")
        lt = prep(leadtimes, (T, self.flat.n_leadtimes), "leadtimes") if self.flat.stochastic_leadtimes else None
        if lt is not None and lt.numel() > 0:
            # the reference draws clip(1 + Poisson(avg - 1), 1, max_leadtime) (supplychain_env.py:670-672); the in-transit
            # ring / queues are sized for exactly that range (the kernels clamp and flag anything else, scg_status)
            lo, hi = int(lt.min()), int(lt.max())
            if lo < 1 or hi > self.flat.max_leadtime:
                raise ValueError("inject: lead times must lie in [1, max_leadtime=%d], got [%d, %d]"
                                 % (self.flat.max_leadtime, lo, hi))

        def lay(x):            # [env][rows][cols...] -> device [rows][cols][env]
            return None if x is None else x.reshape(self.num_envs, x.shape[1], -1).permute(1, 2, 0).contiguous().to(self.device)
        d, lt = lay(d), lay(lt)
        self._inj = (d, lt)                                     # keep the tensors alive
        torch.cuda.current_stream(self.device).synchronize()    # the *_host calls run on the handle's own stream
        _lib.check(self._lib.scg_inject(self._h, self._ptr(d), self._ptr(lt)))

    def _to_single_obs(self, obs):
        return obs[0].double().cpu().numpy()

    def _to_single_reward(self, rew):
        return float(rew[0].item())

    def _info(self):
        """info dict of step(): {} or, with build_info=True, {'sc_episode': {'rewards', 'costs', 'units'}} with the
        reference's 8 keys (supplychain_env.py:684-695, :793-796).  Single env: python floats / lists per product;
        vector env: tensors (rewards [N], each key [P, N])."""
        if not self.build_info:
            return {}
        torch = _torch()
        P, n = self.flat.P, self.num_envs
        buf = torch.empty((16 * P, n), dtype=torch.float64, device=self.device)
        _lib.check(self._lib.scg_get_info(self._h, self._ptr(buf), self._stream()))
        rewards = self.state_dict()["ep_ret"].clone()
        costs = {k: buf[i * P:(i + 1) * P] for i, k in enumerate(_lib.INFO_KEYS)}
        units = {k: buf[(8 + i) * P:(9 + i) * P] for i, k in enumerate(_lib.INFO_KEYS)}
        if self._single:
            costs = {k: [float(x) for x in v[:, 0].tolist()] for k, v in costs.items()}
            units = {k: [float(x) for x in v[:, 0].tolist()] for k, v in units.items()}
            rewards = float(rewards[0])
        return {"sc_episode": {"rewards": rewards, "costs": costs, "units": units}}

    def state_dict(self):
        """Decode the raw state blob (scg_state_layout): time step, episode, stocks [N,P,n] and either the
        lead-time ring [N,P,ring,n] (fast modes; slot k = arriving at t+1+k) or the in-transit queues
        (exact mode: heap_len [N,P,n], heap_time / heap_amount [N,P,cap,n])."""
        torch = _torch()
        blob = self.get_state()
        torch.cuda.synchronize(self.device)
        n, N, P = self.num_envs, self.flat.N, self.flat.P
        lay, info = self._layout, self.scg_info
        real = torch.float32 if self._mode_name == "fast_f32" else torch.float64
        rs = 4 if real == torch.float32 else 8

        def view(off, count, dtype, size):
            return blob[off:off + count * size].view(dtype)
        out = {"t": view(lay[4], n, torch.int32, 4), "episode": view(lay[5], n, torch.int32, 4),
               "ep_ret": view(lay[0], n, torch.float64, 8),
               "stock": view(lay[2], N * P * n, real, rs).view(N, P, n)}
        if self._mode_name == "exact_f64":
            cap = info.heap_capacity
            out["heap_len"] = view(lay[6], N * P * n, torch.int32, 4).view(N, P, n)
            out["heap_time"] = view(lay[7], N * P * cap * n, torch.int32, 4).view(N, P, cap, n)
            out["heap_amount"] = view(lay[1], N * P * cap * n, torch.float64, 8).view(N, P, cap, n)
        else:
            ring = info.ring_depth
            out["slot"] = view(lay[3], N * P * ring * n, real, rs).view(N, P, ring, n)
        return out


class SupplyChain2perStageEnv(SupplyChainEnv):
    def __init__(self, num_envs=None, device=None, mode=None, auto_reset=None, first_env=0, force_generic=False, **kwargs):
        info, kw = presets.two_per_stage(**kwargs)
        super().__init__(info, num_envs=num_envs, device=device, mode=mode, auto_reset=auto_reset, first_env=first_env,
                         force_generic=force_generic, **kw)


class SupplyChain2perStageSeasonalEnv(SupplyChainEnv):
    def __init__(self, num_envs=None, device=None, mode=None, auto_reset=None, first_env=0, force_generic=False, **kwargs):
        info, kw = presets.two_per_stage_seasonal(**kwargs)
        super().__init__(info, num_envs=num_envs, device=device, mode=mode, auto_reset=auto_reset, first_env=first_env,
                         force_generic=force_generic, **kw)


class SupplyChainNPerStage(SupplyChainEnv):
    def __init__(self, num_envs=None, device=None, mode=None, auto_reset=None, first_env=0, force_generic=False, **kwargs):
        info, kw = presets.n_per_stage(**kwargs)
        super().__init__(info, num_envs=num_envs, device=device, mode=mode, auto_reset=auto_reset, first_env=first_env,
                         force_generic=force_generic, **kw)


class SupplyChainMultiProduct(SupplyChainEnv):
    _preset = staticmethod(presets.multi_product)

    def __init__(self, num_envs=None, device=None, mode=None, auto_reset=None, first_env=0, force_generic=False, **kwargs):
        info, kw = self._preset(**kwargs)
        super().__init__(info, num_envs=num_envs, device=device, mode=mode, auto_reset=auto_reset, first_env=first_env,
                         force_generic=force_generic, **kw)


class SupplyChainMultiProduct_IncreasingCosts(SupplyChainMultiProduct):
    _preset = staticmethod(presets.multi_product_increasing_costs)


class SupplyChainMultiProduct_DemConfigByProd(SupplyChainMultiProduct):
    _preset = staticmethod(presets.multi_product_dem_by_prod)


class SupplyChainMultiProduct_DemConfigByProd_IncCosts(SupplyChainMultiProduct):
    _preset = staticmethod(presets.multi_product_dem_by_prod_inc_costs)


class SupplyChainMainDemo(SupplyChainEnv):
    """supplychain-v0 with the reference's own demo chain (supplychain_env.py:842-875) as default nodes_info."""

    def __init__(self, nodes_info=None, num_envs=None, device=None, mode=None, auto_reset=None, first_env=0,
                 force_generic=False, **kwargs):
        if nodes_info is None:
            nodes_info, kw = presets.main_demo_chain()
            kw.update(kwargs)
        else:
            kw = kwargs
        super().__init__(nodes_info, num_envs=num_envs, device=device, mode=mode, auto_reset=auto_reset, first_env=first_env,
                         force_generic=force_generic, **kw)


# --------------------------------------------------------------------------------------------
class _BeerGameBase(_ScgEnvBase):
    def _setup(self, version, levels, weeks, customer_demand, shipment_delays, initial_inventory, inv_cost, backlog_cost,
               max_stock, max_order, penalty, initial_shipment, initial_orders, seed, num_envs, device, auto_reset, first_env):
        torch = _torch()
        s = _lib.ScgBeerGameSpec()
        s.version, s.levels, s.weeks = version, levels, weeks
        s.inv_cost, s.backlog_cost, s.max_stock, s.exceeded_capacity_penalty = inv_cost, backlog_cost, max_stock, penalty
        s.initial_shipment, s.initial_orders, s.max_order = initial_shipment, initial_orders, max_order
        keep = []

        def iptr(v):
            a = np.ascontiguousarray(np.asarray(v, dtype=np.int32))
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int32))
        s.initial_inventory = iptr(initial_inventory)
        if isinstance(customer_demand, tuple):
            s.demand_lo, s.demand_hi = int(customer_demand[0]), int(customer_demand[1])
            s.customer_demand = iptr([0] * weeks)
        else:
            s.customer_demand = iptr(list(customer_demand)[:weeks])
        if isinstance(shipment_delays, tuple):
            s.delay_lo, s.delay_hi = int(shipment_delays[0]), int(shipment_delays[1])
            s.shipment_delays = iptr([2] * weeks)
        else:
            s.shipment_delays = iptr(list(shipment_delays)[:weeks])
        self.levels, self.max_weeks = levels, weeks
        self._io_dtype = torch.int32
        if auto_reset is None:
            auto_reset = num_envs is not None
        self._init_handle(_lib.load().scg_bg_create, s, num_envs, device, "fast_f32", auto_reset, seed, first_env)

    def inject(self, demand=None, delays=None):
        """demand / delays: int [T] (weeks 1..T) for every env, or [N, T]."""
        torch = _torch()

        def prep(x):
            if x is None:
                return None
            x = torch.as_tensor(np.asarray(x)).to(torch.int32)
            if x.dim() == 1:
                x = x.unsqueeze(0).expand(self.num_envs, -1)
            if tuple(x.shape) != (self.num_envs, self.max_weeks):
                raise ValueError("inject: need shape (%d,) or (%d, %d), got %s"
                                 % (self.max_weeks, self.num_envs, self.max_weeks, tuple(x.shape)))
            return x.t().contiguous().to(self.device)
        d, dl = prep(demand), prep(delays)
        if dl is not None and (int(dl.max()) > self.scg_info.ring_depth or int(dl.min()) < 0):
            raise ValueError("injected delays must lie in [0, %d] (the in-transit ring depth of this env), got [%d, %d]"
                             % (self.scg_info.ring_depth, int(dl.min()), int(dl.max())))
        self._inj = (d, dl)
        torch.cuda.current_stream(self.device).synchronize()    # the *_host calls run on the handle's own stream
        _lib.check(self._lib.scg_inject(self._h, self._ptr(d), self._ptr(dl)))

    def _to_single_obs(self, obs):
        return obs[0].cpu().numpy().astype(np.int64)

    def _to_single_reward(self, rew):
        return int(rew[0].item())

    def _np_to_obs(self):
        return self._np_obs.astype(np.int64)

    def _np_to_reward(self):
        return int(self._np_rew[0])


class BeerGameEnv(_BeerGameBase):
    """beergame-v0 (beergame_env.py:6-181): one dict of settings; no spaces (the reference defines none)."""

    def __init__(self, env_init_info={}, num_envs=None, device=None, auto_reset=None, first_env=0, seed=None):
        g = env_init_info.get
        levels = g("levels", 4)
        demand = list(g("customer_demand", [4] * 4 + [8] * 31))
        weeks = len(demand)
        self.action_space = None
        self.observation_space = None
        self._setup(0, levels, weeks, demand, list(g("shipment_delays", [2] * weeks)),
                    list(np.asarray(g("initial_inventory", [12] * levels), dtype=int)), g("inv_cost", 1), g("backlog_cost", 2),
                    0, 0, 0, g("initial_shipment_value", 4), g("initial_orders_value", 4), seed, num_envs, device,
                    auto_reset, first_env)

    def seed(self, seed=None):
        _lib.check(self._lib.scg_seed(self._h, int(seed if seed is not None else 0)))


class BeerGameEnv2(_BeerGameBase):
    """beergame-v2 (beergame2_env.py:5-211)."""

    def __init__(self, max_stock=100, max_order=30, weeks=35, levels=4, customer_demand=[4] * 4 + [8] * 31,
                 initial_inventory=[12, 12, 12, 12], inv_cost=1, backlog_cost=2, exceeded_capacity_penalty=100,
                 shipment_delays=2, initial_shipment=4, initial_orders=4, seed=None,
                 num_envs=None, device=None, auto_reset=None, first_env=0):
        self.action_space = MultiDiscrete(levels * [max_order])                         # :27
        self.observation_space = MultiDiscrete(levels * [2 * max_stock])                # :28
        if isinstance(customer_demand, list) and len(customer_demand) == 2:             # :41
            customer_demand = tuple(customer_demand)
        if isinstance(shipment_delays, int):                                            # :49-50
            shipment_delays = [shipment_delays] * weeks
        elif isinstance(shipment_delays, list) and len(shipment_delays) == 2:           # :51
            shipment_delays = tuple(shipment_delays)
        self._setup(2, levels, weeks, customer_demand, shipment_delays, initial_inventory, inv_cost, backlog_cost,
                    max_stock, max_order, exceeded_capacity_penalty, initial_shipment, initial_orders, seed, num_envs,
                    device, auto_reset, first_env)
