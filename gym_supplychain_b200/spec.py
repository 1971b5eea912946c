"""nodes_info + kwargs  ->  flat arrays for ScgSupplyChainSpec (include/scg.h).

Host-side mirror of what the reference constructor derives
(/root/reference/gym_supplychain/envs/supplychain_env.py:482-626 and SC_Node.__init__ :120-206):
same kwargs, same validation errors (asserts / ValueError), same action / observation sizes.
"""
import ctypes as C

import numpy as np

from . import _lib


def _int_or_list(value, P):
    """supplychain_env.py:178-191: an int is replicated per product; a list must have P entries
    (an empty list means zeros); anything else (e.g. a float) is a ValueError."""
    if type(value) is list:
        if len(value) == 0:
            return [0] * P
        assert len(value) == P
        return list(value)
    if type(value) is int:
        return [value] * P
    raise ValueError(f"Invalid param: '{value}' should be an int or a list with one value per product")


class FlatSpec(object):
    """Flat description of one SupplyChainEnv; `.c_struct()` gives the ctypes view for scg_sc_create."""

    def __init__(self, nodes_info, num_products=1, unmet_demand_cost=1000, exceeded_stock_capacity_cost=1000,
                 exceeded_process_capacity_cost=1000, exceeded_ship_capacity_cost=1000,
                 demand_config_by_product=False, demand_range=(10, 20), demand_std=None, demand_sen_peaks=None,
                 avg_demand_range=None, processing_ratio=3, stochastic_leadtimes=False, avg_leadtime=2,
                 max_leadtime=2, total_time_steps=360, seed=None, build_info=False, demand_perturb_norm=False):
        P = int(num_products)
        names = list(nodes_info.keys())
        index = {name: i for i, name in enumerate(names)}
        N = len(names)
        self.node_names = names
        self.N, self.P, self.T = N, P, int(total_time_steps)
        self.avg_leadtime, self.max_leadtime = int(avg_leadtime), int(max_leadtime)
        self.stochastic_leadtimes = bool(stochastic_leadtimes)
        self.costs = (float(unmet_demand_cost), float(exceeded_stock_capacity_cost),
                      float(exceeded_process_capacity_cost), float(exceeded_ship_capacity_cost))

        f8 = lambda *shape: np.zeros(shape, dtype=np.float64)
        self.last_level = np.zeros(N, dtype=np.int32)
        self.processing_capacity = f8(N)
        self.initial_stock, self.stock_capacity, self.stock_cost = f8(N, P), f8(N, P), f8(N, P)
        self.supply_capacity, self.supply_cost = f8(N, P), f8(N, P)
        self.processing_ratio, self.processing_cost = f8(N, P), f8(N, P)
        edge_begin, edge_dest, edge_cap, edge_cost = [0], [], [], []
        init_begin, init_time, init_amount = [0], [], []

        for n, name in enumerate(names):
            info = nodes_info[name]
            pc = info.get("processing_cost", 0)
            zero_pc = (type(pc) is int and pc == 0) or (type(pc) is list and sum(pc) == 0)      # :518-522
            self.processing_ratio[n] = _int_or_list(0 if zero_pc else processing_ratio, P)
            self.processing_cost[n] = _int_or_list(pc, P)
            self.processing_capacity[n] = info.get("processing_capacity", 0)
            self.last_level[n] = 1 if info.get("last_level", False) else 0
            self.initial_stock[n] = _int_or_list(info.get("initial_stock", 0), P)
            self.stock_capacity[n] = _int_or_list(info.get("stock_capacity", float("inf")), P)   # float -> ValueError
            self.stock_cost[n] = _int_or_list(info.get("stock_cost", 0), P)
            self.supply_capacity[n] = _int_or_list(info.get("supply_capacity", 0), P)
            self.supply_cost[n] = _int_or_list(info.get("supply_cost", 0), P)
            if "destinations" in info:                                                            # :551-554
                dests = info["destinations"]
                caps, dcost = info["ship_capacity"], info["dest_costs"]
                for i, d in enumerate(dests):
                    edge_dest.append(index[d])
                    edge_cap.append(float(caps[i]))
                    edge_cost.append([float(dcost[p][i]) for p in range(P)])
            edge_begin.append(len(edge_dest))
            for p in range(P):                                                                    # :405-412
                for src in (info.get("initial_supply", None), info.get("initial_shipments", None)):
                    if src:
                        for i, amount in enumerate(src[p]):
                            init_time.append(i + 1)
                            init_amount.append(float(amount))
                init_begin.append(len(init_time))

        self.edge_begin = np.asarray(edge_begin, dtype=np.int32)
        self.edge_dest = np.asarray(edge_dest, dtype=np.int32)
        self.edge_ship_capacity = np.asarray(edge_cap, dtype=np.float64)
        self.edge_dest_cost = np.asarray(edge_cost, dtype=np.float64).reshape(len(edge_dest), P)
        self.init_begin = np.asarray(init_begin, dtype=np.int32)
        self.init_time = np.asarray(init_time, dtype=np.int32)
        self.init_amount = np.asarray(init_amount, dtype=np.float64)
        self.E = len(edge_dest)

        # ---- demand model (:566-595, demands_generator.py:20-31) ----
        self.demand_config_by_product = bool(demand_config_by_product)
        kind, lo, hi = np.zeros(P, np.int32), f8(P), f8(P)
        std, peaks, mn, mx = f8(P), f8(P), f8(P), f8(P)
        for p in range(P):
            if not demand_config_by_product:
                rng_p, std_p, peaks_p, avg_p, norm_p = demand_range, demand_std, demand_sen_peaks, avg_demand_range, demand_perturb_norm
            else:
                rng_p, std_p, peaks_p = demand_range[p], demand_std[p], demand_sen_peaks[p]
                avg_p, norm_p = avg_demand_range[p], demand_perturb_norm[p]
            if isinstance(std_p, (list, tuple)):          # supplychain_multiproduct_env.py:243,250 wraps it in a list
                std_p = std_p[0]
            assert rng_p[0] != rng_p[1]                   # fixed demand is rejected (:591-595)
            lo[p], hi[p] = rng_p[0], rng_p[1]
            if peaks_p is None:
                kind[p] = _lib.DEMAND_UNIFORM if std_p is None else _lib.DEMAND_NORMAL
                std[p] = 0 if std_p is None else std_p
            else:
                kind[p] = _lib.DEMAND_SIN_NORMAL if norm_p else _lib.DEMAND_SIN_UNIFORM
                std[p] = 0 if std_p is None else std_p
                peaks[p] = peaks_p
                mn[p], mx[p] = avg_p[0], avg_p[1]
        self.demand_kind, self.demand_lo, self.demand_hi = kind, lo, hi
        self.demand_std, self.demand_peaks, self.demand_minavg, self.demand_maxavg = std, peaks, mn, mx

        # ---- sizes (:597-621) ----
        D = np.diff(self.edge_begin)
        n_sup = (self.supply_capacity > 0).sum(axis=1)
        n_shipprod = np.where(D > 0, (self.stock_capacity > 0).sum(axis=1), 0)
        self.n_retailers = int(self.last_level.sum())
        self.n_actions = int((n_sup + n_shipprod * D).sum())
        self.n_obs = self.n_retailers * P + N * P * (1 + self.avg_leadtime) + 1
        self.n_leadtimes = int((np.where(n_sup > 0, P, 0) + D).sum()) if self.stochastic_leadtimes else 0
        self._keep = None

    def c_struct(self):
        s = _lib.ScgSupplyChainSpec()
        s.n_nodes, s.n_products, s.n_edges, s.total_time_steps = self.N, self.P, self.E, self.T
        s.avg_leadtime, s.max_leadtime = self.avg_leadtime, self.max_leadtime
        s.stochastic_leadtimes = 1 if self.stochastic_leadtimes else 0
        (s.unmet_demand_cost, s.exceeded_stock_capacity_cost, s.exceeded_process_capacity_cost,
         s.exceeded_ship_capacity_cost) = self.costs
        keep = []

        def ptr(arr, ctype):
            a = np.ascontiguousarray(arr)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(ctype)) if a.size else C.cast(None, C.POINTER(ctype))
        for name in ("last_level", "edge_begin", "edge_dest", "init_begin", "init_time", "demand_kind"):
            setattr(s, name, ptr(getattr(self, name).astype(np.int32), C.c_int32))
        for name in ("processing_capacity", "initial_stock", "stock_capacity", "stock_cost", "supply_capacity",
                     "supply_cost", "processing_ratio", "processing_cost", "edge_ship_capacity", "edge_dest_cost",
                     "init_amount", "demand_lo", "demand_hi", "demand_std", "demand_peaks", "demand_minavg",
                     "demand_maxavg"):
            setattr(s, name, ptr(getattr(self, name).astype(np.float64), C.c_double))
        self._keep = keep
        return s

    def negative_ship_capacity_reachable(self):
        """Can the reference's ship-capacity quirk (supplychain_env.py:316-328) leave a factory edge with a
        NEGATIVE capacity for the next product?  It needs a multi-product factory whose processing ratio exceeds 1 and an
        edge it can saturate: the most it can send is min(stock_capacity, processing_capacity) / ratio (:298-310).  From
        then on whether an amount counts as `> 0` decides a whole clip, so only `exact_f64` follows the reference step
        for step; the tolerance modes may diverge on such specs.  None of the reference's presets can
        reach it.  Returns the list of (node name, product, edge destination name) that can."""
        hits = []
        if self.P < 2:
            return hits
        for n in range(self.N):
            pcap = self.processing_capacity[n]
            if not pcap > 0:
                continue
            for e in range(self.edge_begin[n], self.edge_begin[n + 1]):
                for p in range(self.P):
                    r = self.processing_ratio[n, p]
                    if r > 1 and min(self.stock_capacity[n, p], pcap) / r > self.edge_ship_capacity[e]:
                        hits.append((self.node_names[n], p, self.node_names[self.edge_dest[e]]))
        return hits

    def precompile(self, mode="fast_f32", in_kernel_actions=False):
        """Make sure a compile-time kernel exists for this chain before the first rollout (needs no GPU): chains without a
        pre-built instantiation get one compiled with NVRTC into the on-disk cache (scg_sc_precompile).  Returns False
        when the spec can only run on the runtime-topology kernel (not a layered full-bipartite chain)."""
        lib = _lib.load()
        rc = lib.scg_sc_precompile(C.byref(self.c_struct()), _lib.MODES[mode], 1 if in_kernel_actions else 0)
        if rc == -1:                        # SCG_EINVAL: no compile-time kernel for this spec
            return False
        _lib.check(rc)
        return True
