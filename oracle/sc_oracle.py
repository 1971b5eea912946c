"""CPU oracle for the SupplyChainEnv family -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module.  The product package (gym_supplychain_b200) never does: its hot path
is the CUDA extension and it fails loudly when that is missing.

This is a restatement (not a copy) of the reference's reset()/step() dynamics:

  action scaling ............ /root/reference/gym_supplychain/envs/supplychain_env.py:49-96
  node step ................. supplychain_env.py:208-396
  node reset ................ supplychain_env.py:402-412
  node observation .......... supplychain_env.py:428-463
  env construction .......... supplychain_env.py:499-626
  env reset / step / obs .... supplychain_env.py:630-682, 703-748, 762-791
  demand generators ......... demands_generator.py:3-89
  Poisson lead times ........ supplychain_env.py:664-672

It is deliberately scalar Python on IEEE doubles, with the stdlib `heapq` for the
in-transit queues, because two things are only reproducible that way:
(1) the order of every floating-point addition (arrivals are summed in heappop
order, costs in node/product order), and (2) the reference observation walks the
*heap array* by index, so the exact CPython heap layout matters.

Parity pinning: tests/test_oracle_vs_golden.py checks this module against the
golden vectors in tests/golden/ (produced by oracle/make_golden.py from the
UNMODIFIED reference run in the build container), and
tests/test_oracle_vs_reference.py checks it live against /root/reference when
that tree is present.

Besides the reference-layout observation it also computes the *canonical*
observation (true per-arrival-time sums) that the fast GPU mode emits.
"""
import heapq
import math

import numpy as np

__all__ = ["SupplyChainOracle", "generate_demand_table", "generate_leadtime_table", "INFO_KEYS"]

# keys of info['sc_episode']['costs'|'units'] in the reference's order (supplychain_env.py:416-417, :685-686)
INFO_KEYS = ("stock", "stock_pen", "supply", "process", "process_pen", "ship", "ship_pen", "unmet_dem")


# --------------------------------------------------------------------------
# stochastic inputs (demands_generator.py:3-89, supplychain_env.py:664-672)
# --------------------------------------------------------------------------
def generate_demand_table(rng, shape, horizon, lo, hi, std=None, sen_peaks=None,
                          minavg=None, maxavg=None, perturb_norm=True):
    """One whole-episode demand table, consuming `rng` exactly like the reference.

    demands_generator.py:20-31 selects the family: sinusoid when `sen_peaks` is
    given, else clipped/rounded normal when `std` is given, else uniform ints.
    """
    if sen_peaks is None and std is None:
        return rng.randint(low=lo, high=hi + 1, size=shape)            # :33-36
    if sen_peaks is None:
        x = rng.normal((hi + lo) / 2, std, size=shape)                  # :41
        return np.rint(np.clip(x, lo, hi)).astype(int)                  # :44-47
    std = 0 if std is None else std
    half = (maxavg - minavg) / 2                                        # :72
    w = sen_peaks * 2 * np.pi / horizon                                 # :73
    if perturb_norm:
        noise = rng.normal(0, std, size=shape)                          # :75
    else:
        noise = rng.randint(low=-3 * std, high=3 * std + 1, size=shape)  # :77
    out = np.zeros(shape)
    # the reference only fills [period, d] for 2-D shapes and broadcasts the
    # trailing product axis for 3-D ones (data[period, d] = <P-vector>), :79-84
    for period in range(shape[0]):
        base = minavg + half * (1 + np.sin(w * period))
        for d in range(shape[1]):
            out[period, d] = np.clip(base + noise[period, d], lo, hi)
    return np.rint(out).astype(int)


def generate_leadtime_table(rng, avg_leadtime, max_leadtime, T, K):
    """`clip(1 + Poisson(avg-1), 1, max)` of shape (T, K) (supplychain_env.py:670-672)."""
    return np.clip(1 + rng.poisson(lam=avg_leadtime - 1, size=(T, K)), 1, max_leadtime)


def _lsum(values):
    """Plain left-to-right IEEE sum starting from int 0.

    NOT the builtin sum(): CPython >= 3.12 uses compensated (Neumaier) summation when
    every item is an exact Python float.  In the reference the items are np.float64
    (a float subclass), which takes sum()'s generic `+` path on every Python version,
    i.e. exactly this loop (supplychain_env.py:331,352)."""
    acc = 0
    for v in values:
        acc = acc + v
    return acc


def _per_product(value, P):
    """int -> [int]*P; list -> checked list ([] -> zeros) (supplychain_env.py:178-191)."""
    if type(value) is list:
        if len(value) == 0:
            return [0] * P
        assert len(value) == P
        return list(value)
    if type(value) is int:
        return [value] * P
    raise ValueError("Invalid param: %r should be an int or a list with one value per product" % (value,))


class _Node(object):
    __slots__ = ("name", "P", "sup_cap", "sup_cost", "is_supplier", "proc_cap", "proc_ratio",
                 "proc_cost", "last", "stock0", "init_supply", "init_ship", "cap", "hold",
                 "dests", "ship_cap", "dest_cost", "can_ship", "max_ship", "stock", "q")


class SupplyChainOracle(object):
    """Single-env CPU oracle with the reference's constructor kwargs (supplychain_env.py:482-489)."""

    def __init__(self, nodes_info, num_products=1, unmet_demand_cost=1000,
                 exceeded_stock_capacity_cost=1000, exceeded_process_capacity_cost=1000,
                 exceeded_ship_capacity_cost=1000, demand_config_by_product=False,
                 demand_range=(10, 20), demand_std=None, demand_sen_peaks=None, avg_demand_range=None,
                 processing_ratio=3, stochastic_leadtimes=False, avg_leadtime=2, max_leadtime=2,
                 total_time_steps=360, seed=None, build_info=False, demand_perturb_norm=False):
        P = self.P = num_products
        self.T = total_time_steps
        self.avg_lt, self.max_lt, self.stoch_lt = avg_leadtime, max_leadtime, stochastic_leadtimes
        self.c_unmet = unmet_demand_cost
        self.c_xstock = exceeded_stock_capacity_cost
        self.c_xproc = exceeded_process_capacity_cost
        self.c_xship = exceeded_ship_capacity_cost
        self.by_product = demand_config_by_product
        self.d_range, self.d_std, self.d_peaks = demand_range, demand_std, demand_sen_peaks
        self.d_avg, self.d_norm = avg_demand_range, demand_perturb_norm
        self.build_info = build_info
        self.rng = np.random.RandomState(seed)

        # ---- nodes, in dict insertion order (= act order, :503,544) ----
        self.nodes, by_name = [], {}
        for name, info in nodes_info.items():
            nd = _Node()
            nd.name, nd.P = name, P
            nd.sup_cap = _per_product(info.get("supply_capacity", 0), P)
            nd.sup_cost = _per_product(info.get("supply_cost", 0), P)
            nd.is_supplier = max(nd.sup_cap) > 0                              # :137
            pc = info.get("processing_cost", 0)
            zero_pc = (type(pc) is int and pc == 0) or (type(pc) is list and sum(pc) == 0)   # :518-522
            nd.proc_ratio = _per_product(0 if zero_pc else processing_ratio, P)
            nd.proc_cost = _per_product(pc, P)
            nd.proc_cap = info.get("processing_capacity", 0)
            nd.last = info.get("last_level", False)
            nd.stock0 = _per_product(info.get("initial_stock", 0), P)
            nd.init_supply = info.get("initial_supply", None)
            nd.init_ship = info.get("initial_shipments", None)
            nd.cap = _per_product(info.get("stock_capacity", float("inf")), P)
            nd.hold = _per_product(info.get("stock_cost", 0), P)
            nd.dests, nd.ship_cap, nd.dest_cost = None, None, None
            nd.can_ship = [False] * P
            nd.max_ship = list(nd.sup_cap) if nd.is_supplier else [0] * P    # :147,152
            self.nodes.append(nd)
            by_name[name] = nd
        for name, info in nodes_info.items():                                 # :548-554
            if "destinations" in info:
                nd = by_name[name]
                nd.dests = [by_name[d] for d in info["destinations"]]
                nd.ship_cap = info["ship_capacity"]
                nd.dest_cost = info["dest_costs"]
                nd.can_ship = [c > 0 for c in nd.cap]                         # :196-201
                for i, d in enumerate(nd.dests):                              # :204-206
                    for p in range(P):
                        d.max_ship[p] += nd.ship_cap[i]
        self.retailers = [nd for nd in self.nodes if nd.last]

        # ---- demand config (:575-595) ----
        if not self.by_product:
            assert demand_range[0] != demand_range[1]
            self.d_span = demand_range[1] - demand_range[0]
        else:
            for p in range(P):
                assert demand_range[p][0] != demand_range[p][1]
            self.d_span = [demand_range[p][1] - demand_range[p][0] for p in range(P)]

        # ---- sizes (:597-621) ----
        self.n_leadtimes = sum((P if self._n_supply(nd) > 0 else 0) + (len(nd.dests) if nd.dests else 0)
                               for nd in self.nodes)
        self.n_actions = sum(self._n_actions(nd) for nd in self.nodes)
        self.n_obs = len(self.retailers) * P + len(self.nodes) * P * (1 + avg_leadtime) + 1
        self.t = 0
        self.demand = None
        self.leadtimes = None

    # ----- per-node action counts (:139-142, :196-199, :419-420) -----
    @staticmethod
    def _n_supply(nd):
        return sum(1 for c in nd.sup_cap if c > 0) if nd.is_supplier else 0

    @staticmethod
    def _n_ship(nd):
        return sum(1 for s in nd.can_ship if s) * len(nd.dests) if nd.dests else 0

    def _n_actions(self, nd):
        return self._n_supply(nd) + self._n_ship(nd)

    def seed(self, seed=None):
        self.rng = np.random.RandomState(seed)                                # :811-812

    # ------------------------------------------------------------------
    def draw_inputs(self):
        """Draw (demand, leadtimes) for one episode from self.rng like reset() does (:641-672)."""
        P, R, T = self.P, len(self.retailers), self.T
        if not self.by_product:
            mn, mx = (self.d_avg if self.d_avg else (None, None))
            demand = generate_demand_table(self.rng, (T + 1, R, P), T, self.d_range[0], self.d_range[1],
                                           std=self.d_std, sen_peaks=self.d_peaks, minavg=mn, maxavg=mx,
                                           perturb_norm=self.d_norm)
        else:
            per_prod = []
            for p in range(P):
                mn, mx = (self.d_avg[p] if self.d_avg[p] else (None, None))
                per_prod.append(generate_demand_table(self.rng, (T + 1, R), T, self.d_range[p][0],
                                                      self.d_range[p][1], std=self.d_std[p],
                                                      sen_peaks=self.d_peaks[p], minavg=mn, maxavg=mx,
                                                      perturb_norm=self.d_norm[p]))
            demand = np.stack(per_prod, axis=-1)          # canonical [T+1, R, P]
        lts = None
        if self.stoch_lt:
            lts = generate_leadtime_table(self.rng, self.avg_lt, self.max_lt, T, self.n_leadtimes)
        return demand, lts

    def reset(self, demand=None, leadtimes=None):
        """Reset; `demand` [T+1,R,P] / `leadtimes` [T,K] are injected when given, else drawn."""
        if demand is None:
            demand, lts = self.draw_inputs()
            if leadtimes is None:
                leadtimes = lts
        self.demand = np.asarray(demand)
        self.leadtimes = None if leadtimes is None else np.asarray(leadtimes)
        if self.stoch_lt:
            assert self.leadtimes is not None
        for nd in self.nodes:                                                 # :402-412
            nd.stock = [float(x) for x in nd.stock0]
            nd.q = [[] for _ in range(self.P)]
            for src in (nd.init_supply, nd.init_ship):
                if src:
                    for p in range(self.P):
                        for i, amount in enumerate(src[p]):
                            heapq.heappush(nd.q[p], (i + 1, amount))
        self.t = 0
        self.episode_return = 0.0
        # cumulative statistics of the episode (:684-695); filled only when build_info
        self.ep_costs = {k: [0] * self.P for k in INFO_KEYS}
        self.ep_units = {k: [0] * self.P for k in INFO_KEYS}
        self.ep_rewards = 0
        return self.observation()

    # ------------------------------------------------------------------
    def _ship_cuts(self, values, limit):
        """SHIP scaling: sorted cut points of `limit` (:58-96); limit > 0 guaranteed by caller."""
        out = [0] * len(values)
        left = limit
        prev = 0
        for v, i in sorted((values[i], i) for i in range(len(values))):
            amount = (v - prev) * limit
            if amount > left:
                amount = left
            out[i] = amount
            left -= amount
            prev = v
        return out

    def _node_step(self, nd, acts, lts, t, demand):
        """One node, one period (:208-396).  Returns the node's cost."""
        P = self.P
        cost = 0
        # per-node, per-period statistics (:214-218): every key starts at 0 for every product
        ec = {k: [0] * P for k in INFO_KEYS}
        eu = {k: [0] * P for k in INFO_KEYS}
        # 1. arrivals, popped in ascending (time, amount) (:220-228)
        for p in range(P):
            got = 0.0
            q = nd.q[p]
            while q and q[0][0] == t:
                got += heapq.heappop(q)[1]
            nd.stock[p] = nd.stock[p] + got
        # 2. over-capacity: pay and discard (:232-240)
        for p in range(P):
            if nd.stock[p] > nd.cap[p]:
                cost += self.c_xstock * (nd.stock[p] - nd.cap[p])
                ec["stock_pen"][p] = self.c_xstock * (nd.stock[p] - nd.cap[p])
                eu["stock_pen"][p] = nd.stock[p] - nd.cap[p]
                nd.stock[p] = nd.cap[p]
        ai = 0   # next action value
        li = 0   # next lead-time value
        # 3. supply (:244-256): the lead-time cursor only moves when something was supplied
        if nd.is_supplier and self._n_supply(nd) > 0:
            for p in range(P):
                if nd.sup_cap[p] > 0:
                    amount = acts[ai] * nd.sup_cap[p]
                    ai += 1
                    if amount > 0:
                        heapq.heappush(nd.q[p], (t + lts[li], amount))
                        li += 1
                    cost += amount * nd.sup_cost[p]
                    ec["supply"][p] = amount * nd.sup_cost[p]
                    eu["supply"][p] = amount
        if not nd.last:
            # 4. ship (:262-375)
            D = len(nd.dests)
            room = list(nd.ship_cap)           # per destination, shared by products only via the quirk
            proc_left = nd.proc_cap            # shared by products
            base = li
            for p in range(P):
                if not nd.can_ship[p]:
                    continue
                over_ship = 0
                over_proc = 0
                if nd.stock[p] > 0:
                    out = self._ship_cuts(acts[ai:ai + D], min(nd.cap[p], nd.stock[p]))
                    send = list(out)
                    if nd.proc_cap > 0:                                        # :298-310
                        for i in range(D):
                            if out[i] > 0:
                                if out[i] > proc_left:
                                    over_proc += out[i] - proc_left
                                    out[i] = proc_left
                                proc_left -= out[i]
                            send[i] = out[i] / nd.proc_ratio[p]
                    for i in range(D):                                         # :312-328
                        if send[i] > 0 and send[i] > room[i]:
                            over_ship += send[i] - room[i]
                            send[i] = room[i]
                            out[i] = send[i] * nd.proc_ratio[p] if nd.proc_cap > 0 else send[i]
                            room[i] -= out[i]          # only here, and in raw-material units (F4a)
                    leaving = _lsum(out)                                       # :331
                    nd.stock[p] -= leaving
                    if nd.proc_cap > 0:
                        cost += leaving * nd.proc_cost[p]                      # :338
                        ec["process"][p] = leaving * nd.proc_cost[p]
                        eu["process"][p] = leaving
                    for i in range(D):                                         # :344-348
                        if send[i] > 0:
                            heapq.heappush(nd.dests[i].q[p], (t + lts[base + i], send[i]))
                    ship_cost = _lsum([send[i] * nd.dest_cost[p][i] for i in range(D)])   # :352-353
                    cost += ship_cost
                    ec["ship"][p] = ship_cost
                    eu["ship"][p] = _lsum(send)                                # :356
                cost += self.c_xproc * over_proc                               # :361
                cost += self.c_xship * over_ship                               # :366
                ec["process_pen"][p], eu["process_pen"][p] = self.c_xproc * over_proc, over_proc
                ec["ship_pen"][p], eu["ship_pen"][p] = self.c_xship * over_ship, over_ship
                ai += D
        else:
            # 5. retailers serve what they can; the rest is lost (:380-384)
            for p in range(P):
                served = min(nd.stock[p], demand[p])
                nd.stock[p] -= served
                cost += self.c_unmet * (demand[p] - served)
                ec["unmet_dem"][p] = self.c_unmet * (demand[p] - served)
                eu["unmet_dem"][p] = demand[p] - served
        # 6. holding (:390-391)
        for p in range(P):
            cost += nd.stock[p] * nd.hold[p]
            ec["stock"][p] = nd.stock[p] * nd.hold[p]
            eu["stock"][p] = nd.stock[p]
        if self.build_info:                                                    # :750-760, node by node
            for k in INFO_KEYS:
                for p in range(P):
                    self.ep_costs[k][p] += ec[k][p]
                    self.ep_units[k][p] += eu[k][p]
        return cost

    def step(self, action):
        """(obs, reward, done, info) like the reference (:703-748).  `action` in [-1, 1]."""
        a = (np.asarray(action, dtype=np.float64) + 1) / 2                    # :697-698 on float64 input
        a = [float(x) for x in a]
        self.t += 1
        t = self.t
        total = 0
        ai = li = ri = 0
        for nd in self.nodes:
            na = self._n_actions(nd)
            acts = a[ai:ai + na]
            ai += na
            if self.stoch_lt:                                                  # :719-722
                nl = self._n_supply(nd) + self._n_ship(nd) // self.P
                lts = [int(x) for x in self.leadtimes[t - 1, li:li + nl]]
                li += nl
            else:
                lts = [self.avg_lt] * na                                       # :724
            dem = None
            if nd.last:
                dem = [int(x) for x in self.demand[t - 1, ri]]                 # :728
                ri += 1
            total += self._node_step(nd, acts, lts, t, dem)
        reward = -total
        self.episode_return += reward
        info = {}
        if self.build_info:
            self.ep_rewards += reward                                          # :751
            info = {"sc_episode": {"rewards": self.ep_rewards, "costs": self.ep_costs, "units": self.ep_units}}
        return self.observation(), reward, t == self.T, info

    # ------------------------------------------------------------------
    def _obs_head(self):
        t = self.t
        head = []
        for r in range(len(self.retailers)):
            for p in range(self.P):
                if not self.by_product:
                    head.append((int(self.demand[t, r, p]) - self.d_range[0]) / self.d_span)
                else:
                    head.append((int(self.demand[t, r, p]) - self.d_range[p][0]) / self.d_span[p])
        return head

    def _finish(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.clip(x * 2 - 1, -1.0, 1.0)                                   # :700-701

    def observation(self):
        """Reference observation: walks each heap ARRAY in index order (:428-463, F3)."""
        t, L = self.t, self.avg_lt
        lo, hi = t + 1, t + L
        x = self._obs_head()
        for nd in self.nodes:
            x += [nd.stock[p] / nd.cap[p] for p in range(self.P)]
            for p in range(self.P):
                q = nd.q[p]
                if not q:
                    x += [0] * L
                    continue
                k = 0
                for when in range(lo, hi):
                    s = 0
                    while k < len(q) and q[k][0] == when:
                        s += q[k][1]
                        k += 1
                    x.append(s / nd.max_ship[p])
                s = 0
                while k < len(q):
                    s += q[k][1]
                    k += 1
                x.append(s / (nd.max_ship[p] * (self.max_lt - (hi - lo))))
        x.append((self.T - t) / self.T)
        return self._finish(x)

    def canonical_observation(self):
        """True per-arrival-time sums (ascending (time, amount) order inside a bucket).

        Equals observation() whenever every heap array is time-sorted (always at
        reset; every step of single-source constant-lead-time chains)."""
        t, L = self.t, self.avg_lt
        x = self._obs_head()
        for nd in self.nodes:
            x += [nd.stock[p] / nd.cap[p] for p in range(self.P)]
            for p in range(self.P):
                entries = sorted(nd.q[p])
                if not entries:
                    x += [0] * L
                    continue
                for when in range(t + 1, t + L):
                    s = 0
                    for (w, amount) in entries:
                        if w == when:
                            s += amount
                    x.append(s / nd.max_ship[p])
                s = 0
                for (w, amount) in entries:
                    if w >= t + L:
                        s += amount
                x.append(s / (nd.max_ship[p] * (self.max_lt - (L - 1))))
        x.append((self.T - t) / self.T)
        return self._finish(x)

    # ----- state snapshots used by the parity tests -----
    def stocks(self):
        return np.array([[nd.stock[p] for p in range(self.P)] for nd in self.nodes], dtype=np.float64)

    def pipeline_sums(self, depth):
        """[N, P, depth]: total in transit arriving at t+1 .. t+depth (ascending-sorted sums)."""
        out = np.zeros((len(self.nodes), self.P, depth))
        for n, nd in enumerate(self.nodes):
            for p in range(self.P):
                for (w, amount) in sorted(nd.q[p]):
                    k = w - self.t - 1
                    assert 0 <= k < depth, (w, self.t, depth)
                    out[n, p, k] += amount
        return out

    def heap_arrays(self, cap):
        """(len[N,P], time[N,P,cap], amount[N,P,cap]) -- the raw heap layout."""
        N = len(self.nodes)
        ln = np.zeros((N, self.P), dtype=np.int32)
        tm = np.zeros((N, self.P, cap), dtype=np.int32)
        am = np.zeros((N, self.P, cap), dtype=np.float64)
        for n, nd in enumerate(self.nodes):
            for p in range(self.P):
                q = nd.q[p]
                assert len(q) <= cap
                ln[n, p] = len(q)
                for k, (w, amount) in enumerate(q):
                    tm[n, p, k] = w
                    am[n, p, k] = amount
        return ln, tm, am


def _selftest():  # pragma: no cover - tiny smoke run
    info = {"S": {"initial_stock": 10, "stock_capacity": 100, "stock_cost": 1, "supply_capacity": 20,
                  "supply_cost": 5, "destinations": ["R"], "dest_costs": [[2]], "ship_capacity": [50]},
            "R": {"initial_stock": 10, "stock_capacity": 100, "stock_cost": 1, "last_level": True}}
    env = SupplyChainOracle(info, total_time_steps=5, seed=0)
    env.reset()
    done = False
    while not done:
        _, r, done, _ = env.step(np.zeros(env.n_actions))
        assert math.isfinite(r)


if __name__ == "__main__":
    _selftest()
    print("ok")
