"""CPU oracle for beergame-v0 / beergame-v2 -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates /root/reference/gym_supplychain/envs/beergame_env.py:66-156 (v0, action =
delta on the incoming order) and beergame2_env.py:79-192 (v2, action = the order,
capacity penalty, optional randint demand/delays).  All quantities are Python ints.

The reference has no tests for the beer game ("parity unpinned" by the reference's
own suite); this oracle is pinned against golden trajectories generated from the
unmodified reference (tests/golden/beergame_*.npz, oracle/make_golden.py) and against
the known answers derived in SURVEY.md Appendix E.

Unlike the reference (a table of shipments indexed by absolute week) the in-transit
stock is a dict keyed by arrival week; the arithmetic and its order are the same.
"""
import numpy as np

__all__ = ["BeerGameOracle"]


class BeerGameOracle(object):
    def __init__(self, version=0, levels=4, customer_demand=None, initial_inventory=None,
                 shipment_delays=None, initial_shipment=4, initial_orders=4,
                 inv_cost=1, backlog_cost=2, max_stock=100, exceeded_capacity_penalty=100,
                 weeks=35, seed=None):
        """version 0: beergame_env.py:11-56; version 2: beergame2_env.py:10-74.

        customer_demand: list of per-week demands, or (lo, hi) for v2's randint(lo, hi).
        shipment_delays: list of per-week delays (week 1..T) or an int (v2) or (lo, hi) (v2).
        The delay in force at week 0 (initial pipeline depth) is always 2 (beergame_env.py:39,
        beergame2_env.py:50,55,92).
        """
        self.version = version
        self.L = levels
        self.inv_cost, self.backlog_cost = inv_cost, backlog_cost
        self.max_stock, self.penalty = max_stock, exceeded_capacity_penalty
        self.init_ship, self.init_orders = initial_shipment, initial_orders
        self.demand_range = None
        self.delay_range = None
        if customer_demand is None:
            customer_demand = [4] * 4 + [8] * 31
        if version == 2 and (isinstance(customer_demand, tuple) or
                             (isinstance(customer_demand, list) and len(customer_demand) == 2)):
            self.demand_range = customer_demand
            self.T = weeks
            self.demand = None
        else:
            self.demand = [int(x) for x in customer_demand]
            self.T = len(self.demand) if version == 0 else weeks
        if shipment_delays is None:
            shipment_delays = 2 if version == 2 else [2] * self.T
        if isinstance(shipment_delays, int):
            self.delays = [2] + [shipment_delays] * self.T
        elif version == 2 and (isinstance(shipment_delays, tuple) or len(shipment_delays) == 2):
            self.delay_range = shipment_delays
            self.delays = None
        else:
            self.delays = [2] + [int(x) for x in shipment_delays]
        self.inventory0 = [12] * levels if initial_inventory is None else [int(x) for x in initial_inventory]
        self.rng = np.random.RandomState(seed)

    def seed(self, seed=None):
        self.rng = np.random.RandomState(seed)                 # beergame2_env.py:210-211

    def reset(self, demand=None, delays=None):
        """`demand` [T] / `delays` [T] (weeks 1..T) may be injected; v2 draws them otherwise."""
        if demand is not None:
            self.demand = [int(x) for x in demand]
        elif self.demand_range is not None:                     # beergame2_env.py:86-87
            self.demand = [int(x) for x in self.rng.randint(self.demand_range[0], self.demand_range[1], self.T)]
        if delays is not None:
            self.delays = [2] + [int(x) for x in delays]
        elif self.delay_range is not None:                      # beergame2_env.py:89-92
            self.delays = [2] + [int(x) for x in self.rng.randint(self.delay_range[0], self.delay_range[1], self.T)]
        self.week = 0
        self.inventory = list(self.inventory0)
        self.backlog = [0] * self.L
        self.placed = [self.init_orders] * self.L
        self.incoming = [self.init_orders] * self.L
        self.transit = {}
        for w in range(1, 1 + self.delays[0]):                   # beergame_env.py:52
            self.transit[w] = [self.init_ship] * self.L
        return self._obs()

    def _obs(self):
        off = self.max_stock if self.version == 2 else 0         # beergame2_env.py:111-112
        return np.array([off + self.inventory[i] - self.backlog[i] for i in range(self.L)], dtype=np.int64)

    def step(self, action):
        L = self.L
        action = [int(x) for x in action]
        self.week += 1
        w = self.week
        arriving = self.transit.pop(w, [0] * L)                   # beergame_env.py:72
        for i in range(L):
            self.inventory[i] += arriving[i]
        self.incoming = [self.demand[w - 1]] + self.placed[:-1]  # :79-81
        to_fill = [self.incoming[i] + self.backlog[i] for i in range(L)]
        deliver = [min(self.inventory[i], to_fill[i]) for i in range(L)]
        delay = self.delays[w]
        if delay == 0:                                            # :93-96 (before the decrement below)
            for i in range(L - 1):
                self.inventory[i] += deliver[i + 1]
        else:
            slot = self.transit.setdefault(w + delay, [0] * L)
            for i in range(L - 1):
                slot[i] += deliver[i + 1]
        for i in range(L):                                        # :101-103
            self.inventory[i] -= deliver[i]
            self.backlog[i] = to_fill[i] - deliver[i]
        if delay == 0:                                            # :111-114, the brewery's own order
            self.inventory[L - 1] += self.placed[L - 1]
        else:
            self.transit.setdefault(w + delay, [0] * L)[L - 1] += self.placed[L - 1]
        if self.version == 0:
            self.placed = [self.incoming[i] + action[i] for i in range(L)]    # beergame_env.py:121
        else:
            self.placed = list(action)                                        # beergame2_env.py:168
        reward = -sum(self.inv_cost * self.inventory[i] + self.backlog_cost * self.backlog[i] for i in range(L))
        if self.version == 2:                                                 # beergame2_env.py:179-180
            reward -= sum(self.penalty * (max(self.inventory[i] - self.max_stock, 0) +
                                          max(self.backlog[i] - self.max_stock, 0)) for i in range(L))
        return self._obs(), reward, w == self.T, {}
