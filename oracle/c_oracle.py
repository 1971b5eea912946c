"""ctypes front end of oracle/sc_oracle.c (the plain-C CPU oracle) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__ and bench.py's cpu legs import this.  The chain description is parsed by the Python oracle
(oracle/sc_oracle.py, itself a restatement of supplychain_env.py:499-626) and handed to C as flat arrays; stochastic inputs
are injected tables.  `build()` compiles the C file with gcc into oracle/_build/ (git-ignored).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .sc_oracle import SupplyChainOracle

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sc_oracle.c")
LIB = os.path.join(HERE, "_build", "libsc_oracle.so")
_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)


class ScoSpec(C.Structure):
    _fields_ = [("N", C.c_int32), ("P", C.c_int32), ("E", C.c_int32), ("T", C.c_int32), ("avgL", C.c_int32), ("maxL", C.c_int32),
                ("stochL", C.c_int32), ("n_init", C.c_int32),
                ("c_unmet", C.c_double), ("c_xstock", C.c_double), ("c_xproc", C.c_double), ("c_xship", C.c_double),
                ("last", _pi), ("ebeg", _pi), ("proc_cap", _pd),
                ("stock0", _pd), ("cap", _pd), ("hold", _pd), ("sup_cap", _pd), ("sup_cost", _pd), ("ratio", _pd), ("proc_cost", _pd),
                ("edst", _pi), ("ecap", _pd), ("ecost", _pd), ("ibeg", _pi), ("itime", _pi), ("iamt", _pd), ("dlo", _pd), ("dspan", _pd)]


def build(force=False):
    """gcc -O2 -ffp-contract=off: no fused multiply-adds, every operation rounds like numpy's."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    tmp = LIB + ".tmp%d" % os.getpid()
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c99", "-shared", "-fPIC", "-o", tmp, SRC], check=True)
    os.replace(tmp, LIB)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.sco_create.restype = C.c_void_p; L.sco_create.argtypes = [C.POINTER(ScoSpec)]
        L.sco_destroy.argtypes = [C.c_void_p]
        for f in ("sco_n_actions", "sco_n_obs", "sco_n_leadtimes", "sco_n_retailers"):
            getattr(L, f).restype = C.c_int; getattr(L, f).argtypes = [C.c_void_p]
        L.sco_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sco_step.restype = C.c_int; L.sco_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sco_observation.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.sco_stocks.argtypes = [C.c_void_p, C.c_void_p]
        L.sco_heap.restype = C.c_int; L.sco_heap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.sco_rollout.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.bgo_create.restype = C.c_void_p; L.bgo_create.argtypes = [C.c_int] * 9 + [C.c_void_p]
        L.bgo_destroy.argtypes = [C.c_void_p]
        L.bgo_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bgo_obs.argtypes = [C.c_void_p, C.c_void_p]
        L.bgo_step.restype = C.c_int; L.bgo_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class CSupplyChainOracle(object):
    """Same constructor as SupplyChainOracle / the reference's SupplyChainEnv; demand and lead times are injected."""

    def __init__(self, nodes_info, **kw):
        py = SupplyChainOracle(nodes_info, **kw)            # parses nodes_info exactly like the reference's constructor
        self.py = py
        P, nodes = py.P, py.nodes
        idx = {id(nd): i for i, nd in enumerate(nodes)}
        f64 = lambda rows: np.ascontiguousarray(np.array(rows, dtype=np.float64).reshape(-1))  # noqa: E731
        i32 = lambda rows: np.ascontiguousarray(np.array(rows, dtype=np.int32).reshape(-1))    # noqa: E731
        ebeg, edst, ecap, ecost = [0], [], [], []
        ibeg, itime, iamt = [0], [], []
        for nd in nodes:
            D = len(nd.dests) if nd.dests else 0
            for i in range(D):
                edst.append(idx[id(nd.dests[i])]); ecap.append(nd.ship_cap[i])
                ecost.append([nd.dest_cost[p][i] for p in range(P)])
            ebeg.append(ebeg[-1] + D)
            for p in range(P):
                for src in (nd.init_supply, nd.init_ship):            # push order of SC_Node.reset (:402-412)
                    if src:
                        for i, amount in enumerate(src[p]):
                            itime.append(i + 1); iamt.append(amount)
                ibeg.append(len(itime))
        if not py.by_product:
            dlo, dspan = [py.d_range[0]] * P, [py.d_span] * P
        else:
            dlo, dspan = [py.d_range[p][0] for p in range(P)], list(py.d_span)
        self._keep = dict(
            last=i32([1 if nd.last else 0 for nd in nodes]), ebeg=i32(ebeg), proc_cap=f64([nd.proc_cap for nd in nodes]),
            stock0=f64([nd.stock0 for nd in nodes]), cap=f64([nd.cap for nd in nodes]), hold=f64([nd.hold for nd in nodes]),
            sup_cap=f64([nd.sup_cap for nd in nodes]), sup_cost=f64([nd.sup_cost for nd in nodes]),
            ratio=f64([nd.proc_ratio for nd in nodes]), proc_cost=f64([nd.proc_cost for nd in nodes]),
            edst=i32(edst if edst else [0]), ecap=f64(ecap if ecap else [0]), ecost=f64(ecost if ecost else [0]),
            ibeg=i32(ibeg), itime=i32(itime if itime else [0]), iamt=f64(iamt if iamt else [0]), dlo=f64(dlo), dspan=f64(dspan))
        s = ScoSpec()
        s.N, s.P, s.E, s.T, s.avgL, s.maxL = len(nodes), P, len(edst), py.T, py.avg_lt, py.max_lt
        s.stochL, s.n_init = (1 if py.stoch_lt else 0), len(itime)
        s.c_unmet, s.c_xstock, s.c_xproc, s.c_xship = py.c_unmet, py.c_xstock, py.c_xproc, py.c_xship
        for k, a in self._keep.items():
            setattr(s, k, C.cast(_ptr(a), _pd if a.dtype == np.float64 else _pi))
        self._L = lib()
        self._h = self._L.sco_create(C.byref(s))
        self.n_actions, self.n_obs = self._L.sco_n_actions(self._h), self._L.sco_n_obs(self._h)
        self.n_leadtimes, self.n_retailers = self._L.sco_n_leadtimes(self._h), self._L.sco_n_retailers(self._h)
        self.P, self.T, self.N, self.stoch_lt = P, py.T, len(nodes), py.stoch_lt
        assert (self.n_actions, self.n_obs) == (py.n_actions, py.n_obs)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.sco_destroy(self._h)
            self._h = None

    def reset(self, demand, leadtimes=None):
        self._dem = np.ascontiguousarray(np.asarray(demand).reshape(self.T + 1, self.n_retailers, self.P), dtype=np.int32)
        self._lts = None if leadtimes is None else np.ascontiguousarray(leadtimes, dtype=np.int32)
        assert self._lts is not None or not self.stoch_lt
        self._L.sco_reset(self._h, _ptr(self._dem), None if self._lts is None else _ptr(self._lts))
        return self.observation()

    def observation(self, canonical=False):
        obs = np.empty(self.n_obs, dtype=np.float64)
        self._L.sco_observation(self._h, _ptr(obs), 1 if canonical else 0)
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        assert a.shape == (self.n_actions,)
        r = C.c_double()
        done = self._L.sco_step(self._h, _ptr(a), C.byref(r))
        return self.observation(), r.value, bool(done), {}

    def rollout(self, actions, canonical=False, want_obs=True):
        """actions [T, A] -> (obs [T, O] | None, reward [T], done [T]) in one C call."""
        a = np.ascontiguousarray(actions, dtype=np.float64)
        T = a.shape[0]
        obs = np.empty((T, self.n_obs), dtype=np.float64) if want_obs else None
        rew = np.empty(T, dtype=np.float64); done = np.empty(T, dtype=np.uint8)
        self._L.sco_rollout(self._h, T, _ptr(a), None if obs is None else _ptr(obs), _ptr(rew), _ptr(done), 1 if canonical else 0)
        return obs, rew, done

    def stocks(self):
        out = np.empty((self.N, self.P), dtype=np.float64)
        self._L.sco_stocks(self._h, _ptr(out))
        return out

    def heap_arrays(self, cap):
        ln = np.zeros((self.N, self.P), dtype=np.int32)
        tm = np.zeros((self.N, self.P, cap), dtype=np.int32); am = np.zeros((self.N, self.P, cap), dtype=np.float64)
        for n in range(self.N):
            for p in range(self.P):
                t1 = np.zeros(cap, dtype=np.int32); a1 = np.zeros(cap, dtype=np.float64)
                ln[n, p] = self._L.sco_heap(self._h, n * self.P + p, cap, _ptr(t1), _ptr(a1))
                assert ln[n, p] <= cap
                tm[n, p], am[n, p] = t1, a1
        return ln, tm, am


class CBeerGameOracle(object):
    """version 0 / 2 beer game in C; demand [T] and delays [T] (weeks 1..T) are injected at reset()."""

    def __init__(self, version=0, levels=4, weeks=35, initial_inventory=None, initial_shipment=4, initial_orders=4, inv_cost=1,
                 backlog_cost=2, max_stock=100, exceeded_capacity_penalty=100):
        self.L, self.T = levels, weeks
        inv0 = np.ascontiguousarray([12] * levels if initial_inventory is None else initial_inventory, dtype=np.int32)
        self._L = lib()
        self._h = self._L.bgo_create(version, levels, weeks, inv_cost, backlog_cost, max_stock, exceeded_capacity_penalty,
                                     initial_shipment, initial_orders, _ptr(inv0))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.bgo_destroy(self._h)
            self._h = None

    def _obs(self):
        o = np.empty(self.L, dtype=np.int64)
        self._L.bgo_obs(self._h, _ptr(o))
        return o

    def reset(self, demand, delays):
        self._dem = np.ascontiguousarray(demand, dtype=np.int32); self._del = np.ascontiguousarray(delays, dtype=np.int32)
        assert self._dem.shape == (self.T,) and self._del.shape == (self.T,)
        self._L.bgo_reset(self._h, _ptr(self._dem), _ptr(self._del))
        return self._obs()

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.int32)
        r = C.c_longlong()
        done = self._L.bgo_step(self._h, _ptr(a), C.byref(r))
        return self._obs(), int(r.value), bool(done), {}
