// scg_sc_device.cuh -- device-side SupplyChainEnv dynamics (one thread = one env).
//
// One templated node-step (`sc_step_nodes`) serves every kernel; three policies vary:
//   Topo  : where topology and per-node constants come from
//             DynTopo      runtime CSR + arrays in global memory (any nodes_info DAG)
//             LayeredTopo  compile-time S-F-W-R full-bipartite chain, constants in the kernel
//                          parameter bank (scg_sc_static.cuh) -> loops unroll, state in registers
//   Pipe  : how stock and in-transit material are stored
//             RingGlobal   pre-summed lead-time ring in global memory   (FAST modes, generic)
//             RingRegs     pre-summed lead-time ring in registers       (FAST modes, static)
//             HeapGlobal   CPython-heapq-exact queues of (time, amount) (EXACT mode)
//   Real  : float or double.
//
// Semantics restated from the reference (paths relative to its repo root):
//   supplychain_env.py:49-96    SC_Action.apply (SUPPLY scaling, SHIP sorted cuts)
//   supplychain_env.py:208-396  SC_Node.act
//   supplychain_env.py:402-412  SC_Node.reset
//   supplychain_env.py:428-463  SC_Node.build_observation
//   supplychain_env.py:697-748  step; :762-791 _build_observation
// including its quirks.  The fp64 translation unit is compiled with
// -fmad=false so every product and sum rounds exactly as numpy's does.
#pragma once
#include "scg_common.cuh"

namespace scg {

#ifndef SCG_TRIM
#define SCG_TRIM 63             // compile-time kernels: clips written as max(x, 0) arithmetic (fewer fp64 selects)
#endif
#ifndef SCG_LOOSE_RELU
#define SCG_LOOSE_RELU 1        // one-instruction fp64 max(x, 0) where a tiny positive residue is harmless
#endif

constexpr int SC_MAXD = 32;     // max destinations per node in the generic kernels
constexpr int SC_MAXP = 8;      // max products

// ---------------------------------------------------------------------------------
// Device copy of the spec (arrays in global memory, typed in the handle's Real).
// ---------------------------------------------------------------------------------
template <class Real>
struct ScDevSpec {
    int N, P, E, T, avgL, maxL, stochL, RP, A, O, K, ring, heapcap, n_init;
    Real c_unmet, c_xstock, c_xproc, c_xship;
    const int* last;        // [N]
    const int* ebeg;        // [N+1]
    const int* act_off;     // [N] first action index of the node
    const int* lt_off;      // [N] first lead-time index of the node (stochastic lead times)
    const int* ret_idx;     // [N] retailer ordinal (or -1)
    const int* is_sup;      // [N] node has at least one SUPPLY action
    const Real* proc_cap;   // [N]
    const Real *stock0, *cap, *hold, *sup_cap, *sup_cost, *ratio, *proc_cost, *max_ship;   // [N*P]
    const int* edst;        // [E]
    const Real* ecap;       // [E]
    const Real* ecost;      // [E*P]
    const int* ibeg;        // [N*P+1]
    const int* itime;       // [n_init]
    const Real* iamt;       // [n_init]
    const unsigned long long* lt_thr;   // [maxL] stochastic lead times: lt = 1 + #{j < maxL-1 : word >= lt_thr[j]}
    const float *inv_cap, *inv_ms;      // [N*P] 1/stock_capacity, 1/max_ship (0 where max_ship == 0): fp32 obs of the split kernel
    const double *q_scale, *q_inv;      // [N*P] fixed-point ring of the fp64-state split kernel: slot = amount * q_scale (a power of two)
    const float* inv_ms_q;              // [N*P] q_inv / max_ship: fp32 obs straight from a fixed-point slot
    const int* dkind;       // [P]
    const double *dlo, *dhi, *dstd, *dpeaks, *dminavg, *dmaxavg;   // [P]
};

// Per-launch inputs common to all kernels.
struct ScInputs {
    const int32_t* inj_demand;      // [T+1][RP][n] or nullptr
    const int32_t* inj_leadtimes;   // [T][K][n] or nullptr
    RngKey key;
    long long first_env;
    int auto_reset;
    int* err;                       // sticky SCG_ERR_* bits of the handle (scg_status); never null
};

// Sticky error bits a kernel can raise (read back by scg_status / the *_host calls).
enum : int { SCG_ERR_LEADTIME_RANGE = 1, SCG_ERR_QUEUE_OVERFLOW = 2, SCG_ERR_DELAY_RANGE = 4 };
__device__ __forceinline__ void raise_error(int* err, int bit) {
    if ((*(volatile int*)err & bit) == 0) atomicOr(err, bit);
}
// An injected lead time outside [1, max_leadtime] would index past the ring / overfill a queue: clamp it and flag it.
__device__ __forceinline__ int checked_leadtime(int lt, int maxL, int* err) {
    if (lt < 1 || lt > maxL) { raise_error(err, SCG_ERR_LEADTIME_RANGE); lt = lt < 1 ? 1 : maxL; }
    return lt;
}

// Global env state, env-minor ("SoA"): element i of env e lives at base[i*n + e].
template <class Real>
struct ScState {
    long long n;
    int* t;            // [n] period of the current episode (0 after reset)
    int* episode;      // [n] episode counter (Philox counter word)
    double* ep_ret;    // [n] return accumulated so far in the episode
    Real* stock;       // [N*P][n]
    Real* slot;        // FAST:  [N*P*ring][n]  slot k = material arriving at t+1+k
    int* hlen;         // EXACT: [N*P][n]
    int* htime;        // EXACT: [N*P*heapcap][n]
    double* hamt;      // EXACT: [N*P*heapcap][n]
    double* stats;     // [5] count, sum, sumsq, min, max of finished-episode returns
    double* info;      // build_info: [16*P][n] cumulative costs (8 keys x P) then units (8 keys x P); else nullptr
};

// ---------------------------------------------------------------------------------
// stochastic inputs
// ---------------------------------------------------------------------------------
template <class Real>
__device__ __forceinline__ int sc_demand(const ScDevSpec<Real>& sp, const ScInputs& in, long long n_envs, long long env,
                                         unsigned episode, int row, int rp, int p) {
    if (in.inj_demand) return in.inj_demand[((long long)row * sp.RP + rp) * n_envs + env];
    const unsigned long long genv = (unsigned long long)(in.first_env + env);
    const int kind = sp.dkind[p];
    const double lo = sp.dlo[p], hi = sp.dhi[p];
    if (kind == 0) {
        Philox4 b = rng_block(in.key, genv, episode, (unsigned)row, STREAM_DEMAND, (unsigned)(rp >> 2));
        return uniform_int(pick_word(b, rp & 3), (int)lo, (unsigned)((int)hi - (int)lo + 1));
    }
    Philox4 b = rng_block(in.key, genv, episode, (unsigned)row, STREAM_DEMAND_NOISE, (unsigned)(rp >> 1));
    const unsigned w0 = pick_word(b, 2 * (rp & 1)), w1 = pick_word(b, 2 * (rp & 1) + 1);
    const double sd = sp.dstd[p];
    double x;
    if (kind == 1) {
        x = (hi + lo) / 2 + sd * normal_from(w0, w1);                       // demands_generator.py:41
    } else {
        double base = sp.dminavg[p] + (sp.dmaxavg[p] - sp.dminavg[p]) / 2 *
                      (1.0 + sin(sp.dpeaks[p] * 2.0 * 3.141592653589793 / (double)sp.T * (double)row));   // :72-82
        double noise = (kind == 2) ? sd * normal_from(w0, w1)
                                   : (double)uniform_int(w0, -3 * (int)sd, (unsigned)(6 * (int)sd + 1));  // :75-77
        x = base + noise;
    }
    x = fmin(fmax(x, lo), hi);
    return (int)rint(x);                                                    // np.rint: half-to-even
}

template <class Real>
__device__ __forceinline__ int sc_leadtime(const ScDevSpec<Real>& sp, const ScInputs& in, long long n_envs, long long env,
                                           unsigned episode, int t, int k) {
    if (!sp.stochL) return sp.avgL;                                         // supplychain_env.py:724
    if (in.inj_leadtimes) return checked_leadtime(in.inj_leadtimes[((long long)(t - 1) * sp.K + k) * n_envs + env], sp.maxL, in.err);
    const unsigned long long genv = (unsigned long long)(in.first_env + env);
    Philox4 b = rng_block(in.key, genv, episode, (unsigned)t, STREAM_LEADTIME, (unsigned)(k >> 2));
    const unsigned word = pick_word(b, k & 3);
    int lt = 1;                                      // 1 + Poisson(avg-1) clipped to maxL, by inverse CDF (:670-672)
    for (int q = 0; q + 1 < sp.maxL; ++q) lt += (word > (unsigned)(sp.lt_thr[q] - 1ull)) ? 1 : 0;   // thr in [1, 2^32]
    return lt;
}

// uniform action in [-1, 1) from one Philox word; the float value is what every mode sees
__device__ __forceinline__ float action_from_word(unsigned w) { return (float)(int)w * 4.656612873077393e-10f; }

// ---------------------------------------------------------------------------------
// arithmetic helpers that differ between the exact fp64 path and the fp32 path
// ---------------------------------------------------------------------------------
template <class Real> __device__ __forceinline__ Real denorm_action(Real a);
// (a+1)/2 (:697-698).  fma(a, .5, .5) rounds (a+1)/2 once; halving is exact, so it equals fl(a+1)/2 bit for bit.
template <> __device__ __forceinline__ double denorm_action<double>(double a) { return fma(a, 0.5, 0.5); }
template <> __device__ __forceinline__ float denorm_action<float>(float a) { return fmaf(a, 0.5f, 0.5f); }

// max(x, 0).  fp64 has no min/max instruction on sm_100 (a compare plus two selects plus a NaN fix-up on the half-rate
// FP64 pipe), so the double flavour clears the words of a negative value with three integer operations instead.
template <class Real> __device__ __forceinline__ Real relu(Real x);
template <> __device__ __forceinline__ float relu<float>(float x) { return fmaxf(x, 0.0f); }
template <> __device__ __forceinline__ double relu<double>(double x) {
    const int hi = __double2hiint(x), keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(x) & keep);
}

// "Loose" max(x, 0) for amounts that are only ever ADDED to or subtracted from O(1)+ quantities, multiplied, or cast to
// fp32 -- never compared with 0: ONE instruction for a double.  The sign lives in the high word, and for |x| < 2^1017
// the high word read as an fp32 has that sign and is never a NaN, so fmaxf(high, 0) clears it exactly when x < 0.
// The low word is left alone: a negative x becomes a denormal < 5e-315 instead of 0, which vanishes in every sum
// with a normal number and casts to 0.0f.
template <class Real> __device__ __forceinline__ Real relu_loose(Real x);
template <> __device__ __forceinline__ float relu_loose<float>(float x) { return fmaxf(x, 0.0f); }
template <> __device__ __forceinline__ double relu_loose<double>(double x) {
#if SCG_LOOSE_RELU
    return __hiloint2double(__float_as_int(fmaxf(__int_as_float(__double2hiint(x)), 0.0f)), __double2loint(x));
#else
    return relu<double>(x);
#endif
}

// v, or "loosely" zero when d > 0: the double flavour tests and clears the high word only (one integer compare and one
// select instead of a compare on the FP64 pipe and two selects).  What stays behind is a non-negative denormal, and a
// positive denormal d does not count as positive -- both harmless for a capacity that is only ever subtracted from amounts.
template <class Real> __device__ __forceinline__ Real clear_if_positive_loose(Real v, Real d);
template <> __device__ __forceinline__ float clear_if_positive_loose<float>(float v, float d) { return d > 0.0f ? 0.0f : v; }
template <> __device__ __forceinline__ double clear_if_positive_loose<double>(double v, double d) {
#if SCG_LOOSE_RELU
    return __hiloint2double(__double2hiint(d) > 0 ? 0 : __double2hiint(v), __double2loint(v));
#else
    return d > 0.0 ? 0.0 : v;
#endif
}

// everything relu_loose can leave behind is below this; no real quantity of the simulation is
template <class Real> __device__ __forceinline__ constexpr Real loose_zero() { return sizeof(Real) == 8 ? (Real)1e-300 : (Real)0; }

template <class Real> __device__ __forceinline__ Real clamp_pm1(Real x) {
    return x < (Real)-1 ? (Real)-1 : (x > (Real)1 ? (Real)1 : x);           // np.clip(.., -1, 1), :700-701
}
template <class Real> __device__ __forceinline__ Real obs_norm(Real x) { return clamp_pm1<Real>(x * (Real)2 - (Real)1); }

// ---------------------------------------------------------------------------------
// Topo policy: runtime topology
// ---------------------------------------------------------------------------------
template <class Real>
struct DynTopo {
    static constexpr bool kStatic = false;
    static constexpr int kMaxD = SC_MAXD;
    const ScDevSpec<Real>& sp;
    __device__ __forceinline__ explicit DynTopo(const ScDevSpec<Real>& s) : sp(s) {}
    __device__ __forceinline__ int N() const { return sp.N; }
    __device__ __forceinline__ int P() const { return sp.P; }
    __device__ __forceinline__ int avgL() const { return sp.avgL; }
    __device__ __forceinline__ int maxL() const { return sp.maxL; }
    __device__ __forceinline__ int ring() const { return sp.ring; }
    __device__ __forceinline__ int T() const { return sp.T; }
    __device__ __forceinline__ bool last(int n) const { return sp.last[n] != 0; }
    __device__ __forceinline__ bool supplier(int n) const { return sp.is_sup[n] != 0; }
    __device__ __forceinline__ int ebeg(int n) const { return sp.ebeg[n]; }
    __device__ __forceinline__ int ndest(int n) const { return sp.ebeg[n + 1] - sp.ebeg[n]; }
    __device__ __forceinline__ int edst(int e) const { return sp.edst[e]; }
    __device__ __forceinline__ int dest_np(int e0, int i, int p) const { return sp.edst[e0 + i] * sp.P + p; }   // where edge e0+i delivers product p
    __device__ __forceinline__ Real ecap(int e) const { return sp.ecap[e]; }
    __device__ __forceinline__ Real ecost(int e, int p) const { return sp.ecost[e * sp.P + p]; }
    __device__ __forceinline__ Real proc_cap(int n) const { return sp.proc_cap[n]; }
    __device__ __forceinline__ Real cap(int n, int p) const { return sp.cap[n * sp.P + p]; }
    __device__ __forceinline__ Real hold(int n, int p) const { return sp.hold[n * sp.P + p]; }
    __device__ __forceinline__ Real sup_cap(int n, int p) const { return sp.sup_cap[n * sp.P + p]; }
    __device__ __forceinline__ Real sup_cost(int n, int p) const { return sp.sup_cost[n * sp.P + p]; }
    __device__ __forceinline__ Real ratio(int n, int p) const { return sp.ratio[n * sp.P + p]; }
    __device__ __forceinline__ Real proc_cost(int n, int p) const { return sp.proc_cost[n * sp.P + p]; }
    __device__ __forceinline__ Real max_ship(int n, int p) const { return sp.max_ship[n * sp.P + p]; }
    __device__ __forceinline__ int act_off(int n) const { return sp.act_off[n]; }
    __device__ __forceinline__ int lt_off(int n) const { return sp.lt_off[n]; }
    __device__ __forceinline__ int ret_idx(int n) const { return sp.ret_idx[n]; }
    __device__ __forceinline__ Real c_unmet() const { return sp.c_unmet; }
    __device__ __forceinline__ Real c_xstock() const { return sp.c_xstock; }
    __device__ __forceinline__ Real c_xproc() const { return sp.c_xproc; }
    __device__ __forceinline__ Real c_xship() const { return sp.c_xship; }
    __device__ __forceinline__ bool has_supply(int n, int p) const { return sup_cap(n, p) > (Real)0; }
    __device__ __forceinline__ bool can_ship(int n, int p) const { return cap(n, p) > (Real)0 && ndest(n) > 0; }
    __device__ __forceinline__ Real div_ratio(Real x, int n, int p) const { return x / ratio(n, p); }   // :310
};

// ---------------------------------------------------------------------------------
// Pipe policies (global-memory flavours; the register flavour lives in scg_sc_static.cuh)
// ---------------------------------------------------------------------------------
template <class Real>
struct RingGlobal {
    static constexpr bool kTwoPhase = true;    // all arrivals first, then all node actions
    static constexpr bool kExactObs = false;
    Real* stock; Real* slot; long long n; int ring;     // n = element stride (n_envs in HBM, blockDim.x in shared memory)
    __device__ __forceinline__ RingGlobal(Real* stock_, Real* slot_, long long stride, int ring_)
        : stock(stock_), slot(slot_), n(stride), ring(ring_) {}
    __device__ __forceinline__ Real get_stock(int np) const { return stock[(long long)np * n]; }
    __device__ __forceinline__ void set_stock(int np, Real v) { stock[(long long)np * n] = v; }
    __device__ __forceinline__ Real slot_at(int np, int k) const { return slot[((long long)np * ring + k) * n]; }
    // material arriving now (slot 0), and shift the ring by one period
    __device__ __forceinline__ Real arrive(int np, int /*t*/) {
        Real* s = slot + (long long)np * ring * n;
        Real a = s[0];
        for (int k = 0; k + 1 < ring; ++k) s[(long long)k * n] = s[(long long)(k + 1) * n];
        s[(long long)(ring - 1) * n] = (Real)0;
        return a;
    }
    __device__ __forceinline__ void push(int np, int /*t*/, int lt, Real amount) {
        slot[((long long)np * ring + (lt - 1)) * n] += amount;
    }
    __device__ __forceinline__ void push_own(int np, int t, int lt, Real amount) { push(np, t, lt, amount); }   // a supplier's own order
    template <class Tp>
    __device__ __forceinline__ void reset(const Tp& tp, const ScDevSpec<Real>& sp) {
        const int NP = sp.N * sp.P;
        for (int np = 0; np < NP; ++np) {
            set_stock(np, sp.stock0[np]);
            for (int k = 0; k < ring; ++k) slot[((long long)np * ring + k) * n] = (Real)0;
            for (int i = sp.ibeg[np]; i < sp.ibeg[np + 1]; ++i)
                slot[((long long)np * ring + (sp.itime[i] - 1)) * n] += sp.iamt[i];
        }
    }
};

// EXACT: per (node, product) a binary heap of (time, amount) laid out exactly like CPython's
// heapq list (Lib/heapq.py: heappush/_siftdown, heappop/_siftup), because the reference
// observation reads that list by index (supplychain_env.py:445-461).
struct HeapGlobal {
    static constexpr bool kTwoPhase = false;   // reference order: node by node
    static constexpr bool kExactObs = true;
    double* stock; int* hlen; int* htime; double* hamt; long long n; int cap; int* err;
    __device__ __forceinline__ HeapGlobal(double* stock_, int* hlen_, int* htime_, double* hamt_, long long stride, int cap_, int* err_)
        : stock(stock_), hlen(hlen_), htime(htime_), hamt(hamt_), n(stride), cap(cap_), err(err_) {}
    __device__ __forceinline__ double get_stock(int np) const { return stock[(long long)np * n]; }
    __device__ __forceinline__ void set_stock(int np, double v) { stock[(long long)np * n] = v; }
    __device__ __forceinline__ int& len(int np) { return hlen[(long long)np * n]; }
    __device__ __forceinline__ int& tm(int np, int k) { return htime[((long long)np * cap + k) * n]; }
    __device__ __forceinline__ double& am(int np, int k) { return hamt[((long long)np * cap + k) * n]; }
    static __device__ __forceinline__ bool lt(int ta, double aa, int tb, double ab) {     // tuple '<'
        return ta < tb || (ta == tb && aa < ab);
    }
    __device__ __forceinline__ void sift_down(int np, int start, int pos, int nt, double na) {
        while (pos > start) {
            int par = (pos - 1) >> 1;
            int pt = tm(np, par); double pa = am(np, par);
            if (lt(nt, na, pt, pa)) { tm(np, pos) = pt; am(np, pos) = pa; pos = par; } else break;
        }
        tm(np, pos) = nt; am(np, pos) = na;
    }
    __device__ __forceinline__ void push_item(int np, int when, double amount) {
        int L = len(np);
        // cap = n_src*(maxL+1) + n_init holds every legal trajectory (lead times are range-checked on read); if a
        // caller-built state blob (scg_set_state) overfills a queue the entry is dropped and the handle flagged
        if (L >= cap) { raise_error(err, SCG_ERR_QUEUE_OVERFLOW); return; }
        len(np) = L + 1;
        sift_down(np, 0, L, when, amount);
    }
    __device__ __forceinline__ double pop_item(int np) {
        int L = len(np) - 1;
        len(np) = L;
        int lt_t = tm(np, L); double lt_a = am(np, L);
        if (L == 0) return lt_a;
        double ret = am(np, 0);
        // _siftup(heap, 0) with newitem = last element
        int pos = 0, child = 1;
        while (child < L) {
            int right = child + 1;
            if (right < L && !lt(tm(np, child), am(np, child), tm(np, right), am(np, right))) child = right;
            tm(np, pos) = tm(np, child); am(np, pos) = am(np, child);
            pos = child; child = 2 * pos + 1;
        }
        sift_down(np, 0, pos, lt_t, lt_a);
        return ret;
    }
    __device__ __forceinline__ double arrive(int np, int t) {                              // :220-225
        double got = 0.0;
        while (len(np) > 0 && tm(np, 0) == t) got += pop_item(np);
        return got;
    }
    __device__ __forceinline__ void push(int np, int t, int lt_, double amount) { push_item(np, t + lt_, amount); }
    __device__ __forceinline__ void push_own(int np, int t, int lt_, double amount) { push_item(np, t + lt_, amount); }
    template <class Tp>
    __device__ __forceinline__ void reset(const Tp& tp, const ScDevSpec<double>& sp) {     // :402-412
        const int NP = sp.N * sp.P;
        for (int np = 0; np < NP; ++np) {
            set_stock(np, sp.stock0[np]);
            len(np) = 0;
            for (int i = sp.ibeg[np]; i < sp.ibeg[np + 1]; ++i) push_item(np, sp.itime[i], sp.iamt[i]);
        }
    }
};

// ---------------------------------------------------------------------------------
// SHIP action scaling (supplychain_env.py:58-96): destinations sorted by (action value, index);
// each gets the slice of `limit` between consecutive cut points, never more than what is left.
// ---------------------------------------------------------------------------------
template <class Real, int MAXD>
__device__ __forceinline__ void ship_cuts_dyn(const Real* v, int D, Real limit, Real* out) {
    int order[MAXD];
    for (int i = 0; i < D; ++i) {            // stable insertion sort on (value, index)
        int j = i;
        while (j > 0 && v[order[j - 1]] > v[i]) { order[j] = order[j - 1]; --j; }
        order[j] = i;
    }
    Real left = limit, prev = (Real)0;
    for (int k = 0; k < D; ++k) {
        int i = order[k];
        Real amount = (v[i] - prev) * limit;
        if (amount > left) amount = left;
        out[i] = amount;
        left -= amount;
        prev = v[i];
    }
}

// ---------------------------------------------------------------------------------
// build_info statistics (supplychain_env.py:214-218 per node, :750-760 cumulative): 8 keys x P products of
// costs and of units.  Key order = the reference's dict order.
// ---------------------------------------------------------------------------------
enum : int { INFO_STOCK = 0, INFO_STOCK_PEN = 1, INFO_SUPPLY = 2, INFO_PROCESS = 3, INFO_PROCESS_PEN = 4, INFO_SHIP = 5,
             INFO_SHIP_PEN = 6, INFO_UNMET = 7 };

struct NoInfo {
    static constexpr bool kOn = false;
    template <class R> __device__ __forceinline__ void add(int, int, R, R) const {}
};
struct GlobalInfo {                       // env-minor cumulative arrays; node order = accumulation order of the reference
    static constexpr bool kOn = true;
    double* base; long long n; int P;
    template <class R> __device__ __forceinline__ void add(int key, int p, R cost, R units) const {
        base[(long long)(key * P + p) * n] += (double)cost;
        base[(long long)((8 + key) * P + p) * n] += (double)units;
    }
    __device__ __forceinline__ void clear() const {
        for (int i = 0; i < 16 * P; ++i) base[(long long)i * n] = 0.0;
    }
};

// ---------------------------------------------------------------------------------
// One period for all nodes of one env.  Returns the total cost (reward = -cost).
//   Ar: action reader  Real get(int idx)   (raw action in [-1,1])
//   Lr: lead-time reader  int get(int k)
//   Dr: demand reader  int cur(int rp, int p)   (row t-1)
// ---------------------------------------------------------------------------------
template <class Real, class Tp, class Pp, class Ar, class Lr, class Dr, class Inf = NoInfo>
__device__ __forceinline__ Real sc_step_nodes(const Tp& tp, Pp& pp, int t, Ar& ar, Lr& lr, Dr& dr, const Inf& inf = Inf()) {
    constexpr int MAXD = Tp::kMaxD;
    const int N = tp.N(), P = tp.P();
    if (Pp::kTwoPhase) {
#pragma unroll(Tp::kStatic ? 64 : 1)
        for (int n = 0; n < N; ++n)
#pragma unroll(Tp::kStatic ? 8 : 1)
            for (int p = 0; p < P; ++p) {
                const int np = n * P + p;
                pp.set_stock(np, pp.get_stock(np) + pp.arrive(np, t));
            }
    }
    Real total = (Real)0;
#pragma unroll(Tp::kStatic ? 64 : 1)
    for (int n = 0; n < N; ++n) {
        Real cost = (Real)0;
        // 1. arrivals (:220-228), 2. over-capacity: pay and discard (:232-240)
#pragma unroll(Tp::kStatic ? 8 : 1)
        for (int p = 0; p < P; ++p) {
            const int np = n * P + p;
            Real s = pp.get_stock(np);
            if (!Pp::kTwoPhase) s = s + pp.arrive(np, t);
            const Real c = tp.cap(n, p);
            if constexpr (Tp::kStatic && (SCG_TRIM & 1)) {
                // same update as arithmetic: over = max(s - c, 0) is paid for and discarded
                const Real over = relu_loose<Real>(s - c);
                cost += tp.c_xstock() * over;
                inf.add(INFO_STOCK_PEN, p, tp.c_xstock() * over, over);
                s -= over;
            } else if (s > c) {
                const Real pen = tp.c_xstock() * (s - c);
                cost += pen;
                inf.add(INFO_STOCK_PEN, p, pen, s - c);
                s = c;
            }
            pp.set_stock(np, s);
        }
        int ai = tp.act_off(n);
        int li = tp.lt_off(n);
        // 3. supply (:244-256); the lead-time cursor moves only when something was supplied (F4b)
        if (tp.supplier(n)) {
#pragma unroll(Tp::kStatic ? 8 : 1)
            for (int p = 0; p < P; ++p) {
                if (tp.has_supply(n, p)) {
                    const Real amount = denorm_action<Real>(ar.get(ai)) * tp.sup_cap(n, p);
                    ++ai;
                    if constexpr (Tp::kStatic && (SCG_TRIM & 32) && Lr::kConst) pp.push_own(n * P + p, t, lr.get(li), relu_loose<Real>(amount));
                    else if (amount > (Real)0) { pp.push_own(n * P + p, t, lr.get(li), amount); ++li; }
                    const Real sup = amount * tp.sup_cost(n, p);
                    cost += sup;
                    inf.add(INFO_SUPPLY, p, sup, amount);
                }
            }
        }
        if (!tp.last(n)) {
            // 4. ship (:262-375)
            const int D = tp.ndest(n), e0 = tp.ebeg(n);
            Real room[MAXD];
#pragma unroll(Tp::kStatic ? MAXD : 1)
            for (int i = 0; i < D; ++i) room[i] = tp.ecap(e0 + i);
            Real proc_left = tp.proc_cap(n);
            bool factory = proc_left > (Real)0;
            if constexpr (Tp::kStatic) factory = tp.is_factory(n);          // known at compile time (sc_is_layered)
            const int base = li;
#pragma unroll(Tp::kStatic ? 8 : 1)
            for (int p = 0; p < P; ++p) {
                const int np = n * P + p;
                if (!tp.can_ship(n, p)) continue;                            // no SHIP action (:196-201)
                Real over_ship = (Real)0, over_proc = (Real)0;
                const Real s = pp.get_stock(np);
                typename Ar::raw_t rv[MAXD];
                // (read unconditionally so the staged-row reader sees straight-line code)
#pragma unroll(Tp::kStatic ? MAXD : 1)
                for (int i = 0; i < D; ++i) rv[i] = ar.get_raw(ai + i);
                // The reference skips the whole block when stock <= 0 (:283).  The compile-time kernels run it with
                // limit = 0 instead (every amount 0, nothing pushed, costs += 0): same result, no divergent branch.
                if (Tp::kStatic || s > (Real)0) {
                    // limit = min(stock_capacity, stock) (:61-65) is just the stock: it was clipped to the capacity above
                    Real limit = s;
                    if (Tp::kStatic) limit = (SCG_TRIM & 2) ? relu_loose<Real>(limit) : (limit > (Real)0 ? limit : (Real)0);
                    Real out[MAXD], send[MAXD];
                    tp.ship_cuts(rv, D, limit, out);
                    const Real r = tp.ratio(n, p);
#pragma unroll(Tp::kStatic ? MAXD : 1)
                    for (int i = 0; i < D; ++i) send[i] = out[i];
                    if (factory) {                                           // :298-310
#pragma unroll(Tp::kStatic ? MAXD : 1)
                        for (int i = 0; i < D; ++i) {
                            if constexpr (Tp::kStatic && (SCG_TRIM & 4)) {
                                // arithmetic form (proc_left >= 0, so a non-positive amount has no excess and
                                // consumes nothing): ex = what does not fit; the rest is processed.  A clipped
                                // amount must be EXACTLY proc_left (and the budget then exactly 0), not
                                // out - (out - proc_left): a residue of 1e-13 where the reference has 0 flips
                                // `send > 0` below, and with a negative ship capacity (F4a) that is a whole clip.
                                const Real d = out[i] - proc_left, ex = relu_loose<Real>(d);
                                over_proc = i == 0 ? ex : over_proc + ex;        // (0 + x is not folded by the compiler)
                                out[i] = d > (Real)0 ? proc_left : out[i];
                                proc_left -= relu_loose<Real>(out[i]);
                            } else {
                                if (out[i] > (Real)0) {
                                    if (out[i] > proc_left) { over_proc += out[i] - proc_left; out[i] = proc_left; }
                                    proc_left -= out[i];
                                }
                            }
                            send[i] = tp.div_ratio(out[i], n, p);
                        }
                    }
#pragma unroll(Tp::kStatic ? MAXD : 1)
                    for (int i = 0; i < D; ++i) {                            // :312-328 (F4a)
                        if constexpr (Tp::kStatic && (SCG_TRIM & 8)) {
                            if (!factory) {
                                // no processing ratio: the capacity left never goes negative, so the clip is
                                // ex = max(send - room, 0) and a clipped destination is left with no capacity
                                const Real d = send[i] - room[i], ex = relu_loose<Real>(d);
                                over_ship = i == 0 ? ex : over_ship + ex;
                                send[i] -= ex;
                                out[i] = send[i];
                                room[i] = clear_if_positive_loose<Real>(room[i], d);
                            } else if (send[i] > loose_zero<Real>() && send[i] > room[i]) {
                                // a factory's capacity can be NEGATIVE here (after a clip of an earlier product the
                                // reference subtracts raw-material units, F4a), so this `send > 0` decides a whole
                                // clip: the denormals that relu_loose leaves where the reference has exact zeros
                                // (unpushed negative amounts, emptied stocks) must not count as shipments
                                over_ship += send[i] - room[i];
                                send[i] = room[i];
                                out[i] = send[i] * r;
                                room[i] -= out[i];
                            }
                        } else {
                            if (send[i] > (Real)0 && send[i] > room[i]) {
                                over_ship += send[i] - room[i];
                                send[i] = room[i];
                                out[i] = factory ? send[i] * r : send[i];
                                room[i] -= out[i];
                            }
                        }
                    }
                    Real leaving = (Real)0;
#pragma unroll(Tp::kStatic ? MAXD : 1)
                    for (int i = 0; i < D; ++i) leaving = (Tp::kStatic && i == 0) ? out[0] : leaving + out[i];   // :331
                    pp.set_stock(np, s - leaving);
                    if (factory) {                                           // :338
                        const Real pc = leaving * tp.proc_cost(n, p);
                        cost += pc;
                        inf.add(INFO_PROCESS, p, pc, leaving);
                    }
                    Real sc = (Real)0;
#pragma unroll(Tp::kStatic ? MAXD : 1)
                    for (int i = 0; i < D; ++i) {
                        if constexpr (Tp::kStatic && (SCG_TRIM & 16)) pp.push(tp.dest_np(e0, i, p), t, lr.get(base + i), relu_loose<Real>(send[i]));
                        else if (send[i] > (Real)0) pp.push(tp.dest_np(e0, i, p), t, lr.get(base + i), send[i]);   // :344-348
                        sc += send[i] * tp.ecost(e0 + i, p);                 // :352-353
                    }
                    cost += sc;
                    if constexpr (Inf::kOn) {
                        Real shipped = (Real)0;                              // sum(amounts_to_ship), :356
                        for (int i = 0; i < D; ++i) shipped += send[i];
                        inf.add(INFO_SHIP, p, sc, shipped);
                    }
                }
                const Real ppen = tp.c_xproc() * over_proc, spen = tp.c_xship() * over_ship;
                if (!Tp::kStatic || factory) cost += ppen;                   // :361 (over_proc is identically 0 elsewhere)
                cost += spen;                                                // :366
                inf.add(INFO_PROCESS_PEN, p, ppen, over_proc);
                inf.add(INFO_SHIP_PEN, p, spen, over_ship);
                ai += D;
            }
        } else {
            // 5. retailers: serve what is possible, the rest is lost (:380-384)
            const int r0 = tp.ret_idx(n) * P;
#pragma unroll(Tp::kStatic ? 8 : 1)
            for (int p = 0; p < P; ++p) {
                const int np = n * P + p;
                const Real d = (Real)dr.cur(r0 + p, p);
                const Real s = pp.get_stock(np);
                const Real served = s < d ? s : d;
                pp.set_stock(np, s - served);
                const Real lost = tp.c_unmet() * (d - served);
                cost += lost;
                inf.add(INFO_UNMET, p, lost, d - served);
            }
        }
        // 6. holding cost (:390-391)
#pragma unroll(Tp::kStatic ? 8 : 1)
        for (int p = 0; p < P; ++p) {
            const Real held = pp.get_stock(n * P + p), hc = held * tp.hold(n, p);
            cost += hc;
            inf.add(INFO_STOCK, p, hc, held);
        }
        total += cost;
    }
    return total;
}

// Generic (runtime) helpers selected through the Topo policy.
template <class Real>
struct DynTopoOps : DynTopo<Real> {
    __device__ __forceinline__ explicit DynTopoOps(const ScDevSpec<Real>& s) : DynTopo<Real>(s) {}
    template <class RawT>
    __device__ __forceinline__ void ship_cuts(const RawT* rv, int D, Real limit, Real* out) const {
        Real v[SC_MAXD];
        for (int i = 0; i < D; ++i) v[i] = denorm_action<Real>((Real)rv[i]);      // :697-698
        ship_cuts_dyn<Real, SC_MAXD>(v, D, limit, out);
    }
};

// ---------------------------------------------------------------------------------
// Observation (:428-463, :762-791).  Ow: void put(int idx, Real v).
// ---------------------------------------------------------------------------------
// FAST modes: canonical per-arrival-time sums out of the ring.
template <class Real, class Tp, class Pp, class Dr, class Ow>
__device__ __forceinline__ void sc_obs_ring(const Tp& tp, const ScDevSpec<Real>& sp, const Pp& pp, int t, Dr& dr, Ow& ow) {
    const int N = tp.N(), P = tp.P(), L = tp.avgL(), ring = tp.ring();
    int o = 0;
    for (int rp = 0; rp < sp.RP; ++rp) {
        const int p = rp % P;
        const Real lo = (Real)sp.dlo[p], span = (Real)(sp.dhi[p] - sp.dlo[p]);
        ow.put(o++, obs_norm<Real>(((Real)dr.next(rp, p) - lo) / span));
    }
    for (int n = 0; n < N; ++n) {
        for (int p = 0; p < P; ++p) ow.put(o++, obs_norm<Real>(pp.get_stock(n * P + p) / tp.cap(n, p)));
        for (int p = 0; p < P; ++p) {
            const int np = n * P + p;
            // a (node, product) nothing can be shipped to (max_ship == 0) has an empty queue: zeros, like the
            // reference's empty-heap branch (:438-439), not 0/0
            const Real ms = tp.max_ship(n, p);
            for (int k = 0; k + 1 < L; ++k) ow.put(o++, obs_norm<Real>(ms > (Real)0 ? pp.slot_at(np, k) / ms : (Real)0));
            Real rest = (Real)0;
            for (int k = L - 1; k < ring; ++k) rest += pp.slot_at(np, k);
            ow.put(o++, obs_norm<Real>(ms > (Real)0 ? rest / (ms * (Real)(tp.maxL() - (L - 1))) : (Real)0));
        }
    }
    ow.put(o++, obs_norm<Real>((Real)(tp.T() - t) / (Real)tp.T()));
}

// EXACT mode: walk each heap ARRAY in index order exactly like the reference does.
template <class Tp, class Dr, class Ow>
__device__ __forceinline__ void sc_obs_heap(const Tp& tp, const ScDevSpec<double>& sp, HeapGlobal& pp, int t, Dr& dr, Ow& ow) {
    const int N = tp.N(), P = tp.P(), L = tp.avgL();
    int o = 0;
    for (int rp = 0; rp < sp.RP; ++rp) {
        const int p = rp % P;
        ow.put(o++, obs_norm<double>(((double)dr.next(rp, p) - sp.dlo[p]) / (sp.dhi[p] - sp.dlo[p])));
    }
    for (int n = 0; n < N; ++n) {
        for (int p = 0; p < P; ++p) ow.put(o++, obs_norm<double>(pp.get_stock(n * P + p) / tp.cap(n, p)));
        for (int p = 0; p < P; ++p) {
            const int np = n * P + p;
            const int len = pp.len(np);
            const double ms = tp.max_ship(n, p);
            if (len == 0) { for (int k = 0; k < L; ++k) ow.put(o++, obs_norm<double>(0.0)); continue; }
            int k = 0;
            for (int when = t + 1; when < t + L; ++when) {
                double s = 0.0;
                while (k < len && pp.tm(np, k) == when) { s += pp.am(np, k); ++k; }
                ow.put(o++, obs_norm<double>(s / ms));
            }
            double s = 0.0;
            while (k < len) { s += pp.am(np, k); ++k; }
            ow.put(o++, obs_norm<double>(s / (ms * (double)(tp.maxL() - (L - 1)))));
        }
    }
    ow.put(o++, obs_norm<double>((double)(tp.T() - t) / (double)tp.T()));
}

// ---------------------------------------------------------------------------------
// readers/writers used by the generic kernels
// ---------------------------------------------------------------------------------
template <class Real, class Io>
struct GlobalActions {                 // actions[(step*n + env)*A + idx]  or Philox when null
    const Io* row; bool random; RngKey key; unsigned long long genv; unsigned episode; int t;
    Philox4 blk; int blk_id;
    using raw_t = Io;
    __device__ __forceinline__ Io get_raw(int idx) {
        if (!random) return row[idx];
        const int b = idx >> 2;
        if (b != blk_id) { blk = rng_block(key, genv, episode, (unsigned)t, STREAM_ACTION, (unsigned)b); blk_id = b; }
        return (Io)action_from_word(pick_word(blk, idx & 3));
    }
    __device__ __forceinline__ Real get(int idx) { return (Real)get_raw(idx); }
};

template <class Real>
struct DynLeadtimes {
    static constexpr bool kConst = false;
    const ScDevSpec<Real>& sp; const ScInputs& in; long long n, env; unsigned episode; int t;
    __device__ __forceinline__ int get(int k) const { return sc_leadtime(sp, in, n, env, episode, t, k); }
};

template <class Real>
struct DynDemand {
    const ScDevSpec<Real>& sp; const ScInputs& in; long long n, env; unsigned episode; int t;   // t = period just simulated
    __device__ __forceinline__ int cur(int rp, int p) const { return sc_demand(sp, in, n, env, episode, t - 1, rp, p); }
    __device__ __forceinline__ int next(int rp, int p) const { return sc_demand(sp, in, n, env, episode, t, rp, p); }
};

template <class Real, class Io>
struct GlobalObs {
    Io* row;
    __device__ __forceinline__ void put(int idx, Real v) { if (row) row[idx] = (Io)v; }
};

// contiguous warp copy, 16 B per lane when both sides are 16-byte aligned
template <class T>
__device__ __forceinline__ void warp_copy(T* __restrict__ dst, const T* __restrict__ src, int count, int lane) {
    const bool vec = ((((size_t)dst) | ((size_t)src)) & 15) == 0;
    constexpr int VW = 16 / (int)sizeof(T);
    if (vec) {
        const int nv = count / VW;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
        for (int i = lane; i < nv; i += 32) d4[i] = s4[i];
        for (int i = nv * VW + lane; i < count; i += 32) dst[i] = src[i];
    } else {
        for (int i = lane; i < count; i += 32) dst[i] = src[i];
    }
}

// ---------------------------------------------------------------------------------
// Generic kernels: any topology (runtime CSR), one thread per env.
//   use_smem : the env state (stocks + ring, or stocks + heapq arrays) is copied into shared memory for the
//              whole launch ([element][thread] layout: conflict free) and written back at the end; otherwise
//              every access goes to HBM (huge specs that do not fit 200 KB per block).
//   stage_io : each warp moves its 32 action rows / obs rows through a shared-memory tile with coalesced
//              16-byte accesses instead of 32 strided row accesses per instruction.
//   MODE 0: FAST ring (Real = float|double);  MODE 1: EXACT heap (Real = double)
// n_steps = 0 with a mask gives reset(); the same kernel serves step() (n_steps = 1).
// ---------------------------------------------------------------------------------
template <class Real, int MODE> struct PipeSel { using type = RingGlobal<Real>; };
template <> struct PipeSel<double, 1> { using type = HeapGlobal; };

template <class Real, int MODE, class Pp, class Dr, class Ow>
__device__ __forceinline__ void sc_write_obs(const DynTopoOps<Real>& tp, const ScDevSpec<Real>& sp, Pp& pp, int t, Dr& dr, Ow& ow) {
    if constexpr (MODE == 1) sc_obs_heap(tp, sp, pp, t, dr, ow);
    else sc_obs_ring<Real>(tp, sp, pp, t, dr, ow);
}

template <class Real, class Io, int MODE, bool INFO>
__global__ void __launch_bounds__(128) sc_generic_kernel(ScDevSpec<Real> sp, ScState<Real> st, ScInputs in, int n_steps,
                                                         const unsigned char* __restrict__ reset_mask, int do_reset,
                                                         const Io* __restrict__ actions, Io* __restrict__ obs,
                                                         Io* __restrict__ reward, unsigned char* __restrict__ done,
                                                         int use_smem, int stage_io) {
    extern __shared__ __align__(16) unsigned char g_smem[];
    const int tid = threadIdx.x, bd = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const long long env = (long long)blockIdx.x * bd + tid;
    const long long n = st.n;
    const bool active = env < n;
    const long long env0 = (long long)blockIdx.x * bd + warp * 32;          // first env of this warp
    const int n_here = env0 >= n ? 0 : (int)((n - env0) < 32 ? (n - env0) : 32);
    const int NP = sp.N * sp.P;
    ReturnStats rs; rs.init();

    // ---- shared-memory carve-up: [doubles][Reals][ints][per-warp io tiles] ----
    const int n_amt = (MODE == 1) ? NP * sp.heapcap : 0;                       // heap amounts (double)
    const int n_real = NP + ((MODE == 1) ? 0 : NP * sp.ring);                  // stocks (+ ring slots)
    const int n_int = (MODE == 1) ? NP + NP * sp.heapcap : 0;                  // heap lengths + times
    double* sm_amt = reinterpret_cast<double*>(g_smem);
    Real* sm_real = reinterpret_cast<Real*>(sm_amt + (use_smem ? (size_t)n_amt * bd : 0));
    int* sm_int = reinterpret_cast<int*>(sm_real + (use_smem ? (size_t)n_real * bd : 0));
    size_t io_off = (size_t)(reinterpret_cast<unsigned char*>(sm_int + (use_smem ? (size_t)n_int * bd : 0)) - g_smem);
    io_off = (io_off + 15) & ~(size_t)15;
    Io* s_act = reinterpret_cast<Io*>(g_smem + io_off) + (size_t)warp * 32 * (sp.A + sp.O);
    Io* s_obs = s_act + 32 * sp.A;

    // ---- state pointers of this env: shared memory (stride bd) or HBM (stride n) ----
    const long long stride = use_smem ? bd : n;
    Real* p_stock = use_smem ? sm_real + tid : st.stock + env;
    Real* p_slot = use_smem ? sm_real + (size_t)NP * bd + tid : st.slot + env;
    double* p_amt = use_smem ? sm_amt + tid : st.hamt + env;
    int* p_len = use_smem ? sm_int + tid : st.hlen + env;
    int* p_time = use_smem ? sm_int + (size_t)NP * bd + tid : st.htime + env;
    if (use_smem && active && !(do_reset && (reset_mask == nullptr || reset_mask[env]))) {
        for (int i = 0; i < NP; ++i) p_stock[(size_t)i * bd] = st.stock[(long long)i * n + env];
        if constexpr (MODE == 1) {
            for (int i = 0; i < NP; ++i) {
                const int len = st.hlen[(long long)i * n + env];
                p_len[(size_t)i * bd] = len;
                for (int k = 0; k < len; ++k) {
                    p_time[((size_t)i * sp.heapcap + k) * bd] = st.htime[((long long)i * sp.heapcap + k) * n + env];
                    p_amt[((size_t)i * sp.heapcap + k) * bd] = st.hamt[((long long)i * sp.heapcap + k) * n + env];
                }
            }
        } else {
            for (int i = 0; i < NP * sp.ring; ++i) p_slot[(size_t)i * bd] = st.slot[(long long)i * n + env];
        }
    }

    DynTopoOps<Real> tp(sp);
    typename PipeSel<Real, MODE>::type pp = [&]() {
        if constexpr (MODE == 1) return HeapGlobal(p_stock, p_len, p_time, p_amt, stride, sp.heapcap, in.err);
        else return RingGlobal<Real>(p_stock, p_slot, stride, sp.ring);
    }();
    int t = 0; unsigned episode = 0; double ep_ret = 0.0;
    GlobalInfo ginf{(INFO && active) ? st.info + env : nullptr, n, sp.P};
    if (active) {
        t = st.t[env]; episode = (unsigned)st.episode[env]; ep_ret = st.ep_ret[env];
        if (do_reset && (reset_mask == nullptr || reset_mask[env])) {
            ++episode; t = 0; ep_ret = 0.0;
            pp.reset(tp, sp);
            if constexpr (INFO) ginf.clear();
        }
    }
    const bool staged_out = stage_io && obs != nullptr;
    const bool staged_in = stage_io && actions != nullptr;

    if (n_steps == 0 && obs != nullptr) {       // reset(): emit the current observation
        if (active) {
            DynDemand<Real> dr{sp, in, n, env, episode, t};
            GlobalObs<Real, Io> ow{staged_out ? s_obs + lane * sp.O : obs + env * sp.O};
            sc_write_obs<Real, MODE>(tp, sp, pp, t, dr, ow);
        }
        if (staged_out) {
            __syncwarp();
            if (n_here > 0) warp_copy<Io>(obs + env0 * sp.O, s_obs, n_here * sp.O, lane);
        }
    }
    for (int s = 0; s < n_steps; ++s) {
        const long long io = (long long)s * n + env;
        const long long io0 = (long long)s * n + env0;
        if (staged_in) {
            __syncwarp();
            if (n_here > 0) warp_copy<Io>(s_act, actions + io0 * sp.A, n_here * sp.A, lane);
            __syncwarp();
        }
        if (active) {
            GlobalObs<Real, Io> ow{obs == nullptr ? nullptr : (staged_out ? s_obs + lane * sp.O : obs + io * sp.O)};
            if (t >= sp.T) {      // finished and not auto-reset: frozen
                DynDemand<Real> dr{sp, in, n, env, episode, t};
                sc_write_obs<Real, MODE>(tp, sp, pp, t, dr, ow);
                reward[io] = (Io)0; done[io] = 1;
            } else {
                ++t;
                const Io* arow = actions == nullptr ? nullptr : (staged_in ? s_act + lane * sp.A : actions + io * sp.A);
                GlobalActions<Real, Io> ar{arow, actions == nullptr, in.key, (unsigned long long)(in.first_env + env), episode, t,
                                           Philox4{0, 0, 0, 0}, -1};
                DynLeadtimes<Real> lr{sp, in, n, env, episode, t};
                DynDemand<Real> dr{sp, in, n, env, episode, t};
                Real cost;
                if constexpr (INFO) cost = sc_step_nodes<Real>(tp, pp, t, ar, lr, dr, ginf);
                else cost = sc_step_nodes<Real>(tp, pp, t, ar, lr, dr);
                const Real r = -cost;
                ep_ret += (double)r;
                const bool fin = (t == sp.T);
                if (fin) {
                    rs.add(ep_ret);
                    if (in.auto_reset) {
                        ++episode; t = 0; ep_ret = 0.0;
                        pp.reset(tp, sp);
                        if constexpr (INFO) ginf.clear();
                        dr.episode = episode; dr.t = 0;
                    }
                }
                sc_write_obs<Real, MODE>(tp, sp, pp, t, dr, ow);
                reward[io] = (Io)r; done[io] = fin ? 1 : 0;
            }
        }
        if (staged_out) {
            __syncwarp();
            if (n_here > 0) warp_copy<Io>(obs + io0 * sp.O, s_obs, n_here * sp.O, lane);
            __syncwarp();                  // the tile is rewritten by the next period's sc_write_obs
        }
    }
    if (active) {
        st.t[env] = t; st.episode[env] = (int)episode; st.ep_ret[env] = ep_ret;
        if (use_smem) {
            for (int i = 0; i < NP; ++i) st.stock[(long long)i * n + env] = p_stock[(size_t)i * bd];
            if constexpr (MODE == 1) {
                for (int i = 0; i < NP; ++i) {
                    const int len = p_len[(size_t)i * bd];
                    st.hlen[(long long)i * n + env] = len;
                    for (int k = 0; k < len; ++k) {
                        st.htime[((long long)i * sp.heapcap + k) * n + env] = p_time[((size_t)i * sp.heapcap + k) * bd];
                        st.hamt[((long long)i * sp.heapcap + k) * n + env] = p_amt[((size_t)i * sp.heapcap + k) * bd];
                    }
                }
            } else {
                for (int i = 0; i < NP * sp.ring; ++i) st.slot[(long long)i * n + env] = p_slot[(size_t)i * bd];
            }
        }
    }
    rs.flush(st.stats);
}

}  // namespace scg
