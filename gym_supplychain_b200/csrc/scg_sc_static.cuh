// scg_sc_static.cuh -- compile-time "layered chain" kernels (FAST modes).
//
// For S suppliers -> F factories -> W wholesalers -> R retailers, full bipartite between
// consecutive echelons, P products, constant lead time L (the reference's 2perstage /
// NPerStage / linear presets: supplychain_2perstage_env.py:27-53, supplychain_Nperstage_env.py:4-35),
// every loop of sc_step_nodes unrolls and the whole env state (N*P stocks + N*P*L ring slots)
// lives in registers for the whole T-step rollout.  Per-node constants come from the kernel
// parameter bank (constant memory, immediate operands).
//
// HBM traffic per env-step is just the I/O: the action row in, obs row + reward + done out.
// Rows are env-major in global memory ([n_envs, A] / [n_envs, O]); a warp owns 32 consecutive
// envs, i.e. one contiguous 32*A (32*O) element chunk per period, which it moves with 16-byte
// coalesced accesses through a per-warp shared-memory tile; each lane then reads (writes) its
// own row from (to) the tile with vector accesses chosen so that they are bank-conflict free.
#pragma once
#include "scg_sc_device.cuh"

namespace scg {

#ifndef SCG_SYNC_PERIOD
#define SCG_SYNC_PERIOD 0       // experiment: a block barrier per period (shared instruction fetches for loops beyond the I-cache)
#endif
#ifndef SCG_FIXED_RING
#define SCG_FIXED_RING 1        // mixed mode, constant lead time: in-transit material as 32-bit fixed point (RingFixed)
#endif

// L = avg_leadtime (constant lead time when !STOCH; number of obs buckets), ML = max_leadtime = ring depth.
template <class Real, int S, int F, int W, int R, int P, int L, int ML, bool STOCH>
struct LayeredConsts {
    static constexpr int N = S + F + W + R, NP = N * P, E = S * F + F * W + W * R;
    Real cap[NP], hold[NP], stock0[NP], max_ship[NP], init_slot[NP * ML];
    Real q_scale[NP], q_inv[NP];                         // fixed-point ring (RingFixed): slot = amount * q_scale, a power of two
    unsigned long long lt_thr[ML];                       // stochastic lead times: lt = 1 + #{j < ML-1 : word >= lt_thr[j]}
    float obs_stock[NP], obs_mid[NP], obs_last[NP];      // 1/cap, 1/max_ship, 1/(max_ship*(maxL-avgL+1)) (non-exact obs path;
                                                         // the last two times q_inv where the ring is fixed-point)
    Real sup_cap[S * P], sup_cost[S * P];
    Real proc_cap[F], ratio[F * P], inv_ratio[F * P], proc_cost[F * P];
    Real ecap[E], ecost[E * P];
    Real c_unmet, c_xstock, c_xproc, c_xship;
    Real dlo[P], dspan[P];
    float obs_dem[P];                                    // 1/span
    float obs_time;                                      // 1/T
    int dlo_i[P], dcount[P];                             // uniform generator: lo, hi-lo+1
    int all_uniform;
    int T;
};

}  // namespace scg
#ifdef SCG_BAKED_CONSTS_INC     // run-time instantiation with the spec's constants baked in: defines scg::kBaked
#include SCG_BAKED_CONSTS_INC
#endif
namespace scg {

// EXACT = true keeps IEEE divisions (fp64 I/O modes: obs/rewards comparable bit-for-bit with the generic
// fp64 kernel); false multiplies by host-computed reciprocals (fp32 and mixed modes).
template <class Real, bool EXACT, int S, int F, int W, int R, int P_, int L_, int ML_, bool STOCH_>
struct LayeredTopo {
    using C = LayeredConsts<Real, S, F, W, R, P_, L_, ML_, STOCH_>;
    static constexpr bool kStatic = true;
    static constexpr int kN = S + F + W + R, kP = P_, kL = L_, kML = ML_;
    static constexpr int kK = S * (P_ + F) + F * W + W * R;     // lead-time values per period (:597-605)
    static constexpr int kMaxD = (F > W ? (F > R ? F : R) : (W > R ? W : R));
    static constexpr int kA = S * (P_ + P_ * F) + F * P_ * W + W * P_ * R;
    static constexpr int kO = R * P_ + kN * P_ * (1 + L_) + 1;   // (:617-621)
    static constexpr int kRP = R * P_;
    const C& c;
    __device__ __forceinline__ explicit LayeredTopo(const C& cc) : c(cc) {}
    static __device__ __forceinline__ constexpr int ech(int n) { return n < S ? 0 : (n < S + F ? 1 : (n < S + F + W ? 2 : 3)); }
    static __device__ __forceinline__ constexpr int nbase(int e) { return e == 0 ? 0 : (e == 1 ? S : (e == 2 ? S + F : S + F + W)); }
    static __device__ __forceinline__ constexpr int loc(int n) { return n - nbase(ech(n)); }
    static __device__ __forceinline__ constexpr int cnt(int e) { return e == 0 ? S : (e == 1 ? F : (e == 2 ? W : R)); }
    __device__ __forceinline__ constexpr int N() const { return kN; }
    __device__ __forceinline__ constexpr int P() const { return kP; }
    __device__ __forceinline__ constexpr int avgL() const { return kL; }
    __device__ __forceinline__ constexpr int maxL() const { return kML; }
    __device__ __forceinline__ constexpr int ring() const { return kML; }
    __device__ __forceinline__ int T() const { return c.T; }
    __device__ __forceinline__ constexpr bool last(int n) const { return ech(n) == 3; }
    __device__ __forceinline__ constexpr bool supplier(int n) const { return ech(n) == 0; }
    __device__ __forceinline__ constexpr bool is_factory(int n) const { return ech(n) == 1; }   // processing capacity > 0 there only
    __device__ __forceinline__ constexpr int ndest(int n) const { return ech(n) == 3 ? 0 : cnt(ech(n) + 1); }
    __device__ __forceinline__ constexpr int ebeg(int n) const {
        return ech(n) == 0 ? loc(n) * F : (ech(n) == 1 ? S * F + loc(n) * W : S * F + F * W + loc(n) * R);
    }
    __device__ __forceinline__ constexpr int edst(int e) const {
        return e < S * F ? S + e % F : (e < S * F + F * W ? S + F + (e - S * F) % W : S + F + W + (e - S * F - F * W) % R);
    }
    __device__ __forceinline__ constexpr int dest_np(int e0, int i, int p) const { return edst(e0 + i) * kP + p; }
    __device__ __forceinline__ Real ecap(int e) const { return c.ecap[e]; }
    __device__ __forceinline__ Real ecost(int e, int p) const { return c.ecost[e * kP + p]; }
    __device__ __forceinline__ Real proc_cap(int n) const { return ech(n) == 1 ? c.proc_cap[loc(n)] : (Real)0; }
    __device__ __forceinline__ Real cap(int n, int p) const { return c.cap[n * kP + p]; }
    __device__ __forceinline__ Real hold(int n, int p) const { return c.hold[n * kP + p]; }
    __device__ __forceinline__ Real sup_cap(int n, int p) const { return c.sup_cap[loc(n) * kP + p]; }
    __device__ __forceinline__ Real sup_cost(int n, int p) const { return c.sup_cost[loc(n) * kP + p]; }
    __device__ __forceinline__ Real ratio(int n, int p) const { return ech(n) == 1 ? c.ratio[loc(n) * kP + p] : (Real)1; }
    __device__ __forceinline__ Real proc_cost(int n, int p) const { return ech(n) == 1 ? c.proc_cost[loc(n) * kP + p] : (Real)0; }
    __device__ __forceinline__ Real max_ship(int n, int p) const { return c.max_ship[n * kP + p]; }
    __device__ __forceinline__ constexpr int act_off(int n) const {
        return ech(n) == 0 ? loc(n) * (kP + kP * F)
                           : (ech(n) == 1 ? S * (kP + kP * F) + loc(n) * kP * W
                                          : (ech(n) == 2 ? S * (kP + kP * F) + F * kP * W + loc(n) * kP * R : kA));
    }
    __device__ __forceinline__ constexpr int lt_off(int n) const {       // n_supply + n_ship/P values per node (:719-722)
        return ech(n) == 0 ? loc(n) * (kP + F)
                           : (ech(n) == 1 ? S * (kP + F) + loc(n) * W : (ech(n) == 2 ? S * (kP + F) + F * W + loc(n) * R : kK));
    }
    __device__ __forceinline__ constexpr int ret_idx(int n) const { return loc(n); }
    __device__ __forceinline__ Real c_unmet() const { return c.c_unmet; }
    __device__ __forceinline__ Real c_xstock() const { return c.c_xstock; }
    __device__ __forceinline__ Real c_xproc() const { return c.c_xproc; }
    __device__ __forceinline__ Real c_xship() const { return c.c_xship; }

    __device__ __forceinline__ constexpr bool has_supply(int, int) const { return true; }
    __device__ __forceinline__ constexpr bool can_ship(int, int) const { return true; }
    __device__ __forceinline__ Real div_ratio(Real x, int n, int p) const {
        if constexpr (EXACT) return x / c.ratio[loc(n) * kP + p];
        else return x * c.inv_ratio[loc(n) * kP + p];
    }
    // D folds to a constant after unrolling, so exactly one branch survives
    template <class RawT>
    __device__ __forceinline__ void ship_cuts(const RawT* rv, int D, Real limit, Real* out) const;
};

// SHIP cuts for a compile-time D, registers only.
//
// Reference semantics (supplychain_env.py:72-94): walk the destinations in ascending (value, index) order;
// destination i gets (v_i - prev)*limit where prev is the value of its predecessor in that order (0 for the
// first), never more than what is left of `limit`.  In exact arithmetic the "never more than what is left"
// clamp is the same as capping every cut point at 1, so with
//     before(j, i) := v_j < v_i or (v_j == v_i and j < i),   prev_i := min(v_j, 1) of i's predecessor, or 0
// the amount is (min(v_i, 1) - prev_i) * limit -- no sort, no scatter.  v = (a+1)/2 is monotone in the raw
// action a, so the order predicates and the cap are evaluated on the RAW actions (fp32 compares even when
// the arithmetic type is fp64) and only the final subtraction and product run in Real.
// (The fp64-exact mode does not use this: it runs the reference's sequential loop, ship_cuts_dyn.)
// The predecessor itself is chosen among the RAW (capped) actions -- fp32 selects when the actions are fp32 -- with
// "no predecessor" encoded as the raw action -1 (whose cut point is 0):
//     amount_i = (capped_i - capped_pred(i)) * (limit / 2)
// The difference of two fp32 values is exact in fp64 (up to 2^-53 when their exponents are > 29 apart), so this is
// at least as accurate as subtracting two rounded cut points.
template <class Real, class RawT, int D>
__device__ __forceinline__ void ship_cuts_static(const RawT* r, Real limit, Real* out) {
    RawT cpd[D], pr[D];
#pragma unroll
    for (int i = 0; i < D; ++i) cpd[i] = r[i] < (RawT)1 ? r[i] : (RawT)1;
    if constexpr (D == 1) {
        pr[0] = (RawT)-1;
    } else if constexpr (D == 2) {
        const bool c01 = r[0] <= r[1];
        pr[0] = c01 ? (RawT)-1 : cpd[1];
        pr[1] = c01 ? cpd[0] : (RawT)-1;
    } else if constexpr (D == 3) {
        const bool c01 = r[0] <= r[1], c02 = r[0] <= r[2], c12 = r[1] <= r[2];
        if constexpr (sizeof(RawT) == 4) {
            // only the VALUE of the predecessor matters: with both others before i it is the larger of the two
            // (one FMNMX), so three selects and no predicate logic.  "None" must stay a separate case: the start
            // cut (raw -1) precedes every destination even when an out-of-range action lies below it.
            auto prev_of = [&](bool bj, bool bk, RawT wj, RawT wk) {
                const RawT both = fmaxf(wj, wk);
                return bj ? (bk ? both : wj) : (bk ? wk : (RawT)-1);
            };
            pr[0] = prev_of(!c01, !c02, cpd[1], cpd[2]);
            pr[1] = prev_of(c01, !c12, cpd[0], cpd[2]);
            pr[2] = prev_of(c02, c12, cpd[0], cpd[1]);
        } else {                                                             // no fp64 max instruction on sm_100
            auto prev_of = [&](bool bj, bool bk, bool bjk, RawT wj, RawT wk) {
                const RawT cand = (bk && (!bj || bjk)) ? wk : wj;
                return (bj || bk) ? cand : (RawT)-1;
            };
            pr[0] = prev_of(!c01, !c02, c12, cpd[1], cpd[2]);
            pr[1] = prev_of(c01, !c12, c02, cpd[0], cpd[2]);
            pr[2] = prev_of(c02, c12, c01, cpd[0], cpd[1]);
        }
    } else if constexpr (sizeof(RawT) == 4 && D <= 4 && sizeof(Real) == 8) {
        // D = 4, fp32 actions, fp64 state: the predecessor's value is the largest capped action of those before i (a compare,
        // a select and an FMNMX per ordered pair); "nobody before i" stays a separate case, as above.  (The pairwise form
        // below keeps D running maxima alive at once: measured to spill at D = 4 in the fp64-state kernels -- 4-4-4-4 7.0 ms
        // against 8.7 -- while the fp32 kernels prefer it, 5.2 ms against 5.6.)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const RawT none = __int_as_float((int)0xff800000u);          // -inf
            RawT m = none;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                if (j == i) continue;
                const bool before = (j < i) ? (r[j] <= r[i]) : (r[j] < r[i]);
                m = fmaxf(m, before ? cpd[j] : none);
            }
            pr[i] = m == none ? (RawT)-1 : m;
        }
    } else if constexpr (sizeof(RawT) == 4) {
        // larger D, fp32 actions: the predecessor's value is the largest capped action of those before i.  ONE compare per
        // unordered pair {a < b} decides both directions (a before b, ties to the lower index -- else b before a), each
        // side then takes a conditional max: 3 instructions per pair instead of 6.  "Nobody before i" stays a separate
        // case (-inf sentinel), as above.
        const RawT none = __int_as_float((int)0xff800000u);              // -inf
        RawT m[D];
#pragma unroll
        for (int i = 0; i < D; ++i) m[i] = none;
        // (rectangular loops with a folded guard: a triangular `b = a + 1` nest is not always fully unrolled, and one
        // surviving run-time index demotes the whole env state to local memory)
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int b = 0; b < D; ++b)
                if (b > a) {
                    const bool ab = r[a] <= r[b];
                    m[b] = ab ? fmaxf(m[b], cpd[a]) : m[b];
                    m[a] = ab ? m[a] : fmaxf(m[a], cpd[b]);
                }
#pragma unroll
        for (int i = 0; i < D; ++i) pr[i] = m[i] == none ? (RawT)-1 : m[i];
    } else {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            RawT prev = (RawT)-1, rprev = (RawT)0;
            bool has = false;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                if (j == i) continue;
                const bool before = (j < i) ? (r[j] <= r[i]) : (r[j] < r[i]);
                const bool later = !has || (j > 0 && (rprev <= r[j]));
                const bool take = before && later;
                prev = take ? cpd[j] : prev;
                rprev = take ? r[j] : rprev;
                has = has || before;
            }
            pr[i] = prev;
        }
    }
    const Real half = limit * (Real)0.5;
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] = ((Real)cpd[i] - (Real)pr[i]) * half;
}

template <class Real, bool EXACT, int S, int F, int W, int R, int P_, int L_, int ML_, bool STOCH_>
template <class RawT>
__device__ __forceinline__ void LayeredTopo<Real, EXACT, S, F, W, R, P_, L_, ML_, STOCH_>::ship_cuts(const RawT* rv, int D, Real limit, Real* out) const {
    if (D == F) ship_cuts_static<Real, RawT, F>(rv, limit, out);
    else if (D == W) ship_cuts_static<Real, RawT, W>(rv, limit, out);
    else ship_cuts_static<Real, RawT, R>(rv, limit, out);
}

// ---------------------------------------------------------------------------------
// register-resident state
// ---------------------------------------------------------------------------------
template <class Real, int NP, int L, bool STOCH>
struct RingRegs {
    static constexpr bool kTwoPhase = true;
    static constexpr bool kExactObs = false;
    static constexpr bool kFixed = false;
    Real stock[NP];
    Real slot[NP * L];
    const Real* onehot = nullptr;          // STOCH: L x L identity in shared memory, row lt-1 = where a lead time lt lands
    __device__ __forceinline__ RingRegs(const Real*, const Real*) {}
    __device__ __forceinline__ void commit() {}
    __device__ __forceinline__ void load_slot(int i, Real v) { slot[i] = v; }
    __device__ __forceinline__ Real store_slot(int i) const { return slot[i]; }
    __device__ __forceinline__ float obs_slot(int i) const { return (float)slot[i]; }
    __device__ __forceinline__ Real get_stock(int np) const { return stock[np]; }
    __device__ __forceinline__ void set_stock(int np, Real v) { stock[np] = v; }
    __device__ __forceinline__ Real slot_at(int np, int k) const { return slot[np * L + k]; }
    __device__ __forceinline__ Real arrive(int np, int) {
        const Real a = slot[np * L];
#pragma unroll
        for (int k = 0; k + 1 < L; ++k) slot[np * L + k] = slot[np * L + k + 1];
        slot[np * L + L - 1] = (Real)0;
        return a;
    }
    __device__ __forceinline__ void push(int np, int, int lt, Real amount) {
        if constexpr (!STOCH) slot[np * L + L - 1] += amount;              // constant lead time == ring depth
        else {
            // data-dependent slot: one vector load of the one-hot row and an FMA per slot (amount * 1 or amount * 0 -- the same
            // values as a compare and a select per slot, a third of the instructions; amounts are finite)
            const Real* row = onehot + (lt - 1) * L;
#pragma unroll
            for (int k = 0; k < L; ++k) slot[np * L + k] = fma(amount, row[k], slot[np * L + k]);
        }
    }
    __device__ __forceinline__ void push_own(int np, int t, int lt, Real amount) { push(np, t, lt, amount); }
};

// Fixed-point lead-time ring for the fp64-state / fp32-I/O kernels with a constant lead time (SCG_MODE_MIXED).
//
// Stocks integrate over the whole episode and stay fp64.  In-transit material does not integrate: a slot is written in
// one period (the sum of what its <= n_src sources push) and consumed L periods later, so it only has to be held to a
// fixed ABSOLUTE resolution.  Each slot is a 32-bit unsigned integer in units of q_inv[np] = 2^-k, k chosen on the host
// from the largest amount that can be in flight to that (node, product) (sc_fill_consts: q = bound * 2^-31, e.g.
// 2.8e-6 for an ntom retailer -- fp32 would resolve 3e-5 there).  That halves the registers of the ring (48 instead of
// 96 for ntom -> the kernel fits 10-11 warps per SM instead of 8), the per-period shift becomes one move per slot, and
// the observation reads the slot with an ALU int->float conversion instead of an fp64->fp32 one.  This period's pushes
// are accumulated in fp64 (`acc`, live only between the first and the last source of a destination) and rounded ONCE
// in commit(); conversions saturate, so an out-of-contract supply action far outside [-1, 1] clips at 8x capacity.
// uint <-> double through the 2^52 trick: (2^52 + u) is exactly the double whose low word is u, so one DADD / DFMA on the FP64
// pipe replaces an I2F / F2I on the quarter-rate conversion pipe (the busiest pipe of the fp64-state kernels).  The to-integer
// direction rounds to nearest-even like cvt.rni but WRAPS instead of saturating: used only where the amount is bounded by a
// capped stock (shipments); a supplier's own order (bounded only by the action) keeps the saturating conversion.
__device__ __forceinline__ double u2d(unsigned u) { return __hiloint2double(0x43300000, (int)u) - 4503599627370496.0; }
__device__ __forceinline__ unsigned d2u_wrap(double x, double scale) { return (unsigned)__double2loint(fma(x, scale, 4503599627370496.0)); }

template <int NP, int L, int NSUP = NP>      // NSUP: the first NSUP (node, product) pairs belong to suppliers
struct RingFixed {
    static constexpr bool kTwoPhase = true;
    static constexpr bool kExactObs = false;
    static constexpr bool kFixed = true;
    double stock[NP];
    unsigned slot[NP * L];
    double acc[NP];
    bool has[NP];                      // compile-time after unrolling: the first push of a period is an assignment
    const double* q_scale; const double* q_inv;
    __device__ __forceinline__ RingFixed(const double* qs, const double* qi) : q_scale(qs), q_inv(qi) {
#pragma unroll
        for (int i = 0; i < NP; ++i) { acc[i] = 0.0; has[i] = false; }
    }
    __device__ __forceinline__ double get_stock(int np) const { return stock[np]; }
    __device__ __forceinline__ void set_stock(int np, double v) { stock[np] = v; }
    // measured (profiles/r2_ring_conversion_2p52_ab.log, bit-identical results): ntom +1.1 %, 2-2-2-2 / 2 products +0.5 %, but the
    // small chains, which do not wait for the conversion pipe, lose (linear -1.2 %, 1-1-1-3 -0.3 %) -> from 8 (node, product) pairs on;
    // and a chain at the edge of the register file spills more with it (2-3-2-4, 3 products, lead time 3: 165 state registers,
    // 0.57 -> 0.48, profiles/r2_midsize_chains_v5.log) -> only while stocks + ring stay within 150 registers
    static constexpr bool kMagic = NP >= 8 && NP * (2 + L) <= 150;
    __device__ __forceinline__ double arrive(int np, int) {
        const double a = (kMagic ? u2d(slot[np * L]) : (double)slot[np * L]) * q_inv[np];
#pragma unroll
        for (int k = 0; k + 1 < L; ++k) slot[np * L + k] = slot[np * L + k + 1];
        has[np] = false;
        return a;
    }
    __device__ __forceinline__ void push(int np, int, int, double amount) {
        acc[np] = has[np] ? acc[np] + amount : amount;
        has[np] = true;
    }
    __device__ __forceinline__ void push_own(int np, int t, int lt, double amount) { push(np, t, lt, amount); }
    __device__ __forceinline__ void commit() {
#pragma unroll
        for (int np = 0; np < NP; ++np)
            slot[np * L + L - 1] = has[np] ? ((!kMagic || np < NSUP) ? __double2uint_rn(acc[np] * q_scale[np]) : d2u_wrap(acc[np], q_scale[np])) : 0u;
    }
    __device__ __forceinline__ void load_slot(int i, double v) { slot[i] = __double2uint_rn(v * q_scale[i / L]); }
    __device__ __forceinline__ double store_slot(int i) const { return (double)slot[i] * q_inv[i / L]; }
    __device__ __forceinline__ float obs_slot(int i) const { return (float)slot[i]; }     // host folds q_inv into obs_mid / obs_last
};

// The same fixed-point ring for STOCHASTIC lead times (L = max lead time slots).  A push lands in a data-dependent slot, so it
// cannot wait for a per-period commit: each amount is rounded to the (node, product) quantum when it is pushed and booked with one
// integer multiply-add per slot against a row of an L x L 0/1 table in shared memory (the integer twin of RingRegs::push).  The
// host leaves log2(L) bits of headroom in the scale (sc_ring_scales): a slot can collect the pushes of up to L periods.
template <int NP, int L, int NSUP = NP>
struct RingFixedStoch {
    static constexpr bool kTwoPhase = true;
    static constexpr bool kExactObs = false;
    static constexpr bool kFixed = true;
    double stock[NP];
    unsigned slot[NP * L];
    const double* q_scale; const double* q_inv;
    const unsigned* onehot = nullptr;  // L x L identity in shared memory, row lt-1 = where a lead time lt lands
    __device__ __forceinline__ RingFixedStoch(const double* qs, const double* qi) : q_scale(qs), q_inv(qi) {}
    __device__ __forceinline__ void commit() {}
    __device__ __forceinline__ double get_stock(int np) const { return stock[np]; }
    __device__ __forceinline__ void set_stock(int np, double v) { stock[np] = v; }
    __device__ __forceinline__ double arrive(int np, int) {
        const double a = (double)slot[np * L] * q_inv[np];
#pragma unroll
        for (int k = 0; k + 1 < L; ++k) slot[np * L + k] = slot[np * L + k + 1];
        slot[np * L + L - 1] = 0u;
        return a;
    }
    __device__ __forceinline__ void push(int np, int, int lt, double amount) {
        const unsigned q = __double2uint_rn(amount * q_scale[np]);
        const unsigned* row = onehot + (lt - 1) * L;
#pragma unroll
        for (int k = 0; k < L; ++k) slot[np * L + k] += q * row[k];
    }
    __device__ __forceinline__ void push_own(int np, int t, int lt, double amount) { push(np, t, lt, amount); }
    __device__ __forceinline__ void load_slot(int i, double v) { slot[i] = __double2uint_rn(v * q_scale[i / L]); }
    __device__ __forceinline__ double store_slot(int i) const { return (double)slot[i] * q_inv[i / L]; }
    __device__ __forceinline__ float obs_slot(int i) const { return (float)slot[i]; }
};

template <bool C, class A, class B> struct TypeIf { using type = A; };           // (no <type_traits> under NVRTC)
template <class A, class B> struct TypeIf<false, A, B> { using type = B; };
#ifndef SCG_FIXED_RING_STOCH
#define SCG_FIXED_RING_STOCH 1  // mixed mode, stochastic lead times: RingFixedStoch (0 = the fp64 ring)
#endif
template <bool FIXED, class Real, int NP, int L, bool STOCH, int NSUP = NP> struct PipeChoice { using type = RingRegs<Real, NP, L, STOCH>; };
template <int NP, int L, int NSUP> struct PipeChoice<true, double, NP, L, false, NSUP> { using type = RingFixed<NP, L, NSUP>; };
template <int NP, int L, int NSUP> struct PipeChoice<true, double, NP, L, true, NSUP> { using type = RingFixedStoch<NP, L, NSUP>; };

// Reads one env's row of LEN elements out of the warp tile, 16/8/4 bytes at a time.  All
// indices are compile-time constants after unrolling, so the chunk cache folds away.
template <class Real, int LEN>
struct RowReader {
    static constexpr int BYTES = LEN * (int)sizeof(Real);
    static constexpr int VB = (BYTES % 16 == 0) ? 16 : ((BYTES % 8 == 0) ? 8 : (int)sizeof(Real));
    static constexpr int VW = VB / (int)sizeof(Real);     // elements per vector access
    const Real* row;
    Real cur[VW];
    int cur_chunk;
    __device__ __forceinline__ explicit RowReader(const Real* r) : row(r), cur_chunk(-1) {}
    __device__ __forceinline__ void load(int chunk) {
        if constexpr (VB == 16 && sizeof(Real) == 4) {
            const float4 q = reinterpret_cast<const float4*>(row)[chunk];
            cur[0] = q.x; cur[1] = q.y; cur[2] = q.z; cur[3] = q.w;
        } else if constexpr (VB == 16 && sizeof(Real) == 8) {
            const double2 d = reinterpret_cast<const double2*>(row)[chunk];
            cur[0] = d.x; cur[1] = d.y;
        } else if constexpr (VB == 8 && sizeof(Real) == 4) {
            const float2 q = reinterpret_cast<const float2*>(row)[chunk]; cur[0] = q.x; cur[1] = q.y;
        } else {
            cur[0] = row[chunk];
        }
        cur_chunk = chunk;
    }
    __device__ __forceinline__ Real get(int idx) {
        const int chunk = idx / VW;
        if (chunk != cur_chunk) load(chunk);
        Real r = cur[0];
#pragma unroll
        for (int j = 1; j < VW; ++j) r = (idx % VW == j) ? cur[j] : r;
        return r;
    }
};

template <class Real, class Io, bool RANDOM, int A>
struct StaticActions {
    RowReader<Io, A> rd;
    RngKey key; unsigned long long genv; unsigned episode; int t;
    Philox4 blk; int blk_id;
    __device__ __forceinline__ StaticActions(const Io* row, RngKey k, unsigned long long g, unsigned ep, int tt)
        : rd(row), key(k), genv(g), episode(ep), t(tt), blk_id(-1) {}
    using raw_t = Io;
    __device__ __forceinline__ Io get_raw(int idx) {
        if constexpr (!RANDOM) return rd.get(idx);
        else {
            const int b = idx >> 2;
            if (b != blk_id) { blk = rng_block(key, genv, episode, (unsigned)t, STREAM_ACTION, (unsigned)b); blk_id = b; }
            return (Io)action_from_word(pick_word(blk, idx & 3));
        }
    }
    __device__ __forceinline__ Real get(int idx) { return (Real)get_raw(idx); }
};

template <int L> struct ConstLeadtime {
    static constexpr bool kConst = true;        // no data-dependent lead-time cursor (F4b) to maintain
    __device__ __forceinline__ constexpr int get(int) const { return L; }
};

// K lead times of one period packed 4 bits each (value - 1), so that the data-dependent cursor of
// SC_Node.act (:252-254, F4b) can index them at run time without touching local memory.
template <int K> struct PackedLeadtimes {
    static constexpr bool kConst = false;
    static constexpr int NW = (K + 15) / 16;
    unsigned long long w[NW];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = 0ull;
    }
    __device__ __forceinline__ void set(int k, int lt) { w[k >> 4] |= (unsigned long long)(lt - 1) << ((k & 15) * 4); }   // k static
    __device__ __forceinline__ int get(int k) const {
        unsigned long long v = w[0];
#pragma unroll
        for (int i = 1; i < NW; ++i) v = ((k >> 4) == i) ? w[i] : v;
        return (int)((v >> ((k & 15) * 4)) & 15ull) + 1;
    }
};

template <int RP> struct RegDemand {
    int cur_[RP], next_[RP];
    __device__ __forceinline__ int cur(int rp, int) const { return cur_[rp]; }
    __device__ __forceinline__ int next(int rp, int) const { return next_[rp]; }
};

template <class Real, class Io> struct SmemObs {
    Io* row;
    template <class V> __device__ __forceinline__ void put(int idx, V v) { row[idx] = (Io)v; }
};

// ---------------------------------------------------------------------------------
// Blackwell/Hopper bulk async copies (TMA, 1-D): one elected lane moves a whole warp tile.
//   global -> shared : cp.async.bulk + mbarrier complete_tx   (SASS UBLKCP)
//   shared -> global : cp.async.bulk ... bulk_group
//   L2 prefetch      : cp.async.bulk.prefetch.L2
// ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// One lane of a converged warp (elect.sync): the guard of the bulk-copy issue.  Unlike `lane == 0` the predicate is known to
// the compiler as "exactly one lane", so it does not wrap the uniform-datapath copy instruction in a loop over the active lanes.
__device__ __forceinline__ bool elect_one() {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok));
    return ok != 0;
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* g, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}

// Action tiles per warp.  The fp64-state kernels with fp32 I/O are register-bound at 8-12 warps per SM, which leaves shared
// memory for TWO action tiles: the bulk copy of tile s+1 is issued at the top of period s, a whole period ahead, instead of
// after the step with an L2 prefetch to hide its latency (the prefetch instruction alone held the warp for 2.8 % of the
// headline kernel's time, profiles/r2_ncu_ntom_mixed_v12_summary.txt).  SCG_DBUF_ACTIONS=0 builds the single-tile form.
#ifndef SCG_DBUF_ACTIONS
#define SCG_DBUF_ACTIONS 1
#endif
#ifndef SCG_G2S_AT_TOP
#define SCG_G2S_AT_TOP 0        // two action tiles: 1 = request tile s+1 at the top of period s; 0 = tile s+2 at its end, with the obs store
#endif
// Interleaved A/B (profiles/r2_action_tile_double_buffer_ab.log): ntom +3.1 %, supplychain-v0 +1.8 %, 2-2-2-2 with two
// products +1 %; the chains with tiny tiles lose (linear -8 %, 1-1-1-3 -1 %), hence the bound on the (node, product) count.
template <class Real, class Io, int NP>
__host__ __device__ constexpr int layered_action_tiles() {
    return (SCG_DBUF_ACTIONS != 0 && sizeof(Real) == 8 && sizeof(Io) == 4 && NP >= 8) ? 2 : 1;
}

// Bytes in front of the per-warp tiles: the mbarriers and, for stochastic lead times, the one-hot table of RingRegs::push.
template <class Real, int ML, bool STOCH>
__host__ __device__ constexpr int layered_smem_header() {
    return 128 + (STOCH ? (ML * ML * (int)sizeof(Real) + 127) / 128 * 128 : 0);
}

// ---------------------------------------------------------------------------------
// the fused rollout kernel
// ---------------------------------------------------------------------------------
template <class Real, class Io, int S, int F, int W, int R, int P, int L, int ML, bool STOCH, bool RANDOM_ACT, int THREADS, int MINB>
#ifdef SCG_MAXNREG
__global__ void __maxnreg__(SCG_MAXNREG)
#else
__global__ void __launch_bounds__(THREADS, MINB)
#endif
sc_layered_rollout(const __grid_constant__ LayeredConsts<Real, S, F, W, R, P, L, ML, STOCH> c_arg, ScDevSpec<Real> sp, ScState<Real> st,
                   ScInputs in, int n_steps, const Io* __restrict__ actions, Io* __restrict__ obs,
                   Io* __restrict__ reward, unsigned char* __restrict__ done) {
    constexpr bool kExact = sizeof(Io) == 8;
#ifdef SCG_BAKED_CONSTS_INC
    const LayeredConsts<Real, S, F, W, R, P, L, ML, STOCH>& c = kBaked;        // compile-time constants of THIS spec
#else
    const LayeredConsts<Real, S, F, W, R, P, L, ML, STOCH>& c = c_arg;         // the kernel parameter block
#endif
    using Tp = LayeredTopo<Real, kExact, S, F, W, R, P, L, ML, STOCH>;
    constexpr int K = Tp::kK;
    constexpr int N = Tp::kN, NP = N * P, A = Tp::kA, O = Tp::kO, RP = Tp::kRP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // one warp per block: every tile / barrier / bulk-copy address derives from blockIdx alone, so the compiler keeps it in
    // uniform registers (a warp index read from threadIdx costs a register-to-uniform loop around every bulk copy)
    const int lane = threadIdx.x & 31, warp = THREADS == 32 ? 0 : (int)(threadIdx.x >> 5);
    // smem: [two mbarriers per warp (one per action tile), padded to 128 B][per warp: action tile(s) | obs tile]
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw) + warp * 2;     // two per warp (one per action tile)
    // [128 B: the mbarriers][STOCH: ML x ML 0/1 table (Real, or unsigned for the fixed-point ring), padded to 128 B][per warp: action tile(s) | obs tile]
    constexpr int kHdr = layered_smem_header<Real, ML, STOCH>();
    constexpr int kActTiles = layered_action_tiles<Real, Io, NP>();
    Io* s_act0 = reinterpret_cast<Io*>(smem_raw + kHdr) + (size_t)warp * 32 * (kActTiles * A + O);
    Io* s_obs = s_act0 + kActTiles * 32 * A;
    const long long env0 = ((long long)blockIdx.x * (THREADS / 32) + warp) * 32;
    const long long env = env0 + lane;
    const long long n = st.n;
    const bool active = env < n;
    const int n_here = env0 >= n ? 0 : (int)((n - env0) < 32 ? (n - env0) : 32);
    const Tp tp(c);
    ReturnStats rs; rs.init();

    // bulk (TMA) copies need 16-byte aligned addresses and sizes; otherwise fall back to plain warp copies
    const unsigned act_bytes = (unsigned)(n_here * A * (int)sizeof(Io)), obs_bytes = (unsigned)(n_here * O * (int)sizeof(Io));
#ifndef SCG_NO_BULK
#define SCG_NO_BULK 0           // experiment: plain warp copies instead of the bulk (TMA) copies (bank-conflict attribution)
#endif
    const bool bulk_in = !SCG_NO_BULK && !RANDOM_ACT && n_here > 0 && (((size_t)actions) & 15) == 0 &&
                         (((size_t)n * A * sizeof(Io)) & 15) == 0 && (act_bytes & 15) == 0 && ((32 * A * sizeof(Io)) & 15) == 0;
    const bool bulk_out = !SCG_NO_BULK && obs != nullptr && n_here > 0 && (((size_t)obs) & 15) == 0 && (((size_t)n * O * sizeof(Io)) & 15) == 0 &&
                          (obs_bytes & 15) == 0 && ((32 * (A + O) * sizeof(Io)) & 15) == 0;
    if (lane == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_mbar_init(); }
    __syncwarp();

    // fp64 state with fp32 I/O and a constant lead time: in-transit material as 32-bit fixed point (RingFixed)
    constexpr bool kFixed = (SCG_FIXED_RING != 0) && sizeof(Real) == 8 && sizeof(Io) == 4 && (!STOCH || SCG_FIXED_RING_STOCH != 0);
    using Pp = typename PipeChoice<kFixed, Real, NP, ML, STOCH, S * P>::type;
    Pp pp(c.q_scale, c.q_inv);
    if constexpr (STOCH) {
        using Tab = typename TypeIf<kFixed, unsigned, Real>::type;                 // the header is sized for Real: enough for both
        Tab* tab = reinterpret_cast<Tab*>(smem_raw + 128);
        for (int i = threadIdx.x; i < ML * ML; i += THREADS) tab[i] = (i / ML == i % ML) ? (Tab)1 : (Tab)0;
        __syncthreads();
        pp.onehot = tab;
    }
    RegDemand<RP> dm;
    int t = 0; unsigned episode = 0; double ep_ret = 0.0;
    const unsigned long long genv = (unsigned long long)(in.first_env + env);

    auto fetch_next_demand = [&](int row) {
        if (c.all_uniform && in.inj_demand == nullptr) {
#pragma unroll
            for (int b = 0; b < (RP + 3) / 4; ++b) {
                const Philox4 blk = rng_block(in.key, genv, episode, (unsigned)row, STREAM_DEMAND, (unsigned)b);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int rp = 4 * b + j;
                    if (rp < RP) dm.next_[rp] = uniform_int(pick_word(blk, j), c.dlo_i[rp % P], (unsigned)c.dcount[rp % P]);
                }
            }
        } else {
#pragma unroll
            for (int rp = 0; rp < RP; ++rp) dm.next_[rp] = sc_demand(sp, in, n, env, episode, row, rp, rp % P);
        }
    };

    // prologue: start moving the first action tile while the state is being loaded
    if (bulk_in && n_steps > 0 && lane == 0) {
        mbar_expect_tx(bar, act_bytes);
        bulk_g2s(s_act0, actions + env0 * A, act_bytes, bar);
        if (kActTiles == 2 && !SCG_G2S_AT_TOP && n_steps > 1) {               // two tiles: tile s+2 is requested at the end of period s
            mbar_expect_tx(bar + 1, act_bytes);
            bulk_g2s(s_act0 + 32 * A, actions + (n + env0) * A, act_bytes, bar + 1);
        }
    }
    if (active) {
        t = st.t[env]; episode = (unsigned)st.episode[env]; ep_ret = st.ep_ret[env];
#pragma unroll
        for (int i = 0; i < NP; ++i) pp.stock[i] = st.stock[(long long)i * n + env];
#pragma unroll
        for (int i = 0; i < NP * ML; ++i) pp.load_slot(i, st.slot[(long long)i * n + env]);
        fetch_next_demand(t);
    }

#pragma unroll 1
    for (int s = 0; s < n_steps; ++s) {
        const long long io0 = (long long)s * n + env0;
#if SCG_SYNC_PERIOD
        __syncthreads();                   // keep the block's warps in the same stretch of the (larger than I-cache) loop body
#endif
        Io* s_act = s_act0 + ((kActTiles == 2 && bulk_in) ? (s & 1) * 32 * A : 0);
        if (bulk_in) {
            // tile s was requested one period ago (or in the prologue)
            if constexpr (kActTiles == 1) {
                if (s + 1 < n_steps && elect_one()) bulk_prefetch_l2(actions + (io0 + n) * A, act_bytes);     // ask L2 for tile s+1 now
                while (!mbar_try_wait(bar, (unsigned)(s & 1))) {}
            } else {
                // tile s sits in buffer s & 1 (requested two periods ago, at the end of period s-2, or in the prologue)
                while (!mbar_try_wait(bar + (s & 1), (unsigned)((s >> 1) & 1))) {}
#if SCG_G2S_AT_TOP
                if (s + 1 < n_steps && elect_one()) {      // every lane finished reading that buffer in period s-1
                    mbar_expect_tx(bar + ((s + 1) & 1), act_bytes);
                    bulk_g2s(s_act0 + ((s + 1) & 1) * 32 * A, actions + (io0 + n) * A, act_bytes, bar + ((s + 1) & 1));
                }
#endif
            }
        } else if (!RANDOM_ACT && n_here > 0) {
            __syncwarp();
            warp_copy<Io>(s_act, actions + io0 * A, n_here * A, lane);
            __syncwarp();
        }
        Real r = (Real)0;
        bool fin = true, stepped = false;
        if (active && t < c.T) {
            stepped = true;
            ++t;
#pragma unroll
            for (int rp = 0; rp < RP; ++rp) dm.cur_[rp] = dm.next_[rp];
            StaticActions<Real, Io, RANDOM_ACT, A> ar(s_act + lane * A, in.key, genv, episode, t);
            Real cost;
            if constexpr (!STOCH) {
                ConstLeadtime<L> lr;
                cost = sc_step_nodes<Real>(tp, pp, t, ar, lr, dm);
                pp.commit();
            } else {
                // this period's K lead times: injected table, or 1 + Poisson(avg-1) clipped to ML by integer
                // thresholds on one Philox word each (same words and thresholds as the generic kernel)
                PackedLeadtimes<K> lr;
                lr.clear();
                if (in.inj_leadtimes != nullptr) {
#pragma unroll
                    for (int k = 0; k < K; ++k) lr.set(k, checked_leadtime(in.inj_leadtimes[((long long)(t - 1) * K + k) * n + env], ML, in.err));
                } else {
#pragma unroll
                    for (int b = 0; b < (K + 3) / 4; ++b) {
                        const Philox4 blk = rng_block(in.key, genv, episode, (unsigned)t, STREAM_LEADTIME, (unsigned)b);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k = 4 * b + j;
                            if (k < K) {
                                // word >= thr with thr in [1, 2^32] ("never" = 2^32)  <=>  word > thr - 1 in 32 bits
                                const unsigned word = pick_word(blk, j);
                                int lt = 1;
#pragma unroll
                                for (int q = 0; q + 1 < ML; ++q) lt += (word > (unsigned)(c.lt_thr[q] - 1ull)) ? 1 : 0;
                                lr.set(k, lt);
                            }
                        }
                    }
                }
                cost = sc_step_nodes<Real>(tp, pp, t, ar, lr, dm);
            }
            r = -cost;
            ep_ret += (double)r;
            fin = (t == c.T);
            if (fin) {
                rs.add(ep_ret);
                if (in.auto_reset) {
                    ++episode; t = 0; ep_ret = 0.0;
#pragma unroll
                    for (int i = 0; i < NP; ++i) pp.stock[i] = c.stock0[i];
#pragma unroll
                    for (int i = 0; i < NP * ML; ++i) pp.load_slot(i, c.init_slot[i]);
                }
            }
        }
        if constexpr (kActTiles == 1) {
            __syncwarp();                      // every lane is done with the action tile of period s
            if (bulk_in && s + 1 < n_steps && elect_one()) {
                // request tile s+1 now: it lands while the observation below is computed and stored
                mbar_expect_tx(bar, act_bytes);
                bulk_g2s(s_act, actions + (io0 + n) * A, act_bytes, bar);
            }
        } else if (bulk_in && !bulk_out && obs == nullptr) {
            __syncwarp();                      // no observation output: nothing else separates the periods' tile reads
        }
        if (bulk_out) {
            // the previous obs tile has left shared memory.  Every lane waits on its OWN bulk groups (a no-op for the lanes that
            // never issued one): no divergent region around the wait
            bulk_wait_read_all();
            __syncwarp();
        }
        if (active) {
            if (stepped) fetch_next_demand(t);   // demand row t: shown in this obs, consumed by the next period
            if (obs != nullptr) {
                SmemObs<Real, Io> ow{s_obs + lane * O};
                // observation (canonical per-arrival-time sums), :428-463 / :762-791
                int o = 0;
                auto unit = [](float x, float inv) { return fmaf(__saturatef(x * inv), 2.0f, -1.0f); };   // clip(2x-1, -1, 1)
#pragma unroll
                for (int rp = 0; rp < RP; ++rp) {
                    const int p = rp % P;
                    if constexpr (kExact) ow.put(o++, obs_norm<Real>(((Real)dm.next_[rp] - c.dlo[p]) / c.dspan[p]));
                    else ow.put(o++, unit((float)(dm.next_[rp] - c.dlo_i[p]), c.obs_dem[p]));
                }
#pragma unroll
                for (int nn = 0; nn < N; ++nn) {
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int np = nn * P + p;
                        if constexpr (kExact) ow.put(o++, obs_norm<Real>(pp.stock[np] / c.cap[np]));
                        else ow.put(o++, unit((float)pp.stock[np], c.obs_stock[np]));
                    }
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int np = nn * P + p;
#pragma unroll
                        for (int k = 0; k + 1 < L; ++k) {
                            if constexpr (kExact) ow.put(o++, obs_norm<Real>(c.max_ship[np] > (Real)0 ? pp.slot[np * ML + k] / c.max_ship[np] : (Real)0));
                            else ow.put(o++, unit(pp.obs_slot(np * ML + k), c.obs_mid[np]));
                        }
                        if constexpr (kFixed) {
                            unsigned rest = 0u;                    // everything arriving at t+L or later: an exact integer sum
#pragma unroll
                            for (int k = L - 1; k < ML; ++k) rest += pp.slot[np * ML + k];
                            ow.put(o++, unit((float)rest, c.obs_last[np]));
                        } else {
                            Real rest = (Real)0;                   // everything arriving at t+L or later
#pragma unroll
                            for (int k = L - 1; k < ML; ++k) rest += pp.slot[np * ML + k];
                            if constexpr (kExact) ow.put(o++, obs_norm<Real>(c.max_ship[np] > (Real)0 ? rest / (c.max_ship[np] * (Real)(ML - (L - 1))) : (Real)0));
                            else ow.put(o++, unit((float)rest, c.obs_last[np]));
                        }
                    }
                }
                if constexpr (kExact) ow.put(o++, obs_norm<Real>((Real)(c.T - t) / (Real)c.T));
                else ow.put(o++, unit((float)(c.T - t), c.obs_time));
            }
            reward[io0 + lane] = (Io)r;
            done[io0 + lane] = fin ? 1 : 0;
        }
        const bool next_tile = kActTiles == 2 && !SCG_G2S_AT_TOP && bulk_in && s + 2 < n_steps;     // tile s+2 goes where tile s was
        if (bulk_out) {
            fence_proxy_async_smem();      // generic-proxy writes of the tile -> visible to the async proxy
            __syncwarp();
        } else if (obs != nullptr && n_here > 0) {
            __syncwarp();
            warp_copy<Io>(obs + io0 * O, s_obs, n_here * O, lane);
            __syncwarp();
        }
        // ONE single-lane region per period issues both bulk copies: the obs tile of period s out, the action tile of period
        // s+2 in (every lane is past the warp barriers above, i.e. done reading tile s)
        if (bulk_out || next_tile) {
            if (elect_one()) {
                if (bulk_out) { bulk_s2g(obs + io0 * O, s_obs, obs_bytes); bulk_commit(); }
                if (next_tile) {
                    mbar_expect_tx(bar + (s & 1), act_bytes);
                    bulk_g2s(s_act, actions + (io0 + 2 * n) * A, act_bytes, bar + (s & 1));
                }
            }
        }
    }
    if (bulk_out) bulk_wait_all();                 // shared memory must outlive the last bulk store (every lane: its own groups)

    if (active) {
        st.t[env] = t; st.episode[env] = (int)episode; st.ep_ret[env] = ep_ret;
#pragma unroll
        for (int i = 0; i < NP; ++i) st.stock[(long long)i * n + env] = pp.stock[i];
#pragma unroll
        for (int i = 0; i < NP * ML; ++i) st.slot[(long long)i * n + env] = pp.store_slot(i);
    }
    rs.flush(st.stats);
}

}  // namespace scg
