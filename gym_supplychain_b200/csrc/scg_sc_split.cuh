// scg_sc_split.cuh -- node-parallel ("split") rollout kernel for layered chains whose state does not fit one thread.
//
// The compile-time kernel of scg_sc_static.cuh maps one thread to one env; with 26 nodes x 2 products x (1 + 2) state
// values (SupplyChainNPerStage([5,4,7,10]), supplychain_Nperstage_env.py:37-66) or 40 x 2 x 3 (10 per stage) the register
// file overflows and the kernel ran at 3 % of the HBM roofline.  Here G lanes (2, 4 or 8) share one env: local node j of
// every echelon belongs to lane j % G, so a lane owns ceil(cnt/G) nodes per echelon -- their stocks and lead-time ring in
// registers, every loop unrolled -- and a warp simulates 32/G envs.  What crosses lanes:
//   * shipments: a lane adds what ITS nodes send to each node of the next echelon into per-destination partial sums
//     (static register indices), the partials are summed over the G lanes with warp shuffles (xor butterfly inside the
//     env's lane group) and the owner lane of each destination adds the total to its ring;
//   * the per-node costs: one shuffle reduction of the period's cost -> reward;
// Everything inside a node (SHIP cut points, processing capacity and ratio, ship capacity shared by the products of an
// edge, lost sales, holding cost -- supplychain_env.py:208-396) is lane-local: it is the SAME sc_step_nodes as every other
// kernel, instantiated per echelon with a Topo policy that reads the owned node's constants from global memory (L1) and
// a Pipe policy over the lane's registers.  A slot without a node (cnt not a multiple of G) is simulated as an empty
// node: zero stock, zero supply capacity, zero demand -> it pushes nothing and costs nothing, no divergence.
//
// I/O is the same per-warp shared-memory tile as the static kernel (32/G consecutive envs = one contiguous chunk of
// actions / obs per period, moved with TMA bulk copies); lanes read their nodes' actions and write their nodes' obs
// entries at run-time offsets inside the env's row.  Constant lead times, fp32 I/O (modes fast_f32 and mixed).
#pragma once
#include "scg_sc_static.cuh"

namespace scg {

template <int S, int F, int W, int R>
struct SplitShape {
    static __host__ __device__ constexpr int cnt(int e) { return e == 0 ? S : (e == 1 ? F : (e == 2 ? W : R)); }
    static __host__ __device__ constexpr int base(int e) { return e == 0 ? 0 : (e == 1 ? S : (e == 2 ? S + F : S + F + W)); }
    static __host__ __device__ constexpr int ebase(int e) { return e == 0 ? 0 : (e == 1 ? S * F : S * F + F * W); }   // first edge of the echelon
};

// Per-lane registers: the owned nodes of all four echelons, echelon-major, plus the partial sums of what this lane's nodes
// send to each node of the next echelon (one echelon at a time).  ONE plain struct of arrays, so that every element is a
// register after unrolling; the per-echelon Pipe policies below are method-only views of it.
// fp64 state (mixed mode): the ring is 32-bit fixed point exactly as in RingFixed (scg_sc_static.cuh) -- per (node, product)
// power-of-two scales from the host (ScDevSpec::q_scale / q_inv = sc_ring_scales), stocks stay fp64.  Half the ring registers
// of a kernel that otherwise sits at the 255-register limit, 32-bit shuffles for the cross-lane sums.
template <class Real> struct SplitSlot { using type = Real; static constexpr bool kFixed = false; };
#if SCG_FIXED_RING
template <> struct SplitSlot<double> { using type = unsigned; static constexpr bool kFixed = true; };
#endif

template <class Real, int KT, int P, int L, int MAXPART, int KOWN = 1>
struct SplitRegs {
    using Slot = typename SplitSlot<Real>::type;
    static constexpr bool kFixed = SplitSlot<Real>::kFixed;
    Real stock[KT * P];
    Slot slot[KT * P * L];
    Real part[MAXPART];
    bool has[MAXPART];                     // compile-time after unrolling: the first push is an assignment
    Real own[KOWN * P];                    // fixed-point ring: what the lane's suppliers ordered this period (booked after the step)
};

// Topo policy of ONE echelon as seen by one lane (sc_step_nodes iterates its K owned slots).
template <class Real, int E, int S, int F, int W, int R, int P_, int L_, int G>
struct SplitTopo {
    using Sh = SplitShape<S, F, W, R>;
    static constexpr int CNT = Sh::cnt(E), NEXT = E < 3 ? Sh::cnt(E + 1) : 0, K = (CNT + G - 1) / G;
    static constexpr int kP = P_, kL = L_;
    static constexpr bool kStatic = true;
    static constexpr int kMaxD = NEXT > 0 ? NEXT : 1;
    const ScDevSpec<Real>& sp;
    int q;                                   // this lane's index inside the env's group of G lanes
    Real inv_ratio[K * P_];
    __device__ __forceinline__ SplitTopo(const ScDevSpec<Real>& s, int q_) : sp(s), q(q_) {
#pragma unroll
        for (int l = 0; l < K; ++l)
#pragma unroll
            for (int p = 0; p < P_; ++p) inv_ratio[l * P_ + p] = E == 1 ? (Real)1 / __ldg(sp.ratio + node(l) * P_ + p) : (Real)1;
    }
    __device__ __forceinline__ bool valid(int l) const { return q + G * l < CNT; }
    __device__ __forceinline__ int local(int l) const { const int j = q + G * l; return j < CNT ? j : CNT - 1; }   // clamped: finite constants
    __device__ __forceinline__ int node(int l) const { return Sh::base(E) + local(l); }
    __device__ __forceinline__ constexpr int N() const { return K; }
    __device__ __forceinline__ constexpr int P() const { return P_; }
    __device__ __forceinline__ constexpr int avgL() const { return L_; }
    __device__ __forceinline__ constexpr int maxL() const { return L_; }
    __device__ __forceinline__ constexpr int ring() const { return L_; }
    __device__ __forceinline__ int T() const { return sp.T; }
    __device__ __forceinline__ constexpr bool last(int) const { return E == 3; }
    __device__ __forceinline__ constexpr bool supplier(int) const { return E == 0; }
    __device__ __forceinline__ constexpr bool is_factory(int) const { return E == 1; }
    __device__ __forceinline__ constexpr int ndest(int) const { return NEXT; }
    __device__ __forceinline__ int ebeg(int l) const { return Sh::ebase(E) + local(l) * NEXT; }
    __device__ __forceinline__ constexpr int dest_np(int, int i, int p) const { return i * P_ + p; }     // index into the partial sums
    __device__ __forceinline__ Real ecap(int e) const { return __ldg(sp.ecap + e); }
    __device__ __forceinline__ Real ecost(int e, int p) const { return __ldg(sp.ecost + e * P_ + p); }
    __device__ __forceinline__ Real proc_cap(int l) const { return E == 1 ? __ldg(sp.proc_cap + node(l)) : (Real)0; }
    __device__ __forceinline__ Real cap(int l, int p) const { return __ldg(sp.cap + node(l) * P_ + p); }
    __device__ __forceinline__ Real hold(int l, int p) const { return __ldg(sp.hold + node(l) * P_ + p); }
    __device__ __forceinline__ Real sup_cap(int l, int p) const { return valid(l) ? __ldg(sp.sup_cap + node(l) * P_ + p) : (Real)0; }
    __device__ __forceinline__ Real sup_cost(int l, int p) const { return __ldg(sp.sup_cost + node(l) * P_ + p); }
    __device__ __forceinline__ Real ratio(int l, int p) const { return E == 1 ? __ldg(sp.ratio + node(l) * P_ + p) : (Real)1; }
    __device__ __forceinline__ Real proc_cost(int l, int p) const { return E == 1 ? __ldg(sp.proc_cost + node(l) * P_ + p) : (Real)0; }
    __device__ __forceinline__ Real max_ship(int l, int p) const { return __ldg(sp.max_ship + node(l) * P_ + p); }
    __device__ __forceinline__ int act_off(int l) const {            // action layout of a layered chain (supplychain_env.py:714-717)
        const int j = local(l);
        return E == 0 ? j * (P_ + P_ * F) : (E == 1 ? S * (P_ + P_ * F) + j * P_ * W : S * (P_ + P_ * F) + F * P_ * W + j * P_ * R);
    }
    __device__ __forceinline__ constexpr int lt_off(int) const { return 0; }
    __device__ __forceinline__ constexpr int ret_idx(int l) const { return l; }     // the lane's own demand registers
    __device__ __forceinline__ Real c_unmet() const { return sp.c_unmet; }
    __device__ __forceinline__ Real c_xstock() const { return sp.c_xstock; }
    __device__ __forceinline__ Real c_xproc() const { return sp.c_xproc; }
    __device__ __forceinline__ Real c_xship() const { return sp.c_xship; }
    __device__ __forceinline__ constexpr bool has_supply(int, int) const { return true; }
    __device__ __forceinline__ constexpr bool can_ship(int, int) const { return true; }
    __device__ __forceinline__ Real div_ratio(Real x, int l, int p) const { return x * inv_ratio[l * P_ + p]; }
    template <class RawT>
    __device__ __forceinline__ void ship_cuts(const RawT* rv, int, Real limit, Real* out) const {
        ship_cuts_static<Real, RawT, kMaxD>(rv, limit, out);
    }
};

// Pipe policy of one echelon: a view of the lane's registers (arrivals are applied before the echelons act) -- no data
// of its own, so the kernel may look at its SplitRegs through it.
template <class Real, int OFF, int P, int L, int NEXT, int KT, int MAXPART, int KOWN>
struct SplitPipe : SplitRegs<Real, KT, P, L, MAXPART, KOWN> {
    using Base = SplitRegs<Real, KT, P, L, MAXPART, KOWN>;
    static constexpr bool kTwoPhase = false;
    static constexpr bool kExactObs = false;
    static constexpr int NPART = NEXT * P;
    __device__ __forceinline__ void begin() {
#pragma unroll
        for (int i = 0; i < MAXPART; ++i) { this->part[i] = (Real)0; this->has[i] = false; }
        if constexpr (Base::kFixed && OFF == 0) {
#pragma unroll
            for (int i = 0; i < KOWN * P; ++i) this->own[i] = (Real)0;
        }
    }
    __device__ __forceinline__ Real get_stock(int np) const { return this->stock[OFF * P + np]; }
    __device__ __forceinline__ void set_stock(int np, Real v) { this->stock[OFF * P + np] = v; }
    __device__ __forceinline__ Real arrive(int, int) const { return (Real)0; }
    __device__ __forceinline__ void push_own(int np, int, int, Real amount) {
        if constexpr (Base::kFixed) this->own[np] = amount;          // suppliers only (OFF == 0): np < KOWN * P
        else this->slot[(OFF * P + np) * L + L - 1] += amount;
    }
    __device__ __forceinline__ void push(int idx, int, int, Real amount) {
        this->part[idx] = this->has[idx] ? this->part[idx] + amount : amount;
        this->has[idx] = true;
    }
};

template <int G, class Real>
__device__ __forceinline__ Real group_sum(Real v) {          // the G lanes of an env are consecutive: xor butterfly
#pragma unroll
    for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <class Real, class Io>
struct SplitActions {                      // the env's action row in the warp tile, or Philox (random policy)
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

template <int KR, int P> struct SplitDemand {
    int cur_[KR * P], next_[KR * P];
    __device__ __forceinline__ int cur(int rp, int) const { return cur_[rp]; }
    __device__ __forceinline__ int next(int rp, int) const { return next_[rp]; }
};


template <int S, int F, int W, int R, int G>
__device__ __forceinline__ int split_gnode(int e, int l, int q) {       // global node of local slot (echelon e, l), clamped to an existing one
    using Sh = SplitShape<S, F, W, R>;
    const int j = q + G * l, c = Sh::cnt(e);
    return Sh::base(e) + (j < c ? j : c - 1);
}

// The lane's registers from the state arrays in HBM (FROM_SPEC = false) or from the spec's initial state (reset, :402-412).
template <class Real, int S, int F, int W, int R, int P, int L, int G, bool FROM_SPEC, int KT, int MAXPART, int KOWN>
__device__ __forceinline__ void split_load_state(const ScDevSpec<Real>& sp, const ScState<Real>& st, SplitRegs<Real, KT, P, L, MAXPART, KOWN>& rg, int q,
                                                 long long n, long long env, bool active) {
    constexpr bool kFixed = SplitRegs<Real, KT, P, L, MAXPART, KOWN>::kFixed;
    using Sh = SplitShape<S, F, W, R>;
    constexpr int KS = (S + G - 1) / G, KF = (F + G - 1) / G, KW = (W + G - 1) / G, KR = (R + G - 1) / G;
    static_assert(KT == KS + KF + KW + KR, "register layout");
    constexpr int koff[4] = {0, KS, KS + KF, KS + KF + KW}, kcnt[4] = {KS, KF, KW, KR};
    constexpr int KMAX = (KS > KF ? KS : KF) > (KW > KR ? KW : KR) ? (KS > KF ? KS : KF) : (KW > KR ? KW : KR);
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int l = 0; l < KMAX; ++l)
#pragma unroll
            for (int p = 0; p < P; ++p) if (l < kcnt[e]) {
                const int li = (koff[e] + l) * P + p, np = split_gnode<S, F, W, R, G>(e, l, q) * P + p;
                const bool ok = (q + G * l < Sh::cnt(e)) && active;
                Real qs = (Real)1;
                if constexpr (kFixed) qs = (Real)__ldg(sp.q_scale + np);
                Real tmp[L];                                              // the slots in Real, converted below
                if constexpr (FROM_SPEC) {
                    rg.stock[li] = ok ? __ldg(sp.stock0 + np) : (Real)0;
#pragma unroll
                    for (int k = 0; k < L; ++k) tmp[k] = (Real)0;
                    const int i0 = ok ? sp.ibeg[np] : 0, i1 = ok ? sp.ibeg[np + 1] : 0;
                    for (int i = i0; i < i1; ++i) {                       // initial in-transit material (:405-412)
                        const int when = sp.itime[i] - 1; const Real amt = sp.iamt[i];
#pragma unroll
                        for (int k = 0; k < L; ++k) tmp[k] += (when == k) ? amt : (Real)0;
                    }
                } else {
                    rg.stock[li] = ok ? st.stock[(long long)np * n + env] : (Real)0;
#pragma unroll
                    for (int k = 0; k < L; ++k) tmp[k] = ok ? st.slot[((long long)np * L + k) * n + env] : (Real)0;
                }
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    if constexpr (kFixed) rg.slot[li * L + k] = __double2uint_rn(tmp[k] * qs);
                    else rg.slot[li * L + k] = tmp[k];
                }
            }
}

// Demand of this lane's retailers for period `row`: same Philox words and mapping as every other kernel.
template <class Real, int S, int F, int W, int R, int P, int G, int KR>
__device__ __forceinline__ void split_fetch_demand(const ScDevSpec<Real>& sp, const ScInputs& in, SplitDemand<KR, P>& dm, int q, long long n,
                                                   long long env, unsigned long long genv, unsigned episode, int row, bool all_uniform,
                                                   const int* d_lo_i, const unsigned* d_cnt) {
    Philox4 blk = Philox4{0, 0, 0, 0}; int have = -1;
#pragma unroll
    for (int l = 0; l < KR; ++l)
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int j = q + G * l;
            const int rp = (j < R ? j : R - 1) * P + p;
            int d;
            if (all_uniform && in.inj_demand == nullptr) {
                if ((rp >> 2) != have) { have = rp >> 2; blk = rng_block(in.key, genv, episode, (unsigned)row, STREAM_DEMAND, (unsigned)have); }
                d = uniform_int(pick_word(blk, rp & 3), d_lo_i[p], d_cnt[p]);
            } else {
                d = sc_demand(sp, in, n, env, episode, row, rp, p);
            }
            dm.next_[l * P + p] = j < R ? d : 0;
        }
}

template <class Real, class Io, int S, int F, int W, int R, int P, int L, int G, bool RANDOM_ACT, int THREADS, int MINB, int ACT_TILES = 1>
__global__ void __launch_bounds__(THREADS, MINB)
sc_split_rollout(ScDevSpec<Real> sp, ScState<Real> st, ScInputs in, int n_steps, const Io* __restrict__ actions, Io* __restrict__ obs,
                 Io* __restrict__ reward, unsigned char* __restrict__ done) {
    static_assert(sizeof(Io) == 4, "the split kernel serves the fp32-I/O modes");
    static_assert(G == 2 || G == 4 || G == 8, "lanes per env");
    using Sh = SplitShape<S, F, W, R>;
    constexpr int KS = (S + G - 1) / G, KF = (F + G - 1) / G, KW = (W + G - 1) / G, KR = (R + G - 1) / G, KT = KS + KF + KW + KR;
    constexpr int N = S + F + W + R, RP = R * P;
    constexpr int A = S * (P + P * F) + F * P * W + W * P * R, O = RP + N * P * (1 + L) + 1;
    constexpr int EPW = 32 / G;                                   // envs per warp
    using T0 = SplitTopo<Real, 0, S, F, W, R, P, L, G>;
    using T1 = SplitTopo<Real, 1, S, F, W, R, P, L, G>;
    using T2 = SplitTopo<Real, 2, S, F, W, R, P, L, G>;
    using T3 = SplitTopo<Real, 3, S, F, W, R, P, L, G>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = THREADS == 32 ? 0 : (int)(threadIdx.x >> 5);   // one-warp blocks: uniform addresses
    const int q = lane % G, e_in_warp = lane / G;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw) + warp;
    static_assert(ACT_TILES == 1 || ACT_TILES == 2, "action tiles per warp");
    Io* s_act0 = reinterpret_cast<Io*>(smem_raw + 128) + (size_t)warp * EPW * (ACT_TILES * A + O);
    Io* s_obs = s_act0 + ACT_TILES * EPW * A;
    const long long env0 = ((long long)blockIdx.x * (THREADS / 32) + warp) * EPW;
    const long long env = env0 + e_in_warp;
    const long long n = st.n;
    const bool active = env < n;
    const int n_here = env0 >= n ? 0 : (int)((n - env0) < EPW ? (n - env0) : EPW);
    ReturnStats rs; rs.init();

    const unsigned act_bytes = (unsigned)(n_here * A * 4), obs_bytes = (unsigned)(n_here * O * 4);
    const bool bulk_in = !RANDOM_ACT && n_here > 0 && (((size_t)actions) & 15) == 0 && (((size_t)n * A * 4) & 15) == 0 &&
                         (act_bytes & 15) == 0 && ((EPW * A * 4) & 15) == 0;
    const bool bulk_out = obs != nullptr && n_here > 0 && (((size_t)obs) & 15) == 0 && (((size_t)n * O * 4) & 15) == 0 &&
                          (obs_bytes & 15) == 0 && ((EPW * (ACT_TILES * A + O) * 4) & 15) == 0;
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();

    const T0 tp0(sp, q); const T1 tp1(sp, q); const T2 tp2(sp, q); const T3 tp3(sp, q);
    constexpr int MAXPART = P * (F > W ? (F > R ? F : R) : (W > R ? W : R));
    using Regs = SplitRegs<Real, KT, P, L, MAXPART, KS>;
    constexpr bool kFixed = Regs::kFixed;          // fp64 state: the ring as 32-bit fixed point
    Regs rg;
    SplitDemand<KR, P> dm;
    int t = 0; unsigned episode = 0; double ep_ret = 0.0;
    const unsigned long long genv = (unsigned long long)(in.first_env + env);
    bool all_uniform = true;
    float d_lo[P], d_inv[P];
    int d_lo_i[P]; unsigned d_cnt[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        all_uniform = all_uniform && (sp.dkind[p] == 0);
        d_lo[p] = (float)sp.dlo[p]; d_inv[p] = (float)(1.0 / (sp.dhi[p] - sp.dlo[p]));
        d_lo_i[p] = (int)sp.dlo[p]; d_cnt[p] = (unsigned)((int)sp.dhi[p] - (int)sp.dlo[p] + 1);
    }

    auto gnode = [&](int e, int l) { return split_gnode<S, F, W, R, G>(e, l, q); };
    auto gvalid = [&](int e, int l) { return q + G * l < Sh::cnt(e); };
    constexpr int koff[4] = {0, KS, KS + KF, KS + KF + KW}, kcnt[4] = {KS, KF, KW, KR};
    constexpr int KMAX = (KS > KF ? KS : KF) > (KW > KR ? KW : KR) ? (KS > KF ? KS : KF) : (KW > KR ? KW : KR);

    if (bulk_in && n_steps > 0 && lane == 0) {
        mbar_expect_tx(bar, act_bytes);
        bulk_g2s(s_act0, actions + env0 * A, act_bytes, bar);
    }
    if (active) { t = st.t[env]; episode = (unsigned)st.episode[env]; ep_ret = st.ep_ret[env]; }
    split_load_state<Real, S, F, W, R, P, L, G, false>(sp, st, rg, q, n, env, active);
    if (active) split_fetch_demand<Real, S, F, W, R, P, G>(sp, in, dm, q, n, env, genv, episode, t, all_uniform, d_lo_i, d_cnt);
    else {
#pragma unroll
        for (int i = 0; i < KR * P; ++i) dm.next_[i] = 0;
    }
    const int T_end = sp.T;

#pragma unroll 1
    for (int s = 0; s < n_steps; ++s) {
        const long long io0 = (long long)s * n + env0;
#if SCG_SYNC_PERIOD
        __syncthreads();                   // keep the block's warps in the same stretch of the (larger than I-cache) loop body
#endif
        Io* s_act = s_act0 + ((ACT_TILES == 2 && bulk_in) ? (s & 1) * EPW * A : 0);
        if (bulk_in) {
            if constexpr (ACT_TILES == 1) {
                if (s + 1 < n_steps && elect_one()) bulk_prefetch_l2(actions + (io0 + n) * A, act_bytes);
                while (!mbar_try_wait(bar, (unsigned)(s & 1))) {}
            } else {
                // two tiles: request tile s+1 into the buffer every lane finished reading in period s-1 (warp barriers below)
                while (!mbar_try_wait(bar, (unsigned)(s & 1))) {}
                if (s + 1 < n_steps && elect_one()) {
                    mbar_expect_tx(bar, act_bytes);
                    bulk_g2s(s_act0 + ((s + 1) & 1) * EPW * A, actions + (io0 + n) * A, act_bytes, bar);
                }
            }
        } else if (!RANDOM_ACT && n_here > 0) {
            __syncwarp();
            warp_copy<Io>(s_act, actions + io0 * A, n_here * A, lane);
            __syncwarp();
        }
        const bool stepping = active && t < T_end;               // the same for the G lanes of an env
        Real cost = (Real)0;
        if (stepping) {
            ++t;
#pragma unroll
            for (int i = 0; i < KR * P; ++i) dm.cur_[i] = dm.next_[i];
            // all arrivals first (pushes carry time >= t+1: no same-period dependence), :220-228
            if constexpr (!kFixed) {
#pragma unroll
                for (int i = 0; i < KT * P; ++i) {
                    rg.stock[i] += rg.slot[i * L];
#pragma unroll
                    for (int k = 0; k + 1 < L; ++k) rg.slot[i * L + k] = rg.slot[i * L + k + 1];
                    rg.slot[i * L + L - 1] = (Real)0;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int l = 0; l < KMAX; ++l)
#pragma unroll
                        for (int p = 0; p < P; ++p) if (l < kcnt[e]) {
                            const int i = (koff[e] + l) * P + p;
                            rg.stock[i] = fma(u2d(rg.slot[i * L]), __ldg(sp.q_inv + gnode(e, l) * P + p), rg.stock[i]);
#pragma unroll
                            for (int k = 0; k + 1 < L; ++k) rg.slot[i * L + k] = rg.slot[i * L + k + 1];
                            rg.slot[i * L + L - 1] = 0u;
                        }
            }
        }
        SplitActions<Real, Io> ar{s_act + e_in_warp * A, RANDOM_ACT, in.key, genv, episode, t, Philox4{0, 0, 0, 0}, -1};
        ConstLeadtime<L> lr;
        // echelon by echelon: the lane's nodes act, the partial shipments are summed over the env's lanes (all lanes of the
        // warp take part in the shuffles) and each destination's owner books the total into its ring
        {
            auto& pp = static_cast<SplitPipe<Real, koff[0], P, L, F, KT, MAXPART, KS>&>(rg);
            pp.begin();
            if (stepping) cost += sc_step_nodes<Real>(tp0, pp, t, ar, lr, dm);
            if constexpr (kFixed) if (stepping) {  // what the lane's suppliers ordered: one rounding per (node, product)
#pragma unroll
                for (int l = 0; l < KS; ++l)
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        rg.slot[(l * P + p) * L + L - 1] = __double2uint_rn(rg.own[l * P + p] * __ldg(sp.q_scale + gnode(0, l) * P + p));
            }
#pragma unroll
            for (int i = 0; i < F; ++i)
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const int tail = ((koff[1] + i / G) * P + p) * L + L - 1;
                    if constexpr (kFixed) {
                        // each lane's partial in units of the DESTINATION's quantum, summed as integers (order-independent)
                        const unsigned tot = group_sum<G>(d2u_wrap(pp.part[i * P + p], __ldg(sp.q_scale + (Sh::base(1) + i) * P + p)));
                        rg.slot[tail] = (stepping && q == i % G) ? tot : rg.slot[tail];     // a frozen env keeps its ring
                    } else {
                        const Real tot = group_sum<G>(pp.part[i * P + p]);
                        rg.slot[tail] += (q == i % G) ? tot : (Real)0;
                    }
                }
        }
        {
            auto& pp = static_cast<SplitPipe<Real, koff[1], P, L, W, KT, MAXPART, KS>&>(rg);
            pp.begin();
            if (stepping) cost += sc_step_nodes<Real>(tp1, pp, t, ar, lr, dm);
#pragma unroll
            for (int i = 0; i < W; ++i)
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const int tail = ((koff[2] + i / G) * P + p) * L + L - 1;
                    if constexpr (kFixed) {
                        // each lane's partial in units of the DESTINATION's quantum, summed as integers (order-independent)
                        const unsigned tot = group_sum<G>(d2u_wrap(pp.part[i * P + p], __ldg(sp.q_scale + (Sh::base(2) + i) * P + p)));
                        rg.slot[tail] = (stepping && q == i % G) ? tot : rg.slot[tail];     // a frozen env keeps its ring
                    } else {
                        const Real tot = group_sum<G>(pp.part[i * P + p]);
                        rg.slot[tail] += (q == i % G) ? tot : (Real)0;
                    }
                }
        }
        {
            auto& pp = static_cast<SplitPipe<Real, koff[2], P, L, R, KT, MAXPART, KS>&>(rg);
            pp.begin();
            if (stepping) cost += sc_step_nodes<Real>(tp2, pp, t, ar, lr, dm);
#pragma unroll
            for (int i = 0; i < R; ++i)
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const int tail = ((koff[3] + i / G) * P + p) * L + L - 1;
                    if constexpr (kFixed) {
                        // each lane's partial in units of the DESTINATION's quantum, summed as integers (order-independent)
                        const unsigned tot = group_sum<G>(d2u_wrap(pp.part[i * P + p], __ldg(sp.q_scale + (Sh::base(3) + i) * P + p)));
                        rg.slot[tail] = (stepping && q == i % G) ? tot : rg.slot[tail];     // a frozen env keeps its ring
                    } else {
                        const Real tot = group_sum<G>(pp.part[i * P + p]);
                        rg.slot[tail] += (q == i % G) ? tot : (Real)0;
                    }
                }
        }
        {
            auto& pp = static_cast<SplitPipe<Real, koff[3], P, L, 0, KT, MAXPART, KS>&>(rg);
            pp.begin();
            if (stepping) cost += sc_step_nodes<Real>(tp3, pp, t, ar, lr, dm);
        }
        const Real r = -group_sum<G>(cost);
        bool fin = true;
        if (stepping) {
            ep_ret += (double)r;
            fin = (t == T_end);
            if (fin) {
                if (q == 0) rs.add(ep_ret);
                if (in.auto_reset) {
                    ++episode; t = 0; ep_ret = 0.0;
                    split_load_state<Real, S, F, W, R, P, L, G, true>(sp, st, rg, q, n, env, true);
                }
            }
        }
        __syncwarp();                      // every lane is done with the action tile of period s
        if (ACT_TILES == 1 && bulk_in && s + 1 < n_steps && elect_one()) {
            mbar_expect_tx(bar, act_bytes);
            bulk_g2s(s_act, actions + (io0 + n) * A, act_bytes, bar);
        }
        if (bulk_out) {
            bulk_wait_read_all();          // every lane: its own bulk groups (a no-op for the lanes that never issued one)
            __syncwarp();
        }
        if (active) {
            if (stepping) split_fetch_demand<Real, S, F, W, R, P, G>(sp, in, dm, q, n, env, genv, episode, t, all_uniform, d_lo_i, d_cnt);
            if (obs != nullptr) {
                Io* row = s_obs + e_in_warp * O;
                auto unit = [](float x, float inv) { return fmaf(__saturatef(x * inv), 2.0f, -1.0f); };     // clip(2x-1, -1, 1)
#pragma unroll
                for (int l = 0; l < KR; ++l)
                    if (gvalid(3, l)) {
                        const int j = gnode(3, l) - Sh::base(3);
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            row[j * P + p] = unit((float)dm.next_[l * P + p] - d_lo[p], d_inv[p]);
                        }
                    }
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int l = 0; l < KMAX; ++l)
                        if (l < kcnt[e] && gvalid(e, l)) {
                            const int nd = gnode(e, l);
                            Io* o = row + RP + nd * P * (1 + L);
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                const int li = (koff[e] + l) * P + p;
                                o[p] = unit((float)rg.stock[li], __ldg(sp.inv_cap + nd * P + p));
                                const float ims = kFixed ? __ldg(sp.inv_ms_q + nd * P + p) : __ldg(sp.inv_ms + nd * P + p);   // fixed point: quantum folded in
#pragma unroll
                                for (int k = 0; k < L; ++k) o[P + p * L + k] = unit((float)rg.slot[li * L + k], ims);   // ring == lead time: last bucket / 1
                            }
                        }
                if (q == 0) row[O - 1] = unit((float)(T_end - t), 1.0f / (float)T_end);
            }
            if (q == 0) {
                reward[io0 + e_in_warp] = (Io)r;
                done[io0 + e_in_warp] = fin ? 1 : 0;
            }
        }
        if (bulk_out) {
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) { bulk_s2g(obs + io0 * O, s_obs, obs_bytes); bulk_commit(); }
        } else if (obs != nullptr && n_here > 0) {
            __syncwarp();
            warp_copy<Io>(obs + io0 * O, s_obs, n_here * O, lane);
            __syncwarp();
        }
    }
    if (bulk_out) bulk_wait_all();     // every lane: its own bulk groups

    if (active) {
        if (q == 0) { st.t[env] = t; st.episode[env] = (int)episode; st.ep_ret[env] = ep_ret; }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int l = 0; l < KMAX; ++l)
                if (l < kcnt[e] && gvalid(e, l))
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int li = (koff[e] + l) * P + p, np = gnode(e, l) * P + p;
                        st.stock[(long long)np * n + env] = rg.stock[li];
#pragma unroll
                        for (int k = 0; k < L; ++k) {
                            if constexpr (kFixed) st.slot[((long long)np * L + k) * n + env] = (double)rg.slot[li * L + k] * __ldg(sp.q_inv + np);
                            else st.slot[((long long)np * L + k) * n + env] = rg.slot[li * L + k];
                        }
                    }
    }
    rs.flush(st.stats);
}

}  // namespace scg
