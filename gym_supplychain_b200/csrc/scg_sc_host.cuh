// scg_sc_host.cuh -- host-side launch logic for the supply-chain kernels, templated on Real.
// Included by scg_sc_f32.cu (default FMA contraction) and scg_sc_f64.cu (-fmad=false).
#pragma once
#include <cmath>
#include <cstring>

#include "scg_internal.h"
#include "scg_sc_static.cuh"

namespace scg {

// Integer thresholds of the truncated-Poisson lead-time draw: with u = word / 2^32,
// lt = 1 + #{j < maxL-1 : u >= CDF_Poisson(avg-1)(j)}  <=>  word >= ceil(CDF(j) * 2^32)   (supplychain_env.py:670-672)
static inline std::vector<unsigned long long> sc_leadtime_thresholds(const ScHostSpec& hs) {
    std::vector<unsigned long long> thr((size_t)(hs.maxL > 1 ? hs.maxL : 1), ~0ull);
    const double lam = (double)(hs.avgL - 1);
    double p = std::exp(-lam), F = p;
    for (int j = 0; j + 1 < hs.maxL; ++j) {
        const double x = std::ceil(F * 4294967296.0);
        thr[j] = x >= 4294967296.0 ? (1ull << 32) : (x < 1.0 ? 1ull : (unsigned long long)x);   // 2^32: never reached by a 32-bit word
        p *= lam / (double)(j + 1);
        F += p;
    }
    return thr;
}

// Power-of-two scales of the fixed-point lead-time ring (RingFixed, scg_sc_static.cuh), one per (node, product):
// 2^k with k the largest exponent that keeps the biggest amount that can be in flight to that (node, product) in one
// period below 2^32.  The bound follows from SC_Action.apply (supplychain_env.py:49-96): a SHIP amount never exceeds the
// source's stock, which was clipped to its capacity at the start of the period (:232-240) -- divided by the processing
// ratio at a factory (:310), capped by the edge's ship capacity elsewhere (:316-328); a SUPPLY amount is
// action * supply_capacity (:56), given 8x headroom for out-of-range actions.  Everything else gets 2x headroom.
static inline std::vector<double> sc_ring_scales(const ScHostSpec& hs) {
    const int NP = hs.N * hs.P;
    std::vector<double> bound((size_t)NP, 0.0), scale((size_t)NP, 1.0);
    for (int n = 0; n < hs.N; ++n) {
        const bool factory = hs.proc_cap[n] > 0;
        for (int p = 0; p < hs.P; ++p) {
            const int np = n * hs.P + p;
            if (hs.sup_cap[np] > 0) bound[np] += 8.0 * hs.sup_cap[np];
            for (int e = hs.ebeg[n]; e < hs.ebeg[n + 1]; ++e) {
                double b = hs.cap[np];
                if (factory && hs.ratio[np] > 0) b /= hs.ratio[np];
                else if (hs.ecap[e] < b) b = hs.ecap[e];
                bound[hs.edst[e] * hs.P + p] += 2.0 * (b > 0 ? b : 0.0);
            }
        }
    }
    for (int np = 0; np < NP; ++np) {
        double init = 0.0;
        for (int k = hs.ibeg[np]; k < hs.ibeg[np + 1]; ++k) init += std::fabs(hs.iamt[k]);
        if (2.0 * init > bound[np]) bound[np] = 2.0 * init;
        if (hs.stochL) bound[np] *= (double)hs.maxL;        // stochastic lead times: one arrival slot collects up to maxL periods' pushes
        int k = 0;
        if (bound[np] > 0 && std::isfinite(bound[np])) k = std::ilogb(4294967295.0 / bound[np]);
        k = k < -60 ? -60 : (k > 60 ? 60 : k);
        scale[np] = std::ldexp(1.0, k);
    }
    return scale;
}

// ---- device spec ---------------------------------------------------------------
template <class Real>
static int sc_build_devspec_t(const ScHostSpec& hs, void** blob_out, void* devspec_out) {
    // pack every array into one host buffer (8-byte aligned sections), upload once
    std::vector<unsigned char> buf;
    auto put = [&buf](const void* src, size_t bytes) -> size_t {
        size_t off = (buf.size() + 15) & ~(size_t)15;
        buf.resize(off + (bytes ? bytes : 16));
        if (bytes) memcpy(buf.data() + off, src, bytes);
        return off;
    };
    auto put_real = [&](const std::vector<double>& v) -> size_t {
        std::vector<Real> tmp(v.begin(), v.end());
        return put(tmp.data(), tmp.size() * sizeof(Real));
    };
    auto put_int = [&](const std::vector<int>& v) -> size_t { return put(v.data(), v.size() * sizeof(int)); };
    auto put_dbl = [&](const std::vector<double>& v) -> size_t { return put(v.data(), v.size() * sizeof(double)); };

    size_t o_last = put_int(hs.last), o_ebeg = put_int(hs.ebeg), o_act = put_int(hs.act_off), o_lt = put_int(hs.lt_off);
    size_t o_ret = put_int(hs.ret_idx), o_sup = put_int(hs.is_sup), o_pc = put_real(hs.proc_cap);
    size_t o_s0 = put_real(hs.stock0), o_cap = put_real(hs.cap), o_hold = put_real(hs.hold), o_sc = put_real(hs.sup_cap);
    size_t o_scost = put_real(hs.sup_cost), o_ratio = put_real(hs.ratio), o_pcost = put_real(hs.proc_cost);
    size_t o_ms = put_real(hs.max_ship), o_edst = put_int(hs.edst), o_ecap = put_real(hs.ecap), o_ecost = put_real(hs.ecost);
    size_t o_ibeg = put_int(hs.ibeg), o_itime = put_int(hs.itime), o_iamt = put_real(hs.iamt), o_dkind = put_int(hs.dkind);
    size_t o_dlo = put_dbl(hs.dlo), o_dhi = put_dbl(hs.dhi), o_dstd = put_dbl(hs.dstd), o_dpk = put_dbl(hs.dpeaks);
    size_t o_dmin = put_dbl(hs.dminavg), o_dmax = put_dbl(hs.dmaxavg);
    const std::vector<unsigned long long> thr = sc_leadtime_thresholds(hs);
    size_t o_thr = put(thr.data(), thr.size() * sizeof(unsigned long long));
    std::vector<float> inv_cap(hs.cap.size()), inv_ms(hs.cap.size());
    for (size_t i = 0; i < hs.cap.size(); ++i) {
        inv_cap[i] = hs.cap[i] > 0 ? (float)(1.0 / hs.cap[i]) : 0.0f;
        inv_ms[i] = hs.max_ship[i] > 0 ? (float)(1.0 / hs.max_ship[i]) : 0.0f;
    }
    size_t o_icap = put(inv_cap.data(), inv_cap.size() * sizeof(float)), o_ims = put(inv_ms.data(), inv_ms.size() * sizeof(float));
    const std::vector<double> q_scale = sc_ring_scales(hs);
    std::vector<double> q_inv(q_scale.size());
    std::vector<float> inv_ms_q(q_scale.size());
    for (size_t i = 0; i < q_scale.size(); ++i) {
        q_inv[i] = 1.0 / q_scale[i];
        inv_ms_q[i] = hs.max_ship[i] > 0 ? (float)(q_inv[i] / hs.max_ship[i]) : 0.0f;
    }
    size_t o_qs = put_dbl(q_scale), o_qi = put_dbl(q_inv), o_imq = put(inv_ms_q.data(), inv_ms_q.size() * sizeof(float));

    unsigned char* d = nullptr;
    SCG_CUDA(cudaMalloc(&d, buf.size()));
    SCG_CUDA(cudaMemcpy(d, buf.data(), buf.size(), cudaMemcpyHostToDevice));
    ScDevSpec<Real> sp;
    memset(&sp, 0, sizeof(sp));
    sp.N = hs.N; sp.P = hs.P; sp.E = hs.E; sp.T = hs.T; sp.avgL = hs.avgL; sp.maxL = hs.maxL; sp.stochL = hs.stochL;
    sp.RP = hs.RP; sp.A = hs.A; sp.O = hs.O; sp.K = hs.K; sp.ring = hs.ring; sp.heapcap = hs.heapcap;
    sp.n_init = (int)hs.itime.size();
    sp.c_unmet = (Real)hs.c_unmet; sp.c_xstock = (Real)hs.c_xstock; sp.c_xproc = (Real)hs.c_xproc; sp.c_xship = (Real)hs.c_xship;
#define SCG_P(T_, off) reinterpret_cast<const T_*>(d + (off))
    sp.last = SCG_P(int, o_last); sp.ebeg = SCG_P(int, o_ebeg); sp.act_off = SCG_P(int, o_act); sp.lt_off = SCG_P(int, o_lt);
    sp.ret_idx = SCG_P(int, o_ret); sp.is_sup = SCG_P(int, o_sup); sp.proc_cap = SCG_P(Real, o_pc);
    sp.stock0 = SCG_P(Real, o_s0); sp.cap = SCG_P(Real, o_cap); sp.hold = SCG_P(Real, o_hold); sp.sup_cap = SCG_P(Real, o_sc);
    sp.sup_cost = SCG_P(Real, o_scost); sp.ratio = SCG_P(Real, o_ratio); sp.proc_cost = SCG_P(Real, o_pcost);
    sp.max_ship = SCG_P(Real, o_ms); sp.edst = SCG_P(int, o_edst); sp.ecap = SCG_P(Real, o_ecap); sp.ecost = SCG_P(Real, o_ecost);
    sp.ibeg = SCG_P(int, o_ibeg); sp.itime = SCG_P(int, o_itime); sp.iamt = SCG_P(Real, o_iamt); sp.dkind = SCG_P(int, o_dkind);
    sp.dlo = SCG_P(double, o_dlo); sp.dhi = SCG_P(double, o_dhi); sp.dstd = SCG_P(double, o_dstd); sp.dpeaks = SCG_P(double, o_dpk);
    sp.dminavg = SCG_P(double, o_dmin); sp.dmaxavg = SCG_P(double, o_dmax);
    sp.lt_thr = SCG_P(unsigned long long, o_thr);
    sp.inv_cap = SCG_P(float, o_icap); sp.inv_ms = SCG_P(float, o_ims);
    sp.q_scale = SCG_P(double, o_qs); sp.q_inv = SCG_P(double, o_qi); sp.inv_ms_q = SCG_P(float, o_imq);
#undef SCG_P
    memcpy(devspec_out, &sp, sizeof(sp));
    *blob_out = d;
    return SCG_OK;
}

// ---- layered-variant matching ----------------------------------------------------
static inline bool sc_is_layered(const ScHostSpec& hs, int S, int F, int W, int R, int P, int L, int ML = 0, bool stoch = false) {
    if (ML == 0) ML = L;
    if (hs.N != S + F + W + R || hs.P != P || (hs.stochL != 0) != stoch || hs.avgL != L || hs.maxL != ML || hs.ring != ML) return false;
    for (double cap : hs.ecap) if (!(cap >= 0)) return false;       // the kernels rely on non-negative ship capacities
    const int cnt[4] = {S, F, W, R};
    const int base[4] = {0, S, S + F, S + F + W};
    for (int e = 0; e < 4; ++e)
        for (int j = 0; j < cnt[e]; ++j) {
            const int n = base[e] + j;
            const int D = hs.ebeg[n + 1] - hs.ebeg[n];
            if ((hs.last[n] != 0) != (e == 3)) return false;
            if (e == 3) { if (D != 0) return false; }
            else {
                if (D != cnt[e + 1]) return false;
                for (int i = 0; i < D; ++i) if (hs.edst[hs.ebeg[n] + i] != base[e + 1] + i) return false;
            }
            if ((hs.proc_cap[n] > 0) != (e == 1)) return false;
            for (int p = 0; p < P; ++p) {
                if ((hs.sup_cap[n * P + p] > 0) != (e == 0)) return false;
                if (!(hs.cap[n * P + p] > 0)) return false;
                if (e == 1 && !(hs.ratio[n * P + p] > 0)) return false;
            }
        }
    return true;
}

template <class Real, int S, int F, int W, int R, int P, int L, int ML, bool STOCH>
static void sc_fill_consts(const ScHostSpec& hs, LayeredConsts<Real, S, F, W, R, P, L, ML, STOCH>& c) {
    memset(&c, 0, sizeof(c));
    constexpr int N = S + F + W + R, NP = N * P, E = S * F + F * W + W * R;
    constexpr bool fixed_ring = sizeof(Real) == 8 && (!STOCH || SCG_FIXED_RING_STOCH != 0);   // RingFixed*: the fp32-I/O kernel reads obs_mid / obs_last
    const std::vector<double> qs = sc_ring_scales(hs);
    for (int i = 0; i < NP; ++i) {
        c.cap[i] = (Real)hs.cap[i]; c.hold[i] = (Real)hs.hold[i]; c.stock0[i] = (Real)hs.stock0[i];
        c.max_ship[i] = (Real)hs.max_ship[i];
        c.q_scale[i] = (Real)qs[i]; c.q_inv[i] = (Real)(1.0 / qs[i]);
        const double q = fixed_ring ? 1.0 / qs[i] : 1.0;
        c.obs_stock[i] = (float)(1.0 / hs.cap[i]);
        c.obs_mid[i] = (float)(hs.max_ship[i] > 0 ? q / hs.max_ship[i] : 0.0);
        c.obs_last[i] = (float)(hs.max_ship[i] > 0 ? q / (hs.max_ship[i] * (double)(hs.maxL - (hs.avgL - 1))) : 0.0);
        for (int k = hs.ibeg[i]; k < hs.ibeg[i + 1]; ++k) c.init_slot[i * ML + hs.itime[k] - 1] += (Real)hs.iamt[k];
    }
    for (int i = 0; i < S * P; ++i) { c.sup_cap[i] = (Real)hs.sup_cap[i]; c.sup_cost[i] = (Real)hs.sup_cost[i]; }
    for (int j = 0; j < F; ++j) {
        c.proc_cap[j] = (Real)hs.proc_cap[S + j];
        for (int p = 0; p < P; ++p) {
            c.ratio[j * P + p] = (Real)hs.ratio[(S + j) * P + p];
            c.inv_ratio[j * P + p] = (Real)(1.0 / hs.ratio[(S + j) * P + p]);
            c.proc_cost[j * P + p] = (Real)hs.proc_cost[(S + j) * P + p];
        }
    }
    for (int e = 0; e < E; ++e) {
        c.ecap[e] = (Real)hs.ecap[e];
        for (int p = 0; p < P; ++p) c.ecost[e * P + p] = (Real)hs.ecost[e * P + p];
    }
    c.c_unmet = (Real)hs.c_unmet; c.c_xstock = (Real)hs.c_xstock; c.c_xproc = (Real)hs.c_xproc; c.c_xship = (Real)hs.c_xship;
    c.all_uniform = 1;
    for (int p = 0; p < P; ++p) {
        c.dlo[p] = (Real)hs.dlo[p]; c.dspan[p] = (Real)(hs.dhi[p] - hs.dlo[p]);
        c.obs_dem[p] = (float)(1.0 / (hs.dhi[p] - hs.dlo[p]));
        c.dlo_i[p] = (int)hs.dlo[p]; c.dcount[p] = (int)hs.dhi[p] - (int)hs.dlo[p] + 1;
        if (hs.dkind[p] != 0) c.all_uniform = 0;
    }
    const std::vector<unsigned long long> thr = sc_leadtime_thresholds(hs);
    for (int j = 0; j < ML; ++j) c.lt_thr[j] = j < (int)thr.size() ? thr[j] : ~0ull;
    c.T = hs.T;
    c.obs_time = (float)(1.0 / (double)hs.T);
}

template <class Real, class Io, int S, int F, int W, int R, int P, int L, int THREADS, int MINB, int ML = L, bool STOCH = false>
static int sc_launch_layered(const ScRun& r) {
    using C = LayeredConsts<Real, S, F, W, R, P, L, ML, STOCH>;
    using Tp = LayeredTopo<Real, sizeof(Io) == 8, S, F, W, R, P, L, ML, STOCH>;
    static_assert(sizeof(C) <= 3800, "constants must fit the kernel parameter bank");
    // The constant block depends only on the handle's spec: built on the first launch, reused afterwards (the step()
    // API launches once per period -- rebuilding it cost more host time than the launch itself).
    C local;
    const C* cp = &local;
    if (r.consts_cache && r.consts_cache->size() == sizeof(C)) {
        cp = reinterpret_cast<const C*>(r.consts_cache->data());
    } else {
        sc_fill_consts<Real, S, F, W, R, P, L, ML, STOCH>(*r.hs, local);
        if (r.consts_cache) {
            r.consts_cache->assign(reinterpret_cast<const unsigned char*>(&local), reinterpret_cast<const unsigned char*>(&local) + sizeof(C));
            cp = reinterpret_cast<const C*>(r.consts_cache->data());
        }
    }
    const C& c = *cp;
    const ScDevSpec<Real>& sp = *reinterpret_cast<const ScDevSpec<Real>*>(r.dev_spec);
    const ScState<Real>& st = *reinterpret_cast<const ScState<Real>*>(r.state);
    ScInputs in{r.inj_demand, r.inj_leadtimes, r.key, r.first_env, r.auto_reset, r.err};
    const size_t smem = layered_smem_header<Real, ML, STOCH>() +
                        (size_t)(THREADS / 32) * 32 * (layered_action_tiles<Real, Io, Tp::kN * P>() * Tp::kA + Tp::kO) * sizeof(Io);
    const long long per_block = THREADS;
    const unsigned blocks = (unsigned)((st.n + per_block - 1) / per_block);
    static unsigned long long attr_set[2] = {0ull, 0ull};         // [random actions?] bit d: shared-memory size set on device d
    const int rnd = r.actions == nullptr ? 1 : 0;
    const bool need_attr = r.device < 0 || r.device >= 64 || !((attr_set[rnd] >> r.device) & 1ull);
    if (rnd) {
        auto k = sc_layered_rollout<Real, Io, S, F, W, R, P, L, ML, STOCH, true, THREADS, MINB>;
        if (need_attr) SCG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<blocks, THREADS, smem, r.stream>>>(c, sp, st, in, r.n_steps, nullptr, (Io*)r.obs, (Io*)r.reward, r.done);
    } else {
        auto k = sc_layered_rollout<Real, Io, S, F, W, R, P, L, ML, STOCH, false, THREADS, MINB>;
        if (need_attr) SCG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<blocks, THREADS, smem, r.stream>>>(c, sp, st, in, r.n_steps, (const Io*)r.actions, (Io*)r.obs, (Io*)r.reward, r.done);
    }
    if (need_attr && r.device >= 0 && r.device < 64) attr_set[rnd] |= 1ull << r.device;
    SCG_CUDA(cudaGetLastError());
    return SCG_OK;
}

template <class Real, class Io>
static int sc_launch_generic(const ScRun& r) {
    const ScDevSpec<Real>& sp = *reinterpret_cast<const ScDevSpec<Real>*>(r.dev_spec);
    const ScState<Real>& st = *reinterpret_cast<const ScState<Real>*>(r.state);
    ScInputs in{r.inj_demand, r.inj_leadtimes, r.key, r.first_env, r.auto_reset, r.err};
    const int do_reset = r.op == SC_OP_RESET ? 1 : 0;
    const int n_steps = r.op == SC_OP_RESET ? 0 : r.n_steps;
    const bool exact = r.mode == SCG_MODE_EXACT_F64;
    // shared-memory plan: env state per thread + one action/obs tile per warp, within ~200 KB per block
    const size_t NP = (size_t)sp.N * sp.P;
    const size_t state_bytes = exact ? NP * sizeof(double) + NP * sizeof(int) + NP * sp.heapcap * (sizeof(double) + sizeof(int))
                                     : NP * (1 + (size_t)sp.ring) * sizeof(Real);
    const size_t tile_bytes = (size_t)32 * (sp.A + sp.O) * sizeof(Io);
    const size_t budget = 200 * 1024;
    int threads = 128, use_smem = 0, stage_io = 0;
    // copying the state into shared memory and back only pays when it is reused over several periods: reset() and
    // the single-period step() touch every element about once and work on HBM directly
    // (also measured for the drop-in single env, where every access is a dependent one: no gain, 182 us per ntom step either
    // way -- one thread's instruction latency, not the L2 round trips, is what a single heapq-exact env waits for)
    const bool want_smem = n_steps >= 3;
    for (int th : {128, 64, 32}) {
        if (!want_smem) break;
        if ((size_t)th * state_bytes + (size_t)(th / 32) * tile_bytes + 64 <= budget) { threads = th; use_smem = 1; stage_io = 1; break; }
    }
    if (!use_smem && want_smem) {
        for (int th : {128, 64, 32})
            if ((size_t)th * state_bytes + 64 <= budget) { threads = th; use_smem = 1; break; }
    }
    if (!use_smem && (size_t)4 * tile_bytes + 64 <= budget) stage_io = 1;
    const size_t smem = (use_smem ? (size_t)threads * state_bytes : 0) + (stage_io ? (size_t)(threads / 32) * tile_bytes : 0) + 64;
    const unsigned blocks = (unsigned)((st.n + threads - 1) / threads);
#define SCG_GEN_LAUNCH(REAL, IO, MODE_, INFO_)                                                                         \
    do {                                                                                                               \
        auto k = sc_generic_kernel<REAL, IO, MODE_, INFO_>;                                                            \
        static size_t smem_set[64] = {0};                /* largest shared-memory size granted so far, per device */    \
        const bool known = r.device >= 0 && r.device < 64;                                                             \
        if (!known || smem_set[r.device] < smem) {                                                                     \
            SCG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
            if (known) smem_set[r.device] = smem;                                                                      \
        }                                                                                                              \
        k<<<blocks, threads, smem, r.stream>>>(sp, st, in, n_steps, r.reset_mask, do_reset, (const IO*)r.actions,      \
                                               (IO*)r.obs, (IO*)r.reward, r.done, use_smem, stage_io);                 \
    } while (0)
    if (exact) {
        // only the -fmad=false translation unit (Real = Io = double) may instantiate the exact kernel:
        // a second copy compiled with FMA contraction would be an ODR clash and break bit-exactness
        if constexpr (sizeof(Real) == 8 && sizeof(Io) == 8) {
            if (st.info) SCG_GEN_LAUNCH(double, double, 1, true);
            else SCG_GEN_LAUNCH(double, double, 1, false);
        } else {
            set_error("EXACT mode needs the fp64 translation unit");
            return SCG_EINVAL;
        }
    } else {
        if (st.info) SCG_GEN_LAUNCH(Real, Io, 0, true);
        else SCG_GEN_LAUNCH(Real, Io, 0, false);
    }
#undef SCG_GEN_LAUNCH
    SCG_CUDA(cudaGetLastError());
    return SCG_OK;
}

}  // namespace scg
