// scg_bg.cu -- beergame-v0 / beergame-v2 kernels (all-integer dynamics, one thread = one env).
//
// Restated from the reference (paths relative to its repo root):
//   gym_supplychain/envs/beergame_env.py:66-138    BeerGameEnv.step   (action = delta on incoming order)
//   gym_supplychain/envs/beergame_env.py:140-156   reset
//   gym_supplychain/envs/beergame2_env.py:114-192  BeerGameEnv2.step  (action = order; capacity penalty)
//   gym_supplychain/envs/beergame2_env.py:79-109   reset, randint demand / delays
// The reference's table of shipments indexed by absolute week becomes a ring of `ring` slots
// (slot k = what arrives in week+1+k), shifted once per week, so the state is a few dozen int32
// that stay in registers for the whole rollout.  With 4 levels an action/obs row is 16 bytes:
// every lane moves its row with one 128-bit access and a warp touches 512 contiguous bytes.
#include "scg_internal.h"

namespace scg {

// PF = how many weeks ahead the action row of the 4-level fast path is requested (memory-level parallelism:
// one 16-byte load per lane per week would leave only ~8 KB in flight per SM).
#ifndef SCG_BG_PF
#define SCG_BG_PF 2
#endif
// Blocks of 256 threads per SM: 3 (80 registers).  At 4 (64 registers) the kernels with a ring of 4 and Philox demand / delays
// spill -- with the branch-free bookings below 0.455 of the HBM peak instead of 0.643 -- and the others are within 1 % either way
// (profiles/r2_beergame_onehot_bookings_ab.log, r2_beergame_launch_shape_ab.log).
#ifndef SCG_BG_MINB
#define SCG_BG_MINB 3
#endif
#ifndef SCG_BG_THREADS
#define SCG_BG_THREADS 256
#endif
#ifndef SCG_BG_PF_TABLE
#define SCG_BG_PF_TABLE 3       // kernels whose demand / delays come from the spec's tables: registers to spare at 80 per thread
#endif

// VER: -1 = the game version is read at run time, 0 / 2 = known (beergame-v0 / -v2).
// INP:  0 = where demand / delay come from is decided at run time (injected tables, Philox, the spec's tables, any mix),
//       1 = both from the spec's per-week tables (the reference's default configuration), 2 = both from Philox
//       (customer_demand=(lo,hi), shipment_delays=(lo,hi): beergame2_env.py:86-92).  The specialisations drop ~30 uniform
//       tests and branches per week from a kernel that moves only 37 bytes per env-week.
template <int LMAX, int RMAX, bool STATIC, int VER, int INP>
__global__ void __launch_bounds__(SCG_BG_THREADS, STATIC ? SCG_BG_MINB : 1) bg_kernel(BgDevSpec sp, BgState st, const int32_t* __restrict__ inj_demand_,
                                                 const int32_t* __restrict__ inj_delays_, RngKey key, long long first_env,
                                                 int auto_reset, int n_steps, const unsigned char* __restrict__ reset_mask,
                                                 int do_reset, const int32_t* __restrict__ actions, int32_t* __restrict__ obs,
                                                 int32_t* __restrict__ reward, unsigned char* __restrict__ done, int* err) {
    const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = st.n;
    const int L = STATIC ? LMAX : sp.L;
    const int RING = STATIC ? RMAX : sp.ring;
    const bool v2 = VER < 0 ? sp.version == 2 : VER == 2;
    const int32_t* inj_demand = INP == 0 ? inj_demand_ : nullptr;
    const int32_t* inj_delays = INP == 0 ? inj_delays_ : nullptr;
    // the per-week demand / delay tables are shared by every env: stage them in shared memory once per block
    // (a dependent global load per week, even an L1 hit, was the dominant stall of the first version)
    extern __shared__ int s_tab[];
    int* s_dem = s_tab;
    int* s_del = s_tab + sp.T;
    for (int i = threadIdx.x; i < sp.T; i += blockDim.x) { s_dem[i] = sp.demand[i]; s_del[i] = sp.delays[i]; }
    __syncthreads();
    ReturnStats rs; rs.init();
    if (env < n) {
        int inv[LMAX], back[LMAX], placed[LMAX], tr[RMAX][LMAX];
        int week = st.week[env];
        unsigned episode = (unsigned)st.episode[env];
        double ep_ret = st.ep_ret[env];
        const unsigned long long genv = (unsigned long long)(first_env + env);

        auto do_reset_state = [&]() {                                  // beergame_env.py:140-148
#pragma unroll
            for (int i = 0; i < LMAX; ++i) if (i < L) {
                inv[i] = sp.inv0[i]; back[i] = 0; placed[i] = sp.init_orders;
#pragma unroll
                for (int k = 0; k < RMAX; ++k) if (k < RING) tr[k][i] = (k < 2) ? sp.init_ship : 0;   // weeks 1..2 (:52)
            }
        };
        if (do_reset && (reset_mask == nullptr || reset_mask[env])) {
            ++episode; week = 0; ep_ret = 0.0;
            do_reset_state();
        } else {
#pragma unroll
            for (int i = 0; i < LMAX; ++i) if (i < L) {
                inv[i] = st.inv[(long long)i * n + env];
                back[i] = st.backlog[(long long)i * n + env];
                placed[i] = st.placed[(long long)i * n + env];
#pragma unroll
                for (int k = 0; k < RMAX; ++k) if (k < RING) tr[k][i] = st.transit[((long long)k * L + i) * n + env];
            }
        }
        const int off = v2 ? sp.max_stock : 0;                           // beergame2_env.py:111-112
        auto write_obs = [&](long long io) {
            if (obs == nullptr) return;
            if (STATIC && LMAX == 4) {
                int4 o = make_int4(off + inv[0] - back[0], off + inv[1] - back[1], off + inv[2] - back[2], off + inv[3] - back[3]);
                __stcs(reinterpret_cast<int4*>(obs) + io, o);        // streaming store: written once, never re-read here
            } else {
                for (int i = 0; i < L; ++i) obs[io * L + i] = off + inv[i] - back[i];
            }
        };
        if (n_steps == 0) write_obs(env);
        constexpr bool kFast4 = STATIC && LMAX == 4;
        const bool pipelined = kFast4 && actions != nullptr;
        const int4* act4 = reinterpret_cast<const int4*>(actions);

        // one week of one env; `a_now` is the (pre-loaded) action row of the 4-level fast path
        long long io = env;                                            // running index of (week s, this env) in the [T, n] outputs
        Philox4 blk = Philox4{0, 0, 0, 0}; unsigned blk_q = 0xFFFFFFFFu;   // cached random block of weeks 4q+1 .. 4q+4
        auto week_step = [&](int s, int4 a_now) {
            if (week >= sp.T) {                                        // finished, not auto-reset: frozen
                write_obs(io); reward[io] = 0; done[io] = 1;
                return;
            }
            int act[LMAX];
            if (actions == nullptr) {                                  // random policy, keyed (env, episode, week)
                Philox4 b0 = rng_block(key, genv, episode, (unsigned)(week + 1), STREAM_ACTION, 0);
                Philox4 b1 = rng_block(key, genv, episode, (unsigned)(week + 1), STREAM_ACTION, 1);
#pragma unroll
                for (int i = 0; i < LMAX; ++i)
                    if (i < L) act[i] = v2 ? uniform_int(pick_word(i < 4 ? b0 : b1, i & 3), 0, (unsigned)sp.max_order)
                                                        : uniform_int(pick_word(i < 4 ? b0 : b1, i & 3), -2, 7);
            } else if (STATIC && LMAX == 4) {
                act[0] = a_now.x; act[1] = a_now.y; act[2] = a_now.z; act[3] = a_now.w;
            } else {
                for (int i = 0; i < L; ++i) act[i] = actions[io * L + i];
            }
            ++week;
            // 1. receive what is due this week and advance the ring (:67-72)
#pragma unroll
            for (int i = 0; i < LMAX; ++i) if (i < L) {
                inv[i] += tr[0][i];
#pragma unroll
                for (int k = 0; k + 1 < RMAX; ++k) if (k + 1 < RING) tr[k][i] = tr[k + 1][i];
                tr[RING - 1][i] = 0;
            }
            // customer demand and this week's delay; when both are stochastic they share one Philox block
            // keyed (env, episode, week): word x -> demand, word y -> delay (randint(lo, hi): hi exclusive)
            int dem, delay;
            const bool rnd_dem = INP == 0 ? (!inj_demand && sp.dem_hi > sp.dem_lo) : INP == 2;
            const bool rnd_del = INP == 0 ? (!inj_delays && sp.del_hi > sp.del_lo) : INP == 2;
            // One Philox block feeds FOUR consecutive weeks of this (env, episode): block q = (week-1) >> 2, word (week-1) & 3,
            // demand from the low 16 bits, delay from the high 16 (multiply-shift; bias < span / 65536).  The block is kept
            // in registers and refilled every fourth week (a 37-byte env-week cannot afford ten Philox rounds per week).
            unsigned wk = 0u;
            if (rnd_dem || rnd_del) {
                const unsigned q = (unsigned)(week - 1) >> 2;
                if (q != blk_q) { blk = rng_block(key, genv, episode, q, STREAM_BG_DEMAND, 0); blk_q = q; }
                wk = pick_word(blk, (week - 1) & 3);
            }
            if (inj_demand) dem = inj_demand[(long long)(week - 1) * n + env];
            else if (rnd_dem) dem = sp.dem_lo + (int)(((wk & 0xFFFFu) * (unsigned)(sp.dem_hi - sp.dem_lo)) >> 16);
            else dem = s_dem[week - 1];
            if (inj_delays) {
                delay = inj_delays[(long long)(week - 1) * n + env];
                if (delay < 0 || delay > RING) {          // would land outside the in-transit ring: clamp and flag (scg_status)
                    if ((*(volatile int*)err & 4) == 0) atomicOr(err, 4);
                    delay = delay < 0 ? 0 : RING;
                }
            }
            else if (rnd_del) delay = sp.del_lo + (int)(((wk >> 16) * (unsigned)(sp.del_hi - sp.del_lo)) >> 16);
            else delay = s_del[week - 1];
            // 2. fill orders (:79-89)
            int incoming[LMAX], fill[LMAX], deliver[LMAX];
#pragma unroll
            for (int i = 0; i < LMAX; ++i) if (i < L) {
                incoming[i] = (i == 0) ? dem : placed[i - 1];
                fill[i] = incoming[i] + back[i];
                deliver[i] = min(inv[i], fill[i]);
            }
            // Where this week's shipments land: ONE delay for all levels, so the one-hot of the ring slot (and of "delay 0":
            // straight into the inventory) is built once per week and every booking is a multiply-add by 1 or 0 -- no branch on
            // a per-env random delay (lanes of a warp diverged on it), no compare + select per (slot, level).
            int oh0 = delay == 0 ? 1 : 0, oh[RMAX];
#pragma unroll
            for (int k = 0; k < RMAX; ++k) oh[k] = (k == delay - 1) ? 1 : 0;
            // ship downstream; delay 0 lands in the lower level's inventory before its own decrement (:93-96)
#pragma unroll
            for (int i = 0; i < LMAX; ++i) if (i + 1 < L) {
                inv[i] += deliver[i + 1] * oh0;
#pragma unroll
                for (int k = 0; k < RMAX; ++k) if (k < RING) tr[k][i] += deliver[i + 1] * oh[k];
            }
            // 3. inventory / backlog (:101-103)
#pragma unroll
            for (int i = 0; i < LMAX; ++i) if (i < L) { inv[i] -= deliver[i]; back[i] = fill[i] - deliver[i]; }
            // 4. the brewery's own order comes back after the delay (:111-114)
            inv[L - 1] += placed[L - 1] * oh0;
#pragma unroll
            for (int k = 0; k < RMAX; ++k) if (k < RING) tr[k][L - 1] += placed[L - 1] * oh[k];
            // 5. place orders (beergame_env.py:121 / beergame2_env.py:168)
#pragma unroll
            for (int i = 0; i < LMAX; ++i) if (i < L) placed[i] = v2 ? act[i] : incoming[i] + act[i];
            // 6. reward (:130 / beergame2_env.py:177-180)
            int r = 0;
#pragma unroll
            for (int i = 0; i < LMAX; ++i) if (i < L) {
                r -= sp.inv_cost * inv[i] + sp.backlog_cost * back[i];
                if (v2) r -= sp.penalty * (max(inv[i] - sp.max_stock, 0) + max(back[i] - sp.max_stock, 0));
            }
            ep_ret += (double)r;
            const bool fin = (week == sp.T);
            if (fin) {
                rs.add(ep_ret);
                if (auto_reset) { ++episode; week = 0; ep_ret = 0.0; blk_q = 0xFFFFFFFFu; do_reset_state(); }
            }
            write_obs(io);
            reward[io] = r;
            done[io] = fin ? 1 : 0;
        };

        if (pipelined) {
            // groups of BG_PF weeks, double-buffered: the rows of group g+1 are requested before group g is computed
            // and only copied into `cur` afterwards, so BG_PF independent 16-byte loads per lane stay in flight
            // depth: 3 weeks for the table-driven kernels (v2 0.684 -> 0.708, v0 0.700 -> 0.709 of the HBM peak), 2 where Philox
            // demand / delays need the registers (3 costs them 6 %, 4 spills everywhere; profiles/r2_beergame_prefetch_depth_ab.log)
            constexpr int BG_PF = INP == 1 ? SCG_BG_PF_TABLE : SCG_BG_PF;
            int4 cur[BG_PF], nxt[BG_PF];
#pragma unroll
            for (int u = 0; u < BG_PF; ++u) {
                cur[u] = make_int4(0, 0, 0, 0);
                if (u < n_steps) cur[u] = __ldcs(act4 + ((long long)u * n + env));
            }
            for (int g = 0; g < n_steps; g += BG_PF) {
#pragma unroll
                for (int u = 0; u < BG_PF; ++u) {
                    nxt[u] = make_int4(0, 0, 0, 0);
                    if (g + BG_PF + u < n_steps) nxt[u] = __ldcs(act4 + ((long long)(g + BG_PF + u) * n + env));
                }
#pragma unroll
                for (int u = 0; u < BG_PF; ++u)
                    if (g + u < n_steps) { week_step(g + u, cur[u]); io += n; }
#pragma unroll
                for (int u = 0; u < BG_PF; ++u) cur[u] = nxt[u];
            }
        } else {
            for (int s = 0; s < n_steps; ++s) { week_step(s, make_int4(0, 0, 0, 0)); io += n; }
        }
        st.week[env] = week; st.episode[env] = (int)episode; st.ep_ret[env] = ep_ret;
#pragma unroll
        for (int i = 0; i < LMAX; ++i) if (i < L) {
            st.inv[(long long)i * n + env] = inv[i];
            st.backlog[(long long)i * n + env] = back[i];
            st.placed[(long long)i * n + env] = placed[i];
#pragma unroll
            for (int k = 0; k < RMAX; ++k) if (k < RING) st.transit[((long long)k * L + i) * n + env] = tr[k][i];
        }
    }
    rs.flush(st.stats);
}

int bg_run(const BgRun& r) {
    const int threads = SCG_BG_THREADS;
    const unsigned blocks = (unsigned)((r.st.n + threads - 1) / threads);
    const int do_reset = r.op == SC_OP_RESET ? 1 : 0;
    const int n_steps = r.op == SC_OP_RESET ? 0 : r.n_steps;
    const size_t smem = (size_t)2 * r.sp.T * sizeof(int);
    if (smem > 48 * 1024) { set_error("beer game: more than 6144 weeks per episode not supported"); return SCG_EINVAL; }
#define SCG_BG_LAUNCH(LM, RM, ST, VER, INP)                                                                              \
    bg_kernel<LM, RM, ST, VER, INP><<<blocks, threads, smem, r.stream>>>(r.sp, r.st, r.inj_demand, r.inj_delays, r.key,       \
                                                                      r.first_env, r.auto_reset, n_steps, r.reset_mask,    \
                                                                      do_reset, r.actions, r.obs, r.reward, r.done, r.err)
    // where this launch's demand / delays come from: 1 = the spec's tables, 2 = Philox for both, 0 = anything else
    const bool rnd_dem = r.sp.dem_hi > r.sp.dem_lo, rnd_del = r.sp.del_hi > r.sp.del_lo;
    const int inp = (r.inj_demand || r.inj_delays) ? 0 : ((!rnd_dem && !rnd_del) ? 1 : ((rnd_dem && rnd_del) ? 2 : 0));
#define SCG_BG_PICK(RM)                                                                                                     \
    do {                                                                                                                    \
        if (inp == 1 && r.sp.version == 2) SCG_BG_LAUNCH(4, RM, true, 2, 1);                                                \
        else if (inp == 1) SCG_BG_LAUNCH(4, RM, true, 0, 1);                                                                \
        else if (inp == 2 && r.sp.version == 2) SCG_BG_LAUNCH(4, RM, true, 2, 2);                                           \
        else SCG_BG_LAUNCH(4, RM, true, -1, 0);                                                                             \
    } while (0)
    if (r.sp.L == 4 && r.sp.ring == 2) SCG_BG_PICK(2);
    else if (r.sp.L == 4 && r.sp.ring == 4) SCG_BG_PICK(4);
    else SCG_BG_LAUNCH(8, 8, false, -1, 0);
#undef SCG_BG_PICK
#undef SCG_BG_LAUNCH
    SCG_CUDA(cudaGetLastError());
    return SCG_OK;
}

}  // namespace scg
