// scg_sc_mixed.cu -- SCG_MODE_MIXED: fp32 I/O (actions, obs, reward) with fp64 state and arithmetic.
//
// Why it exists: inventories and in-transit amounts are integrated over hundreds of periods, so
// fp32 rounding noise on them random-walks to ~2e-6 relative on late-episode rewards (measured,
// DESIGN.md).  Keeping the integrating path in fp64 while streaming fp32 through HBM keeps the
// 561 B/env-step traffic of the fp32 mode and brings rewards within ~1e-7 of the reference.
// Compiled WITH FMA contraction (bit-exactness is the job of scg_sc_f64.cu).
#include "scg_sc_host.cuh"

#ifndef SCG_V1_MINB_F64
#define SCG_V1_MINB_F64 5
#endif
#ifndef SCG_V2_MINB_F64
#define SCG_V2_MINB_F64 4
#endif

#ifndef SCG_V6_MINB_F64
#define SCG_V6_MINB_F64 3
#endif
#ifndef SCG_V3_THREADS_F64
#define SCG_V3_THREADS_F64 64
#endif
#ifndef SCG_V3_MINB_F64
#define SCG_V3_MINB_F64 4
#endif

// Launch shape of a pre-built kernel: (threads per block, minimum blocks per SM).  SCG_SHAPE1 is the one-warp form of a shape:
// the same warps per SM in blocks of 32 threads, so that every bulk-copy address derives from blockIdx and lives in uniform
// registers (scg_sc_static.cuh) -- measured faster wherever the kernel is not already HBM-bound (profiles/r2_onewarp_blocks_ab.log).
// SCG_ONE_WARP (experiment builds): 1 = every shape one-warp, -1 = none, 0 = the per-kernel choice below.
#ifndef SCG_ONE_WARP
#define SCG_ONE_WARP 0
#endif
#define SCG_SHAPE_SEL(w1, threads, minb) ((w1) ? 32 : (threads)), ((w1) ? ((minb) * (threads) / 32 > 32 ? 32 : (minb) * (threads) / 32) : (minb))
#define SCG_SHAPE(threads, minb) SCG_SHAPE_SEL(SCG_ONE_WARP == 1, threads, minb)
#define SCG_SHAPE1(threads, minb) SCG_SHAPE_SEL(SCG_ONE_WARP != -1, threads, minb)

namespace scg {

int sc_match_variant_mixed(const ScHostSpec& hs) {
    if (sc_is_layered(hs, 1, 1, 1, 1, 1, 2)) return 1;
    if (sc_is_layered(hs, 2, 2, 2, 2, 1, 2)) return 2;
    if (sc_is_layered(hs, 3, 3, 3, 3, 2, 2)) return 3;
    if (sc_is_layered(hs, 1, 1, 1, 3, 1, 2)) return 4;
    if (sc_is_layered(hs, 2, 2, 2, 2, 2, 2)) return 5;
    if (sc_is_layered(hs, 2, 2, 2, 2, 1, 2, 4, true)) return 6;   // stochastic lead times avg 2 / max 4 (supplychain-v0 demo chain)
    return 0;
}

int sc_run_mixed(const ScRun& r) {
    if (r.op == SC_OP_ROLLOUT) {
        switch (r.variant) {
            case 1: return sc_launch_layered<double, float, 1, 1, 1, 1, 1, 2, SCG_SHAPE1(128, SCG_V1_MINB_F64)>(r);
            case 2: return sc_launch_layered<double, float, 2, 2, 2, 2, 1, 2, SCG_SHAPE1(128, SCG_V2_MINB_F64)>(r);
            case 3: return sc_launch_layered<double, float, 3, 3, 3, 3, 2, 2, SCG_SHAPE1(SCG_V3_THREADS_F64, SCG_V3_MINB_F64)>(r);
            case 4: return sc_launch_layered<double, float, 1, 1, 1, 3, 1, 2, SCG_SHAPE1(128, 3)>(r);
            case 5: return sc_launch_layered<double, float, 2, 2, 2, 2, 2, 2, SCG_SHAPE1(64, 4)>(r);
            case 6: return sc_launch_layered<double, float, 2, 2, 2, 2, 1, 2, SCG_SHAPE1(128, SCG_V6_MINB_F64), 4, true>(r);
            default: break;
        }
    }
    return sc_launch_generic<double, float>(r);
}

}  // namespace scg
