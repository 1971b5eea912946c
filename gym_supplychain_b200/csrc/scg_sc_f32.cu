// scg_sc_f32.cu -- fp32 instantiations of the supply-chain kernels (SCG_MODE_FAST_F32).
// Compiled with the default FMA contraction: this is the throughput path; its results agree
// with the reference to ~1e-7 relative per step (tests state the tolerance).
#include "scg_sc_host.cuh"

// launch shapes (threads per block, min blocks per SM = register cap) of the small variants: their state is a few
// registers, so what matters is enough resident warps to hide the per-period TMA round trips
#ifndef SCG_V1_MINB_F32
#define SCG_V1_MINB_F32 8
#endif
#ifndef SCG_V2_MINB_F32
#define SCG_V2_MINB_F32 4
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

// static layered instantiations:  id -> (S, F, W, R, P, L)
//   1: 1-1-1-1 P=1 L=2  (supplychain-linear-v0)
//   2: 2-2-2-2 P=1 L=2  (supplychain-2perstage-v0 / sc-2perstage-v0)
//   3: 3-3-3-3 P=2 L=2  (supplychain-ntom-v0 / sc-Nperstage-multiproduct-v0 defaults)
//   4: 1-1-1-3 P=1 L=2  (supplychain-oneonen-v0)
//   5: 2-2-2-2 P=2 L=2  (sc-2perstage-multiproduct-v0 defaults)
int sc_match_variant_f32(const ScHostSpec& hs) {
    if (sc_is_layered(hs, 1, 1, 1, 1, 1, 2)) return 1;
    if (sc_is_layered(hs, 2, 2, 2, 2, 1, 2)) return 2;
    if (sc_is_layered(hs, 3, 3, 3, 3, 2, 2)) return 3;
    if (sc_is_layered(hs, 1, 1, 1, 3, 1, 2)) return 4;
    if (sc_is_layered(hs, 2, 2, 2, 2, 2, 2)) return 5;
    if (sc_is_layered(hs, 2, 2, 2, 2, 1, 2, 4, true)) return 6;   // stochastic lead times avg 2 / max 4 (supplychain-v0 demo chain)
    return 0;
}

int sc_run_f32(const ScRun& r) {
    if (r.op == SC_OP_ROLLOUT && r.mode == SCG_MODE_FAST_F32) {
        switch (r.variant) {
            case 1: return sc_launch_layered<float, float, 1, 1, 1, 1, 1, 2, SCG_SHAPE1(128, SCG_V1_MINB_F32)>(r);
            case 2: return sc_launch_layered<float, float, 2, 2, 2, 2, 1, 2, SCG_SHAPE(128, SCG_V2_MINB_F32)>(r);
            case 3: return sc_launch_layered<float, float, 3, 3, 3, 3, 2, 2, SCG_SHAPE(128, 3)>(r);
            case 4: return sc_launch_layered<float, float, 1, 1, 1, 3, 1, 2, SCG_SHAPE1(128, 4)>(r);
            case 5: return sc_launch_layered<float, float, 2, 2, 2, 2, 2, 2, SCG_SHAPE(128, 3)>(r);
            case 6: return sc_launch_layered<float, float, 2, 2, 2, 2, 1, 2, SCG_SHAPE1(128, 4), 4, true>(r);
            default: break;
        }
    }
    return sc_launch_generic<float, float>(r);
}

size_t sc_devspec_size_f32() { return sizeof(ScDevSpec<float>); }
int sc_build_devspec_f32(const ScHostSpec& hs, void** blob_out, void* devspec_out) {
    return sc_build_devspec_t<float>(hs, blob_out, devspec_out);
}

}  // namespace scg
