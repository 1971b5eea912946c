// scg_sc_f64.cu -- fp64 instantiations (SCG_MODE_FAST_F64 and SCG_MODE_EXACT_F64).
// MUST be compiled with -fmad=false: numpy never fuses a*b+c, and bit-exact parity with the
// reference needs every product and every sum rounded separately (the build scripts assert it).
#include "scg_sc_host.cuh"

namespace scg {

int sc_match_variant_f64(const ScHostSpec& hs) {
    if (sc_is_layered(hs, 1, 1, 1, 1, 1, 2)) return 1;
    if (sc_is_layered(hs, 2, 2, 2, 2, 1, 2)) return 2;
    if (sc_is_layered(hs, 1, 1, 1, 3, 1, 2)) return 4;
    return 0;
}

int sc_run_f64(const ScRun& r) {
    if (r.op == SC_OP_ROLLOUT && r.mode == SCG_MODE_FAST_F64) {
        switch (r.variant) {
            case 1: return sc_launch_layered<double, double, 1, 1, 1, 1, 1, 2, 128, 3>(r);
            case 2: return sc_launch_layered<double, double, 2, 2, 2, 2, 1, 2, 128, 3>(r);
            case 4: return sc_launch_layered<double, double, 1, 1, 1, 3, 1, 2, 128, 3>(r);
            default: break;
        }
    }
    return sc_launch_generic<double, double>(r);
}

size_t sc_devspec_size_f64() { return sizeof(ScDevSpec<double>); }
int sc_build_devspec_f64(const ScHostSpec& hs, void** blob_out, void* devspec_out) {
    return sc_build_devspec_t<double>(hs, blob_out, devspec_out);
}

}  // namespace scg
