// scg_internal.h -- host-side structures shared by the translation units of libscg.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/scg.h"
#include "scg_common.cuh"

namespace scg {

void set_error(const char* fmt, ...);

#define SCG_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ::scg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SCG_ECUDA;                                                                  \
        }                                                                                      \
    } while (0)

// Host copy of a SupplyChainEnv spec plus everything derived from it.
struct ScHostSpec {
    int N = 0, P = 0, E = 0, T = 0, avgL = 0, maxL = 0, stochL = 0;
    double c_unmet = 0, c_xstock = 0, c_xproc = 0, c_xship = 0;
    std::vector<int> last, ebeg, edst, ibeg, itime, dkind;
    std::vector<double> proc_cap, stock0, cap, hold, sup_cap, sup_cost, ratio, proc_cost, ecap, ecost, iamt;
    std::vector<double> dlo, dhi, dstd, dpeaks, dminavg, dmaxavg;
    // derived
    int R = 0, RP = 0, A = 0, O = 0, K = 0, ring = 0, heapcap = 0;
    std::vector<int> act_off, lt_off, ret_idx, is_sup;
    std::vector<double> max_ship;
};

enum ScOp { SC_OP_RESET = 0, SC_OP_ROLLOUT = 1 };

// One request to a supply-chain translation unit (f32 or f64).
struct ScRun {
    const ScHostSpec* hs;
    const void* dev_spec;       // ScDevSpec<Real>* (host copy of the struct, pointers are device pointers)
    const void* state;          // ScState<Real>*   (idem)
    int mode;                   // SCG_MODE_*
    int variant;                // 0 generic, >0 static layered instantiation (index into the TU's table)
    const int32_t* inj_demand;
    const int32_t* inj_leadtimes;
    RngKey key;
    long long first_env;
    int auto_reset;
    int op;
    int n_steps;
    const unsigned char* reset_mask;
    const void* actions;
    void* obs;
    void* reward;
    unsigned char* done;
    cudaStream_t stream;
    int* err;                   // device word of sticky SCG_ERR_* bits (scg_status)
    std::vector<unsigned char>* consts_cache;   // per handle: the kernel-parameter image of the spec, built on the first launch
    int device;
};

// Implemented in scg_sc_f32.cu / scg_sc_f64.cu
int sc_match_variant_f32(const ScHostSpec& hs);
int sc_match_variant_f64(const ScHostSpec& hs);
int sc_run_f32(const ScRun& r);
int sc_run_f64(const ScRun& r);
int sc_match_variant_mixed(const ScHostSpec& hs);
int sc_run_mixed(const ScRun& r);
size_t sc_devspec_size_f32();
size_t sc_devspec_size_f64();
// Build the device spec: uploads arrays into one device blob, fills the ScDevSpec<Real> struct (host memory).
int sc_build_devspec_f32(const ScHostSpec& hs, void** blob_out, void* devspec_out);
int sc_build_devspec_f64(const ScHostSpec& hs, void** blob_out, void* devspec_out);

// Run-time instantiation of the layered kernel for shapes without a pre-built one (scg_jit.cu)
struct ScJit;
bool sc_jit_match(const ScHostSpec& hs, int mode);      // layered chain + NVRTC and the driver entry points are available
int sc_jit_split_lanes(const ScHostSpec& hs, int mode);  // > 0: the chain gets the node-parallel kernel with that many lanes per env
int sc_jit_run(ScJit** cache, const ScRun& r);          // SCG_OK, or an error after which the caller uses the generic kernel
void sc_jit_free(ScJit* cache);
int sc_jit_precompile(const ScHostSpec& hs, int mode, bool random_act);   // NVRTC only, no GPU needed
constexpr int SC_VARIANT_JIT = 100;
constexpr int SC_VARIANT_JIT_SPLIT = 101;     // reported in ScgInfo.kernel_variant only (dispatch uses SC_VARIANT_JIT)

// Beer game
struct BgHostSpec {
    ScgBeerGameSpec s;
    std::vector<int> inv0, demand, delays;
    int ring = 2;
};
struct BgDevSpec {
    int version, L, T, inv_cost, backlog_cost, max_stock, penalty, init_ship, init_orders;
    int dem_lo, dem_hi, del_lo, del_hi, ring, max_order;
    const int* inv0; const int* demand; const int* delays;
};
struct BgState {
    long long n;
    int* week; int* episode; double* ep_ret;
    int* inv; int* backlog; int* placed;     // [L][n]
    int* transit;                            // [ring][L][n]  slot k = arriving at week+1+k
    double* stats;
};
struct BgRun {
    BgDevSpec sp; BgState st;
    const int32_t* inj_demand; const int32_t* inj_delays;
    RngKey key; long long first_env; int auto_reset;
    int op; int n_steps; const unsigned char* reset_mask;
    const int32_t* actions; int32_t* obs; int32_t* reward; unsigned char* done;
    cudaStream_t stream;
    int* err;
};
int bg_run(const BgRun& r);

}  // namespace scg
