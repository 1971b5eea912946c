// scg_api.cu -- the extern "C" boundary declared in include/scg.h.
//
// Host-side only: validates and flattens specs, owns the device allocations (spec blob, env
// state, statistics, staging buffers of the *_host calls) and dispatches to the kernels in
// scg_sc_f32.cu / scg_sc_f64.cu / scg_bg.cu.  No CPU implementation of the dynamics exists
// in this library: if CUDA is unavailable every call fails with SCG_ECUDA.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "scg_internal.h"
#include "scg_sc_device.cuh"

namespace scg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace scg

using namespace scg;

struct ScgEnv {
    int family = 0;                 // 0 supply chain, 1 beer game
    ScgOptions opt{};
    ScgInfo info{};
    RngKey key{};
    long long launches = 0;
    // supply chain
    ScHostSpec hs;
    void* spec_blob = nullptr;
    alignas(16) unsigned char devspec[sizeof(ScDevSpec<double>) > sizeof(ScDevSpec<float>) ? sizeof(ScDevSpec<double>) : sizeof(ScDevSpec<float>)];
    ScState<float> st32{};
    ScState<double> st64{};
    int variant = 0;
    ScJit* jit = nullptr;           // run-time instantiated kernels (variant == SC_VARIANT_JIT)
    std::vector<unsigned char> consts_cache;   // kernel-parameter image of the spec for the compile-time kernel
    // beer game
    BgHostSpec bhs;
    BgDevSpec bsp{};
    BgState bst{};
    void* bg_blob = nullptr;
    // common
    void* state_blob = nullptr;
    size_t state_bytes = 0;
    int64_t layout[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double* stats = nullptr;        // [5] live  + [5] init pattern (+ the error word behind them)
    int* err = nullptr;             // sticky SCG_ERR_* bits raised by kernels (pinned, device-mapped; scg_status)
    cudaEvent_t ev_user = nullptr;  // last device-pointer call on a caller stream: the *_host calls wait for it
    double* info_buf = nullptr;     // build_info: [16*P][n]
    const int32_t* inj_demand = nullptr;
    const int32_t* inj_leadtimes = nullptr;
    // staging for *_host calls
    void* stage = nullptr;
    size_t stage_bytes = 0;
    void* bounce = nullptr;         // pinned host buffer of the small-batch scg_step_host path
    size_t bounce_bytes = 0;
    cudaStream_t hstream = nullptr;                 // kernels of the *_host calls
    cudaStream_t s_in = nullptr, s_out = nullptr;   // H2D / D2H copy streams (PCIe is full duplex)
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

static int check_handle(ScgHandle h) {
    if (!h) { set_error("null handle"); return SCG_EINVAL; }
    return SCG_OK;
}

static int set_device(ScgHandle h) {
    SCG_CUDA(cudaSetDevice(h->opt.device));
    return SCG_OK;
}

static RngKey key_from_seed(uint64_t seed) {
    RngKey k;
    k.k0 = (uint32_t)seed;
    k.k1 = (uint32_t)(seed >> 32);
    return k;
}

// ---------------------------------------------------------------------------------------
// supply chain: spec -> host spec
// ---------------------------------------------------------------------------------------
static int sc_flatten(const ScgSupplyChainSpec* s, ScHostSpec& hs) {
    if (!s) { set_error("null spec"); return SCG_EINVAL; }
    const int N = s->n_nodes, P = s->n_products, E = s->n_edges;
    if (N < 1 || N > 4096 || P < 1 || P > SC_MAXP || E < 0) { set_error("bad sizes N=%d P=%d E=%d (P<=%d)", N, P, E, SC_MAXP); return SCG_EINVAL; }
    if (s->total_time_steps < 1) { set_error("total_time_steps must be >= 1"); return SCG_EINVAL; }
    if (s->avg_leadtime < 1 || s->max_leadtime < 1) { set_error("lead times must be >= 1"); return SCG_EINVAL; }
    if (s->stochastic_leadtimes && s->max_leadtime < s->avg_leadtime) { set_error("max_leadtime < avg_leadtime"); return SCG_EINVAL; }
    const void* req[] = {s->last_level, s->edge_begin, s->processing_capacity, s->initial_stock, s->stock_capacity, s->stock_cost,
                         s->supply_capacity, s->supply_cost, s->processing_ratio, s->processing_cost, s->init_begin,
                         s->demand_kind, s->demand_lo, s->demand_hi, s->demand_std, s->demand_peaks, s->demand_minavg, s->demand_maxavg};
    for (const void* p : req) if (!p) { set_error("spec has a null array"); return SCG_EINVAL; }
    if (E > 0 && (!s->edge_dest || !s->edge_ship_capacity || !s->edge_dest_cost)) { set_error("spec has null edge arrays"); return SCG_EINVAL; }
    hs.N = N; hs.P = P; hs.E = E; hs.T = s->total_time_steps;
    hs.avgL = s->avg_leadtime; hs.maxL = s->max_leadtime; hs.stochL = s->stochastic_leadtimes ? 1 : 0;
    hs.c_unmet = s->unmet_demand_cost; hs.c_xstock = s->exceeded_stock_capacity_cost;
    hs.c_xproc = s->exceeded_process_capacity_cost; hs.c_xship = s->exceeded_ship_capacity_cost;
    hs.last.assign(s->last_level, s->last_level + N);
    hs.ebeg.assign(s->edge_begin, s->edge_begin + N + 1);
    hs.proc_cap.assign(s->processing_capacity, s->processing_capacity + N);
    const int NP = N * P;
    hs.stock0.assign(s->initial_stock, s->initial_stock + NP);
    hs.cap.assign(s->stock_capacity, s->stock_capacity + NP);
    hs.hold.assign(s->stock_cost, s->stock_cost + NP);
    hs.sup_cap.assign(s->supply_capacity, s->supply_capacity + NP);
    hs.sup_cost.assign(s->supply_cost, s->supply_cost + NP);
    hs.ratio.assign(s->processing_ratio, s->processing_ratio + NP);
    hs.proc_cost.assign(s->processing_cost, s->processing_cost + NP);
    if (hs.ebeg[0] != 0 || hs.ebeg[N] != E) { set_error("edge_begin must run from 0 to n_edges"); return SCG_EINVAL; }
    for (int n = 0; n < N; ++n) {
        const int D = hs.ebeg[n + 1] - hs.ebeg[n];
        if (D < 0 || D > SC_MAXD) { set_error("node %d has %d destinations (max %d)", n, D, SC_MAXD); return SCG_EINVAL; }
    }
    if (E > 0) {
        hs.edst.assign(s->edge_dest, s->edge_dest + E);
        hs.ecap.assign(s->edge_ship_capacity, s->edge_ship_capacity + E);
        hs.ecost.assign(s->edge_dest_cost, s->edge_dest_cost + (size_t)E * P);
        for (int e = 0; e < E; ++e) if (hs.edst[e] < 0 || hs.edst[e] >= N) { set_error("edge %d: bad destination", e); return SCG_EINVAL; }
    }
    hs.ibeg.assign(s->init_begin, s->init_begin + NP + 1);
    const int n_init = hs.ibeg[NP];
    if (hs.ibeg[0] != 0 || n_init < 0) { set_error("bad init_begin"); return SCG_EINVAL; }
    if (n_init > 0) {
        if (!s->init_time || !s->init_amount) { set_error("null init arrays"); return SCG_EINVAL; }
        hs.itime.assign(s->init_time, s->init_time + n_init);
        hs.iamt.assign(s->init_amount, s->init_amount + n_init);
    }
    hs.dkind.assign(s->demand_kind, s->demand_kind + P);
    hs.dlo.assign(s->demand_lo, s->demand_lo + P); hs.dhi.assign(s->demand_hi, s->demand_hi + P);
    hs.dstd.assign(s->demand_std, s->demand_std + P); hs.dpeaks.assign(s->demand_peaks, s->demand_peaks + P);
    hs.dminavg.assign(s->demand_minavg, s->demand_minavg + P); hs.dmaxavg.assign(s->demand_maxavg, s->demand_maxavg + P);
    for (int p = 0; p < P; ++p) {
        if (hs.dlo[p] == hs.dhi[p]) { set_error("fixed demand (min == max) is not supported (supplychain_env.py:591-595)"); return SCG_EINVAL; }
        if (hs.dkind[p] < 0 || hs.dkind[p] > 3) { set_error("bad demand kind"); return SCG_EINVAL; }
    }

    // ---- derived quantities (supplychain_env.py:597-621, :147, :204-206, :419-420) ----
    hs.act_off.assign(N, 0); hs.lt_off.assign(N, 0); hs.ret_idx.assign(N, -1); hs.is_sup.assign(N, 0);
    hs.max_ship.assign(NP, 0.0);
    std::vector<int> indeg(N, 0);
    int A = 0, LT = 0, K = 0, R = 0;
    for (int n = 0; n < N; ++n) {
        int n_sup = 0, n_shipprod = 0;
        const int D = hs.ebeg[n + 1] - hs.ebeg[n];
        for (int p = 0; p < P; ++p) {
            if (hs.sup_cap[n * P + p] > 0) ++n_sup;
            if (D > 0 && hs.cap[n * P + p] > 0) ++n_shipprod;
        }
        hs.is_sup[n] = n_sup > 0;
        hs.act_off[n] = A; hs.lt_off[n] = LT;
        A += n_sup + n_shipprod * D;
        LT += n_sup + (n_shipprod * D) / P;
        K += (n_sup > 0 ? P : 0) + D;
        if (hs.last[n]) hs.ret_idx[n] = R++;
        if (hs.is_sup[n]) for (int p = 0; p < P; ++p) hs.max_ship[n * P + p] += hs.sup_cap[n * P + p];
        if (hs.last[n] && D > 0) { set_error("node %d: a last_level node cannot have destinations", n); return SCG_EINVAL; }
    }
    for (int e = 0; e < E; ++e) {
        ++indeg[hs.edst[e]];
        for (int p = 0; p < P; ++p) hs.max_ship[hs.edst[e] * P + p] += hs.ecap[e];
    }
    hs.R = R; hs.RP = R * P; hs.A = A; hs.K = hs.stochL ? K : 0;
    hs.O = R * P + N * P * (1 + hs.avgL) + 1;
    int ring = hs.maxL > hs.avgL ? hs.maxL : hs.avgL, heapcap = 1;
    for (int np = 0; np < NP; ++np) {
        const int n = np / P;
        const int n_init_np = hs.ibeg[np + 1] - hs.ibeg[np];
        if (n_init_np < 0) { set_error("bad init_begin"); return SCG_EINVAL; }
        for (int k = hs.ibeg[np]; k < hs.ibeg[np + 1]; ++k) {
            if (hs.itime[k] < 1) { set_error("initial arrival periods are 1-based"); return SCG_EINVAL; }
            if (hs.itime[k] > ring) ring = hs.itime[k];
        }
        // every source can have maxL entries in flight, plus the one it pushes in the current period BEFORE this
        // node pops its arrivals (upstream nodes act first, supplychain_env.py:714)
        const int cap = (indeg[n] + (hs.is_sup[n] ? 1 : 0)) * ((hs.maxL > hs.avgL ? hs.maxL : hs.avgL) + 1) + n_init_np;
        if (cap > heapcap) heapcap = cap;
    }
    if (ring > 64) { set_error("ring depth %d too large", ring); return SCG_EINVAL; }
    hs.ring = ring; hs.heapcap = heapcap;
    if (R == 0) { set_error("the chain has no last_level (retailer) node"); return SCG_EINVAL; }
    return SCG_OK;
}

static bool is_f64(int mode) { return mode != SCG_MODE_FAST_F32; }

static int sc_alloc_state(ScgEnv* h) {
    const ScHostSpec& hs = h->hs;
    const size_t n = (size_t)h->opt.n_envs;
    if (h->opt.flags & SCG_FLAG_BUILD_INFO) {
        SCG_CUDA(cudaMalloc(&h->info_buf, (size_t)16 * hs.P * n * sizeof(double)));
        SCG_CUDA(cudaMemset(h->info_buf, 0, (size_t)16 * hs.P * n * sizeof(double)));
    }
    const size_t NP = (size_t)hs.N * hs.P;
    const bool exact = h->opt.mode == SCG_MODE_EXACT_F64;
    const size_t rs = is_f64(h->opt.mode) ? 8 : 4;
    size_t off = 0;
    auto take = [&off](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_ret = take(n * 8);
    const size_t o_hamt = take(exact ? NP * hs.heapcap * n * 8 : 0);
    const size_t o_stock = take(NP * n * rs);
    const size_t o_slot = take(exact ? 0 : NP * hs.ring * n * rs);
    const size_t o_t = take(n * 4), o_ep = take(n * 4);
    const size_t o_hlen = take(exact ? NP * n * 4 : 0);
    const size_t o_htime = take(exact ? NP * hs.heapcap * n * 4 : 0);
    h->state_bytes = off;
    const size_t lay[8] = {o_ret, o_hamt, o_stock, o_slot, o_t, o_ep, o_hlen, o_htime};
    for (int i = 0; i < 8; ++i) h->layout[i] = (int64_t)lay[i];
    SCG_CUDA(cudaMalloc(&h->state_blob, off));
    SCG_CUDA(cudaMemset(h->state_blob, 0, off));
    unsigned char* b = (unsigned char*)h->state_blob;
    if (is_f64(h->opt.mode)) {
        ScState<double>& s = h->st64;
        s.n = (long long)n; s.t = (int*)(b + o_t); s.episode = (int*)(b + o_ep); s.ep_ret = (double*)(b + o_ret);
        s.stock = (double*)(b + o_stock); s.slot = (double*)(b + o_slot);
        s.hlen = (int*)(b + o_hlen); s.htime = (int*)(b + o_htime); s.hamt = (double*)(b + o_hamt);
        s.stats = h->stats; s.info = h->info_buf;
    } else {
        ScState<float>& s = h->st32;
        s.n = (long long)n; s.t = (int*)(b + o_t); s.episode = (int*)(b + o_ep); s.ep_ret = (double*)(b + o_ret);
        s.stock = (float*)(b + o_stock); s.slot = (float*)(b + o_slot);
        s.hlen = nullptr; s.htime = nullptr; s.hamt = nullptr;
        s.stats = h->stats; s.info = h->info_buf;
    }
    // episode = -1 so that the first reset starts episode 0
    SCG_CUDA(cudaMemset(b + o_ep, 0xFF, n * 4));
    return SCG_OK;
}

static int alloc_stats(ScgEnv* h) {
    SCG_CUDA(cudaMalloc(&h->stats, 10 * sizeof(double)));
    const double init[10] = {0, 0, 0, 1e300, -1e300, 0, 0, 0, 1e300, -1e300};
    SCG_CUDA(cudaMemcpy(h->stats, init, sizeof(init), cudaMemcpyHostToDevice));
    // the error word lives in pinned, device-mapped host memory: kernels only touch it when something IS wrong, and the
    // host reads it without a copy
    SCG_CUDA(cudaHostAlloc((void**)&h->err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *h->err = 0;
    SCG_CUDA(cudaEventCreateWithFlags(&h->ev_user, cudaEventDisableTiming));
    return SCG_OK;
}

static int sc_dispatch(ScgEnv* h, int op, int n_steps, const unsigned char* mask, const void* actions, void* obs,
                       void* reward, unsigned char* done, cudaStream_t stream) {
    ScRun r{};
    r.hs = &h->hs; r.dev_spec = h->devspec;
    r.state = is_f64(h->opt.mode) ? (const void*)&h->st64 : (const void*)&h->st32;
    r.mode = h->opt.mode; r.variant = h->variant;
    r.inj_demand = h->inj_demand; r.inj_leadtimes = h->inj_leadtimes;
    r.key = h->key; r.first_env = h->opt.first_env; r.auto_reset = h->opt.auto_reset;
    r.op = op; r.n_steps = n_steps; r.reset_mask = mask; r.actions = actions; r.obs = obs; r.reward = reward; r.done = done;
    r.stream = stream; r.err = h->err; r.consts_cache = &h->consts_cache; r.device = h->opt.device;
    ++h->launches;
    if (op == SC_OP_ROLLOUT && h->variant == SC_VARIANT_JIT) {
        if (sc_jit_run(&h->jit, r) == SCG_OK) return SCG_OK;
        h->variant = 0; h->info.kernel_variant = 0; r.variant = 0;      // reported once on stderr; the generic kernel takes over
    }
    if (h->opt.mode == SCG_MODE_MIXED) return sc_run_mixed(r);
    return is_f64(h->opt.mode) ? sc_run_f64(r) : sc_run_f32(r);
}

static int bg_dispatch(ScgEnv* h, int op, int n_steps, const unsigned char* mask, const void* actions, void* obs,
                       void* reward, unsigned char* done, cudaStream_t stream) {
    BgRun r{};
    r.sp = h->bsp; r.st = h->bst; r.inj_demand = h->inj_demand; r.inj_delays = h->inj_leadtimes;
    r.key = h->key; r.first_env = h->opt.first_env; r.auto_reset = h->opt.auto_reset;
    r.op = op; r.n_steps = n_steps; r.reset_mask = mask;
    r.actions = (const int32_t*)actions; r.obs = (int32_t*)obs; r.reward = (int32_t*)reward; r.done = done; r.stream = stream; r.err = h->err;
    ++h->launches;
    return bg_run(r);
}

static int dispatch(ScgEnv* h, int op, int n_steps, const unsigned char* mask, const void* actions, void* obs, void* reward,
                    unsigned char* done, cudaStream_t stream) {
    const int rc = h->family == 0 ? sc_dispatch(h, op, n_steps, mask, actions, obs, reward, done, stream)
                                  : bg_dispatch(h, op, n_steps, mask, actions, obs, reward, done, stream);
    // work enqueued on a CALLER stream: the *_host entry points (own non-blocking streams) order themselves behind it
    if (rc == SCG_OK && stream != h->hstream && h->ev_user) SCG_CUDA(cudaEventRecord(h->ev_user, stream));
    return rc;
}

// Sticky error bits raised by kernels -> message + SCG_ESTATE.  The caller has synchronised the work in question.
static int report_device_errors(ScgEnv* h) {
    const int e = *(volatile int*)h->err;
    if (e == 0) return SCG_OK;
    set_error("device-side checks failed:%s%s%s",
              (e & 1) ? " an injected lead time is outside [1, max_leadtime] (clamped);" : "",
              (e & 2) ? " an in-transit queue overflowed its capacity (entry dropped);" : "",
              (e & 4) ? " an injected beer-game delay is outside [0, ring depth] (clamped);" : "");
    return SCG_ESTATE;
}

// ---------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------
extern "C" {

const char* scg_last_error(void) { return g_err; }
int scg_abi_version(void) { return SCG_ABI_VERSION; }

static int check_options(const ScgOptions* opt) {
    if (!opt) { set_error("null options"); return SCG_EINVAL; }
    if (opt->n_envs < 1) { set_error("n_envs must be >= 1"); return SCG_EINVAL; }
    if (opt->mode < 0 || opt->mode > 3) { set_error("bad mode %d", opt->mode); return SCG_EINVAL; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s): this library has no CPU fallback", cudaGetErrorString(e));
        return SCG_ECUDA;
    }
    if (opt->device < 0 || opt->device >= ndev) { set_error("bad device %d (have %d)", opt->device, ndev); return SCG_EINVAL; }
    return SCG_OK;
}

int scg_sc_create(const ScgSupplyChainSpec* spec, const ScgOptions* opt, ScgHandle* out) {
    if (!out) { set_error("null out"); return SCG_EINVAL; }
    *out = nullptr;
    ScgEnv* h = new (std::nothrow) ScgEnv();
    if (!h) { set_error("out of host memory"); return SCG_ENOMEM; }
    int rc = sc_flatten(spec, h->hs);                       // spec errors first: they need no GPU
    if (rc == SCG_OK) rc = check_options(opt);
    if (rc != SCG_OK) { delete h; return rc; }
    h->family = 0; h->opt = *opt; h->key = key_from_seed(opt->seed);
    rc = set_device(h);
    if (rc == SCG_OK) rc = alloc_stats(h);
    if (rc == SCG_OK) rc = is_f64(opt->mode) ? sc_build_devspec_f64(h->hs, &h->spec_blob, h->devspec)
                                             : sc_build_devspec_f32(h->hs, &h->spec_blob, h->devspec);
    if (rc == SCG_OK) rc = sc_alloc_state(h);
    if (rc != SCG_OK) { scg_destroy(h); return rc; }
    h->variant = opt->mode == SCG_MODE_FAST_F32 ? sc_match_variant_f32(h->hs)
               : opt->mode == SCG_MODE_FAST_F64 ? sc_match_variant_f64(h->hs)
               : opt->mode == SCG_MODE_MIXED ? sc_match_variant_mixed(h->hs) : 0;
    if (const char* e = getenv("SCG_JIT_FORCE")) if (e[0] == '1') h->variant = 0;      // experiments: ignore the pre-built kernels
    if (h->variant == 0 && sc_jit_match(h->hs, opt->mode)) h->variant = SC_VARIANT_JIT;
    if (opt->flags & (SCG_FLAG_FORCE_GENERIC | SCG_FLAG_BUILD_INFO)) h->variant = 0;
    ScgInfo& I = h->info;
    I.n_envs = opt->n_envs; I.family = 0; I.mode = opt->mode; I.n_actions = h->hs.A; I.n_obs = h->hs.O;
    I.n_demands = h->hs.RP; I.n_leadtimes = h->hs.K; I.total_time_steps = h->hs.T; I.elem_size = (is_f64(opt->mode) && opt->mode != SCG_MODE_MIXED) ? 8 : 4;
    I.ring_depth = h->hs.ring; I.heap_capacity = h->hs.heapcap; I.kernel_variant = h->variant; I.state_bytes = (int64_t)h->state_bytes;
    if (h->variant == SC_VARIANT_JIT && sc_jit_split_lanes(h->hs, opt->mode) > 0) I.kernel_variant = SC_VARIANT_JIT_SPLIT;
    I.info_values = (opt->flags & SCG_FLAG_BUILD_INFO) ? 16 * h->hs.P : 0;
    rc = dispatch(h, SC_OP_RESET, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
    if (rc == SCG_OK) { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { set_error("initial reset failed: %s", cudaGetErrorString(e)); rc = SCG_ECUDA; } }
    if (rc != SCG_OK) { scg_destroy(h); return rc; }
    *out = h;
    return SCG_OK;
}

int scg_sc_precompile(const ScgSupplyChainSpec* spec, int32_t mode, int32_t in_kernel_actions) {
    if (mode < 0 || mode > 3) { set_error("bad mode %d", mode); return SCG_EINVAL; }
    ScHostSpec hs;
    int rc = sc_flatten(spec, hs);
    if (rc != SCG_OK) return rc;
    const int v = mode == SCG_MODE_FAST_F32 ? sc_match_variant_f32(hs) : mode == SCG_MODE_FAST_F64 ? sc_match_variant_f64(hs)
                : mode == SCG_MODE_MIXED ? sc_match_variant_mixed(hs) : 0;
    const char* force = getenv("SCG_JIT_FORCE");
    if (v > 0 && !(force && force[0] == '1')) return SCG_OK;    // already in the library
    return sc_jit_precompile(hs, mode, in_kernel_actions != 0);
}

int scg_bg_create(const ScgBeerGameSpec* s, const ScgOptions* opt, ScgHandle* out) {
    if (!out) { set_error("null out"); return SCG_EINVAL; }
    *out = nullptr;
    if (!s) { set_error("null spec"); return SCG_EINVAL; }
    if (s->version != 0 && s->version != 2) { set_error("beer game version must be 0 or 2"); return SCG_EINVAL; }
    if (s->levels < 1 || s->levels > 8) { set_error("levels must be in 1..8"); return SCG_EINVAL; }
    if (s->weeks < 1) { set_error("weeks must be >= 1"); return SCG_EINVAL; }
    if (!s->initial_inventory) { set_error("null initial_inventory"); return SCG_EINVAL; }
    const bool sd = s->demand_hi > s->demand_lo, sl = s->delay_hi > s->delay_lo;
    if (!sd && !s->customer_demand) { set_error("null customer_demand"); return SCG_EINVAL; }
    if (!sl && !s->shipment_delays) { set_error("null shipment_delays"); return SCG_EINVAL; }
    int rc = check_options(opt);
    if (rc != SCG_OK) return rc;
    ScgEnv* h = new (std::nothrow) ScgEnv();
    if (!h) { set_error("out of host memory"); return SCG_ENOMEM; }
    h->family = 1; h->opt = *opt; h->key = key_from_seed(opt->seed);
    BgHostSpec& b = h->bhs;
    b.s = *s;
    b.inv0.assign(s->initial_inventory, s->initial_inventory + s->levels);
    b.demand.assign(s->weeks, 0); b.delays.assign(s->weeks, 2);
    if (!sd) b.demand.assign(s->customer_demand, s->customer_demand + s->weeks);
    if (!sl) b.delays.assign(s->shipment_delays, s->shipment_delays + s->weeks);
    int ring = 2;
    if (sl) { if (s->delay_lo < 0) { delete h; set_error("negative delay"); return SCG_EINVAL; } if (s->delay_hi - 1 > ring) ring = s->delay_hi - 1; }
    else for (int d : b.delays) { if (d < 0) { delete h; set_error("negative delay"); return SCG_EINVAL; } if (d > ring) ring = d; }
    if (ring > 2 && ring <= 4) ring = 4;
    if (ring > 8) { delete h; set_error("shipment delay > 8 not supported"); return SCG_EINVAL; }
    b.ring = ring;
    rc = set_device(h);
    if (rc == SCG_OK) rc = alloc_stats(h);
    if (rc != SCG_OK) { scg_destroy(h); return rc; }
    // spec blob
    const int L = s->levels, T = s->weeks;
    std::vector<int> blob(L + 2 * T);
    memcpy(blob.data(), b.inv0.data(), L * 4); memcpy(blob.data() + L, b.demand.data(), T * 4); memcpy(blob.data() + L + T, b.delays.data(), T * 4);
    cudaError_t e = cudaMalloc(&h->bg_blob, blob.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(h->bg_blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("cudaMalloc/cudaMemcpy failed: %s", cudaGetErrorString(e)); scg_destroy(h); return SCG_ECUDA; }
    BgDevSpec& d = h->bsp;
    d.version = s->version; d.L = L; d.T = T; d.inv_cost = s->inv_cost; d.backlog_cost = s->backlog_cost;
    d.max_stock = s->max_stock; d.penalty = s->exceeded_capacity_penalty; d.init_ship = s->initial_shipment; d.init_orders = s->initial_orders;
    d.dem_lo = s->demand_lo; d.dem_hi = s->demand_hi; d.del_lo = s->delay_lo; d.del_hi = s->delay_hi; d.ring = ring;
    d.max_order = s->max_order > 0 ? s->max_order : 30;
    d.inv0 = (const int*)h->bg_blob; d.demand = d.inv0 + L; d.delays = d.demand + T;
    // state
    const size_t n = (size_t)opt->n_envs;
    size_t off = 0;
    auto take = [&off](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_ret = take(n * 8), o_week = take(n * 4), o_ep = take(n * 4), o_inv = take(n * 4 * L), o_back = take(n * 4 * L),
                 o_pl = take(n * 4 * L), o_tr = take(n * 4 * L * ring);
    h->state_bytes = off;
    { const size_t lay[8] = {o_ret, o_week, o_ep, o_inv, o_back, o_pl, o_tr, 0}; for (int i = 0; i < 8; ++i) h->layout[i] = (int64_t)lay[i]; }
    e = cudaMalloc(&h->state_blob, off);
    if (e == cudaSuccess) e = cudaMemset(h->state_blob, 0, off);
    if (e == cudaSuccess) e = cudaMemset((unsigned char*)h->state_blob + o_ep, 0xFF, n * 4);
    if (e != cudaSuccess) { set_error("state allocation failed: %s", cudaGetErrorString(e)); scg_destroy(h); return SCG_ECUDA; }
    unsigned char* sb = (unsigned char*)h->state_blob;
    BgState& st = h->bst;
    st.n = (long long)n; st.week = (int*)(sb + o_week); st.episode = (int*)(sb + o_ep); st.ep_ret = (double*)(sb + o_ret);
    st.inv = (int*)(sb + o_inv); st.backlog = (int*)(sb + o_back); st.placed = (int*)(sb + o_pl); st.transit = (int*)(sb + o_tr);
    st.stats = h->stats;
    ScgInfo& I = h->info;
    I.n_envs = opt->n_envs; I.family = 1; I.mode = -1; I.n_actions = L; I.n_obs = L; I.n_demands = 1; I.n_leadtimes = 1;
    I.total_time_steps = T; I.elem_size = 4; I.ring_depth = ring; I.heap_capacity = 0;
    I.kernel_variant = (L == 4 && (ring == 2 || ring == 4)) ? 1 : 0; I.state_bytes = (int64_t)off;
    rc = dispatch(h, SC_OP_RESET, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
    if (rc == SCG_OK) { e = cudaDeviceSynchronize(); if (e != cudaSuccess) { set_error("initial reset failed: %s", cudaGetErrorString(e)); rc = SCG_ECUDA; } }
    if (rc != SCG_OK) { scg_destroy(h); return rc; }
    *out = h;
    return SCG_OK;
}

void scg_destroy(ScgHandle h) {
    if (!h) return;
    cudaSetDevice(h->opt.device);
    if (h->hstream) { cudaStreamSynchronize(h->hstream); cudaStreamDestroy(h->hstream); }
    if (h->s_in) { cudaStreamSynchronize(h->s_in); cudaStreamDestroy(h->s_in); }
    if (h->s_out) { cudaStreamSynchronize(h->s_out); cudaStreamDestroy(h->s_out); }
    for (int i = 0; i < 2; ++i) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    sc_jit_free(h->jit);
    if (h->ev_user) cudaEventDestroy(h->ev_user);
    if (h->err) cudaFreeHost(h->err);
    if (h->bounce) cudaFreeHost(h->bounce);
    cudaFree(h->info_buf); cudaFree(h->stage); cudaFree(h->state_blob); cudaFree(h->spec_blob); cudaFree(h->bg_blob); cudaFree(h->stats);
    delete h;
}

int scg_info(ScgHandle h, ScgInfo* out) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (!out) { set_error("null out"); return SCG_EINVAL; }
    *out = h->info;
    return SCG_OK;
}

int scg_seed(ScgHandle h, uint64_t seed) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    h->key = key_from_seed(seed);
    int* ep = h->family == 0 ? (is_f64(h->opt.mode) ? h->st64.episode : h->st32.episode) : h->bst.episode;
    // a rare, host-synchronous call: everything enqueued on ANY stream (the caller's, the handle's own non-blocking ones)
    // finishes with the old key and episode counters, and the counters are rewritten before anything new can start
    SCG_CUDA(cudaDeviceSynchronize());
    SCG_CUDA(cudaMemset(ep, 0xFF, (size_t)h->opt.n_envs * 4));
    SCG_CUDA(cudaDeviceSynchronize());
    return SCG_OK;
}

int scg_inject(ScgHandle h, const int32_t* demand, const int32_t* leadtimes) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    h->inj_demand = demand;
    h->inj_leadtimes = leadtimes;
    return SCG_OK;
}

int scg_reset(ScgHandle h, const uint8_t* mask, void* obs, void* stream) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    return dispatch(h, SC_OP_RESET, 0, mask, nullptr, obs, nullptr, nullptr, (cudaStream_t)stream);
}

int scg_step(ScgHandle h, const void* actions, void* obs, void* reward, uint8_t* done, void* stream) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (!actions || !reward || !done) { set_error("step: actions, reward and done are required"); return SCG_EINVAL; }
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    return dispatch(h, SC_OP_ROLLOUT, 1, nullptr, actions, obs, reward, done, (cudaStream_t)stream);
}

int scg_rollout(ScgHandle h, int32_t T, const void* actions, void* obs, void* reward, uint8_t* done, void* stream) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (T < 1) { set_error("rollout: T must be >= 1"); return SCG_EINVAL; }
    if (!reward || !done) { set_error("rollout: reward and done are required"); return SCG_EINVAL; }
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    return dispatch(h, SC_OP_ROLLOUT, T, nullptr, actions, obs, reward, done, (cudaStream_t)stream);
}

// ---- host-buffer variants ----------------------------------------------------------------
static int ensure_stage(ScgEnv* h, size_t bytes) {
    if (!h->hstream) {
        SCG_CUDA(cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
        SCG_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        SCG_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            SCG_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            SCG_CUDA(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming));
            SCG_CUDA(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
        }
    }
    if (h->stage_bytes >= bytes) return SCG_OK;
    if (h->stage) { SCG_CUDA(cudaDeviceSynchronize()); cudaFree(h->stage); h->stage = nullptr; h->stage_bytes = 0; }
    SCG_CUDA(cudaMalloc(&h->stage, bytes));
    h->stage_bytes = bytes;
    return SCG_OK;
}

int scg_reset_host(ScgHandle h, const uint8_t* mask, void* obs) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    const size_t n = (size_t)h->opt.n_envs, es = (size_t)h->info.elem_size;
    const size_t b_obs = align_up(n * h->info.n_obs * es, 256), b_mask = align_up(n, 256);
    rc = ensure_stage(h, b_obs + b_mask);
    if (rc != SCG_OK) return rc;
    unsigned char* d = (unsigned char*)h->stage;
    SCG_CUDA(cudaStreamWaitEvent(h->hstream, h->ev_user, 0));
    if (mask) SCG_CUDA(cudaMemcpyAsync(d + b_obs, mask, n, cudaMemcpyHostToDevice, h->hstream));
    rc = dispatch(h, SC_OP_RESET, 0, mask ? d + b_obs : nullptr, nullptr, obs ? d : nullptr, nullptr, nullptr, h->hstream);
    if (rc != SCG_OK) return rc;
    if (obs) SCG_CUDA(cudaMemcpyAsync(obs, d, n * h->info.n_obs * es, cudaMemcpyDeviceToHost, h->hstream));
    SCG_CUDA(cudaStreamSynchronize(h->hstream));
    return report_device_errors(h);
}

int scg_rollout_host(ScgHandle h, int32_t T, const void* actions, void* obs, void* reward, uint8_t* done) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (T < 1 || !reward || !done) { set_error("rollout_host: T >= 1, reward and done are required"); return SCG_EINVAL; }
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    const size_t n = (size_t)h->opt.n_envs, es = (size_t)h->info.elem_size;
    const size_t A = h->info.n_actions, O = h->info.n_obs;
    // The rollout is cut into chunks of `chunk` periods that flow through a 3-stage pipeline on three streams
    // (H2D of chunk i+1 | kernel of chunk i | D2H of chunk i-1) with double-buffered staging, so that the two
    // PCIe directions and the kernel overlap.  Env state lives in the handle, so chunking does not change results.
    const size_t per_step = n * ((actions ? A : 0) * es + (obs ? O : 0) * es + es + 1);
    size_t chunk = ((size_t)T + 7) / 8;
    const size_t max_chunk = per_step ? ((size_t)1 << 30) / per_step : (size_t)T;
    if (chunk > max_chunk) chunk = max_chunk;
    if (chunk < 1) chunk = 1;
    const size_t b_act = align_up(actions ? chunk * n * A * es : 0, 256), b_obs = align_up(obs ? chunk * n * O * es : 0, 256);
    const size_t b_rew = align_up(chunk * n * es, 256), b_done = align_up(chunk * n, 256);
    const size_t b_set = b_act + b_obs + b_rew + b_done;
    rc = ensure_stage(h, 2 * b_set);
    if (rc != SCG_OK) return rc;
    SCG_CUDA(cudaStreamWaitEvent(h->hstream, h->ev_user, 0));
    size_t i = 0;
    for (size_t s0 = 0; s0 < (size_t)T; s0 += chunk, ++i) {
        const int b = (int)(i & 1);
        const size_t c = ((size_t)T - s0) < chunk ? ((size_t)T - s0) : chunk;
        unsigned char* d_act = (unsigned char*)h->stage + (size_t)b * b_set;
        unsigned char* d_obs = d_act + b_act;
        unsigned char* d_rew = d_obs + b_obs;
        unsigned char* d_done = d_rew + b_rew;
        if (actions) {
            if (i >= 2) SCG_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_k[b], 0));        // kernel i-2 has consumed this buffer
            SCG_CUDA(cudaMemcpyAsync(d_act, (const unsigned char*)actions + s0 * n * A * es, c * n * A * es,
                                     cudaMemcpyHostToDevice, h->s_in));
            SCG_CUDA(cudaEventRecord(h->ev_in[b], h->s_in));
            SCG_CUDA(cudaStreamWaitEvent(h->hstream, h->ev_in[b], 0));
        }
        if (i >= 2) SCG_CUDA(cudaStreamWaitEvent(h->hstream, h->ev_out[b], 0));      // D2H i-2 has drained this buffer
        rc = dispatch(h, SC_OP_ROLLOUT, (int)c, nullptr, actions ? d_act : nullptr, obs ? d_obs : nullptr, d_rew, d_done, h->hstream);
        if (rc != SCG_OK) return rc;
        SCG_CUDA(cudaEventRecord(h->ev_k[b], h->hstream));
        SCG_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_k[b], 0));
        if (obs) SCG_CUDA(cudaMemcpyAsync((unsigned char*)obs + s0 * n * O * es, d_obs, c * n * O * es, cudaMemcpyDeviceToHost, h->s_out));
        SCG_CUDA(cudaMemcpyAsync((unsigned char*)reward + s0 * n * es, d_rew, c * n * es, cudaMemcpyDeviceToHost, h->s_out));
        SCG_CUDA(cudaMemcpyAsync(done + s0 * n, d_done, c * n, cudaMemcpyDeviceToHost, h->s_out));
        SCG_CUDA(cudaEventRecord(h->ev_out[b], h->s_out));
    }
    SCG_CUDA(cudaStreamSynchronize(h->s_out));
    SCG_CUDA(cudaStreamSynchronize(h->hstream));
    SCG_CUDA(cudaStreamSynchronize(h->s_in));
    return report_device_errors(h);
}

int scg_step_host(ScgHandle h, const void* actions, void* obs, void* reward, uint8_t* done) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (!actions || !reward || !done) { set_error("step_host: actions, reward and done are required"); return SCG_EINVAL; }
    const size_t n = (size_t)h->opt.n_envs, es = (size_t)h->info.elem_size;
    const size_t A = h->info.n_actions, O = h->info.n_obs;
    const size_t b_act = align_up(n * A * es, 256), b_obs = align_up(n * O * es, 256), b_rew = align_up(n * es, 256), b_done = align_up(n, 256);
    const size_t total = b_act + b_obs + b_rew + b_done;
    if (total > ((size_t)1 << 20)) return scg_rollout_host(h, 1, actions, obs, reward, done);
    // small batches (the drop-in single env): latency path.  One stream, one pinned bounce buffer, two copies, one
    // synchronisation: host -> [actions] -> kernel -> [obs | reward | done] -> host.
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    rc = ensure_stage(h, 2 * total);
    if (rc != SCG_OK) return rc;
    if (h->bounce_bytes < total) {
        if (h->bounce) cudaFreeHost(h->bounce);
        h->bounce = nullptr; h->bounce_bytes = 0;
        SCG_CUDA(cudaMallocHost(&h->bounce, total));
        h->bounce_bytes = total;
    }
    unsigned char* hb = (unsigned char*)h->bounce;
    unsigned char* d = (unsigned char*)h->stage;
    SCG_CUDA(cudaStreamWaitEvent(h->hstream, h->ev_user, 0));
    memcpy(hb, actions, n * A * es);
    if (total <= ((size_t)16 << 10)) {
        // a handful of envs: the kernel reads the actions from and writes its results to the pinned buffer directly
        // (unified addressing makes cudaMallocHost memory device-accessible), so a step is one launch and one
        // synchronisation -- no copy calls at all
        rc = dispatch(h, SC_OP_ROLLOUT, 1, nullptr, hb, obs ? hb + b_act : nullptr, hb + b_act + b_obs, hb + b_act + b_obs + b_rew, h->hstream);
        if (rc != SCG_OK) return rc;
    } else {
        SCG_CUDA(cudaMemcpyAsync(d, hb, n * A * es, cudaMemcpyHostToDevice, h->hstream));
        rc = dispatch(h, SC_OP_ROLLOUT, 1, nullptr, d, obs ? d + b_act : nullptr, d + b_act + b_obs, d + b_act + b_obs + b_rew, h->hstream);
        if (rc != SCG_OK) return rc;
        SCG_CUDA(cudaMemcpyAsync(hb + b_act, d + b_act, b_obs + b_rew + b_done, cudaMemcpyDeviceToHost, h->hstream));
    }
    SCG_CUDA(cudaStreamSynchronize(h->hstream));
    if (obs) memcpy(obs, hb + b_act, n * O * es);
    memcpy(reward, hb + b_act + b_obs, n * es);
    memcpy(done, hb + b_act + b_obs + b_rew, n);
    return report_device_errors(h);
}

int scg_get_state(ScgHandle h, void* blob, void* stream) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (!blob) { set_error("null blob"); return SCG_EINVAL; }
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    SCG_CUDA(cudaMemcpyAsync(blob, h->state_blob, h->state_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SCG_OK;
}

int scg_set_state(ScgHandle h, const void* blob, void* stream) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (!blob) { set_error("null blob"); return SCG_EINVAL; }
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    SCG_CUDA(cudaMemcpyAsync(h->state_blob, blob, h->state_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    SCG_CUDA(cudaEventRecord(h->ev_user, (cudaStream_t)stream));
    return SCG_OK;
}

int scg_return_stats(ScgHandle h, double* out, int32_t clear, void* stream) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (!out) { set_error("null out"); return SCG_EINVAL; }
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    SCG_CUDA(cudaMemcpyAsync(out, h->stats, 5 * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (clear) SCG_CUDA(cudaMemcpyAsync(h->stats, h->stats + 5, 5 * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SCG_OK;
}

int scg_get_info(ScgHandle h, double* out, void* stream) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (!out) { set_error("null out"); return SCG_EINVAL; }
    if (!h->info_buf) { set_error("handle was not created with SCG_FLAG_BUILD_INFO"); return SCG_ESTATE; }
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    SCG_CUDA(cudaMemcpyAsync(out, h->info_buf, (size_t)16 * h->hs.P * (size_t)h->opt.n_envs * sizeof(double),
                             cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SCG_OK;
}

int scg_state_layout(ScgHandle h, int64_t offsets[8]) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    if (!offsets) { set_error("null offsets"); return SCG_EINVAL; }
    for (int i = 0; i < 8; ++i) offsets[i] = h->layout[i];
    return SCG_OK;
}

int64_t scg_launch_count(ScgHandle h) { return h ? h->launches : 0; }

int scg_status(ScgHandle h, int32_t clear) {
    if (check_handle(h) != SCG_OK) return SCG_EINVAL;
    int rc = set_device(h);
    if (rc != SCG_OK) return rc;
    SCG_CUDA(cudaDeviceSynchronize());
    rc = report_device_errors(h);
    if (clear) *h->err = 0;
    return rc;
}

}  // extern "C"
