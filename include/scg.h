/* scg.h -- C ABI of the B200-native batched gym_supplychain simulator.
 *
 * This is the drop-in boundary for ONE hot path of caburu/gym-supplychain: the
 * reset()/step() dynamics of its environments, batched over millions of independent
 * envs on one GPU.  Everything here is `extern "C"`, plain pointers and sizes; no
 * torch / C++ types cross the boundary.  The library is loaded with ctypes (see
 * INTEGRATION.md for the binding a maintainer of the reference would add).
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   gym_supplychain/envs/supplychain_env.py:482-628   SupplyChainEnv.__init__  -> scg_sc_create
 *   gym_supplychain/envs/supplychain_env.py:630-682   SupplyChainEnv.reset     -> scg_reset
 *   gym_supplychain/envs/supplychain_env.py:703-748   SupplyChainEnv.step      -> scg_step / scg_rollout
 *   gym_supplychain/envs/supplychain_env.py:811-813   SupplyChainEnv.seed      -> scg_seed
 *   gym_supplychain/envs/beergame_env.py:11-56,66-156 BeerGameEnv              -> scg_bg_create / scg_step / scg_reset
 *   gym_supplychain/envs/beergame2_env.py:10-74,79-192 BeerGameEnv2            -> scg_bg_create(version=2) / ...
 *   gym_supplychain/envs/demands_generator.py:3-89    generate_demand          -> in-kernel Philox generators
 *                                                                               (or scg_inject for exact tables)
 *
 * Conventions
 *   - every function returns 0 on success or a negative SCG_E* code; scg_last_error()
 *     gives the message (thread-local).  Nothing throws across the ABI.
 *   - the CALLER owns every I/O buffer; the library owns only env state and the spec.
 *   - unless the name ends in _host, pointers are DEVICE pointers on the handle's GPU and
 *     all work is enqueued on the caller's stream (a cudaStream_t passed as void*, 0 =
 *     legacy default stream) with no hidden synchronisation.
 *   - batched layouts are env-major ("array of env rows", what a torch user expects):
 *       actions [n_envs, A]   obs [n_envs, O]   reward [n_envs]   done [n_envs] (uint8)
 *     and for rollouts   actions [T, n_envs, A]  obs [T, n_envs, O]  reward/done [T, n_envs].
 *   - element type of actions/obs/reward follows the handle's mode:
 *       SCG_MODE_FAST_F32 / SCG_MODE_MIXED -> float32;  SCG_MODE_FAST_F64 / SCG_MODE_EXACT_F64 -> float64;
 *       beer game -> int32 for all three.
 *   - one handle per GPU shard; a handle is not thread-safe, distinct handles are independent.
 */
#ifndef SCG_H_
#define SCG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCG_ABI_VERSION 1

#if defined(__GNUC__)
#define SCG_API __attribute__((visibility("default")))
#else
#define SCG_API
#endif

/* error codes */
#define SCG_OK            0
#define SCG_EINVAL       -1   /* bad argument / unsupported spec          */
#define SCG_ECUDA        -2   /* a CUDA runtime call failed               */
#define SCG_ENOMEM       -3   /* host or device allocation failed         */
#define SCG_ESTATE       -4   /* call not valid in the handle's state     */

/* arithmetic / state modes */
#define SCG_MODE_FAST_F32   0  /* fp32 state in registers, pre-summed lead-time ring, canonical obs   */
#define SCG_MODE_FAST_F64   1  /* same algorithm in fp64                                              */
#define SCG_MODE_EXACT_F64  2  /* fp64 + CPython-heapq-exact in-transit queues: bit-exact vs reference */
#define SCG_MODE_MIXED      3  /* fp32 actions/obs/reward in HBM, fp64 state and arithmetic on chip       */

/* demand generator kinds (demands_generator.py:20-31) */
#define SCG_DEMAND_UNIFORM         0  /* randint(lo, hi+1)                                   (:33-36) */
#define SCG_DEMAND_NORMAL          1  /* rint(clip(N((hi+lo)/2, std), lo, hi))               (:38-49) */
#define SCG_DEMAND_SIN_NORMAL      2  /* sinusoid + N(0,std), clipped, rounded               (:51-89) */
#define SCG_DEMAND_SIN_UNIFORM     3  /* sinusoid + randint(-3std, 3std+1), clipped, rounded (:76-77) */

typedef struct ScgEnv* ScgHandle;

/* Options common to every env family. */
typedef struct ScgOptions {
    int64_t  n_envs;          /* envs simulated by THIS handle (one GPU shard)                     */
    int64_t  first_env;       /* global index of env 0 of this shard: Philox streams are keyed by   */
                              /* the global index, so results do not depend on the number of GPUs   */
    uint64_t seed;            /* Philox key                                                          */
    int32_t  device;          /* CUDA device ordinal                                                 */
    int32_t  mode;            /* SCG_MODE_* (ignored by the beer game: always int32)                 */
    int32_t  auto_reset;      /* 1: an env that finishes an episode is reset inside the same step;   */
                              /*    the obs returned with done=1 is the first obs of the next episode*/
                              /* 0: finished envs freeze (reward 0, done 1) until scg_reset          */
    int32_t  flags;           /* SCG_FLAG_* bits                                                     */
} ScgOptions;

#define SCG_FLAG_FORCE_GENERIC 1   /* never pick a compile-time-topology kernel (tests / A-B timing)    */
#define SCG_FLAG_BUILD_INFO    2   /* supply chain: keep the reference's build_info statistics (implies the  */
                                   /* runtime-topology kernel); read them with scg_get_info                  */

/* Flat description of a SupplyChainEnv (what supplychain_env.py:499-626 builds from
 * `nodes_info` + kwargs).  N nodes in act order, P products, E directed edges (CSR by source
 * node, in the order of each node's `destinations`).  All arrays are HOST pointers, copied
 * by scg_sc_create. */
typedef struct ScgSupplyChainSpec {
    int32_t n_nodes, n_products, n_edges, total_time_steps;
    int32_t avg_leadtime, max_leadtime, stochastic_leadtimes, reserved;
    double  unmet_demand_cost, exceeded_stock_capacity_cost;
    double  exceeded_process_capacity_cost, exceeded_ship_capacity_cost;
    /* per node [N] */
    const int32_t* last_level;            /* 1 = retailer (serves customer demand)                */
    const int32_t* edge_begin;            /* [N+1] CSR offsets into the edge arrays               */
    const double*  processing_capacity;   /* > 0 marks a factory                                  */
    /* per node x product [N*P], node-major */
    const double*  initial_stock;
    const double*  stock_capacity;
    const double*  stock_cost;
    const double*  supply_capacity;       /* > 0: the node has a SUPPLY action for that product   */
    const double*  supply_cost;
    const double*  processing_ratio;
    const double*  processing_cost;
    /* per edge [E] */
    const int32_t* edge_dest;             /* destination node index                               */
    const double*  edge_ship_capacity;
    const double*  edge_dest_cost;        /* [E*P], edge-major: cost of product p on edge e       */
    /* initial in-transit material, CSR per (node, product) in the reference's push order
     * (initial_supply then initial_shipments, supplychain_env.py:405-412) */
    const int32_t* init_begin;            /* [N*P+1]                                              */
    const int32_t* init_time;             /* arrival period (1-based)                             */
    const double*  init_amount;
    /* demand model per product [P] */
    const int32_t* demand_kind;           /* SCG_DEMAND_*                                         */
    const double*  demand_lo;             /* also the obs normalisation range (:771)              */
    const double*  demand_hi;
    const double*  demand_std;
    const double*  demand_peaks;
    const double*  demand_minavg;
    const double*  demand_maxavg;
} ScgSupplyChainSpec;

/* Beer game (beergame_env.py / beergame2_env.py).  Level 0 = retailer, L-1 = brewery. */
typedef struct ScgBeerGameSpec {
    int32_t version;                      /* 0: action is a delta on the incoming order; 2: the order */
    int32_t levels;                       /* 1..8                                                     */
    int32_t weeks;                        /* episode length T                                         */
    int32_t inv_cost, backlog_cost;
    int32_t max_stock, exceeded_capacity_penalty;   /* v2 only                                        */
    int32_t initial_shipment, initial_orders;
    int32_t demand_lo, demand_hi;         /* stochastic demand randint(lo, hi) (hi EXCLUSIVE) if hi>lo */
    int32_t delay_lo, delay_hi;           /* stochastic delays randint(lo, hi) (hi EXCLUSIVE) if hi>lo */
    int32_t max_order;                    /* v2: MultiDiscrete([max_order]*levels) action space (random policy) */
    const int32_t* initial_inventory;     /* [levels]                                                 */
    const int32_t* customer_demand;       /* [weeks]   (ignored when stochastic)                      */
    const int32_t* shipment_delays;       /* [weeks]   delay in force at week 1..T (ignored when stochastic);
                                             the initial pipeline depth is always 2 (beergame_env.py:39) */
} ScgBeerGameSpec;

/* Shapes a caller needs to size its buffers. */
typedef struct ScgInfo {
    int64_t n_envs;
    int32_t family;            /* 0 supply chain, 1 beer game                                */
    int32_t mode;
    int32_t n_actions;         /* A */
    int32_t n_obs;             /* O */
    int32_t n_demands;         /* R*P values per period (beer game: 1)                       */
    int32_t n_leadtimes;       /* K lead-time values per period (beer game: 1 delay)         */
    int32_t total_time_steps;  /* T */
    int32_t elem_size;         /* bytes per action/obs/reward element                        */
    int32_t ring_depth;        /* FAST modes: slots per (node, product)                      */
    int32_t heap_capacity;     /* EXACT mode: entries per (node, product) queue              */
    int32_t kernel_variant;    /* 0 generic runtime-topology kernel, >0 a compile-time one   */
                               /* (100: instantiated at run time for this chain's shape;     */
                               /*  101: ditto, node-parallel kernel -- several lanes per env) */
    int32_t info_values;       /* 16*P when SCG_FLAG_BUILD_INFO, else 0                      */
    int64_t state_bytes;       /* size of the blob for scg_get_state / scg_set_state         */
} ScgInfo;

SCG_API const char* scg_last_error(void);
SCG_API int  scg_abi_version(void);

SCG_API int  scg_sc_create(const ScgSupplyChainSpec* spec, const ScgOptions* opt, ScgHandle* out);
SCG_API int  scg_bg_create(const ScgBeerGameSpec* spec, const ScgOptions* opt, ScgHandle* out);
/* Chains without a pre-built compile-time kernel (ScgInfo.kernel_variant == 100): the kernel template is instantiated
 * for the chain's shape with NVRTC at the first rollout (10-20 s, cached under $SCG_JIT_CACHE or <library dir>/_jit;
 * SCG_JIT=0 disables it).  scg_sc_precompile does that compilation ahead of time; it needs no GPU.  Returns 0 when a
 * compile-time kernel exists or was compiled, SCG_EINVAL when the spec only runs on the runtime-topology kernel. */
SCG_API int  scg_sc_precompile(const ScgSupplyChainSpec* spec, int32_t mode, int32_t in_kernel_actions);
SCG_API void scg_destroy(ScgHandle h);
SCG_API int  scg_info(ScgHandle h, ScgInfo* out);

/* New Philox key; the episode counters restart (SupplyChainEnv.seed, :811-813).  Host-synchronous: the device is
 * synchronised before and after, so work already enqueued on any stream keeps the old key and nothing new overtakes. */
SCG_API int  scg_seed(ScgHandle h, uint64_t seed);

/* Replace the on-device generators by caller-provided tables for the NEXT episodes
 * (parity tests: "identical injected demand sequences").  Device pointers, int32:
 *   demand    [T+1, n_demands, n_envs]   (beer game: [T, 1, n_envs])
 *   leadtimes [T,   n_leadtimes, n_envs] (beer game: delay per week [T, 1, n_envs])
 * Either may be NULL (= keep generating that input).  The tables are NOT copied: they
 * must stay alive until replaced or the handle is destroyed. */
SCG_API int  scg_inject(ScgHandle h, const int32_t* demand, const int32_t* leadtimes);

/* Device-side checks.  Kernels validate what they cannot trust: an injected lead time outside [1, max_leadtime]
 * (the reference draws clip(1 + Poisson, 1, max), supplychain_env.py:670-672) or an injected beer-game delay outside
 * [0, ring_depth] is CLAMPED (never an out-of-bounds access) and an in-transit queue that overflows its capacity drops
 * the entry; either raises a sticky bit on the handle.  scg_status synchronises the device and returns SCG_OK or
 * SCG_ESTATE (scg_last_error names the check); clear != 0 resets the bits.  The *_host calls, being synchronous,
 * report the same condition through their own return code. */
SCG_API int  scg_status(ScgHandle h, int32_t clear);

/* reset(): `mask` NULL = every env, else uint8[n_envs] (non-zero = reset).  `obs` may be NULL. */
SCG_API int  scg_reset(ScgHandle h, const uint8_t* mask, void* obs, void* stream);

/* step(): one period for every env. */
SCG_API int  scg_step(ScgHandle h, const void* actions, void* obs, void* reward, uint8_t* done, void* stream);

/* Fused T-step rollout: env state stays on-chip between periods.
 *   actions NULL -> uniform random actions drawn in-kernel (Philox, keyed (env, episode, t)).
 *   obs NULL     -> observations are not written (reward/done still are). */
SCG_API int  scg_rollout(ScgHandle h, int32_t T, const void* actions, void* obs, void* reward, uint8_t* done,
                 void* stream);

/* Same calls with HOST buffers (pageable or pinned): the library stages them through its own
 * device buffers (H2D, kernel, D2H) and returns when the results are in the host buffers.  They run on the library's
 * own streams, ordered BEHIND every earlier device-pointer call on this handle (scg_reset / scg_step / scg_rollout /
 * scg_set_state record an event on the caller's stream that the host calls wait for). */
SCG_API int  scg_reset_host(ScgHandle h, const uint8_t* mask, void* obs);
SCG_API int  scg_step_host(ScgHandle h, const void* actions, void* obs, void* reward, uint8_t* done);
SCG_API int  scg_rollout_host(ScgHandle h, int32_t T, const void* actions, void* obs, void* reward, uint8_t* done);

/* Raw env state (device blob of ScgInfo.state_bytes; layout documented in DESIGN.md). */
SCG_API int  scg_get_state(ScgHandle h, void* blob, void* stream);
SCG_API int  scg_set_state(ScgHandle h, const void* blob, void* stream);

/* Episode-return statistics of the episodes finished since the last clear:
 * out[5] = {count, sum, sum of squares, min, max} as doubles (device pointer; the values a
 * multi-GPU caller all-reduces: SUM on the first three, MIN, MAX).  clear != 0 zeroes them after. */
SCG_API int  scg_return_stats(ScgHandle h, double* out, int32_t clear, void* stream);

/* build_info statistics of the current episode (supplychain_env.py:684-695, :750-760, :793-796):
 * out = double[16*P][n_envs] (device pointer), rows 0..8P-1 cumulative costs and rows 8P..16P-1 cumulative units,
 * key-major then product, keys in the reference's order: stock, stock_pen, supply, process, process_pen, ship,
 * ship_pen, unmet_dem.  The cumulative reward ('rewards') is ep_ret in the state blob.  Needs SCG_FLAG_BUILD_INFO. */
SCG_API int  scg_get_info(ScgHandle h, double* out, void* stream);

/* Byte offsets of the arrays inside the state blob (all env-minor: element i of env e at [i*n_envs + e]).
 * supply chain: {ep_ret f64[n], heap_amount f64[NP*cap*n], stock Real[NP*n], slot Real[NP*ring*n],
 *                t i32[n], episode i32[n], heap_len i32[NP*n], heap_time i32[NP*cap*n]}
 * beer game:    {ep_ret f64[n], week i32[n], episode i32[n], inventory i32[L*n], backlog i32[L*n],
 *                orders_placed i32[L*n], in_transit i32[ring*L*n], 0} */
SCG_API int  scg_state_layout(ScgHandle h, int64_t offsets[8]);

/* Number of kernel launches this handle has enqueued so far (bench bookkeeping). */
SCG_API int64_t scg_launch_count(ScgHandle h);

#ifdef __cplusplus
}
#endif
#endif /* SCG_H_ */
