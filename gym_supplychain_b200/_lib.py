"""ctypes binding of libscg.so (include/scg.h).  There is NO Python/CPU fallback: if the CUDA
extension is missing or cannot be loaded, importing the env classes still works (so that specs
and spaces can be inspected on a CPU box) but creating an env raises ScgError loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCG_LIB") or os.path.join(HERE, "libscg.so")     # SCG_LIB: an experiment build (build.py --tag)

MODE_FAST_F32, MODE_FAST_F64, MODE_EXACT_F64, MODE_MIXED = 0, 1, 2, 3
MODES = {"fast_f32": MODE_FAST_F32, "fast_f64": MODE_FAST_F64, "exact_f64": MODE_EXACT_F64, "mixed": MODE_MIXED}
FLAG_FORCE_GENERIC, FLAG_BUILD_INFO = 1, 2
INFO_KEYS = ("stock", "stock_pen", "supply", "process", "process_pen", "ship", "ship_pen", "unmet_dem")
DEMAND_UNIFORM, DEMAND_NORMAL, DEMAND_SIN_NORMAL, DEMAND_SIN_UNIFORM = 0, 1, 2, 3


class ScgError(RuntimeError):
    pass


class ScgOptions(C.Structure):
    _fields_ = [("n_envs", C.c_int64), ("first_env", C.c_int64), ("seed", C.c_uint64), ("device", C.c_int32),
                ("mode", C.c_int32), ("auto_reset", C.c_int32), ("flags", C.c_int32)]


_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)


class ScgSupplyChainSpec(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_products", C.c_int32), ("n_edges", C.c_int32), ("total_time_steps", C.c_int32),
                ("avg_leadtime", C.c_int32), ("max_leadtime", C.c_int32), ("stochastic_leadtimes", C.c_int32),
                ("reserved", C.c_int32),
                ("unmet_demand_cost", C.c_double), ("exceeded_stock_capacity_cost", C.c_double),
                ("exceeded_process_capacity_cost", C.c_double), ("exceeded_ship_capacity_cost", C.c_double),
                ("last_level", _pi), ("edge_begin", _pi), ("processing_capacity", _pd),
                ("initial_stock", _pd), ("stock_capacity", _pd), ("stock_cost", _pd), ("supply_capacity", _pd),
                ("supply_cost", _pd), ("processing_ratio", _pd), ("processing_cost", _pd),
                ("edge_dest", _pi), ("edge_ship_capacity", _pd), ("edge_dest_cost", _pd),
                ("init_begin", _pi), ("init_time", _pi), ("init_amount", _pd),
                ("demand_kind", _pi), ("demand_lo", _pd), ("demand_hi", _pd), ("demand_std", _pd),
                ("demand_peaks", _pd), ("demand_minavg", _pd), ("demand_maxavg", _pd)]


class ScgBeerGameSpec(C.Structure):
    _fields_ = [("version", C.c_int32), ("levels", C.c_int32), ("weeks", C.c_int32), ("inv_cost", C.c_int32),
                ("backlog_cost", C.c_int32), ("max_stock", C.c_int32), ("exceeded_capacity_penalty", C.c_int32),
                ("initial_shipment", C.c_int32), ("initial_orders", C.c_int32), ("demand_lo", C.c_int32),
                ("demand_hi", C.c_int32), ("delay_lo", C.c_int32), ("delay_hi", C.c_int32), ("max_order", C.c_int32),
                ("initial_inventory", _pi), ("customer_demand", _pi), ("shipment_delays", _pi)]


class ScgInfo(C.Structure):
    _fields_ = [("n_envs", C.c_int64), ("family", C.c_int32), ("mode", C.c_int32), ("n_actions", C.c_int32),
                ("n_obs", C.c_int32), ("n_demands", C.c_int32), ("n_leadtimes", C.c_int32),
                ("total_time_steps", C.c_int32), ("elem_size", C.c_int32), ("ring_depth", C.c_int32),
                ("heap_capacity", C.c_int32), ("kernel_variant", C.c_int32), ("info_values", C.c_int32),
                ("state_bytes", C.c_int64)]


# every symbol include/scg.h declares: (name, restype, argtypes)
_vp = C.c_void_p
PROTOTYPES = [
    ("scg_last_error", C.c_char_p, []),
    ("scg_abi_version", C.c_int, []),
    ("scg_sc_create", C.c_int, [C.POINTER(ScgSupplyChainSpec), C.POINTER(ScgOptions), C.POINTER(_vp)]),
    ("scg_bg_create", C.c_int, [C.POINTER(ScgBeerGameSpec), C.POINTER(ScgOptions), C.POINTER(_vp)]),
    ("scg_sc_precompile", C.c_int, [C.POINTER(ScgSupplyChainSpec), C.c_int32, C.c_int32]),
    ("scg_destroy", None, [_vp]),
    ("scg_info", C.c_int, [_vp, C.POINTER(ScgInfo)]),
    ("scg_seed", C.c_int, [_vp, C.c_uint64]),
    ("scg_inject", C.c_int, [_vp, _vp, _vp]),
    ("scg_reset", C.c_int, [_vp, _vp, _vp, _vp]),
    ("scg_step", C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    ("scg_rollout", C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp]),
    ("scg_reset_host", C.c_int, [_vp, _vp, _vp]),
    ("scg_step_host", C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    ("scg_rollout_host", C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp]),
    ("scg_get_state", C.c_int, [_vp, _vp, _vp]),
    ("scg_set_state", C.c_int, [_vp, _vp, _vp]),
    ("scg_return_stats", C.c_int, [_vp, _vp, C.c_int32, _vp]),
    ("scg_get_info", C.c_int, [_vp, _vp, _vp]),
    ("scg_state_layout", C.c_int, [_vp, C.POINTER(C.c_int64 * 8)]),
    ("scg_launch_count", C.c_int64, [_vp]),
    ("scg_status", C.c_int, [_vp, C.c_int32]),
]

_lib = None


def load():
    """Load libscg.so (once).  Raises ScgError if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScgError("CUDA extension %s is missing: run `python -m gym_supplychain_b200.build` "
                       "(there is no CPU fallback)" % LIB_PATH)
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise ScgError("cannot load %s: %s" % (LIB_PATH, e))
    for name, restype, argtypes in PROTOTYPES:
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.scg_abi_version() != 1:
        raise ScgError("libscg.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().scg_last_error()
        raise ScgError("libscg error %d: %s" % (rc, msg.decode() if msg else "?"))
