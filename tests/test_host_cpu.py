"""CPU-side checks of the product's host layer: presets vs the reference's captured nodes_info, spec
sizes, registry ids, the C-ABI library (loads, exports every symbol of include/scg.h, validates specs
and fails loudly without a GPU).  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gym_supplychain_b200 as gsc
from golden_util import load, sc_cases
from gym_supplychain_b200 import _lib, presets
from gym_supplychain_b200.spec import FlatSpec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRESET_OF = {"SupplyChainNPerStage": presets.n_per_stage, "SupplyChain2perStageEnv": presets.two_per_stage,
             "SupplyChain2perStageSeasonalEnv": presets.two_per_stage_seasonal,
             "SupplyChainMultiProduct": presets.multi_product,
             "SupplyChainMultiProduct_IncreasingCosts": presets.multi_product_increasing_costs,
             "SupplyChainMultiProduct_DemConfigByProd": presets.multi_product_dem_by_prod}


def _norm(x):
    if isinstance(x, tuple):
        return [_norm(v) for v in x]
    if isinstance(x, list):
        return [_norm(v) for v in x]
    if isinstance(x, dict):
        return {k: _norm(v) for k, v in x.items()}
    return x


@pytest.mark.parametrize("name", sc_cases())
def test_presets_reproduce_reference_nodes_info(name):
    meta, _ = load("sc", name)
    if meta["cls"] not in PRESET_OF:
        pytest.skip("explicit nodes_info case")
    info, kw = PRESET_OF[meta["cls"]](**meta["ctor_kwargs"])
    assert list(info.keys()) == list(meta["nodes_info"].keys())          # node (= act) order
    assert _norm(info) == meta["nodes_info"]
    ref_kw = meta["env_kwargs"]
    for k, v in ref_kw.items():
        assert _norm(kw[k]) == v, k


def test_main_demo_chain_matches_golden():
    meta, _ = load("sc", "supplychain_v0_main")
    info, kw = presets.main_demo_chain(total_time_steps=120)
    assert _norm(info) == meta["nodes_info"] and list(info) == list(meta["nodes_info"])
    for k, v in meta["ctor_kwargs"].items():
        if k != "nodes_info":
            assert _norm(kw[k]) == v, k


@pytest.mark.parametrize("name", sc_cases())
def test_flat_spec_sizes_match_reference(name):
    meta, _ = load("sc", name)
    f = FlatSpec(meta["nodes_info"], **meta["env_kwargs"])
    assert (f.N, f.P, f.T) == (meta["N"], meta["P"], meta["T"])
    assert f.n_actions == meta["A"] and f.n_obs == meta["O"] and f.n_leadtimes == meta["K"]
    assert f.n_retailers == meta["R"]


def test_flat_spec_rejects_what_the_reference_rejects():
    info, kw = presets.two_per_stage()
    bad = dict(info)
    bad["Supplier1"] = dict(bad["Supplier1"], stock_capacity=1.5)              # float -> ValueError (:186-189)
    with pytest.raises(ValueError):
        FlatSpec(bad, **kw)
    bad["Supplier1"] = dict(info["Supplier1"], supply_cost=[1, 2, 3])          # wrong list length -> assert (:182)
    with pytest.raises(AssertionError):
        FlatSpec(bad, **kw)
    with pytest.raises(AssertionError):                                         # fixed demand (:592)
        FlatSpec(info, **dict(kw, demand_range=(5, 5)))
    missing_cap = {k: {kk: vv for kk, vv in v.items() if kk != "stock_capacity"} for k, v in info.items()}
    with pytest.raises(ValueError):                                             # default float('inf') (:529)
        FlatSpec(missing_cap, **kw)


def test_registry_has_reference_and_readme_ids():
    ids = set(gsc.registered_ids())
    for i in ("beergame-v0", "beergame-v2", "supplychain-v0", "sc-2perstage-v0", "sc-2perstage-seasonal-v0",
              "sc-2perstage-multiproduct-v0", "sc-Nperstage-multiproduct-v0", "sc-2perstage-multiproduct-inccosts-v0",
              "sc-2perstage-multiproduct-v1", "sc-2perstage-multiproduct-inccosts-v1",
              "supplychain-ntom-v0", "supplychain-oneonen-v0", "supplychain-linear-v0", "supplychain-2perstage-v0"):
        assert i in ids
    with pytest.raises(KeyError):
        gsc.make("gym_supplychain:no-such-env-v0")


def test_shapes_of_named_ids():
    for preset, kw, A, O in ((presets.n_per_stage, dict(nodes_per_echelon=1, num_products=1), 4, 14),
                             (presets.two_per_stage, {}, 14, 27),
                             (presets.n_per_stage, dict(nodes_per_echelon=3), 60, 79),
                             (presets.n_per_stage, dict(nodes_per_echelon=[1, 1, 1, 3], num_products=1), 6, 22),
                             (presets.n_per_stage, dict(nodes_per_echelon=10), 620, None),
                             (presets.n_per_stage, dict(nodes_per_echelon=[5, 4, 7, 10]), 246, None),
                             (presets.n_per_stage, dict(nodes_per_echelon=[5, 4, 7, 10], num_products=4), 492, None)):
        info, k = preset(**kw)
        f = FlatSpec(info, **k)
        assert f.n_actions == A
        if O is not None:
            assert f.n_obs == O


# ---------------------------------------------------------------------------------------------
# the C-ABI library
# ---------------------------------------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "scg.h")).read()
    return sorted(set(re.findall(r"SCG_API[^;(]*?\b(scg_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(n for n, _, _ in _lib.PROTOTYPES) == syms      # the ctypes table covers the whole header
    assert lib.scg_abi_version() == 1


def test_create_validates_spec_then_fails_loudly_without_gpu():
    import torch
    lib = _lib.load()
    info, kw = presets.two_per_stage()
    f = FlatSpec(info, **kw)
    opt = _lib.ScgOptions()
    opt.n_envs, opt.device, opt.mode = 4, 0, _lib.MODE_FAST_F32
    h = C.c_void_p()
    s = f.c_struct()
    s.n_products = 99                                            # spec errors are reported before any CUDA call
    assert lib.scg_sc_create(C.byref(s), C.byref(opt), C.byref(h)) == -1
    assert b"bad sizes" in lib.scg_last_error()
    s = f.c_struct()
    rc = lib.scg_sc_create(C.byref(s), C.byref(opt), C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        lib.scg_destroy(h)
    else:
        assert rc == -2 and b"no CPU fallback" in lib.scg_last_error()
        with pytest.raises(gsc.ScgError):
            gsc.make("supplychain-ntom-v0", num_envs=8)
        with pytest.raises(gsc.ScgError):
            gsc.make("beergame-v2")


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import or include it."""
    pkg = os.path.join(ROOT, "gym_supplychain_b200")
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), fn
                assert "oracle/" not in text and "sc_oracle" not in text and "gym_shim" not in text, fn


def test_runtime_instantiation_compiles_without_a_gpu_and_rejects_non_layered_chains(tmp_path, monkeypatch):
    """scg_sc_precompile: NVRTC instantiates the layered kernel template for a custom chain shape on the CPU box (this also
    checks the host-side image of the kernel's constant block against sizeof() inside the generated translation unit);
    a chain that is not full-bipartite between consecutive echelons only has the runtime-topology kernel."""
    import gym_supplychain_b200 as gsc
    from gym_supplychain_b200 import FlatSpec, presets
    monkeypatch.setenv("SCG_JIT_CACHE", str(tmp_path))
    info, kw = presets.n_per_stage(nodes_per_echelon=[2, 1, 2, 3], num_products=3, avg_leadtime=3, max_leadtime=3)
    fs = FlatSpec(info, **kw)
    assert fs.precompile("mixed") is True
    files = sorted(p.name for p in tmp_path.iterdir())
    assert len(files) == 2 and files[0].endswith(".cubin") and files[1].endswith(".name")
    assert fs.precompile("mixed") is True and len(list(tmp_path.iterdir())) == 2          # second call: cache hit
    info2, kw2 = presets.n_per_stage(nodes_per_echelon=2, num_products=1)                   # pre-built (2-2-2-2, P=1)
    assert FlatSpec(info2, **kw2).precompile("fast_f32") is True and len(list(tmp_path.iterdir())) == 2
    broken = {k: dict(v) for k, v in info2.items()}
    first_supplier = next(k for k, v in broken.items() if "supply_capacity" in v)
    broken[first_supplier]["destinations"] = broken[first_supplier]["destinations"][:1]    # no longer full bipartite
    broken[first_supplier]["dest_costs"] = [c[:1] if isinstance(c, (list, tuple)) else c for c in broken[first_supplier]["dest_costs"]]
    broken[first_supplier]["ship_capacity"] = broken[first_supplier]["ship_capacity"][:1]
    assert FlatSpec(broken, **kw2).precompile("fast_f32") is False
    assert fs.precompile("exact_f64") is False                                              # heapq-exact mode: runtime-topology kernel only


def test_large_chains_get_the_node_parallel_kernel(tmp_path, monkeypatch):
    """Chains whose state overflows one thread's registers (SupplyChainNPerStage([5,4,7,10]), the reference's golden-return
    shape test_Nperstage.py:23-53) are instantiated as the node-parallel split kernel (G lanes per env, warp-shuffle sums of
    the shipments per destination); mid-size chains keep the one-thread-per-env kernel.  NVRTC only: no GPU needed."""
    from gym_supplychain_b200 import FlatSpec, presets
    monkeypatch.setenv("SCG_JIT_CACHE", str(tmp_path))
    monkeypatch.delenv("SCG_SPLIT_G", raising=False)
    monkeypatch.delenv("SCG_SPLIT_MINB", raising=False)
    info, kw = presets.n_per_stage(nodes_per_echelon=[5, 4, 7, 10])
    assert FlatSpec(info, **kw).precompile("fast_f32") is True
    names = [p.read_text() for p in tmp_path.iterdir() if p.name.endswith(".name")]
    assert len(names) == 1 and "sc_split_rollout" in names[0]
    info, kw = presets.n_per_stage(nodes_per_echelon=[2, 1, 2, 3], num_products=3, avg_leadtime=3, max_leadtime=3)
    assert FlatSpec(info, **kw).precompile("fast_f32") is True
    names = [p.read_text() for p in tmp_path.iterdir() if p.name.endswith(".name")]
    assert len(names) == 2 and sum("sc_layered_rollout" in x for x in names) == 1


def test_negative_ship_capacity_diagnostic():
    """The tolerance modes cannot follow the reference once a factory edge went negative: the host layer says
    which specs can get there.  None of the reference's presets can; a multi-product factory with small edges can."""
    from gym_supplychain_b200 import FlatSpec, presets
    for info, kw in (presets.n_per_stage(nodes_per_echelon=3), presets.n_per_stage(nodes_per_echelon=[5, 4, 7, 10], num_products=4),
                     presets.multi_product(), presets.multi_product_increasing_costs(num_products=3), presets.two_per_stage(),
                     presets.main_demo_chain()):
        assert FlatSpec(info, **kw).negative_ship_capacity_reachable() == []
    info, kw = presets.n_per_stage(nodes_per_echelon=2, ship_capacity={"suppliers": [1000, 1000], "factories": [100, 100], "wholesalers": [1000, 1000]})
    hits = FlatSpec(info, **kw).negative_ship_capacity_reachable()
    assert hits and all(h[0].startswith("Factory") for h in hits)
    info, kw = presets.n_per_stage(nodes_per_echelon=2, num_products=1, ship_capacity={"suppliers": [1000, 1000], "factories": [100, 100], "wholesalers": [1000, 1000]})
    assert FlatSpec(info, **kw).negative_ship_capacity_reachable() == []          # one product: nothing comes after the clip


@pytest.mark.parametrize("kw,mode", [
    (dict(nodes_per_echelon=[2, 3, 2, 4], num_products=3, avg_leadtime=3, max_leadtime=3), "mixed"),
    (dict(nodes_per_echelon=[5, 4, 7, 10]), "fast_f32"),
    (dict(nodes_per_echelon=[5, 4, 7, 10]), "mixed"),          # node-parallel kernel, fixed-point ring, two action tiles
    (dict(nodes_per_echelon=10), "mixed"),                     # node-parallel kernel, 8 lanes per env
    (dict(nodes_per_echelon=4), "mixed"),
    (dict(nodes_per_echelon=3, num_products=3, stochastic_leadtimes=True, avg_leadtime=2, max_leadtime=4), "mixed"),   # RingFixedStoch
])
def test_instantiated_kernels_keep_the_env_state_in_registers(tmp_path, monkeypatch, kw, mode):
    """A run-time index that survives unrolling silently moves the WHOLE env state to local memory (seen twice while
    writing the kernels: a 10x slowdown with no spill reported).  Guard: the stack frame of the instantiated kernels stays
    what honest register spills need (cuobjdump --dump-resource-usage of the cached cubin; no GPU needed)."""
    import re
    import shutil
    import subprocess
    from gym_supplychain_b200 import FlatSpec, presets
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    monkeypatch.setenv("SCG_JIT_CACHE", str(tmp_path))
    for v in ("SCG_SPLIT_G", "SCG_SPLIT_MINB", "SCG_JIT_BAKE", "SCG_JIT_FORCE", "SCG_JIT_DEFINES"):
        monkeypatch.delenv(v, raising=False)
    info, k = presets.n_per_stage(**kw)
    assert FlatSpec(info, **k).precompile(mode) is True
    cubins = [p for p in tmp_path.iterdir() if p.name.endswith(".cubin")]
    assert len(cubins) == 1
    out = subprocess.run([cuobjdump, "--dump-resource-usage", str(cubins[0])], capture_output=True, text=True).stdout
    m = re.search(r"REG:(\d+) STACK:(\d+)", out)
    assert m, out[-400:]
    regs, stack = int(m.group(1)), int(m.group(2))
    assert stack <= 700, "kernel state demoted to local memory? REG %d STACK %d" % (regs, stack)


def test_bench_config_tables():
    """bench.py's host-side bookkeeping: the strong-scaling shape of configs[3] always covers 8,388,608 envs x 128 periods,
    the algorithmic bytes follow SURVEY 8(d), the mixed CPU rate is the harmonic combination, every config is named."""
    import bench
    for world in (1, 2, 4, 8):
        n_local, T_launch, lps = bench.c4s_shape(world)
        assert n_local * world == 1 << 23 and T_launch * lps == 128 and T_launch >= 16
        assert n_local * T_launch * bench.BYTES_PER_ENV_STEP <= 80e9          # I/O buffers of one launch fit a 180 GB GPU
    assert bench.BYTES_PER_ENV_STEP == 561 and bench.bytes_per_env_step(14, 27) == 169 and bench.bytes_per_env_step(4, 14) == 77
    assert abs(bench.mixed_rate([(100.0, 720.0), (1000.0, 700.0)]) - 1420.0 / (7.2 + 0.7)) < 1e-9
    for cfg in ("c1", "c2", "c3", "c4", "c4s", "c5"):
        c = bench.workload_config(cfg, 1, 1, 1, "mixed")
        assert c["id"] == cfg and "obs_mode" in c and cfg in bench.REF_SAMPLE
    assert set(case for v in bench.REF_SAMPLE.values() for case, _, _ in v) <= set(bench.REF_CASES)
    # both arms print the same `config` object: it is a function of the command line and the world size alone
    args = bench.parse_args(["--gpus", "2"])
    for cfg in ("c1", "c2", "c3", "c4", "c4s", "c5"):
        c = bench.full_config(cfg, args, 2)
        assert c["id"] == cfg and ("l2" in c or cfg == "c1")
    c4 = bench.full_config("c4", args, 2)
    assert c4["envs_total"] == 2 * c4["envs_per_gpu"] and c4["bytes_per_env_step"] == 561 and "all-gather" in c4["collective"]
