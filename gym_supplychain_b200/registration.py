"""`make(id, **kwargs)` with the env ids of the reference, so `gym.make('gym_supplychain:<id>')` call
sites only change their import.

Ids registered by the reference (gym_supplychain/__init__.py:3-51) are kept verbatim.  The four ids its
README advertises but never registers (README.md:17-23) are defined here as:
  supplychain-2perstage-v0  = sc-2perstage-v0
  supplychain-linear-v0     = SupplyChainNPerStage(nodes_per_echelon=1, num_products=1)   (1-1-1-1 chain)
  supplychain-ntom-v0       = SupplyChainNPerStage(nodes_per_echelon=3)                   (3-3-3-3, 2 products)
  supplychain-oneonen-v0    = SupplyChainNPerStage(nodes_per_echelon=[1,1,1,3], num_products=1)
`supplychain-v0` needs `nodes_info` in the reference (TypeError without it); here it falls back to the
reference's own demo chain (supplychain_env.py:842-875) when `nodes_info` is omitted.
"""
from . import envs

_REGISTRY = {}


def register(id, entry_point, kwargs=None):
    _REGISTRY[id] = (entry_point, dict(kwargs or {}))


def make(id, **kwargs):
    if ":" in id:                       # 'gym_supplychain:beergame-v0' style
        id = id.split(":", 1)[1]
    if id not in _REGISTRY:
        raise KeyError("No registered env with id: %s" % id)
    cls, defaults = _REGISTRY[id]
    kw = dict(defaults)
    kw.update(kwargs)
    return cls(**kw)


def registered_ids():
    return sorted(_REGISTRY)


# --- ids the reference registers ---
register("beergame-v0", envs.BeerGameEnv)
register("beergame-v2", envs.BeerGameEnv2)
register("supplychain-v0", envs.SupplyChainMainDemo)
register("sc-2perstage-v0", envs.SupplyChain2perStageEnv)
register("sc-2perstage-seasonal-v0", envs.SupplyChain2perStageSeasonalEnv)
register("sc-2perstage-multiproduct-v0", envs.SupplyChainMultiProduct)
register("sc-Nperstage-multiproduct-v0", envs.SupplyChainNPerStage)
register("sc-2perstage-multiproduct-inccosts-v0", envs.SupplyChainMultiProduct_IncreasingCosts)
register("sc-2perstage-multiproduct-v1", envs.SupplyChainMultiProduct_DemConfigByProd)
register("sc-2perstage-multiproduct-inccosts-v1", envs.SupplyChainMultiProduct_DemConfigByProd_IncCosts)
# --- ids the README / north star name ---
register("supplychain-2perstage-v0", envs.SupplyChain2perStageEnv)
register("supplychain-linear-v0", envs.SupplyChainNPerStage, dict(nodes_per_echelon=1, num_products=1))
register("supplychain-ntom-v0", envs.SupplyChainNPerStage, dict(nodes_per_echelon=3))
register("supplychain-oneonen-v0", envs.SupplyChainNPerStage, dict(nodes_per_echelon=[1, 1, 1, 3], num_products=1))
