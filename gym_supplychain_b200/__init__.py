"""gym_supplychain_b200 -- B200-native batched simulator for the gym_supplychain environments.

Drop-in for ONE hot path of caburu/gym-supplychain: reset()/step() of its envs, batched over millions of
independent envs per GPU in hand-written sm_100a CUDA kernels behind the C ABI of include/scg.h.

    import gym_supplychain_b200 as gsc
    env = gsc.make('supplychain-ntom-v0', num_envs=1 << 20)      # vector env, torch CUDA tensors
    obs = env.reset(); obs, reward, done, info = env.step(actions)
    obs, reward, done = env.rollout(128, actions)                # fused multi-step kernel
    env1 = gsc.make('gym_supplychain:sc-2perstage-v0')           # single env, numpy, old-gym API
"""
from . import presets  # noqa: F401
from ._lib import LIB_PATH, ScgError  # noqa: F401
from .envs import (BeerGameEnv, BeerGameEnv2, SupplyChain2perStageEnv, SupplyChain2perStageSeasonalEnv,  # noqa: F401
                   SupplyChainEnv, SupplyChainMainDemo, SupplyChainMultiProduct,
                   SupplyChainMultiProduct_DemConfigByProd, SupplyChainMultiProduct_DemConfigByProd_IncCosts,
                   SupplyChainMultiProduct_IncreasingCosts, SupplyChainNPerStage)
from .registration import make, register, registered_ids  # noqa: F401
from .spec import FlatSpec  # noqa: F401
from .vector import GpuVecEnv  # noqa: F401

__version__ = "0.1.0"
