"""Minimal stand-in for legacy OpenAI gym (<=0.21) -- TEST INFRASTRUCTURE ONLY.

The image has no `gym` wheel and no network.  This shim provides
exactly the surface the reference package imports (`gym.Env`, `gym.spaces`,
`gym.envs.registration`, `gym.utils.seeding`, `gym.error`) so that the
UNMODIFIED reference under /root/reference can be imported in the build
container to (a) validate the oracle restatement and (b) generate the golden
vectors committed under tests/golden/.  Nothing in the product package
(`gym_supplychain_b200`) imports it.

The only non-trivial behaviour is legacy seeding/sampling (sha512-hashed seed ->
RandomState; Box.sample = uniform(low, high).astype(float32)), which the
reference's golden episode returns were produced through.
"""
from . import error, spaces, utils  # noqa: F401
from .core import Env  # noqa: F401
from .envs.registration import make, register  # noqa: F401

__version__ = "0.21.0-shim"
