"""iwae_b200: the importance-weighted ELBO hot path of nbip/IWAE as hand-written sm_100a CUDA kernels
behind the reference's own model API.  Importing the package loads libiwae_b200.so (no CPU fallback)."""
from . import _abi
from ._abi import IwaeError
from .model import IWAE, IWAE2, IWAEDReG, CIWAE, BasicBlock, DeviceValue, Result, Variable
from .optim import Adam
from . import data, dist, posterior

__all__ = ["IWAE", "IWAE2", "IWAEDReG", "CIWAE", "BasicBlock", "Adam", "DeviceValue", "Result", "Variable",
           "IwaeError", "data", "dist", "posterior"]
