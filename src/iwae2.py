"""Drop-in for the reference's `src/iwae2.py`: `iwae2.IWAE(n_hidden=[h1,h2], n_latent=[d1,d2])`."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from iwae_b200.model import IWAE2 as IWAE, BasicBlock  # noqa: E402,F401
