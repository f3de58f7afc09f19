"""Drop-in for the reference's `src/iwae1.py` (`sys.path.insert(0, './src'); import iwae1`, main.py:10-13):
same class names, constructor and call signatures; the arithmetic runs in libiwae_b200.so."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from iwae_b200.model import IWAE, BasicBlock, IWAEDReG, CIWAE  # noqa: E402,F401
