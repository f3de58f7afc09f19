"""Drop-in for the reference's `src/utils.py`: logmeanexp, get_bias, bernoullisample, MyMetric."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from iwae_b200.data import bernoullisample, MyMetric, synthetic_mnist, DeviceDataset  # noqa: E402,F401
from iwae_b200 import data as _data  # noqa: E402


def logmeanexp(log_w, axis):
    """utils.logmeanexp (src/utils.py:6-8) for host arrays / DeviceValues (the model computes its own
    log-mean-exp inside the kernels; this is the free function the reference exposes)."""
    a = np.asarray(log_w.numpy() if hasattr(log_w, "numpy") else log_w)
    m = np.max(a, axis=axis)
    return np.log(np.mean(np.exp(a - np.expand_dims(m, axis)), axis=axis)) + m


def get_bias(train_mean=None):
    """utils.get_bias (src/utils.py:11-23).  MNIST is not reachable here, so the mean train pixel comes
    from the synthetic stand-in unless the caller supplies one."""
    if train_mean is None:
        (Ptrain, _), _ = synthetic_mnist()
        train_mean = Ptrain.mean(axis=0)
    return _data.get_bias(train_mean)
