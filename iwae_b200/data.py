"""Host-side data helpers mirroring `src/utils.py` of the reference, plus the synthetic stand-in for
MNIST (no dataset is reachable from this image)."""
from __future__ import annotations

import numpy as np


def get_bias(train_mean: np.ndarray) -> np.ndarray:
    """utils.get_bias (src/utils.py:11-23): logit of the clipped mean train pixel.  The reference
    reloads MNIST inside; here the caller passes the mean of whatever train set it has."""
    return -np.log(1.0 / np.clip(np.asarray(train_mean, dtype=np.float64), 0.001, 0.999) - 1.0)


_DEFAULT_BIAS = {}


def default_out_bias(n_x: int = 784) -> np.ndarray:
    """What `Decoder.__init__` of the reference does on every construction (src/iwae1.py:74-75 -> utils.get_bias():
    reload the train set, take the mean image): the output-layer bias of a freshly built model.  MNIST is not reachable
    here, so the train set is the synthetic stand-in (`synthetic_mnist`, 60,000 images); cached per n_x."""
    if n_x not in _DEFAULT_BIAS:
        rs = np.random.RandomState(123)
        mean = np.zeros(n_x)
        for _ in range(6):                                   # 6 x 10,000 rows of the same stream as synthetic_mnist's Ptrain
            mean += (rs.rand(10000, n_x) ** 4).sum(axis=0)
        _DEFAULT_BIAS[n_x] = get_bias(mean / 60000.0)
    return _DEFAULT_BIAS[n_x]


def bernoullisample(x: np.ndarray) -> np.ndarray:
    """utils.bernoullisample (src/utils.py:26-27), on the NumPy global RNG like the reference."""
    return np.random.binomial(1, x, size=x.shape).astype("float32")


def synthetic_mnist(n_train: int = 60000, n_test: int = 10000, seed: int = 123, n_x: int = 784):
    """Intensities rand**4 (mean 0.2), labels uniform over 10 classes; returns
    (Ptrain, ytrain), (Ptest, ytest) with P in [0,1] float64 like `Xtrain/255` (main.py:64-65)."""
    rs = np.random.RandomState(seed)
    Ptrain = rs.rand(n_train, n_x) ** 4
    ytrain = rs.randint(0, 10, n_train)
    Ptest = rs.rand(n_test, n_x) ** 4
    ytest = rs.randint(0, 10, n_test)
    return (Ptrain, ytrain), (Ptest, ytest)


def synthetic_binarised(n: int, seed: int = 123, n_x: int = 784):
    """(P, X, y): n images binarised exactly like utils.bernoullisample; same stream as the oracle's
    `synthetic_data` so tests and bench share inputs."""
    rs = np.random.RandomState(seed)
    P = rs.rand(n, n_x) ** 4
    X = rs.binomial(1, P, size=P.shape).astype("float32")
    y = rs.randint(0, 10, n)
    return P, X, y


def lr_schedule():
    """main.py:44-51: {first epoch: learning rate}, and the total number of epochs (3280)."""
    epochs, d = 0, {}
    for i in range(8):
        d[epochs] = 0.001 * 10 ** (-i / 7)
        epochs += 3 ** i
    return d, epochs


class MyMetric:
    """utils.MyMetric (src/utils.py:30-45): running mean over appended batches of values."""

    def __init__(self):
        self.VALUES, self.N = [], []

    def update_state(self, losses):
        a = np.asarray(losses.numpy() if hasattr(losses, "numpy") else losses, dtype=np.float64)
        a = a.reshape(a.shape[0], -1) if a.ndim else a.reshape(1, 1)
        self.VALUES.append(a)
        self.N.append(a.shape[0])

    def result(self):
        return np.float32(np.sum(np.concatenate(self.VALUES, axis=0)) / np.float32(np.sum(self.N)))

    def reset_states(self):
        self.VALUES, self.N = [], []


class DeviceDataset:
    """The data step in front of `train_step`, resident on the GPU (main.py:117-120 and
    utils.bernoullisample, src/utils.py:26-27): pixel intensities [N, n_x] live in HBM once, every epoch is
    re-binarised by the Philox threshold kernel (`iwae_bernoulli_sample`, keyed on (seed, epoch) so an epoch can
    be regenerated anywhere) and shuffled on the device; `batch(i)` hands `train_step` a device tensor - no
    NumPy binomial draws and no host-to-device copy inside the training loop."""

    def __init__(self, intensities, labels=None, batch_size: int = 20, seed: int = 123, device=None):
        import torch
        from . import _abi  # noqa: F401  (raises ImportError when the CUDA library is not built: no CPU fallback)
        self.device = torch.device(device if device is not None else "cuda")
        self.P = torch.as_tensor(np.asarray(intensities, dtype=np.float32)).to(self.device).contiguous()
        self.y = None if labels is None else torch.as_tensor(np.asarray(labels, dtype=np.float32)).to(self.device)
        self.N, self.batch_size, self.seed = self.P.shape[0], int(batch_size), int(seed)
        self.X = torch.empty_like(self.P)
        self.perm = torch.arange(self.N, device=self.device)
        self._gen = torch.Generator(device=self.device)
        self.epoch = -1

    def __len__(self):
        return self.N // self.batch_size                    # steps_pr_epoch (main.py:72)

    def start_epoch(self, epoch: int, shuffle: bool = True):
        """Re-binarise (main.py:117) and reshuffle (tf.data shuffle(Ntrain), main.py:118-119) for `epoch`."""
        import ctypes as C
        import torch
        from ._abi import lib, check
        with torch.cuda.device(self.device):
            check(lib.iwae_bernoulli_sample(C.c_void_p(self.P.data_ptr()), self.P.numel(), self.seed, int(epoch),
                                            C.c_void_p(self.X.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "iwae_bernoulli_sample")
        if shuffle:
            self._gen.manual_seed(self.seed * 1000003 + int(epoch))
            self.perm = torch.randperm(self.N, device=self.device, generator=self._gen)
        else:
            self.perm = torch.arange(self.N, device=self.device)
        self.epoch = int(epoch)
        return self

    def batch(self, i: int):
        """(x [batch, n_x], y [batch] or None) of step i of the current epoch, as device tensors."""
        idx = self.perm[i * self.batch_size:(i + 1) * self.batch_size]
        return self.X.index_select(0, idx), (None if self.y is None else self.y.index_select(0, idx))
