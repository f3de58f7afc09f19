"""keras.optimizers.Adam(lr, epsilon=1e-4) as the reference drives it (main.py:93-94,130-133;
src/iwae1.py:160): `.learning_rate.numpy()`, `.learning_rate.assign(v)`, `.apply_gradients(zip(g, vars))`.
The update rule is Keras' (epsilon added to the UNcorrected sqrt(v)), executed by the
vectorised kernel behind `iwae_adam_step`."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._abi import check, lib


class _LearningRate:
    def __init__(self, v: float):
        self.value = float(v)

    def numpy(self):
        return np.float32(self.value)

    def assign(self, v):
        self.value = float(v)
        return self

    def __float__(self):
        return self.value


class Adam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **kwargs):
        self.learning_rate = _LearningRate(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.iterations = 0
        self._m = self._v = None
        self._owner_ptr = None

    def _state_for(self, model):
        """Moment buffers laid out like the model's flat parameter buffer."""
        if self._m is None or self._owner_ptr != model.params.data_ptr():
            self._m = torch.zeros_like(model.params)
            self._v = torch.zeros_like(model.params)
            self._owner_ptr = model.params.data_ptr()
            self._base = model.params
        return self._m, self._v

    def apply_gradients(self, grads_and_vars):
        """Keras semantics: one `iterations` increment per call, every (grad, var) pair updated with the
        same step count.  `var` is a `model.Variable` (a view into the flat buffer); `grad` a tensor or
        DeviceValue of the same shape."""
        pairs = list(grads_and_vars)
        if not pairs:
            return
        self.iterations += 1
        stream = torch.cuda.current_stream().cuda_stream
        for g, var in pairs:
            base = var.value._base if var.value._base is not None else var.value
            if self._m is None or self._owner_ptr != base.data_ptr():
                self._m = torch.zeros_like(base)
                self._v = torch.zeros_like(base)
                self._owner_ptr = base.data_ptr()
            g = getattr(g, "tensor", g)
            g = g.to(device=var.value.device, dtype=torch.float32).contiguous()
            n = var.value.numel()
            m = self._m[var.offset:var.offset + n]
            v = self._v[var.offset:var.offset + n]
            check(lib.iwae_adam_step(C.c_void_p(var.value.data_ptr()), C.c_void_p(g.data_ptr()),
                                     C.c_void_p(m.data_ptr()), C.c_void_p(v.data_ptr()), n, self.learning_rate.value,
                                     self.beta_1, self.beta_2, self.epsilon, self.iterations, stream),
                  "iwae_adam_step")
