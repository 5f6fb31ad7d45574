# coding=utf-8
"""Minimal stand-in for tf.keras.Model as the reference's layers use it: weights are created lazily from the first
input (Keras `build`), named exactly like the reference's `add_weight` calls so that a state dict maps 1:1
(kernel, bias, query_kernel, key_kernel, self_kernel, neighbor_kernel, kernel_i, bias_i ...), initialised with
glorot_uniform / zeros, and the layer is invoked as `layer(inputs, cache=..., training=...)`."""
import math
import warnings

import torch

from .. import ops


class Layer(torch.nn.Module):

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.built = False
        self._seed = kwargs.pop("seed", None)
        # Keras weights are trainable by default; here weights only require grad when asked, so that plain forward calls
        # take the fused inference kernels.  Backward passes exist for every convolution (autograd.py).
        self._trainable = bool(kwargs.pop("trainable", False))

    def add_weight(self, name, shape, initializer="glorot_uniform", regularizer=None, device=None):
        shape = [int(s) for s in shape]
        w = torch.empty(shape, dtype=torch.float32, device=device)
        if initializer == "zeros":
            w.zero_()
        elif initializer == "glorot_uniform":
            fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
            limit = math.sqrt(6.0 / (fan_in + fan_out))
            gen = None
            if self._seed is not None:
                gen = torch.Generator(device="cpu")
                gen.manual_seed(self._seed + sum(ord(c) for c in name))
            w.copy_((torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * limit)
        else:
            raise ValueError("unknown initializer {}".format(initializer))
        param = torch.nn.Parameter(w, requires_grad=self._trainable)
        self.__dict__.pop(name, None)       # __init__ pre-declares the slot as None, like the reference's layers
        self.register_parameter(name, param)
        return param

    def build(self, input_shapes, device=None):
        raise NotImplementedError

    def _maybe_build(self, inputs):
        if not self.built:
            x = inputs[0]
            device = x.device if torch.is_tensor(x) and x.is_cuda else ops.default_device()
            self.build([tuple(x.shape)], device=device)
            self.built = True

    def forward(self, inputs, **kwargs):
        self._maybe_build(inputs)
        if kwargs.get("training") and not self._trainable and not getattr(self, "_warned_frozen", False) \
                and any(True for _ in self.parameters()):
            # Keras weights train by default; here they are frozen unless the layer was created with trainable=True
            warnings.warn("{} was called with training=True but its weights were created with trainable=False: they will "
                          "receive no gradient. Create the layer with trainable=True to train it.".format(type(self).__name__),
                          RuntimeWarning, stacklevel=2)
            self._warned_frozen = True
        return self.call(inputs, **kwargs)
