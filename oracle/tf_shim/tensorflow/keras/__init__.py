"""Minimal `tensorflow.keras` stand-in (numpy float64) -- see oracle/tf_shim/README.md.

TEST INFRASTRUCTURE.  Implements exactly the Keras surface the reference's model.py touches
(model.py:1-492): Layer / Model / Sequential, Conv3D (valid|same, strides, groups, bias,
activation), BatchNormalization (inference), Dense, Dropout (inference), Activation, ReLU, Add,
Softmax, GlobalAveragePooling3D, regularizers.L2, plus the object-graph naming TensorFlow uses
for checkpoint keys (attribute paths, list indices, `layer_with_weights-i` for Sequential).
Arithmetic is numpy float64 over sliding windows / einsum.
"""
import inspect
import types

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

_F = np.float64


# ------------------------------------------------------------------------------ activations
def _relu(x):
    return np.maximum(x, 0.0)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _swish(x):                                   # tf.keras.activations.swish = x * sigmoid(x)
    return x * _sigmoid(x)


def _softmax(x, axis=-1):
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return e / np.sum(e, axis=axis, keepdims=True)


_ACT = {None: lambda x: x, "linear": lambda x: x, "relu": _relu, "sigmoid": _sigmoid,
        "swish": _swish, "softmax": _softmax}


# ------------------------------------------------------------------------------ base classes
class Layer:
    _uid = {}

    def __init__(self, name=None, dtype=None, trainable=True, **kwargs):
        if name is None:
            base = type(self).__name__.lower()
            n = Layer._uid.get(base, 0)
            Layer._uid[base] = n + 1
            name = base if n == 0 else "%s_%d" % (base, n)
        self.name = name
        self.built = False
        self._weights = {}                       # variable leaf name -> ndarray (creation order)

    # Keras creates variables at first call from the input shape
    def build(self, input_shape):
        pass

    def add_weight(self, name, shape, init):
        self._weights[name] = init(tuple(int(s) for s in shape))
        return self._weights[name]

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            shp = [np.shape(t) for t in inputs] if isinstance(inputs, (list, tuple)) else np.shape(inputs)
            self.build(shp)
            self.built = True
        params = inspect.signature(self.call).parameters
        if "training" not in params:
            kwargs.pop("training", None)
        return self.call(inputs, *args, **kwargs)

    def call(self, inputs, **kwargs):
        return inputs


class Model(Layer):
    pass


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self.layers = []
        for l in layers or []:
            self.add(l)

    def add(self, layer):
        self.layers.append(layer)

    def call(self, inputs, training=False):
        x = inputs
        for l in self.layers:
            x = l(x, training=training)
        return x


def Input(shape=None, **kwargs):
    raise NotImplementedError("K.Input / Model.summary are not part of the numeric path")


# ------------------------------------------------------------------------------ initialisers
_rng = np.random.default_rng(20260417)


def _glorot_uniform(shape, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return _rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _tf_same_pad(size, k, s):
    """TensorFlow padding='same': (before, after)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


# ------------------------------------------------------------------------------ layers
class _Layers(types.ModuleType):
    pass


class Conv3D(Layer):
    """tf.keras.layers.Conv3D, channels_last; kernel DHWIO `[kd,kh,kw,Cin/groups,Cout]`.
    Call sites: model.py:80,95,178,187,246,259,278,284,292,360."""

    def __init__(self, filters, kernel_size, strides=(1, 1, 1), padding="valid",
                 data_format=None, dilation_rate=(1, 1, 1), groups=1, activation=None,
                 use_bias=True, kernel_regularizer=None, name=None, **kwargs):
        super().__init__(name=name)
        tup = lambda v: (v, v, v) if isinstance(v, int) else tuple(v)
        self.filters, self.kernel_size, self.strides = int(filters), tup(kernel_size), tup(strides)
        self.padding, self.groups, self.use_bias = padding.lower(), int(groups), use_bias
        assert data_format in (None, "channels_last")
        assert tup(dilation_rate) == (1, 1, 1)
        self.activation = _ACT[activation]
        self.kernel_regularizer = kernel_regularizer

    def build(self, input_shape):
        cin = int(input_shape[-1])
        assert cin % self.groups == 0 and self.filters % self.groups == 0
        kshape = self.kernel_size + (cin // self.groups, self.filters)
        rf = int(np.prod(self.kernel_size))
        self.add_weight("kernel", kshape,
                        lambda s: _glorot_uniform(s, rf * kshape[-2], rf * kshape[-1]))
        if self.use_bias:
            self.add_weight("bias", (self.filters,), lambda s: np.zeros(s, np.float32))

    def call(self, inputs):
        x = np.asarray(inputs, _F)
        w = np.asarray(self._weights["kernel"], _F)
        kd, kh, kw = self.kernel_size
        sd, sh, sw = self.strides
        if self.padding == "same":
            pads = [(0, 0)] + [_tf_same_pad(x.shape[1 + i], self.kernel_size[i], self.strides[i])
                               for i in range(3)] + [(0, 0)]
            x = np.pad(x, pads)
        else:
            assert self.padding == "valid"
        # windows: [N, To, Ho, Wo, C, kd, kh, kw]
        win = sliding_window_view(x, (kd, kh, kw), axis=(1, 2, 3))[:, ::sd, ::sh, ::sw]
        cin = x.shape[-1]
        if self.groups == 1:
            y = np.einsum("nthwcdef,defco->nthwo", win, w, optimize=True)
        elif self.groups == cin and self.filters == cin:       # channelwise: kernel [kd,kh,kw,1,C]
            y = np.einsum("nthwcdef,defc->nthwc", win, w[:, :, :, 0, :], optimize=True)
        else:
            raise NotImplementedError("grouped conv other than channelwise")
        if self.use_bias:
            y = y + np.asarray(self._weights["bias"], _F)
        return self.activation(y)


class BatchNormalization(Layer):
    """Inference mode only: (x - moving_mean) * gamma / sqrt(moving_variance + eps) + beta.
    Call sites: model.py:89,196,254,268,300,368."""

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, name=None, **kwargs):
        super().__init__(name=name)
        assert axis == -1
        self.momentum, self.epsilon = momentum, epsilon

    def build(self, input_shape):
        c = int(input_shape[-1])
        self.add_weight("gamma", (c,), lambda s: np.ones(s, np.float32))
        self.add_weight("beta", (c,), lambda s: np.zeros(s, np.float32))
        self.add_weight("moving_mean", (c,), lambda s: np.zeros(s, np.float32))
        self.add_weight("moving_variance", (c,), lambda s: np.ones(s, np.float32))

    def call(self, inputs, training=False):
        if training:
            raise NotImplementedError("shim BatchNormalization is inference-only")
        W = {k: np.asarray(v, _F) for k, v in self._weights.items()}
        x = np.asarray(inputs, _F)
        return (x - W["moving_mean"]) * (W["gamma"] / np.sqrt(W["moving_variance"] + self.epsilon)) + W["beta"]


class Dense(Layer):                              # model.py:104
    def __init__(self, units, activation=None, use_bias=True, kernel_regularizer=None, name=None, **kw):
        super().__init__(name=name)
        self.units, self.use_bias, self.activation = int(units), use_bias, _ACT[activation]

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.add_weight("kernel", (cin, self.units), lambda s: _glorot_uniform(s, cin, self.units))
        if self.use_bias:
            self.add_weight("bias", (self.units,), lambda s: np.zeros(s, np.float32))

    def call(self, inputs):
        y = np.asarray(inputs, _F) @ np.asarray(self._weights["kernel"], _F)
        if self.use_bias:
            y = y + np.asarray(self._weights["bias"], _F)
        return self.activation(y)


class Dropout(Layer):                            # model.py:103; identity unless training
    def __init__(self, rate, name=None, **kw):
        super().__init__(name=name)
        self.rate = rate

    def call(self, inputs, training=False):
        if training:
            raise NotImplementedError("shim Dropout is inference-only")
        return inputs


class Activation(Layer):                         # model.py:93,200,272,382
    def __init__(self, activation, name=None, **kw):
        super().__init__(name=name)
        self.fn = _ACT[activation]

    def call(self, inputs):
        return self.fn(np.asarray(inputs, _F))


class ReLU(Layer):                               # model.py:258
    def call(self, inputs):
        return _relu(np.asarray(inputs, _F))


class Add(Layer):                                # model.py:381
    def call(self, inputs):
        out = np.asarray(inputs[0], _F)
        for t in inputs[1:]:
            out = out + np.asarray(t, _F)
        return out


class Softmax(Layer):                            # model.py:111
    def __init__(self, axis=-1, dtype=None, name=None, **kw):
        super().__init__(name=name)
        self.axis = axis

    def call(self, inputs):
        return _softmax(np.asarray(inputs, _F), axis=self.axis)


class GlobalAveragePooling3D(Layer):             # model.py:471; channels_last -> [N, C]
    def call(self, inputs):
        return np.mean(np.asarray(inputs, _F), axis=(1, 2, 3))


layers = _Layers("tensorflow.keras.layers")
for _c in (Layer, Conv3D, BatchNormalization, Dense, Dropout, Activation, ReLU, Add, Softmax,
           GlobalAveragePooling3D):
    setattr(layers, _c.__name__, _c)


class _L2:
    def __init__(self, l2=0.01):
        self.l2 = l2


regularizers = types.ModuleType("tensorflow.keras.regularizers")
regularizers.L2 = _L2
regularizers.l2 = _L2
regularizers.Regularizer = _L2

import sys as _sys  # noqa: E402
_sys.modules.setdefault("tensorflow.keras.layers", layers)
_sys.modules.setdefault("tensorflow.keras.regularizers", regularizers)


# ------------------------------------------------------------------------------ checkpoint names
def _has_weights(obj, seen=None):
    return any(True for _ in iter_variables(obj, ""))


def iter_variables(obj, prefix=""):
    """Yields (checkpoint attribute path, owner layer, leaf) in TensorFlow's object-graph naming:
    attribute names joined by '/', list elements by index, Sequential children as
    `layer_with_weights-<i>` (i counts only layers that own variables).  Deterministic order."""
    if isinstance(obj, Sequential):
        i = 0
        for l in obj.layers:
            sub = list(iter_variables(l, "%slayer_with_weights-%d/" % (prefix, i)))
            if sub:
                i += 1
                yield from sub
        return
    if isinstance(obj, Layer):
        for leaf in obj._weights:
            yield prefix + leaf, obj, leaf
        for attr, val in vars(obj).items():
            if attr.startswith("__") or attr in ("_weights",):
                continue
            if isinstance(val, Layer):
                yield from iter_variables(val, "%s%s/" % (prefix, attr))
            elif isinstance(val, (list, tuple)) and val and all(isinstance(v, Layer) for v in val):
                for j, v in enumerate(val):
                    yield from iter_variables(v, "%s%s/%d/" % (prefix, attr, j))


def named_variables(model):
    return {name: owner._weights[leaf] for name, owner, leaf in iter_variables(model)}


def assign_variables(model, values):
    missing = []
    for name, owner, leaf in iter_variables(model):
        if name in values:
            v = np.asarray(values[name], np.float32)
            assert v.shape == owner._weights[leaf].shape, (name, v.shape, owner._weights[leaf].shape)
            owner._weights[leaf] = v
        else:
            missing.append(name)
    return missing
