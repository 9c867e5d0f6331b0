"""Minimal `tensorflow` stand-in (numpy) -- see oracle/tf_shim/README.md.  TEST INFRASTRUCTURE."""
import numpy as np

from . import keras  # noqa: F401

float32, float16, uint8, int32, int64 = np.float32, np.float16, np.uint8, np.int32, np.int64
__version__ = "2.4.1-shim"


def function(fn=None, **_kw):                  # utils.py:42,74 decorate at import time
    if fn is None:
        return lambda f: f
    return fn


def constant(value, dtype=None):               # model.py:161-175
    return np.asarray(value, dtype=dtype)


def pad(tensor, paddings, mode="CONSTANT", constant_values=0):      # model.py:203,205
    assert mode == "CONSTANT"
    return np.pad(np.asarray(tensor), [tuple(int(v) for v in p) for p in np.asarray(paddings)],
                  mode="constant", constant_values=constant_values)


def reshape(tensor, shape):                    # model.py:125,127,476-492
    return np.reshape(np.asarray(tensor), tuple(int(s) for s in shape))


def reduce_mean(tensor, axis=None, keepdims=False):                 # model.py:126
    return np.mean(np.asarray(tensor), axis=axis, keepdims=keepdims)


def shape(tensor):
    return np.asarray(np.shape(tensor))


def cast(tensor, dtype):
    return np.asarray(tensor).astype(dtype)
