"""TEST INFRASTRUCTURE (oracle/): a minimal `tensorflow` look-alike on top of PyTorch.

TensorFlow 2.5 — what the reference's Python half runs on — cannot be installed here (Python 3.12, no network). The
reference's element library `wdf_py/lib/tf_wdf.py`, its `layers.py` and the model / loss definitions inside its training
scripts use a small slice of the TF API (Variable, zeros / ones_like / constant, a dozen math ops, TensorArray,
GradientTape, MeanSquaredError). This package provides exactly that slice, eagerly, backed by torch tensors, so that
`tests/golden/make_golden_py_reference.py` can import and execute the UNMODIFIED reference sources from /root/reference
and record their outputs — and, through torch.autograd behind `GradientTape`, the gradients their op graph defines.
Nothing in the product imports this; it exists to pin the oracle (oracle/torch_wdf.py, oracle/nn.py) and the kernels to
the reference's own Python code.

`set_dtype(torch.float64)` runs the same sources in double precision (every `tf.float32` then means float64): the
gradient oracle. Default float32, as the reference.
"""
import numpy as _np
import torch as _torch

_DTYPE = _torch.float32


def set_dtype(dt):
    global _DTYPE, float32
    _DTYPE = dt
    float32 = dt


float32 = _DTYPE
float64 = _torch.float64
int32 = _torch.int32


def _t(x, dtype=None):
    if isinstance(x, _torch.Tensor):
        return x if dtype is None or x.dtype == dtype or not x.dtype.is_floating_point else x.to(dtype)
    return _torch.as_tensor(_np.asarray(x), dtype=dtype if dtype is not None else (_DTYPE if _np.asarray(x).dtype.kind == "f" else None))


class Variable(_torch.Tensor):
    """tf.Variable: a leaf tensor (requires_grad = trainable) with assign() and an optional constraint."""

    @staticmethod
    def __new__(cls, initial_value=None, name=None, dtype=None, trainable=True, constraint=None, **kw):
        data = _t(initial_value, _DTYPE).detach().clone().to(_DTYPE)
        v = _torch.Tensor._make_subclass(cls, data, bool(trainable))
        v._tf_name, v._tf_trainable, v._tf_constraint = name, bool(trainable), constraint
        return v

    def assign(self, value):
        with _torch.no_grad():
            self.data = _t(value, _DTYPE).detach().clone().to(_DTYPE)
        return self

    def numpy(self):
        return self.detach().cpu().numpy()

    @property
    def trainable(self):
        return getattr(self, "_tf_trainable", False)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):  # results of arithmetic are plain tensors
        with _torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **(kwargs or {}))
        return out


class Module:
    def __init__(self, name=None):
        pass

    @property
    def trainable_variables(self):
        """Trainable Variables reachable from this module, in attribute (then list) order — what tf.Module collects."""
        seen, out = set(), []

        def visit(o):
            if id(o) in seen:
                return
            seen.add(id(o))
            if isinstance(o, Variable):
                if o.trainable:
                    out.append(o)
            elif isinstance(o, Module):
                for k in sorted(vars(o)):
                    visit(vars(o)[k])
            elif isinstance(o, (list, tuple)):
                for e in o:
                    visit(e)

        visit(self)
        return out


class _Logger:
    def setLevel(self, *_):
        pass


def get_logger():
    return _Logger()


def zeros(shape, dtype=None):
    return _torch.zeros(shape if isinstance(shape, (tuple, list)) else (int(shape),), dtype=_DTYPE)


def zeros_like(x):
    return _torch.zeros_like(_t(x, _DTYPE))


def ones_like(x):
    return _torch.ones_like(_t(x, _DTYPE))


def constant(v, dtype=None):
    return _t(v, _DTYPE)


def cast(x, dtype=None):
    x = _t(x)
    return x.to(dtype if dtype is not None else _DTYPE)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def concat(values, axis):
    return _torch.cat([_t(v, _DTYPE) for v in values], dim=axis)


def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*perm) if perm is not None else x.t()


def shape(x):
    return list(_t(x).shape)


def sqrt(x):
    return _torch.sqrt(_t(x, _DTYPE))


def matmul(a, b):
    return _torch.matmul(_t(a, _DTYPE), _t(b, _DTYPE))


def clip_by_value(x, lo, hi):
    return _torch.clamp(_t(x, _DTYPE), lo, hi)


class _Math:
    log = staticmethod(lambda x: _torch.log(_t(x, _DTYPE)))
    abs = staticmethod(lambda x: _torch.abs(_t(x, _DTYPE)))
    square = staticmethod(lambda x: _t(x, _DTYPE) ** 2)
    reciprocal = staticmethod(lambda x: 1.0 / _t(x, _DTYPE))
    reduce_sum = staticmethod(lambda x, axis=None: _torch.sum(_t(x, _DTYPE)) if axis is None else _torch.sum(_t(x, _DTYPE), dim=axis))
    reduce_mean = staticmethod(lambda x, axis=None: _torch.mean(_t(x, _DTYPE)) if axis is None else _torch.mean(_t(x, _DTYPE), dim=axis))
    reduce_min = staticmethod(lambda x: _torch.min(_t(x, _DTYPE)))
    reduce_max = staticmethod(lambda x: _torch.max(_t(x, _DTYPE)))


math = _Math()


class _NN:
    tanh = staticmethod(lambda x: _torch.tanh(x))
    relu = staticmethod(lambda x: _torch.relu(x))


nn = _NN()


class TensorArray:
    def __init__(self, dtype=None, size=0, clear_after_read=True, **kw):
        self._items = [None] * int(size)

    def write(self, i, v):
        self._items[i] = v
        return self

    def stack(self):
        return _torch.stack(self._items)


class GradientTape:
    """Eager autograd: gradient(loss, variables) = torch.autograd.grad over the ops recorded since the variables were created."""

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, target, sources):
        return list(_torch.autograd.grad(target, list(sources), allow_unused=True))


class _Losses:
    class MeanSquaredError:
        def __call__(self, y_true, y_pred):
            return _torch.mean((_t(y_pred, _DTYPE) - _t(y_true, _DTYPE)) ** 2)


class _Init:
    class Orthogonal:
        def __call__(self, shape):
            return _np.zeros(shape, _np.float64)

    class Zeros:
        def __call__(self, shape):
            return _np.zeros(shape, _np.float64)


class _Keras:
    losses = _Losses
    initializers = _Init


keras = _Keras()
