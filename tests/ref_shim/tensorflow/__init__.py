"""Minimal `tensorflow` stand-in on torch-CPU float64 (TEST INFRASTRUCTURE, see ../README.md).

Only the ~45 symbols the reference's hot-path modules touch (SURVEY.md §2 third-party table). Tensors are a
torch.Tensor subclass so that eager arithmetic, indexing and numpy interop behave like tf.EagerTensor; torch autograd
stands in for tf.GradientTape."""
import types

import numpy as np
import torch

float64 = torch.float64
float32 = torch.float32
int16 = torch.int16
int32 = torch.int32
int64 = torch.int64


class Tensor(torch.Tensor):
    """Eager tensor: numpy() works on tensors that carry gradients, numpy operands are converted like TF does."""

    @classmethod
    def __torch_function__(cls, func, types_, args=(), kwargs=None):
        kwargs = kwargs or {}
        args = tuple(_lift(a) for a in args)
        kwargs = {k: _lift(v) for k, v in kwargs.items()}
        return super().__torch_function__(func, types_, args, kwargs)

    def numpy(self):
        return torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor))

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def mean(self, *args, **kwargs):
        if "axis" in kwargs or "out" in kwargs:      # numpy's np.mean(tensor, axis=...) protocol → ndarray, like TF
            return np.mean(self.numpy(), axis=kwargs.get("axis"))
        return super().mean(*args, **kwargs)

    def __array_wrap__(self, array, context=None, return_scalar=False):
        return array                                 # np.linalg.* on a tensor gives an ndarray, like TF

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """ndarray ∘ tensor arithmetic is done by torch (keeps gradients, like TF); any other numpy ufunc works on the values."""
        if method == "__call__" and ufunc in _ARITH_UFUNCS and not kwargs:
            a, b = (i if isinstance(i, torch.Tensor) else torch.as_tensor(np.asarray(i)) for i in inputs)
            return _ARITH_UFUNCS[ufunc](a, b)
        inputs = tuple(i.numpy() if isinstance(i, Tensor) else i for i in inputs)
        return getattr(ufunc, method)(*inputs, **kwargs)


_ARITH_UFUNCS = {np.add: torch.add, np.subtract: torch.sub, np.multiply: torch.mul, np.true_divide: torch.true_divide,
                 np.power: torch.pow, np.matmul: torch.matmul, np.less: torch.lt, np.greater: torch.gt,
                 np.less_equal: torch.le, np.greater_equal: torch.ge}


def _tensor_op(fn, reflected=False):
    def op(self, other):
        other = _lift(other)
        return fn(other, self) if reflected else fn(self, other)
    return op


def _as_t(x, like):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=like.dtype)


for _name, _fn in (("add", torch.add), ("sub", torch.sub), ("mul", torch.mul), ("truediv", torch.true_divide),
                   ("pow", torch.pow), ("matmul", torch.matmul)):
    setattr(Tensor, f"__{_name}__", _tensor_op(_fn))
    setattr(Tensor, f"__r{_name}__", _tensor_op(lambda a, b, _f=_fn: _f(_as_t(a, b), b), reflected=True))
    setattr(Tensor, f"__i{_name}__", _tensor_op(_fn))      # tf tensors are immutable: `a += b` rebinds, never writes in place


def _lift(a):
    """numpy arrays, Variables and GPflow Parameters inside torch calls become tensors."""
    if isinstance(a, np.ndarray):
        return torch.as_tensor(a)
    if hasattr(a, "_as_tensor"):
        return a._as_tensor()
    if isinstance(a, (list, tuple)) and any(isinstance(x, np.ndarray) or hasattr(x, "_as_tensor") for x in a):
        return type(a)(_lift(x) for x in a)
    return a


def convert_to_tensor(x, dtype=None):
    if hasattr(x, "_as_tensor"):
        x = x._as_tensor()
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, (list, tuple)) and len(x) and any(isinstance(e, torch.Tensor) or hasattr(e, "_as_tensor") for e in x):
        t = torch.stack([convert_to_tensor(e) for e in x])
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t if isinstance(t, Tensor) else t.as_subclass(Tensor)


_c = convert_to_tensor


class Variable:
    """tf.Variable: a leaf with gradient; reads as a tensor in every shim op."""

    def __init__(self, initial_value, trainable=True, dtype=None, name=None):
        v = _c(initial_value, dtype).detach().clone().as_subclass(torch.Tensor)
        self._leaf = v.requires_grad_(bool(v.is_floating_point()))
        self.trainable = trainable
        self.name = name

    def _as_tensor(self):
        return self._leaf.as_subclass(Tensor)

    def numpy(self):
        return self._leaf.detach().numpy().copy()

    @property
    def shape(self):
        return tuple(self._leaf.shape)

    @property
    def dtype(self):
        return self._leaf.dtype

    def assign(self, value):
        with torch.no_grad():
            self._leaf.copy_(_c(value).detach())
        return self

    def assign_sub(self, delta):
        with torch.no_grad():
            self._leaf.sub_(_c(delta).detach())
        return self

    def assign_add(self, delta):
        with torch.no_grad():
            self._leaf.add_(_c(delta).detach())
        return self


def _binop(name):
    def f(self, other):
        return getattr(self._as_tensor(), name)(other)
    return f


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__", "__pow__",
           "__rpow__", "__matmul__", "__getitem__", "__lt__", "__gt__", "__le__", "__ge__"):
    setattr(Variable, _n, _binop(_n))
Variable.__neg__ = lambda self: -self._as_tensor()


def constant(value, dtype=None):
    return _c(value, dtype)


def cast(x, dtype):
    return _c(x).to(dtype)


def shape(x):
    return tuple(int(s) for s in _c(x).shape)


def reshape(x, shp):
    return _c(x).reshape(tuple(int(s) for s in shp))


def transpose(x, perm=None):
    x = _c(x)
    if perm is None:
        perm = tuple(reversed(range(x.dim())))
    return x.permute(*perm)


def tile(x, multiples):
    return _c(x).repeat(*[int(m) for m in multiples])


def expand_dims(x, axis):
    return _c(x).unsqueeze(axis)


def squeeze(x, axis=None):
    return _c(x).squeeze() if axis is None else _c(x).squeeze(axis)


def stack(xs, axis=0):
    if isinstance(xs, torch.Tensor):
        return _c(xs)
    return torch.stack([_c(x) for x in xs], dim=axis).as_subclass(Tensor)


def concat(xs, axis):
    return torch.cat([_c(x) for x in xs], dim=axis).as_subclass(Tensor)


def zeros(shp, dtype=float64):
    return torch.zeros(tuple(int(s) for s in shp), dtype=dtype).as_subclass(Tensor)


def ones(shp, dtype=float64):
    return torch.ones(tuple(int(s) for s in shp), dtype=dtype).as_subclass(Tensor)


def zeros_like(x):
    return torch.zeros_like(_c(x))


def eye(n, dtype=float64):
    return torch.eye(int(n), dtype=dtype).as_subclass(Tensor)


def fill(dims, value):
    return torch.full(tuple(int(s) for s in dims), float(value), dtype=float64).as_subclass(Tensor)


def identity(x):
    return _c(x)


def _reduce(fn):
    def f(x, axis=None, keepdims=False):
        x = _c(x)
        if axis is None:
            return fn(x)
        return fn(x, dim=axis, keepdim=keepdims)
    return f


reduce_sum = _reduce(torch.sum)
reduce_mean = _reduce(torch.mean)
reduce_prod = _reduce(torch.prod)


def reduce_max(x, axis=None):
    x = _c(x)
    return x.max() if axis is None else x.max(dim=axis).values


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _c(a), _c(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def tensordot(a, b, axes):
    a, b = _c(a), _c(b)
    return torch.tensordot(a, b, dims=([int(i) for i in axes[0]], [int(i) for i in axes[1]]))


def add(a, b):
    return _c(a) + _c(b)


def square(x):
    return _c(x) ** 2


def sqrt(x):
    return torch.sqrt(_c(x))


def exp(x):
    return torch.exp(_c(x))


def maximum(a, b):
    return torch.maximum(_c(a), _c(b, _c(a).dtype))


def where(cond, a, b):
    dtype = next((t.dtype for t in (a, b) if isinstance(t, torch.Tensor)), float64)
    return torch.where(_c(cond), _c(a, dtype), _c(b, dtype))


def cond(pred, true_fn, false_fn):
    return true_fn() if bool(_c(pred)) else false_fn()


def map_fn(fn, elems, dtype=None, **_):
    outs = [fn(e) for e in _c(elems)]
    if isinstance(outs[0], (tuple, list)):
        return tuple(stack([o[i] for o in outs]) for i in range(len(outs[0])))
    return stack(outs)


def function(fn=None, **_):
    """@tf.function: eager execution here (the graph only caches the same operations)."""
    if fn is None:
        return lambda f: f
    return fn


def _cholesky(a):
    return torch.linalg.cholesky(_c(a))


def _triangular_solve(matrix, rhs, lower=True, adjoint=False):
    m = _c(matrix)
    if adjoint:
        m, lower = m.transpose(-1, -2), not lower
    return torch.linalg.solve_triangular(m, _c(rhs), upper=not lower)


def _cholesky_solve(chol, rhs):
    return torch.cholesky_solve(_c(rhs), _c(chol), upper=False)


def _diag_part(x):
    return torch.diagonal(_c(x), dim1=-2, dim2=-1)


def _diag(x):
    return torch.diag_embed(_c(x))


def _adjoint(x):
    return _c(x).transpose(-1, -2)


linalg = types.SimpleNamespace(cholesky=_cholesky, triangular_solve=_triangular_solve, cholesky_solve=_cholesky_solve,
                               diag_part=_diag_part, diag=_diag, matmul=matmul, adjoint=_adjoint, eye=eye)

math = types.SimpleNamespace(log=lambda x: torch.log(_c(x)), exp=exp, sqrt=sqrt, square=square,
                             reduce_mean=reduce_mean, reduce_sum=reduce_sum, reduce_max=reduce_max,
                             reduce_std=lambda x: _c(x).std(unbiased=False), reduce_min=lambda x: _c(x).min(),
                             is_nan=lambda x: torch.isnan(_c(x)), erf=lambda x: torch.erf(_c(x)),
                             softplus=lambda x: torch.nn.functional.softplus(_c(x), threshold=700.0))


class _Random:
    """tf.random.normal draws come from `source` when a test installs one (explicit draws shared with the oracle);
    otherwise from a seeded torch generator. TF's own stateful Philox stream is not reproducible without TF."""

    def __init__(self):
        self.source = None
        self._gen = torch.Generator().manual_seed(0)

    def set_seed(self, seed):
        self._gen = torch.Generator().manual_seed(int(seed))

    def normal(self, shp, dtype=float64, **_):
        shp = tuple(int(s) for s in shp)
        if self.source is not None:
            z = _c(self.source(shp), dtype)
            assert tuple(z.shape) == shp, (tuple(z.shape), shp)
            return z
        return torch.randn(shp, dtype=dtype, generator=self._gen).as_subclass(Tensor)


random = _Random()


class GradientTape:
    """with tf.GradientTape() as tape: ...; tape.gradient(target, sources) → torch.autograd.grad (None for unused)."""

    def __init__(self, persistent=False, watch_accessed_variables=True):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, variables):
        pass

    def gradient(self, target, sources, output_gradients=None):
        single = not isinstance(sources, (list, tuple))
        srcs = [sources] if single else list(sources)
        leaves = [s._leaf if isinstance(s, Variable) else s for s in srcs]
        grads = torch.autograd.grad(_c(target), leaves, grad_outputs=output_gradients, allow_unused=True, retain_graph=True)
        grads = [None if g is None else g.detach().as_subclass(Tensor) for g in grads]
        return grads[0] if single else grads


class _Adam:
    """tf.optimizers.Adam (Keras OptimizerV2, amsgrad=False): lr_t = lr·sqrt(1−β2^t)/(1−β1^t);
    m ← β1 m + (1−β1) g; v ← β2 v + (1−β2) g²; var ← var − lr_t·m/(sqrt(v) + ε)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self._slots = {}

    def apply_gradients(self, grads_and_vars):
        self.iterations += 1
        t = self.iterations
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        for g, var in grads_and_vars:
            if g is None:
                continue
            g = _c(g).detach().as_subclass(torch.Tensor)
            m, v = self._slots.setdefault(id(var), (torch.zeros_like(g), torch.zeros_like(g)))
            m.mul_(self.b1).add_((1.0 - self.b1) * g)
            v.mul_(self.b2).add_((1.0 - self.b2) * g * g)
            var.assign_sub(lr_t * m / (torch.sqrt(v) + self.eps))


optimizers = types.SimpleNamespace(Adam=_Adam)
keras = types.SimpleNamespace(optimizers=optimizers)


class Module:
    """tf.Module: named container whose variables are found by walking attributes (sorted by name), lists, tuples
    and dicts recursively."""

    def __init__(self, name=None):
        self._name = name if name is not None else type(self).__name__.lower()

    @property
    def name(self):
        return self._name

    def _walk(self, seen):
        for key in sorted(vars(self)):
            yield from _walk_value(vars(self)[key], seen)

    @property
    def variables(self):
        return tuple(v for v in self._walk(set()) if isinstance(v, Variable))

    @property
    def trainable_variables(self):
        return tuple(v for v in self.variables if v.trainable)


def _walk_value(v, seen):
    if id(v) in seen:
        return
    if isinstance(v, Variable):
        seen.add(id(v))
        yield v
    elif isinstance(v, Module):
        seen.add(id(v))
        yield v
        yield from v._walk(seen)
    elif isinstance(v, (list, tuple)):
        for e in v:
            yield from _walk_value(e, seen)
    elif isinstance(v, dict):
        for k in sorted(v):
            yield from _walk_value(v[k], seen)
