"""minijax -- the sliver of the JAX API the reference scripts use, on NumPy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  JAX is not installed in this image and
the reference (/root/reference) cannot be imported without it.  This module provides
``jax.numpy`` / ``jit`` / ``vmap`` / ``jacfwd`` / ``jacrev`` / ``grad`` / ``hessian`` /
``jax.config`` with the semantics the reference relies on, so that the reference's OWN source
text can be executed (oracle/refexec/loader.py) and golden vectors minted from it
(tests/golden/make_golden_ref.py).  Nothing here restates the reference's algorithm: it only
supplies mechanism (array ops + automatic differentiation).

Differentiation is exact forward-mode AD with dual numbers: an ``Arr`` carries a value array
``val`` (shape s) and optionally a tangent array ``tan`` (shape s + (n,)), one slice per seeded
input direction.  ``jacfwd`` seeds the identity; ``jacrev`` / ``grad`` return the same Jacobian
(mode does not change the value).  ``hessian`` differentiates the forward-mode gradient once more
by the complex-step method (values may be complex; every op below is analytic, with
``abs`` / ``norm`` extended analytically), which is exact to rounding for step 1e-30.
JAX conventions that matter to the reference and are kept: d|x|/dx = sign(x) (0 at 0),
d||x||/dx = x/||x||, C-order reshapes unless order='F', ``vmap`` maps axis 0 of every argument.
"""
import numpy as np

newaxis = None
inf = np.inf
pi = np.pi
float64 = np.float64


# ------------------------------------------------------------------------------------------
# dual-number array
# ------------------------------------------------------------------------------------------
def _lift(x):
    """-> (val ndarray, tan ndarray | None)"""
    if isinstance(x, Arr):
        return x.val, x.tan
    return np.asarray(x), None


def _wrap(val, tan=None):
    return Arr(val, tan)


def _bt(t, vshape, rshape):
    """broadcast a tangent array of a value of shape ``vshape`` to result shape ``rshape``"""
    n = t.shape[-1]
    t = t.reshape((1,) * (len(rshape) - len(vshape)) + tuple(vshape) + (n,))
    return np.broadcast_to(t, tuple(rshape) + (n,))


def _bv(v, rshape):
    return np.broadcast_to(v, rshape)[..., None]


def _tan_index(idx):
    if not isinstance(idx, tuple):
        idx = (idx,)
    if any(i is Ellipsis for i in idx):
        return idx + (slice(None),)
    return idx


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        a = self.arr
        vv, vt = _lift(value)
        val = a.val.astype(np.result_type(a.val.dtype, vv.dtype), copy=True)
        val[self.idx] = vv
        tan = a.tan
        if tan is None and vt is None:
            return Arr(val)
        n = (vt if vt is not None else tan).shape[-1]
        dt = np.result_type(val.dtype, *(t.dtype for t in (tan, vt) if t is not None))
        tan = np.zeros(val.shape + (n,), dtype=dt) if tan is None else tan.astype(dt, copy=True)
        tidx = _tan_index(self.idx)
        if vt is None:
            tan[tidx] = 0.0
        else:
            tan[tidx] = vt
        return Arr(val, tan)


class Arr:
    """Immutable array with an optional tangent bundle."""
    __array_priority__ = 1000
    __array_ufunc__ = None          # NumPy defers to our reflected operators

    def __init__(self, val, tan=None):
        self.val = np.asarray(val)
        if tan is not None:
            tan = np.asarray(tan)
            assert tan.shape[:-1] == self.val.shape, (tan.shape, self.val.shape)
        self.tan = tan

    # -- protocol ---------------------------------------------------------------------
    shape = property(lambda s: s.val.shape)
    ndim = property(lambda s: s.val.ndim)
    size = property(lambda s: s.val.size)
    dtype = property(lambda s: s.val.dtype)
    at = property(lambda s: _At(s))

    def __len__(self):
        return len(self.val)

    def __iter__(self):
        for i in range(len(self.val)):
            yield self[i]

    def __array__(self, dtype=None, copy=None):
        return np.array(self.val, dtype=dtype) if dtype is not None else np.array(self.val)

    def __array_function__(self, func, types, args, kwargs):
        # plain NumPy functions applied to an Arr (np.max, np.copy, np.mean ...): value only
        def strip(o):
            if isinstance(o, Arr):
                return o.val
            if isinstance(o, (list, tuple)):
                return type(o)(strip(x) for x in o)
            return o
        return func(*strip(args), **{k: strip(v) for k, v in kwargs.items()})

    def __float__(self):
        return float(self.val)

    def __int__(self):
        return int(self.val)

    def __bool__(self):
        return bool(self.val)

    def __repr__(self):
        return f"Arr({self.val!r}{', dual' if self.tan is not None else ''})"

    def to_py(self):
        return np.array(self.val)

    def item(self):
        return self.val.item()

    def __getitem__(self, idx):
        if isinstance(idx, Arr):
            idx = idx.val
        if isinstance(idx, tuple):
            idx = tuple(i.val if isinstance(i, Arr) else i for i in idx)
        return Arr(self.val[idx], None if self.tan is None else self.tan[_tan_index(idx)])

    @property
    def T(self):
        if self.tan is None:
            return Arr(self.val.T)
        nd = self.val.ndim
        return Arr(self.val.T, self.tan.transpose(tuple(range(nd - 1, -1, -1)) + (nd,)))

    def flatten(self):
        return reshape(self, (-1,))

    def reshape(self, *shape, order='C'):
        if len(shape) == 1 and not np.isscalar(shape[0]):
            shape = shape[0]
        return reshape(self, shape, order)

    def astype(self, dt):
        return Arr(self.val.astype(dt), self.tan)

    # -- arithmetic -------------------------------------------------------------------
    def __neg__(self):
        return Arr(-self.val, None if self.tan is None else -self.tan)

    def __pos__(self):
        return self

    def __add__(self, o):
        av, at = self.val, self.tan
        bv, bt = _lift(o)
        rv = av + bv
        if at is None and bt is None:
            return Arr(rv)
        ta = 0.0 if at is None else _bt(at, av.shape, rv.shape)
        tb = 0.0 if bt is None else _bt(bt, bv.shape, rv.shape)
        return Arr(rv, ta + tb)

    __radd__ = __add__

    def __sub__(self, o):
        return self + (-o if isinstance(o, Arr) else -np.asarray(o))

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        av, at = self.val, self.tan
        bv, bt = _lift(o)
        rv = av * bv
        if at is None and bt is None:
            return Arr(rv)
        t = 0.0
        if at is not None:
            t = t + _bt(at, av.shape, rv.shape) * _bv(bv, rv.shape)
        if bt is not None:
            t = t + _bv(av, rv.shape) * _bt(bt, bv.shape, rv.shape)
        return Arr(rv, t)

    __rmul__ = __mul__

    def __truediv__(self, o):
        av, at = self.val, self.tan
        bv, bt = _lift(o)
        rv = av / bv
        if at is None and bt is None:
            return Arr(rv)
        t = 0.0
        if at is not None:
            t = t + _bt(at, av.shape, rv.shape) / _bv(bv, rv.shape)
        if bt is not None:
            t = t - _bv(rv / bv, rv.shape) * _bt(bt, bv.shape, rv.shape)
        return Arr(rv, t)

    def __rtruediv__(self, o):
        return Arr(np.asarray(o)) / self

    def __pow__(self, p):
        if isinstance(p, Arr):
            if p.tan is not None:
                raise NotImplementedError("dual exponent")
            p = p.val
        rv = self.val ** p
        if self.tan is None:
            return Arr(rv)
        return Arr(rv, (p * self.val ** (p - 1))[..., None] * self.tan)

    def __matmul__(self, o):
        return matmul(self, o)

    def __rmatmul__(self, o):
        return matmul(o, self)

    # comparisons act on values
    def __lt__(self, o): return self.val < _lift(o)[0]
    def __le__(self, o): return self.val <= _lift(o)[0]
    def __gt__(self, o): return self.val > _lift(o)[0]
    def __ge__(self, o): return self.val >= _lift(o)[0]
    def __eq__(self, o): return self.val == _lift(o)[0]
    def __ne__(self, o): return self.val != _lift(o)[0]
    __hash__ = None


ndarray = Arr


# ------------------------------------------------------------------------------------------
# linear algebra
# ------------------------------------------------------------------------------------------
def _mm_tan_left(at, bv):
    """d(a @ b) for the a-tangent: at has shape a.shape + (n,)"""
    atf = np.moveaxis(at, -1, 0)                       # (n,) + a.shape
    r = np.matmul(atf, bv)
    return np.moveaxis(r, 0, -1)


def _mm_tan_right(av, bt, b_ndim):
    btf = np.moveaxis(bt, -1, 0)                       # (n,) + b.shape
    if b_ndim == 1:
        r = np.einsum('...k,nk->n...', av, btf)
    else:
        r = np.matmul(av, btf)
    return np.moveaxis(r, 0, -1)


def matmul(a, b):
    av, at = _lift(a)
    bv, bt = _lift(b)
    rv = np.matmul(av, bv)
    if at is None and bt is None:
        return Arr(rv)
    if av.ndim > 2 or bv.ndim > 2:
        raise NotImplementedError("matmul of duals: 1-D / 2-D operands only")
    t = 0.0
    if at is not None:
        t = t + _mm_tan_left(at, bv)
    if bt is not None:
        t = t + _mm_tan_right(av, bt, bv.ndim)
    return Arr(rv, t)


def dot(a, b):
    av, bv = _lift(a)[0], _lift(b)[0]
    if av.ndim == 0 or bv.ndim == 0:
        return _as_arr(a) * b
    return matmul(a, b)


class _Linalg:
    @staticmethod
    def norm(x, ord=None):
        """Euclidean norm of a vector; gradient x/||x|| like JAX (analytic extension sqrt(sum x^2)
        for complex-step inputs)."""
        if ord not in (None, 2):
            raise NotImplementedError("norm: ord 2 only")
        x = _as_arr(x)
        if x.ndim != 1:
            raise NotImplementedError("norm: vectors only")
        return sqrt(sum(x * x))


linalg = _Linalg()


# ------------------------------------------------------------------------------------------
# constructors / shape ops
# ------------------------------------------------------------------------------------------
def _as_arr(x):
    return x if isinstance(x, Arr) else Arr(np.asarray(x))


def _stack_tree(obj):
    """nested lists/tuples of scalars / Arr -> Arr (like jnp.array on a nested list)"""
    if isinstance(obj, Arr):
        return obj
    if isinstance(obj, (list, tuple)):
        return stack([_stack_tree(o) for o in obj], axis=0)
    return Arr(np.asarray(obj))


def array(obj, dtype=None):
    r = _stack_tree(obj)
    if dtype is not None:
        r = r.astype(dtype)
    elif r.val.dtype.kind in 'iub' and not isinstance(obj, (Arr, np.ndarray)):
        pass
    return Arr(np.array(r.val), None if r.tan is None else np.array(r.tan))


asarray = array


def zeros(shape, dtype=np.float64):
    return Arr(np.zeros(shape, dtype=dtype))


def ones(shape, dtype=np.float64):
    return Arr(np.ones(shape, dtype=dtype))


def zeros_like(a):
    return Arr(np.zeros_like(_lift(a)[0]))


def ones_like(a):
    return Arr(np.ones_like(_lift(a)[0]))


def eye(n, m=None):
    return Arr(np.eye(n, m))


def arange(*a, **k):
    return Arr(np.arange(*a, **k))


def linspace(*a, **k):
    return Arr(np.linspace(*a, **k))


def _tan_or_zeros(v, t, n, dt):
    return np.zeros(v.shape + (n,), dtype=dt) if t is None else t


def _combine(parts, fn_val, fn_tan):
    lifted = [_lift(p) for p in parts]
    rv = fn_val([v for v, _ in lifted])
    tans = [t for _, t in lifted if t is not None]
    if not tans:
        return Arr(rv)
    n, dt = tans[0].shape[-1], np.result_type(*[t.dtype for t in tans])
    return Arr(rv, fn_tan([_tan_or_zeros(v, t, n, dt) for v, t in lifted]))


def _axis(ax, ndim):
    return ax if ax >= 0 else ax + ndim


def concatenate(parts, axis=0):
    nd = _lift(parts[0])[0].ndim
    ax = _axis(axis, nd)
    return _combine(list(parts), lambda vs: np.concatenate(vs, axis=ax),
                    lambda ts: np.concatenate(ts, axis=ax))


def stack(parts, axis=0):
    nd = _lift(parts[0])[0].ndim + 1
    ax = _axis(axis, nd)
    return _combine(list(parts), lambda vs: np.stack(vs, axis=ax), lambda ts: np.stack(ts, axis=ax))


def vstack(parts):
    parts = [_as_arr(p) for p in parts]
    parts = [p if p.ndim >= 2 else reshape(p, (1, -1)) for p in parts]
    return concatenate(parts, axis=0)


def hstack(parts):
    parts = [_as_arr(p) for p in parts]
    if parts[0].ndim == 1:
        return concatenate(parts, axis=0)
    return concatenate(parts, axis=1)


def block(rows):
    return concatenate([concatenate([_as_arr(b) for b in r], axis=1) for r in rows], axis=0)


def diag(a):
    a = _as_arr(a)
    if a.tan is not None:
        raise NotImplementedError("diag of a dual")
    return Arr(np.diag(a.val))


def reshape(a, shape, order='C'):
    a = _as_arr(a)
    if np.isscalar(shape):
        shape = (shape,)
    shape = tuple(int(s) for s in shape)
    rv = np.reshape(a.val, shape, order=order)
    if a.tan is None:
        return Arr(rv)
    n = a.tan.shape[-1]
    tf = np.moveaxis(a.tan, -1, 0)                      # (n,) + s
    tf = np.stack([np.reshape(tf[i], shape, order=order) for i in range(n)], axis=0) if order != 'C' \
        else tf.reshape((n,) + rv.shape)
    return Arr(rv, np.moveaxis(tf, 0, -1))


def repeat(a, repeats, axis=None):
    a = _as_arr(a)
    if axis is None:                                   # flattened, like numpy
        a = reshape(a, (-1,))
        axis = 0
    ax = _axis(axis, a.ndim)
    return Arr(np.repeat(a.val, repeats, axis=ax),
               None if a.tan is None else np.repeat(a.tan, repeats, axis=ax))


def transpose(a):
    return _as_arr(a).T


def sum(a, axis=None):                                   # noqa: A001 (mirrors jnp.sum)
    a = _as_arr(a)
    if axis is None:
        tan = None if a.tan is None else a.tan.reshape(-1, a.tan.shape[-1]).sum(axis=0)
        return Arr(a.val.sum(), tan)
    ax = _axis(axis, a.ndim)
    return Arr(a.val.sum(axis=ax), None if a.tan is None else a.tan.sum(axis=ax))


def mean(a, axis=None):
    a = _as_arr(a)
    cnt = a.val.size if axis is None else a.val.shape[_axis(axis, a.ndim)]
    return sum(a, axis) / cnt


def max(a, axis=None):                                   # noqa: A001
    a = _as_arr(a)
    if a.tan is not None:
        if axis is not None:
            raise NotImplementedError("max of a dual along an axis")
        k = int(np.argmax(a.val.real))
        return Arr(a.val.reshape(-1)[k], a.tan.reshape(-1, a.tan.shape[-1])[k])
    return Arr(a.val.max() if axis is None else a.val.max(axis=axis))


def min(a, axis=None):                                   # noqa: A001
    return -max(-_as_arr(a), axis)


def median(a, axis=None):
    return Arr(np.median(_lift(a)[0], axis=axis))


def maximum(a, b):
    a, b = _as_arr(a), _as_arr(b)
    pick = a.val.real >= b.val.real
    return where(pick, a, b)


def where(c, a, b):
    c = _lift(c)[0]
    a, b = _as_arr(a), _as_arr(b)
    rv = np.where(c, a.val, b.val)
    if a.tan is None and b.tan is None:
        return Arr(rv)
    n = (a.tan if a.tan is not None else b.tan).shape[-1]
    ta = _bt(_tan_or_zeros(a.val, a.tan, n, float), a.val.shape, rv.shape)
    tb = _bt(_tan_or_zeros(b.val, b.tan, n, float), b.val.shape, rv.shape)
    return Arr(rv, np.where(np.broadcast_to(c, rv.shape)[..., None], ta, tb))


# ------------------------------------------------------------------------------------------
# elementwise functions
# ------------------------------------------------------------------------------------------
def _unary(fv, fd):
    def f(a):
        a = _as_arr(a)
        rv = fv(a.val)
        if a.tan is None:
            return Arr(rv)
        return Arr(rv, fd(a.val, rv)[..., None] * a.tan)
    return f


sqrt = _unary(np.sqrt, lambda x, r: 0.5 / r)
sin = _unary(np.sin, lambda x, r: np.cos(x))
cos = _unary(np.cos, lambda x, r: -np.sin(x))
exp = _unary(np.exp, lambda x, r: r)
log = _unary(np.log, lambda x, r: 1.0 / x)
tanh = _unary(np.tanh, lambda x, r: 1.0 - r * r)


def _sign(x):
    return np.sign(x.real)


# |x| with derivative sign(x) (0 at 0, as JAX); for complex-step inputs the analytic extension
# sign(Re x) * x
abs = _unary(lambda x: _sign(x) * x if np.iscomplexobj(x) else np.abs(x),      # noqa: A001
             lambda x, r: _sign(x))


def square(a):
    a = _as_arr(a)
    return a * a


# ------------------------------------------------------------------------------------------
# transformations
# ------------------------------------------------------------------------------------------
def jit(fn=None, static_argnums=None, **_):
    if fn is None:
        return lambda f: f
    return fn


def _tree_map(f, *trees):
    t0 = trees[0]
    if isinstance(t0, (tuple, list)):
        return type(t0)(_tree_map(f, *xs) for xs in zip(*trees))
    return f(*trees)


def _tree_stack(trees):
    t0 = trees[0]
    if isinstance(t0, (tuple, list)):
        return type(t0)(_tree_stack([t[i] for t in trees]) for i in range(len(t0)))
    if isinstance(t0, (bool, np.bool_)):
        return Arr(np.array(trees))
    return stack([_as_arr(t) for t in trees], axis=0)


def vmap(fn, in_axes=0, out_axes=0):
    if in_axes != 0 or out_axes != 0:
        raise NotImplementedError("vmap: axis 0 only")

    def mapped(*args):
        n = len(args[0])
        for a in args:
            if len(a) != n:
                raise ValueError("vmap: mismatched leading axes")
        return _tree_stack([fn(*[a[i] for a in args]) for i in range(n)])
    return mapped


def jacfwd(fn, argnums=0):
    if isinstance(argnums, (tuple, list)):
        raise NotImplementedError("jacfwd: single argnum")

    def jac(*args):
        x = _as_arr(args[argnums])
        if x.tan is not None or any(isinstance(a, Arr) and a.tan is not None for a in args):
            raise NotImplementedError("nested forward mode (use hessian / complex step)")
        n = x.val.size
        seed = np.eye(n, dtype=x.val.dtype if x.val.dtype.kind == 'c' else np.float64)
        xs = Arr(x.val, seed.reshape(x.val.shape + (n,)))
        out = fn(*[xs if i == argnums else a for i, a in enumerate(args)])

        def extract(o):
            o = _as_arr(o)
            if o.tan is None:
                return Arr(np.zeros(o.val.shape + x.val.shape))
            return Arr(o.tan.reshape(o.val.shape + x.val.shape))
        return _tree_map(extract, out)
    return jac


jacrev = jacfwd          # same Jacobian, the mode only changes the cost


def grad(fn, argnums=0):
    j = jacfwd(fn, argnums)

    def g(*args):
        out = j(*args)
        return out
    return g


def hessian(fn, argnums=0, step=1e-30):
    """d/dx of the forward-mode gradient by the complex-step method (exact to rounding)."""
    g = jacfwd(fn, argnums)

    def hess(*args):
        x = np.asarray(_lift(args[argnums])[0], dtype=np.float64)
        n = x.size
        cols = []
        for j in range(n):
            xp = x.astype(np.complex128).reshape(-1)
            xp[j] += 1j * step
            gj = g(*[Arr(xp.reshape(x.shape)) if i == argnums else a for i, a in enumerate(args)])
            cols.append(np.imag(gj.val).reshape(-1) / step)
        H = np.stack(cols, axis=-1)                       # (n_out..., n) : d grad_i / d x_j
        return Arr(H.reshape(x.shape + x.shape))
    return hess


class _Config:
    def update(self, *a, **k):
        return None


config = _Config()


def fori_loop(lower, upper, body, init):
    val = init
    for i in range(int(lower), int(upper)):
        val = body(i, val)
    return val
