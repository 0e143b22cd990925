"""Torch-backed stand-in for the slice of the mid-2020 JAX API that
``/root/reference/lib/ode.py`` imports.

TEST INFRASTRUCTURE ONLY.  The reference cannot be imported in this image
(no jax / jaxlib, and it needs the 2020 API: ``jax.ops.index_update``,
``jax.linear_util``).  This module lets ``tests/golden/make_golden.py`` import
the *unmodified* reference solver source and run it eagerly on torch CPU
tensors, so that the reference's own control flow, tableaux, NFE accounting and
adjoint (``_odeint_rev`` through ``jax.vjp`` -> ``torch.func.vjp``) produce the
golden vectors the oracle is pinned against.  Nothing here is a solver: every
function is a thin eager re-spelling of a JAX primitive (``lax.while_loop`` is a
Python ``while``, ``jax.jit`` is the identity, ...).

Semantics kept from JAX with x64 disabled: floating arrays default to the torch
default dtype (float32 unless the caller switched it to float64, which is the
``jax_enable_x64`` analogue); Python scalars are weakly typed.
"""
import sys
import types

import torch


# ----------------------------------------------------------------------------
# pytrees (tuple / list / dict with sorted keys / leaves)
# ----------------------------------------------------------------------------
def _is_leaf(x):
    return not isinstance(x, (tuple, list, dict))


def _as_tensor(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(x, dtype=torch.get_default_dtype()
                           if isinstance(x, float) else None)


def tree_flatten(tree):
    """Returns (leaves, treedef) with dict keys visited in sorted order."""
    if isinstance(tree, dict):
        keys = sorted(tree)
        parts = [tree_flatten(tree[k]) for k in keys]
        return sum((p[0] for p in parts), []), ("dict", keys, [p[1] for p in parts])
    if isinstance(tree, (tuple, list)):
        parts = [tree_flatten(v) for v in tree]
        kind = "tuple" if isinstance(tree, tuple) else "list"
        return sum((p[0] for p in parts), []), (kind, None, [p[1] for p in parts])
    return [tree], ("leaf", None, None)


def _treedef_size(td):
    if td[0] == "leaf":
        return 1
    return sum(_treedef_size(c) for c in td[2])


def tree_unflatten(td, leaves):
    leaves = list(leaves)
    if td[0] == "leaf":
        return leaves[0]
    out, pos = [], 0
    for c in td[2]:
        n = _treedef_size(c)
        out.append(tree_unflatten(c, leaves[pos:pos + n]))
        pos += n
    if td[0] == "dict":
        return dict(zip(td[1], out))
    return tuple(out) if td[0] == "tuple" else out


def tree_map(f, tree, *rest):
    leaves, td = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(td, [f(*xs) for xs in zip(leaves, *others)])


def ravel_pytree(tree):
    leaves, td = tree_flatten(tree)
    leaves = [_as_tensor(l) for l in leaves]
    shapes = [l.shape for l in leaves]
    sizes = [l.numel() for l in leaves]
    if leaves:
        dt = leaves[0].dtype
        for l in leaves[1:]:
            dt = torch.promote_types(dt, l.dtype)
        flat = torch.cat([l.reshape(-1).to(dt) for l in leaves])
    else:
        flat = torch.zeros(0)

    def unravel(v):
        out, pos = [], 0
        for sh, n in zip(shapes, sizes):
            out.append(v[pos:pos + n].reshape(sh))
            pos += n
        return tree_unflatten(td, out)

    return flat, unravel


# ----------------------------------------------------------------------------
# jax.numpy
# ----------------------------------------------------------------------------
class _Arr(torch.Tensor):
    """Tensor that also accepts the one numpy idiom torch lacks: ``x[::-1]``."""

    def __getitem__(self, idx):
        base = self.as_subclass(torch.Tensor)
        if isinstance(idx, slice) and idx.step == -1 and idx.start is None and idx.stop is None:
            return torch.flip(base, [0])
        return base[idx]


def _array(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, (list, tuple)) and len(x) and any(isinstance(v, torch.Tensor) for v in x):
        return torch.stack([_as_tensor(v) for v in x])
    t = torch.as_tensor(x)
    if t.dtype == torch.float64 and dtype is None:
        t = t.to(torch.get_default_dtype())
    return t if dtype is None else t.to(dtype)


def _where(c, a, b):
    c = _as_tensor(c)
    if not isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
        a = _as_tensor(a)
    return torch.where(c, a, b)


def _polyval(p, x):
    y = torch.zeros_like(p[0] * x)
    for c in p:
        y = y * x + c
    return y


def _dot(a, b):
    return torch.matmul(a, b)


def _make_numpy():
    np = types.ModuleType("jax.numpy")
    np.array = _array
    np.asarray = _array
    np.dot = _dot
    np.matmul = torch.matmul
    np.zeros = lambda shape, dtype=None: torch.zeros(shape, dtype=dtype)
    np.ones = lambda shape, dtype=None: torch.ones(shape, dtype=dtype)
    np.empty = lambda shape, dtype=None: torch.zeros(shape, dtype=dtype)
    np.zeros_like = lambda x: torch.zeros_like(_as_tensor(x))
    np.where = _where
    np.concatenate = lambda xs, axis=0: torch.cat([_as_tensor(x) for x in xs], dim=axis)
    np.stack = lambda xs, axis=0: torch.stack(list(xs), dim=axis)
    np.polyval = _polyval
    np.all = lambda x: torch.all(_as_tensor(x))
    np.arange = lambda *a: torch.arange(*a)
    np.linspace = lambda a, b, n: torch.linspace(a, b, n)
    np.max = lambda x: torch.max(_as_tensor(x))
    np.min = lambda x: torch.min(_as_tensor(x))
    np.abs = torch.abs
    np.sin = torch.sin
    np.cos = torch.cos
    np.sqrt = torch.sqrt
    np.square = torch.square
    np.mean = torch.mean
    np.maximum = lambda a, b: torch.maximum(_as_tensor(a).to(_as_tensor(b).dtype) if not isinstance(a, torch.Tensor) else a,
                                            _as_tensor(b).to(a.dtype) if not isinstance(b, torch.Tensor) else b)
    np.minimum = lambda a, b: torch.minimum(_as_tensor(a).to(_as_tensor(b).dtype) if not isinstance(a, torch.Tensor) else a,
                                            _as_tensor(b).to(a.dtype) if not isinstance(b, torch.Tensor) else b)
    np.inf = float("inf")
    np.pi = 3.141592653589793
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.norm = lambda x: torch.linalg.norm(x)
    np.linalg = linalg
    return np


# ----------------------------------------------------------------------------
# jax.lax control flow, eagerly
# ----------------------------------------------------------------------------
def _while_loop(cond, body, init):
    s = init
    while bool(cond(s)):
        s = body(s)
    return s


def _fori_loop(lo, hi, body, init):
    s = init
    for i in range(int(lo), int(hi)):
        s = body(i, s)
    return s


def _scan(f, init, xs):
    carry, ys = init, []
    for x in xs:
        carry, y = f(carry, x)
        ys.append(y)
    leaves0, td = tree_flatten(ys[0])
    cols = [[tree_flatten(y)[0][j] for y in ys] for j in range(len(leaves0))]
    stacked = [torch.stack([_as_tensor(v) for v in col]).as_subclass(_Arr) for col in cols]
    return carry, tree_unflatten(td, stacked)


def _cond(pred, a_operand, a_fun, b_operand, b_fun):
    return a_fun(a_operand) if bool(pred) else b_fun(b_operand)


# ----------------------------------------------------------------------------
# jax.ops
# ----------------------------------------------------------------------------
class _Index:
    def __getitem__(self, idx):
        return idx


def _index_update(x, idx, v):
    out = x.clone()
    out[idx] = v
    return out


# ----------------------------------------------------------------------------
# transforms
# ----------------------------------------------------------------------------
def _jit(f=None, static_argnums=None, **_):
    if f is None:
        return lambda g: g
    return f


class _CustomVjp:
    """Primal evaluation only; ``defvjp`` just records the rules so that the
    golden generator can call the reference's fwd/rev rules explicitly."""

    def __init__(self, f, nondiff_argnums=()):
        self.f = f
        self.nondiff_argnums = nondiff_argnums
        self.fwd = self.bwd = None
        self.__name__ = getattr(f, "__name__", "custom_vjp")

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    def __call__(self, *a, **k):
        return self.f(*a, **k)


def _vmap(f, in_axes=0):
    def mapped(tree):
        leaves, td = tree_flatten(tree)
        n = leaves[0].shape[0]
        outs = [f(tree_unflatten(td, [l[i] for l in leaves])) for i in range(n)]
        o_leaves0, otd = tree_flatten(outs[0])
        cols = [[tree_flatten(o)[0][j] for o in outs] for j in range(len(o_leaves0))]
        return tree_unflatten(otd, [torch.stack([_as_tensor(v) for v in c]) for c in cols])
    return mapped


def _vjp(f, *primals):
    primals = tuple(tree_map(_as_tensor, p) for p in primals)
    out, fn = torch.func.vjp(f, *primals)
    return out, fn


class _WrappedFun:
    def __init__(self, f, transforms=()):
        self.f, self.transforms = f, transforms

    def call_wrapped(self, *args, **kwargs):
        gens = []
        for gen, static in self.transforms:
            g = gen(*static, *args, **kwargs)
            args, kwargs = next(g)
            gens.append(g)
        ans = self.f(*args, **kwargs)
        for g in reversed(gens):
            ans = g.send(ans)
        return ans


def _transformation(gen):
    def wrap(fun, *static):
        return _WrappedFun(fun.f, fun.transforms + ((gen, static),))
    return wrap


def _safe_map(f, *seqs):
    seqs = [list(s) for s in seqs]
    assert all(len(s) == len(seqs[0]) for s in seqs)
    return list(map(f, *seqs))


def _safe_zip(*seqs):
    seqs = [list(s) for s in seqs]
    assert all(len(s) == len(seqs[0]) for s in seqs)
    return list(zip(*seqs))


def install():
    """Registers the stand-in under ``sys.modules['jax']`` (idempotent)."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "_ENO_SHIM", False):
        return sys.modules["jax"]
    jax = types.ModuleType("jax")
    jax._ENO_SHIM = True
    np = _make_numpy()
    lax = types.ModuleType("jax.lax")
    lax.while_loop, lax.fori_loop, lax.scan, lax.cond = _while_loop, _fori_loop, _scan, _cond
    lax.min = np.minimum        # lax.min(cur_t + step_size, ts[-1]) in the fixed-grid loop (lib/ode.py:891)
    ops = types.ModuleType("jax.ops")
    ops.index = _Index()
    ops.index_update = _index_update
    util = types.ModuleType("jax.util")
    util.safe_map, util.safe_zip = _safe_map, _safe_zip
    flatten_util = types.ModuleType("jax.flatten_util")
    flatten_util.ravel_pytree = ravel_pytree
    tree_util = types.ModuleType("jax.tree_util")
    tree_util.tree_map = tree_map
    tree_util.tree_flatten = tree_flatten
    lu = types.ModuleType("jax.linear_util")
    lu.wrap_init = lambda f: _WrappedFun(f)
    lu.transformation = _transformation
    jax.numpy, jax.lax, jax.ops, jax.util = np, lax, ops, util
    jax.flatten_util, jax.tree_util, jax.linear_util = flatten_util, tree_util, lu
    jax.jit = _jit
    jax.custom_vjp = lambda f=None, nondiff_argnums=(): (
        _CustomVjp(f, nondiff_argnums) if f is not None
        else (lambda g: _CustomVjp(g, nondiff_argnums)))
    jax.vmap = _vmap
    jax.vjp = _vjp
    for name, mod in [("jax", jax), ("jax.numpy", np), ("jax.lax", lax), ("jax.ops", ops),
                      ("jax.util", util), ("jax.flatten_util", flatten_util),
                      ("jax.tree_util", tree_util), ("jax.linear_util", lu)]:
        sys.modules[name] = mod
    return jax


def load_reference_ode(path="/root/reference/lib/ode.py"):
    """Imports the unmodified reference solver module on top of the stand-in."""
    import importlib.util
    install()
    spec = importlib.util.spec_from_file_location("_eno_reference_ode", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
