"""Execute the reference's OWN source text (read from /root/reference at call time) on minijax.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden_ref.py (which mints the committed
``tests/golden/ref_*.npz`` fixtures) and by tests/test_reference_exec.py (which re-executes the
reference in this container and compares with those fixtures).  /root/reference does not exist
on the GPU box: nothing that runs there imports this module.

The reference scripts are flat: imports, module constants, ``class Model``, then the experiment
driver (loops over alphas, OSQP / IPOPT solves, plots) at module level.  ``load_script`` executes
the top-level statements up to the first driver statement (the first ``if`` / ``for`` / ``while``
/ ``with`` after ``class Model``), with stub modules standing in for what the image lacks:
``jax`` -> minijax (array ops + forward-mode AD on NumPy), ``osqp`` / ``ipyopt`` / ``matplotlib`` /
``seaborn`` -> inert stubs (the model code never calls them).  No reference source is copied into
this repository; it is read where it lies.
"""
import ast
import contextlib
import io
import os
import sys
import textwrap
import types

from . import minijax

REFERENCE_ROOT = os.environ.get("SAA_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "drone"))


class _Inert:
    """object that absorbs attribute access, calls, item assignment (matplotlib rcParams...)"""

    def __init__(self, name="stub"):
        self.__dict__["_name"] = name

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Inert(f"{self._name}.{k}")

    def __call__(self, *a, **k):
        return _Inert(self._name + "()")

    def __setitem__(self, k, v):
        pass

    def __getitem__(self, k):
        return _Inert(self._name + "[]")

    def __setattr__(self, k, v):
        pass


class _StubModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Inert(f"{self.__name__}.{k}")


def _stub_modules(extra=None):
    import scipy.stats
    jax = types.ModuleType("jax")
    for name in ("jit", "vmap", "jacfwd", "jacrev", "grad", "hessian"):
        setattr(jax, name, getattr(minijax, name))
    jax.numpy = minijax
    jax.config = types.ModuleType("jax.config")
    jax.config.config = minijax.config
    jax.config.update = minijax.config.update
    jax.lax = types.ModuleType("jax.lax")
    jax.lax.fori_loop = minijax.fori_loop
    jax.scipy = types.ModuleType("jax.scipy")
    jax.scipy.stats = types.ModuleType("jax.scipy.stats")
    jax.scipy.stats.norm = scipy.stats.norm
    mods = {"jax": jax, "jax.numpy": minijax, "jax.config": jax.config, "jax.lax": jax.lax,
            "jax.scipy": jax.scipy, "jax.scipy.stats": jax.scipy.stats}
    for name in ("osqp", "ipyopt", "seaborn", "matplotlib", "matplotlib.pylab", "matplotlib.pyplot",
                 "matplotlib.patches", "matplotlib.lines", "matplotlib.colors", "matplotlib.cm"):
        mods[name] = _StubModule(name)
    mods.update(extra or {})
    return mods


@contextlib.contextmanager
def _patched(script_dir, extra=None):
    mods = _stub_modules(extra)
    saved = {k: sys.modules.get(k) for k in mods}
    before = set(sys.modules)
    sys.modules.update(mods)
    sys.path.insert(0, script_dir)
    try:
        yield
    finally:
        sys.path.remove(script_dir)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        # sibling modules of the script (drone_params, drone_utils, driving_params, ...)
        for k in set(sys.modules) - before:
            f = getattr(sys.modules[k], "__file__", None) or ""
            if f.startswith(REFERENCE_ROOT):
                del sys.modules[k]


def _driver_start(tree, after_class="Model"):
    """index of the first top-level driver statement after ``class Model`` (None: keep all)"""
    seen = False
    for i, node in enumerate(tree.body):
        if isinstance(node, ast.ClassDef) and node.name == after_class:
            seen = True
        elif seen and isinstance(node, (ast.If, ast.For, ast.While, ast.With)):
            return i
    return None


def load_script(relpath, inject=None, extra_modules=None, quiet=True):
    """Execute ``/root/reference/<relpath>`` up to its experiment driver.  -> module object."""
    path = os.path.join(REFERENCE_ROOT, relpath)
    src = open(path).read()
    tree = ast.parse(src, filename=path)
    cut = _driver_start(tree)
    if cut is not None:
        tree.body = tree.body[:cut]
    mod = types.ModuleType("reference_" + os.path.splitext(os.path.basename(relpath))[0])
    mod.__file__ = path
    mod.__dict__.update(inject or {})
    code = compile(tree, path, "exec")
    out = io.StringIO()
    with _patched(os.path.dirname(path), extra_modules):
        with contextlib.redirect_stdout(out if quiet else sys.stdout):
            exec(code, mod.__dict__)
    mod.__reference_stdout__ = out.getvalue()
    return mod


def load_nested(relpath, container_type, class_name="Model", inject=None, quiet=True):
    """For scripts that declare ``class Model`` inside a driver loop (drone/drone_times.py:
    ``for M in [20, 30, 50]:``): execute the statements before the loop, then the class
    definition alone (dedented) with ``inject`` (e.g. ``M``) bound as module globals."""
    path = os.path.join(REFERENCE_ROOT, relpath)
    src = open(path).read()
    tree = ast.parse(src, filename=path)
    idx = cls = None
    for i, node in enumerate(tree.body):
        if isinstance(node, container_type):
            for sub in ast.walk(node):
                if isinstance(sub, ast.ClassDef) and sub.name == class_name:
                    idx, cls = i, sub
                    break
        if cls is not None:
            break
    if cls is None:
        raise ValueError(f"{class_name} not found inside a {container_type.__name__} of {relpath}")
    head = ast.Module(body=tree.body[:idx], type_ignores=[])
    mod = types.ModuleType("reference_" + os.path.splitext(os.path.basename(relpath))[0])
    mod.__file__ = path
    out = io.StringIO()
    with _patched(os.path.dirname(path)):
        with contextlib.redirect_stdout(out if quiet else sys.stdout):
            exec(compile(head, path, "exec"), mod.__dict__)
            mod.__dict__.update(inject or {})
            lines = src.splitlines()[cls.lineno - 1 - len(cls.decorator_list):cls.end_lineno]
            csrc = "\n" * (cls.lineno - 1) + textwrap.dedent("\n".join(lines))   # keep line numbers
            exec(compile(csrc, path, "exec"), mod.__dict__)
    return mod


def extract_functions(relpath, names, namespace):
    """Define the named (possibly nested, e.g. inside ``if B_validate_monte_carlo:``) functions of
    a script in ``namespace`` (a dict, normally a loaded module's ``__dict__``)."""
    path = os.path.join(REFERENCE_ROOT, relpath)
    src = open(path).read()
    tree = ast.parse(src, filename=path)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            lines = src.splitlines()[node.lineno - 1 - len(node.decorator_list):node.end_lineno]
            csrc = "\n" * (node.lineno - 1) + textwrap.dedent("\n".join(lines))
            with _patched(os.path.dirname(path)):
                exec(compile(csrc, path, "exec"), namespace)
            found[node.name] = namespace[node.name]
    missing = set(names) - set(found)
    if missing:
        raise ValueError(f"{missing} not found in {relpath}")
    return found
