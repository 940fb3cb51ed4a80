"""Import shim: lets the UNMODIFIED reference (baseline/_ref install, else /root/reference) import without its absent deps.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py and by the optional
oracle-vs-reference checks in tests/ when /root/reference is mounted).  It is never
imported by the product package `qiskit_dynamics_b200`.

It fabricates minimal stand-ins for the third-party packages the reference imports but
that are absent here (arraylias, qiskit, multiset, matplotlib) -- numpy dispatch only.
Origin: SURVEY.md Appendix B (our own stub; contains no reference source).
"""
import sys, types, importlib.abc, importlib.machinery
import numpy as np, scipy

class LibraryError(Exception): pass

class _Proxy:                      # "aliased module": attribute access -> dispatching function / sub-module
    def __init__(self, alias, like=None, prefix=""): self._a, self._like, self._p = alias, like, prefix
    def __getattr__(self, name):
        path = f"{self._p}.{name}" if self._p else name
        obj = self._a._base
        try:
            for part in path.split("."): obj = getattr(obj, part)
        except AttributeError: obj = None
        if isinstance(obj, types.ModuleType): return _Proxy(self._a, self._like, path)
        return self._a._func(path, self._like)

class Alias:
    def __init__(self, base):
        self._base = base
        self._types = {np.ndarray: "numpy", float: "numpy", int: "numpy", complex: "numpy", np.number: "numpy"}
        self._funcs, self._defaults, self._fallbacks = {}, {}, {}
    def register_type(self, t, lib): self._types[t] = lib
    def registered_types(self): return tuple(self._types.keys())
    def registered_libs(self): return tuple(set(self._types.values()))
    def infer_libs(self, obj):
        if isinstance(obj, (list, tuple)): return self.infer_libs(obj[0]) if len(obj) else ()
        for t, lib in self._types.items():
            if isinstance(obj, t): return (lib,)
        return ()
    def _reg(self, table, key):
        def deco(f): table[key] = f; return f
        return deco
    def register_function(self, func=None, lib=None, path=None):
        if func is not None: self._funcs[(lib, path)] = func; return func
        return self._reg(self._funcs, (lib, path))
    def register_default(self, func=None, path=None): return self._reg(self._defaults, path)
    def register_fallback(self, func=None, path=None): return self._reg(self._fallbacks, path)
    def _lib_of(self, like):
        if like is None: return None
        if isinstance(like, str): return like
        libs = self.infer_libs(like); return libs[0] if libs else "numpy"
    def _resolve(self, path, lib):
        if (lib, path) in self._funcs: return self._funcs[(lib, path)]
        obj = self._base
        try:
            for part in path.split("."): obj = getattr(obj, part)
            if lib in (None, "numpy"): return obj
        except AttributeError: obj = None
        if path in self._fallbacks: return self._fallbacks[path]
        if path in self._defaults: return self._defaults[path]
        if obj is not None: return obj
        raise LibraryError(path)
    def _func(self, path, like):
        lib = self._lib_of(like)
        if lib is not None:
            if lib == "numpy" and path in self._defaults and ("numpy", path) not in self._funcs \
               and not hasattr(self._base, path.split(".")[0]): return self._defaults[path]
            if lib == "numpy" and path == "asarray": return self._defaults.get(path, np.asarray)
            return self._resolve(path, lib)
        def dispatch(*args, **kwargs):
            l = None
            for a in list(args) + list(kwargs.values()):
                libs = self.infer_libs(a)
                if libs: l = libs[0]; break
            l = l or "numpy"
            if l == "numpy" and path == "asarray": return self._defaults.get(path, np.asarray)(*args, **kwargs)
            return self._resolve(path, l)(*args, **kwargs)
        return dispatch
    def __call__(self, like=None, path=None):
        if path is not None: return self._func(path, like if like is not None else "numpy")
        return _Proxy(self, like)

al = types.ModuleType("arraylias"); al.numpy_alias = lambda: Alias(np); al.scipy_alias = lambda: Alias(scipy)
ex = types.ModuleType("arraylias.exceptions"); ex.LibraryError = LibraryError; al.exceptions = ex
sys.modules["arraylias"] = al; sys.modules["arraylias.exceptions"] = ex

class QiskitError(Exception): pass
class Operator:                    # array wrapper: enough for tests that build Paulis / pass Operator inputs
    _P = {"I": np.eye(2), "X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1., -1.])}
    def __init__(self, data, *a, **k): self.data = np.asarray(getattr(data, "data", data), dtype=complex)
    def __array__(self, dtype=None, copy=None): return self.data if dtype is None else self.data.astype(dtype)
    @classmethod
    def from_label(cls, label):
        m = np.eye(1, dtype=complex)
        for ch in label: m = np.kron(m, cls._P[ch])
        return cls(m)
    def __mul__(self, o): return Operator(self.data * o)
    __rmul__ = __mul__
    def __add__(self, o): return Operator(self.data + np.asarray(o))
    def __sub__(self, o): return Operator(self.data - np.asarray(o))
    def __truediv__(self, o): return Operator(self.data / o)
    def __neg__(self): return Operator(-self.data)
    def __matmul__(self, o): return Operator(self.data @ np.asarray(o))
    shape = property(lambda s: s.data.shape); ndim = property(lambda s: s.data.ndim)
    def __len__(self): return len(self.data)
    def __getitem__(self, i): return self.data[i]
def is_hermitian_matrix(mat, rtol=1e-5, atol=1e-8):
    mat = np.array(mat); return mat.ndim == 2 and np.allclose(mat, mat.conj().T, rtol=rtol, atol=atol)
_REAL = {"QiskitError": QiskitError, "Operator": Operator, "is_hermitian_matrix": is_hermitian_matrix}
class _Meta(type):                 # fabricated classes tolerate arbitrary attribute access (type annotations)
    def __getattr__(cls, name):
        if name.startswith("__"): raise AttributeError(name)
        sub = _Meta(name, (), {"__init__": lambda self, *a, **k: None}); setattr(cls, name, sub); return sub
class _QMod(types.ModuleType):
    __path__ = []
    def __getattr__(self, name):
        if name.startswith("__"): raise AttributeError(name)
        if name in _REAL: return _REAL[name]
        cls = _Meta(name, (), {"__init__": lambda self, *a, **k: None}); setattr(self, name, cls); return cls
class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name == "qiskit" or name.startswith("qiskit."):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
    def create_module(self, spec): return _QMod(spec.name)
    def exec_module(self, module): pass
sys.meta_path.insert(0, _Finder())
for modname, setup in (("matplotlib", None), ("multiset", None)):
    try: __import__(modname)
    except ImportError:
        m = types.ModuleType(modname); sys.modules[modname] = m
        if modname == "matplotlib":
            p = types.ModuleType("matplotlib.pyplot"); p.axis = object; m.pyplot = p; sys.modules["matplotlib.pyplot"] = p
        else:
            m.Multiset = type("Multiset", (dict,), {}); m.FrozenMultiset = m.Multiset
import os as _os

# Where the UNMODIFIED reference lives: the git-ignored pip install under baseline/_ref (made once with
# `pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>`; it travels to the GPU box
# with gpurun), else the read-only mount of the dev container.
_REPO = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
_CANDIDATES = [_os.path.join(_REPO, "baseline", "_ref"), "/root/reference"]
REFERENCE_ROOT = next((c for c in _CANDIDATES if _os.path.isdir(_os.path.join(c, "qiskit_dynamics"))), _CANDIDATES[-1])


def reference_available() -> bool:
    return _os.path.isdir(_os.path.join(REFERENCE_ROOT, "qiskit_dynamics"))


if reference_available() and REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
