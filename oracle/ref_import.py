"""Import the UNMODIFIED reference (`/root/reference/sres`) in this container.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- used by oracle/gen_golden.py to pin the oracle and by the reference arm of
bench.py (which, on the GPU box, finds the staged copy oracle/_ref made by oracle/make_ref.py); never imported by
the product or by any test that runs on the GPU box.

The reference needs hydra / omegaconf / xarray / parse / netCDF4 / matplotlib ..., none of which
is installed here.  We register stub modules for exactly those third-party packages:
  * a generic stub whose attributes are dummy classes (enough for `class X(initialize)`, type
    annotations and module-level imports), and
  * a ~100-line mini `xarray.DataArray` that implements the handful of operations the hot path's
    data code calls (`values/dims/coords/attrs/shape/sizes`, `isel`, `sel`, `mean/std(dim,skipna)`,
    arithmetic, `copy(data=)`, `transpose`, `xa.concat`), backed by numpy nan-aware reductions --
    the same numpy routines real xarray dispatches to for float32 data without bottleneck.
Everything under `sres.*` is the reference's own code, executed as is.
"""
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np

import os as _os

# the reference tree in the build container, else the byte-for-byte staged copy (oracle/make_ref.py) that travels to
# the GPU box for the reference arm of bench.py
_STAGED = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = "/root/reference" if _os.path.isdir("/root/reference/sres") else _STAGED


def available() -> bool:
    return _os.path.isdir(_os.path.join(REFERENCE_ROOT, "sres"))
_STUB_TOPLEVEL = ("hydra", "omegaconf", "xarray", "parse", "netCDF4", "matplotlib", "ipywidgets", "zarr", "dask",
                  "h5py", "nvidia", "IPython", "cartopy", "ipympl", "modulus", "cftime")


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __class_getitem__(cls, item):
        return cls


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_TOPLEVEL:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        if module.__name__ == "xarray":
            module.DataArray = DataArray
            module.concat = concat


# ------------------------------------------------------------------------------------------
# mini xarray
# ------------------------------------------------------------------------------------------
class _Coords(dict):
    pass


class DataArray:
    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        self.values = np.asarray(data)
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(self.values.ndim))
        self.coords = _Coords()
        for k, v in (coords or {}).items():
            self.coords[k] = DataArray(np.asarray(v), dims=[k]) if not isinstance(v, DataArray) else v
        self.attrs = dict(attrs or {})
        self.name = name

    # -- structure ------------------------------------------------------------------------
    @property
    def shape(self):
        return self.values.shape

    @property
    def ndim(self):
        return self.values.ndim

    @property
    def size(self):
        return self.values.size

    @property
    def sizes(self):
        return dict(zip(self.dims, self.values.shape))

    def _sub_coords(self, dim, index):
        out = {}
        for k, c in self.coords.items():
            if k == dim:
                v = c.values[index]
                if np.ndim(v) > 0:
                    out[k] = v
            elif k in self.dims:
                out[k] = c.values
        return out

    def isel(self, **idx):
        out = self
        for dim, index in idx.items():
            ax = out.dims.index(dim)
            sl = [slice(None)] * out.values.ndim
            sl[ax] = index
            data = out.values[tuple(sl)]
            dims = out.dims if data.ndim == out.values.ndim else tuple(d for d in out.dims if d != dim)
            out = DataArray(data, dims=dims, coords=out._sub_coords(dim, index), attrs=out.attrs)
        return out

    def sel(self, **idx):
        out = self
        for dim, label in idx.items():
            pos = list(out.coords[dim].values.tolist()).index(label)
            out = out.isel(**{dim: pos})
        return out

    def _reduce(self, fn, dim, skipna, keep_attrs=False, **kw):
        dims = [dim] if isinstance(dim, str) else list(dim)
        axes = tuple(self.dims.index(d) for d in dims)
        f = getattr(np, ("nan" + fn) if skipna else fn)
        data = f(self.values, axis=axes, **kw)
        rdims = tuple(d for d in self.dims if d not in dims)
        coords = {k: c.values for k, c in self.coords.items() if k in rdims}
        return DataArray(data, dims=rdims, coords=coords, attrs=self.attrs if keep_attrs else None)

    def mean(self, dim=None, skipna=True, keep_attrs=False):
        if dim is None:
            return DataArray((np.nanmean if skipna else np.mean)(self.values))
        return self._reduce("mean", dim, skipna, keep_attrs)

    def std(self, dim=None, skipna=True, keep_attrs=False):
        if dim is None:
            return DataArray((np.nanstd if skipna else np.std)(self.values))
        return self._reduce("std", dim, skipna, keep_attrs)

    def max(self, dim=None, skipna=True, keep_attrs=False):
        return self._reduce("max", dim, skipna, keep_attrs)

    def min(self, dim=None, skipna=True, keep_attrs=False):
        return self._reduce("min", dim, skipna, keep_attrs)

    def __format__(self, spec):
        return format(float(self.values), spec)

    def __float__(self):
        return float(self.values)

    # -- arithmetic with dimension-name broadcasting --------------------------------------
    def _binary(self, other, op):
        if isinstance(other, DataArray):
            ov = other.values
            if other.dims != self.dims:
                shape = [self.values.shape[i] if d in other.dims else 1 for i, d in enumerate(self.dims)]
                order = [other.dims.index(d) for d in self.dims if d in other.dims]
                ov = np.transpose(ov, order).reshape(shape)
        else:
            ov = other
        return DataArray(op(self.values, ov), dims=self.dims, coords={k: c.values for k, c in self.coords.items()})

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __add__(self, o):
        return self._binary(o, np.add)

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    def __truediv__(self, o):
        return self._binary(o, np.divide)

    def copy(self, data=None, deep=True):
        return DataArray(self.values.copy() if data is None else data, dims=self.dims,
                         coords={k: c.values for k, c in self.coords.items()}, attrs=dict(self.attrs))

    def transpose(self, *dims):
        order = [self.dims.index(d) for d in dims]
        return DataArray(np.transpose(self.values, order), dims=dims,
                         coords={k: c.values for k, c in self.coords.items()}, attrs=self.attrs)

    def squeeze(self):
        keep = [i for i, s in enumerate(self.values.shape) if s != 1]
        return DataArray(self.values.squeeze(), dims=[self.dims[i] for i in keep])

    def __getitem__(self, key):
        return DataArray(self.values[key], dims=self.dims if np.ndim(self.values[key]) == self.ndim else None,
                         attrs=self.attrs)


def concat(arrays, dim):
    """xa.concat(list, dim) where `dim` is a coordinate DataArray naming a NEW leading dimension
    (the only form the hot path uses: swot/raw.py:211)."""
    name = dim.dims[0] if isinstance(dim, DataArray) else dim
    data = np.stack([a.values for a in arrays], axis=0)
    coords = {k: c.values for k, c in arrays[0].coords.items() if k in arrays[0].dims}
    if isinstance(dim, DataArray):
        coords[name] = dim.values
    return DataArray(data, dims=(name,) + tuple(arrays[0].dims), coords=coords)


_installed = False


def install():
    """Put the stubs + the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


class Cfg(dict):
    """Attribute + mapping access, like the DictConfig the reference reads through cfg()."""
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def __setattr__(self, k, v):
        self[k] = v

    def get(self, k, default=None):
        v = dict.get(self, k, default)
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v


def set_cfg(model: dict, task: dict, **groups):
    """Install a configuration as the reference's process-global cfg() (util/config.py:21-22)."""
    install()
    from sres.base.util.config import ConfigContext
    ConfigContext.cfg = Cfg(model=Cfg(model), task=Cfg(task), **{k: Cfg(v) for k, v in groups.items()})
    return ConfigContext.cfg
