"""NetCDF without netCDF4: the subset of the ``netCDF4.Dataset`` interface that the reference's restart,
topography, ocean and routing-network I/O uses, on top of ``scipy.io.netcdf_file`` (NetCDF-3 classic, 64-bit
offsets).  SURVEY 8(f4): boxes that run the drop-in (this image included) have no netCDF4 / HDF5.

Every call site of the reference goes through one of these forms (run_simulation.py:63-246,
topography.py:341-575, routing.py:79-175, generate_hydrology_maps.py:276-330):

    with Dataset(path, "w") as ds:
        ds.createDimension("lat", n); v = ds.createVariable("u", "f4", ("lat", "lon")); v[:] = a
        s = ds.createVariable("t_seconds", "f8"); s[...] = 1.0; ds.setncattr("title", "..."); v.units = "m"
    with Dataset(path, "r") as ds:
        ds.variables["u"][:].data; ds["lat"][:]; float(ds.variables["t_seconds"][...]); ds.getncattr("day")
        ds.dimensions["lat"].size; "lake_id" in ds.variables

``install_netcdf4_shim()`` registers this module as ``netCDF4`` when the real package is missing, so the
reference's own save_restart / load_restart / RiverRouting / load_topography_from_netcdf run unmodified on files
written here, and vice versa (the real netCDF4 reads NetCDF-3 classic files).  NetCDF-4/HDF5 files written by a
real netCDF4 cannot be read without it; opening one raises a clear error.

Type notes: NetCDF-3 has no unsigned or 64-bit integer types.  ``u1`` variables (land_mask) are stored as signed
bytes and ``i8`` (flow_to_index, flow_order) as 32-bit integers -- both with a ``_qd_dtype`` attribute, and are
converted back on read.  An ``i8`` value outside the int32 range raises instead of wrapping.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
from scipy.io import netcdf_file

_STORE = {"u1": "i1", "uint8": "i1", "B": "i1", "i8": "i4", "int64": "i4", "u2": "i4", "u4": "i4",
          "f4": "f4", "float32": "f4", "f": "f4", "f8": "f8", "float64": "f8", "d": "f8",
          "i1": "i1", "b": "i1", "i2": "i2", "h": "i2", "i4": "i4", "int32": "i4", "i": "i4", "S1": "S1", "c": "S1"}
_RESTORE = {"u1": np.uint8, "uint8": np.uint8, "B": np.uint8, "i8": np.int64, "int64": np.int64, "u2": np.uint16, "u4": np.uint32}


class _Masked(np.ndarray):
    """ndarray with the ``.data`` / ``.filled`` / ``.mask`` trio the reference reads off netCDF4's masked arrays."""

    @property
    def data(self):                       # noqa: D401  (netCDF4: ds.variables[name][:].data)
        return np.asarray(self)

    @property
    def mask(self):
        return np.zeros(self.shape, dtype=bool)

    def filled(self, fill_value=None):
        return np.asarray(self)


class Dimension:
    def __init__(self, name, size):
        self.name, self.size = name, int(size)

    def __len__(self):
        return self.size


class Variable:
    def __init__(self, ds, name, raw, logical_dtype):
        object.__setattr__(self, "_ds", ds)
        object.__setattr__(self, "_name", name)
        object.__setattr__(self, "_raw", raw)
        object.__setattr__(self, "_logical", logical_dtype)

    # -- metadata
    @property
    def dimensions(self):
        return tuple(self._raw.dimensions)

    @property
    def shape(self):
        return tuple(self._raw.shape)

    @property
    def dtype(self):
        return np.dtype(self._logical) if self._logical is not None else self._raw.data.dtype.newbyteorder("=")

    def ncattrs(self):
        return [k for k in self._raw._attributes if not k.startswith("_qd_")]

    def getncattr(self, key):
        return _attr_out(self._raw._attributes[key])

    def setncattr(self, key, value):
        setattr(self._raw, key, _attr_in(value))

    def __getattr__(self, key):           # var.units, var.long_name ...
        raw = object.__getattribute__(self, "_raw")
        if key in raw._attributes:
            return _attr_out(raw._attributes[key])
        raise AttributeError(key)

    def __setattr__(self, key, value):
        setattr(self._raw, key, _attr_in(value))

    # -- data
    def __getitem__(self, idx):
        a = self._raw.data
        if self._raw.shape == ():
            out = np.array(a.item() if hasattr(a, "item") else a)
        else:
            out = np.array(a[idx])        # copy: the file is mmapped and closes with the dataset
        native = out.astype(out.dtype.newbyteorder("="), copy=False)
        if self._logical is not None:
            native = native.astype(self._logical)
        return native.view(_Masked)

    def __setitem__(self, idx, value):
        arr = np.asarray(value)
        if np.ma.isMaskedArray(value):
            arr = np.ma.filled(value)
        store = self._raw.data.dtype
        if self._logical is not None and np.dtype(self._logical).itemsize > store.itemsize and arr.size:
            info = np.iinfo(store)
            if arr.min() < info.min or arr.max() > info.max:
                raise OverflowError(f"variable {self._name!r}: values do not fit the NetCDF-3 storage type {store}")
        if self._raw.shape == ():
            self._raw.data[...] = arr.astype(store).reshape(())       # (netcdf_variable.assignValue indexes [:], which a 0-d array refuses)
        else:
            self._raw.data[idx] = arr.astype(store, copy=False)

    def __array__(self, dtype=None, copy=None):
        out = np.asarray(self[...])
        return out.astype(dtype) if dtype is not None else out

    def __len__(self):
        return self.shape[0]


def _attr_in(v):
    if isinstance(v, (bool, np.bool_)):
        return np.int32(int(v))
    if isinstance(v, str):
        return v
    if isinstance(v, (int, np.integer)):
        return np.int32(v) if -2 ** 31 <= int(v) < 2 ** 31 else np.float64(v)
    if isinstance(v, (float, np.floating)):
        return np.float64(v)
    return np.asarray(v)


def _attr_out(v):
    if isinstance(v, bytes):
        return v.decode("utf-8", "replace")
    if isinstance(v, np.ndarray) and v.shape in ((), (1,)):
        return v.reshape(()).item()
    return v


class Dataset:
    """``netCDF4.Dataset`` look-alike over a NetCDF-3 file (see the module docstring for the supported subset)."""

    def __init__(self, path, mode="r", format=None, **_ignored):
        self._path, self._mode = path, mode[0]
        if self._mode == "r":
            with open(path, "rb") as f:
                magic = f.read(4)
            if magic[:3] != b"CDF":
                raise OSError(f"{path!r} is not a NetCDF-3 classic file (magic {magic!r}); NetCDF-4/HDF5 files need the "
                              "real netCDF4 package, which this environment does not have")
        if self._mode in ("a", "r+"):
            raise NotImplementedError("append mode is not supported by the NetCDF-3 shim")
        self._f = netcdf_file(path, self._mode, mmap=False, version=2) if self._mode == "w" else netcdf_file(path, "r", mmap=False)
        self.dimensions = {k: Dimension(k, v if v is not None else 0) for k, v in self._f.dimensions.items()}
        self.variables = {}
        for name, raw in self._f.variables.items():
            logical = raw._attributes.get("_qd_dtype")
            logical = _RESTORE.get(logical.decode() if isinstance(logical, bytes) else logical) if logical is not None else None
            self.variables[name] = Variable(self, name, raw, logical)

    # -- structure
    def createDimension(self, name, size):
        self._f.createDimension(name, None if size is None else int(size))
        self.dimensions[name] = Dimension(name, size or 0)
        return self.dimensions[name]

    def createVariable(self, name, datatype, dimensions=(), **_ignored):      # zlib=, fill_value= ... are accepted and ignored
        key = np.dtype(datatype).name if not isinstance(datatype, str) else datatype
        if key not in _STORE:
            raise TypeError(f"unsupported NetCDF type {datatype!r}")
        raw = self._f.createVariable(name, _STORE[key], tuple(dimensions))
        logical = _RESTORE.get(key)
        if logical is not None:
            raw._qd_dtype = key
        v = Variable(self, name, raw, logical)
        self.variables[name] = v
        return v

    # -- attributes
    def setncattr(self, key, value):
        setattr(self._f, key, _attr_in(value))

    def getncattr(self, key):
        if key not in self._f._attributes:
            raise AttributeError(key)
        return _attr_out(self._f._attributes[key])

    def ncattrs(self):
        return list(self._f._attributes)

    def __getattr__(self, key):
        f = self.__dict__.get("_f")
        if f is not None and key in f._attributes:
            return _attr_out(f._attributes[key])
        raise AttributeError(key)

    def __getitem__(self, name):
        return self.variables[name]

    def __contains__(self, name):
        return name in self.variables

    # -- lifetime
    def sync(self):
        self._f.flush()

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def isopen(self):
        return self._f is not None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def install_netcdf4_shim(force=False):
    """Make ``import netCDF4`` / ``from netCDF4 import Dataset`` resolve to this module when the real package is
    absent (or always with force=True).  Returns True when the shim is what ``netCDF4`` now names."""
    if not force:
        try:
            import netCDF4  # noqa: F401
            return getattr(sys.modules["netCDF4"], "__qd_shim__", False)
        except Exception:
            pass
    mod = types.ModuleType("netCDF4")
    mod.Dataset, mod.Variable, mod.Dimension = Dataset, Variable, Dimension
    mod.__qd_shim__ = True
    mod.__version__ = "0-qd-netcdf3-shim"
    mod.__file__ = os.path.abspath(__file__)
    sys.modules["netCDF4"] = mod
    return True
