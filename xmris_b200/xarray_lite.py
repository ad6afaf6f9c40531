"""A small, dependency-free stand-in for the part of ``xarray`` the FID->spectrum path touches.

xarray is not installable in the build image nor on the GPU box (SURVEY.md finding 6), so the
drop-in accessor needs *some* labelled-array container to operate on.  This module provides
``DataArray`` / ``Variable`` with exactly the behaviour the hot path relies on (SURVEY.md
Appendix D): named dims, 1-D index coordinates with attrs, attrs/name propagation rules of
xarray binary ops (attrs dropped, name kept only when equal), ``pad`` / ``roll`` / ``isel`` /
``rename`` / ``transpose`` / ``copy(data=)`` / ``assign_coords`` / ``assign_attrs``, name-based
broadcasting for arithmetic and numpy ufuncs, and the accessor registry.

It has two users:
  * the product: ``xmris_b200`` operates on these objects when real ``xarray`` is absent
    (when ``xarray`` imports, the accessor is registered there as well);
  * the golden-vector generator (``tests/golden/make_golden.py``), which installs this module as
    ``sys.modules["xarray"]`` and executes the reference's own five hot-path files on top of it.
    Numeric results there depend only on numpy/scipy, not on this stand-in.

It is intentionally not a general xarray replacement: no Dataset, no multi-dimensional
coordinates, no alignment by label (index coordinates of operands must match positionally).
"""

from __future__ import annotations

import warnings
from collections import OrderedDict

import numpy as np

__all__ = [
    "DataArray",
    "Variable",
    "register_dataarray_accessor",
    "register_dataset_accessor",
    "AccessorRegistrationWarning",
]

__version__ = "0-lite"


class AccessorRegistrationWarning(Warning):
    """Raised (as a warning) when an accessor name is registered twice."""


class Variable:
    """``Variable(dims, data, attrs=None)`` -- an array with named dims and attrs, no coords."""

    def __init__(self, dims, data, attrs=None):
        self.dims = (dims,) if isinstance(dims, str) else tuple(dims)
        self._data = np.asarray(data)
        if self._data.ndim != len(self.dims):
            raise ValueError(
                f"dimensions {self.dims} must have the same length as the number of data dimensions, "
                f"ndim={self._data.ndim}"
            )
        self.attrs = dict(attrs) if attrs else {}

    @property
    def values(self):
        return self._data

    @property
    def shape(self):
        return self._data.shape

    def copy(self):
        return Variable(self.dims, self._data.copy(), dict(self.attrs))


class _Coordinates:
    """Mapping view over a DataArray's coordinates; items come back as DataArrays (as xarray does)."""

    def __init__(self, owner: "DataArray"):
        self._owner = owner

    def __contains__(self, key):
        return key in self._owner._coords

    def __iter__(self):
        return iter(self._owner._coords)

    def __len__(self):
        return len(self._owner._coords)

    def keys(self):
        return self._owner._coords.keys()

    def items(self):
        return [(k, self[k]) for k in self._owner._coords]

    def __getitem__(self, key):
        var = self._owner._coords[key]  # KeyError for a bare dimension, like xarray (Appendix B Q11)
        # A coordinate DataArray carries itself as its own index coordinate.
        co = OrderedDict()
        if var.dims == (key,):
            co[key] = var
        return DataArray._construct(var._data, var.dims, co, dict(var.attrs), key)


class _Sizes(dict):
    pass


def _as_coord_variable(name, value, dims_hint=None):
    """Normalise the accepted coordinate spellings to a Variable."""
    if isinstance(value, Variable):
        return Variable(value.dims, value._data, value.attrs)
    if isinstance(value, DataArray):
        return Variable(value.dims, value._data, value.attrs)
    if isinstance(value, tuple) and len(value) >= 2 and isinstance(value[0], (str, tuple, list)):
        dims, data = value[0], value[1]
        attrs = value[2] if len(value) > 2 else None
        return Variable(dims, data, attrs)
    arr = np.asarray(value)
    if arr.ndim == 0:
        return Variable((), arr)
    if arr.ndim == 1:
        return Variable((name,), arr)
    raise ValueError(f"coordinate {name!r}: cannot infer dims for a {arr.ndim}-D array")


class DataArray:
    """Minimal labelled N-D array: ``DataArray(data, coords=None, dims=None, name=None, attrs=None)``."""

    __array_priority__ = 60

    def __init__(self, data, coords=None, dims=None, name=None, attrs=None):
        data = np.asarray(data)
        if dims is None:
            dims = tuple(f"dim_{i}" for i in range(data.ndim))
        dims = (dims,) if isinstance(dims, str) else tuple(dims)
        if len(dims) != data.ndim:
            raise ValueError(f"different number of dimensions on data and dims: {data.ndim} vs {len(dims)}")
        co = OrderedDict()
        if coords is not None:
            items = coords.items() if hasattr(coords, "items") else coords
            for k, v in items:
                var = _as_coord_variable(k, v)
                for d, n in zip(var.dims, var.shape):
                    if d not in dims:
                        raise ValueError(f"coordinate {k!r} has dimension {d!r} not on the data {dims}")
                    if data.shape[dims.index(d)] != n:
                        raise ValueError(
                            f"conflicting sizes for dimension {d!r}: length {data.shape[dims.index(d)]} on the "
                            f"data but length {n} on coordinate {k!r}"
                        )
                co[k] = var
        self._data = data
        self._dims = dims
        self._coords = co
        self._attrs = dict(attrs) if attrs else {}
        self.name = name

    # -- construction helpers -------------------------------------------------------------
    @classmethod
    def _construct(cls, data, dims, coords, attrs, name):
        obj = cls.__new__(cls)
        obj._data = np.asarray(data)
        obj._dims = tuple(dims)
        obj._coords = coords
        obj._attrs = attrs
        obj.name = name
        return obj

    def _replace(self, data=None, dims=None, coords=None, attrs=None, name="__keep__"):
        return DataArray._construct(
            self._data if data is None else data,
            self._dims if dims is None else dims,
            OrderedDict((k, v.copy()) for k, v in self._coords.items()) if coords is None else coords,
            dict(self._attrs) if attrs is None else attrs,
            self.name if name == "__keep__" else name,
        )

    # -- basic properties -----------------------------------------------------------------
    @property
    def dims(self):
        return self._dims

    @property
    def values(self):
        return self._data

    @values.setter
    def values(self, v):
        v = np.asarray(v)
        if v.shape != self._data.shape:
            raise ValueError("replacement data must match the existing shape")
        self._data = v

    @property
    def data(self):
        return self._data

    @property
    def shape(self):
        return self._data.shape

    @property
    def ndim(self):
        return self._data.ndim

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def size(self):
        return self._data.size

    @property
    def sizes(self):
        return _Sizes(zip(self._dims, self._data.shape))

    @property
    def coords(self):
        return _Coordinates(self)

    @property
    def attrs(self):
        return self._attrs

    @attrs.setter
    def attrs(self, value):
        self._attrs = dict(value)

    def __len__(self):
        return self._data.shape[0]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def __float__(self):
        return float(self._data)

    def __int__(self):
        return int(self._data)

    def __complex__(self):
        return complex(self._data)

    def item(self):
        return self._data.item()

    def __repr__(self):
        co = ", ".join(f"{k}{v.dims}" for k, v in self._coords.items())
        return (
            f"<xarray_lite.DataArray {self.name or ''} {dict(self.sizes)} dtype={self.dtype} "
            f"coords=[{co}] attrs={list(self._attrs)}>"
        )

    def get_axis_num(self, dim):
        if isinstance(dim, (list, tuple)):
            return tuple(self.get_axis_num(d) for d in dim)
        try:
            return self._dims.index(dim)
        except ValueError:
            raise ValueError(f"{dim!r} not found in array dimensions {self._dims!r}") from None

    # -- copying / metadata ---------------------------------------------------------------
    def copy(self, deep=True, data=None):
        if data is not None:
            data = np.asarray(data)
            if data.shape != self._data.shape:
                raise ValueError(f"Data shape {data.shape} must match shape of object {self._data.shape}")
            return self._replace(data=data)
        return self._replace(data=self._data.copy() if deep else self._data)

    def assign_attrs(self, *args, **kwargs):
        new = dict(self._attrs)
        for a in args:
            new.update(a)
        new.update(kwargs)
        return self._replace(attrs=new)

    def assign_coords(self, coords=None, **kwargs):
        items = dict(coords or {})
        items.update(kwargs)
        co = OrderedDict((k, v.copy()) for k, v in self._coords.items())
        for k, v in items.items():
            var = _as_coord_variable(k, v)
            for d, n in zip(var.dims, var.shape):
                if d not in self._dims:
                    raise ValueError(f"cannot add coordinate {k!r} with new dimension {d!r}")
                if self.sizes[d] != n:
                    raise ValueError(
                        f"conflicting sizes for dimension {d!r}: length {self.sizes[d]} on the data but "
                        f"length {n} on coordinate {k!r}"
                    )
            co[k] = var
        return self._replace(coords=co)

    def rename(self, new_name_or_name_dict=None, **names):
        if new_name_or_name_dict is None or isinstance(new_name_or_name_dict, dict):
            mapping = dict(new_name_or_name_dict or {})
            mapping.update(names)
            for k in mapping:
                if k not in self._dims and k not in self._coords:
                    raise ValueError(f"cannot rename {k!r} because it is not a dimension or coordinate")
            dims = tuple(mapping.get(d, d) for d in self._dims)
            co = OrderedDict()
            for k, v in self._coords.items():
                co[mapping.get(k, k)] = Variable(tuple(mapping.get(d, d) for d in v.dims), v._data, v.attrs)
            return self._replace(dims=dims, coords=co)
        return self._replace(name=new_name_or_name_dict)

    def swap_dims(self, dims_dict=None, **kw):
        """Make another 1-D coordinate the dimension (index) coordinate; the old one stays as a non-index coord."""
        mapping = dict(dims_dict or {})
        mapping.update(kw)
        for old, new in mapping.items():
            if old not in self._dims:
                raise ValueError(f"cannot swap from dimension {old!r} because it is not one of the dimensions {self._dims}")
            if new in self._coords and self._coords[new].dims != (old,):
                raise ValueError(f"replacement dimension {new!r} is not a 1D variable along the old dimension {old!r}")
        dims = tuple(mapping.get(d, d) for d in self._dims)
        co = OrderedDict()
        for k, v in self._coords.items():
            co[k] = Variable(tuple(mapping.get(d, d) for d in v.dims), v._data, v.attrs)
        return self._replace(dims=dims, coords=co)

    def pipe(self, func, *args, **kwargs):
        return func(self, *args, **kwargs)

    # -- shape manipulation ---------------------------------------------------------------
    def transpose(self, *dims):
        if not dims:
            dims = self._dims[::-1]
        if ... in dims:
            rest = [d for d in self._dims if d not in dims]
            i = dims.index(...)
            dims = tuple(dims[:i]) + tuple(rest) + tuple(dims[i + 1 :])
        if set(dims) != set(self._dims) or len(dims) != len(self._dims):
            raise ValueError(f"{dims} must be a permuted list of {self._dims}")
        order = [self._dims.index(d) for d in dims]
        return self._replace(data=np.transpose(self._data, order), dims=tuple(dims))

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {})
        indexers.update(kw)
        for d in indexers:
            if d not in self._dims:
                raise ValueError(f"Dimensions {{{d!r}}} do not exist. Expected one or more of {self._dims}")
        key = tuple(indexers.get(d, slice(None)) for d in self._dims)
        data = self._data[key]
        new_dims = tuple(
            d for d in self._dims if not (d in indexers and np.ndim(indexers[d]) == 0 and not isinstance(indexers[d], slice))
        )
        co = OrderedDict()
        for k, v in self._coords.items():
            ckey = tuple(indexers.get(d, slice(None)) for d in v.dims)
            cdims = tuple(
                d for d in v.dims if not (d in indexers and np.ndim(indexers[d]) == 0 and not isinstance(indexers[d], slice))
            )
            co[k] = Variable(cdims, v._data[ckey], v.attrs)  # scalar coords are kept, like xarray
        return DataArray._construct(data, new_dims, co, dict(self._attrs), self.name)

    def pad(self, pad_width=None, mode="constant", constant_values=None, **kw):
        pad_width = dict(pad_width or {})
        pad_width.update(kw)
        if mode != "constant":
            raise NotImplementedError("xarray_lite.pad only implements mode='constant'")
        np_pad = []
        for d in self._dims:
            w = pad_width.get(d, (0, 0))
            if isinstance(w, int):
                w = (w, w)
            np_pad.append(tuple(w))
        for d in pad_width:
            if d not in self._dims:
                raise ValueError(f"cannot pad along missing dimension {d!r}")
        cv = 0 if constant_values is None else constant_values
        data = np.pad(self._data, np_pad, mode="constant", constant_values=cv)
        co = OrderedDict()
        for k, v in self._coords.items():
            if any(d in pad_width for d in v.dims):
                # index coordinates along a padded dim are filled with NaN
                cpad = []
                for d in v.dims:
                    w = pad_width.get(d, (0, 0))
                    cpad.append((w, w) if isinstance(w, int) else tuple(w))
                cdata = np.pad(v._data.astype(np.result_type(v._data.dtype, np.float64)), cpad,
                               mode="constant", constant_values=np.nan)
                co[k] = Variable(v.dims, cdata, v.attrs)
            else:
                co[k] = v.copy()
        return DataArray._construct(data, self._dims, co, dict(self._attrs), self.name)

    def roll(self, shifts=None, roll_coords=False, **kw):
        shifts = dict(shifts or {})
        shifts.update(kw)
        data = self._data
        for d, s in shifts.items():
            data = np.roll(data, s, axis=self.get_axis_num(d))
        co = OrderedDict()
        for k, v in self._coords.items():
            cdata = v._data
            if roll_coords:
                for d, s in shifts.items():
                    if d in v.dims:
                        cdata = np.roll(cdata, s, axis=v.dims.index(d))
            co[k] = Variable(v.dims, cdata, v.attrs)
        return DataArray._construct(data, self._dims, co, dict(self._attrs), self.name)

    # -- reductions -----------------------------------------------------------------------
    def _reduce(self, fn, dim=None):
        if dim is None:
            return DataArray._construct(fn(self._data), (), OrderedDict(), {}, self.name)
        dims = [dim] if isinstance(dim, str) else list(dim)
        axes = tuple(self.get_axis_num(d) for d in dims)
        data = fn(self._data, axis=axes)
        new_dims = tuple(d for d in self._dims if d not in dims)
        co = OrderedDict((k, v.copy()) for k, v in self._coords.items() if not any(d in dims for d in v.dims))
        return DataArray._construct(data, new_dims, co, {}, self.name)

    def min(self, dim=None):
        return self._reduce(np.min, dim)

    def max(self, dim=None):
        return self._reduce(np.max, dim)

    def sum(self, dim=None):
        return self._reduce(np.sum, dim)

    def mean(self, dim=None):
        return self._reduce(np.mean, dim)

    @property
    def real(self):
        return self._replace(data=self._data.real)

    @property
    def imag(self):
        return self._replace(data=self._data.imag)

    def astype(self, dtype):
        return self._replace(data=self._data.astype(dtype))

    def conj(self):
        return self._replace(data=np.conj(self._data))

    # -- arithmetic with broadcasting by dimension name -------------------------------------
    @staticmethod
    def _broadcast(operands):
        """Return (arrays aligned to a common dim order, dims, merged coords, result name)."""
        dims = []
        for op in operands:
            if isinstance(op, DataArray):
                for d in op._dims:
                    if d not in dims:
                        dims.append(d)
        arrays = []
        coords = OrderedDict()
        names = []
        for op in operands:
            if isinstance(op, DataArray):
                names.append(op.name)
                order = [op._dims.index(d) for d in dims if d in op._dims]
                arr = np.transpose(op._data, order)
                shape = [op._data.shape[op._dims.index(d)] if d in op._dims else 1 for d in dims]
                arrays.append(arr.reshape(shape))
                for k, v in op._coords.items():
                    if k in coords:
                        if v.dims == (k,) and coords[k]._data.shape == v._data.shape and not np.array_equal(
                            coords[k]._data, v._data, equal_nan=True
                        ):
                            raise ValueError(
                                f"xarray_lite does not align by label: index coordinate {k!r} differs between operands"
                            )
                    else:
                        coords[k] = v.copy()
            elif isinstance(op, Variable):
                raise TypeError("arithmetic between DataArray and bare Variable is not supported in xarray_lite")
            else:
                arr = np.asarray(op)
                if arr.ndim not in (0,):
                    # plain ndarray operands broadcast positionally against the first DataArray
                    arrays.append(arr)
                else:
                    arrays.append(arr)
        # xarray keeps the name only when all DataArray operands agree on it
        name = names[0] if names and all(n == names[0] for n in names) else None
        return arrays, tuple(dims), coords, name

    def _binary(self, other, fn, reflexive=False):
        if isinstance(other, Variable):
            return NotImplemented
        ops = (other, self) if reflexive else (self, other)
        arrays, dims, coords, name = DataArray._broadcast(ops)
        data = fn(arrays[0], arrays[1])
        return DataArray._construct(data, dims, coords, {}, name)  # attrs are dropped (keep_attrs=False)

    def __add__(self, o):
        return self._binary(o, np.add)

    def __radd__(self, o):
        return self._binary(o, np.add, True)

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __rsub__(self, o):
        return self._binary(o, np.subtract, True)

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    def __rmul__(self, o):
        return self._binary(o, np.multiply, True)

    def __truediv__(self, o):
        return self._binary(o, np.true_divide)

    def __rtruediv__(self, o):
        return self._binary(o, np.true_divide, True)

    def __pow__(self, o):
        return self._binary(o, np.power)

    def __rpow__(self, o):
        return self._binary(o, np.power, True)

    def __neg__(self):
        return self._replace(data=-self._data, attrs={})

    def __abs__(self):
        return self._replace(data=np.abs(self._data), attrs={})

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        if any(isinstance(x, Variable) for x in inputs):
            return NotImplemented
        arrays, dims, coords, name = DataArray._broadcast(inputs)
        result = ufunc(*arrays, **kwargs)
        if isinstance(result, tuple):
            return tuple(DataArray._construct(r, dims, coords, {}, name) for r in result)
        return DataArray._construct(result, dims, coords, {}, name)

    # -- accessor registry --------------------------------------------------------------------
    _accessors: dict = {}

    def __getattr__(self, item):
        # only reached when normal lookup fails -> accessor namespaces
        registry = type(self)._accessors
        if item in registry and not item.startswith("_"):
            cache = self.__dict__.setdefault("_accessor_cache", {})
            if item not in cache:
                cache[item] = registry[item](self)
            return cache[item]
        raise AttributeError(f"{type(self).__name__!r} object has no attribute {item!r}")


def register_dataarray_accessor(name):
    """Class decorator mirroring ``xarray.register_dataarray_accessor``."""

    def decorator(accessor):
        if name in DataArray._accessors or hasattr(DataArray, name):
            warnings.warn(
                f"registration of accessor {accessor!r} under name {name!r} for type {DataArray!r} is "
                "overriding a preexisting attribute with the same name.",
                AccessorRegistrationWarning,
                stacklevel=2,
            )
        DataArray._accessors[name] = accessor
        return accessor

    return decorator


class Dataset:  # placeholder so that `xr.Dataset` annotations in reference files resolve
    _accessors: dict = {}


def register_dataset_accessor(name):
    def decorator(accessor):
        Dataset._accessors[name] = accessor
        return accessor

    return decorator


def apply_ufunc(func, *args, kwargs=None, input_core_dims=None, output_core_dims=((),), vectorize=False, dask="forbidden",
                **_ignored):
    """The one calling pattern the reference uses (``processing/baseline.py:88-96``): a single DataArray argument, one
    input and one output core dimension, ``vectorize=True`` -- ``func`` is applied to every 1-D slice along the core
    dimension and, as in xarray, the core dimension ends up LAST in the result."""
    if len(args) != 1 or not isinstance(args[0], DataArray):
        raise NotImplementedError("xarray_lite.apply_ufunc supports a single DataArray argument")
    da = args[0]
    in_core = list((input_core_dims or [[]])[0])
    out_core = list(list(output_core_dims)[0]) if output_core_dims else []
    if len(in_core) != 1 or out_core != in_core or not vectorize:
        raise NotImplementedError("xarray_lite.apply_ufunc supports one shared input/output core dimension with vectorize=True")
    dim = in_core[0]
    axis = da.get_axis_num(dim)
    moved = np.moveaxis(np.asarray(da.values), axis, -1)
    flat = moved.reshape(-1, moved.shape[-1])
    out = np.stack([np.asarray(func(row, **(kwargs or {}))) for row in flat]) if len(flat) else flat.copy()
    out = out.reshape(moved.shape[:-1] + (out.shape[-1],))
    dims = tuple(d for d in da.dims if d != dim) + (dim,)
    coords = {k: da.coords[k] for k in da.coords}
    return DataArray(out, dims=dims, coords=coords, name=da.name)
