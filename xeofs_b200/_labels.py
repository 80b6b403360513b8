"""Minimal labelled-array shim (xarray is not installed in the target image).

``DataArray(data, dims, coords)`` carries what the hot path needs from an ``xarray.DataArray``: the dim
names (to split sample from feature dims, utils/xarray_utils.py:162-211), the latitude coordinate (coslat
weights, utils/xarray_utils.py:144-159, 256-270) and the shapes to unstack results (stacker.py:216-275).
Real ``xarray.DataArray`` objects are accepted by duck typing (``.dims``, ``.coords``, ``.values``) and results
are returned as ``xarray.DataArray`` when xarray is importable and the input was one.
"""
from __future__ import annotations

import numpy as np
import torch

VALID_LATITUDE_NAMES = [  # utils/constants.py:1-11
    "latitude", "lats", "lat", "Latitude", "Lats", "Lat", "LATITUDE", "LATS", "LAT",
]


class DataArray:
    def __init__(self, data, dims, coords=None, name=None, attrs=None):
        self.data = data
        self.dims = tuple(dims)
        if len(self.dims) != data.ndim:
            raise ValueError(f"dims {self.dims} do not match data of rank {data.ndim}")
        self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
        self.name = name
        self.attrs = dict(attrs or {})

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def values(self):
        d = self.data
        return d.detach().cpu().numpy() if isinstance(d, torch.Tensor) else np.asarray(d)

    def __repr__(self):
        return f"<xeofs_b200.DataArray {self.name or ''} {dict(zip(self.dims, self.shape))}>"


def _is_xarray(X):
    return type(X).__module__.startswith("xarray") and hasattr(X, "dims") and hasattr(X, "values")


def validate_input_type(X):
    """utils/sanity_checks.py:85-95."""
    if not (isinstance(X, DataArray) or _is_xarray(X)):
        raise TypeError(
            f"Invalid input type: {type(X).__name__}. Expected one of the following: DataArray "
            "(xeofs_b200.DataArray or xarray.DataArray)."
        )


def unpack(X):
    """-> (data, dims, coords, was_xarray)."""
    validate_input_type(X)
    if isinstance(X, DataArray):
        return X.data, X.dims, X.coords, False
    coords = {str(k): np.asarray(v.values) for k, v in X.coords.items() if getattr(v, "ndim", 1) == 1}
    return X.values, tuple(str(d) for d in X.dims), coords, True


def wrap(data, dims, coords, name, as_xarray=False, attrs=None):
    coords = {k: v for k, v in (coords or {}).items() if k in dims and len(v) == data.shape[dims.index(k)]}
    if as_xarray:
        import xarray as xr

        return xr.DataArray(np.asarray(data), dims=dims, coords=coords, name=name, attrs=attrs or {})
    return DataArray(data, dims, coords, name=name, attrs=attrs)


def split_dims(dims, sample_dims):
    """utils/xarray_utils.py:162-211: feature dims = every dim not named as a sample dim, in data order."""
    if isinstance(sample_dims, str):
        sample_dims = (sample_dims,)
    sample_dims = tuple(sample_dims)
    for d in sample_dims:
        if d not in dims:
            raise ValueError(f"Sample dimension {d!r} not found in the data dimensions {dims}.")
    feature_dims = tuple(d for d in dims if d not in sample_dims)
    if not feature_dims:
        raise ValueError("No feature dimension left after removing the sample dimensions.")
    return sample_dims, feature_dims


def sqrt_cos_lat_vector(feature_dims, coords):
    """utils/xarray_utils.py:103-159, 256-270: sqrt(clip(cos(deg2rad(lat)), 0, 1)) along the ONE feature dim named
    like a latitude (float64), and the shape that broadcasts it over the feature dims."""
    lat_dims = [d for d in feature_dims if d in VALID_LATITUDE_NAMES]
    if len(lat_dims) == 0:
        raise ValueError(
            f"No latitude coordinate was found to compute coslat weights. Must be one of the following: {VALID_LATITUDE_NAMES}"
        )
    if len(lat_dims) > 1:
        raise ValueError(
            f"Found ambiguous latitude dimensions: {lat_dims}. Only ONE of the following is allowed for computing coslat weights: {VALID_LATITUDE_NAMES}"
        )
    lat_dim = lat_dims[0]
    if lat_dim not in coords:
        raise ValueError(f"No coordinate values for latitude dimension {lat_dim!r}.")
    lat = np.asarray(coords[lat_dim], dtype=np.float64)
    w = np.sqrt(np.cos(np.deg2rad(lat)).clip(0, 1))
    shape = [1] * len(feature_dims)
    shape[feature_dims.index(lat_dim)] = lat.size
    return w, shape


def sqrt_cos_lat_weights(feature_dims, feature_shape, coords):
    """The same, broadcast over the feature shape on the host (a view)."""
    w, shape = sqrt_cos_lat_vector(feature_dims, coords)
    return np.broadcast_to(w.reshape(shape), feature_shape)


def mode_indices(coords, n, k):
    """Row indices (0-based) of the modes a score array names with its 'mode' coordinate (1-based, default 1..n).
    The reference selects them with ``.sel(mode=...)`` (single/eof.py:150-152), which raises KeyError for a mode
    the model does not hold; the device kernel indexes the components buffer with them, so they are checked here."""
    modes = np.asarray(coords.get("mode", np.arange(1, n + 1))).astype(int) - 1
    if modes.shape != (n,):
        raise ValueError(f"the 'mode' coordinate has {modes.size} entries, the scores have {n} modes")
    bad = modes[(modes < 0) | (modes >= k)]
    if bad.size:
        raise KeyError(f"modes {sorted(set((bad + 1).tolist()))} are not in the model (modes 1..{k})")
    return modes
