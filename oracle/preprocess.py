"""Oracle (test infrastructure only): numpy restatement of the xeofs Preprocessor
arithmetic — Scaler -> Stacker -> Sanitizer — for a single DataArray-like input.

Reference lines followed (all under /root/reference/xeofs):
  preprocessing/preprocessor.py:208-228   order scaler -> renamer -> stacker -> sanitizer
  preprocessing/scaler.py:100-116         mean_ (nan-skipping), std_ ddof=0 clipped at float32 eps,
                                          coslat_weights_, weights_ (float64 ones when None)
  preprocessing/scaler.py:146-153         ((X - mean) / std) * coslat * weights, in this order
  utils/xarray_utils.py:78-87             feature_ones_like -> np.ones(dtype=float)  (=> fp64 promotion)
  utils/xarray_utils.py:256-270           sqrt(clip(cos(deg2rad(lat)), 0, 1))
  utils/xarray_utils.py:144-159           exactly one feature dim named like a latitude
  preprocessing/stacker.py:157-214        row-major stacking, output order (sample, feature)
  preprocessing/sanitizer.py:46-56        valid feature / valid sample / valid-per-sample counts
  preprocessing/sanitizer.py:108-124      isolated-NaN ValueError, drop all-NaN rows / cols
"""
from __future__ import annotations

import warnings

import numpy as np

VALID_LATITUDE_NAMES = [  # utils/constants.py:1-11
    "latitude", "lats", "lat", "Latitude", "Lats", "Lat", "LATITUDE", "LATS", "LAT",
]


def sqrt_cos_lat(lat):
    """utils/xarray_utils.py:256-270."""
    return np.sqrt(np.cos(np.deg2rad(lat)).clip(0, 1))


def split_dims(dims, sample_dims):
    """utils/xarray_utils.py:162-211: feature dims = every dim not named as sample dim, data order."""
    if isinstance(sample_dims, str):
        sample_dims = (sample_dims,)
    sample_dims = tuple(sample_dims)
    for d in sample_dims:
        if d not in dims:
            raise ValueError(f"dimension {d!r} not in {dims}")
    feature_dims = tuple(d for d in dims if d not in sample_dims)
    return sample_dims, feature_dims


def scaler_fit(X, dims, sample_dims, feature_dims, coords, center, standardize, use_coslat, weights):
    """preprocessing/scaler.py:69-126.  Returns feature-shaped arrays (dims order = feature_dims)."""
    sample_axes = tuple(dims.index(d) for d in sample_dims)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)  # all-NaN slices, as xarray silences them
        if center:
            out["mean"] = np.nanmean(X, axis=sample_axes)  # X.mean(sample_dims), skipna
        if standardize:
            std = np.nanstd(X, axis=sample_axes)  # ddof = 0
            out["std"] = np.clip(std, np.finfo(np.float32).eps, None)
    if use_coslat:
        lat_dims = set(feature_dims) & set(VALID_LATITUDE_NAMES)
        if len(lat_dims) == 0:
            raise ValueError("No latitude coordinate was found to compute coslat weights.")
        if len(lat_dims) > 1:
            raise ValueError(f"Found ambiguous latitude dimensions: {lat_dims}.")
        lat_dim = lat_dims.pop()
        lat = np.asarray(coords[lat_dim])
        w = sqrt_cos_lat(lat)
        shape = [1] * len(feature_dims)
        shape[feature_dims.index(lat_dim)] = lat.size
        out["coslat"] = w.reshape(shape)
    feat_shape = tuple(X.shape[dims.index(d)] for d in feature_dims)
    if weights is None:
        out["weights"] = np.ones(feat_shape, dtype=float)  # float64 -> promotes the product
    else:
        out["weights"] = np.asarray(weights)
    return out


def _to_sample_feature_order(X, dims, sample_dims, feature_dims):
    order = [dims.index(d) for d in sample_dims] + [dims.index(d) for d in feature_dims]
    return np.transpose(X, order)


def preprocess(
    X,
    dims,
    sample_dims,
    coords=None,
    center=True,
    standardize=False,
    use_coslat=False,
    weights=None,
    check_nans=True,
):
    """Full fit_transform of the preprocessor for one array.

    Returns dict with
      A                 (n', S') float64 (float32 only if weights were float32 and no coslat)
      is_valid_feature  (S,) bool, is_valid_sample (T,) bool
      scaler            dict of feature-shaped mean/std/coslat/weights
      sample_shape, feature_shape
    """
    X = np.asarray(X)
    dims = tuple(dims)
    coords = coords or {}
    sample_dims, feature_dims = split_dims(dims, sample_dims)
    sc = scaler_fit(X, dims, sample_dims, feature_dims, coords, center, standardize, use_coslat, weights)

    Xo = _to_sample_feature_order(X, dims, sample_dims, feature_dims)
    ns = len(sample_dims)
    # scaler.transform (scaler.py:146-153) — broadcasting over the leading sample axes
    if center:
        Xo = Xo - sc["mean"]
    if standardize:
        Xo = Xo / sc["std"]
    if use_coslat:
        Xo = Xo * sc["coslat"]
    Xo = Xo * sc["weights"]

    sample_shape = Xo.shape[:ns]
    feature_shape = Xo.shape[ns:]
    A = Xo.reshape(int(np.prod(sample_shape)), int(np.prod(feature_shape)))  # stacker.py:157-214

    notnull = ~np.isnan(A)
    is_valid_feature = notnull.any(axis=0)  # sanitizer.py:46-47
    is_valid_sample = notnull.any(axis=1)   # sanitizer.py:49-50
    if check_nans:
        per_sample = notnull.sum(axis=1)    # sanitizer.py:52-56
        ok = np.isin(per_sample, [0, is_valid_feature.sum()])
        if (~ok).any():  # sanitizer.py:115-122
            raise ValueError(
                "Input data contains partial NaN entries, which will cause the the SVD to fail."
            )
        A = A[is_valid_sample][:, is_valid_feature]  # sanitizer.py:124
    return {
        "A": A,
        "is_valid_feature": is_valid_feature,
        "is_valid_sample": is_valid_sample,
        "scaler": sc,
        "sample_shape": sample_shape,
        "feature_shape": feature_shape,
        "sample_dims": sample_dims,
        "feature_dims": feature_dims,
    }


def transform_new(Xnew, dims, fitted, center, standardize, use_coslat, check_nans=True):
    """Preprocessor.transform on unseen data with fitted scaler state
    (preprocessor.py:232-259; sanitizer.py:86-126 checks the NaN pattern)."""
    Xnew = np.asarray(Xnew)
    dims = tuple(dims)
    sc = fitted["scaler"]
    Xo = _to_sample_feature_order(Xnew, dims, fitted["sample_dims"], fitted["feature_dims"])
    if center:
        Xo = Xo - sc["mean"]
    if standardize:
        Xo = Xo / sc["std"]
    if use_coslat:
        Xo = Xo * sc["coslat"]
    Xo = Xo * sc["weights"]
    ns = len(fitted["sample_dims"])
    A = Xo.reshape(int(np.prod(Xo.shape[:ns])), -1)
    notnull = ~np.isnan(A)
    if check_nans:
        if not np.array_equal(notnull.any(axis=0), fitted["is_valid_feature"]):
            raise ValueError("Input data had NaN features in different locations than the original data.")
        per_sample = notnull.sum(axis=1)
        if (~np.isin(per_sample, [0, fitted["is_valid_feature"].sum()])).any():
            raise ValueError("Input data contains partial NaN entries, which will cause the the SVD to fail.")
        A = A[notnull.any(axis=1)][:, fitted["is_valid_feature"]]
    return A


def inverse_scale(A2d, fitted, center, standardize, use_coslat):
    """Sanitizer.inverse (reindex -> NaN at dropped features, sanitizer.py:128-153), unstack, then
    Scaler.inverse_transform_data (scaler.py:165-190): / weights / coslat * std + mean."""
    vf = fitted["is_valid_feature"]
    full = np.full((A2d.shape[0], vf.size), np.nan, dtype=A2d.dtype)
    full[:, vf] = A2d
    full = full.reshape((A2d.shape[0],) + tuple(fitted["feature_shape"]))
    sc = fitted["scaler"]
    full = full / sc["weights"]
    if use_coslat:
        full = full / sc["coslat"]
    if standardize:
        full = full * sc["std"]
    if center:
        full = full + sc["mean"]
    return full
