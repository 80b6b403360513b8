"""Oracle (test infrastructure only): EOF.fit / transform / inverse_transform on numpy arrays.

Reference lines followed (/root/reference/xeofs):
  single/base_model_single_set.py:123-161   fit = preprocessor.fit_transform -> _fit_algorithm
  single/eof.py:85-118                      total variance, Decomposer, scores = U*s,
                                            explained_variance = s^2/(n-1), result keys
  utils/xarray_utils.py:236-253             total_variance = X.var(sample, ddof=1).sum()
  single/eof.py:123-132                     transform = X . components
  single/eof.py:134-156                     inverse_transform = components . scores
  single/eof.py:221-240                     explained_variance_ratio
  preprocessing/sanitizer.py:128-153        components get NaN at dropped features
"""
from __future__ import annotations

import numpy as np

from . import preprocess as pp
from .decomposer import decompose


def eof_fit(
    X,
    dims,
    sample_dims,
    coords=None,
    n_modes=2,
    center=True,
    standardize=False,
    use_coslat=False,
    check_nans=True,
    weights=None,
    random_state=None,
    solver="auto",
    solver_kwargs=None,
):
    fitted = pp.preprocess(
        X, dims, sample_dims, coords=coords, center=center, standardize=standardize,
        use_coslat=use_coslat, weights=weights, check_nans=check_nans,
    )
    A = fitted["A"]
    total_variance = A.var(axis=0, ddof=1).sum()
    U, s, V = decompose(A, n_modes=n_modes, solver=solver, random_state=random_state,
                        solver_kwargs=solver_kwargs)
    n = A.shape[0]
    expvar = s**2 / (n - 1)
    vf = fitted["is_valid_feature"]
    comps_full = np.full((vf.size, V.shape[1]), np.nan, dtype=V.dtype)
    comps_full[vf] = V
    return {
        "A": A,
        "fitted": fitted,
        "components_2d": V,                      # (S', k) valid features only
        "components": comps_full.reshape(tuple(fitted["feature_shape"]) + (V.shape[1],)),
        "scores": U * s,                         # (n', k)
        "norms": s,
        "singular_values": s,
        "explained_variance": expvar,
        "total_variance": total_variance,
        "explained_variance_ratio": expvar / total_variance,
        "params": dict(center=center, standardize=standardize, use_coslat=use_coslat, check_nans=check_nans),
    }


def eof_transform(result, Xnew, dims):
    p = result["params"]
    A = pp.transform_new(Xnew, dims, result["fitted"], p["center"], p["standardize"], p["use_coslat"], p["check_nans"])
    return A @ result["components_2d"]


def eof_inverse_transform(result, scores):
    p = result["params"]
    rec = scores @ result["components_2d"].conj().T
    return pp.inverse_scale(rec, result["fitted"], p["center"], p["standardize"], p["use_coslat"])


def eof_fit_list(Xs, dims_list, sample_dims, coords_list=None, n_modes=2, center=True, standardize=False,
                 use_coslat=False, check_nans=True, weights=None, random_state=None, solver="auto", solver_kwargs=None):
    """EOF.fit on a LIST of arrays (DataList): every array goes through its own Scaler / Stacker / Sanitizer
    (preprocessing/preprocessor.py:208-228), the 2D matrices are concatenated along the feature axis
    (preprocessing/concatenator.py:58-81) and decomposed as one; the components are split back per array."""
    coords_list = coords_list or [None] * len(Xs)
    weights = weights or [None] * len(Xs)
    fits = [pp.preprocess(X, d, sample_dims, coords=c, center=center, standardize=standardize, use_coslat=use_coslat,
                          weights=w, check_nans=check_nans)
            for X, d, c, w in zip(Xs, dims_list, coords_list, weights)]
    A = np.concatenate([f["A"] for f in fits], axis=1)
    total_variance = A.var(axis=0, ddof=1).sum()
    U, s, V = decompose(A, n_modes=n_modes, solver=solver, random_state=random_state, solver_kwargs=solver_kwargs)
    n = A.shape[0]
    offs = np.concatenate([[0], np.cumsum([f["A"].shape[1] for f in fits])])
    return {
        "A": A, "fitted": fits,
        "components_2d": [V[a:b] for a, b in zip(offs[:-1], offs[1:])],
        "scores": U * s, "singular_values": s, "explained_variance": s**2 / (n - 1), "total_variance": total_variance,
    }
