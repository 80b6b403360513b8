"""Oracle (test infrastructure only): restatement of xeofs.linalg.decomposer.Decomposer.fit on a
plain 2D numpy array.

Reference lines followed (/root/reference/xeofs):
  linalg/decomposer.py:86-100    rank, n_modes_precompute (float n_modes -> int(rank*init_rank_reduction)),
                                 ValueError when n_modes > rank
  linalg/decomposer.py:112-131   solver policy: "auto" -> exact iff max(shape) < 500 and k > int(0.8*rank)
  linalg/decomposer.py:134-138   exact: np.linalg.svd(X, **solver_kwargs) truncated to k
  linalg/decomposer.py:141-146   randomized: sklearn.utils.extmath.randomized_svd(X, **(solver_kwargs |
                                 {n_components, random_state}))   [third party, called, not restated]
  linalg/decomposer.py:188-216   variance-fraction truncation (float n_modes)
  linalg/decomposer.py:219-222   sign rule  (utils/xarray_utils.py:273-301; numpy twin at
                                 linalg/_numpy/_svd.py:13-32: +1 iff |max| >= |min| along the feature axis)
  linalg/decomposer.py:224-226   U_, s_, V_ = VT^H
The numpy twin of the whole class, linalg/_numpy/_svd.py:35-243, is what tests/golden/make_golden.py
executes to pin this file.
"""
from __future__ import annotations

import warnings

import numpy as np
from sklearn.utils.extmath import randomized_svd


def sign_multiplier(VT):
    """+1 where |max| >= |min| over the feature axis (axis=1 of VT), else -1."""
    mx = VT.max(axis=1)
    mn = VT.min(axis=1)
    return np.where(np.abs(mx) >= np.abs(mn), 1, -1)


def decompose(
    X,
    n_modes=2,
    init_rank_reduction=0.3,
    flip_signs=True,
    solver="auto",
    random_state=None,
    solver_kwargs=None,
):
    """Returns U (n, k), s (k,), V (S, k)."""
    solver_kwargs = dict(solver_kwargs or {})
    rank = min(X.shape)
    based_on_variance = isinstance(n_modes, float)
    k = n_modes
    if based_on_variance:
        k = int(rank * init_rank_reduction)
        if k < 1:
            warnings.warn("`init_rank_reduction` is too low resulting in zero components.")
            k = 1
    if k > rank:
        raise ValueError(
            f"n_modes must be less than or equal to the rank of the dataset (rank = {rank})."
        )
    is_small = max(X.shape) < 500
    if solver == "auto":
        use_exact = bool(is_small and k > int(0.8 * rank))
    elif solver == "full":
        use_exact = True
    elif solver == "randomized":
        use_exact = False
    else:
        raise ValueError(f"Unrecognized solver '{solver}'. Valid options are 'auto', 'full', and 'randomized'.")

    if use_exact:
        U, s, VT = np.linalg.svd(X, **solver_kwargs)
        U, s, VT = U[:, :k], s[:k], VT[:k, :]
    else:
        kw = solver_kwargs | {"n_components": k, "random_state": random_state}
        U, s, VT = randomized_svd(X, **kw)

    if based_on_variance:
        N = X.shape[0] - 1
        total_variance = X.var(axis=0, ddof=1).sum()
        expvar = s**2 / N / total_variance
        cum = expvar.cumsum()
        n_req = k - int((cum >= n_modes).sum()) + 1
        if n_req > k:
            warnings.warn("requested explained variance not reached; consider increasing `init_rank_reduction`.")
            n_req = k
        U, s, VT = U[:, :n_req], s[:n_req], VT[:n_req, :]

    if flip_signs:
        sgn = sign_multiplier(VT)
        VT = VT * sgn[:, None]
        U = U * sgn[None, :]
    return U, s, VT.conj().T
