"""Oracle (test infrastructure only): MCA.fit (= CPCCA with alpha=1) on numpy arrays, use_pca=False
or the default PCA pre-projection.

Reference lines followed (/root/reference/xeofs):
  cross/mca.py:104-123                      MCA = CPCCA(alpha=1.0)
  cross/cpcca.py:143-146                    center=True hard-coded
  cross/base_model_cross_set.py:304-315     preprocess -> PCA -> (augment) -> whiten -> _fit_algorithm
  preprocessing/whitener.py:96-98           alpha == 1 -> identity
  preprocessing/pca.py:94-131               optional PCA: rSVD with int(rank*0.3) modes, 99.9 % truncation, X <- X V
  cross/cpcca.py:176-184, 1008-1015         C = X^H Y / (n - 1)
  cross/cpcca.py:187-194                    Decomposer.fit(C) -> s, Q1 = U_, Q2 = V_
  cross/cpcca.py:197, 991-1000              total_squared_covariance = sum |C|^2
  cross/cpcca.py:200                        idx_modes_sorted = argsort(s)[::-1]
  cross/cpcca.py:204-208                    scores = X Q, norm = sqrt(diag(scores^H scores))
"""
from __future__ import annotations

import numpy as np

from . import preprocess as pp
from .decomposer import decompose


def cross_covariance(X, Y):
    """cross/cpcca.py:1008-1015 (assumes centred data)."""
    if X.shape[0] != Y.shape[0]:
        raise ValueError(
            f"Both data matrices must have the same number of samples but found {X.shape[0]} in the first and {Y.shape[0]} in the second."
        )
    return X.conj().T @ Y / (X.shape[0] - 1)


def whitener_transform(X, alpha):
    """preprocessing/whitener.py:111-133 with linalg/_numpy/_utils.py:6-33: T = C^((alpha-1)/2) for C = X^H X / n
    through the SVD of C (singular values <= eps cut), Tinv = inv(T) (pinv when singular)."""
    C = X.conj().T @ X / X.shape[0]
    _, sv, Vh = np.linalg.svd(C)
    keep = sv > np.finfo(sv.dtype).eps
    Vk, sk = Vh[keep].conj().T, sv[keep]
    T = (Vk * sk ** ((alpha - 1.0) / 2.0)) @ Vk.conj().T
    T = T if np.iscomplexobj(C) else T.real
    try:
        Tinv = np.linalg.inv(T)
    except np.linalg.LinAlgError:
        Tinv = np.linalg.pinv(T)
    return T, Tinv


def residual_squared_covariance(X, Y, Xrec, Yrec):
    """cross/cpcca.py:436-443: squared Frobenius norm of the cross-covariance of the residuals."""
    dX, dY = X - Xrec, Y - Yrec
    return np.linalg.norm(dX.conj().T @ dY / (dX.shape[0] - 1)) ** 2


def pearson_correlation(X, Y):
    """utils/optional/statistics.py:51-55: columns scaled by their population standard deviation, X^H Y / n
    (centred data assumed)."""
    return (X / X.std(0)).conj().T @ (Y / Y.std(0)) / X.shape[0]


def mca_fit(
    X, Y, dims_x, dims_y, sample_dims,
    coords_x=None, coords_y=None,
    n_modes=2, standardize=False, use_coslat=False, check_nans=True,
    weights_x=None, weights_y=None,
    random_state=None, solver="auto", solver_kwargs=None,
    use_pca=False, n_pca_modes=0.999, pca_init_rank_reduction=0.3, pca_random_state=None, alpha=1.0,
):
    """use_pca=False path (the configuration BASELINE.json config 3 is built on; the default
    use_pca=True path is unseeded in the reference, cross/base_model_cross_set.py:165-179)."""
    def _pair(v):
        return (v, v) if not isinstance(v, (list, tuple)) else tuple(v)
    std, cos, chk = _pair(standardize), _pair(use_coslat), _pair(check_nans)
    f1 = pp.preprocess(X, dims_x, sample_dims, coords=coords_x, center=True, standardize=std[0],
                       use_coslat=cos[0], weights=weights_x, check_nans=chk[0])
    f2 = pp.preprocess(Y, dims_y, sample_dims, coords=coords_y, center=True, standardize=std[1],
                       use_coslat=cos[1], weights=weights_y, check_nans=chk[1])
    A1, A2 = f1["A"], f2["A"]
    V1 = V2 = None
    if use_pca:
        # cross/base_model_cross_set.py:165-179, 307-308; preprocessing/pca.py:94-131: the fields are replaced by their
        # projections on the leading principal components (SVD class of linalg/_numpy/_svd.py = decompose(); the
        # reference leaves its random_state unset, pca_random_state makes this restatement reproducible)
        _, _, V1 = decompose(A1, n_modes=n_pca_modes, init_rank_reduction=pca_init_rank_reduction,
                             random_state=pca_random_state)
        _, _, V2 = decompose(A2, n_modes=n_pca_modes, init_rank_reduction=pca_init_rank_reduction,
                             random_state=pca_random_state)
        A1, A2 = A1 @ V1, A2 @ V2
    # fractional whitening (preprocessing/whitener.py:111-133; linalg/_numpy/_utils.py:6-33): T = C^((alpha-1)/2) with
    # C = X^T X / n_samples through the SVD of C, singular values <= eps cut; identity for alpha = 1
    al = _pair(alpha)
    Tinv = [None, None]
    mats = [A1, A2]
    for i in range(2):
        if float(al[i]) == 1.0:
            continue
        T, Tinv[i] = whitener_transform(mats[i], float(al[i]))
        mats[i] = mats[i] @ T
    A1u, A2u = A1, A2   # un-whitened (PCA-space or physical) data
    A1, A2 = mats
    C = cross_covariance(A1, A2)
    Q1, s, Q2 = decompose(C, n_modes=n_modes, solver=solver, random_state=random_state,
                          solver_kwargs=solver_kwargs)
    # cpcca.py:991-1000: total squared covariance of the UN-whitened cross-covariance
    tsc = (np.abs(cross_covariance(A1u, A2u)) ** 2).sum()
    scores1 = A1 @ Q1
    scores2 = A2 @ Q2
    Q1_w, Q2_w = Q1, Q2   # singular vectors in the (whitened, PCA) space they were computed in
    # the accessors return the patterns un-whitened (whitener.py:201-213) and in physical space (pca.py:161-171)
    if Tinv[0] is not None:
        Q1 = Tinv[0].conj().T @ Q1
    if Tinv[1] is not None:
        Q2 = Tinv[1].conj().T @ Q2
    if use_pca:
        Q1, Q2 = V1 @ Q1, V2 @ Q2
    # cpcca.py:418-512 squared covariance fraction; :331-416 correlation coefficients of the scores
    Q1w, Q2w = Q1_w, Q2_w
    scf = []
    for m_ in range(Q1w.shape[1]):
        X1r = np.outer(scores1[:, m_], Q1w[:, m_].conj())
        X2r = np.outer(scores2[:, m_], Q2w[:, m_].conj())
        if Tinv[0] is not None:
            X1r = X1r @ Tinv[0]
        if Tinv[1] is not None:
            X2r = X2r @ Tinv[1]
        res = residual_squared_covariance(A1u, A2u, X1r, X2r)
        scf.append(max(0.0, 1.0 - res / tsc))

    def _corr(A, B):
        A = A / A.std(axis=0)
        B = B / B.std(axis=0)
        return A.conj().T @ B / (A.shape[0] - 1)
    # homogeneous / heterogeneous patterns (cpcca.py:726-898; utils/optional/statistics.py:51-106): Pearson
    # correlation of the (back-transformed) input data with the scores, two-sided p-values from the beta distribution
    import scipy.stats
    P1 = A1u @ V1.conj().T if use_pca else A1u
    P2 = A2u @ V2.conj().T if use_pca else A2u

    def _pearson(X, Y):
        r = pearson_correlation(X, Y)
        a = X.shape[0] / 2 - 1
        return r, 2 * scipy.stats.beta(a, a, loc=-1, scale=2).cdf(-np.abs(r))
    hom1, phom1 = _pearson(P1, scores1)
    hom2, phom2 = _pearson(P2, scores2)
    het1, _ = _pearson(P1, scores2)
    het2, _ = _pearson(P2, scores1)
    # transform / predict / inverse_transform (cpcca.py:227-306; base_model_cross_set.py:323-463)
    G_pred = scores1.conj().T @ scores2 / np.linalg.norm(scores1, axis=0) ** 2

    def _to_model_space(A, V, T):
        A = A @ V if V is not None else A
        return A @ T if T is not None else A
    T1 = None if Tinv[0] is None else np.linalg.inv(Tinv[0])
    T2 = None if Tinv[1] is None else np.linalg.inv(Tinv[1])
    helpers = {
        "transform1": lambda A: _to_model_space(A, V1 if use_pca else None, T1) @ Q1_w,
        "transform2": lambda A: _to_model_space(A, V2 if use_pca else None, T2) @ Q2_w,
        "predict": lambda A: _to_model_space(A, V1 if use_pca else None, T1) @ Q1_w @ G_pred,
    }
    return {
        "helpers": helpers, "components_model": (Q1_w, Q2_w),
        "homogeneous_patterns": (hom1, hom2), "pvalues_homogeneous": (phom1, phom2),
        "heterogeneous_patterns": (het1, het2),
        "squared_covariance_fraction": np.array(scf),
        "cross_correlation_coefficients": np.diag(_corr(scores1, scores2)).real,
        "correlation_coefficients_X": _corr(scores1, scores1),
        "n_pca_modes": None if V1 is None else (V1.shape[1], V2.shape[1]),
        "A1": A1, "A2": A2, "fitted1": f1, "fitted2": f2, "C": C,
        "components1_2d": Q1, "components2_2d": Q2,
        "scores1": scores1, "scores2": scores2,
        "singular_values": s,
        "squared_covariance": s**2,
        "total_squared_covariance": tsc,
        "idx_modes_sorted": np.argsort(s)[::-1],
        "norm1": np.sqrt((scores1.conj() * scores1).sum(axis=0)).real,
        "norm2": np.sqrt((scores2.conj() * scores2).sum(axis=0)).real,
    }
