"""Oracle (test infrastructure only): varimax / promax and the EOFRotator post-processing.

Written fresh from the algorithm's description; behaviour follows (/root/reference/xeofs):
  linalg/_numpy/_rotation.py:95-187   varimax: Kaiser row normalisation with +eps, alpha = gamma / n_rows,
                                      R <- U V^T of svd(X^H (B o (|B|^2 - alpha*colsum|B|^2))), delta = sum(svals),
                                      stop when |delta - delta_old| / delta < rtol, else RuntimeError
  linalg/_numpy/_rotation.py:6-92     promax: varimax, re-normalise rows, column max-normalise, P = Xn|Xn|^(p-1),
                                      L = (X^H X)^-1 X^H P, rescale by sqrt(diag((L^H L)^-1)), phi = L^-1 L^-H
  single/eof_rotator.py:119-209       loadings = V sqrt(expvar); expvar = sum |L|^2; idx sort desc;
                                      components = L / sqrt(expvar); norms = sqrt(expvar (n-1));
                                      scores = (scores / svals) R^-T * norms; sign rule on components
  single/eof_rotator.py:215-225       re-ordering of every mode-dimensioned array by idx_modes_sorted
tests/golden/make_golden.py executes the reference's own _varimax/_promax source to pin this file.
"""
from __future__ import annotations

import numpy as np

from .decomposer import sign_multiplier


def varimax(X, gamma=1.0, max_iter=1000, rtol=1e-8):
    L = np.array(X, copy=True)
    p, m = L.shape
    if m < 2:
        raise ValueError(f"Cannot rotate {m} modes (columns), but must be 2 or more.")
    h = np.sqrt((L * L.conj()).sum(axis=1))
    eps = np.finfo(L.dtype).eps
    Ln = L / (h + eps)[:, None]
    R = np.eye(m)
    alpha = gamma / p
    d_old, d = 0.0, 0.0
    LH = Ln.conj().T
    converged = False
    for _ in range(max_iter):
        d_old = d
        B = Ln @ R
        B2 = B * B.conj()
        G = LH @ (B * (B2 - alpha * B2.sum(axis=0)))
        U, sv, VT = np.linalg.svd(G)
        R = U @ VT
        d = sv.sum()
        if abs(d - d_old) / d < rtol:
            converged = True
            break
    if not converged and abs(d - d_old) / d > rtol:
        raise RuntimeError("Rotation process did not converge.")
    return (h[:, None] * Ln) @ R, R


def promax(X, power=1, max_iter=1000, rtol=1e-8):
    Lv, R = varimax(X, max_iter=max_iter, rtol=rtol)
    h = np.sqrt((Lv * Lv.conj()).sum(axis=1))
    eps = np.finfo(Lv.dtype).eps
    Z = Lv / (h + eps)[:, None]
    Zn = Z / np.abs(Z).max(axis=0)
    P = Zn * np.abs(Zn) ** (power - 1)
    Lr = np.linalg.inv(Z.conj().T @ Z) @ Z.conj().T @ P
    try:
        sig = np.diag(np.diag(np.linalg.inv(Lr.conj().T @ Lr)))
    except np.linalg.LinAlgError:
        sig = np.diag(np.diag(np.linalg.pinv(Lr.conj().T @ Lr)))
    Lr = Lr @ np.sqrt(sig)
    Xrot = h[:, None] * (Z @ Lr)
    R = R @ Lr
    Li = np.linalg.inv(Lr)
    return Xrot, R, Li @ Li.conj().T


def eof_rotator_fit(components_2d, explained_variance, scores, norms, n_samples,
                    n_modes=2, power=1, max_iter=None, rtol=1e-8):
    """components_2d (S', k), explained_variance (k), scores (n, k) = U*s, norms = s."""
    if max_iter is None:
        max_iter = 1000  # compute=True default (100 when compute=False), eof_rotator.py:65-66
    m = n_modes
    V = components_2d[:, :m]
    ev = explained_variance[:m]
    loadings = V * np.sqrt(ev)
    Lrot, R, phi = promax(loadings, power=power, max_iter=max_iter, rtol=rtol)
    expvar = (np.abs(Lrot) ** 2).sum(axis=0)
    idx = np.argsort(expvar)[::-1]
    comps = Lrot / np.sqrt(expvar)
    nrm = (expvar * (n_samples - 1)) ** 0.5
    sc = scores[:, :m] / norms[:m]
    RinvT = R
    if power > 1:
        RinvT = np.linalg.inv(R).conj().T
    sc = sc @ RinvT
    sc = sc * nrm
    sgn = sign_multiplier(comps.T)
    comps = comps * sgn
    sc = sc * sgn
    return {
        "components_2d": comps[:, idx],
        "scores": sc[:, idx],
        "norms": nrm[idx],
        "explained_variance": expvar[idx],
        "idx_modes_sorted": idx,
        "rotation_matrix": R,
        "phi_matrix": phi,
        "modes_sign": sgn[idx],
    }


def mca_rotator_fit(components1_2d, components2_2d, singular_values, scores1, scores2,
                    n_modes=2, power=1, max_iter=None, rtol=1e-8, model_components=None):
    """MCARotator / CPCCARotator with identity whitening and no PCA stage (cross/cpcca_rotator.py:122-305;
    cross/mca_rotator.py:5): varimax/promax of the concatenated, sqrt(s)-weighted singular vectors.
    components*_2d (S', k) valid features only, scores* (n, k) = X Q."""
    if max_iter is None:
        max_iter = 1000  # compute=True default, cpcca_rotator.py:86-87
    m = n_modes
    scaling = np.sqrt(singular_values[:m])                                   # :154-155
    S1 = components1_2d.shape[0]
    loadings = np.concatenate([components1_2d[:, :m], components2_2d[:, :m]], axis=0) * scaling   # :171
    Lrot, R, phi = promax(loadings, power=power, max_iter=max_iter, rtol=rtol)   # :175-180
    Q1r, Q2r = Lrot[:S1], Lrot[S1:]                                          # :193-198
    if model_components is None:
        n1 = np.linalg.norm(Q1r, axis=0)                                     # :211-232
        n2 = np.linalg.norm(Q2r, axis=0)
    else:
        # whitened / PCA models: the rotated vectors go back into the whitened PCA space first (:203-208), where they
        # are (Q_model sqrt(s)) R
        n1 = np.linalg.norm((model_components[0][:, :m] * scaling) @ R, axis=0)
        n2 = np.linalg.norm((model_components[1][:, :m] * scaling) @ R, axis=0)
    Q1r, Q2r = Q1r / n1, Q2r / n2                                            # :235-236
    sqcov = (n1 * n2) ** 2                                                   # :239-240
    idx = np.argsort(sqcov)[::-1]                                            # :243
    RinvT = R
    if power > 1:                                                            # :445-469
        RinvT = np.linalg.inv(R).conj().T
    sc1 = (scores1[:, :m] / scaling) @ RinvT * n1                            # :254-268
    sc2 = (scores2[:, :m] / scaling) @ RinvT * n2
    sgn = sign_multiplier(Lrot.T)                                            # :271 (rule on the combined loadings)

    def _transform(A, comps, nrm):
        """cpcca_rotator.py:322-427: preprocessed data A projected on the UN-rotated components — un-whitened and back
        in physical space (:359-366), which is what ``components*_2d`` holds — / sqrt(s), rotated, reordered, signed,
        scaled with the rotated norms."""
        return (((A @ comps[:, :m]) / scaling) @ RinvT)[:, idx] * sgn[idx] * nrm[idx]

    return {
        "transform1": lambda A: _transform(A, components1_2d, n1),
        "transform2": lambda A: _transform(A, components2_2d, n2),
        "components1_2d": (Q1r * sgn)[:, idx], "components2_2d": (Q2r * sgn)[:, idx],
        "scores1": (sc1 * sgn)[:, idx], "scores2": (sc2 * sgn)[:, idx],
        "squared_covariance": sqcov[idx], "norm1": n1[idx], "norm2": n2[idx],
        "idx_modes_sorted": idx, "rotation_matrix": R, "phi_matrix": phi, "modes_sign": sgn[idx],
    }
