"""Varimax / Promax rotation of an MCA / CCA / RDA / CPCCA solution on B200 — drop-in for
``xeofs.cross.CPCCARotator`` and ``xeofs.cross.MCARotator`` (cross/cpcca_rotator.py:57-469, cross/mca_rotator.py:5),
with or without the PCA stage — the rotation works on the un-whitened physical-space singular vectors either way.

The singular vectors of both fields, weighted with sqrt(singular value), are rotated as ONE (S1 + S2) x m loadings
matrix (cpcca_rotator.py:154-180): the same one-pass-per-iteration varimax sweep as ``EOFRotator`` runs on the
concatenated space-side block; everything after the rotation is m x m algebra plus one streaming apply per field.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _labels as L
from .._lib import lpad
from ..single.eof_rotator import EOFRotator


class CPCCARotator:
    def __init__(self, n_modes=10, power=1, max_iter=None, rtol=1e-8, compute=True):
        if max_iter is None:
            max_iter = 1000 if compute else 100  # cpcca_rotator.py:86-87
        self._params = dict(n_modes=n_modes, power=power, max_iter=max_iter, rtol=rtol, compute=compute)
        self.attrs = {"model": "Rotated MCA", "backend": "xeofs_b200"}
        self.data = {}
        self.n_iter_ = 0

    def fit(self, model):
        ops, comm = model.ops, model.comm
        self.model, self.ops, self.comm = model, ops, comm
        self.preprocessor1, self.preprocessor2 = model.preprocessor1, model.preprocessor2
        p = self._params
        m = int(p["n_modes"])
        if m > model.k:
            raise ValueError(f"n_modes={m} exceeds the {model.k} modes of the MCA model")
        f1, f2 = model._f1, model._f2
        S1, S2, T = f1.S, f2.S, f1.T
        s = model._s[:m]
        scaling = torch.sqrt(s)                                                    # cpcca_rotator.py:154-155
        eye = torch.eye(m, dtype=torch.float64, device=ops.device)
        # loadings = concat(Q1, Q2) * sqrt(s): one space-side block (:171)
        # (the second field starts at a multiple of 32 columns so that both blocks keep 128-byte aligned rows; the
        # zero columns in between are features without loading, which the rotation ignores)
        lp = lpad(m)
        S1p = (S1 + 31) // 32 * 32
        L0 = ops.space_side(lp, S1p + S2, zero=True)
        L1, L2 = L0[:, :S1], L0[:, S1p:S1p + S2]
        ops.apply(model._Q1t, S1, m, 1, eye, m, colscale=scaling, out=L1)
        ops.apply(model._Q2t, S2, m, 1, eye, m, colscale=scaling, out=L2)
        rot = EOFRotator(n_modes=m, power=p["power"], max_iter=p["max_iter"], rtol=p["rtol"])
        Rt, phi = rot._rotate(ops, comm, L0, S1p + S2, f1.n_features + f2.n_features, m)   # :175-180
        self.n_iter_, self.n_iter_tc_ = rot.n_iter_, getattr(rot, "n_iter_tc_", 0)
        # norms of the rotated, loaded vectors of each field: diag(Rt^T (L^T L) Rt)   (:211-232)
        if getattr(model, "_alpha", (1.0, 1.0)) != (1.0, 1.0):
            # whitened models (CCA / RDA / CPCCA): the reference takes the norms after transforming the rotated vectors
            # back into the whitened PCA space (cpcca_rotator.py:203-232), where they are Q sqrt(s) R with orthonormal
            # Q: the Gram matrix of the loadings there is diag(s) for both fields
            G1 = G2 = torch.diag(s.double())
        else:
            G1, G2 = ops.gram(L1, S1, m, 1), ops.gram(L2, S2, m, 1)
            comm.sum_(G1)
            comm.sum_(G2)
        n1 = torch.sqrt(torch.diagonal(Rt.t() @ G1 @ Rt)).clone()
        n2 = torch.sqrt(torch.diagonal(Rt.t() @ G2 @ Rt)).clone()
        sqcov = (n1 * n2) ** 2                                                     # :239-240
        idx = torch.argsort(sqcov, descending=True)                                # :243
        n1s, n2s = n1[idx], n2[idx]
        # rotated, normalised singular vectors, written already in sorted order (:235-236, 433-443)
        Rs = Rt[:, idx].contiguous()
        Q1r = ops.apply(L1, S1, m, 1, Rs, m, colscale=1.0 / n1s)
        Q2r = ops.apply(L2, S2, m, 1, Rs, m, colscale=1.0 / n2s)
        # sign rule on the combined rotated loadings (:271): extrema of each field's block, re-weighted by its norm
        mx1, mn1 = ops.row_minmax(Q1r, m, S1)
        mx2, mn2 = ops.row_minmax(Q2r, m, S2)
        for t in (mx1, mx2):
            comm.max_(t)
        for t in (mn1, mn2):
            comm.min_(t)
        vmax = torch.maximum(mx1.double() * n1s, mx2.double() * n2s)
        vmin = torch.minimum(mn1.double() * n1s, mn2.double() * n2s)
        sign = torch.where(vmax.abs() >= vmin.abs(), 1.0, -1.0).to(torch.float32)
        ops.finish_components(Q1r, m, S1, sign, None)
        ops.finish_components(Q2r, m, S2, sign, None)
        # scores = (scores / sqrt(s)) R^-T * norm * sign   (:246-275)
        RinvT = Rt if p["power"] == 1 else torch.linalg.inv(Rt).t()
        Mat = (RinvT[:, idx] / scaling[:, None]).contiguous()
        sc1 = ops.apply(model._sc1, T, m, 0, Mat, m, colscale=n1s * sign.double())
        sc2 = ops.apply(model._sc2, T, m, 0, Mat, m, colscale=n2s * sign.double())
        self.k = m
        self._Q1t, self._Q2t, self._sc1, self._sc2 = Q1r, Q2r, sc1, sc2
        self.data = {
            "squared_covariance": sqcov[idx], "norm1": n1s, "norm2": n2s, "idx_modes_sorted": idx,
            "rotation_matrix": Rt, "phi_matrix": phi, "modes_sign": sign,
        }
        if "total_squared_covariance" in model.data:
            self.data["total_squared_covariance"] = model.data["total_squared_covariance"]
        return self

    # ------------------------------------------------------------------ accessors
    def components(self, normalized=True):
        """cpcca.py:308-316 (inherited by the rotator): normalized=False scales every mode by norm1 / norm2."""
        Q1t, Q2t = self._Q1t, self._Q2t
        if not normalized:
            Q1t = Q1t[: self.k] * self.data["norm1"].to(torch.float32)[:, None]
            Q2t = Q2t[: self.k] * self.data["norm2"].to(torch.float32)[:, None]
        return (self.preprocessor1.components_to_nd(Q1t, self.k, "components1"),
                self.preprocessor2.components_to_nd(Q2t, self.k, "components2"))

    def scores(self, normalized=False):
        s1, s2 = self._sc1, self._sc2
        if normalized:
            s1, s2 = s1.clone(), s2.clone()
            s1[:, : self.k] /= self.data["norm1"].to(torch.float32)[None, :]
            s2[:, : self.k] /= self.data["norm2"].to(torch.float32)[None, :]
        return (self.preprocessor1.scores_to_nd(s1, self.k, "scores1"),
                self.preprocessor2.scores_to_nd(s2, self.k, "scores2"))

    def transform(self, X=None, Y=None, normalized=False):
        """cpcca_rotator.py:322-427: project the preprocessed data on the UN-rotated singular vectors — taken, as the
        reference takes them (:359-366), un-whitened and back in physical space, i.e. the model's components, not its
        whitening projection — divide by sqrt(s), rotate (R^-T), reorder, sign, scale with the rotated norms."""
        if X is None and Y is None:
            raise ValueError("No data provided. Please provide X and/or Y.")
        model, p, m = self.model, self._params, self.k
        Rt = self.data["rotation_matrix"]
        RinvT = Rt if p["power"] == 1 else torch.linalg.inv(Rt).t()
        idx = self.data["idx_modes_sorted"]
        scaling = torch.sqrt(model._s[:m])
        Mat = (RinvT[:, idx] / scaling[:, None]).contiguous()
        out = []
        for i, d in enumerate((X, Y)):
            if d is None:
                continue
            L.validate_input_type(d)
            pp = (self.preprocessor1, self.preprocessor2)[i]
            new, shp, crd, vs = pp.transform(d)
            Z = self.ops.project_T(new, (model._Q1t, model._Q2t)[i], m, algo=self.ops.accurate_algo)
            self.comm.sum_(Z)
            scale = self.data["modes_sign"].double()
            if not normalized:
                scale = scale * self.data["norm1" if i == 0 else "norm2"]
            Zr = self.ops.apply(Z, int(Z.shape[0]), m, 0, Mat, m, colscale=scale)
            out.append(pp.scores_to_nd(Zr, m, "scores1" if i == 0 else "scores2", shp, crd, vs))
        return out[0] if len(out) == 1 else out

    def _mode_array(self, t, name):
        return L.wrap(t.cpu().numpy(), ("mode",), {"mode": np.arange(1, self.k + 1)}, name,
                      self.preprocessor1.as_xarray)

    def squared_covariance(self):
        return self._mode_array(self.data["squared_covariance"], "squared_covariance")

    def total_squared_covariance(self):
        if "total_squared_covariance" not in self.data:
            self.data["total_squared_covariance"] = self.model.total_squared_covariance()
        return self.data["total_squared_covariance"]

    def squared_covariance_fraction(self):
        return self._mode_array(self.data["squared_covariance"] / self.total_squared_covariance(),
                                "squared_covariance_fraction")

    def rotation_matrix(self):
        return self.data["rotation_matrix"].cpu().numpy()

    def phi_matrix(self):
        return self.data["phi_matrix"].cpu().numpy()

    def get_params(self):
        return dict(self._params)


class MCARotator(CPCCARotator):
    """cross/mca_rotator.py:5: the rotator of MCA models (same algorithm)."""
