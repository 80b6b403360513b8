"""Varimax / Promax rotation of an EOF solution on B200 — drop-in for ``xeofs.single.EOFRotator``
(single/eof_rotator.py:57-225; algorithm linalg/_numpy/_rotation.py:6-187).

Each varimax iteration is ONE streaming pass over the (S x m) loadings: with Ln the Kaiser-normalised loadings,
B = Ln R and W = colsum(B^2),
    Ln^H (B o (B^2 - W/S)) = Ln^H B^3 - (1/S) (Ln^H Ln) R diag(W),
so the kernel accumulates Ln^H B^3 and W in fp64 and the m x m remainder is done on the small matrices.
The polar factor R = U V^T of svd(G) and delta = sum(svals) come from the eigen-decomposition of G^T G.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import _engine as E
from .. import _labels as L


TC_WINDOW = 8  # iterations the stopping test of the tensor-core phase of the varimax iteration looks back
TC_SYNC = 8    # tensor-core iterations between two reads of delta on the host
X1_RTOL = 1e-5  # the single-TF32 sweeps hand over to the 3xTF32 ones when delta moves less than this per iteration
# relative off-diagonal norm at which the Jacobi solver of the m x m step stops during the 3xTF32 phase
TC_EIG_TOL = float(os.environ.get("XEOFS_TC_EIG_TOL", "1e-9"))


class EOFRotator:
    def __init__(self, n_modes=2, power=1, max_iter=None, rtol=1e-8, compute=True):
        if max_iter is None:
            max_iter = 1000 if compute else 100  # eof_rotator.py:65-66
        self._params = dict(n_modes=n_modes, power=power, max_iter=max_iter, rtol=rtol, compute=compute)
        self.attrs = {"model": "Rotated EOF analysis", "backend": "xeofs_b200"}
        self.data = {}
        self.n_iter_ = 0

    # ------------------------------------------------------------------ rotation of a space-side loadings block
    def _rotate(self, ops, comm, L0, S_local, n_rows, m):
        """promax(loadings) of linalg/_numpy/_rotation.py:6-92.  Returns R_total (m x m fp64), phi, iterations.

        One iteration = one sweep over the loadings (``varimax_accumulate``: Ln^H B^3 and W) + the m x m step
        (``varimax_update``: G, its polar factor R = U V^T and delta = sum(svals), all on the device).  Nothing comes
        back to the host inside an iteration: delta is appended to a device array that is read every TC_SYNC iterations
        during the tensor-core phase."""
        p = self._params
        if m < 2:
            raise ValueError(f"Cannot rotate {m} modes (columns), but must be 2 or more.")
        if m > 128:
            raise NotImplementedError(f"rotation of {m} modes: the varimax kernels of this build take at most 128 "
                                      "(one TMEM lane per mode)")
        _, _, Ln = ops.col_norms(L0, S_local, m, normalized_out=True)
        XtX = ops.gram(Ln, S_local, m, 1)
        comm.sum_(XtX)
        R = torch.eye(m, dtype=torch.float64, device=ops.device)
        basis = torch.eye(m, dtype=torch.float64, device=ops.device)
        alpha = 1.0 / n_rows  # gamma = 1 (varimax)
        max_iter, rtol, test = int(p["max_iter"]), p["rtol"], bool(p["compute"])
        hist_dev = torch.zeros(max_iter, dtype=torch.float64, device=ops.device)
        # The reference stops when sum(svals) changes by less than rtol (1e-8) between two iterations — a test that only
        # fp64 sweeps can decide.  The tcgen05 sweep (fp32-level noise in delta, ~20x faster) therefore does the bulk
        # of the iterations with the same test taken over a window of TC_WINDOW iterations (the noise of the windowed
        # mean change is 2/TC_WINDOW of the noise of delta); the fp64 sweep then takes over until the reference's own
        # test holds, so the iteration ends where the reference's ends.  compute=False: no test at all, max_iter
        # iterations (_rotation.py:176-180), the last two in fp64.
        tc = getattr(ops, "_varimax_tc_applies", None)
        use_tc = bool(tc and rtol >= 1e-9 and tc(Ln, S_local, m, False))
        # Far from convergence the tensor-core phase starts with single-TF32 sweeps (a third of the tensor work): their
        # rounding noise (1e-3 per term, averaged over the features) is far below the change of the rotation there.
        x1 = use_tc and rtol < X1_RTOL
        # the loadings are constant over the iterations: the tensor-core sweeps read a tile-by-tile copy of them
        packed = ops.varimax_pack(Ln, S_local, m) if use_tc else None
        hist, read, d, d_old, converged, it = [], 0, None, None, False, 0
        self.n_iter_tc_ = self.n_iter_x1_ = 0
        for it in range(1, max_iter + 1):
            if use_tc and not test and it > max_iter - 2:
                use_tc, self.n_iter_tc_ = False, it - 1
            G3, W, _ = ops.varimax_accumulate(Ln, S_local, m, R, exact=not use_tc, products=1 if (use_tc and x1) else 3,
                                              packed=packed)
            comm.sum_(G3)
            comm.sum_(W)
            ops.varimax_update(G3, W, XtX, alpha, R, basis, hist_dev[it - 1:it],
                               eig_tol=(1e-6 if x1 else TC_EIG_TOL) if use_tc else 0.0)
            if use_tc:
                if not test or (it - read < TC_SYNC and it < max_iter):
                    continue
                vals = hist_dev[read:it].cpu().tolist()  # one host sync per TC_SYNC iterations
                read = it
                for v in vals:
                    hist.append(v)
                    w = min(TC_WINDOW, len(hist) - 1)
                    if x1:
                        if w >= 1 and abs(v - hist[-1 - w]) / (w * v) < X1_RTOL:
                            x1, self.n_iter_x1_ = False, it
                            hist = []  # the two arithmetics differ by a constant in delta: a fresh window
                            break
                    elif w >= 1 and abs(v - hist[-1 - w]) / (w * v) < rtol:
                        use_tc = False  # the next fp64 sweep has no fp64 predecessor to compare with
                        self.n_iter_tc_ = it
                        break
                d = None
                continue
            if not test:
                continue
            d_old, d = d, float(hist_dev[it - 1].item())
            if d_old is not None and abs(d - d_old) / d < rtol:
                converged = True
                break
        if test and not converged and not use_tc and self.n_iter_tc_ >= max_iter - 2:
            # the windowed test of the tensor-core phase passed within the last iterations allowed: let the fp64 sweeps
            # that confirm it (normally two) run before the verdict
            for _ in range(3):
                G3, W, _ = ops.varimax_accumulate(Ln, S_local, m, R, exact=True)
                comm.sum_(G3)
                comm.sum_(W)
                ops.varimax_update(G3, W, XtX, alpha, R, basis, hist_dev[0:1])
                d_old, d = d, float(hist_dev[0].item())
                if d_old is not None and abs(d - d_old) / d < rtol:
                    converged = True
                    break
        if test and not converged:
            raise RuntimeError("Rotation process did not converge.")  # _rotation.py:179-180
        self.n_iter_ = it
        eye = torch.eye(m, dtype=torch.float64, device=ops.device)
        power = p["power"]
        if power == 1:  # Lr = I, phi = I up to round-off (_rotation.py:57-90)
            return R, eye
        _, _, amax = ops.varimax_accumulate(Ln, S_local, m, R, want_absmax=True)
        comm.max_(amax)
        Gp, _, _ = ops.varimax_accumulate(Ln, S_local, m, R, power=float(power), colscale=(1.0 / amax.double()))
        comm.sum_(Gp)
        ZtP = R.t() @ Gp
        ZtZ = R.t() @ XtX @ R
        Lr = torch.linalg.solve(ZtZ, ZtP)
        sig = torch.diagonal(torch.linalg.inv(Lr.t() @ Lr))
        Lr = Lr * torch.sqrt(sig)[None, :]
        Li = torch.linalg.inv(Lr)
        return R @ Lr, Li @ Li.t()

    # ------------------------------------------------------------------ fit (eof_rotator.py:103-225)
    def fit(self, model):
        ops, comm = model.ops, model.comm
        self.model, self.ops, self.comm = model, ops, comm
        self.preprocessor = model.preprocessor
        ff = model.preprocessor.fitted
        p = self._params
        m = int(p["n_modes"])
        if m > model.k:
            raise ValueError(f"n_modes={m} exceeds the {model.k} modes of the EOF model")
        S, T = ff.S, ff.T
        s = model._s[:m]
        ev = (s**2 / (ff.n_samples - 1))
        eye = torch.eye(m, dtype=torch.float64, device=ops.device)
        # loadings = components * sqrt(expvar)   (eof_rotator.py:131-135)
        L0 = ops.apply(model._Vt, S, m, 1, eye, m, colscale=torch.sqrt(ev))
        Rt, phi = self._rotate(ops, comm, L0, S, ff.n_features, m)
        # expvar = sum |L_rot|^2 = diag(Rt^T (L0^T L0) Rt)   (eof_rotator.py:155)
        G0 = ops.gram(L0, S, m, 1)
        comm.sum_(G0)
        expvar = torch.diagonal(Rt.t() @ G0 @ Rt).clone()
        idx = torch.argsort(expvar, descending=True)  # :156
        expvar_s = expvar[idx]
        norms = torch.sqrt(expvar_s * (ff.n_samples - 1))  # :163-164
        # components = L_rot / sqrt(expvar), written already in sorted order (:160, 215-225)
        Vt = ops.apply(L0, S, m, 1, Rt[:, idx].contiguous(), m, colscale=1.0 / torch.sqrt(expvar_s))
        sign = E.sign_flip(ops, Vt, m, S, comm)  # :184-188
        ops.finish_components(Vt, m, S, sign, None)
        # scores = (scores / svals) R^-T * norms * sign   (:168-181)
        RinvT = Rt if p["power"] == 1 else torch.linalg.inv(Rt).t()
        Mat = (RinvT[:, idx] / s[:, None]).contiguous()
        scores = ops.apply(model._scores, T, m, 0, Mat, m, colscale=norms * sign.double())
        self.k = m
        self._Vt, self._scores, self._s = Vt, scores, norms
        self.data = {
            "norms": norms, "explained_variance": expvar_s, "total_variance": model.data["total_variance"],
            "idx_modes_sorted": idx, "rotation_matrix": Rt, "phi_matrix": phi, "modes_sign": sign,
        }
        return self

    # ------------------------------------------------------------------ accessors
    def components(self, normalized=True):
        Vt = self._Vt
        if not normalized:
            Vt = Vt[: self.k] * self._s.to(torch.float32)[:, None]
        return self.preprocessor.components_to_nd(Vt, self.k, "components")

    def scores(self, normalized=False):
        Sc = self._scores
        if normalized:
            Sc = Sc.clone()
            Sc[:, : self.k] /= self._s.to(torch.float32)[None, :]
        return self.preprocessor.scores_to_nd(Sc, self.k, "scores")

    def _mode_array(self, t, name):
        return L.wrap(t.cpu().numpy(), ("mode",), {"mode": np.arange(1, self.k + 1)}, name,
                      self.preprocessor.as_xarray)

    def singular_values(self):
        return self._mode_array(self.data["norms"], "singular_values")

    def explained_variance(self):
        return self._mode_array(self.data["explained_variance"], "explained_variance")

    def explained_variance_ratio(self):
        return self._mode_array(self.data["explained_variance"] / self.data["total_variance"], "explained_variance_ratio")

    def rotation_matrix(self):
        return self.data["rotation_matrix"].cpu().numpy()

    def phi_matrix(self):
        return self.data["phi_matrix"].cpu().numpy()

    def inverse_transform(self, scores, normalized=False):
        """Inherited from EOF in the reference (single/eof.py:134-156 with the rotated components): scores .
        components^H, un-scaled.  For an orthogonal rotation the m rotated modes span what the first m EOFs span."""
        from .eof import EOF
        return EOF.inverse_transform(self, scores, normalized)

    def get_params(self):
        return dict(self._params)

    def transform(self, data, normalized=False):
        """eof_rotator.py:227-263: project on the un-rotated components, rotate, reorder, scale, sign."""
        model, p = self.model, self._params
        new, sample_shape, sample_coords, valid_sample = self.preprocessor.transform(data)
        m = self.k
        Z = self.ops.project_T(new, model._Vt, m, algo=self.ops.accurate_algo)
        self.comm.sum_(Z)
        Rt = self.data["rotation_matrix"]
        RinvT = Rt if p["power"] == 1 else torch.linalg.inv(Rt).t()
        idx = self.data["idx_modes_sorted"]
        Mat = (RinvT[:, idx] / model._s[:m, None]).contiguous()
        scale = self.data["modes_sign"].double() * (1.0 if normalized else self.data["norms"])
        Zr = self.ops.apply(Z, int(Z.shape[0]), m, 0, Mat, m, colscale=scale * torch.ones_like(self.data["norms"]))
        return self.preprocessor.scores_to_nd(Zr, m, "scores", sample_shape, sample_coords, valid_sample)
