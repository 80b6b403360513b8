"""Bootstrap of a fitted EOF model on B200 — drop-in for ``xeofs.validation.EOFBootstrapper``
(validation/bootstrapper.py:46-135).

Every bootstrap member is one more pass of the hot path: the preprocessed field (the model's ``input_data``) is
materialised once on the device, its samples are re-drawn with replacement by a row gather, ``EOF.fit`` runs on the
gathered matrix and the original samples are projected on the member's components.  The members are independent; on a
multi-GPU box they run either feature-sharded like the model itself (``distributed=True`` models) or as replicas
(different ``seed`` per rank, caller-side).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _labels as L
from .._cuda_ops import Field
from .._lib import lpad
from ..single.eof import EOF


class EOFBootstrapper:
    """``seed`` drives the resampling as in the reference.  The reference fits its members with an unseeded randomized
    SVD (bootstrapper.py:89); ``random_state`` (keyword-only, not in the reference) seeds it for reproducible runs."""

    def __init__(self, n_bootstraps=20, seed=None, *, random_state=None):
        self._params = dict(n_bootstraps=n_bootstraps, seed=seed, random_state=random_state)
        self.attrs = {"model": "Bootstrapped EOF analysis", "backend": "xeofs_b200"}
        self.data = {}

    def fit(self, model):
        """bootstrapper.py:56-135."""
        ops, comm = model.ops, model.comm
        self.model, self.ops, self.preprocessor = model, ops, model.preprocessor
        ff = model.preprocessor.fitted
        k = model.k
        n_boot = int(self._params["n_bootstraps"])
        # the model's input_data: scaled, weighted samples; dropped features stay as zero columns, dropped samples go
        A = ops.scaled_rows(ff.field, 0, ff.T)[: ff.T]
        if ff.n_samples < ff.T:
            A = A[ff.valid_sample]
        n = int(A.shape[0])
        rng = np.random.default_rng(self._params["seed"])  # :72
        dims = ("sample", "feature")
        model_scores = model._scores[:, :k] if ff.n_samples == ff.T else model._scores[ff.valid_sample][:, :k]
        ms = model_scores.double()
        expvar, totvar, comps, scores = [], [], [], []
        for _ in range(n_boot):
            idx = rng.choice(n, n, replace=True)  # :81
            Ab = A.index_select(0, torch.as_tensor(idx, device=A.device))
            # :89-90 — no scaling, the data are the model's pre-scaled samples
            bm = EOF(n_modes=k, standardize=False, use_coslat=False, ops=ops, distributed=comm.active,
                     random_state=self._params["random_state"])
            bm.fit(L.DataArray(Ab, dims), dim="sample")
            del Ab
            # :95 — scores of the ORIGINAL samples on the member's components (its own centring, no re-scan of A:
            # the NaN pattern of A is the one the member was fitted on)
            bf = bm.preprocessor.fitted.field
            new = Field(A, bf.pivot, bf.dscale, bf.ccorr, bf.valid, bf.mean, bf.std, None)
            Z = ops.project_T(new, bm._Vt, k, algo=ops.accurate_algo)
            comm.sum_(Z)
            # :117-123 — sign of every mode from its correlation with the model's scores
            z = Z[:, :k].double()
            corr = ((z * ms).mean(0) / z.std(0, unbiased=False) / ms.std(0, unbiased=False))
            sgn = torch.sign(corr)
            Vt = bm._Vt[:k] * sgn.to(torch.float32)[:, None]
            Z = Z[:, :k] * sgn.to(torch.float32)[None, :]
            expvar.append(bm.data["explained_variance"][:k])
            totvar.append(float(bm.data["total_variance"]))
            comps.append(Vt)
            scores.append(Z)
        self.k, self.n = k, n_boot
        self._components, self._scores = comps, scores
        self.data = {
            "explained_variance": torch.stack(expvar), "total_variance": torch.as_tensor(totvar),
            "norms": model.data["norms"],
        }
        return self

    # ------------------------------------------------------------------ accessors (dimension "n" leads)
    def _coords_n(self):
        return np.arange(1, self.n + 1)

    def explained_variance(self):
        return L.wrap(self.data["explained_variance"].cpu().numpy(), ("n", "mode"),
                      {"n": self._coords_n(), "mode": np.arange(1, self.k + 1)}, "explained_variance",
                      self.preprocessor.as_xarray)

    def total_variance(self):
        return L.wrap(self.data["total_variance"].cpu().numpy(), ("n",), {"n": self._coords_n()}, "total_variance",
                      self.preprocessor.as_xarray)

    def components(self):
        """List over bootstrap members of feature-shaped component arrays (NaN at the dropped features)."""
        out = []
        for Vt in self._components:
            buf = self.ops.space_side(lpad(self.k), int(Vt.shape[1]), zero=True)
            buf[: self.k] = Vt
            out.append(self.preprocessor.components_to_nd(buf, self.k, "components"))
        return out

    def scores(self):
        out = []
        ff = self.preprocessor.fitted
        for Z in self._scores:
            buf = self.ops.zeros((ff.T, lpad(self.k)))
            if ff.n_samples < ff.T:
                buf[ff.valid_sample, : self.k] = Z
            else:
                buf[:, : self.k] = Z
            out.append(self.preprocessor.scores_to_nd(buf, self.k, "scores"))
        return out

    def get_params(self):
        return dict(self._params)
