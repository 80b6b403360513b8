"""Maximum Covariance Analysis on B200 — drop-in for ``xeofs.cross.MCA`` (cross/mca.py:88-123; fit template
cross/base_model_cross_set.py:269-321; algorithm cross/cpcca.py:168-225) for the ``use_pca=False`` path.

The cross-covariance C = X^T Y / (n-1) is never materialised (at BASELINE config 3 it would be 269 GB): the
range finder applies it as two streaming passes, C Q = X^T (Y Q) / (n-1).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _engine as E
from .. import _labels as L
from .._cuda_ops import CudaOps
from .._lib import lpad
from .._preprocessor import Preprocessor


def _pair(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v)


class MCA:
    """Same parameters as the reference (cross/mca.py:88-123).  ``use_pca`` defaults to False here: the
    reference's default PCA pre-projection (rank-0.3*T randomized SVD of each field, unseeded —
    cross/base_model_cross_set.py:165-179) is outside this build's scope and ``use_pca=True`` raises."""

    def __init__(self, n_modes=2, standardize=False, use_coslat=False, check_nans=True, use_pca=False,
                 n_pca_modes=0.999, pca_init_rank_reduction=0.3, compute=True, sample_name="sample",
                 feature_name="feature", solver="auto", random_state=None, solver_kwargs=None, *,
                 device=None, algo="auto", distributed=False, total_squared_covariance=True, ops=None):
        if any(_pair(use_pca)):
            raise NotImplementedError(
                "xeofs_b200.cross.MCA implements the use_pca=False path (implicit cross-covariance operator); "
                "the PCA pre-projection of the reference is not built."
            )
        self._params = dict(n_modes=n_modes, standardize=standardize, use_coslat=use_coslat, check_nans=check_nans,
                            use_pca=use_pca, compute=compute, sample_name=sample_name, feature_name=feature_name,
                            solver=solver, random_state=random_state, solver_kwargs=dict(solver_kwargs or {}))
        self.attrs = {"model": "Maximum Covariance Analysis", "backend": "xeofs_b200"}
        self.ops = ops if ops is not None else CudaOps(device=device, algo=algo)
        self.comm = E.Comm() if distributed else E.NO_COMM
        std, cos, chk = _pair(standardize), _pair(use_coslat), _pair(check_nans)
        mk = lambda i: Preprocessor(self.ops, with_center=True, with_std=std[i], with_coslat=cos[i],  # noqa: E731
                                    check_nans=chk[i], sample_name=sample_name, comm=self.comm)
        self.preprocessor1, self.preprocessor2 = mk(0), mk(1)
        self._want_tsc = total_squared_covariance
        self.data = {}

    def fit(self, X, Y, dim, weights_X=None, weights_Y=None):
        for a in (X, Y):
            L.validate_input_type(a)
        f1 = self.preprocessor1.fit_transform(X, dim, weights_X)
        f2 = self.preprocessor2.fit_transform(Y, dim, weights_Y)
        self._fit_algorithm(f1, f2)
        return self

    def _shard_offset(self, S):
        if not self.comm.active:
            return 0
        sizes = torch.zeros(self.comm.world, dtype=torch.int64, device=self.ops.device)
        sizes[self.comm.rank] = S
        self.comm.sum_(sizes)
        return int(sizes[: self.comm.rank].sum().item())

    def _fit_algorithm(self, f1, f2):
        ops, comm, p = self.ops, self.comm, self._params
        op = E.CrossOperator(ops, f1, f2, comm)
        k = p["n_modes"]
        rank = min(op.shape)
        if not isinstance(k, (int, np.integer)):
            raise NotImplementedError("variance-based n_modes is not supported for MCA in this build")
        if k > rank:
            raise ValueError(f"n_modes must be less than or equal to the rank of the dataset (rank = {rank}).")
        solver, kw = p["solver"], dict(p["solver_kwargs"])
        if solver == "auto":
            use_exact = max(op.shape) < 500 and k > int(0.8 * rank)
        elif solver in ("full", "randomized"):
            use_exact = solver == "full"
        else:
            raise ValueError(f"Unrecognized solver '{solver}'. Valid options are 'auto', 'full', and 'randomized'.")
        if use_exact:
            n_over, n_iter = rank - k, 2
        else:
            n_over, n_iter = kw.get("n_oversamples", 10), kw.get("n_iter", "auto")
        c_field = f1 if op.transposed else f2
        Ur, s, Vc, infos = E.randomized_svd(ops, op, k, n_oversamples=n_over, n_iter=n_iter,
                                            random_state=p["random_state"], comm=comm)
        E.check_infos(infos)
        # C = Q1 s Q2^T;  M = C^T when transposed
        Q1t, Q2t = (Vc, Ur) if op.transposed else (Ur, Vc)
        sign = E.sign_flip(ops, Q2t, k, f2.S, comm)  # decomposer.py:219-222: rule on V_ = Q2, applied to both
        ops.finish_components(Q2t, k, f2.S, sign, None)
        ops.finish_components(Q1t, k, f1.S, sign, None)
        # scores = X Q (cpcca.py:204-205), norms (207-208)
        sc1 = ops.project_T(f1.field, Q1t, k, algo=ops.accurate_algo)
        sc2 = ops.project_T(f2.field, Q2t, k, algo=ops.accurate_algo)
        comm.sum_(sc1)
        comm.sum_(sc2)
        n1 = torch.sqrt(torch.diagonal(ops.gram(sc1, f1.T, k, 0)))
        n2 = torch.sqrt(torch.diagonal(ops.gram(sc2, f2.T, k, 0)))
        self.k = k
        self._Q1t, self._Q2t, self._sc1, self._sc2, self._s = Q1t, Q2t, sc1, sc2, s
        self._f1, self._f2 = f1, f2
        self.data = {
            "singular_values": s, "squared_covariance": s**2, "norm1": n1, "norm2": n2,
            "idx_modes_sorted": torch.argsort(s, descending=True),
        }
        if self._want_tsc:
            self.data["total_squared_covariance"] = self._total_squared_covariance()
        return self

    def _total_squared_covariance(self):
        """cpcca.py:991-1000: sum |C|^2 = <X X^T, Y Y^T>_F / (n-1)^2, from the two T x T Gram matrices built 128
        columns at a time with the streaming product.  The Gram matrices are symmetric: for the column block starting
        at sample t0 only the samples t >= t0 are streamed, the strictly lower part counts twice."""
        from .._cuda_ops import Field
        ops, comm = self.ops, self.comm
        T = self._f1.T
        acc = torch.zeros((), dtype=torch.float64, device=ops.device)
        # each Gram entry is a sum over all features: with >= 2^16 of them the rounding noise of a single TF32 product
        # (operands rounded to nearest, 2e-4 per term) averages out below 1e-6; smaller fields take the 3xTF32 product
        S_all = min(self._f1.S_global, self._f2.S_global)
        algo = getattr(ops, "sum_algo", ops.accurate_algo) if S_all >= 65536 else ops.accurate_algo

        def tail(f, t0):  # the field from sample t0 on (same Scaler vectors)
            rv = None if f.row_valid is None else f.row_valid[t0:]
            return Field(f.X[t0:], f.pivot, f.dscale, f.ccorr, f.valid, f.mean, f.std, rv)

        for t0 in range(0, T, 128):
            t1 = min(T, t0 + 128)
            w = t1 - t0
            g = []
            for ff in (self._f1, self._f2):
                blk = ops.scaled_rows(ff.field, t0, t1)
                gi = ops.project_T(tail(ff.field, t0), blk, w, algo=algo)
                comm.sum_(gi)
                g.append(gi[:, :w].double())
            prod = g[0] * g[1]
            acc += prod[:w].sum() + 2.0 * prod[w:].sum()
        return float(acc.item()) / float(self._f1.n_samples - 1) ** 2

    # ------------------------------------------------------------------ accessors
    def components(self, normalized=True):
        c1 = self.preprocessor1.components_to_nd(self._Q1t, self.k, "components1")
        c2 = self.preprocessor2.components_to_nd(self._Q2t, self.k, "components2")
        return c1, c2

    def scores(self, normalized=False):
        s1, s2 = self._sc1, self._sc2
        if normalized:
            s1, s2 = s1.clone(), s2.clone()
            s1[:, : self.k] /= self.data["norm1"].to(torch.float32)[None, :]
            s2[:, : self.k] /= self.data["norm2"].to(torch.float32)[None, :]
        return (self.preprocessor1.scores_to_nd(s1, self.k, "scores1"),
                self.preprocessor2.scores_to_nd(s2, self.k, "scores2"))

    def _mode_array(self, t, name):
        return L.wrap(t.cpu().numpy(), ("mode",), {"mode": np.arange(1, self.k + 1)}, name,
                      self.preprocessor1.as_xarray)

    def singular_values(self):
        return self._mode_array(self.data["singular_values"], "singular_values")

    def squared_covariance(self):
        return self._mode_array(self.data["squared_covariance"], "squared_covariance")

    def total_squared_covariance(self):
        if "total_squared_covariance" not in self.data:
            self.data["total_squared_covariance"] = self._total_squared_covariance()
        return self.data["total_squared_covariance"]

    def squared_covariance_fraction(self):
        """cross/mca.py: SCF_i = s_i^2 / sum |C|^2."""
        return self._mode_array(self.data["squared_covariance"] / self.total_squared_covariance(),
                                "squared_covariance_fraction")

    def get_params(self):
        return dict(self._params)
