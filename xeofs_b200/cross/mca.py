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
    """Same parameters and defaults as the reference (cross/mca.py:88-123).

    ``use_pca=False``: the cross-covariance of the two preprocessed fields is decomposed directly, applied implicitly
    as two streaming passes per product (BASELINE config 3).
    ``use_pca=True`` (the reference's default, cross/base_model_cross_set.py:165-179, preprocessing/pca.py:94-131):
    each field is first projected on its leading principal components — int(pca_init_rank_reduction * rank) modes,
    cut at the ``n_pca_modes`` fraction of variance (linalg/_numpy/_svd.py:214-241) — and MCA runs on the two small
    score matrices.  The reference gets those components from an (unseeded) randomized SVD; here, with n_samples <
    n_features, they come from the eigen-decomposition of the n x n sample Gram matrix A A^T (streamed in 128-sample
    blocks through the tensor-core product), i.e. the converged limit of that randomized SVD.  ``use_pca`` must be
    the same for both fields."""

    def __init__(self, n_modes=2, standardize=False, use_coslat=False, check_nans=True, use_pca=True,
                 n_pca_modes=0.999, pca_init_rank_reduction=0.3, compute=True, sample_name="sample",
                 feature_name="feature", solver="auto", random_state=None, solver_kwargs=None, *,
                 device=None, algo="auto", distributed=False, total_squared_covariance=True, ops=None):
        if len(set(bool(v) for v in _pair(use_pca))) != 1:
            raise NotImplementedError("xeofs_b200.cross.MCA needs the same use_pca setting for both fields")
        self._params = dict(n_modes=n_modes, standardize=standardize, use_coslat=use_coslat, check_nans=check_nans,
                            use_pca=use_pca, n_pca_modes=n_pca_modes, pca_init_rank_reduction=pca_init_rank_reduction,
                            compute=compute, sample_name=sample_name, feature_name=feature_name,
                            solver=solver, random_state=random_state, solver_kwargs=dict(solver_kwargs or {}))
        self.attrs = {"model": "Maximum Covariance Analysis", "backend": "xeofs_b200"}
        self.ops = ops if ops is not None else CudaOps(device=device, algo=algo)
        self.comm = E.Comm() if distributed else E.NO_COMM
        std, cos, chk = _pair(standardize), _pair(use_coslat), _pair(check_nans)
        mk = lambda i: Preprocessor(self.ops, with_center=True, with_std=std[i], with_coslat=cos[i],  # noqa: E731
                                    check_nans=chk[i], sample_name=sample_name, comm=self.comm)
        self.preprocessor1, self.preprocessor2 = mk(0), mk(1)
        self._want_tsc = total_squared_covariance
        self._alpha = (1.0, 1.0)  # identity whitening = MCA (cross/mca.py:112-114)
        self.data = {}

    def fit(self, X, Y, dim, weights_X=None, weights_Y=None):
        for a in (X, Y):
            L.validate_input_type(a)
        self.data, self._tsc_cache = {}, None
        self.__dict__.pop("_pca_ctx", None)
        self.__dict__.pop("_proj_cache", None)
        f1 = self.preprocessor1.fit_transform(X, dim, weights_X)
        f2 = self.preprocessor2.fit_transform(Y, dim, weights_Y)
        if bool(_pair(self._params["use_pca"])[0]):
            self._fit_algorithm_pca(f1, f2)
        elif self._alpha != (1.0, 1.0):
            raise NotImplementedError(
                "fractional whitening (alpha < 1) is built on the PCA scores: use_pca=True is required (the reference "
                "itself warns that whitening the full feature space with n_samples < n_features is ill-conditioned, "
                "preprocessing/whitener.py:101-105)"
            )
        else:
            self._fit_algorithm(f1, f2)
        return self

    # ------------------------------------------------------------------ PCA stage (preprocessing/pca.py:94-131)
    def _gram_block_columns(self, ff, algo, rounded=None):
        """Block columns G[t0:, t0:t1] (t1 - t0 <= 128) of the sample Gram matrix A A^T of a fitted field, as
        ((t0, t1), (T - t0) x (t1 - t0) fp32 device block).  The matrix is symmetric: only the samples t >= t0 are
        streamed for the block starting at t0.  ``rounded``: a materialised TF32-rounded copy of the preprocessed
        matrix (ops.materialize) — its rows are both operands, moved as they are (XEOFS_ALGO_TF32X1F)."""
        from .._cuda_ops import Field
        from .. import _lib
        ops, comm = self.ops, self.comm
        f = ff.field
        if rounded is not None:
            A = rounded.X  # (rows padded to 128) x S view of the copy
            for t0 in range(0, ff.T, 128):
                t1 = min(ff.T, t0 + 128)
                w = t1 - t0
                tail = Field(A[t0:ff.T], rounded.pivot, rounded.dscale, None, None, no_nan=True)
                gi = ops.project_T(tail, A[t0:t0 + 128], w, algo=_lib.ALGO_TF32X1F)
                comm.sum_(gi)
                yield (t0, t1), gi[:, :w]
            return
        for t0 in range(0, ff.T, 128):
            t1 = min(ff.T, t0 + 128)
            w = t1 - t0
            rv = None if f.row_valid is None else f.row_valid[t0:]
            tail = Field(f.X[t0:], f.pivot, f.dscale, f.ccorr, f.valid, f.mean, f.std, rv, no_nan=f.no_nan)
            blk = ops.scaled_rows(f, t0, t1)
            gi = ops.project_T(tail, blk, w, algo=algo)
            comm.sum_(gi)
            yield (t0, t1), gi[:, :w]

    def _pca_stage(self, ff, n_pca_modes, init_rank_reduction):
        """Principal components of one preprocessed field (preprocessing/pca.py:94-131 -> linalg/_numpy/_svd.py:108-243):
        A = U S V^T by the same randomized range finder as EOF.fit, straight on the field through the streaming products
        (either orientation: n_samples < n_features or not), with the reference's truncation (:214-241) and sign rule
        (:205-210).  Returns (U (T x r) fp64, s (r) fp64, Vt (space-side, >= r rows, unit rows) fp32); the projection
        A V the reference hands on (pca.py:121-131) is U * s.

        The reference asks its (unseeded) randomized SVD for int(init_rank_reduction * rank) modes and then keeps the
        few that carry ``n_pca_modes`` of the variance.  Here the sketch starts one kernel block wide (118 modes + 10)
        and is widened only if the cut does not fall well inside it: the kept modes are the converged leading ones
        either way."""
        import warnings
        ops, comm = self.ops, self.comm
        n = ff.n_samples
        rank = min(n, ff.n_features)
        by_variance = isinstance(n_pca_modes, float)
        if by_variance:
            m0 = int(rank * init_rank_reduction)
            if m0 < 1:
                warnings.warn(f"`init_rank_reduction={init_rank_reduction}` is too low resulting in zero components. "
                              "One component will be computed instead.")
                m0 = 1
        elif n_pca_modes == "all":
            m0 = rank
        elif isinstance(n_pca_modes, (int, np.integer)):
            if n_pca_modes > rank:
                raise ValueError(f"n_modes must be less than or equal to the rank of the dataset (rank = {rank}).")
            m0 = int(n_pca_modes)
        else:
            raise ValueError("`n_modes` must be an integer, float or 'all'")
        op = E.FieldOperator(ops, ff, comm, algo=None)
        seed = self._params["random_state"]
        seed = 0 if seed is None else seed       # the reference leaves its PCA unseeded; a fixed draw here
        k_try = min(m0, E.MAX_L - 10) if by_variance else m0
        while True:
            n_over = min(10, rank - k_try)
            Ur, s, Vc, infos = E.randomized_svd(ops, op, k_try, n_oversamples=n_over, n_iter="auto",
                                                random_state=seed, comm=comm)
            E.check_infos(infos)
            r = k_try
            if not by_variance:
                break
            cum = torch.cumsum(s**2 / (n - 1) / ff.total_variance, 0)
            r = k_try - int((cum >= n_pca_modes).sum().item()) + 1       # _svd.py:223-227
            if r <= k_try - 8 or k_try == m0:
                if r > k_try:
                    warnings.warn(f"Dataset has {m0} components, explaining {float(cum[-1]):.2%} of the variance. "
                                  f"However, {n_pca_modes:.2%} explained variance was requested. Please consider "
                                  "increasing `init_rank_reduction`.")
                    r = k_try
                break
            k_try = min(m0, 2 * k_try + 10, E.MAX_L_TOTAL - 10)           # the cut is too close to the sketch's edge
            if 2 * r > k_try and k_try < min(m0, E.MAX_L_TOTAL - 10):
                k_try = min(m0, 4 * r, E.MAX_L_TOTAL - 10)
        Vt, Ut = (Ur, Vc) if op.transposed else (Vc, Ur)
        sign = E.sign_flip(ops, Vt, r, ff.S, comm)                      # _svd.py:205-210
        ops.finish_components(Vt, r, ff.S, sign, None)
        U = Ut[:, :r].double() * sign.double()[None, :]
        return U, s[:r].contiguous(), Vt

    def _decompose_small(self, C, k):
        """SVD of a small dense fp64 matrix (the cross-covariance of two PCA score matrices, decomposer.py:112-146)
        through the eigen-decomposition of its smaller Gram matrix, all on the library's own kernels.  The
        reference's randomized solver (n_iter = 7 for k < 0.1 rank) reproduces the exact SVD far inside the parity
        tolerance.  Returns (Q1 (r1 x k), s (k), Q2 (r2 x k))."""
        ops = self.ops
        r1, r2 = int(C.shape[0]), int(C.shape[1])
        if r2 <= r1:
            ev, V = ops.sym_eig(ops.dgemm(C, C, trans_a=True))          # C^T C = Q2 s^2 Q2^T
            s = torch.sqrt(torch.clamp(ev[:k], min=0.0))
            Q2 = V[:, :k].contiguous()
            Q1 = ops.dgemm(C, Q2) * torch.where(s > 0, 1.0 / s, torch.zeros_like(s))[None, :]
        else:
            ev, V = ops.sym_eig(ops.dgemm(C, C, trans_b=True))          # C C^T = Q1 s^2 Q1^T
            s = torch.sqrt(torch.clamp(ev[:k], min=0.0))
            Q1 = V[:, :k].contiguous()
            Q2 = ops.dgemm(C, Q1, trans_a=True) * torch.where(s > 0, 1.0 / s, torch.zeros_like(s))[None, :]
        return Q1, s, Q2

    @staticmethod
    def _modes_for_variance(s, frac, k0, total_variance, n_rows):
        """decomposer.py:188-216 for a float ``n_modes``: of the k0 precomputed modes keep the first that reach the
        requested fraction of the matrix' total variance (column variances over its rows, ddof = 1)."""
        import warnings
        cum = torch.cumsum(s[:k0] ** 2 / (n_rows - 1) / total_variance, 0)
        n_req = k0 - int((cum >= frac).sum().item()) + 1
        if n_req > k0:
            warnings.warn(f"Dataset has {k0} components, explaining {float(cum[-1]):.2%} of the variance. However, "
                          f"{frac:.2%} explained variance was requested. Please consider increasing "
                          "`init_rank_reduction`.")
            n_req = k0
        return n_req

    def _fit_algorithm_pca(self, f1, f2):
        """cross/base_model_cross_set.py:307-321 + cross/cpcca.py:168-225 on the PCA scores of both fields."""
        import warnings
        ops, p = self.ops, self._params
        if f1.n_samples != f2.n_samples or f1.T != f2.T:
            raise ValueError(
                f"Both data matrices must have the same number of samples but found {f1.n_samples} in the "
                f"first and {f2.n_samples} in the second."
            )
        npm, irr = _pair(p["n_pca_modes"]), _pair(p["pca_init_rank_reduction"])
        U1, s1, V1t = self._pca_stage(f1, npm[0], irr[0])
        U2, s2, V2t = self._pca_stage(f2, npm[1], irr[1])
        n, k = f1.n_samples, p["n_modes"]
        X1, X2 = U1 * s1, U2 * s2                                   # pca.py:121-131: A V = U S
        # fractional whitening (preprocessing/whitener.py:111-133): T = C^((alpha-1)/2) with C = X^T X / n, which in
        # the PCA space is diagonal, s^2 / n; alpha = 1 is the identity (MCA), 0 full whitening (CCA)
        tw = []
        for sv, a in ((s1, self._alpha[0]), (s2, self._alpha[1])):
            lam = sv**2 / n
            t = torch.where(lam > torch.finfo(torch.float64).eps, lam ** ((a - 1.0) / 2.0), torch.zeros_like(lam))
            tw.append(torch.ones_like(lam) if a == 1.0 else t)
        X1u, X2u = X1, X2
        X1, X2 = X1 * tw[0], X2 * tw[1]
        C = ops.dgemm(X1, X2, trans_a=True) / (n - 1)               # cpcca.py:1008-1015, r1 x r2
        rank = min(C.shape)
        by_variance = isinstance(k, float)
        frac = k
        if by_variance:                                            # decomposer.py:88-94
            k = int(rank * 0.3)
            if k < 1:
                warnings.warn("`init_rank_reduction=0.3` is too low resulting in zero components. One component "
                              "will be computed instead.")
                k = 1
        if k > rank:
            raise ValueError(f"n_modes must be less than or equal to the rank of the dataset (rank = {rank}).")
        if p["solver"] not in ("auto", "full", "randomized"):
            raise ValueError(f"Unrecognized solver '{p['solver']}'. Valid options are 'auto', 'full', and 'randomized'.")
        Q1, s, Q2 = self._decompose_small(C, k)
        if by_variance:
            tv = C.var(dim=0, unbiased=True).sum()
            k = self._modes_for_variance(s, frac, k, tv, int(C.shape[0]))
            Q1, s, Q2 = Q1[:, :k].contiguous(), s[:k].contiguous(), Q2[:, :k].contiguous()
        # decomposer.py:219-222: the sign rule reads Q2 in the space it was computed in, the PCA space
        sign = torch.where(Q2.max(0).values.abs() >= Q2.min(0).values.abs(), 1.0, -1.0).to(torch.float64)
        Q1, Q2 = (Q1 * sign).contiguous(), (Q2 * sign).contiguous()
        kp = lpad(k)
        R1, R2 = ops.dgemm(X1, Q1), ops.dgemm(X2, Q2)               # cpcca.py:204-205
        sc1, sc2 = ops.zeros((f1.T, kp)), ops.zeros((f2.T, kp))
        sc1[:, :k] = R1.to(torch.float32)
        sc2[:, :k] = R2.to(torch.float32)
        # components in physical space: un-whiten (Tinv^T Q, whitener.py:201-213), then back from the PCA space
        # (pca.py:161-171): V (Q / t), one k-column product on the space-side block of principal patterns per field
        comps = []
        for ff, Vt, t, Q in ((f1, V1t, tw[0], Q1), (f2, V2t, tw[1], Q2)):
            Mat = (Q * torch.where(t > 0, 1.0 / t, torch.zeros_like(t))[:, None]).contiguous()
            comps.append(ops.apply(Vt, ff.S, int(Q.shape[0]), 1, Mat, k))
        self.k = k
        self._Q1t, self._Q2t, self._sc1, self._sc2, self._s = comps[0], comps[1], sc1, sc2, s
        self._f1, self._f2 = f1, f2
        self.n_pca_modes_ = (int(s1.numel()), int(s2.numel()))
        self._pca_ctx = dict(X1u=X1u, X2u=X2u, Q1=Q1, Q2=Q2, t1=tw[0], t2=tw[1], R1=R1, R2=R2, n=n,
                             U1=U1, U2=U2, V1t=V1t, V2t=V2t, s1=s1, s2=s2)
        self.__dict__.pop("_proj_cache", None)
        Cu = C if self._alpha == (1.0, 1.0) else ops.dgemm(X1u, X2u, trans_a=True) / (n - 1)
        self.data = {
            "singular_values": s, "squared_covariance": s**2,
            "norm1": torch.linalg.norm(R1, dim=0), "norm2": torch.linalg.norm(R2, dim=0),
            "idx_modes_sorted": torch.argsort(s, descending=True),
            # cpcca.py:991-1000: of the UN-whitened cross-covariance (in the PCA space)
            "total_squared_covariance": float((Cu ** 2).sum().item()),
        }
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
        by_variance, frac = isinstance(k, float), k
        if by_variance:                                            # decomposer.py:88-94
            import warnings
            k = int(rank * 0.3)
            if k < 1:
                warnings.warn("`init_rank_reduction=0.3` is too low resulting in zero components. One component "
                              "will be computed instead.")
                k = 1
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
        self._f1, self._f2 = f1, f2
        if by_variance:                                            # decomposer.py:188-216 on the implicit C
            k = self._modes_for_variance(s, frac, k, self._cross_matrix_variance(), f1.n_features)
            s = s[:k].contiguous()
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

    def _cross_matrix_variance(self):
        """What decomposer.py:203 evaluates on the cross-covariance matrix, C.var(feature1, ddof=1).sum(feature2) =
        (sum |C|^2 - S1 sum_j mean_i(C[i,j])^2) / (S1 - 1), without forming C: the column means are
        1^T C / S1 = (A1 1)^T A2 / ((n-1) S1), two one-column streaming passes."""
        ops, comm, f1, f2 = self.ops, self.comm, self._f1, self._f2
        ones = ops.space_side(lpad(1), f1.S, zero=True)
        ones[0] = 1.0
        w = ops.project_T(f1.field, ones, 1, algo=ops.accurate_algo)            # A1 1  (T x 1)
        comm.sum_(w)
        mu = ops.project_S(f2.field, w, 1, algo=ops.accurate_algo)[0].double()  # (A1 1)^T A2  (S2)
        mu = mu / float(f1.n_samples - 1) / float(f1.n_features)
        msq = (mu ** 2).sum()
        comm.sum_(msq)
        tsc = self.data.get("total_squared_covariance") if self.data else None
        if tsc is None:
            tsc = self._total_squared_covariance()
            self._tsc_cache = tsc
        S1 = float(f1.n_features)
        return (tsc - S1 * float(msq.item())) / (S1 - 1.0)

    def _total_squared_covariance(self):
        if getattr(self, "_tsc_cache", None) is not None:
            return self._tsc_cache
        """cpcca.py:991-1000: sum |C|^2 = <X X^T, Y Y^T>_F / (n-1)^2, from the two T x T Gram matrices built 128
        columns at a time with the streaming product.  The Gram matrices are symmetric: for the column block starting
        at sample t0 only the samples t >= t0 are streamed, the strictly lower part counts twice."""
        ops = self.ops
        acc = torch.zeros((), dtype=torch.float64, device=ops.device)
        # each Gram entry is a sum over all features: with >= 2^16 of them the rounding noise of a single TF32 product
        # (operands rounded to nearest, 2e-4 per term) averages out below 1e-6; smaller fields take the 3xTF32 product
        S_all = min(self._f1.S_global, self._f2.S_global)
        algo = getattr(ops, "sum_algo", ops.accurate_algo) if S_all >= 65536 else ops.accurate_algo
        # wide fields: the two sample Gram matrices as tcgen05 GEMMs on a bf16 copy of each preprocessed matrix
        # (csrc/gram_bf16.cu): every entry sums >= 65 536 products whose rounding errors average out
        if S_all >= 65536 and hasattr(ops, "sample_gram") and algo == getattr(ops, "sum_algo", None):
            G1 = ops.sample_gram(self._f1.field)
            G2 = ops.sample_gram(self._f2.field) if G1 is not None else None
            if G2 is not None:
                self.comm.sum_(G1)
                self.comm.sum_(G2)
                P = G1.double() * G2.double()
                del G1, G2
                tot = 2.0 * torch.tril(P, -1).sum() + torch.diagonal(P).sum()
                self._tsc_cache = float(tot.item()) / float(self._f1.n_samples - 1) ** 2
                return self._tsc_cache
        # else: one TF32-rounded copy of each preprocessed matrix, then every Gram block is a plain streaming product
        # of its rows (the field is read ~T/256 times: the copy pays for itself after the second block)
        r1 = r2 = None
        if S_all >= 65536 and hasattr(ops, "materialize") and algo == getattr(ops, "sum_algo", None):
            r1 = ops.materialize(self._f1.field)
            r2 = ops.materialize(self._f2.field) if r1 is not None else None
            if r2 is None:
                r1 = None
        for ((t0, t1), g1), (_, g2) in zip(self._gram_block_columns(self._f1, algo, r1),
                                           self._gram_block_columns(self._f2, algo, r2)):
            w = t1 - t0
            prod = g1.double() * g2.double()
            acc += prod[:w].sum() + 2.0 * prod[w:].sum()
        self._tsc_cache = float(acc.item()) / float(self._f1.n_samples - 1) ** 2
        return self._tsc_cache

    # ------------------------------------------------------------------ accessors
    def components(self, normalized=True):
        """cpcca.py:308-316 behind base_model_cross_set.py:465-493: normalized=False scales every mode by norm1 /
        norm2 (the scaling commutes with the linear un-whitening and PCA back-projection)."""
        Q1t, Q2t = self._Q1t, self._Q2t
        if not normalized:
            Q1t = Q1t[: self.k] * self.data["norm1"].to(torch.float32)[:, None]
            Q2t = Q2t[: self.k] * self.data["norm2"].to(torch.float32)[:, None]
        c1 = self.preprocessor1.components_to_nd(Q1t, self.k, "components1")
        c2 = self.preprocessor2.components_to_nd(Q2t, self.k, "components2")
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
        """cpcca.py:418-512: 1 - ||(X - X_m)^T (Y - Y_m)||_F^2 / (n-1)^2 / sum |C|^2 with X_m, Y_m the un-whitened
        rank-one reconstructions of mode m, negative values set to zero.  For MCA (alpha = 1) that is s_m^2 / sum |C|^2,
        which is what the implicit (use_pca=False) path returns."""
        ctx = getattr(self, "_pca_ctx", None)
        tsc = self.total_squared_covariance()
        if ctx is None or self._alpha == (1.0, 1.0):
            return self._mode_array(self.data["squared_covariance"] / tsc, "squared_covariance_fraction")
        inv1 = torch.where(ctx["t1"] > 0, 1.0 / ctx["t1"], torch.zeros_like(ctx["t1"]))
        inv2 = torch.where(ctx["t2"] > 0, 1.0 / ctx["t2"], torch.zeros_like(ctx["t2"]))
        out = []
        for m in range(self.k):
            X1r = torch.outer(ctx["R1"][:, m], ctx["Q1"][:, m] * inv1)   # whitener.py:176-188 on the reconstruction
            X2r = torch.outer(ctx["R2"][:, m], ctx["Q2"][:, m] * inv2)
            res = torch.linalg.norm((ctx["X1u"] - X1r).t() @ (ctx["X2u"] - X2r) / (ctx["n"] - 1)) ** 2
            out.append(1.0 - res / tsc)
        scf = torch.clamp(torch.stack(out), min=0.0)
        return self._mode_array(scf, "squared_covariance_fraction")

    def covariance_fraction_CD95(self):
        """cross/mca.py:125-215 (Cheng & Dunkerton 1995): sigma_i / sum_i sigma_i over the retained modes, with the
        reference's warning when the estimate still moves with the number of modes."""
        import warnings
        cov = torch.sqrt(self.data["squared_covariance"])
        cf = cov[0] / torch.cumsum(cov, 0)
        if cf.numel() > 1 and float(cf[-2] - cf[-1]) > 0.001:
            warnings.warn("The curent estimate of CF is sensitive to the number of modes retained. Please increase "
                          "`n_modes` for a better estimate.")
        return self._mode_array(cov / cov.sum(), "covariance_fraction")

    # ------------------------------------------------------------------ score statistics (cpcca.py:331-416)
    def _valid_scores(self):
        vs = self._f1.valid_sample if self._f1.n_samples < self._f1.T else None
        r1, r2 = self._sc1[:, : self.k].double(), self._sc2[:, : self.k].double()
        return (r1, r2) if vs is None else (r1[vs], r2[vs])

    @staticmethod
    def _corr(A, B, diagonal=False):
        """cpcca.py:910-985 with method="correlation": columns divided by their (population) standard deviation,
        then A^T B / (n - 1); centred data assumed."""
        A = A / A.std(0, unbiased=False)
        B = B / B.std(0, unbiased=False)
        if diagonal:
            return (A * B).sum(0) / (A.shape[0] - 1)
        return A.t() @ B / (A.shape[0] - 1)

    # ------------------------------------------------------------------ transform / predict / inverse_transform
    def _projection_patterns(self, i):
        """Space-side (kp x S) patterns P with scores = A P: the singular vectors themselves for the implicit MCA;
        with the PCA stage V diag(t) Q (PCA projection, whitening, singular vectors — pca.py:121-131,
        whitener.py:135-145, cpcca.py:233-252) = A^T U diag(t / s) Q, built with one streaming pass and cached."""
        ctx = getattr(self, "_pca_ctx", None)
        if ctx is None:
            return (self._Q1t, self._Q2t)[i]
        cache = self.__dict__.setdefault("_proj_cache", {})
        if i not in cache:
            ff = (self._f1, self._f2)[i]
            Vt, t, Q = ctx["V1t" if i == 0 else "V2t"], ctx["t1" if i == 0 else "t2"], ctx["Q1" if i == 0 else "Q2"]
            cache[i] = self.ops.apply(Vt, ff.S, int(Q.shape[0]), 1, (Q * t[:, None]).contiguous(), self.k)
        return cache[i]

    def _transform_one(self, i, data, normalized):
        pp = (self.preprocessor1, self.preprocessor2)[i]
        new, sample_shape, sample_coords, valid_sample = pp.transform(data)
        Z = self.ops.project_T(new, self._projection_patterns(i), self.k, algo=self.ops.accurate_algo)
        self.comm.sum_(Z)
        if normalized:
            Z[:, : self.k] /= self.data["norm1" if i == 0 else "norm2"].to(torch.float32)[None, :]
        return Z, (pp, sample_shape, sample_coords, valid_sample)

    def transform(self, X=None, Y=None, normalized=False):
        """cross/base_model_cross_set.py:323-392 + cpcca.py:227-252: scores of new data of either field."""
        if X is None and Y is None:
            raise ValueError("Either X or Y must be given.")
        out = []
        for i, d in enumerate((X, Y)):
            if d is None:
                continue
            L.validate_input_type(d)
            Z, (pp, shp, crd, vs) = self._transform_one(i, d, normalized)
            out.append(pp.scores_to_nd(Z, self.k, "scores1" if i == 0 else "scores2", shp, crd, vs))
        return out[0] if len(out) == 1 else tuple(out)

    def predict(self, X):
        """cpcca.py:281-306: pseudo scores of Y from new X: (X P_x) G with G = R_x^H R_y / |R_x|^2."""
        L.validate_input_type(X)
        Z, (pp, shp, crd, vs) = self._transform_one(0, X, False)
        r1, r2 = self._valid_scores()
        G = (r1.t() @ r2) / (torch.linalg.norm(r1, dim=0) ** 2)[:, None]
        Zp = self.ops.zeros(tuple(Z.shape))
        Zp[:, : self.k] = (Z[:, : self.k].double() @ G).to(torch.float32)
        return pp.scores_to_nd(Zp, self.k, "pseudo_scores_Y", shp, crd, vs)

    def inverse_transform(self, X=None, Y=None):
        """cpcca.py:254-279 + the back-transforms of base_model_cross_set.py:394-463: scores . components^H (the
        un-whitened physical-space components), then un-scaled; ``X`` / ``Y`` are score arrays with a 'mode'
        dimension whose coordinate selects the modes."""
        if X is None and Y is None:
            raise ValueError("Either X or Y must be given.")
        out = []
        for i, scores in enumerate((X, Y)):
            if scores is None:
                continue
            data, dims, coords, _ = L.unpack(scores)
            sc = torch.as_tensor(np.asarray(data) if not isinstance(data, torch.Tensor) else data)
            sc = sc.to(self.ops.device, torch.float32)
            if "mode" not in dims:
                sc, dims = sc.unsqueeze(-1), tuple(dims) + ("mode",)
            sc = sc.movedim(dims.index("mode"), -1)
            sample_dims = tuple(d for d in dims if d != "mode")
            sample_shape = tuple(sc.shape[:-1])
            modes = L.mode_indices(coords, int(sc.shape[-1]), self.k)
            pp = (self.preprocessor1, self.preprocessor2)[i]
            if sample_dims != pp.sample_dims:
                raise ValueError(f"scores have sample dimensions {sample_dims}, the model was fitted with {pp.sample_dims}")
            rec = self.ops.reconstruct(pp.fitted.field, sc.reshape(-1, sc.shape[-1]), (self._Q1t, self._Q2t)[i], modes)
            out.append(pp.data_to_nd(rec, sample_shape, {d: coords[d] for d in sample_dims if d in coords}))
        return out[0] if len(out) == 1 else tuple(out)

    # ------------------------------------------------------------------ homogeneous / heterogeneous patterns
    def _data_score_correlation(self, i, R):
        """Pearson correlation (utils/optional/statistics.py:51-76) between every feature of field i — in physical
        space; with the PCA stage that is the rank-r reconstruction the reference back-transforms (cpcca.py:768-775) —
        and the k score series R (n x k fp64, valid samples).  Returns (k x S) on the device."""
        ops, comm, k = self.ops, self.comm, self.k
        ff = (self._f1, self._f2)[i]
        n = ff.n_samples
        vs = ff.valid_sample if n < ff.T else None
        Rn = R / R.std(0, unbiased=False)
        W = ops.zeros((ff.T, lpad(k)))
        ctx = getattr(self, "_pca_ctx", None)
        if ctx is None:
            # A^T R / (n std(A)): one streaming pass; std(A_s) = raw std * |dscale| (the field is centred)
            if vs is None:
                W[:, :k] = Rn.to(torch.float32)
            else:
                W[vs, :k] = Rn.to(torch.float32)
            num = ops.project_S(ff.field, W, k, algo=ops.accurate_algo)
            std = (ff.std * ff.field.dscale.abs()).double()
        else:
            # the reconstruction is A_r = U S V^T:  A_r^T R = V S (U^T R),  sum_t A_r[t,s]^2 = sum_j (s_j V[s,j])^2
            U, Vt, sv = ctx["U1" if i == 0 else "U2"], ctx["V1t" if i == 0 else "V2t"], ctx["s1" if i == 0 else "s2"]
            Uv = U if vs is None else U[vs]
            r = int(sv.numel())
            Mat = (sv[:, None] * ops.dgemm(Uv, Rn.contiguous(), trans_a=True)).contiguous()
            num = ops.apply(Vt, ff.S, r, 1, Mat, k)
            ss = ((Vt[:r].double() * sv[:, None]) ** 2).sum(0)
            std = torch.sqrt(ss / n)
        return num[:k].double() / (n * std)[None, :]

    def _patterns(self, R1, R2, names, correction, alpha):
        if correction is not None:
            raise NotImplementedError("multiple-test correction of the p-values needs statsmodels, which this build "
                                      "does not carry; pass correction=None")
        import scipy.stats

        n = self._f1.n_samples
        dist = scipy.stats.beta(n / 2 - 1, n / 2 - 1, loc=-1, scale=2)  # statistics.py:92-106
        pats, pvals = [], []
        for i, (R, pp, name) in enumerate(zip((R1, R2), (self.preprocessor1, self.preprocessor2), names)):
            corr = self._data_score_correlation(i, R)
            buf = self.ops.space_side(lpad(self.k), int(corr.shape[1]), zero=True)
            buf[: self.k] = corr.to(torch.float32)
            pats.append(pp.components_to_nd(buf, self.k, name))
            pv = 2.0 * dist.cdf(-np.abs(corr.cpu().numpy()))
            buf[: self.k] = torch.as_tensor(pv, dtype=torch.float32, device=buf.device)
            pvals.append(pp.components_to_nd(buf, self.k, "pvalues_of_" + name))
        return tuple(pats), tuple(pvals)

    def homogeneous_patterns(self, correction=None, alpha=0.05):
        """cpcca.py:726-811: correlation of each field with its OWN scores, and two-sided p-values."""
        r1, r2 = self._valid_scores()
        return self._patterns(r1, r2, ("left_homogeneous_patterns", "right_homogeneous_patterns"), correction, alpha)

    def heterogeneous_patterns(self, correction=None, alpha=0.05):
        """cpcca.py:813-898: correlation of each field with the scores of the OTHER field."""
        r1, r2 = self._valid_scores()
        return self._patterns(r2, r1, ("left_heterogeneous_patterns", "right_heterogeneous_patterns"), correction,
                              alpha)

    def cross_correlation_coefficients(self):
        r1, r2 = self._valid_scores()
        return self._mode_array(self._corr(r1, r2, diagonal=True), "cross_correlation_coefficients")

    def correlation_coefficients_X(self):
        r1, _ = self._valid_scores()
        return self._corr(r1, r1).cpu().numpy()

    def correlation_coefficients_Y(self):
        _, r2 = self._valid_scores()
        return self._corr(r2, r2).cpu().numpy()

    def get_params(self):
        return dict(self._params)


class CPCCA(MCA):
    """Continuum Power CCA (cross/cpcca.py:25-160): MCA on fractionally whitened fields, alpha = 1 MCA, 0 CCA,
    (0, 1) RDA.  Built on the PCA scores (``use_pca=True``, the default)."""

    def __init__(self, n_modes=2, alpha=0.2, standardize=False, use_coslat=False, use_pca=True, n_pca_modes=0.999,
                 pca_init_rank_reduction=0.3, check_nans=True, compute=True, sample_name="sample",
                 feature_name="feature", solver="auto", random_state=None, solver_kwargs=None, **kw):
        super().__init__(n_modes=n_modes, standardize=standardize, use_coslat=use_coslat, check_nans=check_nans,
                         use_pca=use_pca, n_pca_modes=n_pca_modes, pca_init_rank_reduction=pca_init_rank_reduction,
                         compute=compute, sample_name=sample_name, feature_name=feature_name, solver=solver,
                         random_state=random_state, solver_kwargs=solver_kwargs, **kw)
        a = tuple(float(v) for v in _pair(alpha))
        if any(v < 0.0 or v > 1.0 for v in a):
            raise ValueError("alpha must be in the range [0, 1]")  # whitener.py:41-42
        self._alpha = a
        self._params["alpha"] = alpha
        self.attrs["model"] = "Continuum Power CCA"


class CCA(CPCCA):
    """Canonical Correlation Analysis (cross/cca.py:8-117): CPCCA with alpha = (0, 0)."""

    def __init__(self, n_modes=2, standardize=False, use_coslat=False, check_nans=True, use_pca=True,
                 n_pca_modes=0.999, pca_init_rank_reduction=0.3, compute=True, sample_name="sample",
                 feature_name="feature", solver="auto", random_state=None, solver_kwargs=None, **kw):
        super().__init__(n_modes=n_modes, alpha=[0.0, 0.0], standardize=standardize, use_coslat=use_coslat,
                         use_pca=use_pca, n_pca_modes=n_pca_modes, pca_init_rank_reduction=pca_init_rank_reduction,
                         check_nans=check_nans, compute=compute, sample_name=sample_name, feature_name=feature_name,
                         solver=solver, random_state=random_state, solver_kwargs=solver_kwargs, **kw)
        self._params.pop("alpha")
        self.attrs["model"] = "Canonical Correlation Analysis"


class RDA(CPCCA):
    """Redundancy Analysis (cross/rda.py:8-120): CPCCA with alpha = (0, 1)."""

    def __init__(self, n_modes=2, standardize=False, use_coslat=False, check_nans=True, use_pca=True,
                 n_pca_modes=0.999, pca_init_rank_reduction=0.3, compute=True, sample_name="sample",
                 feature_name="feature", solver="auto", random_state=None, solver_kwargs=None, **kw):
        super().__init__(n_modes=n_modes, alpha=[0.0, 1.0], standardize=standardize, use_coslat=use_coslat,
                         use_pca=use_pca, n_pca_modes=n_pca_modes, pca_init_rank_reduction=pca_init_rank_reduction,
                         check_nans=check_nans, compute=compute, sample_name=sample_name, feature_name=feature_name,
                         solver=solver, random_state=random_state, solver_kwargs=solver_kwargs, **kw)
        self._params.pop("alpha")
        self.attrs["model"] = "Redundancy Analysis"
