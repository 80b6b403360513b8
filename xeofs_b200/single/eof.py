"""EOF analysis on B200 — drop-in for ``xeofs.single.EOF`` (single/eof.py:14-240) behind the same constructor
and ``fit(X, dim, weights)`` / accessor API (single/base_model_single_set.py:58-161, 180-336).

The whole fit runs on the device through libxeofs_b200.so; there is no CPU fallback.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from .. import _engine as E
from .. import _labels as L
from .._cuda_ops import CudaOps
from .._lib import lpad
from .._preprocessor import MultiPreprocessor, Preprocessor


class EOF:
    """Same parameters as the reference (single/eof.py:54-83).  ``compute`` is accepted for signature
    parity (the device path is always eager).  Extra, B200-specific keywords: ``device``, ``algo``
    ("auto" | "simt" | "tf32x1" | "tf32x3": arithmetic of the streaming products) and ``distributed``
    (True: the feature axis of X is this rank's shard of a torch.distributed job)."""

    def __init__(self, n_modes=2, center=True, standardize=False, use_coslat=False, check_nans=True,
                 sample_name="sample", feature_name="feature", compute=True, random_state=None,
                 solver="auto", solver_kwargs=None, *, device=None, algo="auto", distributed=False, ops=None):
        self.n_modes = n_modes
        self._params = dict(n_modes=n_modes, center=center, standardize=standardize, use_coslat=use_coslat,
                            check_nans=check_nans, sample_name=sample_name, feature_name=feature_name,
                            compute=compute, random_state=random_state, solver=solver,
                            solver_kwargs=dict(solver_kwargs or {}))
        self.attrs = {"model": "EOF analysis", "backend": "xeofs_b200"}
        self.ops = ops if ops is not None else CudaOps(device=device, algo=algo)
        self.comm = E.Comm() if distributed else E.NO_COMM
        self._pp_kw = dict(with_center=center, with_std=standardize, with_coslat=use_coslat, check_nans=check_nans,
                           sample_name=sample_name, feature_name=feature_name, comm=self.comm)
        self.preprocessor = Preprocessor(self.ops, **self._pp_kw)
        self.data = {}

    # ------------------------------------------------------------------ fit
    def fit(self, X, dim, weights=None):
        if isinstance(X, (list, tuple)):  # DataList: one array per variable, concatenated along the feature axis
            for a in X:
                L.validate_input_type(a)
            for w in (weights if isinstance(weights, (list, tuple)) else []):
                if w is not None:
                    L.validate_input_type(w)
            self.preprocessor = MultiPreprocessor(self.ops, **self._pp_kw)
        else:
            L.validate_input_type(X)
            if weights is not None:
                L.validate_input_type(weights)
            if isinstance(self.preprocessor, MultiPreprocessor):
                self.preprocessor = Preprocessor(self.ops, **self._pp_kw)
        self._predrawn = None
        ff = self.preprocessor.fit_transform(X, dim, weights, first=self._first_sketch)
        self._fit_algorithm(ff)
        return self

    def _first_sketch(self, T, S):
        """The Gaussian sketch of the range finder, drawn before the statistics are known so that the statistics pass
        can compute A^T Omega from the same read of the field.  Possible when the range finder starts on the time
        side (n_samples < n_features), with every sample present (checked afterwards) and l = k + oversamples."""
        p = self._params
        k = p["n_modes"]
        if not isinstance(k, (int, np.integer)) or p["solver"] == "full":
            return None
        kw = p["solver_kwargs"]
        l = k + kw.get("n_oversamples", 10)
        n_iter = kw.get("n_iter", "auto")
        S_glob = S * self.comm.world  # a lower bound is enough here
        if not (T < S and l <= T and l <= E.MAX_L and (n_iter == "auto" or int(n_iter) >= 1)):
            return None
        if p["solver"] == "auto" and max(T, S_glob) < 500 and k > int(0.8 * min(T, S_glob)):
            return None  # exact policy (decomposer.py:112-131)
        if isinstance(p["random_state"], np.random.RandomState):
            return None
        self._predrawn = E.draw_sketch(p["random_state"], T, l)
        lp = lpad(l)
        W = np.zeros((T, lp), dtype=np.float32)
        # (an integer seed: the memoised fp32 image of the same draw; no seed: the draw just made)
        W[:, :l] = E.draw_sketch(p["random_state"], T, l, f32=True) if p["random_state"] is not None else self._predrawn
        return self.ops.to_device(W), l

    @staticmethod
    def _usable_first(ff, op, k, n_over):
        l = min(k + n_over, *op.shape)
        ok = ff.first_product is not None and op.transposed and ff.first_l == l and ff.n_samples == ff.T
        return ff.first_product if ok else None

    def _shard_offset(self, ff):
        if not self.comm.active:
            return 0
        sizes = torch.zeros(self.comm.world, dtype=torch.int64, device=self.ops.device)
        sizes[self.comm.rank] = ff.S
        self.comm.sum_(sizes)
        return int(sizes[: self.comm.rank].sum().item())

    def _fit_algorithm(self, ff):
        """single/eof.py:85-118 with linalg/decomposer.py:76-226 inlined."""
        ops, comm, p = self.ops, self.comm, self._params
        n, S = ff.n_samples, ff.n_features
        rank = min(n, S)
        n_modes = p["n_modes"]
        by_variance = isinstance(n_modes, float)
        k = n_modes
        if by_variance:  # decomposer.py:88-94
            k = int(rank * 0.3)
            if k < 1:
                warnings.warn("`init_rank_reduction` is too low resulting in zero components.")
                k = 1
        if k > rank:  # decomposer.py:97-100
            raise ValueError(f"n_modes must be less than or equal to the rank of the dataset (rank = {rank}).")
        solver, kw = p["solver"], dict(p["solver_kwargs"])
        if solver == "auto":  # decomposer.py:112-131
            use_exact = max(n, S) < 500 and k > int(0.8 * rank)
        elif solver == "full":
            use_exact = True
        elif solver == "randomized":
            use_exact = False
        else:
            raise ValueError(f"Unrecognized solver '{solver}'. Valid options are 'auto', 'full', and 'randomized'.")
        op = E.FieldOperator(ops, ff, comm, algo=None)
        if use_exact:
            # a full-width sketch spans the whole row space: the range finder is then exact
            n_over, n_iter = rank - k, 2
        else:
            n_over, n_iter = kw.get("n_oversamples", 10), kw.get("n_iter", "auto")
        Ur, s, Vc, infos = E.randomized_svd(ops, op, k, n_oversamples=n_over, n_iter=n_iter,
                                            random_state=p["random_state"], comm=comm,
                                            predrawn=getattr(self, "_predrawn", None),
                                            first_product=self._usable_first(ff, op, k, n_over))
        E.check_infos(infos)
        # un-transpose: A = U s V^T with V on the space side
        Vt, Ut = (Ur, Vc) if op.transposed else (Vc, Ur)
        if by_variance:  # decomposer.py:188-216
            expvar_ratio = (s**2 / (n - 1) / ff.total_variance).cpu().numpy()
            cum = expvar_ratio.cumsum()
            n_req = k - int((cum >= n_modes).sum()) + 1
            if n_req > k:
                warnings.warn("Dataset has fewer components than n_modes; consider increasing init_rank_reduction.")
                n_req = k
            k = n_req
            s = s[:k]
        sign = E.sign_flip(ops, Vt, k, ff.S, comm)  # decomposer.py:219-222
        ops.finish_components(Vt, k, ff.S, sign, None)
        scores = Ut[:, : lpad(k)].clone()
        scores[:, :k] *= (s.to(torch.float32) * sign)[None, :]  # eof.py:101, sign rule on U
        self.k = k
        self._Vt, self._scores, self._s = Vt, scores, s
        self.data = {
            "norms": s, "explained_variance": s**2 / (n - 1),  # eof.py:105-106
            "total_variance": ff.total_variance,
        }
        return self

    # ------------------------------------------------------------------ transform / inverse_transform
    def transform(self, data, normalized=False):
        """eof.py:123-132 via base_model_single_set.py:180-203: ((new - mean)/std*w)[:, valid] . V."""
        if not isinstance(data, (list, tuple)):
            L.validate_input_type(data)
        new, sample_shape, sample_coords, valid_sample = self.preprocessor.transform(data)
        Z = self.ops.project_T(new, self._Vt, self.k, algo=self.ops.accurate_algo)
        self.comm.sum_(Z)
        if normalized:
            Z[:, : self.k] /= self._s.to(torch.float32)[None, :]
        return self.preprocessor.scores_to_nd(Z, self.k, "scores", sample_shape, sample_coords, valid_sample)

    def fit_transform(self, data, dim, weights=None, **kwargs):
        return self.fit(data, dim, weights).transform(data, **kwargs)

    def inverse_transform(self, scores, normalized=False):
        """eof.py:134-156 + scaler.py:165-190: (scores . V^H) / weights / coslat * std + mean, NaN at the
        dropped features.  ``scores`` carries a 'mode' dimension whose coordinate selects the modes."""
        data, dims, coords, _ = L.unpack(scores)
        sc = torch.as_tensor(np.asarray(data) if not isinstance(data, torch.Tensor) else data)
        sc = sc.to(self.ops.device, torch.float32)
        if "mode" not in dims:
            sc, dims = sc.unsqueeze(-1), tuple(dims) + ("mode",)
        sc = sc.movedim(dims.index("mode"), -1)
        sample_dims = tuple(d for d in dims if d != "mode")
        sample_shape = tuple(sc.shape[:-1])
        modes = L.mode_indices(coords, int(sc.shape[-1]), self.k)
        sc2 = sc.reshape(-1, sc.shape[-1])
        if normalized:
            sc2 = sc2 * self._s.to(torch.float32)[torch.as_tensor(modes, device=sc2.device)][None, :]
        rec = self.ops.reconstruct(self.preprocessor.fitted.field, sc2, self._Vt, modes)
        pp = self.preprocessor
        if sample_dims != pp.sample_dims:
            raise ValueError(f"scores have sample dimensions {sample_dims}, the model was fitted with {pp.sample_dims}")
        return pp.data_to_nd(rec, sample_shape, {d: coords[d] for d in sample_dims if d in coords})

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
        k = self.k
        return L.wrap(t.cpu().numpy(), ("mode",), {"mode": np.arange(1, k + 1)}, name, self.preprocessor.as_xarray)

    def singular_values(self):
        return self._mode_array(self.data["norms"], "norms")

    def explained_variance(self):
        return self._mode_array(self.data["explained_variance"], "explained_variance")

    def explained_variance_ratio(self):
        return self._mode_array(self.data["explained_variance"] / self.data["total_variance"], "explained_variance_ratio")

    def total_variance(self):
        return self.data["total_variance"]

    def get_params(self):
        return dict(self._params)
