"""Host-side mirror of xeofs.preprocessing.Preprocessor for ONE DataArray input
(preprocessing/preprocessor.py:119-142, 191-354): Scaler -> Stacker -> Sanitizer, with the arithmetic on the
device (fit_field) and only the label bookkeeping here.

The reference materialises the scaled 2D matrix; here the raw field stays as it is in HBM and the fitted
per-feature vectors (pivot / dscale / ccorr) are applied inside the streaming kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _labels as L
from ._engine import NO_COMM, fit_field


def _align_rows(t2):
    """The TMA descriptors of the tensor-core kernels need 16-byte aligned rows: a field whose row length is not a
    multiple of 4 floats (BASELINE configs[0]: 25 x 53 = 1325 features) gets ONE padded-pitch copy here instead of
    running every pass on the CUDA-core path."""
    if not t2.is_cuda or (t2.stride(0) % 4 == 0 and t2.data_ptr() % 16 == 0):
        return t2
    T, S = int(t2.shape[0]), int(t2.shape[1])
    buf = torch.empty((T, (S + 31) // 32 * 32), dtype=torch.float32, device=t2.device)
    buf[:, :S].copy_(t2)
    return buf[:, :S]


def _new_field(ops, comm, ff, X2, check_nans):
    """Unseen data behind the fitted scaling vectors + the Sanitizer's checks on it (sanitizer.py:108-122)."""
    from ._cuda_ops import Field

    f = ff.field
    new = Field(X2, f.pivot, f.dscale, f.ccorr, f.valid, f.mean, f.std, None)
    valid_sample = None
    if check_nans:
        st = ops.col_stats(X2)
        new_valid = st["cnt"] > 0
        row_nan = st["row_nan"].to(torch.int64)
        if comm.active:
            comm.sum_(row_nan)
        mism = (new_valid != ff.valid.bool()).any().to(torch.int32)
        if comm.active:
            comm.max_(mism)
        if bool(mism.item()):
            raise ValueError("Input data had NaN features in different locations than the original data.")
        n_invalid = ff.S_global - ff.n_features
        ok = (row_nan == n_invalid) | (row_nan == ff.S_global)
        if not bool(ok.all().item()):
            raise ValueError("Input data contains partial NaN entries, which will cause the the SVD to fail.")
        valid_sample = row_nan < ff.S_global
        if not bool(valid_sample.all().item()):
            new.row_valid = valid_sample.to(torch.uint8)
        else:
            new.no_nan = ff.n_features == ff.S_global  # same NaN pattern as the fitted field: none at all
    return new, valid_sample


class Preprocessor:
    def __init__(self, ops, with_center=True, with_std=False, with_coslat=False, check_nans=True,
                 sample_name="sample", feature_name="feature", comm=NO_COMM):
        self.ops, self.comm = ops, comm
        self.with_center, self.with_std, self.with_coslat = with_center, with_std, with_coslat
        self.check_nans = check_nans
        self.sample_name, self.feature_name = sample_name, feature_name
        self.fitted = None

    # ------------------------------------------------------------------ stacking (stacker.py:157-214)
    def _to_2d(self, data, dims, fit):
        """(sample..., feature...) order, flattened to 2D on the device.  A view when the sample dims lead."""
        if fit:
            self.sample_dims, self.feature_dims = L.split_dims(dims, self._dim)
            self.dims_in = tuple(dims)
        elif tuple(dims) != self.dims_in and set(dims) != set(self.dims_in):
            raise ValueError(f"Data dimensions {dims} do not match the fitted ones {self.dims_in}.")
        order = [dims.index(d) for d in self.sample_dims] + [dims.index(d) for d in self.feature_dims]
        if isinstance(data, torch.Tensor):
            t = data
            if t.dtype != torch.float32:
                t = t.to(torch.float32)
            if t.device != self.ops.device:
                t = t.to(self.ops.device, non_blocking=True)
        else:
            a = np.asarray(data)
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.ops.device)
        if order != list(range(len(dims))):
            t = t.permute(order)
        ns = len(self.sample_dims)
        sample_shape, feature_shape = tuple(t.shape[:ns]), tuple(t.shape[ns:])
        T, S = int(np.prod(sample_shape)), int(np.prod(feature_shape))
        t2 = t.reshape(T, S)  # copies only when the permutation made it non-viewable
        if t2.stride(1) != 1 or (T > 1 and t2.stride(0) < S):
            t2 = t2.contiguous()
        return _align_rows(t2), sample_shape, feature_shape

    # ------------------------------------------------------------------ fit
    def _prepare(self, X, sample_dims, weights=None):
        """Labels + stacking + the per-feature weight vector (coslat * weights, fp64) of one array: a host array, or —
        coslat weights alone — a device vector broadcast there from the latitude vector (at 4 M features the host
        broadcast and its copy to the device cost 20-50 ms per fit)."""
        data, dims, coords, self.as_xarray = L.unpack(X)
        self._dim = sample_dims
        X2, self.sample_shape, self.feature_shape = self._to_2d(data, dims, fit=True)
        self.coords = {d: coords[d] for d in dims if d in coords}
        featw = None
        if self.with_coslat and weights is None:
            w, shape = L.sqrt_cos_lat_vector(self.feature_dims, self.coords)
            if tuple(np.broadcast_shapes(tuple(shape), tuple(self.feature_shape))) != tuple(self.feature_shape):
                raise ValueError(f"latitude coordinate of length {w.size} does not match the feature shape {self.feature_shape}")
            featw = self.ops.to_device(w, torch.float64).reshape(shape).expand(tuple(self.feature_shape)).reshape(-1)
        elif self.with_coslat:
            featw = np.array(L.sqrt_cos_lat_weights(self.feature_dims, self.feature_shape, self.coords), dtype=np.float64)
        if weights is not None:
            wdata, wdims, _, _ = L.unpack(weights)
            w = wdata.detach().cpu().numpy() if isinstance(wdata, torch.Tensor) else np.asarray(wdata)
            extra = [d for d in wdims if d not in self.feature_dims]
            if extra:
                raise ValueError(f"Weights have dimensions {extra} that are not feature dimensions.")
            # broadcast over the feature dims in data order (scaler.py:116, utils/xarray_utils.py:78-100)
            shape = [1] * len(self.feature_dims)
            perm = sorted(range(len(wdims)), key=lambda i: self.feature_dims.index(wdims[i]))
            w = np.transpose(w, perm)
            for i, d in enumerate(sorted(wdims, key=self.feature_dims.index)):
                shape[self.feature_dims.index(d)] = w.shape[i]
            w = np.broadcast_to(w.reshape(shape).astype(np.float64), self.feature_shape)
            featw = w if featw is None else featw * w
        return X2, featw

    def fit_transform(self, X, sample_dims, weights=None, overlap=None, first=None):
        """``first``: callable (T, S) -> (W, l) or None, the sketch for the fused statistics + first product pass."""
        X2, featw = self._prepare(X, sample_dims, weights)
        if isinstance(featw, torch.Tensor):
            self.featw_host, featw_dev = None, featw
        else:
            self.featw_host = None if featw is None else np.ascontiguousarray(featw.reshape(-1))
            featw_dev = None if featw is None else self.ops.to_device(self.featw_host, torch.float64)
        self.fitted = fit_field(self.ops, X2, featw_dev, center=self.with_center, standardize=self.with_std,
                                check_nans=self.check_nans, comm=self.comm, overlap=overlap,
                                first=first(int(X2.shape[0]), int(X2.shape[1])) if first is not None else None)
        return self.fitted

    # ------------------------------------------------------------------ transform of unseen data
    def transform(self, X):
        """preprocessor.py:232-259: re-apply the fitted scaling; sanitizer.py:108-113 NaN-pattern check."""
        if self.fitted is None:
            raise ValueError("The preprocessor has not been fitted.")
        data, dims, coords, _ = L.unpack(X)
        X2, sample_shape, feature_shape = self._to_2d(data, dims, fit=False)
        if feature_shape != self.feature_shape:
            raise ValueError(f"Feature shape {feature_shape} differs from the fitted one {self.feature_shape}.")
        new, valid_sample = _new_field(self.ops, self.comm, self.fitted, X2, self.check_nans)
        sample_coords = {d: coords[d] for d in self.sample_dims if d in coords}
        return new, sample_shape, sample_coords, valid_sample

    # ------------------------------------------------------------------ inverse transforms (labels only)
    def components_to_nd(self, Vt, k, name="components", nan_invalid=True):
        """(kp x S) device, mode-major -> DataArray feature_dims + ('mode',); NaN at dropped features
        (sanitizer.py:128-153 reindex)."""
        V = Vt[:k].detach().clone()
        if nan_invalid:
            V[:, ~self.fitted.valid.bool()] = float("nan")
        arr = V.t().reshape(*self.feature_shape, k).cpu().numpy()
        dims = self.feature_dims + ("mode",)
        coords = dict(self.coords)
        coords["mode"] = np.arange(1, k + 1)
        return L.wrap(arr, dims, coords, name, self.as_xarray)

    def scores_to_nd(self, Sc, k, name="scores", sample_shape=None, sample_coords=None, valid_sample=None):
        sample_shape = self.sample_shape if sample_shape is None else sample_shape
        Sc = Sc[:, :k].detach().clone()
        vs = self.fitted.valid_sample if (valid_sample is None and sample_coords is None) else valid_sample
        if vs is not None:
            Sc[~vs] = float("nan")
        arr = Sc.reshape(*sample_shape, k).cpu().numpy()
        dims = self.sample_dims + ("mode",)
        coords = dict(self.coords if sample_coords is None else sample_coords)
        coords["mode"] = np.arange(1, k + 1)
        return L.wrap(arr, dims, coords, name, self.as_xarray)

    def data_to_nd(self, A2d, sample_shape, sample_coords=None, name="reconstructed_data"):
        arr = A2d.reshape(*sample_shape, *self.feature_shape)
        nd_dims = self.sample_dims + self.feature_dims
        coords = dict(self.coords)
        if sample_coords is not None:
            for d in self.sample_dims:
                coords.pop(d, None)
            coords.update(sample_coords)
        # back to the input dim order (stacker.py:216-275)
        perm = [nd_dims.index(d) for d in self.dims_in]
        arr = arr.permute(perm) if isinstance(arr, torch.Tensor) else np.transpose(arr, perm)
        if isinstance(arr, torch.Tensor):
            arr = arr.cpu().numpy()
        return L.wrap(arr, self.dims_in, coords, name, self.as_xarray)


class _PartFit:
    """What a part's label bookkeeping reads from the fit: its slice of the feature mask and the sample mask."""

    def __init__(self, valid, valid_sample):
        self.valid, self.valid_sample = valid, valid_sample


class MultiPreprocessor:
    """A LIST of arrays as input (the reference's DataList): every array is scaled and weighted on its own, stacked,
    and the 2D matrices are concatenated along the feature axis (preprocessing/preprocessor.py:208-228,
    concatenator.py:58-81).  Per-feature statistics make "on its own" and "after concatenation" the same arithmetic,
    so the concatenated field goes through the same single pass as one array; feature-shaped results come back as one
    array per input.  All arrays must share the sample dimensions and their lengths."""

    def __init__(self, ops, **kw):
        self.ops, self.kw = ops, kw
        self.comm = kw.get("comm", NO_COMM)
        self.check_nans = kw.get("check_nans", True)
        self.parts, self.fitted = [], None

    def fit_transform(self, Xs, sample_dims, weights=None, overlap=None, first=None):
        if self.comm.active:
            raise NotImplementedError("list inputs are not supported together with distributed=True")
        ws = list(weights) if isinstance(weights, (list, tuple)) else [weights] * len(Xs)
        if len(ws) != len(Xs):
            raise ValueError("weights must be a list with one entry (or None) per input array")
        mats, fws = [], []
        self.parts = []
        for Xi, wi in zip(Xs, ws):
            p = Preprocessor(self.ops, **self.kw)
            X2, fw = p._prepare(Xi, sample_dims, wi)
            self.parts.append(p)
            mats.append(X2)
            fws.append(fw)
        shapes = {p.sample_shape for p in self.parts}
        if len(shapes) != 1:
            raise ValueError(f"All arrays must have the same sample dimensions; found shapes {sorted(shapes)}.")
        sizes = [int(m.shape[1]) for m in mats]
        self.offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
        X2 = _align_rows(torch.cat(mats, dim=1))
        del mats
        featw = None
        if any(fw is not None for fw in fws):
            fws = [fw.detach().cpu().numpy() if isinstance(fw, torch.Tensor) else fw for fw in fws]
            featw = np.concatenate([np.ones(n) if fw is None else np.ascontiguousarray(fw.reshape(-1))
                                    for fw, n in zip(fws, sizes)])
        self.featw_host = featw
        featw_dev = None if featw is None else self.ops.to_device(featw, torch.float64)
        p0 = self.parts[0]
        self.fitted = fit_field(self.ops, X2, featw_dev, center=p0.with_center, standardize=p0.with_std,
                                check_nans=self.check_nans, comm=self.comm, overlap=overlap,
                                first=first(int(X2.shape[0]), int(X2.shape[1])) if first is not None else None)
        for p, a, b in zip(self.parts, self.offsets[:-1], self.offsets[1:]):
            p.fitted = _PartFit(self.fitted.valid[a:b], self.fitted.valid_sample)
        self.as_xarray = p0.as_xarray
        self.sample_shape, self.sample_dims = p0.sample_shape, p0.sample_dims
        return self.fitted

    def transform(self, Xs):
        if self.fitted is None:
            raise ValueError("The preprocessor has not been fitted.")
        if not isinstance(Xs, (list, tuple)) or len(Xs) != len(self.parts):
            raise ValueError(f"Expected a list of {len(self.parts)} arrays.")
        mats = []
        for p, Xi in zip(self.parts, Xs):
            data, dims, coords, _ = L.unpack(Xi)
            X2, sample_shape, feature_shape = p._to_2d(data, dims, fit=False)
            if feature_shape != p.feature_shape:
                raise ValueError(f"Feature shape {feature_shape} differs from the fitted one {p.feature_shape}.")
            mats.append(X2)
        new, valid_sample = _new_field(self.ops, self.comm, self.fitted, _align_rows(torch.cat(mats, dim=1)), self.check_nans)
        sample_coords = {d: coords[d] for d in self.parts[0].sample_dims if d in coords}
        return new, sample_shape, sample_coords, valid_sample

    def components_to_nd(self, Vt, k, name="components", nan_invalid=True):
        return [p.components_to_nd(Vt[:, a:b], k, name, nan_invalid)
                for p, a, b in zip(self.parts, self.offsets[:-1], self.offsets[1:])]

    def scores_to_nd(self, Sc, k, name="scores", sample_shape=None, sample_coords=None, valid_sample=None):
        return self.parts[0].scores_to_nd(Sc, k, name, sample_shape, sample_coords, valid_sample)

    def data_to_nd(self, A2d, sample_shape, sample_coords=None, name="reconstructed_data"):
        return [p.data_to_nd(A2d[:, a:b], sample_shape, sample_coords, name)
                for p, a, b in zip(self.parts, self.offsets[:-1], self.offsets[1:])]
