"""Thin Python layer over the C-ABI: allocates torch CUDA tensors and passes raw pointers.

Nothing here computes on the host; every method enqueues kernels of libxeofs_b200.so on the current
torch CUDA stream.  ``CudaOps`` is the only implementation the product ships (tests inject a numpy
test double with the same interface to exercise the host logic on CPU boxes).
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import check, lpad, ptr


BLOCK_L = 128  # width of the product kernels' accumulators and of the k-column kernels


class Field:
    """One preprocessed field resident in HBM: raw X (T x S fp32, row-major) plus the per-feature
    vectors that fold Scaler.transform (preprocessing/scaler.py:146-153) into the operand load:
    A[t,s] = (X[t,s] - pivot[s]) * dscale[s] + ccorr[s]."""

    def __init__(self, X, pivot, dscale, ccorr, valid, mean=None, std=None, row_valid=None, no_nan=False):
        self.no_nan = no_nan  # True: no NaN anywhere in X (every feature and every sample valid), known to the caller
        self.want_h16 = False  # set by fit_field: the power iterations may run on an fp16 copy of the matrix
        self.h16 = None        # (copy (T x pitch) int16, ic16, cc16 or None) once a pass has written it
        self.X = X
        self.T, self.S = int(X.shape[0]), int(X.shape[1])
        self.ldx = int(X.stride(0))
        self.pivot, self.dscale, self.ccorr = pivot, dscale, ccorr
        self.valid = valid
        self.mean, self.std = mean, std
        self.row_valid = row_valid  # (T,) uint8, None = every sample valid (sanitizer.py:49-50)


class CudaOps:
    name = "cuda"

    def __init__(self, device=None, algo="auto"):
        if not torch.cuda.is_available():
            raise RuntimeError("xeofs_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback.")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        a = _lib.ALGO_NAMES[algo] if isinstance(algo, str) else int(algo)
        # algo: arithmetic of the power-iteration products; accurate_algo: of the two products the singular
        # values are read from (the final range basis and B = Q^T M)
        # exact_algo: the same two products when the caller has made the small operand TF32-exact (round_tf32_)
        # sum_algo: products whose entries are summed as numbers over ~1e5 features or more (the sample Gram matrices
        # behind MCA's total squared covariance): one TF32 product with operands rounded to nearest, unbiased
        self.sum_algo = _lib.ALGO_TF32X1R if a in (_lib.ALGO_AUTO, _lib.ALGO_TF32X1) else a
        if a == _lib.ALGO_AUTO:
            self.algo, self.accurate_algo, self.exact_algo = _lib.ALGO_AUTO_FAST, _lib.ALGO_AUTO, _lib.ALGO_TF32X2
        elif a == _lib.ALGO_TF32X1:
            self.algo, self.accurate_algo, self.exact_algo = _lib.ALGO_TF32X1, _lib.ALGO_TF32X3, _lib.ALGO_TF32X2
        else:
            self.algo = self.accurate_algo = self.exact_algo = a
        self._ws = None
        self._vws = None
        self._unit = None
        self.use_h16 = os.environ.get("XEOFS_H16", "1") != "0"  # fp16 copy of the matrix for the power iterations
        self.h16_stats_copy = os.environ.get("XEOFS_H16_STATS", "1") != "0"  # written by the statistics pass already
        self.launches = 0  # kernels enqueued through this object (bench.py reports it)
        self.time_products = False  # bench.py: CUDA events around every streaming product on the launch stream
        self._prod_events = []
        self._h16_fitted = set()  # (T, pitch) of fp16 copies that fitted in HBM before

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return _lib.C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def space_side(self, rows, n, zero=False):
        """(rows x n) fp32 view whose row stride is a multiple of 32 elements (128 B-aligned rows)."""
        ld = (int(n) + 31) // 32 * 32
        buf = (torch.zeros if zero else torch.empty)((rows, ld), dtype=torch.float32, device=self.device)
        return buf[:, :n]

    def to_device(self, a, dtype=None):
        t = torch.as_tensor(a)
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=True)

    def workspace(self, T, S, l, algo):
        need = int(self.lib.xeofs_b200_project_workspace_bytes(T, S, l, algo))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=self.device)
        return self._ws

    def _timed(self, name, l, fn):
        if not self.time_products:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(self.device))
        rc = fn()
        e1.record(torch.cuda.current_stream(self.device))
        self._prod_events.append((name, e0, e1, l))
        return rc

    @staticmethod
    def _algo_tag(algo):
        """Suffix of the timing tag: the single-TF32 products (the bulk of a fit) carry none."""
        return {_lib.ALGO_AUTO: "_x3", _lib.ALGO_TF32X3: "_x3", _lib.ALGO_TF32X2: "_x2", _lib.ALGO_SIMT: "_simt",
                _lib.ALGO_TF32X1R: "_x1r"}.get(algo, "")

    def product_times(self):
        """[(kernel name, milliseconds, l)] of the timed products (synchronises)."""
        torch.cuda.synchronize(self.device)
        return [(n, e0.elapsed_time(e1), l) for n, e0, e1, l in self._prod_events]

    # ------------------------------------------------------------------ preprocessing
    def col_stats(self, X):
        return self._timed("col_stats", 0, lambda: self._col_stats(X))

    def _col_stats(self, X):
        T, S = int(X.shape[0]), int(X.shape[1])
        st = {
            "shift": self.empty(S), "sum": self.empty(S, torch.float64), "sumsq": self.empty(S, torch.float64),
            "cnt": self.empty(S, torch.int32), "row_nan": self.empty(T, torch.int32),
        }
        check(self.lib.xeofs_b200_col_stats(ptr(X), T, S, int(X.stride(0)), ptr(st["shift"]), ptr(st["sum"]),
                                            ptr(st["sumsq"]), ptr(st["cnt"]), ptr(st["row_nan"]), self._stream()),
              "col_stats")
        self.launches += 5
        return st

    def scaling_finalize(self, st, featw, center, standardize):
        S = int(st["shift"].shape[0])
        out = {
            "mean": self.empty(S), "std": self.empty(S), "valid": self.empty(S, torch.uint8),
            "pivot": self.empty(S), "dscale": self.empty(S), "ccorr": self.empty(S),
            "scalars": self.empty(4, torch.float64),
        }
        flags = (_lib.F_CENTER if center else 0) | (_lib.F_STANDARDIZE if standardize else 0)
        check(self.lib.xeofs_b200_scaling_finalize(S, ptr(st["shift"]), ptr(st["sum"]), ptr(st["sumsq"]),
                                                   ptr(st["cnt"]), ptr(featw), flags, ptr(out["mean"]),
                                                   ptr(out["std"]), ptr(out["valid"]), ptr(out["pivot"]),
                                                   ptr(out["dscale"]), ptr(out["ccorr"]), ptr(out["scalars"]),
                                                   self._stream()), "scaling_finalize")
        self.launches += 2
        return out

    def stats_project_S(self, X, featw, center, standardize, W, l):
        """col_stats + scaling_finalize + project_S(single TF32) from one read of X.  Returns (row_nan, fin, Yt) or
        None where the fused kernel does not apply."""
        T, S = int(X.shape[0]), int(X.shape[1])
        ldx = int(X.stride(0))
        if not (self.algo in (_lib.ALGO_AUTO_FAST, _lib.ALGO_TF32X1) and ldx % 4 == 0 and X.data_ptr() % 16 == 0
                and bool(self.lib.xeofs_b200_has_tcgen05())):
            return None
        out = {
            "mean": self.empty(S), "std": self.empty(S), "valid": self.empty(S, torch.uint8),
            "pivot": self.empty(S), "dscale": self.empty(S), "ccorr": self.empty(S),
            "scalars": self.empty(4, torch.float64),
        }
        row_nan = self.empty(T, torch.int32)
        lp = lpad(l)
        Yt = self.space_side(lp, S)
        ws = self.workspace(T, S, l, _lib.ALGO_TF32X1)
        flags = (_lib.F_CENTER if center else 0) | (_lib.F_STANDARDIZE if standardize else 0)
        # where it pays and fits, the same pass writes the fp16 copy of the (shifted) field for the power iterations
        copy = None
        if center and self.use_h16 and self.h16_stats_copy and T * S * 4 >= self.h16_min_bytes:
            pitch = (S + 127) // 128 * 128
            # the memory queries cost ~0.5 ms of host time: ask once per shape, afterwards just try the allocation
            fits = (T, pitch) in self._h16_fitted
            if not fits:
                free, _ = torch.cuda.mem_get_info(self.device)
                cached = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
                fits = T * pitch * 2 + (4 << 30) <= free + cached
            if fits:
                try:
                    copy = (torch.empty((T, pitch), dtype=torch.int16, device=self.device), self.empty(S),
                            self.empty(S), self.empty(S))
                    self._h16_fitted.add((T, pitch))
                except torch.cuda.OutOfMemoryError:
                    self._h16_fitted.discard((T, pitch))
                    copy = None
        if copy is not None:
            A16, c0, ic16, cc16 = copy
            check(self._timed("project_S_stats_wcopy", l, lambda: self.lib.xeofs_b200_project_S_stats_h16copy(
                ptr(X), T, S, ldx, ptr(featw), flags, ptr(W), int(W.stride(0)), l, ptr(out["mean"]), ptr(out["std"]),
                ptr(out["valid"]), ptr(out["pivot"]), ptr(out["dscale"]), ptr(out["ccorr"]), ptr(out["scalars"]),
                ptr(row_nan), ptr(Yt), int(Yt.stride(0)), ptr(ws), ws.numel(), ptr(A16), pitch, ptr(c0), ptr(ic16),
                ptr(cc16), self._stream())), "project_S_stats_h16copy")
            out["h16"] = (A16, ic16, cc16)
            self.launches += 6
            return row_nan, out, Yt
        check(self._timed("project_S_stats", l, lambda: self.lib.xeofs_b200_project_S_stats(
            ptr(X), T, S, ldx, ptr(featw), flags, ptr(W), int(W.stride(0)), l, ptr(out["mean"]), ptr(out["std"]),
            ptr(out["valid"]), ptr(out["pivot"]), ptr(out["dscale"]), ptr(out["ccorr"]), ptr(out["scalars"]),
            ptr(row_nan), ptr(Yt), int(Yt.stride(0)), ptr(ws), ws.numel(), self._stream())), "project_S_stats")
        self.launches += 5
        return row_nan, out, Yt

    # ------------------------------------------------------------------ streaming products
    def project_S(self, f: Field, W, l, algo=None, out=None, tag="project_S"):
        """Yt (lp x S) = A^T W,  W time-side (T x lp)."""
        algo = self.algo if algo is None else algo
        lp = lpad(l)
        Yt = out if out is not None else self.space_side(lp, f.S)
        if l > BLOCK_L:  # wider than the kernels' accumulators: one streaming pass per block of 128 columns
            for j0 in range(0, l, BLOCK_L):
                w = min(BLOCK_L, l - j0)
                self.project_S(f, W[:, j0:], w, algo=algo, out=Yt[j0:j0 + lpad(w)], tag=tag)
            return Yt
        if f.h16 is not None and algo in self._fast_algos and tag == "project_S":
            A16, ic16, cc16 = f.h16
            ws = self.workspace(f.T, f.S, l, _lib.ALGO_TF32X1)
            check(self._timed("project_S_h16", l, lambda: self.lib.xeofs_b200_project_S16(
                ptr(A16), f.T, f.S, int(A16.stride(0)), ptr(ic16), ptr(cc16), ptr(W), int(W.stride(0)), l, ptr(Yt),
                int(Yt.stride(0)), ptr(ws), ws.numel(), self._stream())), "project_S16")
            self.launches += 6
            return Yt
        ws = self.workspace(f.T, f.S, l, algo)
        check(self._timed(tag + self._algo_tag(algo), l, lambda: self.lib.xeofs_b200_project_S(
            ptr(f.X), f.T, f.S, f.ldx, ptr(f.pivot), ptr(f.dscale), ptr(f.ccorr), ptr(f.row_valid), ptr(W),
            int(W.stride(0)), l,
            ptr(Yt), int(Yt.stride(0)), ptr(ws), ws.numel(), algo, self._stream())), "project_S")
        self.launches += 2
        return Yt

    def project_T(self, f: Field, Yt, l, algo=None, out=None):
        """Z (T x lp) = A Y,  Y space-side (lp x S)."""
        algo = self.algo if algo is None else algo
        lp = lpad(l)
        Z = out if out is not None else self.empty((f.T, lp))
        if l > BLOCK_L:
            for j0 in range(0, l, BLOCK_L):
                w = min(BLOCK_L, l - j0)
                self.project_T(f, Yt[j0:], w, algo=algo, out=Z[:, j0:j0 + lpad(w)])
            return Z
        if algo in self._fast_algos and (f.h16 is not None or f.want_h16):
            ws = self.workspace(f.T, f.S, l, _lib.ALGO_TF32X1)
            if f.h16 is not None:
                A16, ic16, cc16 = f.h16
                check(self._timed("project_T_h16", l, lambda: self.lib.xeofs_b200_project_T16(
                    ptr(A16), f.T, f.S, int(A16.stride(0)), ptr(ic16), ptr(cc16), ptr(Yt), int(Yt.stride(0)), l, ptr(Z),
                    int(Z.stride(0)), ptr(ws), ws.numel(), self._stream())), "project_T16")
                self.launches += 7
                return Z
            made = self._make_h16(f)
            if made is not None:
                A16, e16, ic16 = made
                check(self._timed("project_T_wcopy", l, lambda: self.lib.xeofs_b200_project_T_h16copy(
                    ptr(f.X), f.T, f.S, f.ldx, ptr(f.pivot), ptr(f.dscale), ptr(Yt), int(Yt.stride(0)), l, ptr(Z),
                    int(Z.stride(0)), ptr(ws), ws.numel(), int(f.no_nan), ptr(e16), ptr(A16), int(A16.stride(0)),
                    self._stream())), "project_T_h16copy")
                f.h16 = (A16, ic16, None)
                self.launches += 3
                return Z
        ws = self.workspace(f.T, f.S, l, algo)
        flag = _lib.ALGO_FLAG_NO_NAN if (f.no_nan and f.row_valid is None) else 0
        check(self._timed("project_T" + self._algo_tag(algo), l, lambda: self.lib.xeofs_b200_project_T(
            ptr(f.X), f.T, f.S, f.ldx, ptr(f.pivot), ptr(f.dscale), ptr(f.ccorr), ptr(f.row_valid), ptr(Yt),
            int(Yt.stride(0)), l,
            ptr(Z), int(Z.stride(0)), ptr(ws), ws.numel(), algo | flag, self._stream())), "project_T")
        self.launches += 2 + (2 if f.ccorr is not None else 0)
        return Z

    _fast_algos = (_lib.ALGO_AUTO_FAST, _lib.ALGO_TF32X1)
    h16_min_bytes = 1 << 28  # fields below 256 MB are not worth a copy

    def _make_h16(self, f: Field):
        """Buffers for the fp16 copy of a field's preprocessed matrix (include/xeofs_b200.h, 'a half-precision copy'):
        (copy, e16, ic16), or None where it does not apply — un-centred fields (rank-1 term), missing samples, no
        tensor-core path, not enough free memory — in which case the field is not asked again."""
        f.want_h16 = False
        if (f.ccorr is not None or f.row_valid is not None or f.std is None
                or not bool(self.lib.xeofs_b200_has_tcgen05()) or f.ldx % 4 != 0 or f.X.data_ptr() % 16 != 0):
            return None
        pitch = (f.S + 127) // 128 * 128
        free, _ = torch.cuda.mem_get_info(self.device)
        cached = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        if f.T * pitch * 2 + (4 << 30) > free + cached:
            return None
        A16 = torch.empty((f.T, pitch), dtype=torch.int16, device=self.device)
        e16, ic16 = self.empty(f.S), self.empty(f.S)
        check(self.lib.xeofs_b200_h16_scales(ptr(f.dscale), ptr(f.std), f.S, ptr(e16), ptr(ic16), self._stream()),
              "h16_scales")
        return A16, e16, ic16

    def round_tf32_(self, M, rows, cols):
        """In place: keep the TF32 bits of every value (so that ALGO_TF32X2 products with M are exact)."""
        check(self.lib.xeofs_b200_round_tf32(ptr(M), int(rows), int(cols), int(M.stride(0)), self._stream()), "round_tf32")
        self.launches += 1
        return M

    # ------------------------------------------------------------------ k-column linear algebra
    def gram(self, M, n, l, side, out=None, accumulate=False):
        G = out if out is not None else self.empty((l, l), torch.float64)
        if l > BLOCK_L:  # beyond one block of the k-column kernels: one launch per pair of 128-column blocks
            Gn = G if not accumulate else self.empty((l, l), torch.float64)
            check(self.lib.xeofs_b200_gram_wide(ptr(M), n, l, int(M.stride(0)), side, ptr(Gn), self._stream()),
                  "gram_wide")
            if accumulate:
                G += Gn
            nb = (l + BLOCK_L - 1) // BLOCK_L
            self.launches += 1 + nb * (nb + 1) // 2
            return G
        check(self.lib.xeofs_b200_gram(ptr(M), n, l, int(M.stride(0)), side, ptr(G), int(accumulate), self._stream()),
              "gram")
        self.launches += 2
        return G

    def chol_inv(self, G, info=None):
        l = int(G.shape[0])
        Rinv = self.empty((l, l), torch.float64)
        info = info if info is not None else self.empty(2, torch.int32)
        if l > BLOCK_L:
            # wider than the one-block Cholesky kernel: the orthonormalising factor comes from the eigen-decomposition
            # G = V diag(ev) V^T instead, M V diag(ev^-1/2) — any orthonormal basis of the same span serves the range
            # finder (sklearn's own normalizers, LU and QR, produce different ones).  Directions below the fp32 noise
            # floor of G are dropped like the Cholesky kernel drops them.
            ev, V = self.sym_eig(G)
            ok = ev > 5.7e-14 * ev[0].clamp(min=0.0)
            scale = torch.where(ok, 1.0 / torch.sqrt(torch.where(ok, ev, torch.ones_like(ev))), torch.zeros_like(ev))
            Rinv.copy_(V * scale[None, :])
            info[0] = (~ok).sum().to(torch.int32)
            info[1] = (~torch.isfinite(ev).all()).to(torch.int32)
            return Rinv, info
        check(self.lib.xeofs_b200_chol_inv(ptr(G), l, ptr(Rinv), ptr(info), self._stream()), "chol_inv")
        self.launches += 1
        return Rinv, info

    def _unit_vectors(self, S):
        if self._unit is None or self._unit[0].numel() < S:
            self._unit = (self.zeros(S), torch.ones(S, dtype=torch.float32, device=self.device))
        return self._unit[0][:S], self._unit[1][:S]

    def apply(self, In, n, l, side, Mat, k, colscale=None, out=None):
        """Out(n, j') = sum_j In(n, j) Mat[j, j'] colscale[j'];  same side/layout as In, kp = lpad(k) columns."""
        kp = lpad(k)
        if (side == 1 and self.accurate_algo != _lib.ALGO_SIMT and In.stride(0) % 4 == 0 and In.data_ptr() % 16 == 0
                and bool(self.lib.xeofs_b200_has_tcgen05())):
            # space-side: (kp x n) = Mat^T (k x l) . In (l x n) is the streaming product A^T W of the (l x n) "field" In
            # with W = Mat: the tensor-core kernel in 3xTF32 (fp32-level accuracy), in place if out is In
            W = self.zeros((l, kp))
            M = Mat[:l, :k] if colscale is None else Mat[:l, :k] * colscale[None, :k]
            W[:, :k] = M.to(torch.float32)
            zero, one = self._unit_vectors(n)
            f = Field(In[:l, :n], zero, one, None, None)
            if out is None:
                out = self.space_side(kp, n)
            if k > BLOCK_L and out.data_ptr() == In.data_ptr():
                # in place over several column blocks: a later block would read rows an earlier one has overwritten
                tmp = self.project_S(f, W, k, algo=_lib.ALGO_TF32X3, out=self.space_side(kp, n), tag="apply_S")
                out[:kp].copy_(tmp)
                return out
            return self.project_S(f, W, k, algo=_lib.ALGO_TF32X3, out=out, tag="apply_S")
        if out is None:
            out = self.space_side(kp, n) if side == 1 else self.zeros((n, kp))
        if l > BLOCK_L or k > BLOCK_L:
            # (space-side blocks this wide only get here without the tensor-core path: unaligned views)
            A = (In[:n, :l] if side == 0 else In[:l, :n].t()).double().contiguous()
            Ms = Mat[:l, :k].double()
            if colscale is not None:
                Ms = Ms * colscale.double()[None, :k]
            B = self.dgemm(A, Ms.contiguous())
            if side == 0:
                out[:, :k] = B.float()
                out[:, k:] = 0
            else:
                out[:k] = B.t().float()
                out[k:] = 0
            return out
        check(self.lib.xeofs_b200_apply(ptr(In), n, l, int(In.stride(0)), side, ptr(Mat), int(Mat.stride(0)), k,
                                        ptr(colscale), ptr(out), int(out.stride(0)), self._stream()), "apply")
        self.launches += 1
        return out

    def sym_eig(self, G):
        l = int(G.shape[0])
        if l > BLOCK_L:
            return self.sym_eig_wide(G)
        evals = self.empty(l, torch.float64)
        evecs = self.empty((l, l), torch.float64)
        work = self.empty((l + 1) * (l + 1), torch.float64)
        info = self.empty(1, torch.int32)
        check(self.lib.xeofs_b200_sym_eig(ptr(G), l, ptr(evals), ptr(evecs), ptr(work), ptr(info), self._stream()),
              "sym_eig")
        self.launches += 1
        return evals, evecs

    def sym_eig_wide(self, G, max_sweeps=16):
        """Eigen-decomposition (descending) of a symmetric positive semi-definite matrix wider than one block:
        multi-CTA one-sided Jacobi (csrc/dense64.cu)."""
        n = int(G.shape[0])
        G = G.contiguous()
        evals = self.empty(n, torch.float64)
        evecs = self.empty((n, n), torch.float64)
        need = int(self.lib.xeofs_b200_sym_eig_wide_workspace_bytes(n))
        work = torch.empty(need, dtype=torch.uint8, device=self.device)
        info = self.empty(1, torch.int32)
        check(self.lib.xeofs_b200_sym_eig_wide(ptr(G), n, ptr(evals), ptr(evecs), ptr(work), need, ptr(info),
                                               int(max_sweeps), self._stream()), "sym_eig_wide")
        self.launches += 3 + max_sweeps * (n | 1)
        return evals, evecs

    def dgemm(self, A, B, trans_a=False, trans_b=False, alpha=1.0, out=None, beta=0.0):
        """op(A) op(B) in fp64 (row-major device matrices) on the library's own kernel: the small dense products of
        the k-column algebra (PCA scores, cross-covariance of score matrices, m x m rotation algebra)."""
        A = A if A.stride(-1) == 1 else A.contiguous()
        B = B if B.stride(-1) == 1 else B.contiguous()
        m, ka = (A.shape[1], A.shape[0]) if trans_a else (A.shape[0], A.shape[1])
        kb, n = (B.shape[1], B.shape[0]) if trans_b else (B.shape[0], B.shape[1])
        if ka != kb or A.dtype != torch.float64 or B.dtype != torch.float64:
            raise ValueError(f"dgemm: shapes {tuple(A.shape)} x {tuple(B.shape)} (trans {trans_a}, {trans_b}) / dtypes")
        C = out if out is not None else self.empty((int(m), int(n)), torch.float64)
        check(self.lib.xeofs_b200_dgemm(int(trans_a), int(trans_b), int(m), int(n), int(ka), float(alpha), ptr(A),
                                        int(A.stride(0)), ptr(B), int(B.stride(0)), float(beta), ptr(C),
                                        int(C.stride(0)), self._stream()), "dgemm")
        self.launches += 1
        return C

    def row_minmax(self, Vt, k, n):
        vmax, vmin = self.empty(k), self.empty(k)
        check(self.lib.xeofs_b200_row_minmax(ptr(Vt), k, n, int(Vt.stride(0)), ptr(vmax), ptr(vmin), self._stream()),
              "row_minmax")
        self.launches += 2
        return vmax, vmin

    def finish_components(self, Vt, k, n, sign, valid):
        check(self.lib.xeofs_b200_finish_components(ptr(Vt), k, n, int(Vt.stride(0)), ptr(sign), ptr(valid),
                                                    self._stream()), "finish_components")
        self.launches += 1

    def reconstruct(self, f: Field, scores, Vt, modes):
        """(T x S) = scores . V^T un-scaled (EOF.inverse_transform)."""
        T, m = int(scores.shape[0]), int(scores.shape[1])
        scores = scores.contiguous()
        idx = torch.as_tensor(modes, dtype=torch.int32).to(self.device)
        out = self.empty((T, f.S))
        check(self.lib.xeofs_b200_reconstruct(ptr(scores), T, int(scores.stride(0)), ptr(Vt), f.S, int(Vt.stride(0)),
                                              ptr(idx), m, ptr(f.pivot), ptr(f.dscale), ptr(f.ccorr), ptr(f.valid),
                                              ptr(out), int(out.stride(0)), self._stream()), "reconstruct")
        self.launches += 1
        return out

    def materialize(self, f: Field, round_tf32=True, row_multiple=128):
        """The preprocessed matrix of a field written out once (rows padded with zeros to a multiple of
        ``row_multiple``), optionally rounded to TF32: what the sample Gram matrices of MCA's total squared covariance
        stream ~T/256 times.  Returns a Field over the copy (pivot 0, dscale 1) or None when it does not apply
        (no tensor-core path, or not enough free memory for the copy)."""
        if not bool(self.lib.xeofs_b200_has_tcgen05()) or self.accurate_algo == _lib.ALGO_SIMT:
            return None
        rows = (f.T + row_multiple - 1) // row_multiple * row_multiple
        pitch = (f.S + 31) // 32 * 32
        free, _ = torch.cuda.mem_get_info(self.device)
        if rows * pitch * 4 + (2 << 30) > free + torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device):
            return None
        out = torch.empty((rows, pitch), dtype=torch.float32, device=self.device)
        check(self.lib.xeofs_b200_materialize(ptr(f.X), f.T, f.S, f.ldx, ptr(f.pivot), ptr(f.dscale), ptr(f.ccorr),
                                              ptr(f.row_valid), rows, int(round_tf32), ptr(out), pitch,
                                              self._stream()), "materialize")
        self.launches += 1
        zero, one = self._unit_vectors(f.S)
        return Field(out[:, :f.S], zero, one, None, None, no_nan=True)

    def sample_gram(self, f: Field):
        """Block-lower triangle of the sample Gram matrix A A^T (T x T fp32) of a field's preprocessed matrix: one bf16
        copy of the matrix, then a TMA-fed tcgen05 GEMM of its row tiles against themselves (csrc/gram_bf16.cu).
        Entries above the diagonal are valid only inside the 256 x 256 diagonal blocks: read it through torch.tril.
        None when it does not apply (no tensor-core path / not enough memory for the copy)."""
        if not bool(self.lib.xeofs_b200_has_tcgen05()) or self.accurate_algo == _lib.ALGO_SIMT:
            return None
        Tp, Sp = (f.T + 255) // 256 * 256, (f.S + 63) // 64 * 64
        need_ws = int(self.lib.xeofs_b200_gram_rows_bf16_workspace_bytes(Tp, Sp))
        free, _ = torch.cuda.mem_get_info(self.device)
        cached = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        if Tp * Sp * 2 + Tp * Tp * 4 + need_ws + (2 << 30) > free + cached:
            return None
        Ab = torch.empty((Tp, Sp), dtype=torch.bfloat16, device=self.device)
        check(self.lib.xeofs_b200_materialize_bf16(ptr(f.X), f.T, f.S, f.ldx, ptr(f.pivot), ptr(f.dscale), ptr(f.ccorr),
                                                   ptr(f.row_valid), Tp, Sp, ptr(Ab), Sp, self._stream()),
              "materialize_bf16")
        G = torch.empty((Tp, Tp), dtype=torch.float32, device=self.device)
        ws = torch.empty(need_ws, dtype=torch.uint8, device=self.device)
        check(self._timed("gram_rows_bf16", 256, lambda: self.lib.xeofs_b200_gram_rows_bf16(
            ptr(Ab), Tp, Sp, Sp, ptr(G), Tp, ptr(ws), need_ws, self._stream())), "gram_rows_bf16")
        self.launches += 1 + 2 * (Tp // 256)
        return G[:f.T, :f.T]

    def scaled_rows(self, f: Field, t0, t1):
        """Rows t0:t1 of the preprocessed matrix as a space-side block (pad rows zero)."""
        w = int(t1 - t0)
        out = self.space_side(lpad(w), f.S)
        check(self.lib.xeofs_b200_scaled_rows(ptr(f.X), f.T, f.S, f.ldx, ptr(f.pivot), ptr(f.dscale), ptr(f.ccorr),
                                              ptr(f.row_valid), int(t0), w, lpad(w), ptr(out), int(out.stride(0)),
                                              self._stream()), "scaled_rows")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ rotation
    def col_norms(self, L, S, m, normalized_out=False):
        h, rn = self.empty(S), self.empty(S)
        Ln = self.space_side(lpad(m), S, zero=True) if normalized_out else None
        check(self.lib.xeofs_b200_col_norms(ptr(L), S, m, int(L.stride(0)), ptr(h), ptr(rn), ptr(Ln),
                                            S if Ln is None else int(Ln.stride(0)), self._stream()), "col_norms")
        self.launches += 1
        return h, rn, Ln

    # the tensor-core sweep needs enough features for its fp32 rounding noise to average out below the reference's
    # stopping threshold (rtol 1e-8 on sum(svals)): relative noise of G ~ 1e-6 / sqrt(S)
    varimax_tc_min_S = 65536
    varimax_algo = "auto"  # "auto" | "simt" (fp64 CUDA cores) | "tc" (tcgen05 sweep wherever it applies)

    def _varimax_tc_applies(self, L, S, m, exact):
        if self.varimax_algo == "simt" or exact or m < 2 or m > 128:
            return False
        if not (bool(self.lib.xeofs_b200_has_tcgen05()) and L.stride(0) % 4 == 0 and L.data_ptr() % 16 == 0
                and int(L.shape[0]) >= lpad(m)):
            return False
        return self.varimax_algo == "tc" or (S >= self.varimax_tc_min_S and m >= 8)

    def varimax_update(self, G3, W, XtX, alpha, R, basis, dsum, eig_tol=0.0):
        """The m x m step of a varimax iteration on the device (xeofs_b200_varimax_update): R and basis are updated in
        place, the sum of the singular values lands in the one-element fp64 tensor ``dsum``.  No host round trip."""
        m = int(R.shape[0])
        need = int(self.lib.xeofs_b200_varimax_update_workspace_bytes(m))
        if getattr(self, "_uws", None) is None or self._uws.numel() < need:
            self._uws = torch.empty(need, dtype=torch.uint8, device=self.device)
        check(self.lib.xeofs_b200_varimax_update(ptr(G3), ptr(W), ptr(XtX), float(alpha), m, ptr(R), ptr(basis),
                                                 ptr(dsum), float(eig_tol), ptr(self._uws), need, self._stream()),
              "varimax_update")
        self.launches += 12
        return dsum

    def varimax_pack(self, L, S, m):
        """The tile-by-tile copy of the normalised loadings the tensor-core sweeps read (made once per rotation), or
        None where the packed kernel does not apply."""
        if not self._varimax_tc_applies(L, S, m, False):
            return None
        need = int(self.lib.xeofs_b200_varimax_pack_bytes(S, m))
        if need <= 0:
            return None
        packed = torch.empty(need // 4, dtype=torch.float32, device=self.device)
        check(self.lib.xeofs_b200_varimax_pack(ptr(L), S, m, int(L.stride(0)), ptr(packed), need, self._stream()),
              "varimax_pack")
        self.launches += 1
        return packed

    def varimax_accumulate(self, L, S, m, R, power=3.0, colscale=None, want_absmax=False, exact=False, products=3,
                           packed=None):
        """One sweep over the normalised loadings (linalg/_numpy/_rotation.py:166-170): Gout = Ln^T f(Ln R), W = colsum
        ((Ln R)^2).  exact=True keeps every product in fp64 (stopping thresholds below ~1e-9).  packed: what
        varimax_pack returned for the same L."""
        G = self.empty((m, m), torch.float64)
        Wv = self.empty(m, torch.float64)
        if power == 3.0 and colscale is None and not want_absmax and self._varimax_tc_applies(L, S, m, exact):
            need = int(self.lib.xeofs_b200_varimax_workspace_bytes(S, m))
            if self._vws is None or self._vws.numel() < need:
                self._vws = torch.empty(need, dtype=torch.uint8, device=self.device)
            check(self._timed("varimax_sweep" if products == 3 else "varimax_sweep_x1", m,
                              lambda: self.lib.xeofs_b200_varimax_sweep(
                ptr(L), ptr(packed), S, m, int(L.stride(0)), ptr(R), ptr(G), ptr(Wv), 0, int(products), ptr(self._vws),
                self._vws.numel(), self._stream())), "varimax_sweep")
            self.launches += 4
            return G, Wv, None
        amax = self.empty(m) if want_absmax else None
        check(self.lib.xeofs_b200_varimax_accumulate(ptr(L), S, m, int(L.stride(0)), None, ptr(R), float(power),
                                                     ptr(colscale), ptr(G), ptr(Wv), ptr(amax), 0, self._stream()),
              "varimax_accumulate")
        self.launches += 3
        return G, Wv, amax
