"""TEST DOUBLE (not product code): the CudaOps interface of xeofs_b200/_cuda_ops.py implemented with torch on the
CPU, so that the host logic (preprocessor, range finder driver, EOF / MCA / EOFRotator classes, feature sharding
over torch.distributed) can be exercised on a box without a GPU.  Each method states the same operation as the
C-ABI kernel of the same name (include/xeofs_b200.h) in float64."""
import numpy as np
import torch

from xeofs_b200._cuda_ops import Field  # noqa: F401  (plain container, no CUDA)
from xeofs_b200._lib import lpad

EPS32 = float(np.finfo(np.float32).eps)


class TorchCpuOps:
    name = "cpu-test-double"

    def __init__(self):
        self.device = torch.device("cpu")
        self.algo = self.accurate_algo = self.exact_algo = 1
        self.launches = 0
        self.time_products = False

    # ------------------------------------------------------------------ helpers
    def empty(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype)

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype)

    def space_side(self, rows, n, zero=False):
        ld = (int(n) + 31) // 32 * 32
        return torch.zeros((rows, ld), dtype=torch.float32)[:, :n]

    def to_device(self, a, dtype=None):
        t = torch.as_tensor(a)
        return t.to(dtype) if dtype is not None else t

    # ------------------------------------------------------------------ preprocessing
    def col_stats(self, X):
        Xd = X.double()
        nan = torch.isnan(Xd)
        shift = torch.where(nan[0], torch.zeros_like(Xd[0]), Xd[0]).float()
        d = torch.where(nan, torch.zeros_like(Xd), Xd - shift.double()[None, :])
        return {"shift": shift, "sum": d.sum(0), "sumsq": (d * d).sum(0), "cnt": (~nan).sum(0).to(torch.int32),
                "row_nan": nan.sum(1).to(torch.int32)}

    def scaling_finalize(self, st, featw, center, standardize):
        n = st["cnt"].double()
        ok = st["cnt"] > 0
        a = torch.where(ok, st["sum"] / n.clamp(min=1), torch.zeros_like(n))
        mu = st["shift"].double() + a
        m2 = (st["sumsq"] - st["sum"] * a).clamp(min=0)
        mu32 = mu.float()
        sd = torch.sqrt(m2 / n.clamp(min=1)).float().clamp(min=EPS32)
        d = (featw.double() if featw is not None else torch.ones_like(n))
        d = torch.where(ok, d, torch.zeros_like(d))
        if standardize:
            d = torch.where(ok, d / sd.double(), d)
        d32 = d.float()
        piv = torch.where(ok, mu32 if center else st["shift"], torch.zeros_like(mu32))
        mu_eff = mu32 if center else torch.zeros_like(mu32)
        nanv = torch.full_like(mu32, float("nan"))
        tv = torch.where(ok & (st["cnt"] > 1), d32.double() ** 2 * m2 / (n - 1).clamp(min=1), torch.zeros_like(n)).sum()
        cnt_ok = st["cnt"][ok]
        scalars = torch.tensor([float(tv), float(ok.sum()), float(cnt_ok.max()) if cnt_ok.numel() else 0.0,
                                float(cnt_ok.min()) if cnt_ok.numel() else 2147483647.0], dtype=torch.float64)
        return {"mean": torch.where(ok, mu32, nanv), "std": torch.where(ok, sd, nanv), "valid": ok.to(torch.uint8),
                "pivot": piv, "dscale": d32, "ccorr": torch.where(ok, (piv - mu_eff) * d32, torch.zeros_like(d32)),
                "scalars": scalars}

    # ------------------------------------------------------------------ streaming products
    @staticmethod
    def _A(f):
        A = (f.X.double() - f.pivot.double()[None, :]) * f.dscale.double()[None, :]
        A = torch.nan_to_num(A, nan=0.0)
        if f.ccorr is not None:
            A = A + f.ccorr.double()[None, :]
        if f.row_valid is not None:
            A = A * f.row_valid.double()[:, None]
        return A

    def project_S(self, f, W, l, algo=None, out=None):
        lp = lpad(l)
        Yt = out if out is not None else self.space_side(lp, f.S)
        Yt[:] = (self._A(f).t() @ W[:, :lp].double()).t().float()
        return Yt

    def project_T(self, f, Yt, l, algo=None, out=None):
        lp = lpad(l)
        Z = out if out is not None else self.zeros((f.T, lp))
        Z[:, :lp] = (self._A(f) @ Yt[:lp].double().t()).float()
        return Z

    def round_tf32_(self, M, rows, cols):
        v = M[:rows, :cols].contiguous().view(torch.int32) & -8192  # 0xffffe000
        M[:rows, :cols] = v.view(torch.float32)
        return M

    # ------------------------------------------------------------------ k-column linear algebra
    @staticmethod
    def _cols(M, n, l, side):
        return (M[:n, :l] if side == 0 else M[:l, :n].t()).double()

    def gram(self, M, n, l, side, out=None, accumulate=False):
        A = self._cols(M, n, l, side)
        G = A.t() @ A
        if out is not None:
            out[:] = out + G if accumulate else G
            return out
        return G

    def chol_inv(self, G, info=None):
        l = G.shape[0]
        info = torch.zeros(2, dtype=torch.int32)
        if not torch.isfinite(G).all():
            info[1] = 1
            return torch.zeros_like(G), info
        # right-looking Cholesky that drops columns whose pivot is below the fp32 noise floor (smallmat.cu)
        A = G.clone()
        R = torch.zeros_like(G)
        dead = []
        for k in range(l):
            piv = A[k, k]
            if not (piv > 4 * EPS32 * EPS32 * G[k, k]) or G[k, k] <= 0:
                dead.append(k)
                continue
            R[k, k:] = A[k, k:] / torch.sqrt(piv)
            A[k + 1:, k + 1:] -= torch.outer(R[k, k + 1:], R[k, k + 1:])
        info[0] = len(dead)
        Rinv = torch.zeros_like(G)
        live = [k for k in range(l) if k not in dead]
        if live:
            Rl = R[live][:, live]
            Rinv[np.ix_(live, live)] = torch.linalg.inv(Rl)
        return Rinv, info

    def apply(self, In, n, l, side, Mat, k, colscale=None, out=None):
        kp = lpad(k)
        A = self._cols(In, n, l, side)
        B = A @ Mat[:l, :k].double()
        if colscale is not None:
            B = B * colscale.double()[None, :k]
        if out is None:
            out = self.space_side(kp, n) if side == 1 else self.zeros((n, kp))
        if side == 1:
            out[:kp] = 0
            out[:k] = B.t().float()
        else:
            out[:, :k] = B.float()
        return out

    def sym_eig(self, G):
        ev, V = torch.linalg.eigh(0.5 * (G + G.t()))
        return ev.flip(0).contiguous(), V.flip(1).contiguous()

    def dgemm(self, A, B, trans_a=False, trans_b=False, alpha=1.0, out=None, beta=0.0):
        C = alpha * ((A.t() if trans_a else A).double() @ (B.t() if trans_b else B).double())
        if out is not None:
            out.copy_(C + beta * out if beta != 0.0 else C)
            return out
        return C

    def row_minmax(self, Vt, k, n):
        return Vt[:k, :n].max(dim=1).values.clone(), Vt[:k, :n].min(dim=1).values.clone()

    def finish_components(self, Vt, k, n, sign, valid):
        Vt[:k, :n] *= sign[:k, None]
        if valid is not None:
            Vt[:k, :n][:, ~valid.bool()] = float("nan")

    def reconstruct(self, f, scores, Vt, modes):
        idx = torch.as_tensor(np.asarray(modes), dtype=torch.long)
        rec = scores.double() @ Vt[idx].double()
        cc = f.ccorr.double() if f.ccorr is not None else 0.0
        out = (rec - cc) / f.dscale.double()[None, :] + f.pivot.double()[None, :]
        out[:, ~f.valid.bool()] = float("nan")
        return out.float()

    def scaled_rows(self, f, t0, t1):
        w = int(t1 - t0)
        out = self.space_side(lpad(w), f.S, zero=True)
        out[:w] = self._A(f)[t0:t1].float()
        return out

    # ------------------------------------------------------------------ rotation
    def col_norms(self, L, S, m, normalized_out=False):
        h = torch.sqrt((L[:m, :S].double() ** 2).sum(0)).float()
        rn = 1.0 / (h + 2.220446e-16)
        Ln = None
        if normalized_out:
            Ln = self.space_side(lpad(m), S, zero=True)
            Ln[:m] = L[:m, :S] * rn[None, :]
        return h, rn, Ln

    def varimax_update(self, G3, W, XtX, alpha, R, basis, dsum, eig_tol=0.0):
        G = G3 - alpha * (XtX @ R) * W[None, :]
        U, sv, Vh = torch.linalg.svd(G)
        R.copy_(U @ Vh)
        basis.copy_(Vh.t())
        dsum.fill_(float(sv.sum()))
        return dsum

    def varimax_pack(self, L, S, m):
        return None

    def varimax_accumulate(self, L, S, m, R, power=3.0, colscale=None, want_absmax=False, exact=False, products=3,
                           packed=None):
        X = L[:m, :S].double().t()
        B = X @ R
        Bc = B * colscale.double()[None, :] if colscale is not None else B
        F = Bc * Bc.abs() ** (power - 1)
        return X.t() @ F, (B * B).sum(0), (B.abs().max(0).values.float() if want_absmax else None)
