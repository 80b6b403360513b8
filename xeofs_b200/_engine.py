"""Host-side driver of the device kernels: preprocessing state, the randomized range finder and the small
SVD that follows it.  Nothing here touches array *values* on the host except k-length vectors.

What it stands in for (all paths under /root/reference/xeofs):
  preprocessing/scaler.py:69-154 + sanitizer.py:46-124      -> fit_field()
  linalg/decomposer.py:76-226 (policy, randomized_svd call, sign rule, U_/s_/V_)  -> decompose()
  sklearn.utils.extmath.randomized_svd (third party; call site decomposer.py:141-146) -> randomized_svd()

Layout vocabulary (see include/xeofs_b200.h): a *time-side* matrix is (n x lp) fp32 row-major, a
*space-side* matrix is (lp x n) fp32 (mode-major).  Side 0 = time, side 1 = space.  With several GPUs the
space (feature) axis is sharded across ranks, time-side matrices are replicated.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import lpad

MAX_L = 128         # widest k-column block the kernels take in one go (wider sketches run in blocks of it)
MAX_L_TOTAL = 1024  # widest sketch accepted at all
NORMALIZE_BOTH_HALF_STEPS = False  # True: sklearn's schedule (a normalisation after every half-step of a power iteration)


# ---------------------------------------------------------------------------------------------- comm
class Comm:
    """Feature-axis sharding over torch.distributed.  ``Comm(None)`` is the single-GPU no-op."""

    def __init__(self, group="auto"):
        import torch.distributed as dist

        self.dist = dist if (dist.is_available() and dist.is_initialized() and group is not None) else None
        self.group = None if group == "auto" else group
        self.world = self.dist.get_world_size(self.group) if self.dist else 1
        self.rank = self.dist.get_rank(self.group) if self.dist else 0
        self.collectives = 0

    @property
    def active(self):
        return self.world > 1

    def _reduce(self, t, op):
        if self.active:
            self.dist.all_reduce(t, op=op, group=self.group)
            self.collectives += 1
        return t

    def sum_(self, t):
        return self._reduce(t, self.dist.ReduceOp.SUM) if self.active else t

    def max_(self, t):
        return self._reduce(t, self.dist.ReduceOp.MAX) if self.active else t

    def min_(self, t):
        return self._reduce(t, self.dist.ReduceOp.MIN) if self.active else t


NO_COMM = Comm(None)


# ---------------------------------------------------------------------------------------------- preprocessing
class FittedField:
    """Device-resident field plus everything Scaler/Sanitizer would have fitted for it."""

    def __init__(self):
        self.field = None            # _cuda_ops.Field
        self.mean = self.std = None  # (S,) fp32 (NaN at invalid features)
        self.valid = None            # (S,) uint8
        self.featw = None            # (S,) fp64 coslat*weights or None
        self.valid_sample = None     # (T,) bool tensor
        self.n_samples = 0           # T' (samples kept)
        self.n_features = 0          # S' (features kept, global over ranks)
        self.total_variance = 0.0
        self.first_product = None    # A^T W of the fused statistics pass (space-side) or None
        self.first_l = None
        self.T = self.S = 0          # local shape
        self.S_global = 0
        self.center = self.standardize = False


def fit_field(ops, X, featw=None, center=True, standardize=False, check_nans=True, comm=NO_COMM, overlap=None,
              first=None):
    """One streaming pass of column statistics + the Scaler vectors (scaler.py:100-153), the Sanitizer
    masks and its isolated-NaN check (sanitizer.py:46-56, 108-122), total variance (utils/xarray_utils.py:236-253).

    X: (T, S_local) fp32 CUDA tensor (row stride arbitrary), featw: (S_local,) fp64 CUDA tensor or None.
    ``overlap``: host work to run while the statistics pass is in flight (before the one host sync).
    ``first``: (W, l) — the time-side sketch of the range finder: where the fused kernel applies, the statistics and
    the first product A^T W come from the same read of X (``ff.first_product``, valid if every sample is present).
    """
    from ._cuda_ops import Field

    T, S = int(X.shape[0]), int(X.shape[1])
    fused = ops.stats_project_S(X, featw, center, standardize, *first) if (
        first is not None and hasattr(ops, "stats_project_S")) else None
    if fused is not None:
        row_nan32, fin, first_product = fused
    else:
        st = ops.col_stats(X)
        fin = ops.scaling_finalize(st, featw, center, standardize)
        row_nan32, first_product = st["row_nan"], None
    # scalars: total variance, number of valid features, max / min non-NaN count over valid features
    sc = fin["scalars"].clone()
    row_nan = row_nan32.to(torch.int64)
    S_global = S
    if comm.active:
        head = torch.cat([sc[:2], torch.tensor([float(S)], dtype=torch.float64, device=sc.device)])
        comm.sum_(head)
        mx = sc[2:3].clone()
        comm.max_(mx)
        mn = sc[3:4].clone()
        comm.min_(mn)
        comm.sum_(row_nan)
        sc = torch.cat([head[:2], mx, mn, head[2:3]])
    else:
        sc = torch.cat([sc, torch.tensor([float(S)], dtype=torch.float64, device=sc.device)])
    # a sample is valid when it holds at least one non-NaN value (sanitizer.py:49-50); every sample must have either
    # none or all of the valid features (sanitizer.py:115-122).  Evaluated on the device, fetched with ONE copy.
    s_glob = sc[4]
    n_invalid = s_glob - sc[1]
    rn = row_nan.double()
    valid_sample = rn < s_glob
    pattern_ok = ((rn == n_invalid) | (rn == s_glob)).all()
    packed = torch.cat([sc, valid_sample.sum().double()[None], pattern_ok.double()[None]])
    if overlap is not None:
        overlap()
    h = packed.cpu().numpy()  # the one host sync of the preprocessing
    total_variance, n_valid, S_global = float(h[0]), int(round(h[1])), int(round(h[4]))
    n_samples = int(round(h[5]))
    if n_valid == 0:
        raise ValueError("Input data contains no valid (non-NaN) feature.")
    if check_nans and h[6] == 0.0:
        raise ValueError(
            "Input data contains partial NaN entries, which will cause the the SVD to fail."
        )
    ff = FittedField()
    # centred: pivot == mean, the rank-1 term vanishes;  all-NaN samples are named to the kernels only when present
    ccorr = None if center else fin["ccorr"]
    row_valid = valid_sample.to(torch.uint8) if n_samples < T else None
    ff.field = Field(X, fin["pivot"], fin["dscale"], ccorr, fin["valid"], fin["mean"], fin["std"], row_valid,
                     no_nan=(n_valid == S_global and n_samples == T))
    # the power iterations may stream an fp16 copy of the preprocessed matrix (written by their first time-side pass)
    if fin.get("h16") is not None and row_valid is None:
        ff.field.h16 = fin["h16"]  # the statistics pass wrote it (shifted by the first sample: with its rank-1 term)
    else:
        ff.field.want_h16 = bool(center and row_valid is None and T * S * 4 >= getattr(ops, "h16_min_bytes", 1 << 62)
                                 and getattr(ops, "use_h16", True))
    ff.mean, ff.std, ff.valid, ff.featw = fin["mean"], fin["std"], fin["valid"], featw
    ff.valid_sample, ff.n_samples, ff.n_features = valid_sample, n_samples, n_valid
    ff.total_variance = total_variance
    ff.T, ff.S, ff.S_global = T, S, S_global
    ff.center, ff.standardize = center, standardize
    # the fused first product took every sample as present and l as given
    ff.first_product = first_product if (first_product is not None and n_samples == T) else None
    ff.first_l = first[1] if first is not None else None
    return ff


# ---------------------------------------------------------------------------------------------- operators
class FieldOperator:
    """M = A^T when n_samples < n_features (what sklearn's ``transpose='auto'`` does), else M = A."""

    def __init__(self, ops, ff, comm=NO_COMM, algo=None):
        self.ops, self.ff, self.comm, self.algo = ops, ff, comm, algo
        self.n_rows_A, self.n_cols_A = ff.n_samples, ff.n_features
        self.transposed = ff.n_samples < ff.n_features
        # (global length, local length, side) of M's row and column spaces
        t_dim, s_dim = (ff.T, ff.T, 0), (ff.S_global, ff.S, 1)
        self.r, self.c = (s_dim, t_dim) if self.transposed else (t_dim, s_dim)
        self.shape = (ff.n_features, ff.n_samples) if self.transposed else (ff.n_samples, ff.n_features)
        # which entries of M's column space exist in the reference's (compacted) matrix
        t_mask = ff.valid_sample if ff.n_samples < ff.T else None
        s_mask = ff.valid.bool() if ff.n_features < ff.S_global else None
        self.c_mask = t_mask if self.transposed else s_mask

    def _proj_S(self, W, l, algo):
        return self.ops.project_S(self.ff.field, W, l, algo=algo)

    def _proj_T(self, Yt, l, algo):
        Z = self.ops.project_T(self.ff.field, Yt, l, algo=algo)
        self.comm.sum_(Z)
        return Z

    # products with a TF32-exact small operand need two tensor-core products instead of three for fp32 accuracy
    supports_exact = True

    def _algo(self, accurate, exact):
        return self.ops.exact_algo if exact else (self.ops.accurate_algo if accurate else self.algo)

    def mul(self, Q, l, accurate=False, exact=False):
        algo = self._algo(accurate, exact)
        return self._proj_S(Q, l, algo) if self.transposed else self._proj_T(Q, l, algo)

    def mul_t(self, Q, l, accurate=False, exact=False):
        algo = self._algo(accurate, exact)
        return self._proj_T(Q, l, algo) if self.transposed else self._proj_S(Q, l, algo)

    def sketch_rows(self):
        """Rows of the Gaussian test matrix: M.shape[1] (global)."""
        return self.c[0]


class CrossOperator:
    """Implicit cross-covariance C = X^T Y / (n - 1) (cross/cpcca.py:1008-1015) — never materialised:
    C Q = X^T (Y Q) / (n-1), C^T Q = Y^T (X Q) / (n-1).  M = C^T when S1 < S2 (sklearn's transpose rule)."""

    supports_exact = False  # the intermediate time-side block of C Q = X^T (Y Q) is not TF32-exact

    def __init__(self, ops, fx, fy, comm=NO_COMM, algo=None):
        self.ops, self.fx, self.fy, self.comm, self.algo = ops, fx, fy, comm, algo
        if fx.n_samples != fy.n_samples or fx.T != fy.T:
            raise ValueError(
                f"Both data matrices must have the same number of samples but found {fx.n_samples} in the "
                f"first and {fy.n_samples} in the second."
            )
        self.scale = 1.0 / (fx.n_samples - 1)
        self.transposed = fx.n_features < fy.n_features
        x_dim, y_dim = (fx.S_global, fx.S, 1), (fy.S_global, fy.S, 1)
        self.r, self.c = (y_dim, x_dim) if self.transposed else (x_dim, y_dim)
        self.shape = (fy.n_features, fx.n_features) if self.transposed else (fx.n_features, fy.n_features)
        cf = fx if self.transposed else fy
        self.c_mask = cf.valid.bool() if cf.n_features < cf.S_global else None

    def _through(self, f_in, f_out, Q, l, algo):
        Z = self.ops.project_T(f_in.field, Q, l, algo=algo)
        self.comm.sum_(Z)
        Z.mul_(self.scale)
        return self.ops.project_S(f_out.field, Z, l, algo=algo)

    def mul(self, Q, l, accurate=False):   # M @ Q, Q on M's column space
        algo = self.ops.accurate_algo if accurate else self.algo
        return self._through(self.fx, self.fy, Q, l, algo) if self.transposed else self._through(self.fy, self.fx, Q, l, algo)

    def mul_t(self, Q, l, accurate=False):
        algo = self.ops.accurate_algo if accurate else self.algo
        return self._through(self.fy, self.fx, Q, l, algo) if self.transposed else self._through(self.fx, self.fy, Q, l, algo)

    def sketch_rows(self):
        return self.c[0]


# ---------------------------------------------------------------------------------------------- k-column algebra
def _gram(ops, M, dim, l, comm):
    G = ops.gram(M, dim[1], l, dim[2])
    if dim[2] == 1:
        comm.sum_(G)  # space-side matrices are sharded over ranks
    return G


def orthonormalize(ops, M, dim, l, comm, passes, infos):
    """CholeskyQR (passes=1: the role of sklearn's LU normalizer; passes=2: its final QR), in place."""
    for _ in range(passes):
        G = _gram(ops, M, dim, l, comm)
        Rinv, info = ops.chol_inv(G)
        infos.append(info)
        ops.apply(M, dim[1], l, dim[2], Rinv, l, out=M)
    return M


_SKETCH_CACHE = {}


def draw_sketch(random_state, rows, l, f32=False):
    """The host draw of sklearn's range finder, rng.normal(size=(rows, l)) with numpy's legacy RandomState (the
    stream the reference uses, so both sides project on the same sketch).  The generator fills row-major, so the
    first n rows of a larger draw are the draw for n rows.  The draw is a pure function of an integer seed and the
    shape (about 30 ns per number, single-threaded): the last few are memoised, and so is their fp32 image (f32=True:
    what goes to the device; the conversion of a wide sketch costs milliseconds)."""
    if isinstance(random_state, np.random.RandomState) or random_state is None:
        rs = random_state if random_state is not None else np.random.RandomState(None)
        hit = rs.normal(size=(rows, l))
        return hit.astype(np.float32) if f32 else hit
    key = (int(random_state), int(rows), int(l))
    hit = _SKETCH_CACHE.get(key)
    if hit is None:
        hit = [np.random.RandomState(key[0]).normal(size=(rows, l)), None]
        if len(_SKETCH_CACHE) >= 4:
            _SKETCH_CACHE.pop(next(iter(_SKETCH_CACHE)))
        _SKETCH_CACHE[key] = hit
    if not f32:
        return hit[0]
    if hit[1] is None:
        hit[1] = hit[0].astype(np.float32)
    return hit[1]


def sketch_matrix(ops, op, l, random_state, comm=NO_COMM, predrawn=None):
    """sklearn.utils.extmath.randomized_range_finder: Q = rng.normal(size=(M.shape[1], l)) with M the reference's
    compacted matrix (all-NaN samples / features dropped), generated on the host with numpy's RandomState so that
    the oracle and the device path share the sketch; rows are scattered to the entries that survive the Sanitizer."""
    n_valid_global = op.shape[1]
    n_local, side, mask = op.c[1], op.c[2], op.c_mask
    if predrawn is not None and predrawn.shape[0] >= n_valid_global and predrawn.shape[1] == l:
        Om = predrawn[:n_valid_global]
    else:
        Om = draw_sketch(random_state, n_valid_global, l, f32=True)
    n_loc_valid = n_local if mask is None else int(mask.sum().item())
    offset = 0
    if side == 1 and comm.active:  # features are sharded: this rank's first valid feature in the global order
        counts = torch.zeros(comm.world, dtype=torch.int64, device=ops.device)
        counts[comm.rank] = n_loc_valid
        comm.sum_(counts)
        offset = int(counts[: comm.rank].sum().item())
    Om = ops.to_device(np.ascontiguousarray(Om[offset:offset + n_loc_valid], dtype=np.float32))
    lp = lpad(l)
    if side == 0:
        buf = ops.zeros((n_local, lp))
        if mask is None:
            buf[:, :l] = Om
        else:
            buf[mask, :l] = Om
        return buf
    buf = ops.space_side(lp, n_local, zero=True)
    if mask is None:
        buf[:l] = Om.t()
    else:
        buf[:l, mask] = Om.t()
    return buf


def randomized_svd(ops, op, k, n_oversamples=10, n_iter="auto", random_state=None, comm=NO_COMM, Omega=None,
                   predrawn=None, first_product=None):
    """Halko et al. range finder + small SVD, the arithmetic of sklearn.utils.extmath.randomized_svd
    (power_iteration_normalizer='auto', transpose='auto') with CholeskyQR as the normalizer.

    Returns (Ur, s, Vc, infos): Ur r-side (lp x n or n x lp) left singular vectors of M, s (k,) fp64 device,
    Vc c-side right singular vectors of M, all un-flipped.
    """
    n_r, n_c = op.shape
    l = min(k + n_oversamples, n_r, n_c)
    if l > MAX_L_TOTAL:
        raise NotImplementedError(
            f"n_modes + n_oversamples = {k + n_oversamples} exceeds the widest sketch of this build ({MAX_L_TOTAL})"
        )
    if n_iter == "auto":
        n_iter = 7 if k < 0.1 * min(n_r, n_c) else 4
    infos = []
    if Omega is not None:
        Q = Omega
    elif first_product is not None and int(n_iter) >= 1:
        Q = None  # M @ Omega is in hand already (fused statistics pass): the sketch itself is not needed again
    else:
        Q = sketch_matrix(ops, op, l, random_state, comm, predrawn)
    lp = lpad(l)

    def _shape2d(dim):  # (rows, cols) of a k-column block on this side
        return (dim[1], lp) if dim[2] == 0 else (lp, dim[1])

    for it in range(int(n_iter)):
        # M @ Omega may already exist: the statistics pass computed it from the same read of the field
        Y = first_product if (it == 0 and first_product is not None) else op.mul(Q, l)
        # sklearn normalises after both half-steps.  From the second iteration on the columns of Q are graded (column j
        # is dominated by the j-th singular direction), M^T (M Q) keeps them graded, and one normalisation per full
        # iteration (on M's column side) holds the same span to fp32 accuracy: the row-side one is left out.
        if it == 0 or NORMALIZE_BOTH_HALF_STEPS:
            Y = orthonormalize(ops, Y, op.r, l, comm, 1, infos)
        Q = orthonormalize(ops, op.mul_t(Y, l), op.c, l, comm, 1, infos)
    # The last two passes decide the singular values and run at fp32 accuracy.  For the first of them the small
    # operand can be made TF32-exact beforehand — rounding the current iterate is harmless, any nearby iterate serves
    # the range finder — and then two tensor-core products (field hi/lo x operand) do instead of three.  The range
    # basis itself must not be rounded: a perturbation delta of its span costs (sigma_1 delta)^2 / sigma_j in the small
    # singular values, so B = Q^T M takes the orthonormal Q as it is, in 3xTF32.
    if getattr(op, "supports_exact", False):
        ops.round_tf32_(Q, *_shape2d(op.c))
        Y = op.mul(Q, l, exact=True)
    else:
        Y = op.mul(Q, l, accurate=True)
    Q = orthonormalize(ops, Y, op.r, l, comm, 2, infos)                         # r-side, orthonormal
    Bt = op.mul_t(Q, l, accurate=True)                                          # c-side: B^T = M^T Q
    # svd(B) through the l x l Gram of B^T: B B^T = Uh diag(s^2) Uh^T; V = B^T Uh / s; U = Q Uh
    G = _gram(ops, Bt, op.c, l, comm)
    evals, Uh = ops.sym_eig(G)
    s = torch.sqrt(torch.clamp(evals[:k], min=0.0))
    inv_s = torch.where(s > 0, 1.0 / s, torch.zeros_like(s))
    Uh_k = Uh[:, :k].contiguous()
    Vc = ops.apply(Bt, op.c[1], l, op.c[2], Uh_k, k, colscale=inv_s)
    Ur = ops.apply(Q, op.r[1], l, op.r[2], Uh_k, k)
    return Ur, s, Vc, infos


def check_infos(infos):
    """Raise the reference's LinAlgError (decomposer.py:265-270) when a non-finite pivot was met."""
    if not infos:
        return
    flags = torch.stack(infos)[:, 1]
    if bool(flags.any().item()):
        raise np.linalg.LinAlgError(
            "SVD failed: non-finite values met during the decomposition. Check the input for NaN / Inf "
            "entries that the NaN policy did not remove."
        )


def sign_flip(ops, Vt, k, n, comm):
    """utils/xarray_utils.py:273-301: +1 where |max| >= |min| over the feature axis, else -1 (fp32 device)."""
    vmax, vmin = ops.row_minmax(Vt, k, n)
    comm.max_(vmax)
    comm.min_(vmin)
    return torch.where(vmax.abs() >= vmin.abs(), 1.0, -1.0).to(torch.float32)
