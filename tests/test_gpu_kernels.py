"""GPU: every C-ABI kernel against a float64 numpy/torch statement of the same operation."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from xeofs_b200._cuda_ops import CudaOps
    return CudaOps(algo="simt")


def _field(ops, X, featw=None, center=True, standardize=False):
    from xeofs_b200._cuda_ops import Field
    Xd = torch.as_tensor(X, dtype=torch.float32).cuda()
    st = ops.col_stats(Xd)
    fw = None if featw is None else torch.as_tensor(featw, dtype=torch.float64).cuda()
    fin = ops.scaling_finalize(st, fw, center, standardize)
    return Field(Xd, fin["pivot"], fin["dscale"], fin["ccorr"], fin["valid"], fin["mean"], fin["std"]), st, fin


def _A_ref(X, featw=None, center=True, standardize=False):
    X = X.astype(np.float64)
    A = X.copy()
    if center:
        A = A - np.nanmean(X, axis=0)
    if standardize:
        A = A / np.clip(np.nanstd(X, axis=0).astype(np.float32), np.finfo(np.float32).eps, None)
    if featw is not None:
        A = A * featw
    return A


@pytest.mark.parametrize("T,S", [(25, 20), (300, 1001), (1000, 4096), (77, 130)])
@pytest.mark.parametrize("standardize", [False, True])
def test_col_stats_and_scaling(ops, T, S, standardize):
    rng = np.random.default_rng(T * S)
    X = (280 + 3 * rng.standard_normal((T, S))).astype(np.float32)
    X[:, 3] = np.nan
    if S > 100:
        X[:, 97] = np.nan
    featw = rng.uniform(0.5, 1.5, S)
    f, st, fin = _field(ops, X, featw, True, standardize)
    mean_ref = np.nanmean(X.astype(np.float64), axis=0)
    valid = ~np.isnan(X).all(axis=0)
    np.testing.assert_array_equal(fin["valid"].cpu().numpy().astype(bool), valid)
    np.testing.assert_allclose(fin["mean"].cpu().numpy()[valid], mean_ref[valid], rtol=2e-7)
    std_ref = np.nanstd(X.astype(np.float64), axis=0)
    np.testing.assert_allclose(fin["std"].cpu().numpy()[valid], std_ref[valid], rtol=1e-5)
    A = _A_ref(X, featw, True, standardize)[:, valid]
    tv = np.var(A, axis=0, ddof=1).sum()
    sc = fin["scalars"].cpu().numpy()
    np.testing.assert_allclose(sc[0], tv, rtol=1e-5)
    assert int(sc[1]) == valid.sum() and int(sc[2]) == T and int(sc[3]) == T
    rn = st["row_nan"].cpu().numpy()
    assert (rn == (~valid).sum()).all()


@pytest.mark.parametrize("T,S,l", [(25, 20, 12), (300, 1001, 20), (1000, 4096, 60), (520, 777, 110), (64, 130, 5)])
@pytest.mark.parametrize("center", [True, False])
def test_project_simt(ops, T, S, l, center):
    from xeofs_b200._lib import lpad
    rng = np.random.default_rng(l)
    X = (280 + 3 * rng.standard_normal((T, S))).astype(np.float32)
    X[:, 1] = np.nan
    featw = rng.uniform(0.5, 1.5, S)
    f, st, fin = _field(ops, X, featw, center, True)
    A = np.nan_to_num(_A_ref(X, featw, center, True), nan=0.0)
    lp = lpad(l)
    W = np.zeros((T, lp), np.float32)
    W[:, :l] = rng.standard_normal((T, l))
    Yt = ops.project_S(f, torch.from_numpy(W).cuda(), l).cpu().numpy()
    ref = (A.T @ W[:, :l].astype(np.float64)).T
    scale = np.abs(ref).max()
    np.testing.assert_allclose(Yt[:l], ref, atol=2e-5 * scale)
    assert (Yt[l:] == 0).all()
    Y = np.zeros((lp, S), np.float32)
    Y[:l] = rng.standard_normal((l, S))
    Yd = ops.space_side(lp, S, zero=True)
    Yd.copy_(torch.from_numpy(Y))
    Z = ops.project_T(f, Yd, l).cpu().numpy()
    ref = A @ Y[:l].astype(np.float64).T
    np.testing.assert_allclose(Z[:, :l], ref, atol=2e-5 * np.abs(ref).max())


@pytest.mark.parametrize("n,l,side", [(1000, 20, 0), (5000, 60, 1), (333, 110, 1), (40, 7, 0), (100000, 33, 1),
                                      (20000, 128, 1), (3000, 128, 0)])
@pytest.mark.parametrize("kind", ["simt", "auto"])
def test_gram_chol_apply(kind, n, l, side):
    from xeofs_b200._cuda_ops import CudaOps
    ops = CudaOps(algo=kind)  # "auto": the space-side apply runs on the tensor cores (3xTF32)
    from xeofs_b200._lib import lpad
    rng = np.random.default_rng(n + l)
    M = rng.standard_normal((n, l)) @ np.diag(np.logspace(0, -3, l)) @ rng.standard_normal((l, l))
    lp = lpad(l)
    buf = np.zeros((n, lp) if side == 0 else (lp, n), np.float32)
    if side == 0:
        buf[:, :l] = M
        Md = torch.from_numpy(buf).cuda()
    else:
        buf[:l] = M.T
        Md = ops.space_side(lp, n)
        Md.copy_(torch.from_numpy(buf))
    M32 = (buf[:, :l] if side == 0 else buf[:l].T).astype(np.float64)
    G = ops.gram(Md, n, l, side)
    np.testing.assert_allclose(G.cpu().numpy(), M32.T @ M32, rtol=1e-10, atol=1e-10 * np.abs(M32.T @ M32).max())
    Rinv, info = ops.chol_inv(G)
    assert info.cpu().numpy().tolist() == [0, 0]
    R = np.linalg.cholesky(M32.T @ M32).T
    np.testing.assert_allclose(Rinv.cpu().numpy(), np.linalg.inv(R), rtol=1e-5, atol=1e-7 * np.abs(np.linalg.inv(R)).max())
    Q = ops.apply(Md, n, l, side, Rinv, l)
    Q = ops.apply(Q, n, l, side, *ops.chol_inv(ops.gram(Q, n, l, side))[:1], l)  # second pass
    Qh = Q.cpu().numpy()
    Qh = Qh[:, :l] if side == 0 else Qh[:l].T
    np.testing.assert_allclose(Qh.T @ Qh, np.eye(l), atol=5e-6)


def test_chol_drops_dependent_columns(ops):
    rng = np.random.default_rng(0)
    M = rng.standard_normal((500, 8)).astype(np.float32)
    M[:, 5] = M[:, 1] + M[:, 2]      # exactly dependent in fp32 up to rounding
    M[:, 7] = 0
    Md = torch.zeros((500, 16), device="cuda")
    Md[:, :8] = torch.from_numpy(M).cuda()
    G = ops.gram(Md, 500, 8, 0)
    Rinv, info = ops.chol_inv(G)
    assert info.cpu().numpy().tolist() == [2, 0]
    Q = ops.apply(Md, 500, 8, 0, Rinv, 8).cpu().numpy()[:, :8]
    assert np.abs(Q[:, 5]).max() == 0 and np.abs(Q[:, 7]).max() == 0
    keep = [0, 1, 2, 3, 4, 6]
    np.testing.assert_allclose(Q[:, keep].T @ Q[:, keep], np.eye(6), atol=1e-5)


@pytest.mark.parametrize("l", [2, 5, 30, 60, 110, 128])
def test_sym_eig(ops, l):
    rng = np.random.default_rng(l)
    B = rng.standard_normal((l, 3 * l)) * np.logspace(0, -4, l)[:, None]
    G = B @ B.T
    ev, V = ops.sym_eig(torch.from_numpy(G).cuda())
    ev, V = ev.cpu().numpy(), V.cpu().numpy()
    ref = np.linalg.eigvalsh(G)[::-1]
    np.testing.assert_allclose(ev, ref, rtol=1e-9, atol=1e-14 * ref[0])
    np.testing.assert_allclose(V.T @ V, np.eye(l), atol=1e-12)
    np.testing.assert_allclose(V @ np.diag(ev) @ V.T, G, atol=1e-12 * ref[0])


def test_minmax_finish_reconstruct(ops):
    rng = np.random.default_rng(3)
    k, S, T = 7, 1000, 90
    Vt = rng.standard_normal((16, S)).astype(np.float32)
    Vd = torch.from_numpy(Vt).cuda()
    mx, mn = ops.row_minmax(Vd, k, S)
    np.testing.assert_array_equal(mx.cpu().numpy(), Vt[:k].max(axis=1))
    np.testing.assert_array_equal(mn.cpu().numpy(), Vt[:k].min(axis=1))
    sign = torch.tensor([1, -1, 1, -1, -1, 1, 1], dtype=torch.float32).cuda()
    ops.finish_components(Vd, k, S, sign, None)
    np.testing.assert_array_equal(Vd.cpu().numpy()[:k], Vt[:k] * sign.cpu().numpy()[:, None])
    X = (5 + rng.standard_normal((T, S))).astype(np.float32)
    X[:, 10] = np.nan
    f, st, fin = _field(ops, X, None, True, True)
    sc = rng.standard_normal((T, 3)).astype(np.float32)
    rec = ops.reconstruct(f, torch.from_numpy(sc).cuda(), Vd, [0, 2, 5]).cpu().numpy()
    V = Vd.cpu().numpy()[[0, 2, 5]].astype(np.float64)
    ref = (sc.astype(np.float64) @ V) * np.nanstd(X.astype(np.float64), axis=0) + np.nanmean(X.astype(np.float64), axis=0)
    assert np.isnan(rec[:, 10]).all()
    ok = np.arange(S) != 10
    np.testing.assert_allclose(rec[:, ok], ref[:, ok], rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------- tcgen05 products
TC_SHAPES = [(25, 20, 12), (300, 1000, 20), (1000, 4096, 60), (520, 776, 110), (64, 132, 5), (8760, 2500, 60),
             (200, 70000, 40)]


def _tc_case(T, S, l, center, seed):
    from xeofs_b200._lib import lpad
    rng = np.random.default_rng(seed)
    X = (280 + 3 * rng.standard_normal((T, S))).astype(np.float32)
    X[:, 1] = np.nan
    if T > 30:
        X[7] = np.nan       # an all-NaN sample contributes zeros
    featw = rng.uniform(0.5, 1.5, S)
    lp = lpad(l)
    W = np.zeros((T, lp), np.float32)
    W[:, :l] = rng.standard_normal((T, l))
    Y = np.zeros((lp, S), np.float32)
    Y[:l] = rng.standard_normal((l, S))
    return X, featw, W, Y, lp


@pytest.mark.parametrize("T,S,l", TC_SHAPES)
@pytest.mark.parametrize("center", [True, False])
@pytest.mark.parametrize("algo,tol", [("tf32x3", 2e-5), ("tf32x1", 3e-3), ("tf32x1r", 2e-3), ("simt", 2e-5)])
def test_project_tcgen05(T, S, l, center, algo, tol):
    """tcgen05 kind::tf32 kernels (A operand in TMEM, B by TMA) against the fp64 statement of A^T W and A Y.
    3xTF32 must be as accurate as the fp32 SIMT kernel; single TF32 carries 2^-11 operand rounding."""
    from xeofs_b200 import _lib
    from xeofs_b200._cuda_ops import CudaOps
    ops = CudaOps(algo="simt")
    a = _lib.ALGO_NAMES[algo]
    X, featw, W, Y, lp = _tc_case(T, S, l, center, seed=T + S + l)
    Xd = ops.space_side(T, S)
    Xd.copy_(torch.from_numpy(X))
    from xeofs_b200._cuda_ops import Field
    st = ops.col_stats(Xd)
    fin = ops.scaling_finalize(st, torch.as_tensor(featw, dtype=torch.float64).cuda(), center, True)
    row_valid = (st["row_nan"] < S).to(torch.uint8)
    f = Field(Xd, fin["pivot"], fin["dscale"], fin["ccorr"], fin["valid"], fin["mean"], fin["std"], row_valid)
    A = np.nan_to_num(_A_ref(X, featw, center, True), nan=0.0)
    Yt = ops.project_S(f, torch.from_numpy(W).cuda(), l, algo=a).cpu().numpy()
    ref = (A.T @ W[:, :l].astype(np.float64)).T
    np.testing.assert_allclose(Yt[:l], ref, atol=tol * np.abs(ref).max())
    assert (Yt[l:] == 0).all()
    Yd = ops.space_side(lp, S, zero=True)
    Yd.copy_(torch.from_numpy(Y))
    Z = ops.project_T(f, Yd, l, algo=a).cpu().numpy()
    ref = A @ Y[:l].astype(np.float64).T
    np.testing.assert_allclose(Z[:, :l], ref, atol=tol * np.abs(ref).max())
    # the deterministic split reduction: a second call is bit-identical
    if algo != "simt":  # the SIMT validation kernels combine their split sums with atomics
        Z2 = ops.project_T(f, Yd, l, algo=a).cpu().numpy()
        np.testing.assert_array_equal(Z, Z2)


@pytest.mark.parametrize("T,S,l", [(300, 1000, 20), (1000, 4096, 60), (8760, 2500, 60), (200, 70000, 40), (90, 132, 5)])
@pytest.mark.parametrize("center,standardize", [(True, False), (True, True), (False, True)])
def test_fused_stats_and_first_product(T, S, l, center, standardize):
    """xeofs_b200_project_S_stats (one read of X) against col_stats + scaling_finalize + project_S: the Scaler vectors,
    the masks, total variance, per-sample NaN counts, and A^T W to single-TF32 accuracy.  NaN features (a land mask)
    but every sample present — the case the fused pass is for."""
    from xeofs_b200 import _lib
    from xeofs_b200._cuda_ops import CudaOps, Field
    from xeofs_b200._lib import lpad
    ops = CudaOps(algo="auto")
    rng = np.random.default_rng(T + S + l)
    X = (280 + 3 * rng.standard_normal((T, S))).astype(np.float32)
    X[:, rng.random(S) < 0.1] = np.nan
    X[:, 0] = np.nan
    featw = rng.uniform(0.5, 1.5, S)
    lp = lpad(l)
    W = np.zeros((T, lp), np.float32)
    W[:, :l] = rng.standard_normal((T, l))
    Xd = ops.space_side(T, S)
    Xd.copy_(torch.from_numpy(X))
    fw = torch.as_tensor(featw, dtype=torch.float64).cuda()
    Wd = torch.from_numpy(W).cuda()
    fused = ops.stats_project_S(Xd, fw, center, standardize, Wd, l)
    assert fused is not None
    row_nan, fin, Yt = fused
    st = ops.col_stats(Xd)
    ref = ops.scaling_finalize(st, fw, center, standardize)
    np.testing.assert_array_equal(row_nan.cpu().numpy(), st["row_nan"].cpu().numpy())
    np.testing.assert_array_equal(fin["valid"].cpu().numpy(), ref["valid"].cpu().numpy())
    v = ref["valid"].cpu().numpy().astype(bool)
    for key, tol in (("mean", 3e-7), ("std", 2e-5), ("pivot", 3e-7), ("dscale", 2e-5)):
        np.testing.assert_allclose(fin[key].cpu().numpy()[v], ref[key].cpu().numpy()[v], rtol=tol, err_msg=key)
    assert np.isnan(fin["mean"].cpu().numpy()[~v]).all() and (fin["dscale"].cpu().numpy()[~v] == 0).all()
    sc, sr = fin["scalars"].cpu().numpy(), ref["scalars"].cpu().numpy()
    np.testing.assert_allclose(sc[0], sr[0], rtol=1e-5)
    np.testing.assert_array_equal(sc[1:], sr[1:])
    A = np.nan_to_num(_A_ref(X, featw, center, standardize), nan=0.0)
    want = (A.T @ W[:, :l].astype(np.float64)).T
    got = Yt.cpu().numpy()
    np.testing.assert_allclose(got[:l], want, atol=3e-3 * np.abs(want).max())
    assert (got[l:] == 0).all() and (got[:, ~v] == 0).all()
    # and the statistics against an fp64 numpy statement of Scaler.fit / total_variance (not only against the library's
    # own stand-alone kernels): scaler.py:100-108 (mean, std ddof = 0), utils/xarray_utils.py:236-253 (ddof = 1)
    X64 = X.astype(np.float64)
    np.testing.assert_array_equal(v, ~np.isnan(X64).all(axis=0))
    np.testing.assert_allclose(fin["mean"].cpu().numpy()[v], X64[:, v].mean(axis=0), rtol=3e-7)
    np.testing.assert_allclose(fin["std"].cpu().numpy()[v], X64[:, v].std(axis=0), rtol=3e-5)
    np.testing.assert_allclose(sc[0], A[:, v].var(axis=0, ddof=1).sum(), rtol=2e-5)
    assert sc[1] == v.sum()


def _varimax_case(ops, S, m, seed):
    from xeofs_b200._lib import lpad
    g = torch.Generator(device="cuda").manual_seed(seed)
    Ln = ops.space_side(lpad(m), S, zero=True)
    Ln[:m] = torch.randn((m, S), generator=g, device="cuda") * (0.2 + torch.rand((m, 1), generator=g, device="cuda"))
    Ln[:m] /= Ln[:m].double().norm(dim=0).float()[None, :]  # Kaiser-normalised rows of the S x m loadings
    R = torch.linalg.qr(torch.randn((m, m), generator=g, device="cuda", dtype=torch.float64))[0].contiguous()
    X = Ln[:m].double().t()
    B = X @ R
    return Ln, R, X.t() @ B**3, (B * B).sum(0)


@pytest.mark.parametrize("S,m", [(40, 3), (1001, 10), (5000, 33), (777, 64), (20011, 100), (130, 104), (3000, 110)])
def test_varimax_accumulate_fp64(ops, S, m):
    """linalg/_numpy/_rotation.py:166-170, the fp64 sweep (CUDA cores; 32 < m <= 104: the fp64 mma.sync kernel)."""
    ops.varimax_algo = "simt"
    Ln, R, Gref, Wref = _varimax_case(ops, S, m, seed=S)
    G, W, _ = ops.varimax_accumulate(Ln, S, m, R)
    np.testing.assert_allclose(G.cpu().numpy(), Gref.cpu().numpy(), atol=1e-12 * float(Gref.abs().max()))
    np.testing.assert_allclose(W.cpu().numpy(), Wref.cpu().numpy(), rtol=1e-12)
    ops.varimax_algo = "auto"


@pytest.mark.parametrize("m", [2, 7, 50, 100, 101, 118, 128])
@pytest.mark.parametrize("kind", ["random", "clustered", "warm"])
def test_varimax_update_polar_factor(ops, m, kind):
    """The m x m step of a varimax iteration (_rotation.py:170-175): G = G3 - alpha (XtX R) diag(W), R <- U V^T of
    svd(G), delta = sum(svals) — through the eigen-decomposition of G^T G in the previous iteration's basis — against
    LAPACK."""
    g = torch.Generator(device="cuda").manual_seed(m)
    rnd = lambda *sh: torch.randn(sh, generator=g, device="cuda", dtype=torch.float64)  # noqa: E731
    Q0 = torch.linalg.qr(rnd(m, m))[0]
    Q1 = torch.linalg.qr(rnd(m, m))[0]
    if kind == "clustered":  # nearly equal singular values (the planted patterns of config 5) and a far one
        sv = 1.0 + 1e-9 * torch.arange(m, device="cuda", dtype=torch.float64)
        sv[0] = 50.0
    else:
        sv = torch.logspace(0, -3, m, device="cuda", dtype=torch.float64)
    G3 = (Q0 * sv[None, :]) @ Q1.t()
    W = torch.rand(m, generator=g, device="cuda", dtype=torch.float64)
    XtX = rnd(m, m)
    XtX = XtX @ XtX.t() / m
    R = torch.linalg.qr(rnd(m, m))[0].contiguous()
    alpha = 1e-3
    G = G3 - alpha * (XtX @ R) * W[None, :]
    U, s, Vh = torch.linalg.svd(G)
    basis = torch.eye(m, dtype=torch.float64, device="cuda")
    if kind == "warm":  # the right singular vectors, slightly rotated: what the previous iteration leaves behind
        K = 1e-3 * rnd(m, m)
        basis = (Vh.t() @ torch.linalg.matrix_exp(K - K.t())).contiguous()
    dsum = torch.zeros(1, dtype=torch.float64, device="cuda")
    Rn = R.clone()
    ops.varimax_update(G3.contiguous(), W, XtX.contiguous(), alpha, Rn, basis, dsum)
    ref = U @ Vh
    tol = 1e-6 if kind == "clustered" else 1e-8  # (the route through G^T G squares the condition number: 1e6 here)
    assert float((Rn - ref).abs().max()) < tol
    assert abs(float(dsum.item()) / float(s.sum()) - 1.0) < 1e-12
    assert float((Rn.t() @ Rn - torch.eye(m, dtype=torch.float64, device="cuda")).abs().max()) < 1e-11
    # the basis handed to the next iteration diagonalises G^T G
    D = basis.t() @ (G.t() @ G) @ basis
    off = D - torch.diag(torch.diagonal(D))
    assert float(off.norm() / torch.diagonal(D).norm()) < 1e-9


@pytest.mark.parametrize("S,m", [(64, 8), (33, 5), (4096, 20), (20000, 50), (100037, 100), (30011, 128), (65536 + 17, 97),
                                 (70001, 104), (9999, 112)])
def test_varimax_sweep_tcgen05(S, m):
    """The same sweep on the tensor cores (both products kind::tf32 with hi/lo split operands, fp64 accumulation
    across tiles): fp32-level agreement with the fp64 statement."""
    from xeofs_b200._cuda_ops import CudaOps
    ops = CudaOps()
    ops.varimax_algo = "tc"
    Ln, R, Gref, Wref = _varimax_case(ops, S, m, seed=m)
    packed = ops.varimax_pack(Ln, S, m)  # tiles of 64 features from the packed copy (three-product mode: m <= 104)
    assert packed is not None
    G, W, _ = ops.varimax_accumulate(Ln, S, m, R, packed=packed)
    scale = float(Gref.abs().max())
    np.testing.assert_allclose(G.cpu().numpy(), Gref.cpu().numpy(), atol=2e-6 * scale)
    np.testing.assert_allclose(W.cpu().numpy(), Wref.cpu().numpy(), rtol=2e-6)
    # the sweep is deterministic (fixed tile order per CTA, partial sums added in a fixed order)
    G2, W2, _ = ops.varimax_accumulate(Ln, S, m, R, packed=packed)
    assert torch.equal(G, G2) and torch.equal(W, W2)
    # single-TF32 mode (the first phase of the iteration): operands rounded to 11 bits, errors of ~1e-3 per term that
    # average out over the features
    for pk in (packed, None):
        G1, W1, _ = ops.varimax_accumulate(Ln, S, m, R, products=1, packed=pk)
        np.testing.assert_allclose(G1.cpu().numpy(), Gref.cpu().numpy(), atol=3e-3 * scale)
        np.testing.assert_allclose(W1.cpu().numpy(), Wref.cpu().numpy(), rtol=3e-3)
    # without the packed copy: tiles of 32 features through tensor maps (also what m > 104 falls back to in the
    # three-product mode)
    G3, W3, _ = ops.varimax_accumulate(Ln, S, m, R)
    np.testing.assert_allclose(G3.cpu().numpy(), Gref.cpu().numpy(), atol=2e-6 * scale)
    np.testing.assert_allclose(W3.cpu().numpy(), Wref.cpu().numpy(), rtol=2e-6)


@pytest.mark.parametrize("T,S,nan_cols", [(700, 70000, 0), (300, 66001, 37)])
def test_sample_gram_bf16_gemm(T, S, nan_cols):
    """xeofs_b200_materialize_bf16 + xeofs_b200_gram_rows_bf16 (TMA-fed kind::f16 GEMM of the row tiles of a bf16 copy of
    the preprocessed matrix against themselves) vs the fp64 Gram matrix of the same preprocessed field.  Every entry
    sums S >= 65 536 products with independent bf16 rounding errors (unit roundoff 2^-8): the deviation, relative to
    |a_t| |a_t'|, is ~1e-5 typically and stays below 1.5e-4 over the 245 000 entries (measured maximum 7.4e-5); what is
    read from these matrices, the total squared covariance, is held to 2e-5 by
    test_mca_total_squared_covariance_wide_fields.  Only the lower triangle is specified."""
    from xeofs_b200._cuda_ops import CudaOps
    tc_ops = CudaOps()
    rng = np.random.default_rng(T)
    U = np.linalg.qr(rng.standard_normal((T, 6)))[0]
    X = (280 + (U * (300 * 0.7 ** np.arange(6))) @ rng.standard_normal((6, S)) / np.sqrt(S) * 30 + 0.3 * rng.standard_normal((T, S))).astype(np.float32)
    if nan_cols:
        X[:, rng.choice(S, nan_cols, replace=False)] = np.nan
    f, _, _ = _field(tc_ops, X, center=True, standardize=True)
    G = tc_ops.sample_gram(f)
    assert G is not None and tuple(G.shape) == (T, T)
    A = np.nan_to_num(_A_ref(X, center=True, standardize=True), nan=0.0)
    ref = A @ A.T
    d = np.sqrt(np.diag(ref))
    err = np.tril(np.abs(G.cpu().numpy().astype(np.float64) - ref) / np.outer(d, d))
    assert err.max() < 1.5e-4, err.max()
    assert np.median(err[np.tril_indices(T)]) < 1e-5


@pytest.mark.parametrize("T,S,l,standardize", [(300, 1000, 10, False), (1000, 4100, 60, True), (777, 2048 + 64, 110, True)])
def test_half_precision_copy_passes(T, S, l, standardize):
    """The fp16 copy of the preprocessed matrix (include/xeofs_b200.h): the single-TF32 project_T pass that writes it
    (xeofs_b200_project_T_h16copy), then project_S16 / project_T16 on the copy (kind::f16), all against the fp64
    statement at the accuracy of a single TF32 product; a land mask (NaN features) and a feature 1e4 times larger than
    the rest exercise the per-feature power-of-two scale."""
    from xeofs_b200 import _lib
    from xeofs_b200._cuda_ops import CudaOps
    from xeofs_b200._lib import lpad
    tc_ops = CudaOps()
    rng = np.random.default_rng(T + l)
    X = (280 + 3 * rng.standard_normal((T, S))).astype(np.float32)
    X[:, 5] = 280 + 3e4 * rng.standard_normal(T)
    X[:, rng.random(S) < 0.05] = np.nan
    f, _, _ = _field(tc_ops, X, center=True, standardize=standardize)
    assert f.ccorr is None or float(f.ccorr.abs().max()) == 0.0
    f.ccorr = None
    f.want_h16 = True
    A = np.nan_to_num(_A_ref(X, center=True, standardize=standardize), nan=0.0)
    lp = lpad(l)
    Y = tc_ops.space_side(lp, S, zero=True)
    Y[:l] = torch.randn((l, S), device="cuda") * torch.logspace(0, -4, l, device="cuda")[:, None]  # graded columns
    W = torch.zeros((T, lp), device="cuda")
    W[:, :l] = torch.linalg.qr(torch.randn((T, l), device="cuda", dtype=torch.float64))[0].float()
    wantT = A @ Y[:l].double().cpu().numpy().T
    wantS = (A.T @ W[:, :l].double().cpu().numpy()).T
    colT = np.abs(wantT).max(axis=0)
    # 1. the pass that writes the copy is an ordinary single-TF32 product
    Z = tc_ops.project_T(f, Y, l, algo=_lib.ALGO_TF32X1)
    assert f.h16 is not None and not f.want_h16
    np.testing.assert_allclose(Z.cpu().numpy()[:, :l] / colT, wantT / colT, atol=3e-3)
    # the copy itself: A16 ic16 == A to fp16 accuracy (11 bits), zero at the NaN features
    A16, ic16, cc16 = f.h16
    assert cc16 is None
    got = (A16.view(torch.float16)[:, :S].double() * ic16.double()[None, :]).cpu().numpy()
    scale = np.abs(A).max(axis=0) + 1e-30
    assert np.max(np.abs(got - A) / scale) < 1e-3
    # 2. products on the copy
    Z2 = tc_ops.project_T(f, Y, l, algo=_lib.ALGO_TF32X1)
    np.testing.assert_allclose(Z2.cpu().numpy()[:, :l] / colT, wantT / colT, atol=3e-3)
    assert (Z2.cpu().numpy()[:, l:] == 0).all()
    Yt = tc_ops.project_S(f, W, l, algo=_lib.ALGO_TF32X1)
    np.testing.assert_allclose(Yt.cpu().numpy()[:l], wantS, atol=3e-3 * np.abs(wantS).max())
    assert (Yt.cpu().numpy()[l:] == 0).all()
    # the accurate products never take the copy
    Z3 = tc_ops.project_T(f, Y, l, algo=_lib.ALGO_TF32X3)
    np.testing.assert_allclose(Z3.cpu().numpy()[:, :l] / colT, wantT / colT, atol=2e-5)


@pytest.mark.parametrize("T,S,l,center,standardize", [(300, 1000, 10, True, False), (1000, 4100, 60, True, True)])
def test_statistics_pass_writes_the_half_precision_copy(T, S, l, center, standardize, monkeypatch):
    """xeofs_b200_project_S_stats_h16copy: the fused statistics + first product pass also writes the fp16 copy of the
    field shifted by its first sample, with the factor ic16 and the rank-1 term cc16 that turn it into the fitted matrix:
    A = A16 ic16 + cc16.  Checked: the statistics and the product as without the copy, the copy against the fp64
    preprocessed matrix, and project_S16 / project_T16 with the rank-1 term against fp64 (single-TF32 accuracy).  The
    record has a trend, so that samples far from the first one are much larger than the pre-sample suggests locally."""
    from xeofs_b200 import _lib
    from xeofs_b200._cuda_ops import CudaOps, Field
    from xeofs_b200._lib import lpad
    monkeypatch.setattr(CudaOps, "h16_min_bytes", 0)
    ops = CudaOps(algo="auto")
    rng = np.random.default_rng(T + S)
    X = (280 + 3 * rng.standard_normal((T, S)) + np.linspace(0, 40, T)[:, None] * rng.random(S)[None, :]).astype(np.float32)
    X[:, rng.random(S) < 0.1] = np.nan
    lp = lpad(l)
    W = np.zeros((T, lp), np.float32)
    W[:, :l] = np.linalg.qr(rng.standard_normal((T, l)))[0]
    Xd = ops.space_side(T, S)
    Xd.copy_(torch.from_numpy(X))
    Wd = torch.from_numpy(W).cuda()
    row_nan, fin, Yt = ops.stats_project_S(Xd, None, center, standardize, Wd, l)
    assert fin.get("h16") is not None
    A16, ic16, cc16 = fin["h16"]
    A = np.nan_to_num(_A_ref(X, None, center, standardize), nan=0.0)
    v = fin["valid"].cpu().numpy().astype(bool)
    np.testing.assert_array_equal(v, ~np.isnan(X).all(axis=0))
    want = (A.T @ W[:, :l].astype(np.float64)).T
    np.testing.assert_allclose(Yt.cpu().numpy()[:l], want, atol=3e-3 * np.abs(want).max())
    rec = (A16.view(torch.float16)[:, :S].double() * ic16.double()[None, :] + cc16.double()[None, :]).cpu().numpy()
    d = fin["dscale"].cpu().numpy().astype(np.float64)
    span = np.nanmax(np.abs(X - X[0]), axis=0, initial=0.0, where=~np.isnan(X)) * np.abs(d) + 1e-30
    assert np.max(np.abs(rec[:, v] - A[:, v]) / span[v]) < 1e-3   # 11 bits of the largest deviation from the first sample
    assert (rec[:, ~v] == 0).all()
    f = Field(Xd, fin["pivot"], fin["dscale"], None, fin["valid"], fin["mean"], fin["std"], None, no_nan=False)
    f.h16 = fin["h16"]
    Y = ops.space_side(lp, S, zero=True)
    Y[:l] = torch.randn((l, S), device="cuda") * torch.logspace(0, -3, l, device="cuda")[:, None]
    wantT = A @ Y[:l].double().cpu().numpy().T
    colT = np.abs(wantT).max(axis=0)
    Z = ops.project_T(f, Y, l, algo=_lib.ALGO_TF32X1)
    np.testing.assert_allclose(Z.cpu().numpy()[:, :l] / colT, wantT / colT, atol=4e-3)
    Yt2 = ops.project_S(f, Wd, l, algo=_lib.ALGO_TF32X1)
    np.testing.assert_allclose(Yt2.cpu().numpy()[:l], want, atol=4e-3 * np.abs(want).max())
