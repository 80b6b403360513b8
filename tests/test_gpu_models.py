"""GPU parity through the C-ABI for the other two classes on the path: xeofs_b200.cross.MCA (implicit cross-covariance
operator) and xeofs_b200.single.EOFRotator (varimax / promax, rows R1-R2), against the oracle on
identical seeded inputs; plus size-independent a-posteriori properties of EOF.fit at sizes the oracle cannot reach."""
import numpy as np
import pytest
import torch

from _inputs import planted
from oracle import eof as oeof
from oracle import mca as omca
from oracle import rotation as orot

pytestmark = pytest.mark.gpu
DIMS = ("time", "lat", "lon")


def _coupled_fields(T, S1, S2, r, seed):
    rng = np.random.default_rng(seed)
    U = np.linalg.qr(rng.standard_normal((T, r)))[0]
    sig = 500 * 0.75 ** np.arange(r)
    X = 280 + (U * sig) @ np.linalg.qr(rng.standard_normal((S1, r)))[0].T + 0.05 * rng.standard_normal((T, S1))
    Y = 1000 + (U * sig) @ np.linalg.qr(rng.standard_normal((S2, r)))[0].T + 0.05 * rng.standard_normal((T, S2))
    return X.astype(np.float32), Y.astype(np.float32)


@pytest.mark.parametrize("shape", [(300, 40 * 30, 20 * 36, 6), (400, 10 * 12, 60 * 50, 8)])
@pytest.mark.parametrize("kw", [dict(), dict(standardize=True, use_coslat=True)])
def test_mca_matches_explicit_cross_covariance_oracle(shape, kw):
    """C = X^T Y/(n-1) is formed explicitly by the oracle (cross/cpcca.py:1008-1015) and applied implicitly on the
    device; singular values rtol 1e-4, components up to the reference's sign rule, total squared covariance."""
    import xeofs_b200 as xb
    T, S1, S2, k = shape
    X, Y = _coupled_fields(T, S1, S2, 2 * k, seed=S1)
    n1 = 40 if S1 == 1200 else 10
    n2 = 20 if S2 == 720 else 60
    X = X.reshape(T, n1, S1 // n1)
    Y = Y.reshape(T, n2, S2 // n2)
    X[:, 3, 5] = np.nan
    Y[:, 7, 1] = np.nan
    cx = {"lat": np.linspace(80, -80, n1), "lon": np.arange(S1 // n1) * 1.0}
    cy = {"lat": np.linspace(60, -60, n2), "lon": np.arange(S2 // n2) * 1.0}
    o = omca.mca_fit(X, Y, DIMS, DIMS, "time", coords_x=cx, coords_y=cy, n_modes=k, random_state=3, **kw)
    m = xb.cross.MCA(n_modes=k, random_state=3, use_pca=False, **kw)
    m.fit(xb.DataArray(X, DIMS, cx), xb.DataArray(Y, DIMS, cy), dim="time")
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    np.testing.assert_allclose(m.total_squared_covariance(), o["total_squared_covariance"], rtol=1e-4)
    c1, c2 = m.components()
    for c, oc, f in ((c1, o["components1_2d"], o["fitted1"]), (c2, o["components2_2d"], o["fitted2"])):
        V = c.values.reshape(-1, k)
        np.testing.assert_array_equal(np.isnan(V).any(axis=1), ~f["is_valid_feature"])
        dots = (V[f["is_valid_feature"]] * oc).sum(axis=0)
        assert (dots >= 1 - 1e-4).all(), dots
    s1, s2 = m.scores()
    for sc, osc in ((s1, o["scores1"]), (s2, o["scores2"])):
        scale = np.abs(osc).max(axis=0)
        np.testing.assert_allclose(sc.values / scale, osc / scale, atol=2e-3)


@pytest.mark.parametrize("power", [1, 2])
def test_mca_rotator_matches_oracle(power):
    """MCARotator (cross/cpcca_rotator.py:122-305) on the device against the numpy restatement: squared covariance
    rtol 1e-4, both sets of rotated singular vectors up to the reference's sign rule, rotated scores."""
    import xeofs_b200 as xb
    T, S1, S2, k, mr = 300, 40 * 30, 20 * 37, 8, 5
    X, Y = _coupled_fields(T, S1, S2, 2 * k, seed=5)
    X = X.reshape(T, 40, 30)
    Y = Y.reshape(T, 20, 37)
    X[:, 3, 5] = np.nan
    o = omca.mca_fit(X, Y, DIMS, DIMS, "time", n_modes=k, random_state=3)
    m = xb.cross.MCA(n_modes=k, random_state=3, use_pca=False, total_squared_covariance=False)
    m.fit(xb.DataArray(X, DIMS), xb.DataArray(Y, DIMS), dim="time")
    r = xb.cross.MCARotator(n_modes=mr, power=power).fit(m)
    ro = orot.mca_rotator_fit(o["components1_2d"], o["components2_2d"], o["singular_values"], o["scores1"],
                              o["scores2"], n_modes=mr, power=power)
    np.testing.assert_allclose(r.squared_covariance().values, ro["squared_covariance"], rtol=1e-4)
    c1, c2 = r.components()
    for c, oc, f in ((c1, ro["components1_2d"], o["fitted1"]), (c2, ro["components2_2d"], o["fitted2"])):
        V = c.values.reshape(-1, mr)
        np.testing.assert_array_equal(np.isnan(V).any(axis=1), ~f["is_valid_feature"])
        dots = (V[f["is_valid_feature"]] * oc).sum(axis=0)
        assert (dots >= 1 - 1e-4).all(), dots
    s1, s2 = r.scores()
    for sc, osc in ((s1, ro["scores1"]), (s2, ro["scores2"])):
        scale = np.abs(osc).max(axis=0)
        np.testing.assert_allclose(sc.values / scale, osc / scale, atol=2e-3)


@pytest.mark.parametrize("shapes", [((40, 50), (30, 20)), ((8, 9), (5, 7))])
def test_eof_list_input_matches_oracle(shapes):
    """EOF.fit on a list of arrays (two variables, different grids and units): concatenated along the feature axis
    (preprocessing/preprocessor.py:208-228); tensor-core path for the aligned case, fp32 SIMT for the ragged one."""
    import xeofs_b200 as xb
    T, k = 200, 5
    (a1, b1), (a2, b2) = shapes
    full = planted(T, a1 * b1 + a2 * b2, 2 * k, seed=13)
    X1 = full[:, : a1 * b1].reshape(T, a1, b1).copy()
    X2 = (3.0 * full[:, a1 * b1:] + 1000.0).reshape(T, a2, b2).astype(np.float32)
    X1[:, 1, 2] = np.nan
    c1 = {"lat": np.linspace(40, -40, a1), "lon": np.arange(b1) * 1.0}
    c2 = {"lat": np.linspace(30, -30, a2), "lon": np.arange(b2) * 2.0}
    kw = dict(n_modes=k, standardize=True, use_coslat=True, random_state=4)
    o = oeof.eof_fit_list([X1, X2], [DIMS, DIMS], "time", coords_list=[c1, c2], **kw)
    m = xb.single.EOF(**kw).fit([xb.DataArray(X1, DIMS, c1), xb.DataArray(X2, DIMS, c2)], dim="time")
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    dots = 0.0
    for c, oc, f in zip(m.components(), o["components_2d"], o["fitted"]):
        V = c.values.reshape(-1, k)
        np.testing.assert_array_equal(np.isnan(V).any(axis=1), ~f["is_valid_feature"])
        dots = dots + (V[f["is_valid_feature"]] * oc).sum(axis=0)
    assert (np.abs(dots) > 1 - 1e-4).all(), dots


def test_eof_api_corners_match_oracle():
    """float n_modes, user weights + coslat, two sample dimensions, T-mode — on the device (see the host-logic twin)."""
    import xeofs_b200 as xb
    T, nlat, nlon = 240, 12, 20
    X = planted(T, nlat * nlon, 10, seed=21).reshape(T, nlat, nlon)
    coords = {"lat": np.linspace(50, -50, nlat), "lon": np.arange(nlon) * 10.0}
    da = xb.DataArray(X, DIMS, coords)
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, n_modes=0.9, random_state=3)
    m = xb.single.EOF(n_modes=0.9, random_state=3).fit(da, dim="time")
    assert m.singular_values().values.shape == o["singular_values"].shape
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    w = np.linspace(0.5, 2.0, nlon)
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, n_modes=4, use_coslat=True, random_state=3,
                     weights=np.broadcast_to(w, (nlat, nlon)))
    m = xb.single.EOF(n_modes=4, use_coslat=True, random_state=3).fit(da, dim="time", weights=xb.DataArray(w, ("lon",)))
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    assert ((m.components().values.reshape(-1, 4) * o["components_2d"]).sum(axis=0) > 1 - 1e-4).all()
    X4 = X.reshape(20, 12, nlat, nlon)
    dims4 = ("year", "month", "lat", "lon")
    o = oeof.eof_fit(X4, dims4, ("year", "month"), coords=coords, n_modes=4, random_state=3)
    m = xb.single.EOF(n_modes=4, random_state=3).fit(xb.DataArray(X4, dims4, coords), dim=("year", "month"))
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    assert m.scores().values.shape == (20, 12, 4)
    o = oeof.eof_fit(X, DIMS, ("lat", "lon"), coords=coords, n_modes=4, random_state=3)
    m = xb.single.EOF(n_modes=4, random_state=3).fit(da, dim=("lat", "lon"))
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    np.testing.assert_allclose(np.abs((m.components().values * o["components_2d"]).sum(axis=0)), 1.0, atol=1e-4)


def test_eof_more_modes_than_one_kernel_block():
    """n_modes + oversamples = 160 > 128: products in two column blocks, k-column algebra through the library
    fallbacks; leading singular values against the oracle, orthonormal components, scores = A V."""
    import xeofs_b200 as xb
    T, S, k = 600, 64 * 64, 150
    X = planted(T, S, 40, seed=33).reshape(T, 64, 64)
    o = oeof.eof_fit(X, DIMS, "time", n_modes=k, random_state=3)
    m = xb.single.EOF(n_modes=k, random_state=3).fit(xb.DataArray(X, DIMS), dim="time")
    np.testing.assert_allclose(m.singular_values().values[:30], o["singular_values"][:30], rtol=1e-4)
    V = m.components().values.reshape(-1, k)
    np.testing.assert_allclose(V.T @ V, np.eye(k), atol=2e-4)
    dots = (V[:, :20] * o["components_2d"][:, :20]).sum(axis=0)
    assert (dots > 1 - 1e-4).all(), dots
    A = o["A"]
    sc = m.scores().values
    np.testing.assert_allclose(sc[:, :20], A @ V[:, :20], atol=2e-3 * np.abs(sc[:, :20]).max())


def test_bootstrapper_matches_oracle():
    """EOFBootstrapper (validation/bootstrapper.py:56-135) on the device: resampled fits + projection of the original
    samples against the numpy restatement, members seeded on both sides."""
    import xeofs_b200 as xb
    from oracle import bootstrap as oboot
    T, nlat, nlon, k, nb = 300, 40, 50, 6, 3
    X = planted(T, nlat * nlon, 2 * k, seed=31).reshape(T, nlat, nlon)
    X[:, 5, 7] = np.nan
    coords = {"lat": np.linspace(70, -70, nlat), "lon": np.arange(nlon) * 5.0}
    kw = dict(n_modes=k, use_coslat=True, random_state=1)
    m = xb.single.EOF(**kw).fit(xb.DataArray(X, DIMS, coords), dim="time")
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, **kw)
    b = xb.validation.EOFBootstrapper(n_bootstraps=nb, seed=5, random_state=2).fit(m)
    ob = oboot.eof_bootstrap(o["A"], o["scores"], k, n_bootstraps=nb, seed=5, random_state=2)
    np.testing.assert_allclose(b.explained_variance().values, ob["explained_variance"], rtol=1e-4)
    np.testing.assert_allclose(b.total_variance().values, ob["total_variance"], rtol=1e-5)
    valid = ~np.isnan(X[0]).reshape(-1)
    for i, c in enumerate(b.components()):
        V = c.values.reshape(-1, k)
        np.testing.assert_array_equal(np.isnan(V).any(axis=1), ~valid)
        dots = (V[valid] * ob["components"][i]).sum(axis=0)
        assert (dots > 1 - 1e-4).all(), dots
    for i, sc in enumerate(b.scores()):
        scale = np.abs(ob["scores"][i]).max(axis=0)
        np.testing.assert_allclose(sc.values / scale, ob["scores"][i] / scale, atol=2e-3)


@pytest.mark.parametrize("npm", [0.999, 10])
def test_mca_default_pca_stage_matches_oracle(npm):
    """MCA with the reference's default use_pca=True: PCA of both fields from the sample Gram matrices (device) vs the
    reference's randomized-SVD PCA (oracle, seeded); singular values rtol 1e-4, patterns up to the sign rule."""
    import xeofs_b200 as xb
    T, S1, S2, k = 300, 40 * 30, 20 * 36, 5
    X, Y = _coupled_fields(T, S1, S2, 2 * k, seed=7)
    X = X.reshape(T, 40, 30)
    Y = Y.reshape(T, 20, 36)
    X[:, 3, 5] = np.nan
    o = omca.mca_fit(X, Y, DIMS, DIMS, "time", n_modes=k, random_state=3, use_pca=True, n_pca_modes=npm,
                     pca_random_state=1)
    m = xb.cross.MCA(n_modes=k, random_state=3, n_pca_modes=npm)
    m.fit(xb.DataArray(X, DIMS), xb.DataArray(Y, DIMS), dim="time")
    # the 99.9 % cut falls among noise-level modes, where the reference's randomized PCA and the exact one may differ
    # by a mode or two; an integer n_pca_modes is kept exactly
    assert all(abs(a - b) <= (2 if isinstance(npm, float) else 0) for a, b in zip(m.n_pca_modes_, o["n_pca_modes"]))
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    np.testing.assert_allclose(m.total_squared_covariance(), o["total_squared_covariance"], rtol=1e-4)
    c1, c2 = m.components()
    for c, oc, f in ((c1, o["components1_2d"], o["fitted1"]), (c2, o["components2_2d"], o["fitted2"])):
        V = c.values.reshape(-1, k)
        np.testing.assert_array_equal(np.isnan(V).any(axis=1), ~f["is_valid_feature"])
        dots = (V[f["is_valid_feature"]] * oc).sum(axis=0)
        assert (dots >= 1 - 1e-4).all(), dots
    s1, s2 = m.scores()
    for sc, osc in ((s1, o["scores1"]), (s2, o["scores2"])):
        scale = np.abs(osc).max(axis=0)
        np.testing.assert_allclose(sc.values / scale, osc / scale, atol=2e-3)


def test_mca_pca_stage_wide_sketch_runs_on_own_kernels():
    """2000 samples, noise-dominated tail: the 99.9 % cut is not reached inside one 128-column sketch, so the PCA stage
    widens it up to int(0.3 n) = 600 modes (the reference's own width) — the k-column algebra beyond one kernel block
    (cross-block Gram, multi-CTA Jacobi, eigen-orthonormalisation).  No torch.linalg routine (cuSOLVER / cuBLAS) may run
    during the fit; same oracle (the reference's randomized-SVD PCA), same bars."""
    import xeofs_b200 as xb
    T, S1, S2, k = 2000, 64 * 64, 48 * 64, 5
    X, Y = _coupled_fields(T, S1, S2, 2 * k, seed=19)
    X = X.reshape(T, 64, 64)
    Y = Y.reshape(T, 48, 64)
    o = omca.mca_fit(X, Y, DIMS, DIMS, "time", n_modes=k, random_state=3, use_pca=True, pca_random_state=1)
    names = ["qr", "eigh", "svd", "cholesky_ex", "cholesky", "solve_triangular", "solve", "inv", "eig", "svdvals"]
    orig = {n: getattr(torch.linalg, n) for n in names}
    calls = []
    m = xb.cross.MCA(n_modes=k, random_state=3)
    try:
        for n in names:
            setattr(torch.linalg, n, (lambda nn: lambda *a, **kw: (calls.append(nn), orig[nn](*a, **kw))[1])(n))
        m.fit(xb.DataArray(X, DIMS), xb.DataArray(Y, DIMS), dim="time")
    finally:
        for n in names:
            setattr(torch.linalg, n, orig[n])
    assert not calls, f"library factorizations on the fit path: {calls}"
    assert all(abs(a - b) <= 2 for a, b in zip(m.n_pca_modes_, o["n_pca_modes"])), (m.n_pca_modes_, o["n_pca_modes"])
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    c1, c2 = m.components()
    for c, oc in ((c1, o["components1_2d"]), (c2, o["components2_2d"])):
        dots = (c.values.reshape(-1, k) * oc).sum(axis=0)
        assert (dots >= 1 - 1e-4).all(), dots


@pytest.mark.parametrize("cls,alpha", [("CCA", (0.0, 0.0)), ("RDA", (0.0, 1.0)), ("CPCCA", 0.2)])
def test_cpcca_family_matches_oracle(cls, alpha):
    """CCA / RDA / CPCCA on the device: fractional whitening of the PCA scores (preprocessing/whitener.py:111-133),
    un-whitened patterns in physical space, against the numpy restatement."""
    import xeofs_b200 as xb
    T, S1, S2, k = 300, 40 * 30, 20 * 36, 3
    X, Y = _coupled_fields(T, S1, S2, 2 * k, seed=17)
    X = X.reshape(T, 40, 30)
    Y = Y.reshape(T, 20, 36)
    o = omca.mca_fit(X, Y, DIMS, DIMS, "time", n_modes=k, random_state=3, use_pca=True, n_pca_modes=2 * k,
                     pca_random_state=1, alpha=alpha)
    kw = dict(n_modes=k, random_state=3, n_pca_modes=2 * k)
    m = xb.cross.CPCCA(alpha=alpha, **kw) if cls == "CPCCA" else getattr(xb.cross, cls)(**kw)
    m.fit(xb.DataArray(X, DIMS), xb.DataArray(Y, DIMS), dim="time")
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    c1, c2 = m.components()
    for c, oc in ((c1, o["components1_2d"]), (c2, o["components2_2d"])):
        V = c.values.reshape(-1, k)
        cosang = (V * oc).sum(axis=0) / np.linalg.norm(V, axis=0) / np.linalg.norm(oc, axis=0)
        assert (cosang > 1 - 1e-4).all(), cosang
        np.testing.assert_allclose(np.linalg.norm(V, axis=0), np.linalg.norm(oc, axis=0), rtol=1e-3)
    s1, s2 = m.scores()
    for sc, osc in ((s1, o["scores1"]), (s2, o["scores2"])):
        scale = np.abs(osc).max(axis=0)
        np.testing.assert_allclose(sc.values / scale, osc / scale, atol=2e-3)


@pytest.mark.parametrize("use_pca", [False, True])
def test_mca_patterns_and_score_statistics(use_pca):
    """homogeneous / heterogeneous patterns with p-values, squared covariance fraction and score correlations
    (cpcca.py:331-512, 726-898) on the device against the numpy restatement."""
    import xeofs_b200 as xb
    T, S1, S2, k = 300, 40 * 30, 20 * 36, 4
    X, Y = _coupled_fields(T, S1, S2, 2 * k, seed=23)
    X = X.reshape(T, 40, 30)
    Y = Y.reshape(T, 20, 36)
    X[:, 3, 5] = np.nan
    kw = dict(n_modes=k, random_state=3, use_pca=use_pca, n_pca_modes=2 * k)
    o = omca.mca_fit(X, Y, DIMS, DIMS, "time", pca_random_state=1, **kw)
    m = xb.cross.MCA(**kw).fit(xb.DataArray(X, DIMS), xb.DataArray(Y, DIMS), dim="time")
    np.testing.assert_allclose(m.squared_covariance_fraction().values, o["squared_covariance_fraction"], rtol=1e-3)
    np.testing.assert_allclose(m.cross_correlation_coefficients().values, o["cross_correlation_coefficients"], rtol=1e-4)
    (h1, h2), (p1, p2) = m.homogeneous_patterns()
    (g1, g2), _ = m.heterogeneous_patterns()
    valid1 = o["fitted1"]["is_valid_feature"]
    for got, ref, valid in ((h1, o["homogeneous_patterns"][0], valid1), (h2, o["homogeneous_patterns"][1], None),
                            (g1, o["heterogeneous_patterns"][0], valid1), (g2, o["heterogeneous_patterns"][1], None),
                            (p1, o["pvalues_homogeneous"][0], valid1), (p2, o["pvalues_homogeneous"][1], None)):
        v = got.values.reshape(-1, k)
        if valid is not None:
            assert np.isnan(v[~valid]).all()
            v = v[valid]
        np.testing.assert_allclose(v, ref, atol=5e-4)


def test_mca_total_squared_covariance_wide_fields():
    """cpcca.py:991-1000 at a width (S >= 2^16) where the sample Gram matrices come from the bf16 tcgen05 GEMM
    (csrc/gram_bf16.cu): sum |C|^2 = <A1 A1^T, A2 A2^T> / (n-1)^2 against fp64 torch on the preprocessed fields."""
    import xeofs_b200 as xb
    from xeofs_b200 import _lib
    T, S1, S2, k = 300, 70000, 66000, 4
    X, Y = _coupled_fields(T, S1, S2, 2 * k, seed=11)
    X = X.reshape(T, 70, 1000)
    Y = Y.reshape(T, 66, 1000)
    m = xb.cross.MCA(n_modes=k, random_state=3, use_pca=False, total_squared_covariance=False)
    m.fit(xb.DataArray(X, DIMS), xb.DataArray(Y, DIMS), dim="time")
    assert m.ops.sum_algo == _lib.ALGO_TF32X1R
    tsc = m.total_squared_covariance()
    A1 = torch.from_numpy(X.reshape(T, -1)).cuda().double()
    A2 = torch.from_numpy(Y.reshape(T, -1)).cuda().double()
    A1 -= A1.mean(0)
    A2 -= A2.mean(0)
    ref = float(((A1 @ A1.t()) * (A2 @ A2.t())).sum()) / (T - 1) ** 2
    np.testing.assert_allclose(tsc, ref, rtol=2e-5)


@pytest.mark.parametrize("power", [1, 2])
@pytest.mark.parametrize("m_rot", [2, 10])
def test_rotator_matches_oracle(power, m_rot):
    """Varimax (power=1) / promax rotation of the leading modes: one streaming pass per iteration on the device vs
    linalg/_numpy/_rotation.py restated in the oracle (pinned by golden vectors from the reference's own source)."""
    import xeofs_b200 as xb
    T, nlat, nlon, k = 500, 30, 60, 12
    X = planted(T, nlat * nlon, 2 * k, seed=21).reshape(T, nlat, nlon)
    X[:, 4, 4] = np.nan
    coords = {"lat": np.linspace(85, -85, nlat), "lon": np.arange(nlon) * 6.0}
    kw = dict(n_modes=k, use_coslat=True, random_state=2, solver_kwargs={"n_iter": 4})
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, **kw)
    model = xb.single.EOF(**kw).fit(xb.DataArray(X, DIMS, coords), dim="time")
    r = xb.single.EOFRotator(n_modes=m_rot, power=power).fit(model)
    ro = orot.eof_rotator_fit(o["components_2d"], o["explained_variance"], o["scores"], o["norms"], o["A"].shape[0],
                              n_modes=m_rot, power=power)
    np.testing.assert_allclose(r.explained_variance().values, ro["explained_variance"], rtol=1e-4)
    np.testing.assert_allclose(r.singular_values().values, ro["norms"], rtol=1e-4)
    V = r.components().values.reshape(-1, m_rot)[o["fitted"]["is_valid_feature"]]
    dots = (V * ro["components_2d"]).sum(axis=0)
    assert (dots >= 1 - 1e-4).all(), dots
    sc = r.scores().values.reshape(-1, m_rot)
    scale = np.abs(ro["scores"]).max(axis=0)
    np.testing.assert_allclose(sc / scale, ro["scores"] / scale, atol=2e-3)
    if power == 1:  # tests/models/single/test_eof_rotator.py:98-137
        np.testing.assert_allclose(r.explained_variance().values.sum(), model.explained_variance().values[:m_rot].sum(),
                                   rtol=1e-5)
    assert r.n_iter_ < 1000


def test_rotator_tensor_core_sweep_matches_oracle():
    """EOFRotator on the tcgen05 sweep (forced: S is below the automatic threshold): same oracle, same tolerances as
    the fp64 sweep."""
    import xeofs_b200 as xb
    T, nlat, nlon, k, m_rot = 300, 120, 180, 12, 10
    X = planted(T, nlat * nlon, 2 * k, seed=23).reshape(T, nlat, nlon)
    coords = {"lat": np.linspace(85, -85, nlat), "lon": np.arange(nlon) * 2.0}
    kw = dict(n_modes=k, use_coslat=True, random_state=2, solver_kwargs={"n_iter": 4})
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, **kw)
    model = xb.single.EOF(**kw).fit(xb.DataArray(X, DIMS, coords), dim="time")
    model.ops.varimax_algo = "tc"
    assert model.ops._varimax_tc_applies(model._Vt, nlat * nlon, m_rot, False)
    r = xb.single.EOFRotator(n_modes=m_rot).fit(model)
    model.ops.varimax_algo = "simt"
    r64 = xb.single.EOFRotator(n_modes=m_rot).fit(model)
    model.ops.varimax_algo = "auto"
    # the stopping test (relative change of sum(svals) below 1e-8) sits near the fp32 noise of the sweep at this small
    # S: in a slowly converging case the tensor-core iteration may stop somewhat earlier than the fp64 one
    assert r64.n_iter_ // 2 <= r.n_iter_ <= r64.n_iter_ + max(2, r64.n_iter_ // 10), (r.n_iter_, r64.n_iter_)
    np.testing.assert_allclose(r.explained_variance().values, r64.explained_variance().values, rtol=1e-4)
    ro = orot.eof_rotator_fit(o["components_2d"], o["explained_variance"], o["scores"], o["norms"], o["A"].shape[0],
                              n_modes=m_rot, power=1)
    np.testing.assert_allclose(r.explained_variance().values, ro["explained_variance"], rtol=1e-4)
    V = r.components().values.reshape(-1, m_rot)
    dots = (V * ro["components_2d"]).sum(axis=0)
    assert (dots >= 1 - 1e-4).all(), dots
    np.testing.assert_allclose(r.explained_variance().values.sum(), model.explained_variance().values[:m_rot].sum(),
                               rtol=1e-5)


def test_rotator_not_converged_raises():
    import xeofs_b200 as xb
    X = planted(200, 600, 12, seed=5).reshape(200, 20, 30)
    model = xb.single.EOF(n_modes=6, random_state=1).fit(xb.DataArray(X, DIMS), dim="time")
    with pytest.raises(RuntimeError, match="did not converge"):
        xb.single.EOFRotator(n_modes=6, max_iter=1, rtol=1e-14).fit(model)


def test_eof_large_a_posteriori_properties():
    """At a size the fp64 oracle takes minutes for (8760 x 65 536, 2.3 GB): (s_i, v_i) are singular pairs of the
    preprocessed matrix — ||A v_i|| = s_i through the fp32 SIMT kernels (a different code path from the tcgen05 one
    the fit used), V orthonormal, scores = A V, explained variance below total variance."""
    import xeofs_b200 as xb
    from xeofs_b200 import _lib
    T, nlat, nlon, k = 8760, 64, 1024, 20
    g = torch.Generator(device="cuda").manual_seed(3)
    U = torch.linalg.qr(torch.randn((T, 2 * k), generator=g, device="cuda"))[0]
    V = torch.linalg.qr(torch.randn((nlat * nlon, 2 * k), generator=g, device="cuda"))[0]
    sig = 1e5 * 0.85 ** torch.arange(2 * k, device="cuda")
    X = 280 + (U * sig) @ V.t() + 0.05 * torch.randn((T, nlat * nlon), generator=g, device="cuda")
    coords = {"lat": np.linspace(89, -89, nlat), "lon": np.arange(nlon) * (360.0 / nlon)}
    m = xb.single.EOF(n_modes=k, use_coslat=True, random_state=5, solver_kwargs={"n_iter": 4})
    m.fit(xb.DataArray(X.reshape(T, nlat, nlon), DIMS, coords), dim="time")
    s = m.data["norms"]
    ops, f = m.ops, m.preprocessor.fitted.field
    Z = ops.project_T(f, m._Vt, k, algo=_lib.ALGO_SIMT)[:, :k].double()
    np.testing.assert_allclose(Z.norm(dim=0).cpu().numpy(), s.cpu().numpy(), rtol=1e-4)
    G = ops.gram(m._Vt, f.S, k, 1).cpu().numpy()
    np.testing.assert_allclose(G, np.eye(k), atol=1e-4)
    sc = m._scores[:, :k].double()
    scale = sc.abs().max(dim=0).values
    assert float(((Z - sc) / scale).abs().max()) < 2e-3
    evr = m.explained_variance_ratio().values
    assert evr.sum() <= 1.0 + 1e-6 and (np.diff(evr) <= 1e-9).all()
    # the planted spectrum, seen through the coslat weights, bounds the leading value
    assert 0.3 * 1e5 < float(s[0]) < 1.01 * 1e5


_DIST_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests", "golden"))
from _inputs import planted
import xeofs_b200 as xb
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
T, nlat, nlon, k = 700, 48, 96, 10
X = planted(T, nlat * nlon, 2 * k, seed=12).reshape(T, nlat, nlon)
X[:, 5, 7] = np.nan; X[:, 40, 3] = np.nan
lat = np.linspace(88, -88, nlat)
rows = slice(rank * nlat // world, (rank + 1) * nlat // world)
coords = {"lat": lat[rows], "lon": np.arange(nlon) * 3.75}
m = xb.single.EOF(n_modes=k, use_coslat=True, standardize=True, random_state=7, solver_kwargs={"n_iter": 4}, distributed=True)
m.fit(xb.DataArray(X[:, rows], ("time", "lat", "lon"), coords), dim="time")
r = xb.single.EOFRotator(n_modes=6).fit(m)
np.savez(os.path.join(sys.argv[2], f"rank{rank}.npz"), s=m.singular_values().values, comps=m.components().values,
         scores=m.scores().values, evr=m.explained_variance_ratio().values, rot_ev=r.explained_variance().values,
         collectives=m.comm.collectives)
dist.destroy_process_group()
'''


def test_feature_sharded_fit_two_gpus(tmp_path):
    """§8e on real devices: the feature axis split over 2 GPUs (NCCL allreduce of the projected blocks / Gram
    matrices) must reproduce the oracle like the single-GPU fit does."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_DIST_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", str(script), root, str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    T, nlat, nlon, k = 700, 48, 96, 10
    X = planted(T, nlat * nlon, 2 * k, seed=12).reshape(T, nlat, nlon)
    X[:, 5, 7] = np.nan
    X[:, 40, 3] = np.nan
    coords = {"lat": np.linspace(88, -88, nlat), "lon": np.arange(nlon) * 3.75}
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, n_modes=k, use_coslat=True, standardize=True, random_state=7,
                     solver_kwargs={"n_iter": 4})
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert int(r0["collectives"]) > 0
    for r in (r0, r1):
        np.testing.assert_allclose(r["s"], o["singular_values"], rtol=1e-4)
        np.testing.assert_allclose(r["evr"], o["explained_variance_ratio"], rtol=1e-4)
    comps = np.concatenate([r0["comps"], r1["comps"]], axis=0).reshape(-1, k)
    vf = o["fitted"]["is_valid_feature"]
    dots = (comps[vf] * o["components_2d"]).sum(axis=0)
    assert (dots >= 1 - 1e-4).all(), dots
    ro = orot.eof_rotator_fit(o["components_2d"], o["explained_variance"], o["scores"], o["norms"], o["A"].shape[0], n_modes=6)
    np.testing.assert_allclose(r0["rot_ev"], ro["explained_variance"], rtol=1e-4)


def test_mca_through_the_half_precision_copies(monkeypatch):
    """The implicit cross-covariance products of the power iterations on the fp16 copies of both fields (forced on for
    this small case) against the explicit-C oracle: same tolerances as test_mca_matches_explicit_cross_covariance_oracle."""
    import xeofs_b200 as xb
    from xeofs_b200._cuda_ops import CudaOps
    monkeypatch.setattr(CudaOps, "h16_min_bytes", 0)
    T, S1, S2, k = 400, 40 * 30, 60 * 50, 8
    X, Y = _coupled_fields(T, S1, S2, 2 * k, seed=5)
    X, Y = X.reshape(T, 40, 30), Y.reshape(T, 60, 50)
    cx = {"lat": np.linspace(80, -80, 40), "lon": np.arange(30) * 1.0}
    cy = {"lat": np.linspace(60, -60, 60), "lon": np.arange(50) * 1.0}
    kw = dict(standardize=True, use_coslat=True)
    o = omca.mca_fit(X, Y, DIMS, DIMS, "time", coords_x=cx, coords_y=cy, n_modes=k, random_state=3, **kw)
    m = xb.cross.MCA(n_modes=k, random_state=3, use_pca=False, **kw)
    m.fit(xb.DataArray(X, DIMS, cx), xb.DataArray(Y, DIMS, cy), dim="time")
    assert m._f1.field.h16 is not None and m._f2.field.h16 is not None
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    c1, c2 = m.components()
    for c, oc in ((c1, o["components1_2d"]), (c2, o["components2_2d"])):
        dots = (c.values.reshape(-1, k) * oc).sum(axis=0)
        assert (dots >= 1 - 1e-4).all(), dots
