"""CPU: the host side of the drop-in (preprocessor, range-finder driver, EOF / MCA / EOFRotator classes) driven through
a torch-CPU test double of the kernel interface (tests/cpu_ops.py) against the oracle, the C-ABI library's symbol
table, and the feature-sharded path over torch.distributed (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from _inputs import MOCK_LAT, MOCK_LON, mock_data_array, planted
from cpu_ops import TorchCpuOps
from oracle import eof as oeof
from oracle import mca as omca
from oracle import rotation as orot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIMS = ("time", "lat", "lon")


def make_ops():
    """The kernel interface the model classes run on in this module: the torch-CPU test double.  tests/test_gpu_twins.py
    re-runs the cross-model cases below with this replaced by None (= the real CudaOps)."""
    return TorchCpuOps()


# ---------------------------------------------------------------- the C-ABI library
def test_library_exports_every_declared_symbol():
    """include/xeofs_b200.h is the contract: every function it declares is exported by the built library and
    bound (with a signature) by xeofs_b200/_lib.py.  No compute call is made (no GPU here)."""
    from xeofs_b200 import _lib
    header = open(os.path.join(ROOT, "include", "xeofs_b200.h")).read()
    declared = set(re.findall(r"\b(xeofs_b200_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().xeofs_b200_version() >= 100


def test_product_path_has_no_cpu_fallback():
    import torch
    from xeofs_b200._cuda_ops import CudaOps
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CudaOps()


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xeofs_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


# ---------------------------------------------------------------- EOF host logic vs oracle
def _compare_eof(o, m, k, tol=1e-6):
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-5)
    np.testing.assert_allclose(m.explained_variance_ratio().values, o["explained_variance_ratio"], rtol=1e-5)
    comps = m.components().values
    np.testing.assert_array_equal(np.isnan(comps), np.isnan(o["components"]))
    vf = o["fitted"]["is_valid_feature"]
    dots = (comps.reshape(-1, k)[vf] * o["components_2d"]).sum(axis=0)
    assert (dots > 1 - 1e-5).all(), dots
    sc = m.scores().values.reshape(-1, k)
    vs = o["fitted"]["is_valid_sample"]
    scale = np.abs(o["scores"]).max(axis=0)
    np.testing.assert_allclose(sc[vs] / scale, o["scores"] / scale, atol=2e-4)
    assert np.isnan(sc[~vs]).all()


@pytest.mark.parametrize("kw", [dict(), dict(standardize=True, use_coslat=True), dict(center=False)])
def test_eof_host_logic_wide_with_nans(kw):
    import xeofs_b200 as xb
    T, nlat, nlon, k = 120, 12, 30, 6
    X = planted(T, nlat * nlon, 2 * k, seed=1).reshape(T, nlat, nlon)
    X[:, np.random.default_rng(9).random((nlat, nlon)) < 0.1] = np.nan
    X[17] = np.nan
    coords = {"lat": np.linspace(80, -80, nlat), "lon": np.arange(nlon) * 12.0}
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, n_modes=k, random_state=5, solver_kwargs={"n_iter": 4}, **kw)
    m = xb.single.EOF(n_modes=k, random_state=5, solver_kwargs={"n_iter": 4}, ops=TorchCpuOps(), **kw)
    m.fit(xb.DataArray(X, DIMS, coords), dim="time")
    _compare_eof(o, m, k)


def test_eof_host_logic_tall_and_exact_policy():
    import xeofs_b200 as xb
    X = mock_data_array().astype(np.float32)
    coords = {"lat": MOCK_LAT, "lon": MOCK_LON}
    for k in (3, 18):  # 18 > 0.8 * rank: exact policy (decomposer.py:112-131)
        o = oeof.eof_fit(X, DIMS, "time", coords=coords, n_modes=k, random_state=5)
        m = xb.single.EOF(n_modes=k, random_state=5, ops=TorchCpuOps()).fit(xb.DataArray(X, DIMS, coords), dim="time")
        np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)


def test_error_conventions():
    import xeofs_b200 as xb
    X = mock_data_array().astype(np.float32)
    da = xb.DataArray(X, DIMS, {"lat": MOCK_LAT, "lon": MOCK_LON})
    with pytest.raises(ValueError, match="rank"):
        xb.single.EOF(n_modes=21, ops=TorchCpuOps()).fit(da, dim="time")
    with pytest.raises(TypeError):
        xb.single.EOF(n_modes=2, ops=TorchCpuOps()).fit(X, dim="time")
    Xn = X.copy()
    Xn[3, 2, 1] = np.nan
    with pytest.raises(ValueError, match="partial NaN"):
        xb.single.EOF(n_modes=2, ops=TorchCpuOps()).fit(xb.DataArray(Xn, DIMS), dim="time")
    with pytest.raises(ValueError, match="latitude"):
        xb.single.EOF(n_modes=2, use_coslat=True, ops=TorchCpuOps()).fit(xb.DataArray(X, ("time", "y", "x")), dim="time")
    with pytest.raises(NotImplementedError):
        xb.cross.MCA(n_modes=2, use_pca=(True, False), ops=TorchCpuOps())


def test_transform_inverse_transform_host_logic():
    import xeofs_b200 as xb
    X = mock_data_array().astype(np.float32)
    da = xb.DataArray(X, DIMS, {"lat": MOCK_LAT, "lon": MOCK_LON})
    m = xb.single.EOF(n_modes=20, standardize=True, use_coslat=True, solver="full", ops=TorchCpuOps()).fit(da, dim="time")
    np.testing.assert_allclose(m.transform(da).values, m.scores().values, rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(m.inverse_transform(m.scores()).values, X, rtol=1e-4, atol=2e-5)


# ---------------------------------------------------------------- MCA / rotator host logic vs oracle
def _check_patterns(m, o, valid1=None, atol=2e-4):
    """homogeneous / heterogeneous patterns and p-values (cpcca.py:726-898) against the oracle."""
    (h1, h2), (p1, p2) = m.homogeneous_patterns()
    (g1, g2), _ = m.heterogeneous_patterns()
    for got, ref, valid in ((h1, o["homogeneous_patterns"][0], valid1), (h2, o["homogeneous_patterns"][1], None),
                            (g1, o["heterogeneous_patterns"][0], valid1), (g2, o["heterogeneous_patterns"][1], None),
                            (p1, o["pvalues_homogeneous"][0], valid1), (p2, o["pvalues_homogeneous"][1], None)):
        v = got.values.reshape(-1, got.values.shape[-1])
        if valid is not None:
            assert np.isnan(v[~valid]).all()
            v = v[valid]
        np.testing.assert_allclose(v, ref, atol=atol)
    with pytest.raises(NotImplementedError, match="statsmodels"):
        m.homogeneous_patterns(correction="fdr_bh")


def test_mca_host_logic():
    import xeofs_b200 as xb
    T, k = 150, 5
    rng = np.random.default_rng(0)
    U = np.linalg.qr(rng.standard_normal((T, 2 * k)))[0]
    sig = 100 * 0.7 ** np.arange(2 * k)
    X = ((U * sig) @ np.linalg.qr(rng.standard_normal((80, 2 * k)))[0].T + 0.01 * rng.standard_normal((T, 80))).astype(np.float32)
    Y = ((U * sig) @ np.linalg.qr(rng.standard_normal((60, 2 * k)))[0].T + 0.01 * rng.standard_normal((T, 60))).astype(np.float32)
    X[:, 5] = np.nan
    o = omca.mca_fit(X, Y, ("time", "x"), ("time", "y"), "time", n_modes=k, random_state=3)
    m = xb.cross.MCA(n_modes=k, random_state=3, use_pca=False, ops=TorchCpuOps())
    m.fit(xb.DataArray(X, ("time", "x")), xb.DataArray(Y, ("time", "y")), dim="time")
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-5)
    np.testing.assert_allclose(m.total_squared_covariance(), o["total_squared_covariance"], rtol=1e-5)
    np.testing.assert_allclose(m.squared_covariance_fraction().values, o["squared_covariance_fraction"], rtol=1e-4)
    np.testing.assert_allclose(m.cross_correlation_coefficients().values, o["cross_correlation_coefficients"], rtol=1e-4)
    np.testing.assert_allclose(m.correlation_coefficients_X(), o["correlation_coefficients_X"], atol=1e-4)
    _check_patterns(m, o, valid1=~np.isnan(X[0]))
    with pytest.warns(UserWarning, match="sensitive to the number of modes"):
        cf = m.covariance_fraction_CD95().values
    np.testing.assert_allclose(cf, o["singular_values"] / o["singular_values"].sum(), rtol=1e-5)
    c1, c2 = m.components()
    v1 = c1.values[~np.isnan(c1.values).any(axis=1)]
    dots = np.abs((v1 * o["components1_2d"]).sum(axis=0))
    assert (dots > 1 - 1e-5).all(), dots
    assert ((c2.values * o["components2_2d"]).sum(axis=0) > 1 - 1e-5).all()


def test_mca_pca_stage_host_logic():
    """use_pca=True (the reference's default): PCA projection of both fields (preprocessing/pca.py:94-131 with the
    variance cut of linalg/_numpy/_svd.py:214-241), MCA on the scores, patterns mapped back to physical space."""
    import xeofs_b200 as xb
    T, k = 150, 4
    rng = np.random.default_rng(2)
    U = np.linalg.qr(rng.standard_normal((T, 2 * k)))[0]
    sig = 100 * 0.7 ** np.arange(2 * k)
    X = ((U * sig) @ np.linalg.qr(rng.standard_normal((300, 2 * k)))[0].T + 0.01 * rng.standard_normal((T, 300))).astype(np.float32)
    Y = ((U * sig) @ np.linalg.qr(rng.standard_normal((200, 2 * k)))[0].T + 0.01 * rng.standard_normal((T, 200))).astype(np.float32)
    X[:, 5] = np.nan
    for npm in (0.999, 12):
        o = omca.mca_fit(X, Y, ("time", "x"), ("time", "y"), "time", n_modes=k, random_state=3, use_pca=True,
                         n_pca_modes=npm, pca_random_state=1)
        m = xb.cross.MCA(n_modes=k, random_state=3, n_pca_modes=npm, ops=TorchCpuOps())   # use_pca defaults to True
        m.fit(xb.DataArray(X, ("time", "x")), xb.DataArray(Y, ("time", "y")), dim="time")
        # the 99.9 % cut falls among noise-level modes, where the reference's randomized PCA and the exact one may
        # differ by a mode or two; an integer n_pca_modes is kept exactly
        assert all(abs(a - b) <= (2 if isinstance(npm, float) else 0) for a, b in zip(m.n_pca_modes_, o["n_pca_modes"]))
        np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
        np.testing.assert_allclose(m.total_squared_covariance(), o["total_squared_covariance"], rtol=1e-4)
        c1, c2 = m.components()
        v1 = c1.values[~np.isnan(c1.values).any(axis=1)]
        assert ((v1 * o["components1_2d"]).sum(axis=0) > 1 - 1e-4).all()
        assert ((c2.values * o["components2_2d"]).sum(axis=0) > 1 - 1e-4).all()
        s1, s2 = m.scores()
        for sc, osc in ((s1, o["scores1"]), (s2, o["scores2"])):
            scale = np.abs(osc).max(axis=0)
            np.testing.assert_allclose(sc.values / scale, osc / scale, atol=1e-3)
        if npm == 12:
            _check_patterns(m, o, valid1=~np.isnan(X[0]))


@pytest.mark.parametrize("cls,alpha", [("CCA", (0.0, 0.0)), ("RDA", (0.0, 1.0)), ("CPCCA", 0.2)])
def test_cpcca_family_host_logic(cls, alpha):
    """CCA / RDA / CPCCA (cross/cca.py, rda.py, cpcca.py): MCA on fractionally whitened PCA scores
    (preprocessing/whitener.py:111-133), patterns un-whitened and mapped back to physical space."""
    import xeofs_b200 as xb
    T, k = 150, 3
    rng = np.random.default_rng(3)
    U = np.linalg.qr(rng.standard_normal((T, 2 * k)))[0]
    sig = 100 * 0.7 ** np.arange(2 * k)
    X = ((U * sig) @ np.linalg.qr(rng.standard_normal((300, 2 * k)))[0].T + 0.05 * rng.standard_normal((T, 300))).astype(np.float32)
    Y = ((U * sig) @ np.linalg.qr(rng.standard_normal((200, 2 * k)))[0].T + 0.05 * rng.standard_normal((T, 200))).astype(np.float32)
    o = omca.mca_fit(X, Y, ("time", "x"), ("time", "y"), "time", n_modes=k, random_state=3, use_pca=True,
                     n_pca_modes=6, pca_random_state=1, alpha=alpha)
    # (6 = the planted rank: whitening weights every retained principal component alike, and the directions of
    # noise-level components — nearly degenerate — are not reproducible between two PCA solvers)
    kw = dict(n_modes=k, random_state=3, n_pca_modes=6, ops=make_ops())
    m = xb.cross.CPCCA(alpha=alpha, **kw) if cls == "CPCCA" else getattr(xb.cross, cls)(**kw)
    m.fit(xb.DataArray(X, ("time", "x")), xb.DataArray(Y, ("time", "y")), dim="time")
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    c1, c2 = m.components()
    for c, oc in ((c1, o["components1_2d"]), (c2, o["components2_2d"])):
        V, R = c.values, oc
        cosang = (V * R).sum(axis=0) / np.linalg.norm(V, axis=0) / np.linalg.norm(R, axis=0)
        assert (cosang > 1 - 1e-4).all(), cosang
        np.testing.assert_allclose(np.linalg.norm(V, axis=0), np.linalg.norm(R, axis=0), rtol=1e-3)
    s1, s2 = m.scores()
    for sc, osc in ((s1, o["scores1"]), (s2, o["scores2"])):
        scale = np.abs(osc).max(axis=0)
        np.testing.assert_allclose(sc.values / scale, osc / scale, atol=1e-3)
    # total squared covariance of the un-whitened cross-covariance, squared covariance fraction, score correlations
    np.testing.assert_allclose(m.total_squared_covariance(), o["total_squared_covariance"], rtol=1e-4)
    np.testing.assert_allclose(m.squared_covariance_fraction().values, o["squared_covariance_fraction"], rtol=1e-3,
                               atol=1e-6)
    np.testing.assert_allclose(m.cross_correlation_coefficients().values, o["cross_correlation_coefficients"], rtol=1e-4)
    np.testing.assert_allclose(m.correlation_coefficients_X(), o["correlation_coefficients_X"], atol=1e-4)
    # (full whitening makes the canonical correlations of the planted factors nearly equal — 1.006675, 1.006633,
    # 1.006572 here — so the individual modes are defined only up to fp32 noise / gap ~ 1e-7 / 4e-5 of rotation inside
    # that cluster: the PCA scores are fp32 on the device)
    _check_patterns(m, o, atol=2e-3 if cls == "CCA" else 5e-4)
    # components(normalized=False) = components * norm (cpcca.py:308-316)
    u1, u2 = m.components(normalized=False)
    np.testing.assert_allclose(u1.values, c1.values * o["norm1"], rtol=1e-3, atol=1e-6 * o["norm1"].max())
    np.testing.assert_allclose(u2.values, c2.values * o["norm2"], rtol=1e-3, atol=1e-6 * o["norm2"].max())
    # rotation of the whitened model (cross/cpcca_rotator.py:122-305): norms in the whitened PCA space
    for power in (1, 2):
        r = xb.cross.CPCCARotator(n_modes=2, power=power).fit(m)
        ro = orot.mca_rotator_fit(o["components1_2d"], o["components2_2d"], o["singular_values"], o["scores1"],
                                  o["scores2"], n_modes=2, power=power, model_components=o["components_model"])
        np.testing.assert_allclose(r.squared_covariance().values, ro["squared_covariance"], rtol=1e-4)
        rc1, rc2 = r.components()
        for c, oc in ((rc1, ro["components1_2d"]), (rc2, ro["components2_2d"])):
            np.testing.assert_allclose(c.values, oc, atol=2e-4 * np.abs(oc).max())
        rs1, _ = r.scores()
        np.testing.assert_allclose(rs1.values, ro["scores1"], atol=2e-3 * np.abs(ro["scores1"]).max())
        # transform of the rotator of a whitened model (cpcca_rotator.py:322-427): the reference projects on the
        # un-whitened physical components, so the training data do NOT reproduce the fitted scores — follow it
        t1, t2 = r.transform(X=xb.DataArray(X, ("time", "x")), Y=xb.DataArray(Y, ("time", "y")))
        for t, ref in ((t1, ro["transform1"](o["fitted1"]["A"])), (t2, ro["transform2"](o["fitted2"]["A"]))):
            np.testing.assert_allclose(t.values, ref, atol=2e-3 * np.abs(ref).max())
        # normalized=False scales the components by norm1 / norm2 (cpcca.py:308-316)
        u1, u2 = r.components(normalized=False)
        np.testing.assert_allclose(u1.values, rc1.values * ro["norm1"], rtol=1e-3, atol=1e-6 * ro["norm1"].max())
    with pytest.raises(NotImplementedError, match="use_pca"):
        xb.cross.CCA(n_modes=k, use_pca=False, ops=make_ops()).fit(
            xb.DataArray(X, ("time", "x")), xb.DataArray(Y, ("time", "y")), dim="time")


@pytest.mark.parametrize("mode", ["implicit", "pca", "cca"])
def test_cross_transform_predict_inverse_host_logic(mode):
    """transform / predict / inverse_transform of the cross models (cpcca.py:227-306 behind the back-transforms of
    base_model_cross_set.py:323-463) against the oracle, on unseen samples."""
    import xeofs_b200 as xb
    from oracle import preprocess as opp
    T, k = 150, 3
    rng = np.random.default_rng(5)
    U = np.linalg.qr(rng.standard_normal((T, 2 * k)))[0]
    sig = 100 * 0.7 ** np.arange(2 * k)
    X = ((U * sig) @ np.linalg.qr(rng.standard_normal((300, 2 * k)))[0].T + 0.05 * rng.standard_normal((T, 300)) + 7).astype(np.float32)
    Y = ((U * sig) @ np.linalg.qr(rng.standard_normal((200, 2 * k)))[0].T + 0.05 * rng.standard_normal((T, 200)) - 3).astype(np.float32)
    X[:, 5] = np.nan
    dx, dy = ("time", "x"), ("time", "y")
    kw = dict(n_modes=k, random_state=3, standardize=True)
    okw = dict(kw)
    if mode == "implicit":
        cls, kw2, okw2 = xb.cross.MCA, dict(use_pca=False), dict(use_pca=False)
    elif mode == "pca":
        cls, kw2, okw2 = xb.cross.MCA, dict(n_pca_modes=6), dict(use_pca=True, n_pca_modes=6, pca_random_state=1)
    else:
        cls, kw2, okw2 = xb.cross.CCA, dict(n_pca_modes=6), dict(use_pca=True, n_pca_modes=6, pca_random_state=1, alpha=0.0)
    o = omca.mca_fit(X, Y, dx, dy, "time", **okw, **okw2)
    m = cls(ops=make_ops(), **kw, **kw2).fit(xb.DataArray(X, dx), xb.DataArray(Y, dy), dim="time")
    # training data reproduce the scores
    s1, s2 = m.scores()
    t1, t2 = m.transform(X=xb.DataArray(X, dx), Y=xb.DataArray(Y, dy))
    np.testing.assert_allclose(t1.values, s1.values, atol=2e-3 * np.abs(s1.values).max())
    np.testing.assert_allclose(t2.values, s2.values, atol=2e-3 * np.abs(s2.values).max())
    # unseen samples
    Xn = X[:40] + (0.3 * rng.standard_normal((40, 300))).astype(np.float32)
    An = opp.transform_new(Xn, dx, o["fitted1"], True, True, False)
    ref = o["helpers"]["transform1"](An)
    got = m.transform(X=xb.DataArray(Xn, dx)).values
    np.testing.assert_allclose(got, ref, atol=2e-3 * np.abs(ref).max())
    refp = o["helpers"]["predict"](An)
    gotp = m.predict(xb.DataArray(Xn, dx)).values
    np.testing.assert_allclose(gotp, refp, atol=2e-3 * np.abs(refp).max())
    # reconstruction from the scores: scores . components^H, un-scaled (NaN at the dropped feature)
    rec = m.inverse_transform(X=s1).values
    ref_rec = opp.inverse_scale(o["scores1"] @ o["components1_2d"].T, o["fitted1"], True, True, False)
    assert np.isnan(rec[:, 5]).all()
    keep = ~np.isnan(X[0])
    np.testing.assert_allclose(rec[:, keep], ref_rec[:, keep] if ref_rec.shape[1] == X.shape[1] else ref_rec,
                               atol=2e-3 * np.abs(ref_rec[np.isfinite(ref_rec)]).max())
    with pytest.raises(ValueError, match="Either X or Y"):
        m.transform()


def test_mca_rotator_host_logic():
    """cross/cpcca_rotator.py:122-305 (identity whitening, no PCA) against its numpy restatement."""
    import xeofs_b200 as xb
    T, k, mr = 150, 6, 4
    rng = np.random.default_rng(1)
    U = np.linalg.qr(rng.standard_normal((T, 2 * k)))[0]
    sig = 100 * 0.8 ** np.arange(2 * k)
    X = ((U * sig) @ np.linalg.qr(rng.standard_normal((83, 2 * k)))[0].T + 0.01 * rng.standard_normal((T, 83))).astype(np.float32)
    Y = ((U * sig) @ np.linalg.qr(rng.standard_normal((60, 2 * k)))[0].T + 0.01 * rng.standard_normal((T, 60))).astype(np.float32)
    X[:, 7] = np.nan
    o = omca.mca_fit(X, Y, ("time", "x"), ("time", "y"), "time", n_modes=k, random_state=3)
    m = xb.cross.MCA(n_modes=k, random_state=3, use_pca=False, ops=make_ops())
    m.fit(xb.DataArray(X, ("time", "x")), xb.DataArray(Y, ("time", "y")), dim="time")
    for power in (1, 2):
        r = xb.cross.MCARotator(n_modes=mr, power=power).fit(m)
        ro = orot.mca_rotator_fit(o["components1_2d"], o["components2_2d"], o["singular_values"], o["scores1"],
                                  o["scores2"], n_modes=mr, power=power)
        np.testing.assert_allclose(r.squared_covariance().values, ro["squared_covariance"], rtol=1e-4)
        np.testing.assert_allclose(r.data["norm1"].cpu().numpy(), ro["norm1"], rtol=1e-4)
        c1, c2 = r.components()
        v1 = c1.values[~np.isnan(c1.values).any(axis=1)]
        assert ((v1 * ro["components1_2d"]).sum(axis=0) > 1 - 1e-5).all()
        assert ((c2.values * ro["components2_2d"]).sum(axis=0) > 1 - 1e-5).all()
        s1, s2 = r.scores()
        for sc, osc in ((s1, ro["scores1"]), (s2, ro["scores2"])):
            scale = np.abs(osc).max(axis=0)
            np.testing.assert_allclose(sc.values / scale, osc / scale, atol=1e-3)
        # transform of the training data reproduces the rotated scores (cpcca_rotator.py:322-427)
        dax, day = xb.DataArray(X, ("time", "x")), xb.DataArray(Y, ("time", "y"))
        t1, t2 = r.transform(X=dax, Y=day)
        np.testing.assert_allclose(t1.values, s1.values, atol=2e-3 * np.abs(s1.values).max())
        np.testing.assert_allclose(t2.values, s2.values, atol=2e-3 * np.abs(s2.values).max())
    with pytest.raises(ValueError, match="exceeds"):
        xb.cross.MCARotator(n_modes=k + 1).fit(m)


def test_rotator_host_logic():
    import xeofs_b200 as xb
    X = planted(200, 300, 12, seed=4).reshape(200, 10, 30)
    coords = {"lat": np.linspace(60, -60, 10), "lon": np.arange(30) * 12.0}
    ops = TorchCpuOps()
    m = xb.single.EOF(n_modes=8, random_state=1, ops=ops).fit(xb.DataArray(X, DIMS, coords), dim="time")
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, n_modes=8, random_state=1)
    for power in (1, 2):
        r = xb.single.EOFRotator(n_modes=5, power=power).fit(m)
        ro = orot.eof_rotator_fit(o["components_2d"], o["explained_variance"], o["scores"], o["norms"],
                                  o["A"].shape[0], n_modes=5, power=power)
        np.testing.assert_allclose(r.explained_variance().values, ro["explained_variance"], rtol=5e-5)
        V = r.components().values.reshape(-1, 5)
        dots = (V * ro["components_2d"]).sum(axis=0)
        assert (dots > 1 - 1e-5).all(), dots
        if power == 1:  # an orthogonal rotation reconstructs what the un-rotated modes reconstruct
            rec_r = r.inverse_transform(r.scores()).values
            sc = m.scores()
            rec_m = m.inverse_transform(xb.DataArray(sc.values[:, :5], ("time", "mode"), {"mode": np.arange(1, 6)})).values
            np.testing.assert_allclose(rec_r, rec_m, rtol=1e-4, atol=1e-4 * np.abs(rec_m).max())
        if power == 1:  # rotation conserves the variance (tests/models/single/test_eof_rotator.py:98-137)
            np.testing.assert_allclose(r.explained_variance().values.sum(), m.explained_variance().values[:5].sum(),
                                       rtol=1e-5)


def test_eof_variance_based_n_modes_weights_and_tmode():
    """Less-travelled corners of the fit API against the oracle: float n_modes (decomposer.py:88-94, 188-216), user
    weights broadcast over the feature dims (scaler.py:100-126), two sample dimensions, and T-mode analysis
    (dim = the spatial dims, docs .../plot_eof-tmode.py:22-23)."""
    import xeofs_b200 as xb
    T, nlat, nlon = 120, 6, 10
    X = planted(T, nlat * nlon, 10, seed=21).reshape(T, nlat, nlon)
    coords = {"lat": np.linspace(50, -50, nlat), "lon": np.arange(nlon) * 10.0}
    # (1) n_modes as a fraction of variance
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, n_modes=0.9, random_state=3)
    m = xb.single.EOF(n_modes=0.9, random_state=3, ops=TorchCpuOps()).fit(xb.DataArray(X, DIMS, coords), dim="time")
    assert m.singular_values().values.shape == o["singular_values"].shape
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    # (2) weights along one feature dim, together with coslat weights
    w = np.linspace(0.5, 2.0, nlon)
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, n_modes=4, use_coslat=True, random_state=3,
                     weights=np.broadcast_to(w, (nlat, nlon)))
    m = xb.single.EOF(n_modes=4, use_coslat=True, random_state=3, ops=TorchCpuOps())
    m.fit(xb.DataArray(X, DIMS, coords), dim="time", weights=xb.DataArray(w, ("lon",)))
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    V = m.components().values.reshape(-1, 4)
    assert ((V * o["components_2d"]).sum(axis=0) > 1 - 1e-5).all()
    with pytest.raises(ValueError, match="not feature dimensions"):
        xb.single.EOF(n_modes=2, ops=TorchCpuOps()).fit(xb.DataArray(X, DIMS, coords), dim="time",
                                                        weights=xb.DataArray(np.ones(T), ("time",)))
    # (3) two sample dimensions
    X4 = X.reshape(10, 12, nlat, nlon)
    dims4 = ("year", "month", "lat", "lon")
    o = oeof.eof_fit(X4, dims4, ("year", "month"), coords=coords, n_modes=4, random_state=3)
    m = xb.single.EOF(n_modes=4, random_state=3, ops=TorchCpuOps()).fit(xb.DataArray(X4, dims4, coords), dim=("year", "month"))
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    assert m.scores().values.shape == (10, 12, 4) and m.components().values.shape == (nlat, nlon, 4)
    # (4) T-mode: the spatial dims are the samples, time is the feature axis
    o = oeof.eof_fit(X, DIMS, ("lat", "lon"), coords=coords, n_modes=4, random_state=3)
    m = xb.single.EOF(n_modes=4, random_state=3, ops=TorchCpuOps()).fit(xb.DataArray(X, DIMS, coords), dim=("lat", "lon"))
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    assert m.components().values.shape == (T, 4) and m.scores().values.shape == (nlat, nlon, 4)
    np.testing.assert_allclose(np.abs((m.components().values * o["components_2d"]).sum(axis=0)), 1.0, atol=1e-5)


def test_eof_more_modes_than_one_kernel_block():
    """n_modes + oversamples beyond the 128-column width of the device kernels: the engine runs the same algorithm,
    the device layer splits the products into column blocks (exercised on the GPU by its twin)."""
    import xeofs_b200 as xb
    T, S, k = 300, 400, 130
    X = planted(T, S, 40, seed=33).reshape(T, 20, 20)
    o = oeof.eof_fit(X, DIMS, "time", n_modes=k, random_state=3)
    m = xb.single.EOF(n_modes=k, random_state=3, ops=TorchCpuOps()).fit(xb.DataArray(X, DIMS), dim="time")
    np.testing.assert_allclose(m.singular_values().values[:30], o["singular_values"][:30], rtol=1e-4)
    V = m.components().values.reshape(-1, k)
    np.testing.assert_allclose(V.T @ V, np.eye(k), atol=1e-4)


def test_one_normalisation_per_power_iteration_matches_sklearn_schedule():
    """From the second power iteration on only the short side is normalised (_engine.randomized_svd); on a steep
    spectrum (sigma_j = 0.3^j, ratio 2e-5 over the first ten modes) the result equals the reference's schedule —
    a normalisation after every half-step — and the oracle, down to the modes the fp32 input resolves."""
    import xeofs_b200 as xb
    from xeofs_b200 import _engine as E
    rng = np.random.default_rng(0)
    T, S, r, k = 400, 3000, 30, 12
    U = np.linalg.qr(rng.standard_normal((T, r)))[0]
    V = np.linalg.qr(rng.standard_normal((S, r)))[0]
    X = (280 + (U * (1e4 * 0.3 ** np.arange(r))) @ V.T + 1e-9 * rng.standard_normal((T, S))).astype(np.float32)
    o = oeof.eof_fit(X, ("time", "x"), "time", n_modes=k, random_state=3, solver_kwargs={"n_iter": 4})
    res = {}
    try:
        for both in (False, True):
            E.NORMALIZE_BOTH_HALF_STEPS = both
            m = xb.single.EOF(n_modes=k, random_state=3, solver_kwargs={"n_iter": 4}, ops=TorchCpuOps())
            res[both] = m.fit(xb.DataArray(X, ("time", "x")), dim="time").singular_values().values
    finally:
        E.NORMALIZE_BOTH_HALF_STEPS = False
    np.testing.assert_allclose(res[False][:10], res[True][:10], rtol=1e-6)
    np.testing.assert_allclose(res[False][:10], o["singular_values"][:10], rtol=1e-4)


def test_eof_list_input_host_logic():
    """A list of arrays (two variables on different grids): each scaled on its own, concatenated along the feature
    axis (preprocessing/preprocessor.py:208-228, concatenator.py:58-81); components come back one array per input."""
    import xeofs_b200 as xb
    T, k = 150, 4
    full = planted(T, 8 * 9 + 5 * 7, 2 * k, seed=12)
    X1 = full[:, :72].reshape(T, 8, 9).copy()
    X2 = (3.0 * full[:, 72:] + 1000.0).reshape(T, 5, 7).astype(np.float32)
    X1[:, 1, 2] = np.nan
    c1 = {"lat": np.linspace(40, -40, 8), "lon": np.arange(9) * 10.0}
    c2 = {"lat": np.linspace(30, -30, 5), "lon": np.arange(7) * 20.0}
    kw = dict(n_modes=k, standardize=True, use_coslat=True, random_state=4)
    o = oeof.eof_fit_list([X1, X2], [DIMS, DIMS], "time", coords_list=[c1, c2], **kw)
    m = xb.single.EOF(ops=TorchCpuOps(), **kw).fit([xb.DataArray(X1, DIMS, c1), xb.DataArray(X2, DIMS, c2)], dim="time")
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    np.testing.assert_allclose(m.explained_variance().values, o["explained_variance"], rtol=1e-4)
    comps = m.components()
    assert isinstance(comps, list) and [c.shape for c in comps] == [(8, 9, k), (5, 7, k)]
    dots = 0.0
    for c, oc, f in zip(comps, o["components_2d"], o["fitted"]):
        V = c.values.reshape(-1, k)
        np.testing.assert_array_equal(np.isnan(V).any(axis=1), ~f["is_valid_feature"])
        dots = dots + (V[f["is_valid_feature"]] * oc).sum(axis=0)
    assert (np.abs(dots) > 1 - 1e-5).all(), dots
    sc = m.scores().values
    scale = np.abs(o["scores"]).max(axis=0)
    np.testing.assert_allclose(np.abs(sc) / scale, np.abs(o["scores"]) / scale, atol=1e-3)
    # transform of the training data reproduces the scores; inverse_transform returns one array per input
    np.testing.assert_allclose(m.transform([xb.DataArray(X1, DIMS, c1), xb.DataArray(X2, DIMS, c2)]).values, sc,
                               rtol=1e-3, atol=1e-3 * np.abs(sc).max())
    rec = m.inverse_transform(m.scores())
    assert isinstance(rec, list) and rec[0].shape == X1.shape and rec[1].shape == X2.shape
    with pytest.raises(ValueError, match="same sample"):
        xb.single.EOF(ops=TorchCpuOps(), **kw).fit([xb.DataArray(X1, DIMS, c1), xb.DataArray(X2[:-1], DIMS, c2)], dim="time")


def test_bootstrapper_host_logic():
    """validation/bootstrapper.py:56-135 against its numpy restatement (members seeded for reproducibility)."""
    import xeofs_b200 as xb
    from oracle import bootstrap as oboot
    T, nlat, nlon, k, nb = 120, 10, 12, 4, 3
    X = planted(T, nlat * nlon, 2 * k, seed=9).reshape(T, nlat, nlon)
    X[:, 2, 3] = np.nan
    coords = {"lat": np.linspace(50, -50, nlat), "lon": np.arange(nlon) * 10.0}
    kw = dict(n_modes=k, use_coslat=True, random_state=1)
    m = xb.single.EOF(ops=TorchCpuOps(), **kw).fit(xb.DataArray(X, DIMS, coords), dim="time")
    o = oeof.eof_fit(X, DIMS, "time", coords=coords, **kw)
    b = xb.validation.EOFBootstrapper(n_bootstraps=nb, seed=5, random_state=2).fit(m)
    ob = oboot.eof_bootstrap(o["A"], o["scores"], k, n_bootstraps=nb, seed=5, random_state=2)
    np.testing.assert_allclose(b.explained_variance().values, ob["explained_variance"], rtol=1e-4)
    np.testing.assert_allclose(b.total_variance().values, ob["total_variance"], rtol=1e-5)
    valid = o["fitted"]["is_valid_feature"] if "fitted" in o else ~np.isnan(X[0]).reshape(-1)
    for i, c in enumerate(b.components()):
        V = c.values.reshape(-1, k)
        np.testing.assert_array_equal(np.isnan(V).any(axis=1), ~valid)
        dots = (V[valid] * ob["components"][i]).sum(axis=0)
        assert (dots > 1 - 1e-4).all(), dots
    for i, sc in enumerate(b.scores()):
        scale = np.abs(ob["scores"][i]).max(axis=0)
        np.testing.assert_allclose(sc.values / scale, ob["scores"][i] / scale, atol=1e-3)


# ---------------------------------------------------------------- feature sharding over gloo, world size 2
_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests")); sys.path.insert(0, os.path.join(sys.argv[1], "tests", "golden"))
from _inputs import planted
from cpu_ops import TorchCpuOps
import xeofs_b200 as xb
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
T, nlat, nlon, k = 90, 8, 20, 5
X = planted(T, nlat * nlon, 2 * k, seed=2).reshape(T, nlat, nlon)
X[:, 1, 3] = np.nan; X[:, 6, 7] = np.nan; X[11] = np.nan
lat = np.linspace(70, -70, nlat)
rows = slice(rank * nlat // world, (rank + 1) * nlat // world)
coords = {"lat": lat[rows], "lon": np.arange(nlon) * 18.0}
m = xb.single.EOF(n_modes=k, use_coslat=True, standardize=True, random_state=7, solver_kwargs={"n_iter": 3},
                  distributed=True, ops=TorchCpuOps())
m.fit(xb.DataArray(X[:, rows], ("time", "lat", "lon"), coords), dim="time")
r = xb.single.EOFRotator(n_modes=4).fit(m)
# MCA of two sharded fields: implicit cross-covariance and the default PCA stage
Y = planted(T, nlat * nlon, 2 * k, seed=2)[:, ::-1].copy().reshape(T, nlat, nlon) * 2.0 + 5.0
dax = xb.DataArray(np.nan_to_num(X[:, rows], nan=1.0), ("time", "lat", "lon"), coords)
day = xb.DataArray(Y[:, rows], ("time", "lat", "lon"), coords)
c0 = m.comm.collectives
mi = xb.cross.MCA(n_modes=3, use_pca=False, random_state=7, distributed=True, ops=TorchCpuOps()).fit(dax, day, dim="time")
mp = xb.cross.MCA(n_modes=3, n_pca_modes=8, random_state=7, distributed=True, ops=TorchCpuOps()).fit(dax, day, dim="time")
np.savez(os.path.join(sys.argv[2], f"rank{rank}.npz"), s=m.singular_values().values, comps=m.components().values,
         scores=m.scores().values, evr=m.explained_variance_ratio().values, collectives=c0,
         rot_ev=r.explained_variance().values, rot_comps=r.components().values,
         mca_s=mi.singular_values().values, mca_tsc=mi.total_squared_covariance(), mca_c1=mi.components()[0].values,
         pca_s=mp.singular_values().values, pca_tsc=mp.total_squared_covariance(), pca_c2=mp.components()[1].values)
dist.destroy_process_group()
'''


def test_feature_sharded_fit_over_gloo(tmp_path):
    """One process per shard, feature axis split by latitude rows; the result must equal the single-process fit."""
    import xeofs_b200 as xb
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(script), ROOT, str(tmp_path)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    T, nlat, nlon, k = 90, 8, 20, 5
    X = planted(T, nlat * nlon, 2 * k, seed=2).reshape(T, nlat, nlon)
    X[:, 1, 3] = np.nan
    X[:, 6, 7] = np.nan
    X[11] = np.nan
    coords = {"lat": np.linspace(70, -70, nlat), "lon": np.arange(nlon) * 18.0}
    ref = xb.single.EOF(n_modes=k, use_coslat=True, standardize=True, random_state=7, solver_kwargs={"n_iter": 3},
                        ops=TorchCpuOps()).fit(xb.DataArray(X, DIMS, coords), dim="time")
    rref = xb.single.EOFRotator(n_modes=4).fit(ref)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert int(r0["collectives"]) > 0
    for r in (r0, r1):
        np.testing.assert_allclose(r["s"], ref.singular_values().values, rtol=1e-6)
        np.testing.assert_allclose(r["evr"], ref.explained_variance_ratio().values, rtol=1e-6)
        np.testing.assert_allclose(r["scores"], ref.scores().values, rtol=1e-4, atol=1e-4, equal_nan=True)
        np.testing.assert_allclose(r["rot_ev"], rref.explained_variance().values, rtol=1e-5)
    comps = np.concatenate([r0["comps"], r1["comps"]], axis=0)
    np.testing.assert_allclose(comps, ref.components().values, atol=2e-5, equal_nan=True)
    rot = np.concatenate([r0["rot_comps"], r1["rot_comps"]], axis=0)
    np.testing.assert_allclose(rot, rref.components().values, atol=5e-5, equal_nan=True)
    # MCA, both routes, against the single-process fits
    Y = planted(T, nlat * nlon, 2 * k, seed=2)[:, ::-1].copy().reshape(T, nlat, nlon) * 2.0 + 5.0
    dax = xb.DataArray(np.nan_to_num(X, nan=1.0), DIMS, coords)
    day = xb.DataArray(Y, DIMS, coords)
    mi = xb.cross.MCA(n_modes=3, use_pca=False, random_state=7, ops=TorchCpuOps()).fit(dax, day, dim="time")
    mp = xb.cross.MCA(n_modes=3, n_pca_modes=8, random_state=7, ops=TorchCpuOps()).fit(dax, day, dim="time")
    for r in (r0, r1):
        np.testing.assert_allclose(r["mca_s"], mi.singular_values().values, rtol=1e-5)
        np.testing.assert_allclose(r["mca_tsc"], mi.total_squared_covariance(), rtol=1e-5)
        np.testing.assert_allclose(r["pca_s"], mp.singular_values().values, rtol=1e-5)
        np.testing.assert_allclose(r["pca_tsc"], mp.total_squared_covariance(), rtol=1e-5)
    np.testing.assert_allclose(np.concatenate([r0["mca_c1"], r1["mca_c1"]], axis=0), mi.components()[0].values, atol=5e-5)
    np.testing.assert_allclose(np.concatenate([r0["pca_c2"], r1["pca_c2"]], axis=0), mp.components()[1].values, atol=5e-5)


def test_mca_more_samples_than_features_default_pca():
    """Default arguments (use_pca=True, n_pca_modes=0.999) on fields with n_samples > n_features — station data, small
    grids with long records: the PCA stage runs the range finder in the other orientation
    (preprocessing/pca.py:94-131; the round-1 build raised here)."""
    import xeofs_b200 as xb
    T, k = 400, 2
    rng = np.random.default_rng(8)
    U = np.linalg.qr(rng.standard_normal((T, 4)))[0]
    sig = 100 * 0.6 ** np.arange(4)
    X = ((U * sig) @ np.linalg.qr(rng.standard_normal((40, 4)))[0].T + 0.01 * rng.standard_normal((T, 40))).astype(np.float32)
    Y = ((U * sig) @ np.linalg.qr(rng.standard_normal((30, 4)))[0].T + 0.01 * rng.standard_normal((T, 30))).astype(np.float32)
    o = omca.mca_fit(X, Y, ("time", "x"), ("time", "y"), "time", n_modes=k, random_state=3, use_pca=True,
                     pca_random_state=1)
    m = xb.cross.MCA(n_modes=k, random_state=3, ops=make_ops())
    m.fit(xb.DataArray(X, ("time", "x")), xb.DataArray(Y, ("time", "y")), dim="time")
    assert m.n_pca_modes_ == o["n_pca_modes"]
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    c1, c2 = m.components()
    for c, oc in ((c1, o["components1_2d"]), (c2, o["components2_2d"])):
        assert ((c.values * oc).sum(axis=0) > 1 - 1e-4).all()
    s1, _ = m.scores()
    np.testing.assert_allclose(s1.values, o["scores1"], atol=1e-3 * np.abs(o["scores1"]).max())


@pytest.mark.parametrize("use_pca", [True, False])
def test_mca_variance_based_n_modes(use_pca):
    """A float ``n_modes`` (linalg/decomposer.py:88-94, 188-216): int(0.3 rank) modes of the cross-covariance matrix are
    computed and the first that reach the fraction of ITS total variance (column variances over feature1, ddof = 1)
    are kept."""
    import xeofs_b200 as xb
    T = 200
    rng = np.random.default_rng(2)
    U = np.linalg.qr(rng.standard_normal((T, 8)))[0]
    sig = 100 * 0.6 ** np.arange(8)
    X = ((U * sig) @ np.linalg.qr(rng.standard_normal((60, 8)))[0].T + 0.01 * rng.standard_normal((T, 60))).astype(np.float32)
    Y = ((U * sig) @ np.linalg.qr(rng.standard_normal((50, 8)))[0].T + 0.01 * rng.standard_normal((T, 50))).astype(np.float32)
    kw = dict(use_pca=True, n_pca_modes=20, pca_random_state=1) if use_pca else dict(use_pca=False)
    o = omca.mca_fit(X, Y, ("time", "x"), ("time", "y"), "time", n_modes=0.95, random_state=3, **kw)
    mk = dict(n_pca_modes=20) if use_pca else dict(use_pca=False)
    m = xb.cross.MCA(n_modes=0.95, random_state=3, ops=make_ops(), **mk)
    m.fit(xb.DataArray(X, ("time", "x")), xb.DataArray(Y, ("time", "y")), dim="time")
    assert m.k == len(o["singular_values"]) and 1 <= m.k < 15
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-4)
    c1, _ = m.components()
    assert c1.values.shape[-1] == m.k
    assert ((c1.values * o["components1_2d"]).sum(axis=0) > 1 - 1e-4).all()


def test_inverse_transform_rejects_modes_the_model_does_not_hold():
    """The reference's .sel(mode=...) raises KeyError (single/eof.py:150-152); the device kernel indexes with them."""
    import xeofs_b200 as xb
    X = mock_data_array().astype(np.float32)
    m = xb.single.EOF(n_modes=3, ops=make_ops()).fit(xb.DataArray(X, DIMS, {"lat": MOCK_LAT, "lon": MOCK_LON}), dim="time")
    sc = m.scores()
    for bad in ([0, 1, 2], [1, 2, 4], [1, 2, 17]):
        with pytest.raises(KeyError):
            m.inverse_transform(xb.DataArray(sc.values, ("time", "mode"), {"mode": np.array(bad)}))
    rec = m.inverse_transform(xb.DataArray(sc.values[:, :2], ("time", "mode"), {"mode": np.array([1, 3])}))
    assert rec.values.shape == X.shape


def test_rotator_compute_false_runs_max_iter_without_raising():
    """linalg/_numpy/_rotation.py:172-180: with compute=False there is no stopping test and no RuntimeError."""
    import xeofs_b200 as xb
    X = planted(200, 300, 12, seed=4).reshape(200, 10, 30)
    coords = {"lat": np.linspace(60, -60, 10), "lon": np.arange(30) * 12.0}
    m = xb.single.EOF(n_modes=6, ops=make_ops()).fit(xb.DataArray(X, DIMS, coords), dim="time")
    r = xb.single.EOFRotator(n_modes=6, max_iter=3, compute=False).fit(m)
    assert r.n_iter_ == 3
    with pytest.raises(RuntimeError, match="did not converge"):
        xb.single.EOFRotator(n_modes=6, max_iter=3, compute=True).fit(m)


def test_sketch_draw_is_the_reference_stream_and_memoised():
    """The range finder's sketch is numpy's legacy RandomState stream (what sklearn draws from, the reference's call
    at linalg/decomposer.py:141-146); the fp32 image that goes to the device is the rounded fp64 draw, made once."""
    from xeofs_b200 import _engine as E
    a = E.draw_sketch(5, 300, 20)
    np.testing.assert_array_equal(a, np.random.RandomState(5).normal(size=(300, 20)))
    b = E.draw_sketch(5, 300, 20, f32=True)
    assert b.dtype == np.float32 and b is E.draw_sketch(5, 300, 20, f32=True)
    np.testing.assert_array_equal(b, a.astype(np.float32))
    # the first rows of a longer draw are the shorter draw (row-major fill)
    np.testing.assert_array_equal(E.draw_sketch(5, 100, 20), a[:100])
    rs = np.random.RandomState(7)
    c = E.draw_sketch(rs, 10, 4, f32=True)  # a generator object is consumed, never memoised
    assert c.dtype == np.float32 and not np.array_equal(c, E.draw_sketch(rs, 10, 4, f32=True))


def test_bench_helpers_without_a_gpu():
    """bench.py's clock sampler degrades to 'no samples' on a box without NVML / nvidia-smi, and the ncu tag parser
    names the product kinds the bench line reports."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    bench = importlib.import_module("bench")
    with bench.ClockSampler(0) as c:
        pass
    out = c.summary()
    assert out["samples"] == 0 and out["sm_mhz"] is None and out["reasons"] == []
    nt = importlib.import_module("ncu_traffic")
    assert nt.tag_of("void xb::project_tc_kernel<1, true, 2, false, 0, false, true, 2>(CUtensorMap_st)") == "project_T_h16"
    assert nt.tag_of("void xb::project_tc_kernel<1, false, 1, true, 0, false, false, 1>(CUtensorMap_st)") == "project_S_stats_wcopy"
    assert nt.tag_of("void xb::project_tc_kernel<3, true, 2, false, 0, true, true, 0>(CUtensorMap_st)") == "project_T_x3"
    assert nt.tag_of("void xb::varimax_tc2_kernel<28>(xb::Vt2Params)") == "varimax_sweep"
    # the algorithmic bytes of a product kind: fp16 passes stream half the field, the copy-writing pass 1.5 x
    assert bench.alg_bytes_of("project_T_h16", 100 + 8, 100) == 58
    assert bench.alg_bytes_of("project_S_stats_wcopy", 100 + 8, 100) == 158
    assert bench.alg_bytes_of("project_T_x3", 108, 100) == 108


@pytest.mark.parametrize("dims", [("time", "lat", "lon"), ("time", "lon", "lat")])
def test_coslat_weights_are_broadcast_on_the_device(dims):
    """utils/xarray_utils.py:256-270: sqrt(cos(lat)) along the latitude dim, broadcast over the other feature dims.
    Without user weights only the latitude vector leaves the host; the result equals the host broadcast."""
    import torch
    import xeofs_b200 as xb
    from xeofs_b200 import _labels as L
    from xeofs_b200._preprocessor import Preprocessor
    rng = np.random.default_rng(3)
    lat, lon = np.linspace(80, -80, 7), np.arange(9) * 10.0
    shape = (30,) + tuple({"lat": 7, "lon": 9}[d] for d in dims[1:])
    X = (280 + rng.standard_normal(shape)).astype(np.float32)
    coords = {"lat": lat, "lon": lon}
    p = Preprocessor(make_ops(), with_coslat=True)
    ff = p.fit_transform(xb.DataArray(X, dims, coords), ("time",))
    want = np.array(L.sqrt_cos_lat_weights(dims[1:], shape[1:], coords), dtype=np.float64).reshape(-1)
    assert isinstance(ff.featw, torch.Tensor) and ff.featw.dtype == torch.float64
    np.testing.assert_array_equal(ff.featw.cpu().numpy(), want)
    # user weights on top of them: the host route, same product
    w = xb.DataArray(rng.random(7).astype(np.float64), ("lat",), {"lat": lat})
    p2 = Preprocessor(make_ops(), with_coslat=True)
    ff2 = p2.fit_transform(xb.DataArray(X, dims, coords), ("time",), weights=w)
    wl = np.broadcast_to(np.asarray(w.values).reshape([7 if d == "lat" else 1 for d in dims[1:]]), shape[1:]).reshape(-1)
    np.testing.assert_allclose(ff2.featw.cpu().numpy(), want * wl, rtol=1e-15)
