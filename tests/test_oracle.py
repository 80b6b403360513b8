"""CPU: pin the oracle against (a) golden vectors produced by executing the reference's own numpy source
(tests/golden/make_golden.py) and (b) the invariants the reference's test-suite asserts for this path."""
import os

import numpy as np
import pytest

from _inputs import MOCK_LAT, MOCK_LON, mock_data_array, planted
from oracle import eof as oeof
from oracle import mca as omca
from oracle import preprocess as opp
from oracle import rotation as orot
from oracle.decomposer import decompose

DIMS = ("time", "lat", "lon")
COORDS = {"lat": MOCK_LAT, "lon": MOCK_LON}


# ---------------------------------------------------------------- golden vectors (reference source executed)
def test_decomposer_exact_policy_matches_reference(golden):
    A = golden["svd_small_A"]
    U, s, V = decompose(A.copy(), n_modes=19, random_state=5)
    np.testing.assert_allclose(s, golden["svd_small_s"], rtol=1e-12)
    np.testing.assert_allclose(U, golden["svd_small_U"], atol=1e-10)
    np.testing.assert_allclose(V, golden["svd_small_V"], atol=1e-10)


def test_decomposer_randomized_small_matches_reference(golden):
    A = golden["svd_small_A"]
    U, s, V = decompose(A.copy(), n_modes=3, random_state=5)
    np.testing.assert_allclose(s, golden["svd_small3_s"], rtol=1e-12)
    np.testing.assert_allclose(V, golden["svd_small3_V"], atol=1e-10)
    np.testing.assert_allclose(U, golden["svd_small3_U"], atol=1e-10)


def _planted_A():
    Xp = planted(600, 900, 24, seed=11)
    Ap = Xp - Xp.mean(axis=0)
    return Ap * np.ones(900, dtype=float)


def test_decomposer_randomized_planted_matches_reference(golden):
    Ap = _planted_A()
    U, s, V = decompose(Ap.copy(), n_modes=12, random_state=5, solver_kwargs={"n_iter": 4})
    np.testing.assert_allclose(s, golden["svd_planted_s"], rtol=1e-12)
    np.testing.assert_allclose(V, golden["svd_planted_V"], atol=1e-9)
    np.testing.assert_allclose(U, golden["svd_planted_U"], atol=1e-9)
    U, s, V = decompose(Ap.T.copy(), n_modes=12, random_state=5)
    np.testing.assert_allclose(s, golden["svd_plantedT_s"], rtol=1e-12)
    np.testing.assert_allclose(V, golden["svd_plantedT_V"], atol=1e-9)


def test_decomposer_variance_truncation_matches_reference(golden):
    Ap = _planted_A()
    U, s, V = decompose(Ap.copy(), n_modes=0.9, init_rank_reduction=0.05, random_state=5)
    assert s.shape == golden["svd_var_s"].shape
    np.testing.assert_allclose(s, golden["svd_var_s"], rtol=1e-12)
    np.testing.assert_allclose(V, golden["svd_var_V"], atol=1e-9)


def test_varimax_promax_match_reference(golden):
    L = golden["rot_L"]
    Xr, R = orot.varimax(L.copy())
    np.testing.assert_allclose(Xr, golden["varimax_X"], atol=1e-10)
    np.testing.assert_allclose(R, golden["varimax_R"], atol=1e-10)
    for p in (1, 2, 4):
        Xr, R, phi = orot.promax(L.copy(), power=p)
        np.testing.assert_allclose(Xr, golden[f"promax{p}_X"], atol=1e-9)
        np.testing.assert_allclose(R, golden[f"promax{p}_R"], atol=1e-9)
        np.testing.assert_allclose(phi, golden[f"promax{p}_phi"], atol=1e-9)
    Xr, R = orot.varimax(golden["rot_L2"].copy())
    np.testing.assert_allclose(Xr, golden["varimax2_X"], atol=1e-9)


def test_cross_covariance_and_coslat_match_reference(golden):
    np.testing.assert_allclose(omca.cross_covariance(golden["xcov_X"], golden["xcov_Y"]), golden["xcov_C"], rtol=1e-13)
    np.testing.assert_allclose(opp.sqrt_cos_lat(golden["coslat_lat"]), golden["coslat_w"], rtol=0, atol=0)


# ---------------------------------------------------------------- reference test invariants
@pytest.mark.parametrize("standardize,use_coslat", [(False, False), (True, False), (True, True)])
def test_scaler_mean_std(standardize, use_coslat):
    """tests/preprocessing/test_scaler_dataarray.py:76-108, test_preprocessor_dataarray.py:61-89."""
    X = mock_data_array()
    f = opp.preprocess(X, DIMS, "time", coords=COORDS, center=True, standardize=standardize, use_coslat=use_coslat)
    A = f["A"]
    assert A.shape == (25, 20)
    np.testing.assert_allclose(A.mean(axis=0), 0, atol=1e-12)
    if standardize and not use_coslat:
        np.testing.assert_allclose(A.std(axis=0), 1, rtol=1e-12)
    if standardize and use_coslat:
        w = np.repeat(opp.sqrt_cos_lat(MOCK_LAT), 4)
        np.testing.assert_allclose(A.std(axis=0), w, rtol=1e-12)
    f = opp.preprocess(X, DIMS, "time", coords=COORDS, standardize=True, weights=np.full((5, 4), 0.5))
    np.testing.assert_allclose(f["A"].std(axis=0), 0.5, rtol=1e-12)


def test_fp32_input_is_promoted_to_fp64():
    """scaler.py:153 with weights_ = float64 ones (utils/xarray_utils.py:83-87)."""
    X = mock_data_array().astype(np.float32)
    f = opp.preprocess(X, DIMS, "time", coords=COORDS)
    assert f["A"].dtype == np.float64
    assert f["scaler"]["mean"].dtype == np.float32


def test_sanitizer_nan_policies():
    """tests/preprocessing/test_sanitizer.py:233-275, tests/models/single/test_eof.py:111-181."""
    X = mock_data_array()
    Xf = X.copy()
    Xf[:, 1, 2] = np.nan          # full-dimensional NaN feature
    Xf[:, 4, 0] = np.nan
    f = opp.preprocess(Xf, DIMS, "time", coords=COORDS)
    assert f["A"].shape == (25, 18)
    assert f["is_valid_feature"].sum() == 18 and not f["is_valid_feature"][1 * 4 + 2]
    Xs = Xf.copy()
    Xs[3] = np.nan                # all-NaN sample
    f = opp.preprocess(Xs, DIMS, "time", coords=COORDS)
    assert f["A"].shape == (24, 18)
    Xi = X.copy()
    Xi[0, 0, 0] = np.nan          # isolated NaN
    with pytest.raises(ValueError, match="partial NaN"):
        opp.preprocess(Xi, DIMS, "time", coords=COORDS)
    r = oeof.eof_fit(Xf, DIMS, "time", coords=COORDS, n_modes=3, random_state=1)
    assert np.isnan(r["components"][1, 2]).all() and np.isfinite(r["components"][0, 0]).all()


def test_total_variance_and_ratio():
    """tests/utils/test_total_variance.py:7-22; tests/models/single/test_eof.py:85-100."""
    X = mock_data_array()
    r = oeof.eof_fit(X, DIMS, "time", coords=COORDS, n_modes=5, random_state=2)
    np.testing.assert_allclose(r["total_variance"], np.var(X.reshape(25, 20), axis=0, ddof=1).sum(), rtol=1e-12)
    assert r["explained_variance_ratio"].sum() <= 1 + 1e-12
    assert (np.diff(r["explained_variance"]) <= 0).all()


def test_transform_equals_scores_and_full_rank_reconstruction():
    """tests/models/single/test_eof.py:364-408 (rtol 1e-3) and :455-488."""
    X = mock_data_array()
    r = oeof.eof_fit(X, DIMS, "time", coords=COORDS, n_modes=20, standardize=True, use_coslat=True, solver="full")
    np.testing.assert_allclose(oeof.eof_transform(r, X, DIMS), r["scores"], rtol=1e-3, atol=1e-8)
    rec = oeof.eof_inverse_transform(r, r["scores"])
    np.testing.assert_allclose(rec, X, rtol=1e-8)
    sgn_pos = np.abs(r["components_2d"].max(axis=0)) >= np.abs(r["components_2d"].min(axis=0))
    assert sgn_pos.all()
    np.testing.assert_allclose(np.linalg.norm(r["components_2d"], axis=0), 1, rtol=1e-12)


def test_seed_determinism():
    """tests/linalg/test_decomposer.py:168-192."""
    A = _planted_A()
    U1, s1, V1 = decompose(A, n_modes=5, random_state=42, solver="randomized")
    U2, s2, V2 = decompose(A, n_modes=5, random_state=42, solver="randomized")
    assert np.array_equal(U1, U2) and np.array_equal(V1, V2)
    with pytest.raises(ValueError, match="rank"):
        decompose(A[:10], n_modes=11)


def test_rotation_conserves_variance():
    """tests/models/single/test_eof_rotator.py:98-137."""
    X = mock_data_array()
    r = oeof.eof_fit(X, DIMS, "time", coords=COORDS, n_modes=10, solver="full")
    for power in (1, 2):
        rot = orot.eof_rotator_fit(r["components_2d"], r["explained_variance"], r["scores"], r["norms"],
                                   n_samples=25, n_modes=4, power=power)
        if power == 1:
            np.testing.assert_allclose(rot["explained_variance"].sum(), r["explained_variance"][:4].sum(), rtol=1e-10)
            np.testing.assert_allclose(rot["phi_matrix"], np.eye(4), atol=1e-8)
        assert (np.diff(rot["explained_variance"]) <= 0).all()
        np.testing.assert_allclose(np.linalg.norm(rot["components_2d"], axis=0), 1, rtol=1e-10)


def test_mca_total_squared_covariance():
    """tests/models/cross/test_cpcca.py:152-164: total squared covariance == sum of squared np.cov cross block."""
    r1, r2 = np.random.default_rng(123), np.random.default_rng(321)
    X = r1.standard_normal((200, 10))
    Y = r2.standard_normal((200, 20))
    m = omca.mca_fit(X, Y, ("sample", "feature"), ("sample", "feature"), "sample", n_modes=5, random_state=0)
    cov = np.cov(X.T, Y.T)[:10, 10:]
    np.testing.assert_allclose(m["total_squared_covariance"], (cov**2).sum(), rtol=1e-10)
    s_exact = np.linalg.svd(cov, compute_uv=False)[:5]
    np.testing.assert_allclose(m["singular_values"], s_exact, rtol=1e-8)
    np.testing.assert_allclose(m["norm1"], np.linalg.norm(m["scores1"], axis=0))


def test_whitener_residual_and_pearson_against_reference_vectors():
    """Fractional whitening (preprocessing/whitener.py:111-133 -> linalg/_numpy/_utils.py:6-33), the residual squared
    covariance of cpcca.py:436-443 and the Pearson correlation of utils/optional/statistics.py:51-55 were executed
    from the reference's source (tests/golden/make_golden.py); the oracle's restatements must reproduce them."""
    from oracle import mca as omca
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
    X = g["whit_X"]
    for alpha in (0.0, 0.2, 0.7):
        T, Tinv = omca.whitener_transform(X, alpha)
        np.testing.assert_allclose(T, g[f"whit_T_alpha{alpha}"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(T @ Tinv, np.eye(X.shape[1]), atol=1e-9)
    np.testing.assert_allclose(omca.residual_squared_covariance(X, g["resid_Y"], g["resid_Xrec"], g["resid_Yrec"]),
                               float(g["resid_value"]), rtol=1e-12)
    np.testing.assert_allclose(omca.pearson_correlation(X, g["resid_Y"]), g["pearson_XY"], rtol=1e-12, atol=1e-15)
