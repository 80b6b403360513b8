"""Generate golden vectors by EXECUTING the reference's own numpy-level source in this container.

Run once from the repo root (needs /root/reference; the produced .npz files are committed and travel):
    python tests/golden/make_golden.py

xeofs itself cannot be imported here (xarray/dask are not installed), but three pieces of its arithmetic
have no xarray dependency and are executed verbatim from /root/reference:
  * xeofs/linalg/_numpy/_svd.py      class _SVD (numpy twin of linalg.decomposer.Decomposer: solver policy,
                                     sklearn randomized_svd call, sign rule, variance truncation)
  * xeofs/linalg/_numpy/_rotation.py _varimax, _promax
  * xeofs/cross/cpcca.py:1008-1015   CPCCA._compute_cross_covariance_numpy  (function body lifted via ast)
  * xeofs/utils/xarray_utils.py:256-270 _np_sqrt_cos_lat_weights             (function body lifted via ast)
  * xeofs/linalg/_numpy/_utils.py     _fractional_matrix_power (the Whitener's transform, whitener.py:111-133)
  * xeofs/cross/cpcca.py:418-443      _compute_residual_variance_numpy      (nested function, lifted via ast)
  * xeofs/utils/optional/statistics.py:51-55 _correlation_coefficients_numpy (nested function, lifted via ast)
`dask` is stubbed: those files import it only for isinstance checks / the dask branch.
Inputs are the reference's own test fixtures re-created with numpy (tests/conftest.py:225-240 mock_data_array,
seed 7; tests/models/cross/test_cpcca_rotator.py:9-22 generate_random_data) plus a planted-spectrum field.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub_modules():
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")
    dal = types.ModuleType("dask.array.linalg")
    dgm = types.ModuleType("dask.graph_manipulation")

    class Array:  # never instantiated: isinstance(x, Array) is always False for numpy input
        pass

    da.Array = Array
    dal.svd_compressed = None
    dgm.wait_on = lambda *a: a
    dask.array = da
    sys.modules.update({"dask": dask, "dask.array": da, "dask.array.linalg": dal,
                        "dask.graph_manipulation": dgm})
    # package skeleton so that `from ...utils.sanity_checks import sanity_check_n_modes` resolves
    for name in ["xeofs", "xeofs.utils", "xeofs.linalg", "xeofs.linalg._numpy"]:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sc = types.ModuleType("xeofs.utils.sanity_checks")
    sc.sanity_check_n_modes = _lift(f"{REF}/xeofs/utils/sanity_checks.py", "sanity_check_n_modes")
    sys.modules["xeofs.utils.sanity_checks"] = sc


def _lift(path, func_name, ns=None):
    """Compile one function definition out of a reference source file, unmodified."""
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func_name:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            env = {"np": np} | (ns or {})
            exec(compile(mod, path, "exec"), env)
            return env[func_name]
    raise KeyError(func_name)


def _load(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


sys.path.insert(0, HERE)
from _inputs import mock_data_array, planted  # noqa: E402


def main():
    _stub_modules()
    svd_mod = _load("xeofs.linalg._numpy._svd", f"{REF}/xeofs/linalg/_numpy/_svd.py")
    rot_mod = _load("xeofs.linalg._numpy._rotation", f"{REF}/xeofs/linalg/_numpy/_rotation.py")
    xcov = _lift(f"{REF}/xeofs/cross/cpcca.py", "_compute_cross_covariance_numpy")
    coslat = _lift(f"{REF}/xeofs/utils/xarray_utils.py", "_np_sqrt_cos_lat_weights")

    out = {}
    # ---- decomposer: exact-policy case (25 x 20, k=19 > int(0.8*20)) and randomized cases
    X = mock_data_array().reshape(25, 20)
    A = X - X.mean(axis=0)
    U, s, V = svd_mod._SVD(n_modes=19, random_state=5).fit_transform(A.copy())
    out.update(svd_small_A=A, svd_small_U=U, svd_small_s=s, svd_small_V=V)
    U, s, V = svd_mod._SVD(n_modes=3, random_state=5).fit_transform(A.copy())
    out.update(svd_small3_U=U, svd_small3_s=s, svd_small3_V=V)

    Xp = planted(600, 900, 24, seed=11)
    Ap = Xp - Xp.mean(axis=0)                      # float32, like the centred-only Scaler output
    Ap = Ap * np.ones(900, dtype=float)            # weights_ promotion -> float64
    U, s, V = svd_mod._SVD(n_modes=12, random_state=5, solver_kwargs={"n_iter": 4}).fit_transform(Ap.copy())
    out.update(svd_planted_U=U, svd_planted_s=s, svd_planted_V=V)
    U, s, V = svd_mod._SVD(n_modes=12, random_state=5).fit_transform(Ap.T.copy())   # tall (no transpose inside sklearn)
    out.update(svd_plantedT_U=U, svd_plantedT_s=s, svd_plantedT_V=V)
    U, s, V = svd_mod._SVD(n_modes=0.9, init_rank_reduction=0.05, random_state=5).fit_transform(Ap.copy())
    out.update(svd_var_U=U, svd_var_s=s, svd_var_V=V)

    # ---- rotation
    rng = np.random.default_rng(3)
    L = rng.standard_normal((200, 6)) * np.array([5, 4, 3, 2, 1.5, 1.0])
    Xr, R = rot_mod._varimax(L.copy(), max_iter=1000, rtol=1e-8)
    out.update(rot_L=L, varimax_X=Xr, varimax_R=R)
    for p in (1, 2, 4):
        Xr, R, phi = rot_mod._promax(L.copy(), power=p, max_iter=1000, rtol=1e-8)
        out.update({f"promax{p}_X": Xr, f"promax{p}_R": R, f"promax{p}_phi": phi})
    # sparse-pattern loadings (varimax has a known simple-structure target)
    S, m = 3000, 8
    pat = np.zeros((S, m))
    for j in range(m):
        pat[j * 300:(j + 1) * 300 + 200, j] = rng.standard_normal(min(500, S - j * 300))[: pat[j * 300:(j + 1) * 300 + 200, j].size]
    Qm, _ = np.linalg.qr(rng.standard_normal((m, m)))
    L2 = pat @ Qm
    Xr, R = rot_mod._varimax(L2.copy(), max_iter=1000, rtol=1e-8)
    out.update(rot_L2=L2, varimax2_X=Xr, varimax2_R=R)

    # ---- cross covariance  (tests/models/cross/test_cpcca_rotator.py:9-22 inputs)
    r1 = np.random.default_rng(123)
    r2 = np.random.default_rng(321)
    X1 = r1.standard_normal((200, 10)); X1 = X1 - X1.mean(axis=0)
    X2 = r2.standard_normal((200, 20)); X2 = X2 - X2.mean(axis=0)
    out.update(xcov_X=X1, xcov_Y=X2, xcov_C=xcov(X1, X2))

    # ---- coslat
    lat = np.array([-95.0, -90.0, -60.0, -33.3, 0.0, 20.0, 45.0, 89.9, 90.0, 91.0])
    out.update(coslat_lat=lat, coslat_w=coslat(lat))

    # ---- fractional whitening (preprocessing/whitener.py:111-133 calls _fractional_matrix_power with solver="full")
    utl_mod = _load("xeofs.linalg._numpy._utils", f"{REF}/xeofs/linalg/_numpy/_utils.py")
    rw = np.random.default_rng(17)
    Xw = rw.standard_normal((120, 7)) * np.array([9.0, 6.0, 4.0, 2.5, 1.5, 1.0, 0.6])
    Xw = Xw @ np.linalg.qr(rw.standard_normal((7, 7)))[0]
    Xw = Xw - Xw.mean(axis=0)
    Cw = Xw.conj().T @ Xw / Xw.shape[0]                      # whitener.py:124-125
    out.update(whit_X=Xw)
    for alpha in (0.0, 0.2, 0.7):
        out[f"whit_T_alpha{alpha}"] = utl_mod._fractional_matrix_power(Cw, (alpha - 1) / 2, random_state=3,
                                                                        solver="full")

    # ---- residual squared covariance of one mode (cpcca.py:418-443) and the Pearson correlation of
    #      utils/optional/statistics.py:51-55 (nested functions, lifted unmodified)
    resid = _lift(f"{REF}/xeofs/cross/cpcca.py", "_compute_residual_variance_numpy")
    pear = _lift(f"{REF}/xeofs/utils/optional/statistics.py", "_correlation_coefficients_numpy")
    Yw = rw.standard_normal((120, 5))
    Yw = Yw - Yw.mean(axis=0)
    Xrec = np.outer(Xw @ rw.standard_normal(7), rw.standard_normal(7))
    Yrec = np.outer(Yw @ rw.standard_normal(5), rw.standard_normal(5))
    out.update(resid_Y=Yw, resid_Xrec=Xrec, resid_Yrec=Yrec, resid_value=np.array(resid(Xw, Yw, Xrec, Yrec)),
               pearson_XY=pear(Xw, Yw))

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
