#!/usr/bin/env python
"""Full-size parity against the oracle, run once per round on the GPU box's host (196 GB RAM, 16 cores):

  c2   EOF n_modes=50 n_iter=4 on the whole 8760 x (721 x 1440) bench field: oracle.eof_fit arithmetic in fp64 on the
       host (Scaler / Sanitizer per latitude-row chunk through oracle.preprocess — the arithmetic is per feature, so
       chunking changes nothing — then oracle.decomposer.decompose = sklearn randomized_svd on the assembled 73 GB
       matrix) against xeofs_b200.single.EOF on the same field: singular values, explained variance ratio,
       components, scores, with the north star's tolerances.
  c3   MCA n_modes=20 (use_pca=False) on two 8760 x (rows x 720) fields, rows chosen so that the explicit
       cross-covariance matrix the oracle forms (S x S fp64) fits in RAM.
  c5   EOFRotator varimax on the 100 modes of the config-5 EOF model (loadings 4 147 200 x 100): oracle.rotation
       (fp64 numpy restatement of _varimax / _promax and the rotator's post-processing) on the device model's own
       components / scores against xeofs_b200.single.EOFRotator.

Writes gpurun_out/r02_parity_full.json (copied to profiles/) and, from the c2 / c4 runs, the expected singular values
bench.py asserts against (profiles/r02_expected_sv.json).  usage: python tools/parity_full.py [c2] [c3] [c5] [c4n1]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import xeofs_b200 as xb  # noqa: E402
from oracle import mca as omca  # noqa: E402
from oracle import preprocess as opp  # noqa: E402
from oracle import rotation as orot  # noqa: E402
from oracle.decomposer import decompose  # noqa: E402

DIMS = bench.DIMS
OUT = os.path.join(ROOT, "gpurun_out", "r02_parity_full.json")
EXP = os.path.join(ROOT, "gpurun_out", "r02_expected_sv.json")


def mem_available_gb():
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            return int(line.split()[1]) / 1e6
    return 0.0


def compare_modes(s, s_ref, V, V_ref, Sc, Sc_ref):
    """Singular values rtol, |<v_ref, v>| and |<score_ref, score>| / (|.||.|) per mode (up to sign), and how many
    modes carry the same sign (the sign rule is applied on both sides)."""
    dots = np.abs((V * V_ref).sum(0)) / (np.linalg.norm(V, axis=0) * np.linalg.norm(V_ref, axis=0))
    sdots = (Sc * Sc_ref).sum(0) / (np.linalg.norm(Sc, axis=0) * np.linalg.norm(Sc_ref, axis=0))
    return {
        "singular_values_max_rel_err": float(np.max(np.abs(s / s_ref - 1))),
        "components_min_abs_dot": float(dots.min()),
        "scores_min_abs_dot": float(np.abs(sdots).min()),
        "modes_with_equal_sign": int((sdots > 0).sum()), "modes": int(len(s)),
        "scores_max_rel_l2_err_sign_aligned": float(np.max(
            np.linalg.norm(Sc * np.sign(sdots) - Sc_ref, axis=0) / np.linalg.norm(Sc_ref, axis=0))),
    }


def run_c2(report, expected, wl="c2"):
    T, n_lat, n_lon, k, n_iter, kw = bench.WORKLOADS[wl]
    dev = torch.device("cuda")
    fp = bench.FIELD.get(wl, dict(seed=1, decay=0.9))
    X = bench.planted_field_device(T, n_lat, n_lon, 0, n_lat, 2 * k, fp["seed"], dev, decay=fp["decay"])
    coords = {"lat": np.linspace(90.0, -90.0, n_lat), "lon": np.arange(n_lon) * (360.0 / n_lon)}
    m = xb.single.EOF(n_modes=k, random_state=bench.RANDOM_STATE, solver_kwargs={"n_iter": n_iter}, **kw)
    t0 = time.perf_counter()
    m.fit(xb.DataArray(X, DIMS, coords), dim="time")
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    s = m.singular_values().values
    evr = m.explained_variance_ratio().values
    V = m.components().values.reshape(-1, k)
    Sc = m.scores().values
    tv = float(m.total_variance())
    S = n_lat * n_lon
    need = T * S * 12 / 1e9 + 12
    if mem_available_gb() < need:
        raise RuntimeError(f"host RAM: {mem_available_gb():.0f} GB available, {need:.0f} GB needed")
    t0 = time.perf_counter()
    Xh = X.cpu().numpy()
    del X, m
    torch.cuda.empty_cache()
    t_d2h = time.perf_counter() - t0
    # ---- the oracle's fit, fp64, on the host
    t0 = time.perf_counter()
    A = np.empty((T, S), dtype=np.float64)
    tv_ref = 0.0
    rows = 16
    for r0 in range(0, n_lat, rows):
        r1 = min(n_lat, r0 + rows)
        f = opp.preprocess(Xh[:, r0:r1], DIMS, "time", coords={"lat": coords["lat"][r0:r1], "lon": coords["lon"]},
                           center=True, standardize=kw.get("standardize", False),
                           use_coslat=kw.get("use_coslat", False))
        assert f["is_valid_feature"].all() and f["is_valid_sample"].all()
        A[:, r0 * n_lon:r1 * n_lon] = f["A"]
        tv_ref += float(f["A"].var(axis=0, ddof=1).sum())      # utils/xarray_utils.py:236-253
    del Xh
    t_pre = time.perf_counter() - t0
    t0 = time.perf_counter()
    U, s_ref, V_ref = decompose(A, n_modes=k, random_state=bench.RANDOM_STATE, solver_kwargs={"n_iter": n_iter})
    t_svd = time.perf_counter() - t0
    del A
    evr_ref = s_ref ** 2 / (T - 1) / tv_ref
    cmp_ = compare_modes(s, s_ref, V, V_ref, Sc, U * s_ref)
    cmp_["explained_variance_ratio_max_rel_err"] = float(np.max(np.abs(evr / evr_ref - 1)))
    cmp_["total_variance_rel_err"] = abs(tv / tv_ref - 1)
    ok = (cmp_["singular_values_max_rel_err"] <= 1e-4 and cmp_["explained_variance_ratio_max_rel_err"] <= 1e-4
          and cmp_["components_min_abs_dot"] >= 1 - 1e-4 and cmp_["scores_min_abs_dot"] >= 1 - 1e-4
          and cmp_["modes_with_equal_sign"] == k)
    report[wl] = {"workload": f"EOF n_modes={k} n_iter={n_iter} on {T}x({n_lat}x{n_lon}) fp32 "
                              f"({T * S * 4 / 1e9:.2f} GB), kwargs {kw}",
                  "tolerances": "singular values / explained variance ratio rtol 1e-4; |<v_ref, v>| >= 1 - 1e-4 "
                                "per mode for components and scores; equal signs",
                  "ok": bool(ok), **cmp_,
                  "oracle": {"cores": os.cpu_count(), "preprocess_s": t_pre, "randomized_svd_s": t_svd,
                             "fit_s": t_pre + t_svd, "GBps": T * S * 4 / 1e9 / (t_pre + t_svd), "d2h_copy_s": t_d2h},
                  "gpu_first_fit_s_cold": t_gpu,
                  "singular_values_head": [float(v) for v in s[:5]],
                  "oracle_singular_values_head": [float(v) for v in s_ref[:5]]}
    expected[wl] = {"source": "oracle.eof_fit arithmetic (host fp64, sklearn randomized_svd) on the full field, "
                              "tools/parity_full.py", "s": [float(v) for v in s_ref]}
    print(json.dumps(report[wl]), flush=True)


def run_c4n1(report, expected):
    """The one-GPU fit of the strong-scaling field: its singular values are what every N is compared with."""
    T, n_lat, n_lon, k, n_iter, kw = bench.WORKLOADS["c4"]
    dev = torch.device("cuda")
    fp = bench.FIELD["c4"]
    X = bench.planted_field_device(T, n_lat, n_lon, 0, n_lat, 2 * k, fp["seed"], dev, decay=fp["decay"])
    coords = {"lat": np.linspace(90.0, -90.0, n_lat), "lon": np.arange(n_lon) * (360.0 / n_lon)}

    def fit():
        m = xb.single.EOF(n_modes=k, random_state=bench.RANDOM_STATE, solver_kwargs={"n_iter": n_iter}, **kw)
        return m.fit(xb.DataArray(X, DIMS, coords), dim="time")

    for _ in range(2):
        m = fit()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m = fit()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    s = m.singular_values().values
    expected["c4_strong"] = {"source": "xeofs_b200 one-GPU fit of the same globally seeded field "
                                       "(tools/parity_full.py c4n1); the oracle cannot hold 145 GB x 3 on the host",
                             "s": [float(v) for v in s], "ms_per_step_n1": ms}
    report["c4n1"] = {"ms_per_step": ms, "singular_values_head": [float(v) for v in s[:5]]}
    print(json.dumps(report["c4n1"]), flush=True)


def run_c3(report, expected):
    T, n_lon, k = bench.C3["T"], bench.C3["n_lon"], bench.C3["k"]
    rows = int(os.environ.get("XEOFS_PARITY_C3_ROWS", 60))  # S = 43 200: C is 14.9 GB in fp64
    dev = torch.device("cuda")
    U = bench.temporal_factors(T, 2 * k, 2, dev)
    mk = lambda seed: bench.planted_field_device(T, rows, n_lon, 0, rows, 2 * k, seed, dev, sigma0=1e5,  # noqa: E731
                                                 decay=0.85, eps=0.05, U=U)
    X, Y = mk(2), mk(3)
    coords = {"lat": np.linspace(60, -60, rows), "lon": np.arange(n_lon) * 0.5}
    m = xb.cross.MCA(n_modes=k, random_state=bench.RANDOM_STATE, use_pca=False)
    m.fit(xb.DataArray(X, DIMS, coords), xb.DataArray(Y, DIMS, coords), dim="time")
    s = m.singular_values().values
    c1, c2 = m.components()
    s1, s2 = m.scores()
    tsc = m.total_squared_covariance()
    Xh, Yh = X.cpu().numpy(), Y.cpu().numpy()
    del X, Y
    t0 = time.perf_counter()
    # oracle.mca.mca_fit line by line, without the diagnostics it also evaluates (20 more S x S products)
    A1 = opp.preprocess(Xh, DIMS, "time", coords=coords, center=True)["A"]
    A2 = opp.preprocess(Yh, DIMS, "time", coords=coords, center=True)["A"]
    del Xh, Yh
    C = omca.cross_covariance(A1, A2)                                         # cpcca.py:1008-1015
    Q1, s_o, Q2 = decompose(C, n_modes=k, random_state=bench.RANDOM_STATE)    # cpcca.py:187-194
    o = {"singular_values": s_o, "components1_2d": Q1, "components2_2d": Q2, "scores1": A1 @ Q1, "scores2": A2 @ Q2,
         "total_squared_covariance": float((np.abs(C) ** 2).sum())}          # cpcca.py:991-1000, 204-205
    del C, A1, A2
    t_o = time.perf_counter() - t0
    S = rows * n_lon
    c_1 = compare_modes(s, o["singular_values"], c1.values.reshape(-1, k), o["components1_2d"], s1.values, o["scores1"])
    c_2 = compare_modes(s, o["singular_values"], c2.values.reshape(-1, k), o["components2_2d"], s2.values, o["scores2"])
    tsc_err = abs(tsc / o["total_squared_covariance"] - 1)
    ok = all(c["singular_values_max_rel_err"] <= 1e-4 and c["components_min_abs_dot"] >= 1 - 1e-4 and
             c["scores_min_abs_dot"] >= 1 - 1e-4 and c["modes_with_equal_sign"] == k for c in (c_1, c_2)) and tsc_err <= 1e-4
    report["c3"] = {"workload": f"MCA n_modes={k} use_pca=False on two {T}x({rows}x{n_lon}) fp32 fields (S = {S}: the "
                                f"explicit {S} x {S} cross-covariance of the oracle is {S * S * 8 / 1e9:.1f} GB)",
                    "ok": bool(ok), "field1": c_1, "field2": c_2, "total_squared_covariance_rel_err": tsc_err,
                    "oracle_fit_s": t_o, "singular_values_head": [float(v) for v in s[:5]]}
    print(json.dumps(report["c3"]), flush=True)


def run_c5(report, expected):
    T, n_lat, n_lon, k = bench.C5["T"], bench.C5["n_lat"], bench.C5["n_lon"], bench.C5["k"]
    dev = torch.device("cuda")
    X = bench.planted_field_device(T, n_lat, n_lon, 0, n_lat, 2 * k, 4, dev, decay=0.995, sparse=0.05)
    coords = {"lat": np.linspace(89.9, -89.9, n_lat), "lon": np.arange(n_lon) * 0.125}
    model = xb.single.EOF(n_modes=k, use_coslat=True, random_state=bench.RANDOM_STATE, solver_kwargs={"n_iter": 4})
    model.fit(xb.DataArray(X, DIMS, coords), dim="time")
    del X
    torch.cuda.empty_cache()
    r = xb.single.EOFRotator(n_modes=k, power=1, max_iter=1000)
    r.fit(model)
    V = model.components().values.reshape(-1, k).astype(np.float64)
    valid = ~np.isnan(V[:, 0])
    ev = model.explained_variance().values.astype(np.float64)
    sc = model.scores().values.astype(np.float64)
    sv = model.singular_values().values.astype(np.float64)
    t0 = time.perf_counter()
    o = orot.eof_rotator_fit(V[valid], ev, sc, sv, T, n_modes=k, power=1, max_iter=1000)
    t_o = time.perf_counter() - t0
    Vr = r.components().values.reshape(-1, k)[valid]
    cm = compare_modes(r.explained_variance().values, o["explained_variance"], Vr, o["components_2d"],
                       r.scores().values, o["scores"])
    cm["explained_variance_max_rel_err"] = cm.pop("singular_values_max_rel_err")
    ok = (cm["explained_variance_max_rel_err"] <= 1e-4 and cm["components_min_abs_dot"] >= 1 - 1e-4 and
          cm["scores_min_abs_dot"] >= 1 - 1e-4 and cm["modes_with_equal_sign"] == k)
    report["c5"] = {"workload": f"EOFRotator varimax power=1 max_iter=1000 on {k} modes, loadings {int(valid.sum())} x {k}",
                    "ok": bool(ok), **cm, "iterations_device": int(r.n_iter_),
                    "iterations_tensor_core": int(getattr(r, "n_iter_tc_", 0)), "oracle_fit_s": t_o}
    print(json.dumps(report["c5"]), flush=True)


def main():
    what = sys.argv[1:] or ["c2", "c3", "c5"]
    report, expected = {}, {}
    for p, store in ((OUT, report), (EXP, expected)):
        if os.path.exists(p):
            store.update(json.load(open(p)))
    # expectations committed earlier survive a partial run
    committed = os.path.join(ROOT, "profiles", "r02_expected_sv.json")
    if not expected and os.path.exists(committed):
        expected.update(json.load(open(committed)))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    for w in what:
        try:
            {"c2": run_c2, "c3": run_c3, "c5": run_c5, "c4n1": run_c4n1,
             "mid": lambda r, e: run_c2(r, e, "mid")}[w](report, expected)
        except Exception as exc:  # keep what the other cases measured
            report[w] = {"ok": False, "error": f"{type(exc).__name__}: {exc}"[:500]}
            print(json.dumps(report[w]), flush=True)
        torch.cuda.empty_cache()
        json.dump(report, open(OUT, "w"), indent=1)
        json.dump(expected, open(EXP, "w"), indent=1)


if __name__ == "__main__":
    main()
