#!/usr/bin/env python
"""BASELINE.json configs[2] (MCA, implicit cross-covariance) and configs[4] (varimax on 100 modes) timed on one GPU.
usage: python tools/bench_models.py [mca|rot|all] [scale]   (scale < 1 shrinks the feature axes)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import xeofs_b200 as xb  # noqa: E402

DIMS = ("time", "lat", "lon")


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def coupled(T, S, r, seed, U, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    V = torch.randn((S, r), generator=g, device=dev)
    V /= V.norm(dim=0, keepdim=True)
    sig = 1e5 * 0.85 ** torch.arange(r, device=dev, dtype=torch.float32)
    X = torch.empty((T, S), device=dev)
    rows = max(1, (1 << 30) // (4 * S))
    Vt = (V * sig).t().contiguous()
    for t0 in range(0, T, rows):
        blk = X[t0:t0 + rows]
        blk.normal_(0.0, 0.05, generator=g)
        blk.addmm_(U[t0:t0 + rows], Vt)
        blk.add_(280.0)
    return X


def mca(scale):
    dev = torch.device("cuda")
    T, nlat, nlon, k = 8760, int(360 * scale), 720, 20
    g = torch.Generator(device=dev).manual_seed(2)
    U = torch.linalg.qr(torch.randn((T, 2 * k), generator=g, device=dev))[0]
    X = coupled(T, nlat * nlon, 2 * k, 2, U, dev).reshape(T, nlat, nlon)
    Y = coupled(T, nlat * nlon, 2 * k, 3, U, dev).reshape(T, nlat, nlon)
    coords = {"lat": np.linspace(89.75, -89.75, nlat), "lon": np.arange(nlon) * 0.5}

    def fit(tsc):
        m = xb.cross.MCA(n_modes=k, random_state=5, use_pca=False, total_squared_covariance=tsc)
        return m.fit(xb.DataArray(X, DIMS, coords), xb.DataArray(Y, DIMS, coords), dim="time")

    ms, m = timed(lambda: fit(False))
    gb = 2 * T * nlat * nlon * 4 / 1e9
    ms_tsc, m2 = timed(lambda: fit(True), n=1)
    s = m.singular_values().values
    if os.environ.get("XEOFS_BENCH_MCA_PCA", "1") != "0":
        def fit_pca():
            mp = xb.cross.MCA(n_modes=k, random_state=5)   # reference defaults: use_pca=True, 99.9 % of variance
            mp.fit(xb.DataArray(X, DIMS, coords), xb.DataArray(Y, DIMS, coords), dim="time")
            return mp
        ms_pca, mp = timed(fit_pca, n=1)
        print(json.dumps({"config": f"MCA n_modes={k} with the default PCA stage (use_pca=True, n_pca_modes=0.999) on the "
                                    f"same fields", "fit_ms": ms_pca, "pca_modes_kept": mp.n_pca_modes_,
                          "singular_values_head": [float(v) for v in mp.singular_values().values[:3]]}), flush=True)
        del mp
    print(json.dumps({"config": f"MCA n_modes={k} (n_iter auto=7) on two {T}x({nlat}x{nlon}) fp32 fields, implicit C",
                      "input_GB": gb, "fit_ms": ms, "GBps_of_input": gb / ms * 1e3, "launches": m.ops.launches,
                      "fit_with_total_squared_covariance_ms": ms_tsc, "tsc": m2.total_squared_covariance(),
                      "singular_values_head": [float(v) for v in s[:3]],
                      "streams_of_both_fields": 2 * 7 + 2 + 1 + 1}), flush=True)


def rot(scale):
    dev = torch.device("cuda")
    T, nlat, nlon, k = 1024, int(1440 * scale), 2880, 100
    # an EOF model with 100 modes on a field of config-4 width (the loadings are S x 100 = 1.66 GB at scale 1)
    # 100 sparse-ish patterns of nearly equal variance (plus a tail): the EOFs come out as mixtures of them and varimax
    # has a simple structure to find
    X = bench.planted_field_device(T, nlat * nlon, 2 * k, 4, dev, decay=0.995, sparse=0.05).reshape(T, nlat, nlon)
    coords = {"lat": np.linspace(89.9, -89.9, nlat), "lon": np.arange(nlon) * 0.125}
    model = xb.single.EOF(n_modes=k, use_coslat=True, random_state=5, solver_kwargs={"n_iter": 4})
    model.fit(xb.DataArray(X, DIMS, coords), dim="time")
    del X
    torch.cuda.empty_cache()
    r = None

    max_iter = int(os.environ.get("XEOFS_BENCH_ROT_MAX_ITER", 1000))  # BASELINE configs[4]: max_iter=1000
    conv = True

    def fit():
        # the number of iterations to convergence is a property of the data (hundreds for 100 random-orthogonal
        # patterns); the cost per iteration is what the kernels determine
        nonlocal r, conv
        r = xb.single.EOFRotator(n_modes=k, power=1, max_iter=max_iter)
        try:
            r.fit(model)
        except RuntimeError:
            conv = False
        return r

    ms, _ = timed(fit, n=1)
    S = nlat * nlon
    it = r.n_iter_ if conv else max_iter
    print(json.dumps({"config": f"EOFRotator varimax power=1 on {k} modes of a {T}x({nlat}x{nlon}) EOF model",
                      "loadings_GB": S * k * 4 / 1e9, "fit_ms": ms, "iterations": it, "ms_per_iteration": ms / max(it, 1),
                      "GBps_per_iteration": S * k * 4 / 1e9 / (ms / max(it, 1)) * 1e3,
                      "iterations_tensor_core": getattr(r, "n_iter_tc_", 0), "converged_within_cap": conv}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    if what in ("mca", "all"):
        mca(scale)
    if what in ("rot", "all"):
        rot(scale)
