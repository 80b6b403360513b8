#!/usr/bin/env python
"""In-situ timeline of one EOF.fit (torch.profiler / CUPTI): wall time, sum of kernel time, top kernels and the
largest idle gaps between consecutive kernels.  usage: python tools/profile_fit.py [c2|c4|mid|c3|c5]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import xeofs_b200 as xb  # noqa: E402


def make_fit(wl):
    """The timed callable of a workload: c2 / c4 / mid ... (EOF.fit), c3 (MCA.fit), c5 (EOFRotator.fit)."""
    import numpy as np
    dev = torch.device("cuda")
    if wl == "c3":
        T, n_lat, n_lon, k = (bench.C3[x] for x in ("T", "n_lat", "n_lon", "k"))
        U = bench.temporal_factors(T, 2 * k, 2, dev)
        mk = lambda seed: bench.planted_field_device(T, n_lat, n_lon, 0, n_lat, 2 * k, seed, dev,  # noqa: E731
                                                     sigma0=1e5, decay=0.85, eps=0.05, U=U)
        X, Y = mk(2), mk(3)
        coords = {"lat": np.linspace(89.75, -89.75, n_lat), "lon": np.arange(n_lon) * 0.5}
        return lambda: xb.cross.MCA(n_modes=k, random_state=5, use_pca=False).fit(
            xb.DataArray(X, bench.DIMS, coords), xb.DataArray(Y, bench.DIMS, coords), dim="time")
    if wl == "c5":
        T, n_lat, n_lon, k = (bench.C5[x] for x in ("T", "n_lat", "n_lon", "k"))
        X = bench.planted_field_device(T, n_lat, n_lon, 0, n_lat, 2 * k, 4, dev, decay=0.995, sparse=0.05)
        coords = {"lat": np.linspace(89.9, -89.9, n_lat), "lon": np.arange(n_lon) * 0.125}
        model = xb.single.EOF(n_modes=k, use_coslat=True, random_state=5, solver_kwargs={"n_iter": 4})
        model.fit(xb.DataArray(X, bench.DIMS, coords), dim="time")
        del X
        torch.cuda.empty_cache()
        return lambda: xb.single.EOFRotator(n_modes=k, power=1, max_iter=1000).fit(model)
    T, n_lat, n_lon, k, n_iter, kw = bench.WORKLOADS[wl]
    X = bench.planted_field_device(T, n_lat, n_lon, 0, n_lat, 2 * k, 1, dev, **bench.FIELD.get(wl, {}))
    coords = {"lat": np.linspace(90, -90, n_lat), "lon": np.arange(n_lon) * (360.0 / n_lon)}

    def fit():
        m = xb.single.EOF(n_modes=k, random_state=5, solver_kwargs={"n_iter": n_iter}, **kw)
        m.fit(xb.DataArray(X, ("time", "lat", "lon"), coords), dim="time")
        return m
    return fit


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    fit = make_fit(wl)
    for _ in range(2 if wl == "c5" else 3):
        fit()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fit()
        e1.record()
        torch.cuda.synchronize()
    print(f"fit wall (events): {e0.elapsed_time(e1):.2f} ms")
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    tot = sum(e.time_range.elapsed_us() for e in evs) / 1e3
    print(f"device activities: {len(evs)}, busy {tot:.2f} ms")
    gaps = []
    for a, b in zip(evs[:-1], evs[1:]):
        g = b.time_range.start - a.time_range.end
        gaps.append((g, a.name[:50], b.name[:50]))
    gaps.sort(reverse=True)
    print("largest gaps (us): ")
    for g, a, b in gaps[:12]:
        print(f"  {g:9.1f}  after {a}  before {b}")
    print(f"sum of gaps: {sum(g for g, _, _ in gaps) / 1e3:.2f} ms")
    agg = {}
    for e in evs:
        a = agg.setdefault(e.name[:70], [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.elapsed_us() / 1e3
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"  {ms:8.3f} ms {n:4d}x  {name}")


if __name__ == "__main__":
    main()
