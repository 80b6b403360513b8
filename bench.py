#!/usr/bin/env python
"""bench.py — EOF.fit throughput (GB/s of fp32 time x space input streamed) on B200.

Contract (one JSON line on rank 0):
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|small]
  N > 1 is launched by torchrun, one rank per GPU, feature (space) axis sharded across ranks.

* a "step" is one EOF(n_modes, n_iter=4 randomized SVD).fit over the synthetic field, X resident in HBM
  when the timed region starts (`value`), or in pinned host memory (`e2e`: H2D copy of the field and D2H of
  the singular values inside the timed region, through the public xeofs_b200.single.EOF API);
* `roofline` is the dominant kernel (the streaming product A^T W / A Y), algorithmic bytes per launch
  T*S*4 + S*lp*4 + T*lp*4 over its CUDA-event duration measured live on the launch stream;
* `cpu_baseline` / `--impl reference`: the oracle (numpy restatement of the reference's fit calling the installed
  sklearn randomized_svd — the reference package itself needs xarray + dask, absent from this image, see
  DESIGN.md) on a bounded column sample of the same field with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (T, n_lat, n_lon, n_modes, n_iter, extra EOF kwargs)
    "c2": (8760, 721, 1440, 50, 4, {}),                                        # BASELINE configs[1]
    "c4": (8760, 1440, 2880, 100, 4, {"use_coslat": True, "standardize": True}),  # configs[3] (strong scaling)
    "small": (2920, 25, 53, 10, 4, {}),                                        # configs[0] shape (plumbing)
    "mid": (8760, 90, 1440, 50, 4, {}),                                        # 1/8 of c2 (quick looks)
}
RANDOM_STATE = 5
CPU_SAMPLE_COLS = 16384  # columns of the field the CPU legs fit per step (8760 x 16384 fp32 = 0.57 GB)


# ------------------------------------------------------------------------------------------------ synthetic field
def planted_field_device(T, S, r, seed, device, sigma0=1.0e6, decay=0.9, eps=0.1, offset=280.0, nan_cols=None,
                         sparse=0.0):
    """offset + sum_i sigma_i u_i v_i^T + eps N(0,1) built on the device in row blocks.
    sparse > 0: every spatial pattern v_i lives on a random fraction `sparse` of the features (simple structure for
    the varimax workload)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    U = torch.linalg.qr(torch.randn((T, r), generator=g, device=device))[0]
    V = torch.randn((S, r), generator=g, device=device)
    if sparse > 0:
        V *= (torch.rand((S, r), generator=g, device=device) < sparse)
    V /= V.norm(dim=0, keepdim=True)  # near-orthonormal for S >> r; avoids a QR of an S x r matrix
    sig = sigma0 * decay ** torch.arange(r, device=device, dtype=torch.float32)
    X = torch.empty((T, S), dtype=torch.float32, device=device)
    Vt = (V * sig[None, :]).t().contiguous()
    rows = max(1, int((1 << 30) // (4 * S)))
    for t0 in range(0, T, rows):
        t1 = min(T, t0 + rows)
        blk = X[t0:t1]
        blk.normal_(0.0, eps, generator=g)
        blk.addmm_(U[t0:t1], Vt)
        blk.add_(offset)
    del V, Vt
    return X


def planted_field_host(T, S, r, seed, sigma0=1.0e6, decay=0.9, eps=0.1, offset=280.0):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((T, r)))
    V = rng.standard_normal((S, r))
    V /= np.linalg.norm(V, axis=0, keepdims=True)
    sig = sigma0 * decay ** np.arange(r)
    X = (U * sig).astype(np.float32) @ V.T.astype(np.float32)
    X += (eps * rng.standard_normal((T, S), dtype=np.float32))
    X += np.float32(offset)
    return X


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_fit_sample(X_host, n_lat_rows, n_lon, k, n_iter, kw, threads):
    """One oracle fit (reference arithmetic: fp64 Scaler passes + sklearn randomized_svd) on a host sample."""
    from threadpoolctl import threadpool_limits

    from oracle import eof as oeof

    T = X_host.shape[0]
    lat = np.linspace(60.0, -60.0, n_lat_rows)
    coords = {"lat": lat, "lon": np.arange(n_lon) * (360.0 / n_lon)}
    t0 = time.perf_counter()
    with threadpool_limits(limits=threads):
        o = oeof.eof_fit(X_host.reshape(T, n_lat_rows, n_lon), ("time", "lat", "lon"), "time", coords=coords,
                         n_modes=k, random_state=RANDOM_STATE, solver_kwargs={"n_iter": n_iter}, **kw)
    dt = time.perf_counter() - t0
    return dt, o["singular_values"]


def sample_geometry(n_lon):
    cols = CPU_SAMPLE_COLS
    n_lon_s = min(n_lon, 1024)
    return max(1, cols // n_lon_s), n_lon_s


def run_reference(args):
    """--impl reference: the reference's CPU fit (oracle port; the package cannot be imported here) timed per step
    on a bounded column sample of the workload, all host threads.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    T, n_lat, n_lon, k, n_iter, kw = WORKLOADS[args.workload]
    rows_s, n_lon_s = sample_geometry(n_lon)
    S_s = rows_s * n_lon_s
    threads = os.cpu_count() or 1
    X = planted_field_host(T, S_s, 2 * k, seed=1)
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = cpu_fit_sample(X, rows_s, n_lon_s, k, n_iter, kw, threads)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    gbs = T * S_s * 4 / t / 1e9
    line = {
        "impl": "reference", "metric": "EOF.fit GB/s (time x space fp32 streamed)", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: EOF n_modes={k} n_iter={n_iter} on {T}x({n_lat}x{n_lon}) fp32",
                   "sample": f"{T}x{S_s} columns of it per step"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": f"{T}x{S_s} fp32 ({T * S_s * 4 / 1e9:.2f} GB) per step, oracle eof_fit "
                                   "(numpy Scaler/Sanitizer passes in fp64 + sklearn randomized_svd)"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import xeofs_b200 as xb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    T, n_lat, n_lon, k, n_iter, kw = WORKLOADS[args.workload]
    # feature sharding: weak = every rank holds one full copy-sized slab of latitude rows; strong = rows split
    if args.scaling == "weak":
        lat_rows, lat0, n_lat_total = n_lat, rank * n_lat, n_lat * world
    else:
        per = (n_lat + world - 1) // world
        lat0 = rank * per
        lat_rows, n_lat_total = max(0, min(n_lat, lat0 + per) - lat0), n_lat
    S_local = lat_rows * n_lon
    lat_all = np.linspace(90.0, -90.0, n_lat_total)
    coords = {"lat": lat_all[lat0:lat0 + lat_rows], "lon": np.arange(n_lon) * (360.0 / n_lon)}
    dims = ("time", "lat", "lon")
    X = planted_field_device(T, S_local, 2 * k, seed=1 + rank, device=device).reshape(T, lat_rows, n_lon)
    if args.land_frac > 0:
        # a land mask: a fixed fraction of the grid points is NaN at every time step (the Sanitizer's full-dimensional
        # NaN case, sanitizer.py:46-56); every 1/frac-th block of 64 points
        mask = (torch.arange(S_local, device=device) // 64) % max(2, int(round(1.0 / args.land_frac))) == 0
        X.view(T, -1)[:, mask] = float("nan")
        del mask
    bytes_local = T * S_local * 4
    total_bytes = T * n_lat_total * n_lon * 4

    def make_model():
        return xb.single.EOF(n_modes=k, random_state=RANDOM_STATE, solver_kwargs={"n_iter": n_iter},
                             distributed=world > 1, algo=args.algo, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_fit(src):
        m = make_model()
        m.fit(xb.DataArray(src, dims, coords), dim="time")
        return m

    for _ in range(args.warmup):
        m = one_fit(X)
    barrier()
    launches0 = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, kernel_n, launches = [], 0, 0
    with ClockSampler(local_rank) as clk:
        ev0.record()
        for _ in range(args.steps):
            m = one_fit(X)
            launches += m.ops.launches
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    s_vals = m.singular_values().values
    # per-kernel timing of the dominant kernel, outside the step timing (events around each product launch)
    m2 = make_model()
    m2.ops.time_products = True
    m2.fit(xb.DataArray(X, dims, coords), dim="time")
    torch.cuda.synchronize()
    prod = m2.ops.product_times()  # list of (name, ms, l)
    tmax = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = total_bytes / (ms * 1e-3) / 1e9

    # ---- e2e: host buffers through the public API (H2D of the field + D2H of the singular values timed)
    e2e = None
    if not args.no_e2e:
        try:
            del m, m2
            Xh = torch.empty((T, lat_rows, n_lon), dtype=torch.float32, pin_memory=True)
            Xh.copy_(X)
            del X
            torch.cuda.empty_cache()
            e_steps = max(1, min(args.steps, args.e2e_steps))
            barrier()
            t_e = []
            for i in range(1 + e_steps):
                barrier()
                t0 = time.perf_counter()
                me = one_fit(Xh)  # the API uploads the host field (pinned -> device) itself
                sv = me.singular_values().values  # D2H of the result
                barrier()
                if i > 0:
                    t_e.append(time.perf_counter() - t0)
                del me
            te = torch.tensor([float(np.mean(t_e))], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e = {"value": total_bytes / float(te.item()) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": bytes_local,
                   "d2h_bytes_per_step": int(sv.nbytes), "steps": e_steps, "ms_per_step": float(te.item()) * 1e3}
        except Exception as exc:  # host RAM too small for the pinned copy, etc.
            e2e = {"value": None, "unit": "GB/s", "error": f"{type(exc).__name__}: {exc}"[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    roof = None
    if prod:
        by = {}
        for name, t_ms, l in prod:
            by.setdefault(name, []).append(t_ms)
        lp = (min(k + 10, T, S_local) + 15) // 16 * 16
        alg = bytes_local + S_local * lp * 4 + T * lp * 4
        streaming = {n: v for n, v in by.items() if n.startswith(("project_S", "project_T", "col_stats"))}
        dom = max(streaming or by, key=lambda n: sum(by[n]))
        avg_ms = float(np.mean(by[dom]))
        ach = alg / (avg_ms * 1e-3) / 1e9
        # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of this workload (a profiler
        # figure, taken once per change — not measured in this run); null when the local slab is not the profiled one
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tj.get("workload") == args.workload and bytes_local == WORKLOADS[args.workload][0] * \
                    WORKLOADS[args.workload][1] * WORKLOADS[args.workload][2] * 4 and dom in tj:
                traffic, traffic_src = float(tj[dom]), "profiles/r01_traffic.json (ncu --set full, dram read + write)"
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launch_ms": avg_ms,
                "launches_timed": len(by[dom]),
                "algorithmic_bytes_per_launch": alg,
                "per_kernel_ms": {n: float(np.mean(v)) for n, v in by.items()},
                "share_of_step": float(sum(sum(v) for v in by.values()) / ms)}
    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        rows_s, n_lon_s = sample_geometry(n_lon)
        S_s = rows_s * n_lon_s
        threads = os.cpu_count() or 1
        Xs = planted_field_host(T, S_s, 2 * k, seed=1)
        dt, _ = cpu_fit_sample(Xs, rows_s, n_lon_s, k, n_iter, kw, threads)
        cpu = {"value": T * S_s * 4 / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": f"one oracle eof_fit on {T}x{S_s} fp32 ({T * S_s * 4 / 1e9:.2f} GB) of the same synthetic "
                         f"recipe: {dt:.1f} s"}
    line = {
        "metric": "EOF.fit GB/s (time x space fp32 streamed)", "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "tf32/f32 (fp64 small matrices)",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: EOF n_modes={k} n_iter={n_iter} randomized SVD on "
                               f"{T}x({n_lat_total}x{n_lon}) fp32 ({total_bytes / 1e9:.2f} GB), "
                               f"{'feature-sharded over %d GPUs' % world if world > 1 else '1 GPU'}",
                   "l2": "inputs larger than L2 (no flush needed)", "algo": args.algo,
                   "extra": dict(kw, **({"land_frac": args.land_frac} if args.land_frac > 0 else {}))},
        "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clk.summary(),
        "singular_values_head": [float(v) for v in s_vals[:3]],
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--land-frac", type=float, default=0.0, help="fraction of grid points that are NaN at every step")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: warmup {args.warmup} < 3", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
