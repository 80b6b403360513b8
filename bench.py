#!/usr/bin/env python
"""bench.py — EOF.fit throughput (GB/s of fp32 time x space input streamed) on B200.

Contract (one JSON line on rank 0):
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5|small|mid]
  N > 1 is launched by torchrun, one rank per GPU, feature (space) axis sharded across ranks.

* headline (`--workload c2`, the default): a "step" is one EOF(n_modes=50, n_iter=4 randomized SVD).fit over the
  synthetic field of BASELINE configs[1], X resident in HBM when the timed region starts (`value`), or in pinned host
  memory (`e2e`: H2D copy of the field and D2H of the singular values inside the timed region, through the public
  xeofs_b200.single.EOF API).  With N GPUs every rank holds one config-2 slab (weak scaling).
* every c2 line also carries `strong_c4`: BASELINE configs[3] (8760 x (1440 x 2880), n_modes=100, coslat +
  standardize, 145 GB) as ONE fixed field sharded over the N ranks — the north star's strong-scaling case — timed the
  same way (>= 10 steps), with its singular values compared across N (the field is seeded by GLOBAL latitude row, so
  every N decomposes the same matrix).
* `parity`: the run checks its own results instead of printing them: singular values against the committed oracle
  values of the same seeded field (profiles/r02_expected_sv.json, written by tools/parity_full.py from a host fp64 run
  of oracle.eof_fit on the full field), and the reference's own invariant transform(X) == scores
  (tests/models/single/test_eof.py:364-391).  A failed check makes the line say so and the process exit non-zero.
* `roofline` is the dominant kernel (the streaming product A^T W / A Y), algorithmic bytes per launch
  T*S*4 + S*lp*4 + T*lp*4 over its CUDA-event duration measured live on the launch stream;
* `cpu_baseline` / `--impl reference`: the oracle (numpy restatement of the reference's fit calling the installed
  sklearn randomized_svd — the reference package itself needs xarray + dask, absent from this image, see
  DESIGN.md) on a bounded column sample of the same field with all host threads.
* `--workload c3` (MCA on two 8760 x (360 x 720) fields, implicit cross-covariance) and `--workload c5` (EOFRotator
  varimax on 100 modes of a config-4-wide model) print the same kind of line for BASELINE configs[2] / configs[4].
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
    "c4mid": (8760, 176, 2880, 100, 4, {"use_coslat": True, "standardize": True}),  # 1/8 of c4 (quick looks)
}
# synthetic-field parameters per workload: every requested mode must stand clear of the noise floor eps (sqrt(T) +
# sqrt(S)) — noise-level singular values are nearly degenerate and no two summation orders (1 GPU vs 8) agree on them
# to 1e-4.  config 4 asks for 100 modes: sigma_i = 1e6 0.95^i keeps sigma_100 = 6.2e3 a factor 29 above it
FIELD = {"c4": dict(seed=4, decay=0.95), "c4mid": dict(seed=4, decay=0.95)}
C3 = dict(T=8760, n_lat=360, n_lon=720, k=20)          # BASELINE configs[2]
C5 = dict(T=1024, n_lat=1440, n_lon=2880, k=100)       # configs[4]: the loadings are (1440 x 2880) x 100
RANDOM_STATE = 5
CPU_SAMPLE_COLS = 32768  # columns of the field the CPU legs fit per step (22 latitude rows of config 2: 1.11 GB, ~6 s;
# the CPU rate grows with the sample — 0.09 GB/s at 16 384 columns, 0.18 at S/8, 0.166 on the whole field
# (profiles/r02_reference_arm_s8.json, r02_parity_full.json) — so the bounded sample is as wide as the time budget allows)
EXPECTED = os.path.join(ROOT, "profiles", "r02_expected_sv.json")
DIMS = ("time", "lat", "lon")


# ------------------------------------------------------------------------------------------------ synthetic field
def _row_generator(device, seed, row):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed) * 1000003 + int(row))
    return g


def temporal_factors(T, r, seed, device):
    """The r orthonormal temporal factors every rank (and every N) shares."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return torch.linalg.qr(torch.randn((T, r), generator=g, device=device))[0]


def planted_field_device(T, lat_rows, n_lon, lat0, n_lat_total, r, seed, device, sigma0=1.0e6, decay=0.9, eps=0.1,
                         offset=280.0, sparse=0.0, U=None):
    """offset + sum_i sigma_i u_i v_i^T + eps N(0,1) for the latitude rows [lat0, lat0 + lat_rows) of a
    field with n_lat_total rows, built on the device.  Everything random about a latitude row — its piece of the
    spatial patterns and its noise — is drawn from a generator seeded by (seed, GLOBAL row index): any sharding of
    the rows over ranks assembles the same global field.  sparse > 0: every spatial pattern lives on a random
    fraction `sparse` of the grid points (simple structure for the varimax workload)."""
    import torch

    if U is None:
        U = temporal_factors(T, r, seed, device)
    sig = sigma0 * decay ** torch.arange(r, device=device, dtype=torch.float32)
    Us = (U * sig[None, :]).contiguous()
    scale = 1.0 / float(np.sqrt(n_lat_total * n_lon * (sparse if sparse > 0 else 1.0)))  # |v_i| ~ 1 without a collective
    X = torch.empty((T, lat_rows, n_lon), dtype=torch.float32, device=device)
    blk = torch.empty((T, n_lon), dtype=torch.float32, device=device)
    for i in range(lat_rows):
        g = _row_generator(device, seed, lat0 + i)
        V = torch.randn((n_lon, r), generator=g, device=device)
        if sparse > 0:
            V *= (torch.rand((n_lon, r), generator=g, device=device) < sparse)
        blk.normal_(0.0, eps, generator=g)
        blk.addmm_(Us, V.t(), alpha=scale)
        blk.add_(offset)
        X[:, i, :] = blk
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


def shard_rows(n_lat, world, rank, scaling):
    """(local rows, first global row, total rows).  weak: one full slab per rank; strong: the rows split."""
    if scaling == "weak":
        return n_lat, rank * n_lat, n_lat * world
    per = (n_lat + world - 1) // world
    lat0 = rank * per
    return max(0, min(n_lat, lat0 + per) - lat0), lat0, n_lat


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line), every 100 ms.  The
    samples are NVML queries made in this process (what nvidia-smi prints); forking one nvidia-smi per sample, as this
    class first did, cost the timed region ~10 % (52.7 -> 59.4 ms per config-2 fit in the same session).  Without the
    NVML binding: ONE background `nvidia-smi -lms 100`, started before and stopped after the region, as in the recipe."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._th, self._proc, self.source = index, [], threading.Event(), None, None, None
        self._h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            self._nv = pynvml
            self.source = "nvml"
        except Exception:
            self._h = None

    def _run(self):
        nv = self._nv
        bits = [getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)]
        reasons_of = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                r = int(reasons_of(self._h))
                self.rows.append([str(sm), str(mx), "", *("Active" if r & b else "Not Active" for b in bits)])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self._h is not None:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()
        else:
            try:
                self._proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                               "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
                self.source = "nvidia-smi -lms 100"
            except Exception:
                self._proc = None
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._th is not None:
            self._th.join(timeout=6)
        if self._proc is not None:
            self._proc.terminate()
            try:
                out, _ = self._proc.communicate(timeout=5)
                self.rows = [[c.strip() for c in ln.split(",")] for ln in out.strip().splitlines() if ln.strip()]
            except Exception:
                pass

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.source}


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
        o = oeof.eof_fit(X_host.reshape(T, n_lat_rows, n_lon), DIMS, "time", coords=coords,
                         n_modes=k, random_state=RANDOM_STATE, solver_kwargs={"n_iter": n_iter}, **kw)
    dt = time.perf_counter() - t0
    return dt, o["singular_values"]


def sample_geometry(n_lon, cols=CPU_SAMPLE_COLS):
    return max(1, cols // n_lon), n_lon


def run_reference(args):
    """--impl reference: the reference's CPU fit (oracle port; the package cannot be imported here) timed per step
    on a bounded column sample of the workload, all host threads.  Rank 0 only.  `--ref-cols` widens the sample
    (S/8 of config 2 = 129 780 columns was run once: profiles/r02_reference_arm_s8.json)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = args.workload if args.workload in WORKLOADS else "c2"
    T, n_lat, n_lon, k, n_iter, kw = WORKLOADS[wl]
    rows_s, n_lon_s = sample_geometry(n_lon, args.ref_cols)
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
    S_full = n_lat * n_lon
    line = {
        "impl": "reference", "metric": "EOF.fit GB/s (time x space fp32 streamed)", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{wl}: EOF n_modes={k} n_iter={n_iter} on {T}x({n_lat}x{n_lon}) fp32",
                   "sample": f"{T}x{S_s} columns of it per step (1/{S_full / S_s:.1f} of the features)",
                   "extrapolated_full_fit_s": t * S_full / S_s,
                   "extrapolation": "linear in the number of features (every pass of the CPU fit is O(T S l)); a label, "
                                    "not a measurement"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": f"{T}x{S_s} fp32 ({T * S_s * 4 / 1e9:.2f} GB) per step, oracle eof_fit "
                                   "(numpy Scaler/Sanitizer passes in fp64 + sklearn randomized_svd)"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ helpers of our arm
class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device(f"cuda:{self.local_rank}")
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def load_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)", peaks
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)", peaks


def load_traffic(key):
    """DRAM bytes per launch (ncu --set full, dram read + write) of the kernels of workload `key`, generated from the
    ncu CSV by tools/ncu_traffic.py; None if not captured."""
    for name in ("r02b_traffic.json", "r02_traffic.json"):  # the capture of the final code of the round, else the first one
        try:
            hit = json.load(open(os.path.join(ROOT, "profiles", name))).get(key)
        except Exception:
            hit = None
        if hit:
            return dict(hit, _file="profiles/" + name)
    return None


def timed_steps(d, fn, warmup, steps, clock=True):
    """W untimed + K timed calls of fn bracketed by barrier + synchronize; ms per step (max over ranks), the last
    return value, the launches of the timed calls and the clock samples."""
    torch = d.torch
    out = None
    for _ in range(warmup):
        out = fn()
    d.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    clk = ClockSampler(d.local_rank) if clock else None
    if clk:
        clk.__enter__()
    ev0.record()
    for _ in range(steps):
        out = fn()
        launches += getattr(getattr(out, "ops", None), "launches", 0)
    ev1.record()
    d.barrier()
    if clk:
        clk.__exit__()
    ms = d.max_(ev0.elapsed_time(ev1) / steps)
    return ms, out, launches, (clk.summary() if clk else None)


def product_breakdown(model_factory, fit):
    """One extra fit with CUDA events around every streaming product on the launch stream."""
    m = model_factory()
    m.ops.time_products = True
    fit(m)
    prod = m.ops.product_times()
    by = {}
    for name, t_ms, l in prod:
        by.setdefault(name, []).append(t_ms)
    return by


def alg_bytes_of(tag, alg_bytes, field_bytes):
    """Algorithmic bytes of one launch of a streaming product: the fp32 field once plus the k-column operand and
    result (alg_bytes); the passes on the fp16 copy read half the field bytes, the pass that writes it 1.5 times."""
    if tag.endswith("_h16"):
        return alg_bytes - field_bytes / 2
    if tag.endswith("_wcopy"):
        return alg_bytes + field_bytes / 2
    return alg_bytes


def roofline_of(by, ms_step, alg_bytes, peak, peak_src, traffic, field_bytes=0):
    if not by:
        return None
    streaming = {n: v for n, v in by.items() if n.startswith(("project_S", "project_T", "col_stats"))}
    dom = max(streaming or by, key=lambda n: sum(by[n]))
    avg_ms = float(np.mean(by[dom]))
    alg_all = alg_bytes
    alg_bytes = alg_bytes_of(dom, alg_all, field_bytes)
    ach = alg_bytes / (avg_ms * 1e-3) / 1e9
    tr = (traffic or {}).get(dom)
    return {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": tr, "traffic_source": (traffic or {}).get("_file", "profiles") + " (ncu --set full on this code, "
                                             "dram read + write per launch, tools/ncu_traffic.py)" if tr else None,
            "peak_source": peak_src, "launch_ms": avg_ms, "launches_timed": len(by[dom]),
            "algorithmic_bytes_per_launch": alg_bytes,
            "per_kernel_ms": {n: float(np.mean(v)) for n, v in by.items()},
            "per_kernel_launches": {n: len(v) for n, v in by.items()},
            "per_kernel_frac_of_peak": {n: alg_bytes_of(n, alg_all, field_bytes) / (float(np.mean(v)) * 1e-3) / 1e9 / peak
                                        for n, v in streaming.items() if not n.startswith("apply")},
            "share_of_step": float(sum(sum(v) for v in by.values()) / ms_step)}


def expected_sv(key):
    try:
        return json.load(open(EXPECTED)).get(key)
    except Exception:
        return None


def check_sv(s, key, rtol=1e-4):
    """Singular values against the committed expectation of the same seeded field."""
    exp = expected_sv(key)
    if not exp:
        return {"checked": False, "why": f"no entry {key!r} in profiles/r02_expected_sv.json"}
    ref = np.asarray(exp["s"], dtype=np.float64)
    n = min(len(ref), len(s))
    rel = np.abs(np.asarray(s[:n], dtype=np.float64) / ref[:n] - 1.0)
    err = float(rel.max())
    return {"checked": True, "against": exp.get("source"), "modes": n, "max_rel_err": err,
            "worst_mode": int(rel.argmax()) + 1, "median_rel_err": float(np.median(rel)), "rtol": rtol,
            "ok": bool(err <= rtol)}


def transform_invariant(m, X, coords):
    """The reference's own invariant (tests/models/single/test_eof.py:364-391): transform(X) == scores."""
    import xeofs_b200 as xb

    sc = m.scores().values
    tr = m.transform(xb.DataArray(X, DIMS, coords)).values
    num = np.linalg.norm(sc - tr, axis=0)
    den = np.linalg.norm(sc, axis=0)
    err = float(np.max(num / den))
    return {"max_rel_err": err, "rtol": 1e-3, "ok": bool(err <= 1e-3)}


# ------------------------------------------------------------------------------------------------ EOF workloads
def eof_case(d, args, wl, scaling, steps, warmup, e2e=True, cpu=True, clock=True):
    """Timed EOF fits of one workload; returns the dict of everything measured (rank 0 fills the line from it)."""
    import xeofs_b200 as xb

    torch = d.torch
    T, n_lat, n_lon, k, n_iter, kw = WORKLOADS[wl]
    lat_rows, lat0, n_lat_total = shard_rows(n_lat, d.world, d.rank, scaling)
    S_local = lat_rows * n_lon
    lat_all = np.linspace(90.0, -90.0, n_lat_total)
    coords = {"lat": lat_all[lat0:lat0 + lat_rows], "lon": np.arange(n_lon) * (360.0 / n_lon)}
    fp = FIELD.get(wl, dict(seed=1, decay=0.9))
    X = planted_field_device(T, lat_rows, n_lon, lat0, n_lat_total, 2 * k, fp["seed"], d.device, decay=fp["decay"])
    if args.land_frac > 0:
        # a land mask: a fixed fraction of the grid points is NaN at every time step (the Sanitizer's full-dimensional
        # NaN case, sanitizer.py:46-56); every 1/frac-th block of 64 points
        mask = (torch.arange(S_local, device=d.device) // 64) % max(2, int(round(1.0 / args.land_frac))) == 0
        X.view(T, -1)[:, mask] = float("nan")
        del mask
    bytes_local = T * S_local * 4
    total_bytes = T * n_lat_total * n_lon * 4

    def make_model():
        return xb.single.EOF(n_modes=k, random_state=RANDOM_STATE, solver_kwargs={"n_iter": n_iter},
                             distributed=d.world > 1, algo=args.algo, **kw)

    hold = {"X": X}  # the one reference the closures keep: dropping it frees the field
    del X

    def one_fit(src=None, m=None):
        m = m or make_model()
        m.fit(xb.DataArray(hold["X"] if src is None else src, DIMS, coords), dim="time")
        return m

    ms, m, launches, clocks = timed_steps(d, one_fit, warmup, steps, clock)
    s_vals = m.singular_values().values
    res = {"ms": ms, "value": total_bytes / (ms * 1e-3) / 1e9, "launches": launches, "clocks": clocks,
           "s": [float(v) for v in s_vals], "total_bytes": total_bytes, "bytes_local": bytes_local,
           "desc": f"{wl}: EOF n_modes={k} n_iter={n_iter} randomized SVD on {T}x({n_lat_total}x{n_lon}) fp32 "
                   f"({total_bytes / 1e9:.2f} GB), "
                   f"{'feature-sharded over %d GPUs' % d.world if d.world > 1 else '1 GPU'}",
           "extra": dict(kw, **({"land_frac": args.land_frac} if args.land_frac > 0 else {}))}
    res["invariant"] = transform_invariant(m, hold["X"], coords)
    by = product_breakdown(make_model, lambda mm: one_fit(None, mm))
    lp = (min(k + 10, T, S_local) + 15) // 16 * 16
    res["by"] = by
    res["alg_bytes"] = bytes_local + S_local * lp * 4 + T * lp * 4
    full_slab = bytes_local == T * n_lat * n_lon * 4
    res["traffic"] = load_traffic(wl) if full_slab else None
    del m

    # ---- e2e: host buffers through the public API (H2D of the field + D2H of the singular values timed)
    res["e2e"] = None
    if e2e:
        try:
            Xh = torch.empty((T, lat_rows, n_lon), dtype=torch.float32, pin_memory=True)
            Xh.copy_(hold["X"])
            hold["X"] = None
            torch.cuda.empty_cache()
            e_steps = max(1, min(steps, args.e2e_steps))
            t_e = []
            for i in range(1 + e_steps):
                d.barrier()
                t0 = time.perf_counter()
                me = one_fit(Xh)  # the API uploads the host field (pinned -> device) itself
                sv = me.singular_values().values  # D2H of the result
                d.barrier()
                if i > 0:
                    t_e.append(time.perf_counter() - t0)
                del me
            te = d.max_(float(np.mean(t_e)))
            res["e2e"] = {"value": total_bytes / te / 1e9, "unit": "GB/s", "h2d_bytes_per_step": bytes_local,
                          "d2h_bytes_per_step": int(sv.nbytes), "steps": e_steps, "ms_per_step": te * 1e3}
            del Xh
        except Exception as exc:  # host RAM too small for the pinned copy, etc.
            res["e2e"] = {"value": None, "unit": "GB/s", "error": f"{type(exc).__name__}: {exc}"[:200]}
    hold["X"] = None
    torch.cuda.empty_cache()

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    res["cpu"] = None
    if cpu and d.world == 1:
        rows_s, n_lon_s = sample_geometry(n_lon)
        S_s = rows_s * n_lon_s
        threads = os.cpu_count() or 1
        Xs = planted_field_host(T, S_s, 2 * k, seed=1)
        dt, _ = cpu_fit_sample(Xs, rows_s, n_lon_s, k, n_iter, kw, threads)
        res["cpu"] = {"value": T * S_s * 4 / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                      "sample": f"one oracle eof_fit on {T}x{S_s} fp32 ({T * S_s * 4 / 1e9:.2f} GB) of the same "
                                f"synthetic recipe: {dt:.1f} s (the whole config-2 field, once: 219.8 s = 0.166 GB/s, "
                                "profiles/r02_parity_full.json)"}
    return res


def run_eof(args):
    d = Dist()
    peak, peak_src, _ = load_peak()
    wl = args.workload
    scaling = args.scaling if d.world > 1 else "weak"
    res = eof_case(d, args, wl, scaling, args.steps, args.warmup, e2e=not args.no_e2e, cpu=not args.no_cpu)
    strong = None
    if wl == "c2" and not args.no_strong_c4:
        # the north star's strong-scaling case rides on every headline line: configs[3] as ONE field over the N ranks
        try:
            st = eof_case(d, args, args.strong_workload, "strong", max(10, args.steps), 3, e2e=False, cpu=False,
                          clock=False)
            sroof = roofline_of(st["by"], st["ms"], st["alg_bytes"], peak, peak_src, st["traffic"], st["bytes_local"])
            key = f"{args.strong_workload}_strong"
            chk = check_sv(st["s"], key, rtol=1e-4)
            n1 = (expected_sv(key) or {}).get("ms_per_step_n1")
            strong = {"workload": st["desc"], "scaling": "strong", "steps": max(10, args.steps), "warmup": 3,
                      "ms_per_step": st["ms"], "value": st["value"], "unit": "GB/s",
                      "n1_ms_per_step_committed": n1,
                      "speedup_vs_n1": (n1 / st["ms"]) if n1 else None,
                      "share_of_step": sroof["share_of_step"] if sroof else None,
                      "per_kernel_ms": sroof["per_kernel_ms"] if sroof else None,
                      "roofline_frac_dominant": sroof["frac"] if sroof else None,
                      "s_head": st["s"][:5], "parity_vs_n1": chk, "invariant_transform_eq_scores": st["invariant"]}
        except Exception as exc:
            strong = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    # BASELINE configs[2] and [4] ride along on the one-GPU headline line in compact form (the full lines:
    # --workload c3 / c5), so that every driver run records them
    models = None
    if wl == "c2" and d.world == 1 and not args.no_models:
        models = {}
        for name, case in (("c3_mca", mca_case), ("c5_varimax", rotator_case)):
            try:
                ln, okm = case(d, args, cpu=False)
                rf = ln.get("roofline") or {}
                models[name] = {"workload": ln["config"]["workload"], "metric": ln["metric"], "value": ln["value"],
                                "unit": ln["unit"], "ms_per_step": ln["ms_per_step"], "steps": ln["steps"],
                                "roofline": {k: rf.get(k) for k in ("bound", "kernel", "achieved", "peak", "unit", "frac",
                                                                    "launch_ms", "share_of_step", "ms_per_iteration")},
                                "parity": ln["parity"], "ok": okm}
                for extra in ("ms_per_step_without_total_squared_covariance", "iterations"):
                    if extra in ln:
                        models[name][extra] = ln[extra]
            except Exception as exc:
                models[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if d.rank != 0:
        d.close()
        return 0
    roof = roofline_of(res["by"], res["ms"], res["alg_bytes"], peak, peak_src, res["traffic"], res["bytes_local"])
    full = wl in ("c2",) and d.world == 1 and args.land_frac == 0 and args.algo == "auto"
    parity = {"singular_values": check_sv(res["s"], wl, 1e-4) if full else
              {"checked": False, "why": "the committed oracle values are for the one-GPU config-2 field"},
              "transform_eq_scores": res["invariant"]}
    ok = parity["transform_eq_scores"]["ok"] and parity["singular_values"].get("ok", True)
    if strong and "error" not in strong:
        ok = ok and strong["invariant_transform_eq_scores"]["ok"] and strong["parity_vs_n1"].get("ok", True)
    if models:
        ok = ok and all(v.get("ok", False) for v in models.values())
    parity["ok"] = bool(ok)
    line = {
        "metric": "EOF.fit GB/s (time x space fp32 streamed)", "value": res["value"], "unit": "GB/s",
        "n_gpus": d.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms"],
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "tf32/f32 (fp64 small matrices; power iterations 3-8 on an fp16 copy of the preprocessed matrix)",
        "data": "synthetic",
        "config": {"workload": res["desc"], "l2": "inputs larger than L2 (no flush needed)", "algo": args.algo,
                   "extra": res["extra"]},
        "e2e": res["e2e"], "gpu_launches": res["launches"], "roofline": roof, "cpu_baseline": res["cpu"],
        "clocks": res["clocks"], "parity": parity, "strong_c4": strong, "models": models,
        "singular_values_head": res["s"][:3],
    }
    print(json.dumps(line), flush=True)
    d.close()
    return 0 if ok else 3


# ------------------------------------------------------------------------------------------------ MCA (config 3)
def run_mca(args):
    d = Dist()
    line, ok = mca_case(d, args, cpu=not args.no_cpu)
    if d.rank == 0:
        print(json.dumps(line), flush=True)
    d.close()
    return 0 if ok else 3


def mca_case(d, args, cpu=True):
    """BASELINE configs[2]: MCA n_modes=20 on two 8760 x (360 x 720) fields, cross-covariance applied implicitly.
    value = bytes of both fields / fit time; the fit includes the total squared covariance (cpcca.py:197)."""
    import xeofs_b200 as xb

    torch = d.torch
    peak, peak_src, _ = load_peak()
    T, n_lat, n_lon, k = C3["T"], int(C3["n_lat"] * args.scale), C3["n_lon"], C3["k"]
    lat_rows, lat0, n_lat_total = shard_rows(n_lat, d.world, d.rank, "weak" if d.world == 1 else args.scaling)
    U = temporal_factors(T, 2 * k, 2, d.device)  # both fields share their temporal factors: a planted cross-covariance
    mk = lambda seed: planted_field_device(T, lat_rows, n_lon, lat0, n_lat_total, 2 * k, seed, d.device,  # noqa: E731
                                           sigma0=1e5, decay=0.85, eps=0.05, U=U)
    X, Y = mk(2), mk(3)
    coords = {"lat": np.linspace(89.75, -89.75, n_lat_total)[lat0:lat0 + lat_rows], "lon": np.arange(n_lon) * 0.5}
    S_local = lat_rows * n_lon
    total_bytes = 2 * T * n_lat_total * n_lon * 4

    def make_model(tsc=True):
        return xb.cross.MCA(n_modes=k, random_state=RANDOM_STATE, use_pca=False, total_squared_covariance=tsc,
                            distributed=d.world > 1)

    def one_fit(m=None, tsc=True):
        m = m or make_model(tsc)
        return m.fit(xb.DataArray(X, DIMS, coords), xb.DataArray(Y, DIMS, coords), dim="time")

    ms_no_tsc, _, _, _ = timed_steps(d, lambda: one_fit(tsc=False), args.warmup, max(2, args.steps // 2), clock=False)
    ms, m, launches, clocks = timed_steps(d, one_fit, args.warmup, args.steps)
    s = m.singular_values().values
    tsc = m.total_squared_covariance()
    by = product_breakdown(lambda: make_model(False), lambda mm: one_fit(mm))
    lp = (k + 10 + 15) // 16 * 16
    alg = T * S_local * 4 + S_local * lp * 4 + T * lp * 4
    roof = roofline_of(by, ms_no_tsc, alg, peak, peak_src, load_traffic("c3"), T * S_local * 4)
    # parity inside the run: sum of squared singular values can not exceed sum |C|^2, the scores reproduce s:
    # s_m = scores1_m . scores2_m / (n - 1)   (cpcca.py:204-208 with C = Q1 s Q2^T)
    s1, s2 = m.scores()
    sdot = (s1.values * s2.values).sum(0) / (T - 1)
    inv_err = float(np.max(np.abs(sdot / s - 1)))
    parity = {"scores_reproduce_singular_values": {"max_rel_err": inv_err, "rtol": 1e-3, "ok": bool(inv_err < 1e-3)},
              "squared_covariance_le_total": bool(float((s ** 2).sum()) <= tsc * (1 + 2e-5)),
              "singular_values": check_sv([float(v) for v in s], "c3", 1e-4) if (d.world == 1 and args.scale == 1.0)
              else {"checked": False}}
    parity["ok"] = bool(parity["scores_reproduce_singular_values"]["ok"] and parity["squared_covariance_le_total"]
                        and parity["singular_values"].get("ok", True))
    want_cpu, cpu = cpu, None
    if d.world == 1 and want_cpu:
        from threadpoolctl import threadpool_limits

        from oracle import mca as omca
        cols = 4096
        rng = np.random.default_rng(0)
        Uh, _ = np.linalg.qr(rng.standard_normal((T, 2 * k)))
        sig = 1e5 * 0.85 ** np.arange(2 * k)
        mkh = lambda: ((Uh * sig) @ (rng.standard_normal((2 * k, cols)) / np.sqrt(cols)) + 0.05 *  # noqa: E731
                       rng.standard_normal((T, cols)) + 280.0).astype(np.float32)
        Xh, Yh = mkh(), mkh()
        t0 = time.perf_counter()
        with threadpool_limits(limits=os.cpu_count() or 1):
            omca.mca_fit(Xh, Yh, ("time", "x"), ("time", "x"), "time", n_modes=k, use_pca=False,
                         random_state=RANDOM_STATE)
        dt = time.perf_counter() - t0
        cpu = {"value": 2 * T * cols * 4 / dt / 1e9, "unit": "GB/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"one oracle mca_fit (explicit C = X^T Y/(n-1), {cols} x {cols}, then sklearn randomized_svd) on "
                         f"two {T}x{cols} fields: {dt:.1f} s; the explicit C grows with S^2 and cannot be formed at "
                         "the full size"}
    del X, Y, m
    torch.cuda.empty_cache()
    line = {
        "metric": "MCA.fit GB/s (both time x space fp32 fields streamed)", "value": total_bytes / (ms * 1e-3) / 1e9,
        "unit": "GB/s", "n_gpus": d.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak" if d.world == 1 else args.scaling, "vs_baseline": None,
        "dtype": "tf32/f32 (fp64 small matrices)", "data": "synthetic",
        "config": {"workload": f"c3: MCA n_modes={k} (use_pca=False, n_iter auto=7, implicit cross-covariance, total "
                               f"squared covariance included) on two {T}x({n_lat_total}x{n_lon}) fp32 fields "
                               f"({total_bytes / 1e9:.2f} GB)", "l2": "inputs larger than L2 (no flush needed)"},
        "ms_per_step_without_total_squared_covariance": ms_no_tsc, "total_squared_covariance": tsc,
        "e2e": None, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
        "parity": parity, "singular_values_head": [float(v) for v in s[:3]],
    }
    return line, bool(parity["ok"])


# ------------------------------------------------------------------------------------------------ varimax (config 5)
def run_rotator(args):
    d = Dist()
    line, ok = rotator_case(d, args, cpu=not args.no_cpu)
    if d.rank == 0:
        print(json.dumps(line), flush=True)
    d.close()
    return 0 if ok else 3


def rotator_case(d, args, cpu=True):
    """BASELINE configs[4]: EOFRotator varimax (power=1, max_iter=1000) on 100 modes of a model as wide as config 4
    (loadings 4 147 200 x 100 = 1.66 GB).  value = loadings bytes x iterations / fit time."""
    import xeofs_b200 as xb

    torch = d.torch
    peak, peak_src, peaks = load_peak()
    T, n_lat, n_lon, k = C5["T"], int(C5["n_lat"] * args.scale), C5["n_lon"], C5["k"]
    # 100 sparse patterns of nearly equal variance (plus a tail): the EOFs come out as mixtures of them and varimax has
    # a simple structure to find
    X = planted_field_device(T, n_lat, n_lon, 0, n_lat, 2 * k, 4, d.device, decay=0.995, sparse=0.05)
    coords = {"lat": np.linspace(89.9, -89.9, n_lat), "lon": np.arange(n_lon) * 0.125}
    model = xb.single.EOF(n_modes=k, use_coslat=True, random_state=RANDOM_STATE, solver_kwargs={"n_iter": 4})
    model.fit(xb.DataArray(X, DIMS, coords), dim="time")
    del X
    torch.cuda.empty_cache()
    S = n_lat * n_lon

    def one_fit():
        r = xb.single.EOFRotator(n_modes=k, power=1, max_iter=1000)
        r.fit(model)
        return r

    ms, r, _, clocks = timed_steps(d, one_fit, max(1, args.warmup - 2), max(2, args.steps // 2))
    iters = int(r.n_iter_)
    model.ops.time_products = True
    model.ops._prod_events = []
    l0 = int(model.ops.launches)
    one_fit()
    launches = (int(model.ops.launches) - l0) * max(2, args.steps // 2)
    times = model.ops.product_times()
    sweeps = [t for n, t, _ in times if n == "varimax_sweep"]
    sweeps_x1 = [t for n, t, _ in times if n == "varimax_sweep_x1"]
    model.ops.time_products = False
    sweep_ms = float(np.mean(sweeps)) if sweeps else None
    alg = S * k * 4
    flops = 4.0 * S * k * k * 3  # both products of a sweep in 3xTF32
    tf_peak = float(peaks.get("bf16_tflops", 1678.2)) / 2.0  # TF32 dense = half the measured bf16 rate
    roof = None
    if sweep_ms:
        roof = {"bound": "tensor", "kernel": "varimax_sweep (tcgen05, 3xTF32)", "achieved": flops / (sweep_ms * 1e-3) / 1e12,
                "peak": tf_peak, "unit": "TFLOP/s", "frac": flops / (sweep_ms * 1e-3) / 1e12 / tf_peak,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 runs at half the bf16 rate)",
                "traffic": (load_traffic("c5") or {}).get("varimax_sweep"), "launch_ms": sweep_ms,
                "launches_timed": len(sweeps), "algorithmic_bytes_per_launch": alg,
                "hbm": {"achieved": alg / (sweep_ms * 1e-3) / 1e9, "peak": peak, "frac": alg / (sweep_ms * 1e-3) / 1e9 / peak},
                "ms_per_iteration": ms / max(iters, 1), "share_of_step": float((sum(sweeps) + sum(sweeps_x1)) / ms),
                "single_tf32_sweep_ms": float(np.mean(sweeps_x1)) if sweeps_x1 else None,
                "single_tf32_sweeps": len(sweeps_x1),
                "tiles": "64 features per tile from the packed copy of the loadings (one bulk copy per tile)"}
    # parity inside the run: rotation conserves the explained variance (tests/models/single/test_eof_rotator.py:98-137)
    ev_rot = float(r.explained_variance().values.sum())
    ev_eof = float(model.explained_variance().values[:k].sum())
    Rm = r.rotation_matrix()
    orth = float(np.max(np.abs(Rm.T @ Rm - np.eye(k))))
    parity = {"variance_conserved": {"rel_err": abs(ev_rot / ev_eof - 1), "rtol": 1e-5,
                                     "ok": bool(abs(ev_rot / ev_eof - 1) < 1e-5)},
              "rotation_orthogonal": {"max_abs_err": orth, "atol": 1e-8, "ok": bool(orth < 1e-8)},
              "iterations": iters, "iterations_tensor_core": int(getattr(r, "n_iter_tc_", 0)),
              "iterations_single_tf32": int(getattr(r, "n_iter_x1_", 0))}
    parity["ok"] = bool(parity["variance_conserved"]["ok"] and parity["rotation_orthogonal"]["ok"])
    want_cpu, cpu = cpu, None
    if want_cpu:
        from threadpoolctl import threadpool_limits

        from oracle import rotation as orot
        Sc = 131072
        Lh = np.random.default_rng(0).standard_normal((Sc, k)) * (np.random.default_rng(1).random((Sc, k)) < 0.05)
        Lh = Lh @ np.linalg.qr(np.random.default_rng(2).standard_normal((k, k)))[0]
        t0 = time.perf_counter()
        with threadpool_limits(limits=os.cpu_count() or 1):
            try:
                orot.varimax(Lh, max_iter=20, rtol=1e-30)
            except RuntimeError:
                pass
        dt = (time.perf_counter() - t0) / 20
        cpu = {"value": Sc * k * 4 / dt / 1e9, "unit": "GB/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"20 iterations of the oracle varimax (fp64, two S x m GEMMs + m x m SVD per iteration) on "
                         f"{Sc} x {k} loadings: {dt * 1e3:.0f} ms per iteration"}
    line = {
        "metric": "EOFRotator.fit GB/s (fp32 loadings streamed per varimax iteration)",
        "value": alg * iters / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": 1, "steps": max(2, args.steps // 2),
        "warmup": max(1, args.warmup - 2), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32x3/f64", "data": "synthetic",
        "config": {"workload": f"c5: EOFRotator varimax power=1 max_iter=1000 on {k} modes of a {T}x({n_lat}x{n_lon}) EOF "
                               f"model (loadings {alg / 1e9:.2f} GB), 1 GPU",
                   "l2": "loadings larger than L2 (no flush needed)"},
        "iterations": iters, "e2e": None, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
        "clocks": clocks, "parity": parity,
    }
    del r, model
    torch.cuda.empty_cache()
    return line, bool(parity["ok"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c3", "c5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong-c4", action="store_true")
    ap.add_argument("--no-models", action="store_true", help="skip the compact config-3 / config-5 objects of the line")
    ap.add_argument("--strong-workload", default="c4", choices=["c4", "c4mid"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0, help="c3 / c5: shrink the latitude axis (quick looks)")
    ap.add_argument("--ref-cols", type=int, default=CPU_SAMPLE_COLS, help="--impl reference: columns of the sample")
    ap.add_argument("--land-frac", type=float, default=0.0, help="fraction of grid points that are NaN at every step")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: warmup {args.warmup} < 3", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
        return 0
    if args.workload == "c3":
        return run_mca(args)
    if args.workload == "c5":
        return run_rotator(args)
    return run_eof(args)


if __name__ == "__main__":
    sys.exit(main())
