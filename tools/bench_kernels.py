#!/usr/bin/env python
"""Time the streaming products alone (CUDA events, inputs larger than L2) and print GB/s of algorithmic bytes.
usage: python tools/bench_kernels.py [T S l] [--env NAME=v1,v2,...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xeofs_b200 import _lib  # noqa: E402
from xeofs_b200._cuda_ops import CudaOps, Field  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    T, S, l = (int(args[0]), int(args[1]), int(args[2])) if len(args) >= 3 else (8760, 1038240, 60)
    sweeps = [a[6:] for a in sys.argv[1:] if a.startswith("--env=")]
    ops = CudaOps(algo="simt")
    lp = _lib.lpad(l)
    X = torch.randn((T, S), device="cuda") * 3 + 280
    st = ops.col_stats(X)
    fin = ops.scaling_finalize(st, None, True, False)
    f = Field(X, fin["pivot"], fin["dscale"], None, fin["valid"], no_nan=True)
    W = torch.zeros((T, lp), device="cuda")
    W[:, :l] = torch.randn((T, l), device="cuda")
    Y = ops.space_side(lp, S, zero=True)
    Y[:l] = torch.randn((l, S), device="cuda")
    alg_bytes = T * S * 4 + S * lp * 4 + T * lp * 4

    def timeit(fn, n=5):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def run(tag):
        for name, algo in (("x1", _lib.ALGO_TF32X1), ("x1r", _lib.ALGO_TF32X1R), ("x2", _lib.ALGO_TF32X2), ("x3", _lib.ALGO_TF32X3)):
            ms_s = timeit(lambda: ops.project_S(f, W, l, algo=algo))
            ms_t = timeit(lambda: ops.project_T(f, Y, l, algo=algo))
            print(f"{tag:28s} {name}  project_S {ms_s:7.3f} ms {alg_bytes / ms_s / 1e6:7.0f} GB/s   "
                  f"project_T {ms_t:7.3f} ms {alg_bytes / ms_t / 1e6:7.0f} GB/s", flush=True)

    print(f"T={T} S={S} l={l} lp={lp}  {T * S * 4 / 1e9:.2f} GB")
    if not sweeps:
        run("default")
    for sw in sweeps:
        name, vals = sw.split("=")
        for v in vals.split(","):
            os.environ[name] = v
            run(f"{name}={v}")
        os.environ.pop(name, None)
    ms = timeit(lambda: ops.col_stats(X))
    print(f"col_stats {ms:7.3f} ms {T * S * 4 / ms / 1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
