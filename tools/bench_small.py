#!/usr/bin/env python
"""Time the k-column kernels at config-2 size.  usage: python tools/bench_small.py [S l]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xeofs_b200 import _lib
from xeofs_b200._cuda_ops import CudaOps
S, l = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1038240, 60)
T = 8760
ops = CudaOps()
lp = _lib.lpad(l)
Y = ops.space_side(lp, S, zero=True); Y[:l] = torch.randn((l, S), device="cuda")
Z = torch.zeros((T, lp), device="cuda"); Z[:, :l] = torch.randn((T, l), device="cuda")
def timeit(fn, n=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
G = ops.gram(Y, S, l, 1)
Rinv, _ = ops.chol_inv(G)
print(f"gram S-side   {timeit(lambda: ops.gram(Y, S, l, 1)):.3f} ms")
print(f"gram T-side   {timeit(lambda: ops.gram(Z, T, l, 0)):.3f} ms")
print(f"chol_inv      {timeit(lambda: ops.chol_inv(G)):.3f} ms")
print(f"apply S-side  {timeit(lambda: ops.apply(Y, S, l, 1, Rinv, l)):.3f} ms")
print(f"apply T-side  {timeit(lambda: ops.apply(Z, T, l, 0, Rinv, l)):.3f} ms")
for thr in (1024,):
    for n in (20, 60, 100, 110, 128):
        A = torch.randn((n, n), device="cuda", dtype=torch.float64)
        A = (A @ A.t()).contiguous()
        ms = timeit(lambda: ops.sym_eig(A))
        ev, V = ops.sym_eig(A)
        err = float(((V * ev[None, :]) @ V.t() - A).abs().max() / A.abs().max())
        D = torch.diag(ev) + 1e-3 * A / A.abs().max() * ev.mean()  # nearly diagonal: the warm-started case
        ms2 = timeit(lambda: ops.sym_eig(D.contiguous()))
        print(f"sym_eig threads={thr} n={n}: {ms:.3f} ms (err {err:.1e}); nearly diagonal: {ms2:.3f} ms")
print(f"row_minmax    {timeit(lambda: ops.row_minmax(Y, l, S)):.3f} ms")
