#!/usr/bin/env python
"""Jacobi sweeps per varimax iteration of config 5 (read from the solver's info word after every m x m step), and the
time of `sym_eig` on nearly diagonal 100 x 100 matrices.  Diagnosis only (one host sync per iteration)."""
import os
import sys
import collections

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import profile_fit  # noqa: E402
import xeofs_b200 as xb  # noqa: E402
from xeofs_b200._cuda_ops import CudaOps  # noqa: E402


def main():
    ops = CudaOps()
    m = 100
    g = torch.Generator(device="cuda").manual_seed(1)
    d = torch.sort(torch.rand(m, generator=g, device="cuda", dtype=torch.float64) + 0.5, descending=True).values
    E = torch.randn((m, m), generator=g, device="cuda", dtype=torch.float64)
    E = (E + E.t()) / 2
    for eps in (1e-1, 1e-3, 1e-5, 1e-7, 1e-10):
        M = (torch.diag(d) + eps * E).contiguous()
        ops.sym_eig(M)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.sym_eig(M)
        e1.record()
        torch.cuda.synchronize()
        print(f"sym_eig 100x100, off-diagonal {eps:.0e}: {e0.elapsed_time(e1) / 10 * 1e3:.0f} us", flush=True)

    fit = profile_fit.make_fit("c5")
    sweeps = []
    orig = CudaOps.varimax_update

    def patched(self, G3, W, XtX, alpha, R, basis, dsum, eig_tol=0.0):
        out = orig(self, G3, W, XtX, alpha, R, basis, dsum, eig_tol=eig_tol)
        mm = int(R.shape[0])
        off = (9 * mm * mm + mm + (mm + 2) * (mm + 2)) * 8
        sweeps.append((int(self._uws[off:off + 4].view(torch.int32).item()), eig_tol))
        return out

    CudaOps.varimax_update = patched
    r = fit()
    CudaOps.varimax_update = orig
    print("iterations", r.n_iter_, "x1", r.n_iter_x1_, "tc", r.n_iter_tc_)
    print("sweeps per iteration:", [s for s, _ in sweeps])
    print("by tolerance:", {t: collections.Counter(s for s, tt in sweeps if tt == t) for t in sorted({t for _, t in sweeps})})


if __name__ == "__main__":
    main()
