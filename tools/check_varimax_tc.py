"""Developer check of the tcgen05 varimax sweep against fp64 torch, over the descriptor variants the kernel can be
switched between by environment (XEOFS_VT_STAGES), plus timings."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xeofs_b200._cuda_ops import CudaOps  # noqa: E402
from xeofs_b200._lib import lpad  # noqa: E402


def case(ops, S, m, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    Ln = ops.space_side(lpad(m), S, zero=True)
    Ln[:m] = torch.randn((m, S), generator=g, device="cuda") * torch.rand((m, 1), generator=g, device="cuda")
    Ln[:m] /= Ln[:m].double().norm(dim=0).float()[None, :]
    R = torch.linalg.qr(torch.randn((m, m), generator=g, device="cuda", dtype=torch.float64))[0].contiguous()
    X = Ln[:m].double().t()
    B = X @ R
    return Ln, R, X.t() @ B**3, (B * B).sum(0)


def run(ops, S, m, env):
    for k in ("XEOFS_VT_STAGES",):
        os.environ.pop(k, None)
    os.environ.update(env)
    Ln, R, Gref, Wref = case(ops, S, m)
    try:
        G, W, _ = ops.varimax_accumulate(Ln, S, m, R)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        return f"ERROR {e}"
    eg = float((G - Gref).abs().max() / Gref.abs().max())
    ew = float((W - Wref).abs().max() / Wref.abs().max())
    return f"relerr G {eg:.3e}  W {ew:.3e}"


def timeit(ops, S, m, env, n=10):
    for k in ("XEOFS_VT_STAGES",):
        os.environ.pop(k, None)
    os.environ.update(env)
    Ln, R, _, _ = case(ops, S, m)
    for _ in range(2):
        ops.varimax_accumulate(Ln, S, m, R)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        ops.varimax_accumulate(Ln, S, m, R)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if __name__ == "__main__":
    ops = CudaOps()
    ops.varimax_algo = "tc"
    variants = [{}]
    for env in variants:
        for (S, m) in ((64 * 3, 20), (5000, 20), (100037, 100), (30011, 128), (20000, 50), (4096, 8)):
            print(env, S, m, run(ops, S, m, env), flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        for env in ({}, {"XEOFS_VT_STAGES": "2"}):
            for (S, m) in ((4147200, 100), (1038240, 50), (4147200, 20)):
                ms = timeit(ops, S, m, env)
                print("time", env, S, m, f"{ms:.3f} ms/sweep  {S * lpad(m) * 4 / ms / 1e6:.0f} GB/s", flush=True)
        ops.varimax_algo = "simt"
        print("simt fp64", timeit(ops, 4147200, 100, {}, n=2), "ms/sweep")
