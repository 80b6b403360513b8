#!/usr/bin/env python
"""Time of one tensor-core varimax sweep at config 5's size (4 147 200 features x 100 modes), three-product and
single-product mode, with tiles of 64 features from the packed copy and of 32 through tensor maps.
usage: python tools/bench_varimax.py [S] [m]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xeofs_b200._cuda_ops import CudaOps  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 1440 * 2880
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    ops = CudaOps()
    ops.varimax_algo = "tc"
    g = torch.Generator(device="cuda").manual_seed(1)
    Ln = ops.space_side((m + 15) // 16 * 16, S, zero=True)
    Ln[:m] = torch.randn((m, S), generator=g, device="cuda") / S**0.5
    packed = ops.varimax_pack(Ln, S, m)
    R = torch.linalg.qr(torch.randn((m, m), generator=g, device="cuda", dtype=torch.float64))[0].contiguous()
    for pair, pf in [("1", "2"), ("1", "0"), ("1", "4"), ("0", "0")]:
        os.environ["XEOFS_VT_PF"] = pf
        print(f"-- L2 prefetch distance {pf}")
        for products in (3, 1):
            for _ in range(3):
                ops.varimax_accumulate(Ln, S, m, R, products=products, packed=packed if pair == "1" else None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.varimax_accumulate(Ln, S, m, R, products=products, packed=packed if pair == "1" else None)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            tf = 4.0 * S * m * m * products / ms / 1e9
            print(f"S={S} m={m} tile={'64' if pair == '1' else '32'} products={products}: {ms:.3f} ms/sweep "
                  f"({tf:.0f} TFLOP/s tf32, {S * m * 4 / ms / 1e6:.0f} GB/s)", flush=True)


def exact():
    """the fp64 sweep (the last iterations of a rotation), mma.sync kernel and the CUDA-core one"""
    S, m = 1440 * 2880, 100
    ops = CudaOps()
    ops.varimax_algo = "simt"
    g = torch.Generator(device="cuda").manual_seed(1)
    Ln = ops.space_side((m + 15) // 16 * 16, S, zero=True)
    Ln[:m] = torch.randn((m, S), generator=g, device="cuda") / S**0.5
    R = torch.linalg.qr(torch.randn((m, m), generator=g, device="cuda", dtype=torch.float64))[0].contiguous()
    for mma in ("1", "0"):
        os.environ["XEOFS_VX_MMA"] = mma
        ops.varimax_accumulate(Ln, S, m, R, exact=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            G, W, _ = ops.varimax_accumulate(Ln, S, m, R, exact=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"fp64 sweep, mma.sync={mma}: {ms:.2f} ms ({4.0 * S * m * m / ms / 1e9:.1f} TFLOP/s fp64)", flush=True)
    del os.environ["XEOFS_VX_MMA"]


if __name__ == "__main__":
    exact()
    main()
