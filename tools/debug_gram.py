"""Developer probe: signed error of the sample Gram block A A[t0:t1]^T through project_T per algo."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xeofs_b200 import _lib
from xeofs_b200._cuda_ops import CudaOps, Field

ops = CudaOps()
T, S = 300, 70000
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn((T, S), generator=g, device="cuda") * 3 + 280
st = ops.col_stats(X)
fin = ops.scaling_finalize(st, None, True, False)
f = Field(X, fin["pivot"], fin["dscale"], None, fin["valid"])
A = (X.double() - X.double().mean(0))
Gref = A @ A[:128].t()
blk = ops.scaled_rows(f, 0, 128)
print("blk vs A:", float((blk[:128].double() - A[:128]).abs().max()))
for name in ("tf32x1", "tf32x1r", "tf32x3", "simt"):
    G = ops.project_T(f, blk, 128, algo=_lib.ALGO_NAMES[name])[:, :128].double()
    d = torch.diagonal(G[:128]) / torch.diagonal(Gref[:128]) - 1
    off = (G - Gref)
    print(f"{name:8s} diag rel err mean {float(d.mean()):+.3e} std {float(d.std()):.3e};  all entries: mean signed err / mean|G| "
          f"{float(off.mean() / Gref.abs().mean()):+.3e}, rms err / rms G {float(off.pow(2).mean().sqrt() / Gref.pow(2).mean().sqrt()):.3e}")
