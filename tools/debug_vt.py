"""Developer probe of the tcgen05 varimax sweep: prints the kernel's first-tile intermediates next to fp64 torch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xeofs_b200._cuda_ops import CudaOps  # noqa: E402
from xeofs_b200._lib import lpad  # noqa: E402

os.environ["XEOFS_VT_FLAGS"] = os.environ.get("XEOFS_VT_FLAGS", "4")
ops = CudaOps()
ops.varimax_algo = "tc"
S, m = 64, 8
g = torch.Generator(device="cuda").manual_seed(0)
Ln = ops.space_side(lpad(m), S, zero=True)
Ln[:m] = torch.randn((m, S), generator=g, device="cuda")
R = torch.linalg.qr(torch.randn((m, m), generator=g, device="cuda", dtype=torch.float64))[0].contiguous()
X = Ln[:m].double().t()
B = X @ R
print("expected D1[j'][s=0..3]:\n", B.t()[:3, :4])
print("tile row0/1 first values:", Ln[0, :2].tolist(), Ln[1, :2].tolist(), " R[0][0], R[1][0]:", R[0, 0].item(), R[1, 0].item())
print("expected G'[j'][i=0..3]:\n", (X.t() @ B**3).t()[:3, :4])
G, W, _ = ops.varimax_accumulate(Ln, S, m, R)
torch.cuda.synchronize()
print("G\n", G[:3, :4], "\nW", W, "\nWref", (B * B).sum(0))
ws = ops._vws
f = ws[:65536].view(torch.float32)
print("rhi nonzero:", int((f != 0).sum()), "first", f[:8].tolist())
gp = ws[131072:131072 + 128 * 32 * 8].view(torch.float64).view(128, 32)
print("gpart[0][:3,:4]\n", gp[:3, :4])
