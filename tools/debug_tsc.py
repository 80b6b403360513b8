"""Developer probe: MCA total squared covariance per algo against fp64 torch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import xeofs_b200 as xb
from xeofs_b200 import _lib
from test_gpu_models import _coupled_fields, DIMS

T, S1, S2, k = 300, 70000, 66000, 4
X, Y = _coupled_fields(T, S1, S2, 2 * k, seed=11)
X = X.reshape(T, 70, 1000); Y = Y.reshape(T, 66, 1000)
A1 = torch.from_numpy(X.reshape(T, -1)).cuda().double(); A2 = torch.from_numpy(Y.reshape(T, -1)).cuda().double()
A1 -= A1.mean(0); A2 -= A2.mean(0)
G1, G2 = A1 @ A1.t(), A2 @ A2.t()
ref = float((G1 * G2).sum()) / (T - 1) ** 2
print("ref", ref, " diag share", float((torch.diagonal(G1) * torch.diagonal(G2)).sum() / (G1 * G2).sum()))
m = xb.cross.MCA(n_modes=k, random_state=3, use_pca=False, total_squared_covariance=False)
m.fit(xb.DataArray(X, DIMS), xb.DataArray(Y, DIMS), dim="time")
for name in ("tf32x1r", "tf32x3", "tf32x1", "simt"):
    m.ops.sum_algo = _lib.ALGO_NAMES[name]
    m.data.pop("total_squared_covariance", None)
    v = m.total_squared_covariance()
    print(f"{name:8s} {v:.6f} rel {v / ref - 1:+.3e}")
# Gram block of field 1 per algo
from xeofs_b200._cuda_ops import Field
f = m._f1.field
blk = m.ops.scaled_rows(f, 0, 128)
for name in ("tf32x1r", "tf32x3", "simt"):
    G = m.ops.project_T(f, blk, 128, algo=_lib.ALGO_NAMES[name])[:, :128].double()
    off = G - G1[:, :128]
    print(f"{name:8s} gram: rms err / rms G {float(off.pow(2).mean().sqrt() / G1[:, :128].pow(2).mean().sqrt()):.3e}  "
          f"<err, G>/<G, G> {float((off * G1[:, :128]).sum() / (G1[:, :128] ** 2).sum()):+.3e}")
