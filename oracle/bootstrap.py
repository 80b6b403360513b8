"""numpy restatement of xeofs.validation.EOFBootstrapper.fit (validation/bootstrapper.py:56-135).
TEST INFRASTRUCTURE ONLY — imported by tests/, never by the product.

The reference fits every member with an UNSEEDED randomized SVD (bootstrapper.py:89: ``EOF(n_modes=...)`` without
random_state); ``random_state`` here makes the restatement reproducible and is the only departure."""
import numpy as np

from . import eof as oeof


def eof_bootstrap(A, model_scores, n_modes, n_bootstraps=20, seed=None, random_state=None, solver_kwargs=None):
    """A: (n, S') the fitted model's preprocessed input_data (valid samples / features), model_scores (n, k)."""
    n = A.shape[0]
    rng = np.random.default_rng(seed)                                   # :72
    dims = ("sample", "feature")
    expvar, totvar, comps, scores = [], [], [], []
    for _ in range(n_bootstraps):
        idx = rng.choice(n, n, replace=True)                            # :81
        m = oeof.eof_fit(A[idx], dims, "sample", n_modes=n_modes, standardize=False, use_coslat=False,
                         random_state=random_state, solver_kwargs=solver_kwargs)          # :89-90
        V = m["components_2d"]
        mean = A[idx].mean(axis=0)
        sc = (A - mean) @ V                                             # :95 transform(input_data)
        expvar.append(m["explained_variance"])
        totvar.append(m["total_variance"])
        comps.append(V)
        scores.append(sc)
    expvar, totvar, comps, scores = map(np.array, (expvar, totvar, comps, scores))
    ms = model_scores[:, :n_modes]
    corr = (scores * ms).mean(axis=1) / scores.std(axis=1) / ms.std(axis=0)          # :117-121
    signs = np.sign(corr)                                               # (n_boot, k)
    return {
        "explained_variance": expvar, "total_variance": totvar,
        "components": comps * signs[:, None, :], "scores": scores * signs[:, None, :],
    }
