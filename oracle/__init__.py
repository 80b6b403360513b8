"""CPU oracle for the xeofs EOF / MCA / EOFRotator hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``xeofs_b200/`` does.

It is a numpy restatement of the reference's arithmetic (xeofs v3.0.4,
``/root/reference``), function by function, each citing the reference
file:line it follows.  The reference itself cannot be imported in this image
(``xarray`` and ``dask`` are hard imports at ``xeofs/base_model.py:7-8`` and are
not installed), and it holds no golden vectors.  What pins it:

* the randomized SVD is not restated at all: the oracle calls the very function
  the reference calls, ``sklearn.utils.extmath.randomized_svd`` (third party,
  un-vendored, unpinned in the reference's pyproject.toml: ``scikit-learn>=1.0.2``;
  installed here: 1.9.0) — reference call site ``xeofs/linalg/decomposer.py:141-146``;
* ``rotation.py`` / ``mca.cross_covariance`` / ``preprocess.sqrt_cos_lat`` are
  checked against golden vectors produced by EXECUTING the reference's own source for
  ``_varimax``/``_promax`` (``xeofs/linalg/_numpy/_rotation.py``),
  ``_compute_cross_covariance_numpy`` (``xeofs/cross/cpcca.py:1008-1015``) and
  ``_np_sqrt_cos_lat_weights`` (``xeofs/utils/xarray_utils.py:256-270``) in this
  container (``tests/golden/make_golden.py`` — dask stubbed, it is only used for
  an isinstance check);
* the xarray-dependent glue (Scaler, Stacker, Sanitizer, sign rule, total variance,
  EOF/MCA/EOFRotator post-processing) cannot be executed without xarray: PARITY
  UNPINNED for those lines beyond the reference's own test invariants, which
  ``tests/test_oracle.py`` re-states (mean 0 / std 1 after scaling, total variance ==
  sum of np.var(ddof=1), full-rank reconstruction, transform == scores,
  rotation conserves explained variance, total squared covariance == sum |cov|^2).
"""
