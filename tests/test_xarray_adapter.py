"""The labelled-array adapter (xeofs_b200/_labels.py) with xarray-typed inputs.  xarray itself is not installed in this
image, so the test plants a minimal stand-in module named ``xarray`` whose DataArray has the attributes the adapter
reads from the real class (``dims``, ``coords`` mapping to objects with ``.values`` / ``.ndim``, ``values``) and the
constructor signature it calls (``DataArray(data, dims=, coords=, name=, attrs=)``).  What is checked is the contract of
the reference's accessors (single/base_model_single_set.py:307-336): xarray in -> xarray out, feature dims + 'mode' for
components, sample dims + 'mode' for scores, coordinates carried over."""
import sys
import types

import numpy as np
import pytest

from _inputs import MOCK_LAT, MOCK_LON, mock_data_array
from cpu_ops import TorchCpuOps
from oracle import eof as oeof


class _Coord:
    def __init__(self, v):
        self.values = np.asarray(v)
        self.ndim = self.values.ndim


class _StubDataArray:
    def __init__(self, data, dims=None, coords=None, name=None, attrs=None):
        self.values = np.asarray(data)
        self.dims = tuple(dims)
        self.coords = {k: _Coord(v) for k, v in (coords or {}).items()}
        self.name, self.attrs = name, dict(attrs or {})


@pytest.fixture
def fake_xarray(monkeypatch):
    mod = types.ModuleType("xarray")
    _StubDataArray.__module__ = "xarray.core.dataarray"
    mod.DataArray = _StubDataArray
    monkeypatch.setitem(sys.modules, "xarray", mod)
    yield mod
    _StubDataArray.__module__ = __name__


def test_xarray_in_xarray_out(fake_xarray):
    import xeofs_b200 as xb
    X = mock_data_array().astype(np.float32)
    time = np.arange(25)
    da = fake_xarray.DataArray(X, dims=("time", "lat", "lon"), coords={"time": time, "lat": MOCK_LAT, "lon": MOCK_LON})
    m = xb.single.EOF(n_modes=3, use_coslat=True, random_state=5, ops=TorchCpuOps()).fit(da, dim="time")
    comps, scores = m.components(), m.scores()
    assert isinstance(comps, _StubDataArray) and isinstance(scores, _StubDataArray)
    assert comps.dims == ("lat", "lon", "mode") and scores.dims == ("time", "mode")
    np.testing.assert_array_equal(comps.coords["lat"].values, MOCK_LAT)
    np.testing.assert_array_equal(scores.coords["time"].values, time)
    np.testing.assert_array_equal(comps.coords["mode"].values, [1, 2, 3])
    o = oeof.eof_fit(X, ("time", "lat", "lon"), "time", coords={"lat": MOCK_LAT, "lon": MOCK_LON}, n_modes=3,
                     use_coslat=True, random_state=5)
    np.testing.assert_allclose(m.singular_values().values, o["singular_values"], rtol=1e-5)
    assert ((comps.values.reshape(-1, 3) * o["components_2d"]).sum(0) > 1 - 1e-5).all()
    # transform / inverse_transform keep the type and the input's dimension order
    rec = m.inverse_transform(m.transform(da))
    assert isinstance(rec, _StubDataArray) and rec.dims == ("time", "lat", "lon")
    # anything else is refused like the reference refuses it (utils/sanity_checks.py:85-95)
    with pytest.raises(TypeError, match="Invalid input type"):
        xb.single.EOF(n_modes=2, ops=TorchCpuOps()).fit(X, dim="time")
