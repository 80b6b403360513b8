"""Seeded input generators shared by make_golden.py and the tests (numpy only)."""
import numpy as np


def mock_data_array():
    """The reference's canonical fixture, /root/reference/tests/conftest.py:225-240 (seed 7)."""
    rng = np.random.default_rng(7)
    noise = rng.normal(5, 3, size=(25, 5, 4))
    signal = 2 * np.sin(np.linspace(0, 2 * np.pi, 25))[:, None, None]
    return signal + noise


MOCK_LAT = np.array([20.0, 30.0, 40.0, 50.0, 60.0])
MOCK_LON = np.array([-10.0, 0.0, 10.0, 20.0])


def planted(T, S, r, seed, sigma0=1000.0, decay=0.8, eps=0.5, offset=280.0, dtype=np.float32):
    """Planted-spectrum field: offset + sum_i sigma_i u_i v_i^T + eps*N(0,1)."""
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((T, r)))
    V, _ = np.linalg.qr(rng.standard_normal((S, r)))
    sig = sigma0 * decay ** np.arange(r)
    return (offset + (U * sig) @ V.T + eps * rng.standard_normal((T, S))).astype(dtype)
