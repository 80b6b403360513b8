"""xeofs_b200 — the EOF / MCA / EOFRotator fit path of xeofs on NVIDIA B200 (sm_100a).

``xeofs_b200.single.EOF``, ``xeofs_b200.cross.MCA`` and ``xeofs_b200.single.EOFRotator`` mirror the
reference classes' constructor and ``fit`` signatures; the arithmetic runs in hand-written CUDA kernels
(``libxeofs_b200.so``, C-ABI in ``include/xeofs_b200.h``).  There is no CPU fallback.
"""
from . import cross, single, validation  # noqa: F401
from ._labels import DataArray  # noqa: F401

__version__ = "0.1.0"
