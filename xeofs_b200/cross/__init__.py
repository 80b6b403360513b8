from .mca import MCA  # noqa: F401
from .mca_rotator import MCARotator  # noqa: F401
