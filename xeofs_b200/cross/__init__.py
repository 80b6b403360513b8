from .mca import CCA, CPCCA, MCA, RDA  # noqa: F401
from .mca_rotator import MCARotator  # noqa: F401
