from .mca import CCA, CPCCA, MCA, RDA  # noqa: F401
from .mca_rotator import CPCCARotator, MCARotator  # noqa: F401
