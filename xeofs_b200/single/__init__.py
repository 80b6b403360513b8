from .eof import EOF  # noqa: F401
from .eof_rotator import EOFRotator  # noqa: F401
