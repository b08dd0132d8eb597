from .dccrn import DCCRN  # noqa: F401
from .tcn import FreqConvTasNet  # noqa: F401
