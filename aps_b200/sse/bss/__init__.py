from .dccrn import DCCRN  # noqa: F401
from .tcn import FreqConvTasNet, TimeConvTasNet  # noqa: F401
from .transformer import FreqXfmr  # noqa: F401
