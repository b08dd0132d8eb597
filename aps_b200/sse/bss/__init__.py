from .tcn import FreqConvTasNet  # noqa: F401
