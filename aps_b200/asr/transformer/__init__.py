from .encoder import TransformerEncoder  # noqa: F401
