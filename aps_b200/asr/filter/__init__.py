from .mvdr import MvdrBeamformer, beamform, estimate_covar  # noqa: F401
