from .objf import pair_objf_matrix, sisnr_objf, snr_objf, permu_invarint_objf, multiple_objf, hybrid_permu_objf
from .sse import (SisnrTask, SnrTask, WaTask, LinearFreqSaTask, MelFreqSaTask, LinearTimeSaTask, MelTimeSaTask,
                  ComplexMappingTask, ComplexMaskingTask)

__all__ = ["pair_objf_matrix", "sisnr_objf", "snr_objf", "permu_invarint_objf", "multiple_objf", "hybrid_permu_objf",
           "SisnrTask", "SnrTask", "WaTask", "LinearFreqSaTask", "MelFreqSaTask", "LinearTimeSaTask", "MelTimeSaTask",
           "ComplexMappingTask", "ComplexMaskingTask"]
