"""B200-native generator-inference path of EmotionGesture.

Public surface:
  Transformer            drop-in for Full_model.Models.Transformer (inference)
  MemoryTransformer      drop-in for Full_model.Models_memory.Transformer (Prior_MemoryEncoder variant)
  install(generator)     swap a live reference generator's forward for libegx
  Engine                 thin object wrapper over the C ABI (include/egx.h)
  GeneratorConfig, TED, BEAT
  fgd                    FGD statistics (mean/covariance) + NCCL all-reduce
"""
from .config import (BEAT, LOGMEL_DB, LOGMEL_FP16_STORAGE, LOGMEL_LOG_IN, LOGMEL_REFERENCE, TED, GeneratorConfig,
                     audio_length, spectrogram_length)
from .generator import MemoryTransformer, Transformer, randomize_norm_stats_

__all__ = ["Transformer", "MemoryTransformer", "install", "Engine", "GeneratorConfig", "TED", "BEAT", "LOGMEL_DB",
           "LOGMEL_LOG_IN", "LOGMEL_FP16_STORAGE", "LOGMEL_REFERENCE", "audio_length", "spectrogram_length", "randomize_norm_stats_"]


def __getattr__(name):
    # engine / drop-in pull in ctypes + the shared library lazily
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "install":
        from .dropin import install
        return install
    raise AttributeError(name)
