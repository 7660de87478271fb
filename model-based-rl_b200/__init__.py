"""B200-native MuZero search-and-target engine (drop-in for the hot path of JimOhman/model-based-rl).

Import as `model_based_rl_b200`.  The CUDA library (csrc/ -> libmzb200.so) is loaded lazily by
`_lib.load()`; nothing here falls back to a CPU implementation.
"""
__version__ = "0.1.0"
