"""jperceiver_b200 — B200-native (sm_100a) training hot path of JPerceiver.

Host side: Python mirror of the reference's ``mono.model`` / ``mono.apis`` surface
(``jperceiver_b200.model``, ``jperceiver_b200.apis``; also importable as ``mono.*``).
Device side: hand-written CUDA in ``csrc/`` behind the C ABI of ``include/jpb200.h``,
reached through ctypes (``_lib``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
