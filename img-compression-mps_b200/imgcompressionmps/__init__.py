"""B200-native drop-in for the NDMPS encode / compress / reconstruct path of
Alandroid/img-compression-mps.  Same import paths and names as the reference
(``imgcompressionmps.core.ndmps.NDMPS``, ``imgcompressionmps.utils.core``,
``imgcompressionmps.utils.metrics``, ``imgcompressionmps.utils.filetools``); all
arithmetic runs in hand-written sm_100a CUDA behind ``libndmps_sm100.so``.
"""
__version__ = "0.1.0"
