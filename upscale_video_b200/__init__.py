"""B200-native drop-in for the frame-upscale worker loop of davlee1972/upscale_video.

Public surface mirrors ``upscale/upscale_processing.py`` of the reference (see ``upscale_processing``
in this package); the arithmetic runs in ``csrc/`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/b2sr.h``).
"""
__version__ = "0.1.0"
