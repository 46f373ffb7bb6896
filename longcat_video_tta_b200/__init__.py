"""B200-native test-time-adaptation step for LongCat-Video (drop-in for the TTA inner step of
FifthEpoch/longcat-video-tta).  Python host over the C ABI in ``include/b200tta.h``; every hot kernel is
hand-written CUDA for sm_100a in ``csrc/``.  There is no CPU fallback: compute entry points raise on any
device that is not a B200."""
__version__ = "0.1.0"
