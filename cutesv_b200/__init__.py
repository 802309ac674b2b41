"""cutesv_b200 -- B200-native hot path of cuteSV (signature extraction -> clustering -> genotype).

The compute lives in libcutesv_b200.so (hand-written sm_100a CUDA behind the C-ABI of
include/cutesv_b200.h).  Importing the package does not need a GPU; creating an Engine does,
and fails loudly otherwise -- there is no CPU fallback.
"""
__version__ = "0.1.0"
