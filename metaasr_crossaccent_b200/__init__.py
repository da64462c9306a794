"""metaasr_crossaccent_b200 -- B200-native meta-training hot path of MetaASR-CrossAccent.

Python host (thin) over the C ABI in include/metaasr_b200.h.  Importing the package does not load
the CUDA library; constructing a CudaBackend / trainer does, and fails loudly if it is missing.
"""
__version__ = "0.1.0"
