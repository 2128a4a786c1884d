"""gptorch_b200 -- B200-native (sm_100a) implementation of gptorch's dense Gaussian-process hot path.

Same module layout and call surface as the reference package (``gptorch``): kernels, functions, models, ...
All numerics run in libgpb200.so (hand-written CUDA behind a C ABI, include/gpb200.h); there is no CPU path.
"""
__version__ = "0.1.0"
