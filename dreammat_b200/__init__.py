"""dreammat_b200 -- B200-native (sm_100a) score-distillation inner loop for DreamMat.

The package holds only what the hot path needs: `csrc/` (CUDA kernels + the C-ABI of
include/dreammat_b200.h), thin torch-facing wrappers, and host-side mirrors of the five
threestudio plugins the path goes through.
"""
__version__ = "0.1.0"
