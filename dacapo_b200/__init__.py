"""dacapo_b200 -- B200-native HEVM execution backend (drop-in for libSEAL_HEVM.so).

The package holds only what the hot path needs: `csrc/` (sm_100a CUDA kernels + the
C-ABI shared library libB200_HEVM.so) and the host-side mirror of the reference's
Python driver (`runner.HEVM`), an HEVM/CST assembler and the cost-profile emitter.
"""
from .runner import HEVM, setLibnHW  # noqa: F401
