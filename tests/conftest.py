import os
import sys
from pathlib import Path

import pytest

# several VMs share one GPU in the in-process multi-rank test (one spins on a flag another one's kernels set): give every
# stream its own hardware queue so that no kernel is ever queued behind a waiting one (must be set before CUDA starts)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import subprocess
    from dacapo_b200 import _binding
    subprocess.run(["make", "-s", "-C", str(REPO / "oracle")], check=True)
    import fixtures
    return _binding.bind(fixtures.ORACLE_LIB)


@pytest.fixture(scope="session")
def b200_lib():
    import os
    from dacapo_b200 import _binding
    return _binding.bind(os.environ.get("HEVM_LIB", _binding.B200_LIB))  # HEVM_LIB: A/B builds of the same library
