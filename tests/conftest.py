import os
import sys

import pytest

# The torch-CPU oracle is the checker of the fp32 parity bar (1e-5): make it run-to-run reproducible.  MKL picks code paths by
# buffer alignment unless conditional numerical reproducibility is requested, and multi-threaded reductions change with the
# thread count -- both showed up as a 2e-5 .. 5e-5 wobble of the oracle's FIRST evaluation in about 1 of 30 processes
# (DESIGN.md section 6).  Must be set before MKL initialises, i.e. before the first `import torch`.
os.environ.setdefault("MKL_CBWR", "AUTO")
ORACLE_THREADS_TIGHT = 1   # fp32-parity tests: fixed summation order
ORACLE_THREADS_BULK = int(os.environ.get("CEEUS_ORACLE_THREADS", "0")) or min(16, os.cpu_count() or 1)  # 1e-3-bar tests at full size

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "bulk_oracle: the CPU oracle may use all host threads (tolerance 1e-3 cases at full size)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _oracle_threads(request):
    """One oracle thread unless the test asks for `bulk_oracle` (full-size tensor-core cases, tolerance 1e-3)."""
    import torch

    torch.set_num_threads(ORACLE_THREADS_BULK if "bulk_oracle" in request.keywords else ORACLE_THREADS_TIGHT)
    yield
