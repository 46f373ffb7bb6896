import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "refonly: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = Path("/root/reference/delta_experiment/scripts/common.py").is_file()
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this box")
    for item in items:
        if "refonly" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
