import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: a minute or more of CPU work in the oracle")


@pytest.fixture(scope="session")
def dummy_goldens():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "dummy_goldens.json")))


@pytest.fixture(scope="session")
def dummy_inputs(dummy_goldens, tmp_path_factory):
    """Materialise the reference's tst/dummy inputs (stored inside the fixture JSON) as files."""
    d = tmp_path_factory.mktemp("dummy")
    for name, txt in dummy_goldens["inputs"].items():
        (d / name).write_text(txt)
    return d
