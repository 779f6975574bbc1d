import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


_REPORT = {}


@pytest.fixture
def report(request):
    """Collect error statistics per test; dumped to gpurun_out/test_report.json at session end."""
    d = {}
    _REPORT[request.node.nodeid] = d
    return d


def pytest_sessionfinish(session, exitstatus):
    if _REPORT:
        out = os.path.join(ROOT, "gpurun_out")
        try:
            os.makedirs(out, exist_ok=True)
            path = os.path.join(out, "test_report.json")
            old = {}
            if os.path.exists(path):
                try:
                    old = json.load(open(path))
                except Exception:
                    old = {}
            old.update(_REPORT)
            json.dump(old, open(path, "w"), indent=1, default=str)
        except OSError:
            pass
