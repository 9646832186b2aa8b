import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_json_gz(rel):
    with gzip.open(os.path.join(GOLDEN, rel), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def config_case(name):
    """One of the BASELINE.json configurations (tests/golden/configs/*.json.gz) as a dict."""
    return load_json_gz(os.path.join("configs", name + ".json.gz"))


CONFIG_NAMES = ["c1_example", "c2_case3", "c3_case7", "c3_case7_evolving", "c4_trappist1", "c5_circumbinary"]
