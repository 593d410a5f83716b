import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def xml_dir():
    return os.path.join(ROOT, "tests", "golden", "xmls")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import pyoracle as po

    po.build()
    return po.Oracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import pyoracle as po

    if not po.Oracle.reference_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return po.Oracle("reference")
