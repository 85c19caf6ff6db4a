import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def chr1_fixture():
    """The reference's 20k-line 1000G chr1 slice (previous_out_check/), decompressed."""
    import gzip

    with gzip.open(os.path.join(ROOT, "tests", "golden", "chr1_20klines.vcf.gz")) as f:
        return f.read()


@pytest.fixture(scope="session")
def query_fixture():
    """The reference's examples/test.query.vcf (60 samples, GT:AD:DP:GQ:PL, '/' separated)."""
    import gzip

    with gzip.open(os.path.join(ROOT, "tests", "golden", "test.query.vcf.gz")) as f:
        return f.read()
