import gzip
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_inputs(tmp_path_factory):
    """Decompress the committed FASTA/FASTQ fixtures; returns (names, paths) in golden order."""
    work = tmp_path_factory.mktemp("gold_in")
    names = open(os.path.join(GOLD, "inputs", "order.txt")).read().split()
    paths = []
    for n in names:
        dst = os.path.join(work, n)
        with gzip.open(os.path.join(GOLD, "inputs", n + ".gz"), "rb") as f, open(dst, "wb") as o:
            shutil.copyfileobj(f, o)
        paths.append(dst)
    return names, paths


def expected(name):
    return os.path.join(GOLD, "expected", name)
