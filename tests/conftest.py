import csv
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_csv(name):
    rows = list(csv.reader(open(os.path.join(GOLDEN, name + ".csv"))))
    out = {}
    for j, c in enumerate(rows[0]):
        v = [r[j] for r in rows[1:]]
        try:
            out[c] = np.array([float(x) for x in v])
        except ValueError:
            out[c] = np.array(v, dtype=object)
    return out


@pytest.fixture(scope="session")
def thetalog2pop():
    return load_csv("thetalog2pop")


@pytest.fixture(scope="session")
def oracle_cases():
    return json.load(open(os.path.join(GOLDEN, "oracle_cases.json")))


def arr(v):
    return np.array([np.nan if x is None else x for x in v], dtype=float)


@pytest.fixture(scope="session")
def gpu_lib():
    import gpedm_b200
    L = gpedm_b200.lib()
    if L.gpedm_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (the product path has no CPU fallback)")
    return L
