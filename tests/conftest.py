import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
DATA_DIR = os.path.join(GOLDEN_DIR, "data")
# (iid_count, sid_count) of the fixture .bed files copied from the reference (tests/golden/make_golden.py)
SHAPES = {"n300": (300, 1015), "toydata": (500, 10000), "dbx": (100, 100), "snpgen": (1000, 5), "gen1": (190, 20), "gen4": (198, 20)}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the shared objects are build artefacts (git-ignored): build them when a fresh checkout has none (nvcc cross-compiles without a GPU)
    import subprocess
    lib = os.path.join(ROOT, "pysnptools_b200", "libpst_b200.so")
    if not os.path.exists(lib):
        subprocess.call(["bash", os.path.join(ROOT, "pysnptools_b200", "csrc", "build.sh")], stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libpst_oracle.so")):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "golden.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import bed_oracle
    return bed_oracle


def i8_to_float(a, dtype=np.float64):
    out = a.astype(dtype)
    out[a == -127] = np.nan
    return out


def fixture_packed(name):
    from oracle import bed_oracle
    n, m = SHAPES[name]
    return bed_oracle.read_packed(os.path.join(DATA_DIR, name + ".bed"), n, m), n, m
