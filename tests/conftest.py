import os
import subprocess
import sys

# The reference's OpenMP build has a row-end store race; one thread keeps it deterministic.
os.environ.setdefault("OMP_NUM_THREADS", "1")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _make(*targets):
    r = subprocess.run(["make", "-C", ROOT, "--no-print-directory", *targets], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("make %s failed:\n%s\n%s" % (" ".join(targets), r.stdout[-4000:], r.stderr[-4000:]))


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Everything is built in-tree by __graft_entry__.build(); make is a no-op when up to date."""
    have_nvcc = os.path.exists("/usr/local/cuda/bin/nvcc")
    _make("oracle")
    if have_nvcc:
        _make("lib", "scenes")
    if os.path.isdir("/root/reference/src"):
        _make("ref")
    return True


@pytest.fixture(scope="session")
def oracle_scenes(built_libraries):
    from checkers import load_oracle_scenes
    return load_oracle_scenes()


def _ref(bfix):
    from checkers import load_reference_scenes
    try:
        return load_reference_scenes(bfix)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (no /root/reference in this environment)")


@pytest.fixture(scope="session")
def ref_scenes(built_libraries):
    return _ref(False)


@pytest.fixture(scope="session")
def ref_bfix_scenes(built_libraries):
    return _ref(True)


@pytest.fixture(scope="session")
def product_scenes(built_libraries):
    from pixelforge_b200 import load_product_scenes
    return load_product_scenes()


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "scenes_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def host_matches_golden(golden, built_libraries):
    """Golden hashes were produced on a host with a particular RCPPS/RSQRTPS table + libm."""
    import hashlib
    import numpy as np
    from checkers import load_oracle_pfcu
    lib = load_oracle_pfcu()
    rcp, rb, rsq, sb = lib.harvest_tables()
    a = np.ctypeslib.as_array(rcp, shape=(1 << rb,)).tobytes() + np.ctypeslib.as_array(rsq, shape=(2 << sb,)).tobytes()
    return hashlib.sha256(a).hexdigest() == golden["host_tables"]["sha256"]
