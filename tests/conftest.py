import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _build_once():
    """Make sure the in-tree libraries exist (compiles on CPU; no-op when up to date)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("agf_build", os.path.join(ROOT, "agri-fly_b200", "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.build_native()
    import orc
    if not (orc.available("port-glibc") and orc.available("port-shared")):
        m.build_oracle()


@pytest.fixture(scope="session", autouse=True)
def built():
    _build_once()


@pytest.fixture(scope="session")
def agf(built):
    import agrifly_b200
    agrifly_b200.lib()
    return agrifly_b200


@pytest.fixture(scope="session")
def orc_mod(built):
    import orc
    return orc


def oracle_or_skip(orc, flavour):
    if not orc.available(flavour):
        pytest.skip("oracle flavour %s not built here" % flavour)
    return orc.Oracle(flavour)


@pytest.fixture(scope="session")
def port_shared(orc_mod):
    return oracle_or_skip(orc_mod, "port-shared")


@pytest.fixture(scope="session")
def port_glibc(orc_mod):
    return oracle_or_skip(orc_mod, "port-glibc")


# The checker of the GPU parity tests: the UNMODIFIED reference sources (oracle/_ref, built where /root/reference exists and
# shipped to the GPU box) when present, else the port (bit-identical to it: tests/test_oracle.py).  AGF_TEST_ORACLE=port forces
# the port.
def _checker(orc, flavour):
    import os
    if os.environ.get("AGF_TEST_ORACLE", "") != "port" and orc.available("ref-" + flavour):
        return orc.Oracle("ref-" + flavour)
    return oracle_or_skip(orc, "port-" + flavour)


@pytest.fixture(scope="session")
def checker_shared(orc_mod):
    return _checker(orc_mod, "shared")


@pytest.fixture(scope="session")
def checker_glibc(orc_mod):
    return _checker(orc_mod, "glibc")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
