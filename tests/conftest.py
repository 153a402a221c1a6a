import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def orc():
    import cpulibs
    return cpulibs.oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own code (oracle/_ref); tests that need it skip when it was never built."""
    import cpulibs
    r = cpulibs.reference()
    if r is None:
        pytest.skip("oracle/_ref/libbbcref.so not built (needs /root/reference, dev container only)")
    return r


@pytest.fixture(scope="session")
def bbx():
    """The product binding.  GPU tests must not pass on a fallback: no device -> hard failure."""
    import bbcat_dsp_b200 as b
    b.lib()
    if b.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need a B200 (libbbx has no CPU fallback)")
    return b


@pytest.fixture(scope="session")
def gpu(bbx):
    import gpulib
    return gpulib.GpuLib(bbx)


# one test body, two implementations: the CPU oracle (always) and the CUDA product (-m gpu)
IMPLS = ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.fixture(params=IMPLS)
def impl(request):
    if request.param == "oracle":
        return request.getfixturevalue("orc")
    return request.getfixturevalue("gpu")


def golden(name):
    import numpy as np
    return np.load(os.path.join(HERE, "golden", name))
