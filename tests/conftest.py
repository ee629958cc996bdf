import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def ref_path(cap=8192):
    return os.path.join(ROOT, "oracle", "_ref", "libferox_ref_%d.so" % cap)


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref, compiled from /root/reference by oracle/Makefile)."""
    from ferox_b200 import capi
    p = ref_path(8192)
    if not os.path.exists(p):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return capi.load_abi(p)


@pytest.fixture(scope="session")
def prod():
    """The product library; loading it needs no GPU, stepping does.  A fresh checkout has no built artefacts
    (they are git-ignored): compile them first, exactly as __graft_entry__.build() does."""
    import ferox_b200 as fb
    if not os.path.exists(fb.LIB_PATH):
        fb.build()
    return fb.lib()


@pytest.fixture(scope="session")
def gpu(prod):
    assert prod.frxSetDevice(0) == 0, "no CUDA device: -m gpu tests need one (there is no CPU fallback)"
    return prod
