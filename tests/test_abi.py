"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares."""
import ctypes
import os

import pytest

from fal_net_b200 import _lib


@pytest.fixture(scope="module")
def built():
    from fal_net_b200.build import build_library
    return build_library()


def test_library_builds_and_loads(built):
    assert os.path.exists(built)
    h = ctypes.CDLL(built)
    h.faln_version.restype = ctypes.c_int
    assert h.faln_version() >= 100


def test_exports_every_declared_symbol(built):
    h = ctypes.CDLL(built)
    names = _lib.declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, missing
    # and the Python binding covers the whole header
    assert sorted(_lib._SIGNATURES) == names


def test_argument_errors_are_reported_without_a_gpu(built):
    L = _lib.lib()
    rc = L.faln_med_fwd(None, None, None, None, None, None, None, None, None, None, None, 1, 49, 4, 64, 64, 0, None)
    assert rc == -1
    assert b"null" in L.faln_last_error()


def test_sass_uses_bulk_async_copy(built):
    """The MED kernels stream planes with the TMA unit (SASS UBLKCP), not with LDG loops."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass
    assert "SYNCS" in sass  # mbarrier
