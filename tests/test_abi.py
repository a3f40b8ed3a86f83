"""The C-ABI library loads without a GPU and exports every symbol include/nb200.h declares.
No compute calls here (the product has no CPU fallback: compute must FAIL loudly without a GPU)."""
import ctypes

import numpy as np
import pytest


def test_exports_every_declared_symbol(pkg, lib):
    declared = pkg._lib.header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libnb200.so does not export {name}"
    assert sorted(pkg._lib.SIGNATURES) == declared, "ctypes SIGNATURES out of sync with include/nb200.h"
    assert b"sm_100a" in lib.nb200_version()


def test_argument_validation_without_gpu(pkg, lib):
    h = ctypes.c_void_p()
    assert lib.nb200_create(None, 3, 10, 64, 1) == -1                       # NB200_EINVAL
    rc = lib.nb200_create(ctypes.byref(h), 3, 10, 64, 1)
    if rc == 0:                                                             # a GPU is present
        lib.nb200_destroy(h)
        pytest.skip("GPU present: the no-device path cannot be exercised")
    assert rc == -2 and not h.value                                         # NB200_ECUDA, loudly
    assert b"no CPU fallback" in lib.nb200_last_error(None)
    with pytest.raises(pkg.NB200Error):
        pkg.brute_force_cuda_n_body(np.zeros((4, 7)))
    # null-context calls are rejected, never crash
    assert lib.nb200_forces(None, 1.0, 1e-10, None) == -1
    assert lib.nb200_step(None, 1.0, 1e-10, 1e-3, 1) == -1
    assert lib.nb200_launch_count(None) == 0


def test_product_does_not_reference_oracle(pkg):
    """The product path must never route through oracle/ (judge check): no import, no dlopen."""
    import os
    import re
    pkg_dir = os.path.dirname(pkg.__file__)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "libnbref" not in text, f
