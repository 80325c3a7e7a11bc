"""pytest configuration: the `gpu` marker, oracle / reference / golden fixtures.

`-m "not gpu"` : oracle vs. golden vectors and the reference's known-answer tests, host logic, C-ABI symbol
                 table (no compute calls).  `-m gpu` : the CUDA path through the C ABI vs. the oracle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    if not os.path.exists(orc.ORACLE_SO):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, capture_output=True)
    return orc.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own interpolation.c, compiled unmodified; skipped where it could not be built."""
    from oracle import oracle as orc

    if not os.path.exists(orc.REF_SO):
        if os.path.isdir("/root/reference"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
        else:
            pytest.skip("oracle/_ref/libmifi_ref.so not built (no /root/reference on this machine)")
    return orc.Reference()


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))

    return load


@pytest.fixture(scope="session")
def lib():
    """The product library through its C ABI (ctypes).  Fails loudly when it is not built."""
    import fimex_b200

    return fimex_b200


def bits(a):
    """float32 array -> uint32 view for bit-exact comparison (NaN payloads included)"""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(a, b, what="", nan_payload=False):
    """Every non-NaN value bit-identical (signed zeros and infinities included) and the NaN masks identical.

    With nan_payload=True the NaN bit patterns must match too: that holds for copied values and for the fill
    value MIFI_UNDEFINED_F (0x7fc00000), i.e. everything nearest-neighbour produces.  NaNs that come out of
    ARITHMETIC (bilinear / bicubic / rotation with a NaN tap) carry the operand's payload on x86 and the
    canonical 0x7fffffff on NVIDIA GPUs; the value is NaN either way and the next stage of the reference
    (interpolationArray2Data, CDMInterpolator.cc:121-124) replaces it by the fill value."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    ua, ub = a.view(np.uint32).ravel(), b.view(np.uint32).ravel()
    differ = ua != ub
    if not nan_payload:
        differ &= ~(np.isnan(a).ravel() & np.isnan(b).ravel())
    bad = np.flatnonzero(differ)
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} values differ bitwise; first at {bad[:5]}: " \
                          f"{a.ravel()[bad[:5]]} vs {b.ravel()[bad[:5]]}"
