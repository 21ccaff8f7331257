"""Shared fixtures. `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Everything compiled (product library, emulation harness, oracle helpers)."""
    import __graft_entry__ as g
    g.build()
    return g


@pytest.fixture(scope="session")
def lib(built):
    from spfft_b200 import capi
    return capi.load()


@pytest.fixture(scope="session")
def ref_lib(built):
    """The UNMODIFIED reference host pipeline (oracle/_ref/libspfft_ref.so), or skip."""
    from spfft_b200 import capi
    path = os.path.join(REF_DIR, "libspfft_ref.so")
    if not os.path.exists(path):
        pytest.skip("reference host library not built (needs /root/reference at build time)")
    return capi.SpfftLib(path)


@pytest.fixture(scope="session")
def ref_indices(built):
    path = os.path.join(REF_DIR, "libref_indices.so")
    if not os.path.exists(path):
        pytest.skip("reference index wrapper not built")
    return C.CDLL(path)


class FixtureGen:
    """The reference tests' random index/value generator (oracle/gen_indices.cpp)."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.spfft_oracle_gen_fixture.restype = C.c_longlong

    def make(self, dim_x, dim_y, dim_z, hermitian=False, center=False, seed=42, num_ranks=1,
             rank=0, stick_distribution=None, stick_fraction=0.7, fill_fraction=0.7,
             real_values_only=False):
        dist = np.ascontiguousarray(
            np.ones(num_ranks) if stick_distribution is None else stick_distribution, dtype=np.float64)
        args = (C.c_uint(seed), C.c_int(num_ranks), dist.ctypes.data_as(C.POINTER(C.c_double)),
                C.c_double(stick_fraction), C.c_double(fill_fraction), C.c_int(dim_x),
                C.c_int(dim_y), C.c_int(dim_z), C.c_int(int(hermitian)),
                C.c_int(int(real_values_only)), C.c_int(int(center)), C.c_int(rank))
        n = self.lib.spfft_oracle_gen_fixture(*args, None, None)
        trip = np.zeros((n, 3), dtype=np.int32)
        vals = np.zeros(2 * n, dtype=np.float64)
        if n:
            self.lib.spfft_oracle_gen_fixture(*args, trip.ctypes.data_as(C.POINTER(C.c_int)),
                                              vals.ctypes.data_as(C.POINTER(C.c_double)))
        return trip, vals.view(np.complex128).copy()

    def plane_split(self, dim_z, distribution):
        dist = np.ascontiguousarray(distribution, dtype=np.float64)
        out = np.zeros(len(dist), dtype=np.int32)
        self.lib.spfft_oracle_plane_split(C.c_int(dim_z), C.c_int(len(dist)),
                                          dist.ctypes.data_as(C.POINTER(C.c_double)),
                                          out.ctypes.data_as(C.POINTER(C.c_int)))
        return out.tolist()


@pytest.fixture(scope="session")
def gen(built):
    return FixtureGen(os.path.join(REF_DIR, "liboracle_gen.so"))


def hermitian_space_values(orc, dim_x, dim_y, dim_z, trip, seed=7):
    """Frequency values on `trip` of a random REAL space field (exactly hermitian input)."""
    rng = np.random.default_rng(seed)
    space = rng.uniform(-1, 1, (dim_z, dim_y, dim_x))
    vals = orc.dense_forward(orc.SPFFT_TRANS_R2C, dim_x, dim_y, dim_z, trip, space)
    return np.ascontiguousarray(vals, dtype=np.complex128)
