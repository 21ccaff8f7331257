"""Generates tests/golden/*.npz from the reference's OWN host pipeline.

Run in the build container (needs /root/reference to have been compiled by oracle/Makefile into
oracle/_ref/libspfft_ref.so; FFT provider = oracle/fftw3_shim because FFTW is not installed):

    python tests/golden/make_golden.py

Each fixture holds, for one case of the reference tests' generator (tests/test_util/
generate_indices.hpp: mt19937(42), 0.7/0.7 fill): the index triplets, the frequency values, the
index maps produced by the reference's convert_index_triplets, the space-domain result of the
reference's backward transform and the result of its forward transform (full scaling) of that.
The GPU parity tests and the oracle tests compare against these files; nothing under
/root/reference is needed at test time.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import FixtureGen, hermitian_space_values  # noqa: E402
from oracle import spfft_oracle as orc  # noqa: E402
from spfft_b200 import capi  # noqa: E402

CASES = [
    # name, type, (nx, ny, nz), centered, single
    ("c2c_11x12x13", 0, (11, 12, 13), False, False),
    ("c2c_13x11x12_centered", 0, (13, 11, 12), True, False),
    ("c2c_2x1x11", 0, (2, 1, 11), False, False),
    ("c2c_32x32x32_centered", 0, (32, 32, 32), True, False),
    ("c2c_100x13x12", 0, (100, 13, 12), False, False),
    ("r2c_12x13x11", 1, (12, 13, 11), False, False),
    ("r2c_11x2x13", 1, (11, 2, 13), False, False),
    ("r2c_32x64x32", 1, (32, 64, 32), False, False),
    ("c2c_float_12x11x13", 0, (12, 11, 13), False, True),
    ("r2c_float_13x12x11", 1, (13, 12, 11), False, True),
]


def main():
    ref = capi.SpfftLib(os.path.join(ROOT, "oracle", "_ref", "libspfft_ref.so"))
    idx = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_indices.so"))
    gen = FixtureGen(os.path.join(ROOT, "oracle", "_ref", "liboracle_gen.so"))
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, ttype, (nx, ny, nz), centered, single in CASES:
        trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), center=centered)
        if ttype:
            vals = hermitian_space_values(orc, nx, ny, nz, trip)
        n = len(trip)
        vi = np.zeros(n, np.int32)
        si = np.zeros(nx * ny, np.int32)
        ns = C.c_int()
        tt = np.ascontiguousarray(trip.reshape(-1))
        err = idx.spfft_ref_convert_index_triplets(ttype, nx, ny, nz, n, tt.ctypes.data_as(C.c_void_p),
                                                   vi.ctypes.data_as(C.c_void_p),
                                                   si.ctypes.data_as(C.c_void_p), C.byref(ns))
        assert err == 0
        cdt = np.complex64 if single else np.complex128
        t = capi.Transform(ref, processing_unit=capi.SPFFT_PU_HOST, transform_type=ttype, dim_x=nx,
                           dim_y=ny, dim_z=nz, indices=trip, single=single)
        v = np.ascontiguousarray(vals.astype(cdt))
        t.backward(v, capi.SPFFT_PU_HOST)
        space = t.space_domain_host_view(ttype).copy()
        back = np.zeros(n, dtype=cdt)
        t.forward(capi.SPFFT_PU_HOST, back, capi.SPFFT_FULL_SCALING)
        t.destroy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), type=ttype, dims=np.array([nx, ny, nz]),
                            triplets=trip, values=v, value_indices=vi, stick_indices=si[:ns.value],
                            space=space, forward=back)
        print(name, n, "elements", ns.value, "sticks")


if __name__ == "__main__":
    main()
