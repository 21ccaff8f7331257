"""Pins the oracle (oracle/spfft_oracle.py + oracle/fftw3_shim) -- CPU only.

  * against the committed golden fixtures (tests/golden/*.npz, produced by the reference's own host
    pipeline, tests/golden/make_golden.py),
  * against the reference host library itself when it is present (oracle/_ref/libspfft_ref.so),
  * against the dense 3-D DFT the reference's own tests use as truth
    (tests/test_util/test_transform.hpp:41-46),
  * index maps bit for bit against the reference's convert_index_triplets.
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from oracle import spfft_oracle as orc
from spfft_b200 import capi

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 10


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_golden(path):
    g = np.load(path)
    ttype = int(g["type"])
    nx, ny, nz = (int(v) for v in g["dims"])
    single = g["values"].dtype == np.complex64
    tol = 1e-5 if single else 1e-12
    param = orc.Parameters(ttype, nx, ny, nz, g["triplets"])
    # integer maps: bit exact
    assert np.array_equal(param.value_indices, g["value_indices"])
    assert np.array_equal(param.stick_indices, g["stick_indices"])
    vals = g["values"].astype(np.complex128)
    space = orc.backward(param, vals)
    assert orc.rel_l2(space, g["space"]) <= tol
    back = orc.forward(param, g["space"].astype(np.float64 if ttype else np.complex128), orc.SPFFT_FULL_SCALING)
    assert orc.rel_l2(back, g["forward"]) <= tol
    # and the dense DFT definition (C2C; a partial R2C index set may hold one of two conjugate
    # partners, which the pipeline does not complete -- see test_r2c_full_set_like_reference_test)
    if ttype == 0:
        assert orc.rel_l2(orc.dense_backward(ttype, nx, ny, nz, g["triplets"], vals), g["space"]) <= tol


SHAPES = [(1, 1, 1), (2, 2, 2), (1, 13, 2), (11, 12, 13), (13, 12, 11), (12, 11, 1), (100, 11, 12), (16, 32, 8)]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("ttype", [0, 1])
@pytest.mark.parametrize("center", [False, True])
def test_oracle_matches_reference_library(ref_lib, gen, shape, ttype, center):
    from conftest import hermitian_space_values
    if ttype and center:
        pytest.skip("the reference's R2C tests use non-centered indices only")
    nx, ny, nz = shape
    trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), center=center)
    if ttype:
        vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(ttype, nx, ny, nz, trip)
    rt = capi.Transform(ref_lib, processing_unit=capi.SPFFT_PU_HOST, transform_type=ttype, dim_x=nx,
                        dim_y=ny, dim_z=nz, indices=trip)
    rt.backward(np.ascontiguousarray(vals), capi.SPFFT_PU_HOST)
    rt.backward(np.ascontiguousarray(vals), capi.SPFFT_PU_HOST)  # twice, like test_transform.hpp:129-131
    ref_space = rt.space_domain_host_view(ttype).copy()
    ref_back = np.zeros(len(trip), np.complex128)
    rt.forward(capi.SPFFT_PU_HOST, ref_back, capi.SPFFT_FULL_SCALING)
    space = orc.backward(param, vals)
    assert orc.rel_l2(space, ref_space) <= 1e-12
    assert orc.rel_l2(orc.forward(param, space, orc.SPFFT_FULL_SCALING), ref_back) <= 1e-12
    # the reference's own test criterion: dense 3-D DFT, 1e-6 absolute (test_check_values.hpp:46-78)
    if ttype == 0:
        dense = orc.dense_backward(ttype, nx, ny, nz, trip, vals)
        assert np.max(np.abs(ref_space - dense), initial=0.0) < 1e-6


@pytest.mark.parametrize("shape", [(2, 2, 2), (11, 12, 13), (12, 13, 11), (13, 11, 12), (16, 8, 32)],
                         ids=lambda s: "x".join(map(str, s)))
def test_r2c_full_set_like_reference_test(ref_lib, gen, shape):
    """tests/test_util/test_transform.hpp:221-290 restated: all sticks of the non-redundant half
    space, random real space field, forward vs dense DFT, then backward vs N^3 * field."""
    nx, ny, nz = shape
    trip, _ = gen.make(nx, ny, nz, hermitian=True, stick_fraction=1.0, fill_fraction=1.0)
    rng = np.random.default_rng(3)
    field = rng.uniform(0, 1, (nz, ny, nx))
    param = orc.Parameters(1, nx, ny, nz, trip)
    dense = orc.dense_forward(1, nx, ny, nz, trip, field)
    rt = capi.Transform(ref_lib, processing_unit=capi.SPFFT_PU_HOST, transform_type=1, dim_x=nx, dim_y=ny,
                        dim_z=nz, indices=trip)
    rt.space_domain_host_view(1)[...] = field
    freq = np.zeros(len(trip), np.complex128)
    rt.forward(capi.SPFFT_PU_HOST, freq, capi.SPFFT_NO_SCALING)
    assert np.max(np.abs(freq - dense)) < 1e-6                      # the reference's own criterion
    assert orc.rel_l2(orc.forward(param, field), freq) <= 1e-12       # oracle == reference
    rt.backward(freq, capi.SPFFT_PU_HOST)
    ref_space = rt.space_domain_host_view(1).copy()
    assert np.max(np.abs(ref_space - field * (nx * ny * nz))) < 1e-6
    assert orc.rel_l2(orc.backward(param, freq), ref_space) <= 1e-12


@pytest.mark.parametrize("hermitian", [False, True])
@pytest.mark.parametrize("center", [False, True])
def test_index_maps_bit_exact_with_reference(lib, ref_indices, gen, hermitian, center):
    """Product (C ABI, no GPU needed), oracle and reference agree on valueIndices / stickIndices."""
    for shape in [(11, 12, 13), (2, 1, 13), (100, 13, 12), (32, 32, 32)]:
        nx, ny, nz = shape
        trip, _ = gen.make(nx, ny, nz, hermitian=hermitian, center=center and not hermitian)
        rng = np.random.default_rng(1)
        trip = trip[rng.permutation(len(trip))]  # arbitrary user order
        n = len(trip)
        rvi = np.zeros(n, np.int32)
        rsi = np.zeros(nx * ny, np.int32)
        ns = C.c_int()
        tt = np.ascontiguousarray(trip.reshape(-1))
        assert ref_indices.spfft_ref_convert_index_triplets(int(hermitian), nx, ny, nz, n, tt.ctypes.data_as(C.c_void_p),
                                                            rvi.ctypes.data_as(C.c_void_p),
                                                            rsi.ctypes.data_as(C.c_void_p), C.byref(ns)) == 0
        ovi, osi = orc.convert_index_triplets(hermitian, nx, ny, nz, trip)
        pvi, psi = capi.convert_index_triplets(lib, hermitian, nx, ny, nz, trip)
        assert np.array_equal(ovi, rvi) and np.array_equal(osi, rsi[:ns.value])
        assert np.array_equal(pvi, rvi) and np.array_equal(psi, rsi[:ns.value])


def test_index_errors_match_reference(lib, ref_indices):
    cases = [
        (False, 4, 4, 4, [[4, 0, 0]], capi.SPFFT_INVALID_INDICES_ERROR),       # out of range
        (False, 4, 4, 4, [[-1, 0, 0], [3, 0, 0]], capi.SPFFT_INVALID_INDICES_ERROR),  # centered => max 2
        (True, 4, 4, 4, [[-1, 0, 0]], capi.SPFFT_INVALID_INDICES_ERROR),       # hermitian: x >= 0
        (True, 4, 4, 4, [[3, 0, 0]], capi.SPFFT_INVALID_INDICES_ERROR),        # hermitian: x <= n/2
        (False, 1, 1, 1, [[0, 0, 0], [0, 0, 0]], capi.SPFFT_INVALID_PARAMETER_ERROR),  # too many values
        (False, 5, 5, 5, [[-2, 2, -2], [2, -2, 0]], 0),
    ]
    for herm, nx, ny, nz, trip, code in cases:
        t = np.array(trip, np.int32)
        n = len(t)
        vi = np.zeros(n, np.int32)
        si = np.zeros(nx * ny + 1, np.int32)
        ns = C.c_int()
        tt = np.ascontiguousarray(t.reshape(-1))
        ref_code = ref_indices.spfft_ref_convert_index_triplets(int(herm), nx, ny, nz, n, tt.ctypes.data_as(C.c_void_p),
                                                                vi.ctypes.data_as(C.c_void_p),
                                                                si.ctypes.data_as(C.c_void_p), C.byref(ns))
        assert ref_code == code
        if code:
            with pytest.raises(capi.SpfftError) as e:
                capi.convert_index_triplets(lib, herm, nx, ny, nz, t)
            assert e.value.code == code
            with pytest.raises((orc.InvalidIndicesError, orc.InvalidParameterError)):
                orc.convert_index_triplets(herm, nx, ny, nz, t)
        else:
            capi.convert_index_triplets(lib, herm, nx, ny, nz, t)


def test_distributed_oracle_equals_local(gen):
    """The distributed restatement (sticks on several ranks, planes split) reproduces the local one."""
    nx, ny, nz = 12, 13, 11
    nranks = 3
    trips, vals = [], []
    for r in range(nranks):
        t, v = gen.make(nx, ny, nz, num_ranks=nranks, rank=r)
        trips.append(t)
        vals.append(v)
    planes = gen.plane_split(nz, [1.0, 0.0, 2.0])  # a rank without planes
    assert sum(planes) == nz
    params = orc.distributed_parameters(0, nx, ny, nz, trips, planes)
    slabs = orc.backward_distributed(params, vals)
    full = orc.dense_backward(0, nx, ny, nz, np.concatenate(trips), np.concatenate(vals))
    assert orc.rel_l2(np.concatenate([s for s in slabs if s.shape[0]], axis=0), full) <= 1e-12
    back = orc.forward_distributed(params, slabs, orc.SPFFT_FULL_SCALING)
    for b, v in zip(back, vals):
        assert orc.rel_l2(b, v) <= 1e-12
    with pytest.raises(orc.DuplicateIndicesError):
        orc.distributed_parameters(0, nx, ny, nz, [trips[0], trips[0]], [nz, 0])


def test_spherical_workload_counts():
    """Element / stick counts of the benchmark workloads."""
    import bench
    for n, herm, ns_expected, ne_expected in [(64, False, 3207, 137062), (128, False, 12851, 1097914)]:
        t = orc.spherical_cutoff_triplets(n, hermitian=herm)
        assert len(t) == ne_expected
        _, si = orc.convert_index_triplets(herm, n, n, n, t)
        assert len(si) == ns_expected
        assert np.array_equal(bench.spherical_triplets(n, herm), t)
    assert np.array_equal(bench.spherical_triplets(32, True), orc.spherical_cutoff_triplets(32, hermitian=True))


def test_index_maps_random_triplets_vs_reference(lib, ref_indices):
    """Randomised: arbitrary triplet lists (any order, duplicates, negative = centered, sometimes out of
    range, hermitian or not, degenerate dimensions) through the reference's own convert_index_triplets
    (src/compression/indices.hpp:120-186) and through the product: same error code, or bit-identical
    valueIndices / stickIndices."""
    rng = np.random.default_rng(12345)
    ok = bad = 0
    for _ in range(300):
        nx, ny, nz = (int(rng.choice([1, 2, 3, 4, 7, 8, 12, 16])) for _ in range(3))
        herm = bool(rng.integers(0, 2))
        n = int(rng.integers(0, min(nx * ny * nz, 60) + 1))
        mode = rng.integers(0, 4)  # 0: storage indices, 1: centered, 2: occasionally out of range, 3: mixed signs
        t = np.zeros((n, 3), np.int32)
        for d, dim in enumerate((nx, ny, nz)):
            if mode == 0:
                t[:, d] = rng.integers(0, dim, n)
            elif mode == 1:
                t[:, d] = rng.integers(dim // 2 - dim + 1, dim // 2 + 1, n)
            elif mode == 2:
                t[:, d] = rng.integers(-dim, dim + 1, n)
            else:
                t[:, d] = rng.integers(-(dim // 2), dim, n)
        if herm and mode in (0, 1) and n:
            t[:, 0] = np.abs(t[:, 0]) % (nx // 2 + 1)
        vi = np.zeros(max(n, 1), np.int32)
        si = np.zeros(nx * ny + 1, np.int32)
        ns = C.c_int()
        tt = np.ascontiguousarray(t.reshape(-1)) if n else np.zeros(3, np.int32)
        code = ref_indices.spfft_ref_convert_index_triplets(int(herm), nx, ny, nz, n, tt.ctypes.data_as(C.c_void_p),
                                                            vi.ctypes.data_as(C.c_void_p),
                                                            si.ctypes.data_as(C.c_void_p), C.byref(ns))
        if code:
            bad += 1
            with pytest.raises(capi.SpfftError) as e:
                capi.convert_index_triplets(lib, herm, nx, ny, nz, t)
            assert e.value.code == code, (herm, nx, ny, nz, t.tolist())
        else:
            ok += 1
            pvi, psi = capi.convert_index_triplets(lib, herm, nx, ny, nz, t)
            assert np.array_equal(pvi, vi[:n]) and np.array_equal(psi, si[:ns.value]), (herm, nx, ny, nz, t.tolist())
    assert ok > 50 and bad > 20
