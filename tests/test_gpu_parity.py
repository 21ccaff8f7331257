"""GPU parity tests: the CUDA path driven through the C ABI (libspfft_b200.so) against
  (1) the numpy oracle (oracle/spfft_oracle.py, restating the reference host pipeline),
  (2) the reference's own host library when it travelled with the snapshot (oracle/_ref),
  (3) the committed golden fixtures (tests/golden),
on the reference tests' index generator (tests/test_util/generate_indices.hpp semantics) and shape
grid (tests/local_tests/test_local_transform.cpp:94-110), plus the spherical-cutoff workloads.

Tolerances (north_star): relative L2 error <= 1e-12 in double, <= 1e-5 in single precision.
"""
import itertools
import os

import numpy as np
import pytest

from oracle import spfft_oracle as orc
from spfft_b200 import capi

pytestmark = pytest.mark.gpu

TOL = {False: 1e-12, True: 1e-5}


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(0)
    return torch


def _to_dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.float32 if a.dtype in (np.complex64, np.float32) else np.float64).reshape(-1).copy()).cuda()


def _run_pair(torch, lib, ttype, nx, ny, nz, trip, vals, single=False, grid=None, device_ptrs=True,
              twice=True):
    """backward (+ again, like the reference tests) and forward with full scaling.
    Returns (space, back) as numpy arrays."""
    cdt = np.complex64 if single else np.complex128
    rdt = np.float32 if single else np.float64
    if grid is not None:
        t = grid.create_transform(capi.SPFFT_PU_GPU, ttype, nx, ny, nz, nz, trip)
    else:
        t = capi.Transform(lib, processing_unit=capi.SPFFT_PU_GPU, transform_type=ttype, dim_x=nx,
                           dim_y=ny, dim_z=nz, indices=trip, single=single)
    n = len(trip)
    v = np.ascontiguousarray(vals.astype(cdt))
    sdt = rdt if ttype == capi.SPFFT_TRANS_R2C else cdt
    if device_ptrs:
        d_v = _to_dev(torch, v) if n else None
        nreal = nz * ny * nx * (1 if ttype == capi.SPFFT_TRANS_R2C else 2)
        d_s = torch.full((max(nreal, 1),), float("nan"), dtype=torch.float32 if single else torch.float64, device="cuda")
        for _ in range(2 if twice else 1):
            t.backward_ptr(d_v, d_s)
        space = d_s.cpu().numpy()[:nreal].view(sdt).reshape(nz, ny, nx)
        d_o = torch.zeros(max(2 * n, 1), dtype=d_s.dtype, device="cuda")
        t.forward_ptr(d_s, d_o, capi.SPFFT_FULL_SCALING)
        back = d_o.cpu().numpy()[:2 * n].view(cdt)
    else:
        for _ in range(2 if twice else 1):
            t.backward(v, capi.SPFFT_PU_HOST)
        space = t.space_domain_host_view(ttype).copy()
        back = np.zeros(n, dtype=cdt)
        t.forward(capi.SPFFT_PU_HOST, back, capi.SPFFT_FULL_SCALING)
    t.destroy()
    return space, back


# the reference's shape grid {1,2,11,12,13,100}^3 has 216 members; keep every combination of the
# small sizes and a few with 100
SMALL = [1, 2, 11, 12, 13]
SHAPES = list(itertools.product(SMALL, SMALL, SMALL)) + [(100, 11, 2), (13, 100, 12), (2, 12, 100), (100, 100, 100)]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_c2c_reference_shapes(torch_cuda, lib, gen, shape):
    nx, ny, nz = shape
    trip, vals = gen.make(nx, ny, nz)
    param = orc.Parameters(orc.SPFFT_TRANS_C2C, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, capi.SPFFT_TRANS_C2C, nx, ny, nz, trip, vals)
    ref = orc.backward(param, vals)
    assert orc.rel_l2(space, ref) <= TOL[False]
    assert orc.rel_l2(space, orc.dense_backward(0, nx, ny, nz, trip, vals)) <= TOL[False]
    assert orc.rel_l2(back, orc.forward(param, ref, orc.SPFFT_FULL_SCALING)) <= TOL[False]
    if len(trip):
        assert orc.rel_l2(back, vals) <= TOL[False]


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 2, 2), (11, 11, 11), (11, 12, 13), (12, 13, 11), (100, 100, 100)],
                         ids=lambda s: "x".join(map(str, s)))
def test_c2c_centered(torch_cuda, lib, gen, shape):
    # tests/local_tests/test_local_transform.cpp:103-110
    nx, ny, nz = shape
    trip, vals = gen.make(nx, ny, nz, center=True)
    param = orc.Parameters(orc.SPFFT_TRANS_C2C, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, capi.SPFFT_TRANS_C2C, nx, ny, nz, trip, vals)
    assert orc.rel_l2(space, orc.backward(param, vals)) <= TOL[False]
    assert orc.rel_l2(back, vals) <= TOL[False]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_r2c_reference_shapes(torch_cuda, lib, gen, shape):
    from conftest import hermitian_space_values
    nx, ny, nz = shape
    trip, _ = gen.make(nx, ny, nz, hermitian=True)
    vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(orc.SPFFT_TRANS_R2C, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, capi.SPFFT_TRANS_R2C, nx, ny, nz, trip, vals)
    ref = orc.backward(param, vals)
    assert space.dtype == np.float64
    assert orc.rel_l2(space, ref) <= TOL[False]
    assert orc.rel_l2(back, orc.forward(param, ref, orc.SPFFT_FULL_SCALING)) <= TOL[False]


@pytest.mark.parametrize("shape", [(12, 11, 13), (32, 64, 32), (64, 32, 128), (96, 192, 96), (192, 96, 32), (33, 96, 64)],
                         ids=lambda s: "x".join(map(str, s)))
def test_r2c_negative_half_input(torch_cuda, lib, gen, shape):
    """Input given at -y on the x=0 plane / negative z on stick (0,0) (details.rst:37-40): the
    hermitian fills run in both directions (generic, power-of-two and 3*2^k kernels: the register-FFT
    kernels complete the column / the stick while gathering)."""
    from conftest import hermitian_space_values
    nx, ny, nz = shape
    trip, _ = gen.make(nx, ny, nz, hermitian=True, stick_fraction=1.0, fill_fraction=1.0)
    trip = trip.copy()
    sel = (trip[:, 0] == 0) & (trip[:, 1] > 0)
    trip[sel, 1] = ny - trip[sel, 1]
    trip[sel, 2] = (nz - trip[sel, 2]) % nz
    sel0 = (trip[:, 0] == 0) & (trip[:, 1] == 0) & (trip[:, 2] > 0)
    trip[sel0, 2] = nz - trip[sel0, 2]
    vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(orc.SPFFT_TRANS_R2C, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, capi.SPFFT_TRANS_R2C, nx, ny, nz, trip, vals)
    assert orc.rel_l2(space, orc.backward(param, vals)) <= TOL[False]
    assert orc.rel_l2(space, orc.dense_backward(1, nx, ny, nz, trip, vals)) <= TOL[False]
    assert orc.rel_l2(back, vals) <= TOL[False]


@pytest.mark.parametrize("ttype", [capi.SPFFT_TRANS_C2C, capi.SPFFT_TRANS_R2C])
@pytest.mark.parametrize("shape", [(11, 12, 13), (100, 100, 100), (64, 64, 64)], ids=lambda s: "x".join(map(str, s)))
def test_single_precision(torch_cuda, lib, gen, ttype, shape):
    from conftest import hermitian_space_values
    nx, ny, nz = shape
    herm = ttype == capi.SPFFT_TRANS_R2C
    trip, vals = gen.make(nx, ny, nz, hermitian=herm)
    if herm:
        vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(ttype, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, ttype, nx, ny, nz, trip, vals, single=True)
    ref = orc.backward(param, vals.astype(np.complex64).astype(np.complex128))
    assert space.dtype == (np.float32 if herm else np.complex64)
    assert orc.rel_l2(space, ref) <= TOL[True]
    assert orc.rel_l2(back, orc.forward(param, ref, orc.SPFFT_FULL_SCALING)) <= TOL[True]


@pytest.mark.parametrize("ttype", [capi.SPFFT_TRANS_C2C, capi.SPFFT_TRANS_R2C])
def test_host_pointers_and_internal_buffers(torch_cuda, lib, gen, ttype):
    """Host pointers in and out, space domain in the transform's own (pinned) host buffer --
    how the reference's tests drive the GPU backend (tests/test_util/test_transform.hpp:251-275)."""
    from conftest import hermitian_space_values
    nx, ny, nz = 13, 12, 11
    herm = ttype == capi.SPFFT_TRANS_R2C
    trip, vals = gen.make(nx, ny, nz, hermitian=herm)
    if herm:
        vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(ttype, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, ttype, nx, ny, nz, trip, vals, device_ptrs=False)
    ref = orc.backward(param, vals)
    assert orc.rel_l2(space, ref) <= TOL[False]
    assert orc.rel_l2(back, orc.forward(param, ref, orc.SPFFT_FULL_SCALING)) <= TOL[False]


def test_grid_shared_by_transforms_and_smaller_transform(torch_cuda, lib, gen):
    """Transforms smaller than their grid, created from one grid, run one after the other
    (transform_internal.cpp:55-70; Appendix C of SURVEY.md)."""
    grid = capi.Grid(lib, 16, 16, 16, 256, capi.SPFFT_PU_GPU, -1)
    assert (grid.max_dim_x(), grid.max_dim_y(), grid.max_dim_z()) == (16, 16, 16)
    assert grid.max_num_local_z_columns() == 256 and grid.max_local_z_length() == 16
    assert grid.processing_unit() == capi.SPFFT_PU_GPU
    for shape in [(16, 16, 16), (11, 12, 13), (2, 16, 1)]:
        nx, ny, nz = shape
        trip, vals = gen.make(nx, ny, nz)
        param = orc.Parameters(0, nx, ny, nz, trip)
        space, back = _run_pair(torch_cuda, lib, 0, nx, ny, nz, trip, vals, grid=grid)
        assert orc.rel_l2(space, orc.backward(param, vals)) <= TOL[False]
        assert orc.rel_l2(back, vals) <= TOL[False]
    # limits: too many sticks / too large dims / wrong local z length
    trip, _ = gen.make(16, 16, 16, stick_fraction=1.0, fill_fraction=0.2)
    small = capi.Grid(lib, 16, 16, 16, 10, capi.SPFFT_PU_GPU, -1)
    with pytest.raises(capi.SpfftError) as e:
        small.create_transform(capi.SPFFT_PU_GPU, 0, 16, 16, 16, 16, trip)
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    with pytest.raises(capi.SpfftError) as e:
        grid.create_transform(capi.SPFFT_PU_GPU, 0, 17, 16, 16, 16, trip)
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    with pytest.raises(capi.SpfftError) as e:
        grid.create_transform(capi.SPFFT_PU_GPU, 0, 16, 16, 16, 8, trip)
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    with pytest.raises(capi.SpfftError) as e:
        grid.create_transform(capi.SPFFT_PU_HOST, 0, 16, 16, 16, 16, trip)
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR


def test_grid_with_both_processing_unit_bits(torch_cuda, lib, gen):
    """A Grid may be created for SPFFT_PU_HOST | SPFFT_PU_GPU (grid_internal.cpp:76-99); a transform picks
    exactly one unit (transform_internal.cpp:71-80). A SPFFT_PU_HOST transform is served by the same device kernels
    with its data staged through host memory (there is no CPU path): it accepts host locations only."""
    both = capi.SPFFT_PU_HOST | capi.SPFFT_PU_GPU
    grid = capi.Grid(lib, 12, 12, 12, 144, both, -1)
    assert grid.processing_unit() == both
    nx, ny, nz = 12, 11, 12
    trip, vals = gen.make(nx, ny, nz)
    param = orc.Parameters(0, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, 0, nx, ny, nz, trip, vals, grid=grid)
    assert orc.rel_l2(space, orc.backward(param, vals)) <= TOL[False]
    for bad in (both, 0):
        with pytest.raises(capi.SpfftError) as e:
            grid.create_transform(bad, 0, nx, ny, nz, nz, trip)
        assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    th = grid.create_transform(capi.SPFFT_PU_HOST, 0, nx, ny, nz, nz, trip)
    th.backward(np.ascontiguousarray(vals), capi.SPFFT_PU_HOST)
    assert orc.rel_l2(th.space_domain_host_view(0), orc.backward(param, vals)) <= TOL[False]
    back = np.zeros(len(trip), np.complex128)
    th.forward(capi.SPFFT_PU_HOST, back, capi.SPFFT_FULL_SCALING)
    assert orc.rel_l2(back, vals) <= TOL[False]
    with pytest.raises(capi.SpfftError) as e:  # transform_internal.cpp:212-232: no device-side space domain
        th.backward(np.ascontiguousarray(vals), capi.SPFFT_PU_GPU)
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    th.destroy()


def test_empty_and_degenerate(torch_cuda, lib):
    # no elements at all: legal, output is zero
    t = capi.Transform(lib, transform_type=0, dim_x=4, dim_y=5, dim_z=6, indices=np.zeros((0, 3), np.int32))
    d_s = torch_cuda.full((2 * 4 * 5 * 6,), float("nan"), dtype=torch_cuda.float64, device="cuda")
    t.backward_ptr(None, d_s)
    assert float(d_s.abs().max()) == 0.0
    t.forward_ptr(d_s, None)
    assert t.num_local_elements() == 0 and t.local_slice_size() == 120 and t.global_size() == 120
    # a single element
    trip = np.array([[1, 2, 3]], np.int32)
    vals = np.array([1.5 - 2j])
    space, back = _run_pair(torch_cuda, lib, 0, 4, 5, 6, trip, vals)
    assert orc.rel_l2(space, orc.dense_backward(0, 4, 5, 6, trip, vals)) <= TOL[False]
    assert orc.rel_l2(back, vals) <= TOL[False]


def test_in_place_and_unscaled_forward(torch_cuda, lib, gen):
    """Same device buffer for the frequency values and the space domain (README.md:175)."""
    nx, ny, nz = 12, 13, 11
    trip, vals = gen.make(nx, ny, nz)
    param = orc.Parameters(0, nx, ny, nz, trip)
    t = capi.Transform(lib, transform_type=0, dim_x=nx, dim_y=ny, dim_z=nz, indices=trip)
    buf = torch_cuda.zeros(2 * nx * ny * nz, dtype=torch_cuda.float64, device="cuda")
    buf[:2 * len(trip)] = _to_dev(torch_cuda, vals)
    t.backward_ptr(buf, buf)
    ref = orc.backward(param, vals)
    assert orc.rel_l2(buf.cpu().numpy().view(np.complex128).reshape(nz, ny, nx), ref) <= TOL[False]
    t.forward_ptr(buf, buf, capi.SPFFT_NO_SCALING)
    back = buf.cpu().numpy()[:2 * len(trip)].view(np.complex128)
    assert orc.rel_l2(back, vals * (nx * ny * nz)) <= TOL[False]


def test_internal_device_space_buffer(torch_cuda, lib, gen):
    """backward(..., SPFFT_PU_GPU) leaves the result in the internal device buffer returned by
    space_domain_data(SPFFT_PU_GPU); forward(SPFFT_PU_GPU, ...) consumes it."""
    for ttype in (0, 1):
        nx, ny, nz = 12, 11, 13
        trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype))
        if ttype:
            from conftest import hermitian_space_values
            vals = hermitian_space_values(orc, nx, ny, nz, trip)
        param = orc.Parameters(ttype, nx, ny, nz, trip)
        t = capi.Transform(lib, transform_type=ttype, dim_x=nx, dim_y=ny, dim_z=nz, indices=trip)
        d_v = _to_dev(torch_cuda, vals)
        t.backward(d_v, capi.SPFFT_PU_GPU)
        addr = t.space_domain_data(capi.SPFFT_PU_GPU)
        nreal = nx * ny * nz * (1 if ttype else 2)
        tmp = _dev_read(torch_cuda, addr, nreal)
        ref = orc.backward(param, vals)
        got = tmp.view(np.float64 if ttype else np.complex128).reshape(nz, ny, nx)
        assert orc.rel_l2(got, ref) <= TOL[False]
        d_o = torch_cuda.zeros(2 * len(trip), dtype=torch_cuda.float64, device="cuda")
        t.forward(capi.SPFFT_PU_GPU, d_o, capi.SPFFT_FULL_SCALING)
        assert orc.rel_l2(d_o.cpu().numpy().view(np.complex128), orc.forward(param, ref, 1)) <= TOL[False]


def _dev_read(torch, addr, nreal, dtype=np.float64):
    """numpy copy of `nreal` reals at raw device address `addr`."""
    import ctypes as C
    out = np.zeros(nreal, dtype=dtype)
    cands = [os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so.12"),
             "/usr/local/cuda/lib64/libcudart.so.12"]
    rt = C.CDLL([c for c in cands if os.path.exists(c)][0])
    torch.cuda.synchronize()
    err = rt.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(addr), C.c_size_t(out.nbytes), C.c_int(2))
    assert err == 0
    return out


def test_clone_multi_transform_and_async(torch_cuda, lib, gen):
    """tests/mpi_tests/test_multi_transform.cpp:22-95 restated: clones, constant inputs (i,i),
    backward then forward => i*NxNyNz; plus shared-grid rejection and asynchronous mode."""
    nx, ny, nz = 11, 12, 13
    trip, _ = gen.make(nx, ny, nz)
    n = len(trip)
    t0 = capi.Transform(lib, transform_type=0, dim_x=nx, dim_y=ny, dim_z=nz, indices=trip)
    ts = [t0, t0.clone(), t0.clone()]
    ins = [_to_dev(torch_cuda, np.full(n, (i + 1) * (1 + 1j))) for i in range(3)]
    spaces = [torch_cuda.zeros(2 * nx * ny * nz, dtype=torch_cuda.float64, device="cuda") for _ in range(3)]
    outs = [torch_cuda.zeros(2 * n, dtype=torch_cuda.float64, device="cuda") for _ in range(3)]
    capi.multi_transform_backward_ptr(ts, ins, spaces)
    capi.multi_transform_forward_ptr(ts, spaces, outs, [capi.SPFFT_NO_SCALING] * 3)
    for i in range(3):
        got = outs[i].cpu().numpy().view(np.complex128)
        assert np.allclose(got, (i + 1) * (1 + 1j) * nx * ny * nz, rtol=0, atol=1e-8)
    # internal-buffer flavour
    capi.multi_transform_backward(ts, ins, [capi.SPFFT_PU_GPU] * 3)
    capi.multi_transform_forward(ts, [capi.SPFFT_PU_GPU] * 3, outs, [capi.SPFFT_FULL_SCALING] * 3)
    for i in range(3):
        assert np.allclose(outs[i].cpu().numpy().view(np.complex128), (i + 1) * (1 + 1j), atol=1e-12)
    # transforms of one grid share buffers -> rejected (multi_transform_internal.hpp:69-76)
    grid = capi.Grid(lib, nx, ny, nz, nx * ny, capi.SPFFT_PU_GPU, -1)
    g1 = grid.create_transform(capi.SPFFT_PU_GPU, 0, nx, ny, nz, nz, trip)
    g2 = grid.create_transform(capi.SPFFT_PU_GPU, 0, nx, ny, nz, nz, trip)
    with pytest.raises(capi.SpfftError) as e:
        capi.multi_transform_backward_ptr([g1, g2], ins[:2], spaces[:2])
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    # asynchronous execution: ordered with the default stream only
    assert t0.execution_mode() == capi.SPFFT_EXEC_SYNCHRONOUS
    t0.set_execution_mode(capi.SPFFT_EXEC_ASYNCHRONOUS)
    assert t0.execution_mode() == capi.SPFFT_EXEC_ASYNCHRONOUS
    t0.backward_ptr(ins[0], spaces[0])
    t0.forward_ptr(spaces[0], outs[0], capi.SPFFT_FULL_SCALING)
    torch_cuda.cuda.synchronize()
    assert np.allclose(outs[0].cpu().numpy().view(np.complex128), 1 + 1j, atol=1e-12)


@pytest.mark.parametrize("case", [(0, (64, 64, 64), False, 3), (0, (32, 12, 96), True, 5), (1, (96, 32, 13), False, 4),
                                  (0, (11, 12, 13), False, 35), (1, (192, 64, 32), False, 2), (0, (256, 32, 128), True, 3)],
                         ids=lambda c: f"{'r2c' if c[0] else 'c2c'}-{'x'.join(map(str, c[1]))}-{'f32' if c[2] else 'f64'}-{c[3]}bands")
def test_batched_multi_transform(torch_cuda, lib, gen, case):
    """Clones of one plan through spfft_multi_transform_*_ptr with device pointers run as ONE launch
    per stage (band_kernels.cu, blockIdx.y = band). Every band must equal the same transform
    executed on its own (same kernels => bit identical) and the oracle; the launch counter proves
    that the batched path ran."""
    from conftest import hermitian_space_values
    torch = torch_cuda
    ttype, (nx, ny, nz), single, bands = case
    trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), center=not ttype, stick_fraction=0.5, fill_fraction=0.6)
    if ttype:
        vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(ttype, nx, ny, nz, trip)
    cdt = np.complex64 if single else np.complex128
    rdt = torch.float32 if single else torch.float64
    n = len(trip)
    t0 = capi.Transform(lib, transform_type=ttype, dim_x=nx, dim_y=ny, dim_z=nz, indices=trip, single=single)
    ts = [t0] + [t0.clone() for _ in range(bands - 1)]
    scale = [1.0 + 0.25 * b for b in range(bands)]
    ins = [_to_dev(torch, (vals * scale[b]).astype(cdt)) for b in range(bands)]
    nreal = nz * ny * nx * (1 if ttype else 2)
    spaces = [torch.full((nreal,), float("nan"), dtype=rdt, device="cuda") for _ in range(bands)]
    outs = [torch.zeros(2 * n, dtype=rdt, device="cuda") for _ in range(bands)]
    l0 = capi.kernel_launch_count(lib)
    capi.multi_transform_backward_ptr(ts, ins, spaces)
    capi.multi_transform_forward_ptr(ts, spaces, outs, [capi.SPFFT_FULL_SCALING] * bands)
    launches = capi.kernel_launch_count(lib) - l0
    assert launches == 6 * ((bands + 31) // 32), f"batched path not taken: {launches} launches"
    # the same transforms one at a time
    ref_space = torch.empty(nreal, dtype=rdt, device="cuda")
    ref_out = torch.empty(2 * n, dtype=rdt, device="cuda")
    sdt = (np.float32 if single else np.float64) if ttype else cdt
    v0 = vals.astype(np.complex64).astype(np.complex128) if single else vals
    oracle_space = orc.backward(param, v0)
    for b in (0, bands // 2, bands - 1):
        t0.backward_ptr(ins[b], ref_space)
        t0.forward_ptr(ref_space, ref_out, capi.SPFFT_FULL_SCALING)
        assert torch.equal(spaces[b], ref_space) and torch.equal(outs[b], ref_out)
        got = spaces[b].cpu().numpy().view(sdt).reshape(nz, ny, nx)
        assert orc.rel_l2(got, oracle_space * scale[b]) <= TOL[single]
    # mixed scaling flags do not qualify for one launch: falls back to the per-transform path, same results
    if bands >= 2:
        flags = [capi.SPFFT_FULL_SCALING, capi.SPFFT_NO_SCALING] + [capi.SPFFT_FULL_SCALING] * (bands - 2)
        outs2 = [torch.zeros(2 * n, dtype=rdt, device="cuda") for _ in range(bands)]
        capi.multi_transform_forward_ptr(ts, spaces, outs2, flags)
        assert torch.equal(outs2[0], outs[0])
        assert float((outs2[1] / (nx * ny * nz) - outs[1]).abs().max()) <= 1e-4 * float(outs[1].abs().max())
    for t in ts:
        t.destroy()


def test_error_codes_on_device(torch_cuda, lib):
    trip = np.array([[0, 0, 0], [5, 0, 0]], np.int32)
    with pytest.raises(capi.SpfftError) as e:
        capi.Transform(lib, transform_type=0, dim_x=4, dim_y=4, dim_z=4, indices=trip)
    assert e.value.code == capi.SPFFT_INVALID_INDICES_ERROR
    with pytest.raises(capi.SpfftError) as e:  # more values than grid points
        capi.Transform(lib, transform_type=0, dim_x=1, dim_y=1, dim_z=1, indices=np.zeros((2, 3), np.int32))
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    # a grid-less SPFFT_PU_HOST transform (what the reference's examples create) works: device kernels, host staging
    th = capi.Transform(lib, processing_unit=capi.SPFFT_PU_HOST, transform_type=0, dim_x=4, dim_y=4, dim_z=4,
                        indices=np.zeros((1, 3), np.int32))
    th.backward(np.array([2.0 + 1.0j]), capi.SPFFT_PU_HOST)
    assert np.allclose(th.space_domain_host_view(0), 2.0 + 1.0j, atol=1e-14)
    th.destroy()
    with pytest.raises(capi.SpfftError) as e:
        capi.Grid(lib, 0, 4, 4, 4, capi.SPFFT_PU_GPU, 1)
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR


def test_index_maps_of_transform_match_reference(torch_cuda, lib, gen, ref_indices):
    import ctypes as C
    nx, ny, nz = 13, 12, 11
    trip, _ = gen.make(nx, ny, nz, center=True)
    t = capi.Transform(lib, transform_type=0, dim_x=nx, dim_y=ny, dim_z=nz, indices=trip)
    vi, si = capi.transform_index_maps(t)
    n = len(trip)
    rvi = np.zeros(n, np.int32)
    rsi = np.zeros(nx * ny, np.int32)
    ns = C.c_int()
    tt = np.ascontiguousarray(trip.reshape(-1))
    assert ref_indices.spfft_ref_convert_index_triplets(0, nx, ny, nz, n, tt.ctypes.data_as(C.c_void_p),
                                                        rvi.ctypes.data_as(C.c_void_p),
                                                        rsi.ctypes.data_as(C.c_void_p), C.byref(ns)) == 0
    assert np.array_equal(vi, rvi) and np.array_equal(si, rsi[:ns.value])


@pytest.mark.parametrize("ttype,n", [(0, 64), (0, 128), (1, 128), (0, 192)])
def test_spherical_cutoff_vs_oracle(torch_cuda, lib, ttype, n):
    """BASELINE.json configs at sizes the oracle finishes in seconds."""
    from conftest import hermitian_space_values
    trip = orc.spherical_cutoff_triplets(n, hermitian=bool(ttype))
    rng = np.random.default_rng(42)
    if ttype:
        vals = hermitian_space_values(orc, n, n, n, trip)
    else:
        vals = rng.uniform(-1, 1, len(trip)) + 1j * rng.uniform(-1, 1, len(trip))
    param = orc.Parameters(ttype, n, n, n, trip)
    space, back = _run_pair(torch_cuda, lib, ttype, n, n, n, trip, vals, twice=False)
    ref = orc.backward(param, vals)
    assert orc.rel_l2(space, ref) <= TOL[False]
    assert orc.rel_l2(back, vals) <= TOL[False]


def test_against_reference_host_library(torch_cuda, lib, ref_lib, gen):
    """Same C ABI calls on the reference's own host pipeline (SPFFT_PU_HOST) and on this library
    (SPFFT_PU_GPU): results within tolerance, index maps identical."""
    from conftest import hermitian_space_values
    for ttype, shape, center in [(0, (11, 12, 13), False), (0, (100, 13, 12), True), (1, (12, 13, 11), False),
                                 (1, (100, 100, 100), False), (0, (64, 64, 64), True)]:
        nx, ny, nz = shape
        trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), center=center)
        if ttype:
            vals = hermitian_space_values(orc, nx, ny, nz, trip)
        rt = capi.Transform(ref_lib, processing_unit=capi.SPFFT_PU_HOST, transform_type=ttype, dim_x=nx,
                            dim_y=ny, dim_z=nz, indices=trip)
        rt.backward(np.ascontiguousarray(vals), capi.SPFFT_PU_HOST)
        ref_space = rt.space_domain_host_view(ttype).copy()
        ref_back = np.zeros(len(trip), np.complex128)
        rt.forward(capi.SPFFT_PU_HOST, ref_back, capi.SPFFT_FULL_SCALING)
        space, back = _run_pair(torch_cuda, lib, ttype, nx, ny, nz, trip, vals)
        assert orc.rel_l2(space, ref_space) <= TOL[False]
        assert orc.rel_l2(back, ref_back) <= TOL[False]


def test_full_size_properties_512(torch_cuda, lib):
    """BASELINE's headline size (512^3 C2C double, pi/6 fill): size-independent properties --
    backward->forward round trip with full scaling is the identity on the sparse set, linearity,
    and Parseval between the sparse values and the dense space slab."""
    torch = torch_cuda
    n = 512
    trip = orc.spherical_cutoff_triplets(n)
    ne = len(trip)
    t = capi.Transform(lib, transform_type=0, dim_x=n, dim_y=n, dim_z=n, indices=trip)
    del trip
    g = torch.Generator(device="cuda").manual_seed(42)
    a = torch.rand(2 * ne, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    b = torch.rand(2 * ne, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    sa = torch.empty(2 * n ** 3, dtype=torch.float64, device="cuda")
    t.backward_ptr(a, sa)
    # Parseval: sum |space|^2 = N^3 * sum |values|^2 (unnormalised backward transform)
    lhs = float((sa * sa).sum())
    rhs = float((a * a).sum()) * n ** 3
    assert abs(lhs - rhs) <= 1e-12 * rhs
    out = torch.empty(2 * ne, dtype=torch.float64, device="cuda")
    t.forward_ptr(sa, out, capi.SPFFT_FULL_SCALING)
    assert float((out - a).norm() / a.norm()) <= TOL[False]
    # linearity: B(a + 2b) = B(a) + 2 B(b), checked on a strided sample to bound memory
    sb_ = torch.empty(2 * n ** 3, dtype=torch.float64, device="cuda")
    t.backward_ptr(b, sb_)
    lin = sa[::97] + 2 * sb_[::97]
    del sb_
    t.backward_ptr(a + 2 * b, sa)
    assert float((sa[::97] - lin).norm() / lin.norm()) <= TOL[False]
    t.destroy()


@pytest.mark.parametrize("shape", [(32, 32, 32), (64, 32, 128), (32, 12, 64), (11, 64, 32), (128, 32, 13),
                                   (256, 32, 32), (32, 512, 32), (32, 32, 1024), (1024, 32, 32)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("ttype", [0, 1])
def test_power_of_two_fast_path(torch_cuda, lib, gen, shape, single, ttype):
    """Register-FFT kernels (power-of-two axes) mixed with generic axes, both precisions."""
    from conftest import hermitian_space_values
    nx, ny, nz = shape
    trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), center=not ttype, stick_fraction=0.5, fill_fraction=0.6)
    if ttype:
        vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(ttype, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, ttype, nx, ny, nz, trip, vals, single=single)
    v = vals.astype(np.complex64).astype(np.complex128) if single else vals
    ref = orc.backward(param, v)
    assert orc.rel_l2(space, ref) <= TOL[single]
    assert orc.rel_l2(back, orc.forward(param, ref, orc.SPFFT_FULL_SCALING)) <= TOL[single]


@pytest.mark.parametrize("shape", [(96, 96, 96), (192, 32, 12), (12, 192, 32), (32, 13, 192), (384, 96, 32), (32, 384, 96),
                                   (96, 32, 384), (768, 12, 32), (12, 768, 32), (33, 32, 768), (192, 192, 192),
                                   (160, 160, 32), (12, 320, 32), (32, 13, 640), (320, 96, 64), (640, 12, 160)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("ttype", [0, 1])
@pytest.mark.parametrize("shuffle", [False, True])
def test_three_times_power_of_two_fast_path(torch_cuda, lib, gen, shape, single, ttype, shuffle):
    """Register-FFT kernels for N = 3 * 2^k and 5 * 2^k axes (fast3_stage_kernels.hpp) mixed with the other kernel
    families, both precisions; shuffle = values in arbitrary user order (scatter-form z kernels)."""
    from conftest import hermitian_space_values
    nx, ny, nz = shape
    trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), center=not ttype, stick_fraction=0.5, fill_fraction=0.6)
    if ttype:
        vals = hermitian_space_values(orc, nx, ny, nz, trip)
    if shuffle:
        perm = np.random.default_rng(7).permutation(len(trip))
        trip, vals = np.ascontiguousarray(trip[perm]), np.ascontiguousarray(vals[perm])
    param = orc.Parameters(ttype, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, ttype, nx, ny, nz, trip, vals, single=single)
    v = vals.astype(np.complex64).astype(np.complex128) if single else vals
    ref = orc.backward(param, v)
    assert orc.rel_l2(space, ref) <= TOL[single]
    assert orc.rel_l2(back, orc.forward(param, ref, orc.SPFFT_FULL_SCALING)) <= TOL[single]


import glob as _glob

_GOLDEN = sorted(_glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", _GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_golden_fixtures(torch_cuda, lib, path):
    """Committed outputs of the reference's own host pipeline (tests/golden/make_golden.py):
    index maps bit exact, values within the north-star tolerance."""
    g = np.load(path)
    ttype = int(g["type"])
    nx, ny, nz = (int(v) for v in g["dims"])
    single = g["values"].dtype == np.complex64
    t = capi.Transform(lib, transform_type=ttype, dim_x=nx, dim_y=ny, dim_z=nz, indices=g["triplets"], single=single)
    vi, si = capi.transform_index_maps(t)
    assert np.array_equal(vi, g["value_indices"]) and np.array_equal(si, g["stick_indices"])
    t.destroy()
    space, back = _run_pair(torch_cuda, lib, ttype, nx, ny, nz, g["triplets"], g["values"], single=single)
    assert orc.rel_l2(space, g["space"]) <= TOL[single]
    assert orc.rel_l2(back, g["forward"]) <= TOL[single]


def test_baseline_config3_256_r2c_vs_reference_library(torch_cuda, lib, ref_lib):
    """BASELINE.json config 3 at size: 256^3 R2C / C2R double, gamma-point half sphere, value by value against
    the reference's own host library (SPFFT_PU_HOST) on the same exactly hermitian input."""
    from conftest import hermitian_space_values
    n = 256
    trip = orc.spherical_cutoff_triplets(n, hermitian=True)
    vals = hermitian_space_values(orc, n, n, n, trip)
    rt = capi.Transform(ref_lib, processing_unit=capi.SPFFT_PU_HOST, transform_type=1, dim_x=n, dim_y=n, dim_z=n, indices=trip)
    rt.backward(np.ascontiguousarray(vals), capi.SPFFT_PU_HOST)
    ref = rt.space_domain_host_view(1).copy()
    ref_back = np.zeros(len(trip), np.complex128)
    rt.forward(capi.SPFFT_PU_HOST, ref_back, capi.SPFFT_FULL_SCALING)
    rt.destroy()
    space, back = _run_pair(torch_cuda, lib, 1, n, n, n, trip, vals, twice=False)
    assert space.dtype == np.float64
    assert orc.rel_l2(space, ref) <= TOL[False]
    assert orc.rel_l2(back, ref_back) <= TOL[False]


@pytest.mark.parametrize("single", [False, True], ids=["f64", "f32"])
def test_baseline_config5_256_bands_192(torch_cuda, lib, single):
    """BASELINE.json config 5 at size: a batch of 256 bands at 192^3 (spherical cutoff, centered indices,
    out-of-place with external space-domain buffers) through spfft_multi_transform_*_ptr; sampled bands against
    the oracle, round trip of every band."""
    torch = torch_cuda
    n, bands = 192, 256
    trip = orc.spherical_cutoff_triplets(n)
    ne = len(trip)
    rng = np.random.default_rng(5)
    base = rng.uniform(-1, 1, ne) + 1j * rng.uniform(-1, 1, ne)
    cdt = np.complex64 if single else np.complex128
    rdt = torch.float32 if single else torch.float64
    t0 = capi.Transform(lib, transform_type=0, dim_x=n, dim_y=n, dim_z=n, indices=trip, single=single)
    ts = [t0] + [t0.clone() for _ in range(bands - 1)]
    d_base = _to_dev(torch, base.astype(cdt))
    ins = [d_base * (1.0 + 0.5 * b / bands) for b in range(bands)]
    spaces = [torch.empty(2 * n ** 3, dtype=rdt, device="cuda") for _ in range(bands)]
    outs = [torch.empty(2 * ne, dtype=rdt, device="cuda") for _ in range(bands)]
    capi.multi_transform_backward_ptr(ts, ins, spaces)
    capi.multi_transform_forward_ptr(ts, spaces, outs, [capi.SPFFT_FULL_SCALING] * bands)
    torch.cuda.synchronize()
    param = orc.Parameters(0, n, n, n, trip)
    v0 = base.astype(np.complex64).astype(np.complex128) if single else base
    oracle_space = orc.backward(param, v0)
    for b in (0, 101, bands - 1):
        got = spaces[b].cpu().numpy().view(cdt).reshape(n, n, n)
        assert orc.rel_l2(got, oracle_space * (1.0 + 0.5 * b / bands)) <= TOL[single]
    for b in range(bands):
        assert float((outs[b] - ins[b]).norm() / ins[b].norm()) <= TOL[single]
    for t in ts:
        t.destroy()


@pytest.mark.parametrize("case", [(0, (512, 512, 3)), (0, (512, 512, 24)), (0, (512, 512, 45)), (0, (32, 32, 512)), (1, (64, 12, 512)),
                                  (0, (512, 512, 512))],
                         ids=lambda c: ("r2c" if c[0] else "c2c") + "x".join(map(str, c[1])))
def test_warp_fft_kernels(torch_cuda, lib, ref_lib, gen, case):
    """Warp-FFT kernels (wfft_xy.cu: fused xy stage with the hand-off in L2, ring slots reused from 11 planes on;
    wfft_z.cu: z stage incl. the hermitian completion of stick (0,0)), double precision, length 512: against the
    numpy oracle and, at 512^3 (the BASELINE headline shape, spherical cutoff), value by value against the
    reference's own host library on the same inputs."""
    from conftest import hermitian_space_values
    ttype, (nx, ny, nz) = case
    if nx * ny * nz >= 512 ** 3:
        trip = orc.spherical_cutoff_triplets(nx)
        rng = np.random.default_rng(42)
        vals = rng.uniform(-1, 1, len(trip)) + 1j * rng.uniform(-1, 1, len(trip))
        rt = capi.Transform(ref_lib, processing_unit=capi.SPFFT_PU_HOST, transform_type=ttype, dim_x=nx,
                            dim_y=ny, dim_z=nz, indices=trip)
        rt.backward(np.ascontiguousarray(vals), capi.SPFFT_PU_HOST)
        ref = rt.space_domain_host_view(ttype).copy()
        ref_back = np.zeros(len(trip), np.complex128)
        rt.forward(capi.SPFFT_PU_HOST, ref_back, capi.SPFFT_FULL_SCALING)
        rt.destroy()
        space, back = _run_pair(torch_cuda, lib, ttype, nx, ny, nz, trip, vals, twice=False)
        assert orc.rel_l2(space, ref) <= TOL[False]
        assert orc.rel_l2(back, ref_back) <= TOL[False]
        return
    trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), center=not ttype, stick_fraction=0.6, fill_fraction=0.7)
    if ttype:
        vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(ttype, nx, ny, nz, trip)
    space, back = _run_pair(torch_cuda, lib, ttype, nx, ny, nz, trip, vals)
    ref = orc.backward(param, vals)
    assert orc.rel_l2(space, ref) <= TOL[False]
    assert orc.rel_l2(back, orc.forward(param, ref, orc.SPFFT_FULL_SCALING)) <= TOL[False]


@pytest.mark.parametrize("shape,device_ptrs", [((512, 512, 3), True), ((512, 512, 24), True), ((512, 512, 45), True),
                                               ((32, 32, 512), True), ((64, 12, 512), True), ((512, 512, 512), True),
                                               ((512, 512, 36), False)],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) else ("dev" if v else "host"))
def test_warp_fft_kernels_single_precision(torch_cuda, lib, gen, monkeypatch, shape, device_ptrs):
    """Single precision on the warp-FFT kernels (SPFFT_B200_WFFT bit 3): a warp runs two transforms packed into
    16-byte units on the packed fp32 pipe (sb::f2, wfft.hpp). Against the numpy oracle, against the round-1 kernels on
    the same inputs, and the launch count proves which kernels ran (4 per pair: z, fused xy, fused xy, z)."""
    nx, ny, nz = shape
    if nx * ny * nz >= 512 ** 3:
        trip = orc.spherical_cutoff_triplets(nx)
        rng = np.random.default_rng(43)
        vals = rng.uniform(-1, 1, len(trip)) + 1j * rng.uniform(-1, 1, len(trip))
    else:
        trip, vals = gen.make(nx, ny, nz, center=True, stick_fraction=0.6, fill_fraction=0.7)
    monkeypatch.setenv("SPFFT_B200_WFFT", "7")
    space_r1, back_r1 = _run_pair(torch_cuda, lib, 0, nx, ny, nz, trip, vals, single=True, device_ptrs=device_ptrs)
    monkeypatch.setenv("SPFFT_B200_WFFT", "15")
    l0 = capi.kernel_launch_count(lib)
    space, back = _run_pair(torch_cuda, lib, 0, nx, ny, nz, trip, vals, single=True, device_ptrs=device_ptrs, twice=False)
    launches = capi.kernel_launch_count(lib) - l0
    warp_z, warp_xy = nz == 512, nx == 512
    if device_ptrs:
        assert launches == (2 if warp_xy else 4) + 2
    assert np.isfinite(space).all()
    assert orc.rel_l2(space, space_r1) <= 2e-6 and orc.rel_l2(back, back_r1) <= 2e-6
    assert orc.rel_l2(back, vals) <= TOL[True]
    if nx * ny * nz < 512 ** 3:
        param = orc.Parameters(0, nx, ny, nz, trip)
        assert orc.rel_l2(space, orc.backward(param, vals)) <= TOL[True]
    assert warp_z or warp_xy


def test_warp_fft_transforms_on_concurrent_streams(torch_cuda, lib, gen):
    """Several transforms that take the warp-FFT kernels, enqueued together on their own streams through
    spfft_multi_transform_* (asynchronous execution inside the call): the fused xy stage is a cooperative persistent
    kernel, so concurrent launches must serialise instead of dead-locking each other; results as when run alone."""
    torch = torch_cuda
    shapes = [(512, 512, 12), (512, 512, 20), (64, 16, 512)]
    ts, ins, spaces, outs, refs = [], [], [], [], []
    for nx, ny, nz in shapes:
        trip, vals = gen.make(nx, ny, nz, center=True, stick_fraction=0.6, fill_fraction=0.7)
        t = capi.Transform(lib, transform_type=0, dim_x=nx, dim_y=ny, dim_z=nz, indices=trip)
        ts.append(t)
        ins.append(_to_dev(torch, vals))
        spaces.append(torch.full((2 * nx * ny * nz,), float("nan"), dtype=torch.float64, device="cuda"))
        outs.append(torch.zeros(2 * len(trip), dtype=torch.float64, device="cuda"))
        refs.append((orc.backward(orc.Parameters(0, nx, ny, nz, trip), vals), vals, (nz, ny, nx)))
    for _ in range(3):
        capi.multi_transform_backward_ptr(ts, ins, spaces)
        capi.multi_transform_forward_ptr(ts, spaces, outs, [capi.SPFFT_FULL_SCALING] * len(ts))
    torch.cuda.synchronize()
    for i, (ref_space, vals, shp) in enumerate(refs):
        assert orc.rel_l2(spaces[i].cpu().numpy().view(np.complex128).reshape(shp), ref_space) <= TOL[False]
        assert orc.rel_l2(outs[i].cpu().numpy().view(np.complex128), vals) <= TOL[False]
    for t in ts:
        t.destroy()


@pytest.mark.parametrize("shape", [(64, 48, 40), (512, 512, 36), (96, 32, 64)], ids=lambda s: "x".join(map(str, s)))
def test_host_pointer_slabs(torch_cuda, lib, gen, shape):
    """Host-pointer calls of local C2C transforms move the space domain in slabs of planes on a second stream while
    the xy stage of the neighbouring slab runs (enqueue_backward / enqueue_forward): same bits as the device-pointer
    call, also when the call is repeated and through the internal host buffer."""
    nx, ny, nz = shape
    trip, vals = gen.make(nx, ny, nz, center=True, stick_fraction=0.6, fill_fraction=0.7)
    space_d, back_d = _run_pair(torch_cuda, lib, 0, nx, ny, nz, trip, vals)
    space_h, back_h = _run_pair(torch_cuda, lib, 0, nx, ny, nz, trip, vals, device_ptrs=False)
    assert np.array_equal(space_d, space_h) and np.array_equal(back_d, back_h)
    param = orc.Parameters(0, nx, ny, nz, trip)
    assert orc.rel_l2(space_h, orc.backward(param, vals)) <= TOL[False]
    # external pinned-or-not host buffers through the _ptr entry points
    t = capi.Transform(lib, transform_type=0, dim_x=nx, dim_y=ny, dim_z=nz, indices=trip)
    out = np.full((nz, ny, nx), np.nan + 0j, dtype=np.complex128)
    t.backward_ptr(np.ascontiguousarray(vals), out)
    assert np.array_equal(out, space_d)
    back = np.zeros(len(trip), np.complex128)
    t.forward_ptr(out, back, capi.SPFFT_FULL_SCALING)
    assert np.array_equal(back, back_d)
    t.destroy()


def test_distributed_two_gpus(torch_cuda):
    """One process per GPU over NCCL (tests/dist_gpu_check.py); skipped on a single-GPU box."""
    import subprocess
    import sys
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tests", "dist_gpu_check.py")], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "DIST_GPU_CHECK PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
