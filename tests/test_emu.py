"""Kernel bodies (spfft_b200/csrc/stage_kernels.hpp, fft_tile.hpp) emulated block by block on the
CPU (tests/emu/emu_stages.cpp) against the numpy oracle. Unit-tests tile/index arithmetic and the
butterflies where no GPU is available; the GPU parity tests proper are tests/test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

from oracle import spfft_oracle as orc


@pytest.fixture(scope="module")
def emu(built):
    lib = C.CDLL(built.EMU_LIB)
    lib.sb_emu_transform.restype = C.c_int
    lib.sb_emu_tile_fft.restype = C.c_int
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 18, 20, 25, 30, 36, 45, 49, 50, 60, 64, 75, 90, 100,
                               120, 128, 144, 150, 192, 200, 210, 225, 240, 250, 360, 400, 450, 625, 720, 2048])
@pytest.mark.parametrize("backward", [0, 1])
def test_tile_fft_double(emu, n, backward):
    rng = np.random.default_rng(n)
    for log2v, swz in ((0, 0), (3, 0), (3, 1)):
        v = 1 << log2v
        data = (rng.standard_normal((n, v)) + 1j * rng.standard_normal((n, v))).astype(np.complex128)
        buf = data.copy()
        if swz:  # element n of sequence `lane` lives at n*V + (lane ^ (n & (V-1)))
            for i in range(n):
                buf[i, np.arange(v) ^ (i & (v - 1))] = data[i]
        assert emu.sb_emu_tile_fft(0, n, log2v, backward, swz, _ptr(buf), 64) == 0
        out = buf.copy()
        if swz:
            for i in range(n):
                out[i] = buf[i, np.arange(v) ^ (i & (v - 1))]
        ref = np.fft.ifft(data, axis=0) * n if backward else np.fft.fft(data, axis=0)
        assert orc.rel_l2(out, ref) < 1e-14 * max(1, np.log2(n + 1))


def test_tile_fft_float(emu):
    rng = np.random.default_rng(0)
    n, log2v = 96, 4
    data = (rng.standard_normal((n, 16)) + 1j * rng.standard_normal((n, 16))).astype(np.complex64)
    buf = data.copy()
    assert emu.sb_emu_tile_fft(1, n, log2v, 1, 0, _ptr(buf), 128) == 0
    assert orc.rel_l2(buf, np.fft.ifft(data.astype(np.complex128), axis=0) * n) < 1e-6


SHAPES = [(1, 1, 1), (2, 2, 2), (11, 12, 13), (13, 11, 2), (12, 1, 11), (1, 13, 12), (20, 18, 16)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("center", [False, True])
def test_emulated_c2c(emu, gen, shape, center):
    nx, ny, nz = shape
    trip, vals = gen.make(nx, ny, nz, center=center)
    param = orc.Parameters(orc.SPFFT_TRANS_C2C, nx, ny, nz, trip)
    out = np.full((nz, ny, nx), np.nan + 0j, dtype=np.complex128)
    t = np.ascontiguousarray(trip.reshape(-1))
    assert emu.sb_emu_transform(0, 0, nx, ny, nz, len(trip), _ptr(t), 0, _ptr(vals), _ptr(out), 0, 32, -1) == 0
    ref = orc.backward(param, vals)
    assert orc.rel_l2(out, ref) < 1e-13
    back = np.zeros(len(trip), dtype=np.complex128)
    assert emu.sb_emu_transform(0, 0, nx, ny, nz, len(trip), _ptr(t), 1, _ptr(out), _ptr(back), 1, 32, 1) == 0
    assert orc.rel_l2(back, vals) < 1e-13


@pytest.mark.parametrize("shape", SHAPES)
def test_emulated_r2c(emu, gen, shape):
    from conftest import hermitian_space_values
    nx, ny, nz = shape
    trip, _ = gen.make(nx, ny, nz, hermitian=True)
    vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(orc.SPFFT_TRANS_R2C, nx, ny, nz, trip)
    out = np.full((nz, ny, nx), np.nan, dtype=np.float64)
    t = np.ascontiguousarray(trip.reshape(-1))
    assert emu.sb_emu_transform(0, 1, nx, ny, nz, len(trip), _ptr(t), 0, _ptr(vals), _ptr(out), 0, 32, -1) == 0
    ref = orc.backward(param, vals)
    assert orc.rel_l2(out, ref) < 1e-13
    back = np.zeros(len(trip), dtype=np.complex128)
    assert emu.sb_emu_transform(0, 1, nx, ny, nz, len(trip), _ptr(t), 1, _ptr(out), _ptr(back), 1, 32, -1) == 0
    # not a round trip: the random index set may hold only one of two conjugate partners on the
    # x = Nx/2 column, which the backward transform does not complete
    assert orc.rel_l2(back, orc.forward(param, out, orc.SPFFT_FULL_SCALING)) < 1e-13


def test_emulated_float_and_duplicates(emu, gen):
    nx, ny, nz = 12, 11, 13
    trip, vals = gen.make(nx, ny, nz)
    # duplicate triplets are legal: last one wins on backward, all receive the value on forward
    trip = np.concatenate([trip, trip[:5]], axis=0)
    vals = np.concatenate([vals, vals[:5] * 3.0])
    param = orc.Parameters(orc.SPFFT_TRANS_C2C, nx, ny, nz, trip)
    v32 = vals.astype(np.complex64)
    out = np.zeros((nz, ny, nx), dtype=np.complex64)
    t = np.ascontiguousarray(trip.reshape(-1))
    assert emu.sb_emu_transform(1, 0, nx, ny, nz, len(trip), _ptr(t), 0, _ptr(v32), _ptr(out), 0, 64, -1) == 0
    ref = orc.backward(param, vals)
    assert orc.rel_l2(out, ref) < 2e-6
    back = np.zeros(len(trip), dtype=np.complex64)
    assert emu.sb_emu_transform(1, 0, nx, ny, nz, len(trip), _ptr(t), 1, _ptr(out), _ptr(back), 1, 64, -1) == 0
    assert orc.rel_l2(back[:5], back[-5:]) < 1e-6


@pytest.mark.parametrize("elem,log2v", [(16, 0), (16, 1), (16, 2), (16, 3), (8, 0), (8, 1), (8, 2), (8, 3), (8, 4)])
def test_x_stage_swizzle(emu, elem, log2v):
    """SwzX (fast_fft.hpp) must be a bijection of the N*V tile for every lane count the x stage kernels
    can pick (a fold that is conflict free but not injective silently corrupts the exchange)."""
    for n in (32, 64, 128, 256, 512, 1024, 2048):
        v = 1 << log2v
        out = np.full(n * v, -2, dtype=np.int32)
        assert emu.sb_emu_swzx(elem, log2v, n, _ptr(out)) == 0
        assert np.array_equal(np.sort(out), np.arange(n * v))


@pytest.mark.parametrize("shape", [(32, 32, 32), (96, 32, 64), (64, 32, 160), (12, 192, 32)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("single", [False, True])
def test_emulated_duplicates_on_register_fft_axes(emu, gen, shape, single):
    """Duplicate triplets (legal: last one wins on backward, all receive the value on forward,
    compression_host.hpp:88-91) switch the z stage of the register-FFT kernels to the scatter form."""
    nx, ny, nz = shape
    trip, vals = gen.make(nx, ny, nz, center=True, stick_fraction=0.5, fill_fraction=0.6)
    trip = np.concatenate([trip, trip[:7]], axis=0)
    vals = np.concatenate([vals, vals[:7] * 3.0])
    param = orc.Parameters(orc.SPFFT_TRANS_C2C, nx, ny, nz, trip)
    cdt = np.complex64 if single else np.complex128
    tol = 3e-6 if single else 1e-13
    v = vals.astype(cdt)
    out = np.full((nz, ny, nx), np.nan, dtype=cdt)
    t = np.ascontiguousarray(trip.reshape(-1))
    assert emu.sb_emu_transform(int(single), 0, nx, ny, nz, len(trip), _ptr(t), 0, _ptr(v), _ptr(out), 0, 64, -1) == 0
    assert orc.rel_l2(out, orc.backward(param, vals)) < tol
    back = np.zeros(len(trip), dtype=cdt)
    assert emu.sb_emu_transform(int(single), 0, nx, ny, nz, len(trip), _ptr(t), 1, _ptr(out), _ptr(back), 1, 64, -1) == 0
    assert orc.rel_l2(back[:7], back[-7:]) < tol


# ---- register-FFT ("fast") bodies: power-of-two lengths, mixed with generic axes ---------------
FAST_SHAPES = [(32, 32, 32), (64, 32, 128), (32, 12, 64), (11, 64, 32), (128, 32, 13), (256, 32, 32),
               (32, 512, 32), (32, 32, 1024), (1024, 32, 12), (512, 12, 32),
               # 3 * 2^k lengths (fast3_stage_kernels.hpp), alone and mixed with the other kernel families
               (96, 96, 96), (192, 32, 12), (12, 192, 32), (32, 13, 192), (384, 96, 32), (32, 384, 96),
               (96, 32, 384), (768, 12, 32), (12, 768, 32), (33, 32, 768),
               # 5 * 2^k lengths (same kernels, radix-5 step, 40 values per thread)
               (160, 160, 32), (12, 320, 32), (32, 13, 640), (320, 96, 64), (640, 12, 160)]


@pytest.mark.parametrize("shape", FAST_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("shuffle", [False, True])
def test_emulated_fast_c2c(emu, gen, shape, single, shuffle):
    nx, ny, nz = shape
    trip, vals = gen.make(nx, ny, nz, center=True, stick_fraction=0.5, fill_fraction=0.6)
    if shuffle:  # arbitrary user order -> scatter-form z kernels instead of the inverse-map form
        perm = np.random.default_rng(5).permutation(len(trip))
        trip, vals = np.ascontiguousarray(trip[perm]), vals[perm]
    param = orc.Parameters(orc.SPFFT_TRANS_C2C, nx, ny, nz, trip)
    cdt = np.complex64 if single else np.complex128
    tol = 2e-6 if single else 1e-13
    v = vals.astype(cdt)
    out = np.full((nz, ny, nx), np.nan, dtype=cdt)
    t = np.ascontiguousarray(trip.reshape(-1))
    assert emu.sb_emu_transform(int(single), 0, nx, ny, nz, len(trip), _ptr(t), 0, _ptr(v), _ptr(out), 0, 64, -1) == 0
    ref = orc.backward(param, vals)
    assert orc.rel_l2(out, ref) < tol
    back = np.zeros(len(trip), dtype=cdt)
    assert emu.sb_emu_transform(int(single), 0, nx, ny, nz, len(trip), _ptr(t), 1, _ptr(out), _ptr(back), 1, 64, -1) == 0
    assert orc.rel_l2(back, vals) < tol


@pytest.mark.parametrize("shape", [(32, 32, 32), (64, 128, 32), (12, 32, 64), (33, 64, 32), (96, 96, 96), (192, 32, 96),
                                   (32, 192, 12), (384, 96, 32), (33, 96, 192), (768, 12, 96),
                                   (160, 160, 32), (320, 12, 160), (33, 640, 32)],
                         ids=lambda s: "x".join(map(str, s)))
def test_emulated_fast_r2c(emu, gen, shape):
    from conftest import hermitian_space_values
    nx, ny, nz = shape
    trip, _ = gen.make(nx, ny, nz, hermitian=True)
    vals = hermitian_space_values(orc, nx, ny, nz, trip)
    param = orc.Parameters(orc.SPFFT_TRANS_R2C, nx, ny, nz, trip)
    out = np.full((nz, ny, nx), np.nan, dtype=np.float64)
    t = np.ascontiguousarray(trip.reshape(-1))
    assert emu.sb_emu_transform(0, 1, nx, ny, nz, len(trip), _ptr(t), 0, _ptr(vals), _ptr(out), 0, 32, -1) == 0
    ref = orc.backward(param, vals)
    assert orc.rel_l2(out, ref) < 1e-13
    back = np.zeros(len(trip), dtype=np.complex128)
    assert emu.sb_emu_transform(0, 1, nx, ny, nz, len(trip), _ptr(t), 1, _ptr(out), _ptr(back), 1, 32, -1) == 0
    assert orc.rel_l2(back, orc.forward(param, out, orc.SPFFT_FULL_SCALING)) < 1e-13


@pytest.mark.parametrize("shape", [(12, 11, 13), (32, 64, 32), (96, 192, 96), (33, 96, 64), (160, 320, 32)],
                         ids=lambda s: "x".join(map(str, s)))
def test_emulated_r2c_negative_half_input(emu, gen, shape):
    """R2C input given at -y on the x = 0 plane and at negative z on stick (0,0) (details.rst:37-40): the
    gather-form kernels complete the hermitian half while loading (hermitian_combine)."""
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
    out = np.full((nz, ny, nx), np.nan, dtype=np.float64)
    t = np.ascontiguousarray(trip.reshape(-1))
    assert emu.sb_emu_transform(0, 1, nx, ny, nz, len(trip), _ptr(t), 0, _ptr(vals), _ptr(out), 0, 32, -1) == 0
    assert orc.rel_l2(out, orc.backward(param, vals)) < 1e-13
    assert orc.rel_l2(out, orc.dense_backward(1, nx, ny, nz, trip, vals)) < 1e-13


# ---- distributed transforms: all ranks emulated in one process, peer-store and block exchange ------
def _emu_distributed(emu, gen, ttype, shape, world, sdist, pdist, center, single, peer, wire_f32=False):
    from conftest import hermitian_space_values
    nx, ny, nz = shape
    trips, vals = [], []
    for r in range(world):
        t, v = gen.make(nx, ny, nz, hermitian=bool(ttype), center=center, num_ranks=world, rank=r,
                        stick_distribution=sdist)
        trips.append(np.ascontiguousarray(t))
        vals.append(v)
    if ttype:
        full = hermitian_space_values(orc, nx, ny, nz, np.concatenate(trips))
        off = 0
        for r in range(world):
            vals[r] = full[off:off + len(trips[r])]
            off += len(trips[r])
    planes = gen.plane_split(nz, pdist)
    params = orc.distributed_parameters(ttype, nx, ny, nz, trips, planes)
    ref_slabs = orc.backward_distributed(params, vals)
    ref_back = orc.forward_distributed(params, ref_slabs, orc.SPFFT_FULL_SCALING)
    cdt = np.complex64 if single else np.complex128
    sdt = (np.float32 if single else np.float64) if ttype else cdt
    tol = 3e-6 if (single or wire_f32) else 1e-13
    n_loc = (C.c_int * world)(*[len(t) for t in trips])
    pl = (C.c_int * world)(*planes)
    flat = [np.ascontiguousarray(t.reshape(-1)) if len(t) else np.zeros(3, np.int32) for t in trips]
    tp = (C.c_void_p * world)(*[f.ctypes.data for f in flat])
    vin = [np.ascontiguousarray(v.astype(cdt)) if len(v) else np.zeros(1, cdt) for v in vals]
    slabs = [np.full(max(planes[r] * ny * nx, 1), np.nan, dtype=sdt) for r in range(world)]
    pin = (C.c_void_p * world)(*[a.ctypes.data for a in vin])
    pout = (C.c_void_p * world)(*[a.ctypes.data for a in slabs])
    assert emu.sb_emu_transform_distributed(int(single), ttype, nx, ny, nz, world, n_loc, tp, pl, 0, pin, pout, 0, 64,
                                            int(peer), int(wire_f32)) == 0
    for r in range(world):
        if planes[r]:
            assert orc.rel_l2(slabs[r][:planes[r] * ny * nx].reshape(planes[r], ny, nx), ref_slabs[r]) < tol
    back = [np.zeros(max(len(v), 1), dtype=cdt) for v in vals]
    pback = (C.c_void_p * world)(*[a.ctypes.data for a in back])
    assert emu.sb_emu_transform_distributed(int(single), ttype, nx, ny, nz, world, n_loc, tp, pl, 1, pout, pback, 1, 64,
                                            int(peer), int(wire_f32)) == 0
    for r in range(world):
        if len(vals[r]):
            assert orc.rel_l2(back[r][:len(vals[r])], ref_back[r]) < tol


DIST_CASES = [
    (0, (11, 12, 13), "uniform", "uniform", False, False),
    (0, (12, 13, 11), "first", "uniform", False, False),
    (0, (13, 11, 12), "first", "last", False, False),
    (1, (12, 11, 13), "uniform", "uniform", False, False),
    (0, (32, 32, 32), "uniform", "uniform", True, False),
    (0, (64, 32, 128), "uniform", "ramp", True, False),
    (1, (64, 64, 32), "uniform", "uniform", False, False),
    (0, (32, 64, 32), "uniform", "uniform", True, True),
    (0, (96, 96, 96), "uniform", "uniform", True, False),
    (1, (192, 96, 32), "uniform", "ramp", False, False),
    (0, (32, 192, 96), "first", "last", True, True),
    (0, (160, 32, 160), "uniform", "ramp", True, False),
]


@pytest.mark.parametrize("case", DIST_CASES, ids=lambda c: f"{'r2c' if c[0] else 'c2c'}-{'x'.join(map(str, c[1]))}-{c[2]}-{c[3]}")
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("peer", [1, 0], ids=["peer-stores", "block-exchange"])
def test_emulated_distributed(emu, gen, case, world, peer):
    """The distributed stage kernels (tables of index_plan.cpp::build_exchange_plan, peer stores of the z
    and y kernels, gather and scatter forms of the distributed y stage) for every kernel family, with the
    fixtures of the reference's MPI tests (tests/mpi_tests/test_transform.cpp: uniform, all sticks on rank 0,
    planes on the last rank) -- the same cases tests/dist_gpu_check.py runs on real GPUs."""
    ttype, shape, sname, pname, center, single = case
    dists = {"uniform": [1.0] * world, "first": [1.0] + [0.0] * (world - 1), "last": [0.0] * (world - 1) + [1.0],
             "ramp": [1.0 + r for r in range(world)]}
    _emu_distributed(emu, gen, ttype, shape, world, dists[sname], dists[pname], center, single, peer)


@pytest.mark.parametrize("case", [c for c in DIST_CASES if not c[5]],
                         ids=lambda c: f"{'r2c' if c[0] else 'c2c'}-{'x'.join(map(str, c[1]))}-{c[2]}-{c[3]}")
@pytest.mark.parametrize("peer", [1, 0], ids=["peer-stores", "block-exchange"])
def test_emulated_distributed_float_wire(emu, gen, case, peer):
    """SPFFT_EXCH_*_FLOAT: a double-precision distributed transform whose exchanged buffers hold cx<float>
    (reference: complex_conversion.cuh:35-54 in the pack / unpack kernels). Accuracy is that of a
    single-precision exchange (details.rst:74), everything else stays double; and the exchange must really
    be single precision (the result differs from the double-precision exchange)."""
    ttype, shape, sname, pname, center, single = case
    world = 2
    dists = {"uniform": [1.0] * world, "first": [1.0] + [0.0] * (world - 1), "last": [0.0] * (world - 1) + [1.0],
             "ramp": [1.0 + r for r in range(world)]}
    _emu_distributed(emu, gen, ttype, shape, world, dists[sname], dists[pname], center, False, peer, wire_f32=True)
    if sname == "uniform":  # data really crosses ranks: a double-precision exchange would pass 1e-13
        with pytest.raises(AssertionError):
            _emu_distributed_strict(emu, gen, ttype, shape, world, dists[sname], dists[pname], center, peer)


def _emu_distributed_strict(emu, gen, ttype, shape, world, sdist, pdist, center, peer):
    """float wire format checked against the double tolerance (expected to fail)."""
    import unittest.mock as mock
    real = orc.rel_l2
    with mock.patch.object(orc, "rel_l2", lambda a, b: real(a, b) * 1e7):  # 3e-6 * 1e-7 < 1e-12
        _emu_distributed(emu, gen, ttype, shape, world, sdist, pdist, center, False, peer, wire_f32=True)


@pytest.mark.parametrize("elem,log2v,sizes", [(16, 0, (1024,)), (16, 1, (512, 1024)), (16, 2, (256, 512)), (16, 3, (64, 128, 256)),
                                              (8, 0, (1024,)), (8, 1, (512,)), (8, 2, (256,)), (8, 3, (64, 128, 256))])
def test_x_stage_swizzle_is_conflict_free(emu, elem, log2v, sizes):
    """Bank-conflict model of tools/check_swizzle.py (32 banks x 4 bytes; 16-byte accesses in phases of 8
    threads, 8-byte ones in phases of 16) on the real SwzX addresses, for the (length, rows per tile)
    combinations the x stage kernels pick (x_lanes_log2): every exchange write and read of the
    power-of-two plan must take one wavefront per phase."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "check_swizzle", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "check_swizzle.py"))
    cs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cs)
    v = 1 << log2v
    phase = 8 if elem == 16 else 16
    slots = 128 // elem
    for n in sizes:
        table = np.zeros(n * v, dtype=np.int32)
        assert emu.sb_emu_swzx(elem, log2v, n, _ptr(table)) == 0
        t = n // 8
        radix, ns = cs.plan(n)
        worst = 1

        def check(addresses):
            seen = {}
            for a in addresses:
                seen.setdefault(a % slots, set()).add(a)
            return max(len(s) for s in seen.values())

        threads = t * v
        for s in range(len(radix) - 1):
            r, nss = radix[s], ns[s]
            m = 8 // r
            for i in range(m):
                for q in range(r):
                    for w0 in range(0, threads, phase):
                        addrs = []
                        for tid in range(w0, min(w0 + phase, threads)):
                            j, lane = tid % t, tid // t
                            b = j + i * t
                            k = b % nss
                            addrs.append(int(table[((b - k) * r + k + q * nss) * v + lane]))
                        worst = max(worst, check(addrs))
            for mm in range(8):
                for w0 in range(0, threads, phase):
                    addrs = [int(table[((tid % t) + t * mm) * v + tid // t]) for tid in range(w0, min(w0 + phase, threads))]
                    worst = max(worst, check(addrs))
        assert worst == 1, f"N={n} V={v} elem={elem}: {worst} wavefronts per phase"


def _fuzz_cases(count, seed):
    rng = np.random.default_rng(seed)
    dims = [1, 2, 3, 5, 8, 12, 13, 20, 32, 33, 64, 96, 100, 128, 160, 192]
    cases = []
    while len(cases) < count:
        shape = tuple(int(rng.choice(dims)) for _ in range(3))
        if shape[0] * shape[1] * shape[2] > 700_000:
            continue
        cases.append((shape, int(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2)),
                      float(rng.choice([0.2, 0.6, 1.0])), float(rng.choice([0.3, 0.7, 1.0])), bool(rng.integers(0, 2)),
                      int(rng.integers(0, 1 << 30))))
    return cases


@pytest.mark.parametrize("case", _fuzz_cases(40, 2024),
                         ids=lambda c: f"{'x'.join(map(str, c[0]))}-{'r2c' if c[1] else 'c2c'}-{'f32' if c[2] else 'f64'}"
                                       f"{'-centered' if c[3] else ''}{'-shuffled' if c[6] else ''}")
def test_emulated_random_mix(emu, gen, case):
    """Random mixes of the kernel families along the three axes (generic, 2^k, 3*2^k, 5*2^k, degenerate lengths),
    transform type, precision, centering, sparsity and value order against the oracle."""
    from conftest import hermitian_space_values
    (nx, ny, nz), ttype, single, center, sf, ff, shuffle, seed = case
    center = center and not ttype
    trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), center=center, stick_fraction=sf, fill_fraction=ff,
                          seed=seed % 100000)
    if len(trip) == 0:
        pytest.skip("empty index set")
    if ttype:
        vals = hermitian_space_values(orc, nx, ny, nz, trip)
    if shuffle:
        perm = np.random.default_rng(seed).permutation(len(trip))
        trip, vals = np.ascontiguousarray(trip[perm]), np.ascontiguousarray(vals[perm])
    param = orc.Parameters(ttype, nx, ny, nz, trip)
    cdt = np.complex64 if single else np.complex128
    sdt = (np.float32 if single else np.float64) if ttype else cdt
    tol = 5e-6 if single else 1e-13
    v = np.ascontiguousarray(vals.astype(cdt))
    out = np.full((nz, ny, nx), np.nan, dtype=sdt)
    t = np.ascontiguousarray(trip.reshape(-1))
    assert emu.sb_emu_transform(int(single), ttype, nx, ny, nz, len(trip), _ptr(t), 0, _ptr(v), _ptr(out), 0, 64, -1) == 0
    ref = orc.backward(param, vals.astype(cdt).astype(np.complex128))
    assert orc.rel_l2(out, ref) < tol
    back = np.zeros(len(trip), dtype=cdt)
    assert emu.sb_emu_transform(int(single), ttype, nx, ny, nz, len(trip), _ptr(t), 1, _ptr(out), _ptr(back), 1, 64, -1) == 0
    assert orc.rel_l2(back, orc.forward(param, ref, orc.SPFFT_FULL_SCALING)) < tol


def _fuzz_dist_cases(count, seed):
    rng = np.random.default_rng(seed)
    dims = [2, 5, 8, 12, 13, 32, 33, 64, 96, 160]
    cases = []
    while len(cases) < count:
        shape = tuple(int(rng.choice(dims)) for _ in range(3))
        if shape[0] * shape[1] * shape[2] > 400_000:
            continue
        world = int(rng.integers(2, 6))
        sdist = [float(x) for x in rng.choice([0.0, 1.0, 2.0, 3.0], size=world)]
        pdist = [float(x) for x in rng.choice([0.0, 1.0, 2.0], size=world)]
        if sum(sdist) == 0:
            sdist[int(rng.integers(0, world))] = 1.0
        if sum(pdist) == 0:
            pdist[int(rng.integers(0, world))] = 1.0
        cases.append((int(rng.integers(0, 2)), shape, world, sdist, pdist, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)),
                      int(rng.integers(0, 2)), bool(rng.integers(0, 4) == 0)))
    return cases


@pytest.mark.parametrize("case", _fuzz_dist_cases(24, 99),
                         ids=lambda c: f"{'r2c' if c[0] else 'c2c'}-{'x'.join(map(str, c[1]))}-P{c[2]}-{'peer' if c[7] else 'block'}"
                                       f"{'-f32' if c[6] else ''}{'-wire' if c[8] else ''}")
def test_emulated_distributed_random_mix(emu, gen, case):
    """Random rank counts (2-5), stick / plane distributions (ranks without sticks or planes included), kernel
    families, precisions, exchange forms and wire formats against the distributed oracle."""
    ttype, shape, world, sdist, pdist, center, single, peer, wire = case
    _emu_distributed(emu, gen, ttype, shape, world, sdist, pdist, center and not ttype, single, peer,
                     wire_f32=wire and not single)


@pytest.mark.parametrize("case", _fuzz_cases(12, 31337), ids=lambda c: f"{'x'.join(map(str, c[0]))}-{'r2c' if c[1] else 'c2c'}")
def test_emulated_kernels_vs_reference_host_library(emu, gen, ref_lib, case):
    """The kernel bodies against the REFERENCE itself: the unmodified reference host pipeline
    (oracle/_ref/libspfft_ref.so, SPFFT_PU_HOST, its own C ABI) on the same random inputs, double precision."""
    from conftest import hermitian_space_values
    from spfft_b200 import capi
    (nx, ny, nz), ttype, _single, center, sf, ff, _shuffle, seed = case
    center = center and not ttype
    trip, vals = gen.make(nx, ny, nz, hermitian=bool(ttype), center=center, stick_fraction=sf, fill_fraction=ff,
                          seed=seed % 100000)
    if len(trip) == 0:
        pytest.skip("empty index set")
    if ttype:
        vals = hermitian_space_values(orc, nx, ny, nz, trip)
    rt = capi.Transform(ref_lib, processing_unit=capi.SPFFT_PU_HOST, transform_type=ttype, dim_x=nx, dim_y=ny,
                        dim_z=nz, indices=trip)
    rt.backward(np.ascontiguousarray(vals), capi.SPFFT_PU_HOST)
    ref_space = rt.space_domain_host_view(ttype).copy()
    ref_back = np.zeros(len(trip), np.complex128)
    rt.forward(capi.SPFFT_PU_HOST, ref_back, capi.SPFFT_FULL_SCALING)
    rt.destroy()
    out = np.full((nz, ny, nx), np.nan, dtype=np.float64 if ttype else np.complex128)
    t = np.ascontiguousarray(trip.reshape(-1))
    v = np.ascontiguousarray(vals.astype(np.complex128))
    assert emu.sb_emu_transform(0, ttype, nx, ny, nz, len(trip), _ptr(t), 0, _ptr(v), _ptr(out), 0, 64, -1) == 0
    assert orc.rel_l2(out, ref_space) <= 1e-12
    back = np.zeros(len(trip), dtype=np.complex128)
    assert emu.sb_emu_transform(0, ttype, nx, ny, nz, len(trip), _ptr(t), 1, _ptr(out), _ptr(back), 1, 64, -1) == 0
    assert orc.rel_l2(back, ref_back) <= 1e-12


# ---- warp FFT (wfft.hpp): the arithmetic bodies, exchange slots and tile address algebra of the length-512 /
# length-256 plans, lane by lane on the CPU (tests/emu/emu_wfft.cpp) ------------------------------------------
@pytest.fixture(scope="module")
def emu_wfft(built):
    lib = C.CDLL(built.EMU_WFFT_LIB)
    for f in ("emu_wfft_f64", "emu_wfft_f32", "emu_wfft_conflicts", "emu_wfft_tile_f64", "emu_wfft_tile_conflicts",
              "emu_wfft_tile_f32x2"):
        getattr(lib, f).restype = C.c_int
    return lib


@pytest.mark.parametrize("n", [512, 256])
@pytest.mark.parametrize("backward", [0, 1])
@pytest.mark.parametrize("single", [False, True])
def test_warp_fft_plan(emu_wfft, n, backward, single):
    """One warp = one (n = 512) or two (n = 256) transforms: 16 values per lane, one exchange, natural order in and
    out; the exchange slots are a bijection (checked inside) and bank-conflict free."""
    rng = np.random.default_rng(n + backward)
    count = 1 if n == 512 else 2
    cdt = np.complex64 if single else np.complex128
    x = (rng.uniform(-1, 1, (count, n)) + 1j * rng.uniform(-1, 1, (count, n))).astype(cdt)
    out = np.zeros_like(x)
    fn = emu_wfft.emu_wfft_f32 if single else emu_wfft.emu_wfft_f64
    assert fn(n, backward, _ptr(x), _ptr(out)) == 0
    ref = (np.fft.ifft(x.astype(np.complex128), axis=1) * n) if backward else np.fft.fft(x.astype(np.complex128), axis=1)
    assert orc.rel_l2(out, ref) < (2e-6 if single else 1e-14)
    assert emu_wfft.emu_wfft_conflicts(n, int(single)) == 1


@pytest.mark.parametrize("W", [2, 4, 8])
@pytest.mark.parametrize("backward", [0, 1])
def test_warp_fft_tile_form(emu_wfft, W, backward):
    """The form the stage kernels run (wfft_kernels.cuh): the warp's column of a [512][W] sub-tile in the TMA swizzle
    of the row width, xor-form addresses equal to the plan's slots, lane twiddles w, w^2, w^4, w^8 with derived
    powers; no other column is touched, no quarter warp hits a 16-byte bank group twice."""
    rng = np.random.default_rng(10 * W + backward)
    for wl in range(W):
        x = rng.uniform(-1, 1, 512) + 1j * rng.uniform(-1, 1, 512)
        out = np.zeros_like(x)
        assert emu_wfft.emu_wfft_tile_f64(W, wl, backward, _ptr(x), _ptr(out)) == 0
        ref = np.fft.ifft(x) * 512 if backward else np.fft.fft(x)
        assert orc.rel_l2(out, ref) < 1e-14
        assert emu_wfft.emu_wfft_tile_conflicts(W, wl) == 1


@pytest.mark.parametrize("W", [4, 8])
@pytest.mark.parametrize("backward", [0, 1])
def test_warp_fft_tile_form_single_pairs(emu_wfft, W, backward):
    """Single precision in the stage kernels: a warp runs TWO transforms packed into 16-byte units (sb::f2, wfft.hpp)
    through the same tile geometry; both come out right and independent of each other."""
    rng = np.random.default_rng(77 * W + backward)
    for wl in range(W):
        x = (rng.uniform(-1, 1, (2, 512)) + 1j * rng.uniform(-1, 1, (2, 512))).astype(np.complex64)
        out = np.zeros_like(x)
        assert emu_wfft.emu_wfft_tile_f32x2(W, wl, backward, _ptr(x), _ptr(out)) == 0
        x64 = x.astype(np.complex128)
        ref = np.fft.ifft(x64, axis=1) * 512 if backward else np.fft.fft(x64, axis=1)
        assert orc.rel_l2(out[0], ref[0]) < 2e-6 and orc.rel_l2(out[1], ref[1]) < 2e-6


@pytest.mark.parametrize("planes,lag,ring", [(1, 8, 1), (3, 8, 3), (8, 8, 8), (9, 8, 9), (24, 8, 18), (64, 8, 18), (512, 8, 18),
                                             (45, 5, 12), (20, 12, 26), (512, 10, 22), (36, 10, 22)])
@pytest.mark.parametrize("tiles", [64, 32])
def test_fused_xy_hand_out_order(emu, planes, lag, ring, tiles):
    """Item order of the fused xy stage (wfft_xy.cu): the dense decode the kernels use enumerates the valid items of
    xy_decode in the same order, dependencies only point backwards (also when there are fewer planes than the lag).
    tiles = items per plane and role: 64 in double precision, 32 in single precision (16 columns / rows per item)."""
    emu.sb_emu_check_xy_order.restype = C.c_int
    assert emu.sb_emu_check_xy_order(planes, lag, ring, tiles) == 0
