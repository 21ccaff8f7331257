"""CPU restatement (numpy) of the reference's sparse 3D transform pipeline.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
(``spfft_b200/`` or ``libspfft_b200.so``); only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and only as the checker.

Every function cites the reference file:line (relative to /root/reference) whose behaviour it
restates.  Integer work (index maps) is bit-exact; floating-point work is done stage by stage in
the reference's order with ``numpy.fft`` (pocketfft, double precision) standing in for FFTW.

Parity pinning: the reference ships no golden vectors.  This oracle is
pinned against outputs of the reference's own host pipeline compiled here from its unmodified
sources (``oracle/_ref/libspfft_ref.so``, recipe ``oracle/Makefile``, FFT provider = the
``oracle/fftw3_shim`` restatement of FFTW's published definition because FFTW is not installed)
and against fixtures generated from it (``tests/golden/*.npz``, script
``tests/golden/make_golden.py``).  See tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np

# --- enums: include/spfft/types.h:33-117, include/spfft/errors.h:37-125 -----------------------
SPFFT_TRANS_C2C = 0
SPFFT_TRANS_R2C = 1
SPFFT_NO_SCALING = 0
SPFFT_FULL_SCALING = 1


class InvalidParameterError(ValueError):
    """src/compression/indices.hpp:124-127 (numValues > dimX*dimY*dimZ)."""


class InvalidIndicesError(ValueError):
    """src/compression/indices.hpp:145-149 (index outside the allowed range)."""


class DuplicateIndicesError(ValueError):
    """src/compression/indices.hpp:105-117 (same z-stick on two ranks)."""


def to_storage_index(dim: int, index: np.ndarray) -> np.ndarray:
    """src/compression/indices.hpp:49-55: negative frequency index -> dim + index."""
    return np.where(index < 0, index + dim, index)


def convert_index_triplets(hermitian: bool, dim_x: int, dim_y: int, dim_z: int,
                           triplets: np.ndarray):
    """src/compression/indices.hpp:120-186.

    Returns (valueIndices[int32, Ne], stickIndices[int32, Ns]):
      stickIndices = sorted unique ``x*dimY + y`` (storage indices),
      valueIndices[i] = position_of_stick(i) * dimZ + z_storage(i).
    """
    t = np.asarray(triplets, dtype=np.int64).reshape(-1, 3)
    n = t.shape[0]
    if n > dim_x * dim_y * dim_z:
        raise InvalidParameterError("more values than grid points")
    x, y, z = t[:, 0], t[:, 1], t[:, 2]
    # :129-135 -- "centered" is decided globally over all three coordinates
    centered = bool(n > 0 and (t < 0).any())
    max_x = (dim_x // 2 + 1 if (hermitian or centered) else dim_x) - 1
    max_y = (dim_y // 2 + 1 if centered else dim_y) - 1
    max_z = (dim_z // 2 + 1 if centered else dim_z) - 1
    min_x = 0 if hermitian else max_x - dim_x + 1
    min_y = max_y - dim_y + 1
    min_z = max_z - dim_z + 1
    if n > 0 and ((x < min_x).any() or (x > max_x).any() or (y < min_y).any() or
                  (y > max_y).any() or (z < min_z).any() or (z > max_z).any()):
        raise InvalidIndicesError("index out of bounds")
    xs = to_storage_index(dim_x, x)
    ys = to_storage_index(dim_y, y)
    zs = to_storage_index(dim_z, z)
    keys = xs * dim_y + ys                      # :152-158 std::map key
    stick_indices = np.unique(keys)             # ordered unique keys (:179-183)
    stick_of_value = np.searchsorted(stick_indices, keys)   # map value = rank of key (:160-165)
    value_indices = stick_of_value * dim_z + zs  # :168-176
    return value_indices.astype(np.int32), stick_indices.astype(np.int32)


def check_stick_duplicates(sticks_per_rank):
    """src/compression/indices.hpp:105-117."""
    seen = set()
    for sticks in sticks_per_rank:
        for s in np.asarray(sticks).tolist():
            if s in seen:
                raise DuplicateIndicesError("z-stick present on two ranks")
            seen.add(s)


class Parameters:
    """Local (single rank) plan parameters -- src/parameters/parameters.cpp:143-180.

    For the distributed constructor (:43-140) use :func:`distributed_parameters`.
    """

    def __init__(self, transform_type: int, dim_x: int, dim_y: int, dim_z: int, triplets):
        self.transform_type = transform_type
        self.dim_x, self.dim_y, self.dim_z = dim_x, dim_y, dim_z
        self.dim_x_freq = dim_x // 2 + 1 if transform_type == SPFFT_TRANS_R2C else dim_x
        self.value_indices, self.stick_indices = convert_index_triplets(
            transform_type == SPFFT_TRANS_R2C, dim_x, dim_y, dim_z, triplets)
        check_stick_duplicates([self.stick_indices])
        self.num_sticks = int(self.stick_indices.size)
        # :173-179 -- position of key 0, or num_sticks if absent
        hit = np.nonzero(self.stick_indices == 0)[0]
        self.zero_zero_stick_index = int(hit[0]) if hit.size else self.num_sticks
        # single rank: all planes local
        self.num_xy_planes = [dim_z]
        self.xy_plane_offsets = [0]
        self.sticks_per_rank = [self.stick_indices]
        self.rank = 0


def distributed_parameters(transform_type, dim_x, dim_y, dim_z, triplets_per_rank,
                           planes_per_rank):
    """All ranks' parameters for a distributed transform -- src/parameters/parameters.cpp:43-140.

    Returns a list of Parameters-like objects (one per rank) sharing the global stick lists.
    """
    size = len(triplets_per_rank)
    assert len(planes_per_rank) == size
    if sum(planes_per_rank) != dim_z:
        raise InvalidParameterError("sum of local z lengths != dimZ")   # :108-110
    out = []
    for r in range(size):
        p = Parameters.__new__(Parameters)
        p.transform_type = transform_type
        p.dim_x, p.dim_y, p.dim_z = dim_x, dim_y, dim_z
        p.dim_x_freq = dim_x // 2 + 1 if transform_type == SPFFT_TRANS_R2C else dim_x
        p.value_indices, p.stick_indices = convert_index_triplets(
            transform_type == SPFFT_TRANS_R2C, dim_x, dim_y, dim_z, triplets_per_rank[r])
        p.num_sticks = int(p.stick_indices.size)
        hit = np.nonzero(p.stick_indices == 0)[0]
        p.zero_zero_stick_index = int(hit[0]) if hit.size else p.num_sticks
        p.rank = r
        out.append(p)
    sticks = [p.stick_indices for p in out]
    check_stick_duplicates(sticks)
    if sum(s.size for s in sticks) > dim_x * dim_y:
        raise InvalidParameterError("more sticks than xy points")      # :104-107
    offs = np.concatenate([[0], np.cumsum(planes_per_rank)[:-1]]).astype(int).tolist()
    for p in out:
        p.num_xy_planes = list(planes_per_rank)
        p.xy_plane_offsets = offs
        p.sticks_per_rank = sticks
    return out


# ----------------------------------------------------------------------------------------------
# Stages (host pipeline order: src/execution/execution_host.cpp:247-352)
# ----------------------------------------------------------------------------------------------

def decompress(param, values: np.ndarray, dtype=np.complex128) -> np.ndarray:
    """src/compression/compression_host.hpp:76-92: zero the sticks, then
    sticks.flat[valueIndices[i]] = values[i] (later i overwrites earlier on duplicates)."""
    sticks = np.zeros((param.num_sticks, param.dim_z), dtype=dtype)
    v = np.asarray(values).reshape(-1)
    if v.dtype.kind != "c":
        v = v.reshape(-1, 2)
        v = v[:, 0] + 1j * v[:, 1]
    flat = sticks.reshape(-1)
    flat[param.value_indices] = v.astype(dtype)   # numpy fancy assignment: last write wins
    return sticks


def compress(param, sticks: np.ndarray, scaling: int) -> np.ndarray:
    """src/compression/compression_host.hpp:55-73; scale = T(1/(NxNyNz)),
    src/execution/execution_host.cpp:53-54."""
    out = sticks.reshape(-1)[param.value_indices]
    if scaling == SPFFT_FULL_SCALING:
        real_t = np.float32 if sticks.dtype == np.complex64 else np.float64
        scale = real_t(1.0 / float(param.dim_x * param.dim_y * param.dim_z))
        out = out * scale
    return out.astype(sticks.dtype)


def _hermitian_fill_1d(vec: np.ndarray) -> None:
    """Sequential low-to-high fill used by both symmetry classes
    (src/symmetry/symmetry_host.hpp:47-58 plane, :73-90 stick):
    for i in 1..n-1: if v[i] != 0: v[n-i] = conj(v[i])."""
    n = vec.shape[0]
    for i in range(1, n):
        val = vec[i]
        if val != 0:
            vec[n - i] = np.conj(val)


def stick_symmetry(param, sticks: np.ndarray) -> None:
    """StickSymmetryHost::apply on the (x=0,y=0) stick, src/symmetry/symmetry_host.hpp:68-94,
    wired at src/execution/execution_host.cpp:91-96 (R2C only, if the stick is local)."""
    if param.transform_type != SPFFT_TRANS_R2C:
        return
    if param.zero_zero_stick_index < param.num_sticks:
        _hermitian_fill_1d(sticks[param.zero_zero_stick_index])


def z_transform(sticks: np.ndarray, backward: bool) -> np.ndarray:
    """Transform1DPlanesHost along z on every stick (src/fft/transform_1d_host.hpp:56-125),
    unnormalised, sign + backward / - forward (docs/source/details.rst:6-13)."""
    if sticks.shape[0] == 0:
        return sticks
    n = sticks.shape[1]
    res = np.fft.ifft(sticks, axis=1) * n if backward else np.fft.fft(sticks, axis=1)
    return res.astype(sticks.dtype)


def sticks_to_planes(param, sticks_per_rank, z_offset: int, num_planes: int) -> np.ndarray:
    """Backward transpose, src/transpose/transpose_host.hpp:75-117 (local) and the compact
    exchange src/transpose/transpose_mpi_compact_buffered_host.cpp:83-175 (distributed):
    zero the planes, plane[z][x][y] = stick[z].  Returned array is (z, y, x) (GPU plane layout,
    src/execution/execution_gpu.cpp:88-89); values are layout independent."""
    dt = sticks_per_rank[0].dtype if len(sticks_per_rank) else np.complex128
    planes = np.zeros((num_planes, param.dim_y, param.dim_x_freq), dtype=dt)
    for keys, sticks in zip(param.sticks_per_rank, sticks_per_rank):
        if len(keys) == 0:
            continue
        x = keys // param.dim_y
        y = keys % param.dim_y
        planes[:, y, x] = sticks[:, z_offset:z_offset + num_planes].T
    return planes


def planes_to_sticks(param, planes_per_rank, rank: int) -> np.ndarray:
    """Forward transpose (src/transpose/transpose_host.hpp:121-153): gather this rank's sticks
    from every rank's planes. planes_per_rank[r] is (nz_r, y, x)."""
    keys = param.sticks_per_rank[rank]
    dt = planes_per_rank[0].dtype
    sticks = np.zeros((len(keys), param.dim_z), dtype=dt)
    if len(keys) == 0:
        return sticks
    x = keys // param.dim_y
    y = keys % param.dim_y
    for r, planes in enumerate(planes_per_rank):
        z0 = param.xy_plane_offsets[r]
        nz = param.num_xy_planes[r]
        if nz:
            sticks[:, z0:z0 + nz] = planes[:, y, x].T
    return sticks


def plane_symmetry(param, planes: np.ndarray) -> None:
    """PlaneSymmetryHost::apply, src/symmetry/symmetry_host.hpp:43-63: on column x=0 of every
    plane, hermitian fill along y (R2C only; wired execution_host.cpp:98)."""
    if param.transform_type != SPFFT_TRANS_R2C:
        return
    for z in range(planes.shape[0]):
        _hermitian_fill_1d(planes[z, :, 0])


def xy_backward(param, planes: np.ndarray) -> np.ndarray:
    """y transform on the x columns then x transform (C2C, or C2R writing unpadded reals):
    src/execution/execution_host.cpp:339-348, transform_1d_host.hpp:161-207,
    transform_real_1d_host.hpp:52-241."""
    if planes.shape[0] == 0:
        return planes
    ny = param.dim_y
    t = np.fft.ifft(planes, axis=1) * ny
    if param.transform_type == SPFFT_TRANS_R2C:
        out = np.fft.irfft(t, n=param.dim_x, axis=2) * param.dim_x
        return out.astype(np.float32 if planes.dtype == np.complex64 else np.float64)
    out = np.fft.ifft(t, axis=2) * param.dim_x
    return out.astype(planes.dtype)


def xy_forward(param, space: np.ndarray) -> np.ndarray:
    """x transform (C2C or R2C) then y transform: src/execution/execution_host.cpp:247-262."""
    if space.shape[0] == 0:
        cdt = np.complex64 if space.dtype in (np.float32, np.complex64) else np.complex128
        return np.zeros((0, param.dim_y, param.dim_x_freq), dtype=cdt)
    if param.transform_type == SPFFT_TRANS_R2C:
        cdt = np.complex64 if space.dtype == np.float32 else np.complex128
        t = np.fft.rfft(space, axis=2)
    else:
        cdt = space.dtype
        t = np.fft.fft(space, axis=2)
    return np.fft.fft(t, axis=1).astype(cdt)


# ----------------------------------------------------------------------------------------------
# Whole transforms
# ----------------------------------------------------------------------------------------------

def backward(param, values, dtype=np.complex128) -> np.ndarray:
    """Local backward transform, stage order of ExecutionHost::backward_z / backward_xy
    (src/execution/execution_host.cpp:298-352). Returns (z, y, x) space data."""
    sticks = decompress(param, values, dtype)
    stick_symmetry(param, sticks)
    sticks = z_transform(sticks, backward=True)
    planes = sticks_to_planes(param, [sticks], 0, param.dim_z)
    plane_symmetry(param, planes)
    return xy_backward(param, planes)


def forward(param, space, scaling=SPFFT_NO_SCALING) -> np.ndarray:
    """Local forward transform (src/execution/execution_host.cpp:247-296). ``space`` is
    (z, y, x) complex (C2C) or real (R2C). Returns the Ne compressed complex values."""
    space = np.asarray(space)
    planes = xy_forward(param, space.reshape(param.dim_z, param.dim_y, param.dim_x))
    sticks = planes_to_sticks(param, [planes], 0)
    sticks = z_transform(sticks, backward=False)
    return compress(param, sticks, scaling)


def backward_distributed(params, values_per_rank, dtype=np.complex128):
    """Distributed backward: every rank's z stage, the all-to-all, every rank's xy stage.
    Returns the list of local slabs."""
    sticks = []
    for p, v in zip(params, values_per_rank):
        s = decompress(p, v, dtype)
        stick_symmetry(p, s)
        sticks.append(z_transform(s, backward=True))
    out = []
    for r, p in enumerate(params):
        planes = sticks_to_planes(p, sticks, p.xy_plane_offsets[r], p.num_xy_planes[r])
        plane_symmetry(p, planes)
        out.append(xy_backward(p, planes))
    return out


def forward_distributed(params, space_per_rank, scaling=SPFFT_NO_SCALING):
    planes = []
    for r, (p, s) in enumerate(zip(params, space_per_rank)):
        s = np.asarray(s).reshape(p.num_xy_planes[r], p.dim_y, p.dim_x)
        planes.append(xy_forward(p, s))
    out = []
    for r, p in enumerate(params):
        sticks = planes_to_sticks(p, planes, r)
        sticks = z_transform(sticks, backward=False)
        out.append(compress(p, sticks, scaling))
    return out


# ----------------------------------------------------------------------------------------------
# Dense truth (what the reference's own tests compare against: fftw_plan_dft_3d,
# tests/test_util/test_transform.hpp:41-46)
# ----------------------------------------------------------------------------------------------

def dense_cube(transform_type, dim_x, dim_y, dim_z, triplets, values, dtype=np.complex128):
    """Scatter the sparse values into a dense (z, y, x) frequency cube; for R2C the missing
    hermitian half is completed (value at -k = conj(value at k)) exactly like the two symmetry
    fills do for consistent input."""
    t = np.asarray(triplets, dtype=np.int64).reshape(-1, 3)
    v = np.asarray(values).reshape(-1)
    if v.dtype.kind != "c":
        v = v.reshape(-1, 2)
        v = v[:, 0] + 1j * v[:, 1]
    cube = np.zeros((dim_z, dim_y, dim_x), dtype=dtype)
    xs = to_storage_index(dim_x, t[:, 0])
    ys = to_storage_index(dim_y, t[:, 1])
    zs = to_storage_index(dim_z, t[:, 2])
    if transform_type == SPFFT_TRANS_R2C:
        cube[(-zs) % dim_z, (-ys) % dim_y, (-xs) % dim_x] = np.conj(v)
    cube[zs, ys, xs] = v
    return cube


def dense_backward(transform_type, dim_x, dim_y, dim_z, triplets, values):
    cube = dense_cube(transform_type, dim_x, dim_y, dim_z, triplets, values)
    out = np.fft.ifftn(cube) * (dim_x * dim_y * dim_z)
    return out.real.copy() if transform_type == SPFFT_TRANS_R2C else out


def dense_forward(transform_type, dim_x, dim_y, dim_z, triplets, space, scaling=SPFFT_NO_SCALING):
    t = np.asarray(triplets, dtype=np.int64).reshape(-1, 3)
    cube = np.fft.fftn(np.asarray(space).reshape(dim_z, dim_y, dim_x))
    xs = to_storage_index(dim_x, t[:, 0])
    ys = to_storage_index(dim_y, t[:, 1])
    zs = to_storage_index(dim_z, t[:, 2])
    out = cube[zs, ys, xs]
    if scaling == SPFFT_FULL_SCALING:
        out = out / (dim_x * dim_y * dim_z)
    return out


def rel_l2(a, b) -> float:
    """Parity metric: ||a-b||_2 / ||b||_2 (0/0 -> 0)."""
    a = np.asarray(a).astype(np.complex128).reshape(-1)
    b = np.asarray(b).astype(np.complex128).reshape(-1)
    nb = np.linalg.norm(b)
    na = np.linalg.norm(a - b)
    if nb == 0:
        return float(na)
    return float(na / nb)


# ----------------------------------------------------------------------------------------------
# Synthetic workloads (SURVEY.md section 8d / BASELINE.md section 3)
# ----------------------------------------------------------------------------------------------

def spherical_cutoff_triplets(n: int, hermitian: bool = False, centered: bool = True,
                              radius: float | None = None) -> np.ndarray:
    """All (kx,ky,kz) in the centered range [-n/2+1, n/2]^3 with |k|^2 <= (n/2)^2 (pi/6 fill),
    grouped by stick in ascending storage key x*Ny+y, z ascending (storage) within a stick.
    hermitian=True keeps the non-redundant half: kx>=0; kx==0 -> ky>=0; kx==ky==0 -> kz>=0."""
    r = n / 2 if radius is None else radius
    k = np.arange(-(n // 2) + (1 if n % 2 == 0 else 0), n // 2 + 1, dtype=np.int64)
    kx, ky = np.meshgrid(k, k, indexing="ij")
    kx = kx.reshape(-1)
    ky = ky.reshape(-1)
    keep = kx * kx + ky * ky <= r * r
    if hermitian:
        keep &= (kx > 0) | ((kx == 0) & (ky >= 0))
    kx, ky = kx[keep], ky[keep]
    key = (kx % n) * n + (ky % n)
    order = np.argsort(key, kind="stable")
    kx, ky = kx[order], ky[order]
    out = []
    # z in storage order 0..n/2, then -n/2+1..-1
    kz_storage_order = np.concatenate([k[k >= 0], k[k < 0]])
    kz2 = kz_storage_order * kz_storage_order
    rem = r * r - (kx * kx + ky * ky)
    for x, y, m in zip(kx.tolist(), ky.tolist(), rem.tolist()):
        zsel = kz_storage_order[kz2 <= m]
        if hermitian and x == 0 and y == 0:
            zsel = zsel[zsel >= 0]
        blk = np.empty((zsel.size, 3), dtype=np.int32)
        blk[:, 0] = x
        blk[:, 1] = y
        blk[:, 2] = zsel
        out.append(blk)
    trip = np.concatenate(out, axis=0) if out else np.zeros((0, 3), np.int32)
    if not centered:
        trip = np.stack([trip[:, 0] % n, trip[:, 1] % n, trip[:, 2] % n], axis=1).astype(np.int32)
    return np.ascontiguousarray(trip, dtype=np.int32)
