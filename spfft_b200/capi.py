"""ctypes binding of the SpFFT C ABI (include/spfft/{grid,transform,multi_transform}[_float].h).

The binding is library-agnostic: it drives ``libspfft_b200.so`` (this repo's product) and -- in
the tests and the CPU-baseline leg of bench.py -- the reference host library compiled under
``oracle/_ref`` through the *same* entry points, which is what "drop-in" means here.

Class and method names mirror the reference's C++ API (include/spfft/grid.hpp:49-203,
include/spfft/transform.hpp:56-315): Grid.create_transform, Transform.backward / forward /
space_domain_data / clone / local_z_length ...  Errors raise SpfftError carrying the C error code
(include/spfft/errors.h:37-125).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

# include/spfft/types.h
SPFFT_EXCH_DEFAULT = 0
SPFFT_EXCH_BUFFERED = 1
SPFFT_EXCH_BUFFERED_FLOAT = 2
SPFFT_EXCH_COMPACT_BUFFERED = 3
SPFFT_EXCH_COMPACT_BUFFERED_FLOAT = 4
SPFFT_EXCH_UNBUFFERED = 5
SPFFT_PU_HOST = 1
SPFFT_PU_GPU = 2
SPFFT_INDEX_TRIPLETS = 0
SPFFT_TRANS_C2C = 0
SPFFT_TRANS_R2C = 1
SPFFT_NO_SCALING = 0
SPFFT_FULL_SCALING = 1
SPFFT_EXEC_SYNCHRONOUS = 0
SPFFT_EXEC_ASYNCHRONOUS = 1

# include/spfft/errors.h
ERROR_NAMES = [
    "SPFFT_SUCCESS", "SPFFT_UNKNOWN_ERROR", "SPFFT_INVALID_HANDLE_ERROR", "SPFFT_OVERFLOW_ERROR",
    "SPFFT_ALLOCATION_ERROR", "SPFFT_INVALID_PARAMETER_ERROR", "SPFFT_DUPLICATE_INDICES_ERROR",
    "SPFFT_INVALID_INDICES_ERROR", "SPFFT_MPI_SUPPORT_ERROR", "SPFFT_MPI_ERROR",
    "SPFFT_MPI_PARAMETER_MISMATCH_ERROR", "SPFFT_HOST_EXECUTION_ERROR", "SPFFT_FFTW_ERROR",
    "SPFFT_GPU_ERROR", "SPFFT_GPU_PRECEDING_ERROR", "SPFFT_GPU_SUPPORT_ERROR",
    "SPFFT_GPU_ALLOCATION_ERROR", "SPFFT_GPU_LAUNCH_ERROR", "SPFFT_GPU_NO_DEVICE_ERROR",
    "SPFFT_GPU_INVALID_VALUE_ERROR", "SPFFT_GPU_INVALID_DEVICE_PTR_ERROR", "SPFFT_GPU_COPY_ERROR",
    "SPFFT_GPU_FFT_ERROR",
]
for _i, _n in enumerate(ERROR_NAMES):
    globals()[_n] = _i


class SpfftError(RuntimeError):
    def __init__(self, code: int, where: str = ""):
        self.code = code
        name = ERROR_NAMES[code] if 0 <= code < len(ERROR_NAMES) else f"error {code}"
        super().__init__(f"{name} ({code}) in {where}")


def _ptr(x):
    """Raw address of a numpy array / torch tensor / int / None."""
    if x is None:
        return C.c_void_p(None)
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if isinstance(x, C.c_void_p):
        return x
    raise TypeError(f"cannot take the address of {type(x)}")


class SpfftLib:
    """One loaded shared library exporting the SpFFT C ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found -- build it first (python -c 'import __graft_entry__ as g; "
                f"g.build()'); there is no fallback implementation")
        self.path = path
        # RTLD_LOCAL: the product and the reference host library export the same symbols
        self.lib = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))

    def call(self, name: str, *args):
        fn = getattr(self.lib, name)
        fn.restype = C.c_int
        err = fn(*args)
        if err != 0:
            raise SpfftError(err, name)

    def has(self, name: str) -> bool:
        return hasattr(self.lib, name)


class Grid:
    """spfft::Grid / spfft::GridFloat (include/spfft/grid.hpp:49-203)."""

    def __init__(self, lib: SpfftLib, max_dim_x, max_dim_y, max_dim_z, max_num_local_z_columns,
                 processing_unit=SPFFT_PU_GPU, max_num_threads=-1, single=False):
        self.lib = lib
        self.single = single
        self._sfx = "_float" if single else ""
        self.handle = C.c_void_p()
        lib.call("spfft" + self._sfx + "_grid_create", C.byref(self.handle), int(max_dim_x),
                 int(max_dim_y), int(max_dim_z), int(max_num_local_z_columns),
                 int(processing_unit), int(max_num_threads))

    def _get_int(self, what):
        v = C.c_int()
        self.lib.call(f"spfft{self._sfx}_grid_{what}", self.handle, C.byref(v))
        return v.value

    def max_dim_x(self): return self._get_int("max_dim_x")
    def max_dim_y(self): return self._get_int("max_dim_y")
    def max_dim_z(self): return self._get_int("max_dim_z")
    def max_num_local_z_columns(self): return self._get_int("max_num_local_z_columns")
    def max_local_z_length(self): return self._get_int("max_local_z_length")
    def processing_unit(self): return self._get_int("processing_unit")
    def device_id(self): return self._get_int("device_id")
    def num_threads(self): return self._get_int("num_threads")

    def create_transform(self, processing_unit, transform_type, dim_x, dim_y, dim_z,
                         local_z_length, indices, index_format=SPFFT_INDEX_TRIPLETS,
                         num_local_elements=None):
        return Transform(self.lib, grid=self, processing_unit=processing_unit,
                         transform_type=transform_type, dim_x=dim_x, dim_y=dim_y, dim_z=dim_z,
                         local_z_length=local_z_length, indices=indices,
                         index_format=index_format, single=self.single,
                         num_local_elements=num_local_elements)

    def destroy(self):
        if self.handle:
            self.lib.call("spfft" + self._sfx + "_grid_destroy", self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Transform:
    """spfft::Transform / spfft::TransformFloat (include/spfft/transform.hpp:56-315)."""

    def __init__(self, lib: SpfftLib, grid=None, processing_unit=SPFFT_PU_GPU,
                 transform_type=SPFFT_TRANS_C2C, dim_x=0, dim_y=0, dim_z=0, local_z_length=None,
                 indices=None, index_format=SPFFT_INDEX_TRIPLETS, max_num_threads=-1,
                 single=False, num_local_elements=None, _handle=None):
        self.lib = lib
        self.single = single
        self._sfx = "_float" if single else ""
        self.real_t = np.float32 if single else np.float64
        self.cplx_t = np.complex64 if single else np.complex128
        self.handle = C.c_void_p()
        if _handle is not None:
            self.handle = _handle
            return
        idx = None
        n = 0
        if indices is not None:
            idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int32).reshape(-1))
            n = idx.size // 3
        if num_local_elements is not None:
            n = int(num_local_elements)
        iptr = idx.ctypes.data_as(C.POINTER(C.c_int)) if idx is not None and idx.size else None
        if grid is not None:
            if local_z_length is None:
                local_z_length = dim_z
            lib.call("spfft" + self._sfx + "_transform_create", C.byref(self.handle), grid.handle,
                     int(processing_unit), int(transform_type), int(dim_x), int(dim_y), int(dim_z),
                     int(local_z_length), int(n), int(index_format), iptr)
        else:
            lib.call("spfft" + self._sfx + "_transform_create_independent", C.byref(self.handle),
                     int(max_num_threads), int(processing_unit), int(transform_type), int(dim_x),
                     int(dim_y), int(dim_z), int(n), int(index_format), iptr)

    # ---- getters ----
    def _get_int(self, what):
        v = C.c_int()
        self.lib.call(f"spfft{self._sfx}_transform_{what}", self.handle, C.byref(v))
        return v.value

    def _get_ll(self, what):
        v = C.c_longlong()
        self.lib.call(f"spfft{self._sfx}_transform_{what}", self.handle, C.byref(v))
        return v.value

    def dim_x(self): return self._get_int("dim_x")
    def dim_y(self): return self._get_int("dim_y")
    def dim_z(self): return self._get_int("dim_z")
    def local_z_length(self): return self._get_int("local_z_length")
    def local_z_offset(self): return self._get_int("local_z_offset")
    def local_slice_size(self): return self._get_int("local_slice_size")
    def num_local_elements(self): return self._get_int("num_local_elements")
    def global_size(self): return self._get_ll("global_size")
    def num_global_elements(self): return self._get_ll("num_global_elements")
    def device_id(self): return self._get_int("device_id")
    def num_threads(self): return self._get_int("num_threads")
    def execution_mode(self): return self._get_int("execution_mode")

    def set_execution_mode(self, mode):
        self.lib.call(f"spfft{self._sfx}_transform_set_execution_mode", self.handle, int(mode))

    def clone(self):
        h = C.c_void_p()
        self.lib.call(f"spfft{self._sfx}_transform_clone", self.handle, C.byref(h))
        return Transform(self.lib, single=self.single, _handle=h)

    # ---- data ----
    def space_domain_data(self, location) -> int:
        """Raw address of the internal space-domain buffer at `location`."""
        p = C.c_void_p()
        self.lib.call(f"spfft{self._sfx}_transform_get_space_domain", self.handle, int(location),
                      C.byref(p))
        return p.value or 0

    def space_domain_host_view(self, transform_type) -> np.ndarray:
        """numpy view (z_local, y, x) of the internal HOST space-domain buffer."""
        addr = self.space_domain_data(SPFFT_PU_HOST)
        nz, ny, nx = self.local_z_length(), self.dim_y(), self.dim_x()
        n = nz * ny * nx
        if transform_type == SPFFT_TRANS_R2C:
            buf = (C.c_float if self.single else C.c_double) * n
            return np.frombuffer(buf.from_address(addr), dtype=self.real_t).reshape(nz, ny, nx)
        buf = (C.c_float if self.single else C.c_double) * (2 * n)
        return np.frombuffer(buf.from_address(addr), dtype=self.real_t).view(self.cplx_t).reshape(
            nz, ny, nx)

    def backward(self, input_values, output_location) -> None:
        self.lib.call(f"spfft{self._sfx}_transform_backward", self.handle, _ptr(input_values),
                      int(output_location))

    def backward_ptr(self, input_values, output) -> None:
        self.lib.call(f"spfft{self._sfx}_transform_backward_ptr", self.handle, _ptr(input_values),
                      _ptr(output))

    def forward(self, input_location, output_values, scaling=SPFFT_NO_SCALING) -> None:
        self.lib.call(f"spfft{self._sfx}_transform_forward", self.handle, int(input_location),
                      _ptr(output_values), int(scaling))

    def forward_ptr(self, input_space, output_values, scaling=SPFFT_NO_SCALING) -> None:
        self.lib.call(f"spfft{self._sfx}_transform_forward_ptr", self.handle, _ptr(input_space),
                      _ptr(output_values), int(scaling))

    def destroy(self):
        if self.handle:
            self.lib.call(f"spfft{self._sfx}_transform_destroy", self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def _ptr_array(objs):
    arr = (C.c_void_p * len(objs))()
    for i, o in enumerate(objs):
        arr[i] = _ptr(o)
    return arr


def _int_array(vals):
    arr = (C.c_int * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr


def multi_transform_backward(transforms, inputs, output_locations):
    """spfft_multi_transform_backward (include/spfft/multi_transform.h:74-82)."""
    t0 = transforms[0]
    hs = (C.c_void_p * len(transforms))(*[t.handle for t in transforms])
    t0.lib.call(f"spfft{t0._sfx}_multi_transform_backward", len(transforms), hs,
                _ptr_array(inputs), _int_array(output_locations))


def multi_transform_backward_ptr(transforms, inputs, outputs):
    t0 = transforms[0]
    hs = (C.c_void_p * len(transforms))(*[t.handle for t in transforms])
    t0.lib.call(f"spfft{t0._sfx}_multi_transform_backward_ptr", len(transforms), hs,
                _ptr_array(inputs), _ptr_array(outputs))


def multi_transform_forward(transforms, input_locations, outputs, scalings):
    """spfft_multi_transform_forward (include/spfft/multi_transform.h:49-60)."""
    t0 = transforms[0]
    hs = (C.c_void_p * len(transforms))(*[t.handle for t in transforms])
    t0.lib.call(f"spfft{t0._sfx}_multi_transform_forward", len(transforms), hs,
                _int_array(input_locations), _ptr_array(outputs), _int_array(scalings))


def multi_transform_forward_ptr(transforms, inputs, outputs, scalings):
    t0 = transforms[0]
    hs = (C.c_void_p * len(transforms))(*[t.handle for t in transforms])
    t0.lib.call(f"spfft{t0._sfx}_multi_transform_forward_ptr", len(transforms), hs,
                _ptr_array(inputs), _ptr_array(outputs), _int_array(scalings))


# ----------------------------------------------------------------------------------------------
# Loading the product library
# ----------------------------------------------------------------------------------------------
_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "lib", "libspfft_b200.so")

# every symbol include/spfft/*.h declares (checked at load time, tests/test_abi.py)
_GRID_FUNCS = ["grid_create", "grid_destroy", "grid_max_dim_x", "grid_max_dim_y", "grid_max_dim_z",
               "grid_max_num_local_z_columns", "grid_max_local_z_length", "grid_processing_unit",
               "grid_device_id", "grid_num_threads"]
_TRANSFORM_FUNCS = ["transform_create", "transform_create_independent", "transform_destroy",
                    "transform_clone", "transform_forward", "transform_forward_ptr",
                    "transform_backward", "transform_backward_ptr", "transform_get_space_domain",
                    "transform_dim_x", "transform_dim_y", "transform_dim_z",
                    "transform_local_z_length", "transform_local_slice_size",
                    "transform_local_z_offset", "transform_global_size",
                    "transform_num_local_elements", "transform_num_global_elements",
                    "transform_device_id", "transform_num_threads", "transform_execution_mode",
                    "transform_set_execution_mode"]
_MULTI_FUNCS = ["multi_transform_forward", "multi_transform_forward_ptr",
                "multi_transform_backward", "multi_transform_backward_ptr"]
_EXT_PER_PRECISION = ["transform_index_maps", "transform_stream", "transform_peer_exchange",
                      "transform_set_profiling",
                      "transform_stage_times"]
_EXT_COMMON = ["spfft_b200_convert_index_triplets", "spfft_b200_kernel_launch_count",
               "spfft_b200_nccl_unique_id", "spfft_b200_comm_create", "spfft_b200_comm_destroy",
               "spfft_b200_comm_size", "spfft_b200_comm_rank", "spfft_b200_exchange_plan",
               "spfft_b200_exchange_plan_peer",
               "spfft_grid_create_distributed_nccl", "spfft_float_grid_create_distributed_nccl",
               "spfft_transform_create_independent_distributed_nccl",
               "spfft_float_transform_create_independent_distributed_nccl"]


def exported_symbols():
    """Names of all C entry points the public headers declare."""
    names = []
    for prefix in ("spfft_", "spfft_float_"):
        names += [prefix + f for f in _GRID_FUNCS + _TRANSFORM_FUNCS + _MULTI_FUNCS]
    for prefix in ("spfft_b200_", "spfft_b200_float_"):
        names += [prefix + f for f in _EXT_PER_PRECISION]
    return names + _EXT_COMMON


_LOADED = None


def load(path: str | None = None) -> SpfftLib:
    """dlopen the product library (no fallback: a missing library or symbol is an error)."""
    global _LOADED
    if path is None and _LOADED is not None:
        return _LOADED
    # SPFFT_B200_LIB: experiment builds of the same library (tools/build_variant.py)
    lib = SpfftLib(path or os.environ.get("SPFFT_B200_LIB") or LIB_PATH)
    missing = [s for s in exported_symbols() if not lib.has(s)]
    if missing:
        raise ImportError(f"{lib.path} lacks symbols: {missing}")
    if path is None:
        _LOADED = lib
    return lib


def kernel_launch_count(lib: SpfftLib) -> int:
    v = C.c_longlong()
    lib.call("spfft_b200_kernel_launch_count", C.byref(v))
    return v.value


def convert_index_triplets(lib: SpfftLib, hermitian, dim_x, dim_y, dim_z, triplets):
    """spfft_b200_convert_index_triplets -> (valueIndices, stickIndices)."""
    t = np.ascontiguousarray(np.asarray(triplets, dtype=np.int32).reshape(-1))
    n = t.size // 3
    vi = np.empty(n, dtype=np.int32)
    si = np.empty(max(min(n, dim_x * dim_y), 1), dtype=np.int32)
    ns = C.c_int()
    lib.call("spfft_b200_convert_index_triplets", int(bool(hermitian)), int(dim_x), int(dim_y),
             int(dim_z), int(n), t.ctypes.data_as(C.POINTER(C.c_int)) if n else None,
             vi.ctypes.data_as(C.POINTER(C.c_int)), si.ctypes.data_as(C.POINTER(C.c_int)),
             C.byref(ns))
    return vi, si[:ns.value].copy()


def transform_index_maps(t: Transform):
    """(valueIndices, stickIndices) of an existing transform (copies)."""
    pv = C.POINTER(C.c_int)()
    ps = C.POINTER(C.c_int)()
    nv = C.c_int()
    ns = C.c_int()
    name = "spfft_b200_float_transform_index_maps" if t.single else "spfft_b200_transform_index_maps"
    t.lib.call(name, t.handle, C.byref(pv), C.byref(nv), C.byref(ps), C.byref(ns))
    vi = np.ctypeslib.as_array(pv, shape=(nv.value,)).copy() if nv.value else np.zeros(0, np.int32)
    si = np.ctypeslib.as_array(ps, shape=(ns.value,)).copy() if ns.value else np.zeros(0, np.int32)
    return vi, si


def transform_stream(t: Transform) -> int:
    p = C.c_void_p()
    name = "spfft_b200_float_transform_stream" if t.single else "spfft_b200_transform_stream"
    t.lib.call(name, t.handle, C.byref(p))
    return p.value or 0


def peer_exchange(t: Transform) -> bool:
    """spfft_b200_transform_peer_exchange: exchange fused into the stage kernels over peer memory?"""
    v = C.c_int()
    name = "spfft_b200_float_transform_peer_exchange" if t.single else "spfft_b200_transform_peer_exchange"
    t.lib.call(name, t.handle, C.byref(v))
    return bool(v.value)


def set_profiling(t: Transform, enable: bool) -> None:
    name = "spfft_b200_float_transform_set_profiling" if t.single else "spfft_b200_transform_set_profiling"
    t.lib.call(name, t.handle, int(bool(enable)))


def stage_times(t: Transform):
    """[(stage name, milliseconds)] of the most recent call (profiling must be enabled)."""
    names = (C.c_char_p * 16)()
    ms = (C.c_float * 16)()
    n = C.c_int()
    name = "spfft_b200_float_transform_stage_times" if t.single else "spfft_b200_transform_stage_times"
    t.lib.call(name, t.handle, 16, C.byref(n), names, ms)
    return [(names[i].decode(), float(ms[i])) for i in range(n.value)]


# ----------------------------------------------------------------------------------------------
# Distributed transforms (include/spfft/b200_ext.h): NCCL communicator in place of MPI_Comm
# ----------------------------------------------------------------------------------------------
def nccl_unique_id(lib: SpfftLib) -> bytes:
    """128-byte id created on rank 0; broadcast it to every rank (e.g. torch.distributed)."""
    buf = C.create_string_buffer(128)
    lib.call("spfft_b200_nccl_unique_id", buf)
    return buf.raw


class Comm:
    """SpfftB200Comm: this rank's NCCL communicator on the current CUDA device."""

    def __init__(self, lib: SpfftLib, num_ranks: int, rank: int, unique_id: bytes):
        self.lib = lib
        self.handle = C.c_void_p()
        lib.call("spfft_b200_comm_create", C.byref(self.handle), int(num_ranks), int(rank),
                 C.c_char_p(unique_id))
        self.size, self.rank = int(num_ranks), int(rank)

    def destroy(self):
        if self.handle:
            self.lib.call("spfft_b200_comm_destroy", self.handle)
            self.handle = C.c_void_p()


def comm_from_torch(lib: SpfftLib) -> Comm:
    """Communicator over the ranks of the default torch.distributed process group (the id travels
    through a broadcast of that group; the current CUDA device must already be set)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = [nccl_unique_id(lib) if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    return Comm(lib, world, rank, payload[0])


class DistributedGrid(Grid):
    """spfft_grid_create_distributed_nccl (mirror of spfft_grid_create_distributed, grid.h:84-96)."""

    def __init__(self, lib: SpfftLib, comm: Comm, max_dim_x, max_dim_y, max_dim_z,
                 max_num_local_z_columns, max_local_z_length, processing_unit=SPFFT_PU_GPU,
                 max_num_threads=-1, exchange_type=SPFFT_EXCH_DEFAULT, single=False):
        self.lib = lib
        self.single = single
        self._sfx = "_float" if single else ""
        self.handle = C.c_void_p()
        self.comm = comm
        lib.call("spfft" + self._sfx + "_grid_create_distributed_nccl", C.byref(self.handle),
                 int(max_dim_x), int(max_dim_y), int(max_dim_z), int(max_num_local_z_columns),
                 int(max_local_z_length), int(processing_unit), int(max_num_threads), comm.handle,
                 int(exchange_type))


def distributed_transform(lib: SpfftLib, comm: Comm, transform_type, dim_x, dim_y, dim_z,
                          local_z_length, indices, processing_unit=SPFFT_PU_GPU, max_num_threads=-1,
                          exchange_type=SPFFT_EXCH_DEFAULT, single=False) -> Transform:
    """spfft_transform_create_independent_distributed_nccl (mirror of transform.h:115-128)."""
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int32).reshape(-1))
    n = idx.size // 3
    h = C.c_void_p()
    sfx = "_float" if single else ""
    lib.call(f"spfft{sfx}_transform_create_independent_distributed_nccl", C.byref(h),
             int(max_num_threads), comm.handle, int(exchange_type), int(processing_unit),
             int(transform_type), int(dim_x), int(dim_y), int(dim_z), int(local_z_length), int(n),
             SPFFT_INDEX_TRIPLETS, idx.ctypes.data_as(C.POINTER(C.c_int)) if n else None)
    return Transform(lib, single=single, _handle=h)


def exchange_plan(lib: SpfftLib, transform_type, single, dim_x, dim_y, dim_z, comm_rank,
                  sticks_per_rank, planes_per_rank):
    """spfft_b200_exchange_plan: host-only view of one rank's stick<->slab exchange (dict)."""
    size = len(sticks_per_rank)
    ns = np.array([len(s) for s in sticks_per_rank], dtype=np.int32)
    allsticks = np.ascontiguousarray(np.concatenate([np.asarray(s, np.int32) for s in sticks_per_rank])
                                     if ns.sum() else np.zeros(1, np.int32))
    planes = np.ascontiguousarray(planes_per_rank, dtype=np.int32)
    total = int(ns.sum())
    out = {"pitch": np.zeros(size, np.int32), "stick_offset": np.zeros(size, np.int64),
           "stick_count": np.zeros(size, np.int64), "plane_offset": np.zeros(size, np.int64),
           "plane_count": np.zeros(size, np.int64), "xt_start": np.zeros(dim_x + 2, np.int32),
           "stick_slot": np.zeros(max(total, 1), np.int32), "src_base": np.zeros(max(total, 1), np.int32),
           "src_pitch": np.zeros(max(total, 1), np.int32)}
    nxt, l2vy = C.c_int(), C.c_int()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.call("spfft_b200_exchange_plan", int(transform_type), int(bool(single)), int(dim_x), int(dim_y),
             int(dim_z), size, int(comm_rank), p(ns), p(allsticks), p(planes), p(out["pitch"]),
             p(out["stick_offset"]), p(out["stick_count"]), p(out["plane_offset"]),
             p(out["plane_count"]), C.byref(nxt), C.byref(l2vy), p(out["xt_start"]),
             p(out["stick_slot"]), p(out["src_base"]), p(out["src_pitch"]))
    out["num_x_tiles"], out["log2_vy"] = nxt.value, l2vy.value
    out["xt_start"] = out["xt_start"][:nxt.value + 1]
    for k in ("stick_slot", "src_base", "src_pitch"):
        out[k] = out[k][:total]
    # peer-memory form of the same exchange (where the fused kernels store)
    peer = {"row_rank": np.zeros(dim_z, np.int32), "row_off": np.zeros(dim_z, np.int64),
            "stick_rank": np.zeros(max(total, 1), np.int32), "fwd_base": np.zeros(max(total, 1), np.int32)}
    rot = C.c_int()
    lib.call("spfft_b200_exchange_plan_peer", int(transform_type), int(bool(single)), int(dim_x), int(dim_y),
             int(dim_z), size, int(comm_rank), p(ns), p(allsticks), p(planes), p(peer["row_rank"]),
             p(peer["row_off"]), p(peer["stick_rank"]), p(peer["fwd_base"]), C.byref(rot))
    out["row_rank"], out["row_off"] = peer["row_rank"], peer["row_off"]
    out["stick_rank"], out["fwd_base"] = peer["stick_rank"][:total], peer["fwd_base"][:total]
    out["fwd_tile_rotate"] = rot.value
    return out
