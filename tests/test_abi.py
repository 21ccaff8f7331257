"""The drop-in boundary on a machine without a GPU: the library loads, exports every symbol the
public headers declare, the headers compile as C and as C++, and handle / parameter validation
that needs no device behaves like the reference (src/spfft/transform.cpp:246-248 etc.)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from spfft_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared_symbols():
    """Every SPFFT_FN(...) / spfft_b200_* symbol the headers declare, expanded per precision."""
    names = set()
    detail = os.path.join(INC, "spfft", "detail")
    for inc in ("grid_api.inc", "transform_api.inc", "multi_transform_api.inc"):
        for m in re.finditer(r"SPFFT_FN\((\w+)\)\(", open(os.path.join(detail, inc)).read()):
            names.add("spfft_" + m.group(1))
            names.add("spfft_float_" + m.group(1))
    for m in re.finditer(r"SPFFT_EXPORT\s+SpfftError\s+(spfft_\w+)\s*\(", open(os.path.join(INC, "spfft", "b200_ext.h")).read()):
        names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol(lib):
    declared = _declared_symbols()
    assert len(declared) > 80
    missing = [s for s in declared if not lib.has(s)]
    assert not missing, missing
    # and the python binding's own list is the same set
    assert set(capi.exported_symbols()) <= set(declared)


def test_only_api_symbols_are_exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib.path], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    stray = [n for n in names if not (n.startswith("spfft_") or n.startswith("_ZN5spfft") or n.startswith("_ZNK5spfft"))]
    assert not stray, stray[:10]


def test_headers_compile_as_c_and_cpp(tmp_path):
    c_src = tmp_path / "t.c"
    c_src.write_text('#include "spfft/spfft.h"\nint main(void){ SpfftGrid g = 0; SpfftTransform t = 0; '
                     'SpfftFloatTransform ft = 0; (void)g; (void)t; (void)ft; '
                     'return (int)SPFFT_SUCCESS + (int)SPFFT_PU_GPU - 2; }\n')
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", INC, "-c", str(c_src), "-o",
                    str(tmp_path / "t.o")], check=True)
    cpp_src = tmp_path / "t.cpp"
    cpp_src.write_text('#include "spfft/spfft.hpp"\n#include "spfft/spfft.h"\n'
                       'int f(spfft::Transform& t){ return t.dim_x() + (int)sizeof(spfft::GridFloat); }\n')
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I", INC, "-c", str(cpp_src), "-o",
                    str(tmp_path / "tpp.o")], check=True)


def test_enum_values_match_reference_abi():
    # include/spfft/types.h:33-117, errors.h:37-125 of the reference
    assert (capi.SPFFT_PU_HOST, capi.SPFFT_PU_GPU) == (1, 2)
    assert (capi.SPFFT_TRANS_C2C, capi.SPFFT_TRANS_R2C) == (0, 1)
    assert capi.SPFFT_EXCH_COMPACT_BUFFERED == 3 and capi.SPFFT_EXCH_UNBUFFERED == 5
    assert capi.SPFFT_INVALID_HANDLE_ERROR == 2 and capi.SPFFT_INVALID_PARAMETER_ERROR == 5
    assert capi.SPFFT_GPU_PRECEDING_ERROR == 14 and capi.SPFFT_GPU_FFT_ERROR == 22
    hdr = open(os.path.join(INC, "spfft", "errors.h")).read()
    order = re.findall(r"^\s*(SPFFT_\w+),?\s*/\*", hdr, flags=re.M)
    assert order == capi.ERROR_NAMES


def test_null_handles_are_rejected(lib):
    null = C.c_void_p(None)
    out = C.c_int()
    for sfx in ("", "_float"):
        for fn in ("grid_max_dim_x", "grid_device_id", "transform_dim_x", "transform_local_z_length",
                   "transform_num_local_elements", "transform_execution_mode"):
            f = getattr(lib.lib, f"spfft{sfx}_{fn}")
            f.restype = C.c_int
            assert f(null, C.byref(out)) == capi.SPFFT_INVALID_HANDLE_ERROR
        for fn in ("grid_destroy", "transform_destroy"):
            f = getattr(lib.lib, f"spfft{sfx}_{fn}")
            f.restype = C.c_int
            assert f(null) == capi.SPFFT_INVALID_HANDLE_ERROR
        f = getattr(lib.lib, f"spfft{sfx}_transform_backward_ptr")
        f.restype = C.c_int
        assert f(null, null, null) == capi.SPFFT_INVALID_HANDLE_ERROR
        f = getattr(lib.lib, f"spfft{sfx}_transform_create")
        f.restype = C.c_int
        h = C.c_void_p()
        assert f(C.byref(h), null, 2, 0, 4, 4, 4, 4, 0, 0, null) == capi.SPFFT_INVALID_HANDLE_ERROR


def test_parameter_validation_without_device(lib):
    """Checks that fail before any CUDA call (transform.cpp:50-53, grid_internal.cpp:60-67)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    for args in [(-1, 4, 4), (4, 4, -1)]:
        with pytest.raises(capi.SpfftError) as e:
            capi.Transform(lib, transform_type=0, dim_x=args[0], dim_y=args[1], dim_z=args[2],
                           indices=np.zeros((1, 3), np.int32))
        assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    with pytest.raises(capi.SpfftError) as e:  # null indices with numLocalElements > 0
        capi.Transform(lib, transform_type=0, dim_x=4, dim_y=4, dim_z=4, indices=None, num_local_elements=3)
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    with pytest.raises(capi.SpfftError) as e:
        capi.Grid(lib, 4, 4, 0, 4, capi.SPFFT_PU_GPU, 1)
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    with pytest.raises(capi.SpfftError) as e:  # SPFFT_PU_HOST grids are served by the device kernels: without a
        capi.Grid(lib, 4, 4, 4, 4, capi.SPFFT_PU_HOST, 1)  # device the call fails loudly (there is no CPU path)
    assert e.value.code in (capi.SPFFT_GPU_ERROR, capi.SPFFT_GPU_SUPPORT_ERROR, capi.SPFFT_GPU_ALLOCATION_ERROR)
    with pytest.raises(capi.SpfftError) as e:  # neither unit bit
        capi.Grid(lib, 4, 4, 4, 4, 0, 1)
    assert e.value.code == capi.SPFFT_INVALID_PARAMETER_ERROR
    # a valid request fails LOUDLY without a device -- there is no CPU fallback
    with pytest.raises(capi.SpfftError) as e:
        capi.Grid(lib, 4, 4, 4, 4, capi.SPFFT_PU_GPU, 1)
    assert e.value.code in (capi.SPFFT_GPU_ERROR, capi.SPFFT_GPU_NO_DEVICE_ERROR, capi.SPFFT_GPU_SUPPORT_ERROR)


def test_product_does_not_reference_the_oracle():
    """Nothing under spfft_b200/ (python or C++) may import, include or link the oracle."""
    pkg = os.path.join(ROOT, "spfft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh", ".h", ".inc")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle/|libspfft_ref", text, flags=re.M), f
    out = subprocess.run(["ldd", os.path.join(pkg, "lib", "libspfft_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "spfft_ref" not in out and "fftw" not in out.lower()


def test_fortran_module_matches_the_c_api(built):
    """include/spfft/spfft.f90 (generated by tools/gen_fortran_module.py; no Fortran compiler in the image):
    up to date with the generator, every bind(C) interface names a symbol the library exports, and its
    dummy-argument list is exactly the argument list of the C declaration."""
    import importlib.util
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_fortran_module", os.path.join(root, "tools", "gen_fortran_module.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text, count = gen.generate()
    committed = open(os.path.join(root, "include", "spfft", "spfft.f90")).read()
    assert committed == text, "run python tools/gen_fortran_module.py"
    lib = C.CDLL(built.LIB)
    joined = re.sub(r"&\s*\n\s*", "", committed)
    found = re.findall(r"integer\(c_int\) function (\w+)\(([^)]*)\) bind\(C\)", joined)
    assert len(found) == count == 72
    protos = {name: [a for _, a in args] for name, args in gen.prototypes()}
    for name, args in found:
        assert hasattr(lib, name), name
        base = name.replace("spfft_float_", "").replace("spfft_", "", 1) if name.startswith("spfft_float_") else name[len("spfft_"):]
        assert [a.strip() for a in args.split(",")] == protos[base], name
    # structure: balanced blocks, no line beyond the free-form limit
    assert committed.count("end function") == count and committed.count("\ninterface\n") == 1
    assert max(len(l) for l in committed.splitlines()) <= 132


def test_trace_variant_of_the_fused_xy_kernels_compiles():
    """The experiment build with phase timers (-DSB_WTRACE, tools/build_variant.py + tools/r02/wtrace.py) must keep
    compiling for every instantiation of the fused xy kernels, single precision included (front end + PTX only)."""
    import __graft_entry__ as g
    if not os.path.exists(g.NVCC):
        pytest.skip("no nvcc")
    src = os.path.join(g.CSRC, "wfft_xy.cu")
    cmd = [g.NVCC, "-DSB_WTRACE", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--expt-relaxed-constexpr",
           "-ccbin", g.HOST_CXX] + g.INCLUDES + ["-ptx", "-o", os.devnull, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]


def test_default_512_kernels_use_tma_and_packed_fp32(built):
    """SASS of the warp-FFT objects that go into the library (cuobjdump, no GPU needed): the double- and single-
    precision stage kernels of the headline shape move their tiles with bulk tensor copies / bulk copies, prefetch
    the inverse maps with LDGSTS, exchange across lanes with SHFL, and the single-precision instantiations compute
    on the packed fp32 pipe (FADD2 / FMUL2 / FFMA2) -- the evidence summarised in profiles/r02_sass_summary.txt."""
    import shutil
    import __graft_entry__ as g
    cuobjdump = os.path.join(g.CUDA_HOME, "bin", "cuobjdump")
    if not os.path.exists(cuobjdump) or shutil.which("c++filt") is None:
        pytest.skip("no cuobjdump / c++filt")
    ops = {}
    for obj in ("wfft_xy.cu.o", "wfft_z.cu.o"):
        out = subprocess.run([cuobjdump, "-sass", os.path.join(g.OBJDIR, obj)], capture_output=True, text=True, check=True).stdout
        name = None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.split("(")[0].replace("void sb::", "")
                ops[name] = set()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
            if m and name:
                ops[name].add(m.group(1))
    for prec in ("double", "float"):
        assert {"UTMASTG", "LDGSTS", "SHFL"} <= ops[f"k_wz_bwd<{prec}>"]
        assert {"UTMALDG", "SHFL"} <= ops[f"k_wz_fwd<{prec}>"]
        for W in (2, 4, 8):
            assert {"UTMASTG", "LDGSTS", "SHFL"} <= ops[f"k_wxy_bwd<{prec}, {W}>"]
            assert {"UTMALDG", "UBLKCP", "LDGSTS", "SHFL"} <= ops[f"k_wxy_fwd<{prec}, {W}>"]
    for k, v in ops.items():
        if "<float" in k:
            assert {"FADD2", "FMUL2", "FFMA2"} <= v and "DFMA" not in v, k
        else:
            assert "DFMA" in v, k
