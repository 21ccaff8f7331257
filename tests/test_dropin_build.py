"""Drop-in boundary as a C / C++ user sees it: consumers compile against include/spfft and link with
spfft_b200/lib/libspfft_b200.so using nothing but gcc / g++.

* our own consumers tests/callers/caller_c.c (C99) and caller_cpp.cpp (C++17): compiled and linked on
  the CPU, executed on the GPU box (checked against a direct DFT inside the programs);
* the reference's own example programs, compiled from where they lie under /root/reference (never
  copied), against OUR headers and library: proves that unmodified user code builds (skipped on the
  GPU box, where /root/reference does not exist). They select SPFFT_PU_HOST, which this library serves with
  its device kernels and host staging (there is no CPU path), so the binaries built here also RUN on the GPU box."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "spfft_b200", "lib")
OUT = os.path.join(ROOT, "tests", "callers", "_build")
GCC, GXX = "/usr/bin/gcc", "/usr/bin/g++"


def _build(compiler, std, src, exe):
    os.makedirs(OUT, exist_ok=True)
    cmd = [compiler, std, "-O1", "-Wall", "-I" + INC, src, "-L" + LIBDIR, "-lspfft_b200", "-lm",
           "-Wl,-rpath," + LIBDIR, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, " ".join(cmd) + "\n" + res.stdout + res.stderr
    return exe


@pytest.fixture(scope="module")
def callers(built):
    return {"c": _build(GCC, "-std=c99", os.path.join(ROOT, "tests", "callers", "caller_c.c"), os.path.join(OUT, "caller_c")),
            "cpp": _build(GXX, "-std=c++17", os.path.join(ROOT, "tests", "callers", "caller_cpp.cpp"), os.path.join(OUT, "caller_cpp"))}


def test_own_consumers_compile_and_link(callers):
    for exe in callers.values():
        assert os.path.exists(exe)


@pytest.mark.parametrize("name,compiler,std", [("example.c", GCC, "-std=c99"), ("example.cpp", GXX, "-std=c++17")])
def test_reference_examples_build_unmodified(built, name, compiler, std):
    src = os.path.join("/root/reference/examples", name)
    if not os.path.exists(src):
        pytest.skip("/root/reference is not present on this machine")
    _build(compiler, std, src, os.path.join(OUT, "ref_" + name.replace(".", "_")))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["example.c", "example.cpp"])
def test_reference_examples_run(name):
    """The reference's unmodified example programs (built on the CPU box by the test above, they travel with the
    snapshot) run against this library: they create SPFFT_PU_HOST grids / transforms and print the space domain."""
    exe = os.path.join(OUT, "ref_" + name.replace(".", "_"))
    if not os.path.exists(exe):
        pytest.skip("reference example not built (needs /root/reference at build time)")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert len(res.stdout.splitlines()) > 4


@pytest.fixture(scope="module")
def benchmark_exe(built):
    """tools/spfft_benchmark.cpp: the reference's benchmark CLI (tests/programs/benchmark.cpp) on the public C++ API."""
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "spfft_benchmark")
    cmd = [GXX, "-std=c++17", "-O2", "-Wall", "-I" + INC, "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "tools", "spfft_benchmark.cpp"), "-L" + LIBDIR, "-lspfft_b200",
           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + LIBDIR,
           "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, " ".join(cmd) + "\n" + res.stdout + res.stderr
    return exe


def test_benchmark_program_builds_and_validates_arguments(benchmark_exe):
    res = subprocess.run([benchmark_exe, "-d", "8", "8", "8", "-r", "1", "-o", "/dev/null", "-e", "compact", "-p", "cpu"],
                         capture_output=True, text=True)
    assert res.returncode == 2 and "no host execution path" in res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["-d", "64", "64", "64", "-r", "3", "-p", "gpu-gpu", "-e", "compact"],
                                  ["-d", "48", "32", "20", "-r", "2", "-p", "gpu", "-e", "compact", "-t", "r2c", "-s", "0.5"],
                                  ["-d", "32", "32", "32", "-r", "2", "-p", "gpu-gpu", "-e", "compact", "-m", "3"]],
                         ids=["c2c-device", "r2c-host-sparse", "multi"])
def test_benchmark_program_runs(benchmark_exe, tmp_path, args):
    import json
    out = tmp_path / "bench.json"
    res = subprocess.run([benchmark_exe] + args + ["-o", str(out)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    j = json.loads(out.read_text())
    assert j["parameters"]["dim_x"] == int(args[1]) and j["parameters"]["num_repeats"] == int(args[5])
    assert j["timings"]["pairs_per_s"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["c", "cpp"])
def test_own_consumers_run(callers, which):
    res = subprocess.run([callers[which]], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "PASS" in res.stdout


def test_pkg_config_file(built):
    """spfft_b200.pc (written by build()) gives a consumer its flags without knowing the tree."""
    if not shutil.which("pkg-config"):
        pytest.skip("pkg-config not installed")
    env = dict(os.environ, PKG_CONFIG_PATH=os.path.join(LIBDIR, "pkgconfig"))
    cflags = subprocess.run(["pkg-config", "--cflags", "spfft_b200"], capture_output=True, text=True, env=env)
    libs = subprocess.run(["pkg-config", "--libs", "spfft_b200"], capture_output=True, text=True, env=env)
    assert cflags.returncode == 0 and libs.returncode == 0, cflags.stderr + libs.stderr
    assert "-I" + INC in cflags.stdout and "-lspfft_b200" in libs.stdout
