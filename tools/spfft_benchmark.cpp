// spfft_benchmark.cpp -- command-line benchmark with the interface of the reference's
// tests/programs/benchmark.cpp (flags -d -r -o -m -s -t -e -p, JSON file with a "parameters" and a
// "timings" object), written from scratch against the public C++ API (spfft/spfft.hpp) only.
//
//   spfft_benchmark -d 256 256 256 -r 20 -o out.json -p gpu-gpu -e compact [-m 4] [-s 0.5] [-t r2c]
//
// Workload (same definition as the reference program, benchmark.cpp:171-199): all z of the sticks
// (x, y) with x < dimXFreq * sparsity (R2C: y < dimY/2+1 on x = 0), values zero-initialised; one
// warm-up pair, then -r repeats of backward + forward (SPFFT_NO_SCALING), -m transforms through
// multi_transform_*. Differences: single process (no MPI in this build; -e is recorded only),
// -p cpu is rejected (this library has no host execution path), timings are CUDA-synchronised
// wall-clock totals instead of an rt_graph tree.
//
// Build: g++ -std=c++17 -Iinclude -I$CUDA_HOME/include tools/spfft_benchmark.cpp -Lspfft_b200/lib
//        -lspfft_b200 -L$CUDA_HOME/lib64 -lcudart   (tests/test_dropin_build.py does it)
#include <cuda_runtime.h>

#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "spfft/spfft.hpp"

namespace {

struct Options {
  int dims[3] = {0, 0, 0};
  int repeats = 1;
  int transforms = 1;
  std::string output;
  std::string exchange = "compact";
  std::string proc;
  std::string type = "c2c";
  double sparsity = 1.0;
};

[[noreturn]] void usage(const char* why) {
  std::fprintf(stderr,
               "%s\nusage: spfft_benchmark -d X Y Z -r REPEATS -o FILE -p {gpu|gpu-gpu} "
               "-e {all|compact|compactFloat|buffered|bufferedFloat|unbuffered} [-m TRANSFORMS] [-s SPARSITY] "
               "[-t {c2c|r2c}]\n",
               why);
  std::exit(2);
}

Options parse(int argc, char** argv) {
  Options o;
  bool haveD = false, haveR = false, haveO = false, haveP = false, haveE = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* {
      if (i + 1 >= argc) usage(("missing value after " + a).c_str());
      return argv[++i];
    };
    if (a == "-d") {
      for (int k = 0; k < 3; ++k) o.dims[k] = std::atoi(next());
      haveD = true;
    } else if (a == "-r") {
      o.repeats = std::atoi(next());
      haveR = true;
    } else if (a == "-o") {
      o.output = next();
      haveO = true;
    } else if (a == "-m") {
      o.transforms = std::atoi(next());
    } else if (a == "-s") {
      o.sparsity = std::atof(next());
    } else if (a == "-t") {
      o.type = next();
    } else if (a == "-e") {
      o.exchange = next();
      haveE = true;
    } else if (a == "-p") {
      o.proc = next();
      haveP = true;
    } else {
      usage(("unknown option " + a).c_str());
    }
  }
  if (!haveD || !haveR || !haveO || !haveP || !haveE) usage("-d, -r, -o, -e and -p are required");
  if (o.dims[0] <= 0 || o.dims[1] <= 0 || o.dims[2] <= 0 || o.repeats < 1 || o.transforms < 1)
    usage("dimensions, repeats and transform count must be positive");
  if (o.type != "c2c" && o.type != "r2c") usage("-t must be c2c or r2c");
  if (o.proc == "cpu") usage("-p cpu: this library has no host execution path (SPFFT_PU_GPU only)");
  if (o.proc != "gpu" && o.proc != "gpu-gpu") usage("-p must be gpu or gpu-gpu");
  return o;
}

void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    std::fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e));
    std::exit(3);
  }
}

}  // namespace

int main(int argc, char** argv) {
  const Options o = parse(argc, argv);
  const int nx = o.dims[0], ny = o.dims[1], nz = o.dims[2];
  const bool r2c = o.type == "r2c";
  const SpfftTransformType type = r2c ? SPFFT_TRANS_R2C : SPFFT_TRANS_C2C;
  const int nxFreq = r2c ? nx / 2 + 1 : nx;
  const int nyFreq = r2c ? ny / 2 + 1 : ny;

  // index triplets: whole z-sticks for the first `sparsity` share of the x range
  std::vector<int> triplets;
  int numSticks = 0;
  for (int x = 0; x < nxFreq * o.sparsity; ++x) {
    const int yEnd = (x == 0) ? nyFreq : ny;
    for (int y = 0; y < yEnd; ++y, ++numSticks) {
      for (int z = 0; z < nz; ++z) {
        triplets.push_back(x);
        triplets.push_back(y);
        triplets.push_back(z);
      }
    }
  }
  const int numElements = static_cast<int>(triplets.size() / 3);
  const bool dataOnGpu = o.proc == "gpu-gpu";
  const SpfftProcessingUnitType target = dataOnGpu ? SPFFT_PU_GPU : SPFFT_PU_HOST;

  std::cout << "Num MPI ranks: 1\nGrid size: " << nx << ", " << ny << ", " << nz << "\nTransform type: " << o.type
            << "\nSparsity: " << o.sparsity << "\nProc: " << o.proc << "\nGPU Direct: Disabled" << std::endl;

  double seconds = 0.0;
  try {
    // one Grid per transform: transforms of one Grid share work buffers and cannot run together
    std::vector<spfft::Transform> transforms;
    for (int t = 0; t < o.transforms; ++t) {
      spfft::Grid grid(nx, ny, nz, numSticks, SPFFT_PU_GPU, -1);
      if (t == 0)
        transforms.push_back(grid.create_transform(SPFFT_PU_GPU, type, nx, ny, nz, nz, numElements,
                                                   SPFFT_INDEX_TRIPLETS, triplets.data()));
      else
        transforms.push_back(transforms.front().clone());
    }
    // frequency values: pinned host memory or device memory (gpu-gpu)
    std::vector<double*> values(o.transforms, nullptr);
    const size_t bytes = sizeof(double) * 2 * static_cast<size_t>(numElements);
    for (int t = 0; t < o.transforms; ++t) {
      if (dataOnGpu) {
        check(cudaMalloc(reinterpret_cast<void**>(&values[t]), bytes), "cudaMalloc");
        check(cudaMemset(values[t], 0, bytes), "cudaMemset");
      } else {
        check(cudaMallocHost(reinterpret_cast<void**>(&values[t]), bytes), "cudaMallocHost");
        std::memset(values[t], 0, bytes);
      }
    }
    std::vector<SpfftProcessingUnitType> targets(o.transforms, target);
    std::vector<SpfftScalingType> scalings(o.transforms, SPFFT_NO_SCALING);
    auto pair = [&]() {
      if (o.transforms == 1) {
        transforms.front().backward(values[0], target);
        transforms.front().forward(target, values[0], SPFFT_NO_SCALING);
      } else {
        spfft::multi_transform_backward(o.transforms, transforms.data(), values.data(), targets.data());
        spfft::multi_transform_forward(o.transforms, transforms.data(), targets.data(), values.data(),
                                       scalings.data());
      }
    };
    pair();  // warm-up
    check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < o.repeats; ++r) pair();
    check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (double* p : values) {
      if (dataOnGpu)
        cudaFree(p);
      else
        cudaFreeHost(p);
    }
  } catch (const spfft::GenericError& e) {
    std::fprintf(stderr, "SpFFT error %d: %s\n", static_cast<int>(e.error_code()), e.what());
    return 1;
  }

  const double pairs = static_cast<double>(o.repeats) * o.transforms;
  std::cout << "Total: " << seconds << " s for " << pairs << " backward+forward pairs (" << pairs / seconds
            << " pairs/s)" << std::endl;

  const std::time_t now = std::time(nullptr);
  std::string stamp(std::ctime(&now));
  if (!stamp.empty() && stamp.back() == '\n') stamp.pop_back();
  std::ofstream file(o.output);
  file << "{\n  \"parameters\": {\n"
       << "    \"proc\": \"" << o.proc << "\",\n"
       << "    \"data_on_gpu\": " << (dataOnGpu ? "true" : "false") << ",\n"
       << "    \"gpu_direct\": false,\n"
       << "    \"num_ranks\": 1,\n"
       << "    \"num_threads\": 1,\n"
       << "    \"dim_x\": " << nx << ",\n    \"dim_y\": " << ny << ",\n    \"dim_z\": " << nz << ",\n"
       << "    \"exchange_type\": \"" << o.exchange << "\",\n"
       << "    \"num_repeats\": " << o.repeats << ",\n"
       << "    \"num_transforms\": " << o.transforms << ",\n"
       << "    \"sparsity\": " << o.sparsity << ",\n"
       << "    \"transform_type\": \"" << o.type << "\",\n"
       << "    \"time\": \"" << stamp << "\"\n  },\n"
       << "  \"timings\": {\n"
       << "    \"total_s\": " << seconds << ",\n"
       << "    \"pair_ms\": " << 1e3 * seconds / pairs << ",\n"
       << "    \"pairs_per_s\": " << pairs / seconds << ",\n"
       << "    \"num_elements\": " << numElements << ",\n"
       << "    \"num_sticks\": " << numSticks << "\n  }\n}\n";
  return 0;
}
