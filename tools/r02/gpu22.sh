# round 2, GPU session 22: SPFFT_PU_HOST served on the device (reference examples run), concurrent warp-FFT transforms
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "concurrent_streams or processing_unit or empty_and_degenerate or reference_examples_run or grid_shared" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.log
tail -15 gpurun_out/pytest_new.log
tests/callers/_build/ref_example_cpp | tail -12
