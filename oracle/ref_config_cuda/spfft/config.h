/* Hand-written stand-in for the CMake-generated spfft/config.h (reference template:
 * include/spfft/config.h.in). Used ONLY by oracle/Makefile to compile the unmodified reference
 * sources WITH its CUDA backend (ExecutionGPU + cuFFT) into oracle/_ref/libspfft_ref_cuda.so: the
 * on-box GPU comparator of bench.py --impl reference-gpu. OpenMP on, CUDA on, MPI off. */
#ifndef SPFFT_CONFIG_H
#define SPFFT_CONFIG_H
#define SPFFT_CUDA
#define SPFFT_OMP
#define SPFFT_SINGLE_PRECISION
#include "spfft/spfft_export.h"
#endif
