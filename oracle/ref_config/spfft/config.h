/* Hand-written stand-in for the CMake-generated spfft/config.h (reference template:
 * include/spfft/config.h.in). Used ONLY by oracle/Makefile to compile the unmodified reference
 * host sources into oracle/_ref/libspfft_ref.so. Host-only build: OpenMP on, MPI/CUDA off. */
#ifndef SPFFT_CONFIG_H
#define SPFFT_CONFIG_H
#define SPFFT_OMP
#define SPFFT_SINGLE_PRECISION
#include "spfft/spfft_export.h"
#endif
