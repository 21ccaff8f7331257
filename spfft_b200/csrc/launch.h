/* launch.h -- the thin internal C ABI between the host C++ (transform_engine.cpp) and the CUDA
 * translation unit (kernels.cu). POD argument structs (stage_kernels.hpp), int error return
 * (a cudaError_t value, 0 = success). `stream` is a cudaStream_t. */
#pragma once
#include "fast_stage_kernels.hpp"
#include "stage_kernels.hpp"

extern "C" {
/* z stage: forward == 0 -> decompress + symmetry + inverse z-FFT; else z-FFT + compress. */
int sb_launch_z_f64(int forward, const sb::ZArgs<double>* args, void* stream);
int sb_launch_z_f32(int forward, const sb::ZArgs<float>* args, void* stream);
/* y stage on every (x tile, local plane). */
int sb_launch_y_f64(int forward, const sb::YArgs<double>* args, void* stream);
int sb_launch_y_f32(int forward, const sb::YArgs<float>* args, void* stream);
/* x stage on every (row tile, local plane). */
int sb_launch_x_f64(int forward, const sb::XArgs<double>* args, void* stream);
int sb_launch_x_f32(int forward, const sb::XArgs<float>* args, void* stream);
/* Barrier over the ranks of a distributed transform through peer-mapped flag arrays
 * (flags[r] = rank r's array of numRanks ints, zero-initialised; epoch increases by one per call). */
int sb_launch_peer_barrier(int* const* flags, int numRanks, int me, int epoch, void* stream);
/* Warp-FFT kernels (wfft_xy.cu, wfft_z.cu: one warp per transform -- two in single precision --, TMA-staged tiles,
 * transform length 512). sb_wxy_config: plan-time query of the fused xy stage (C2C, dimX == dimY == 512,
 * local slab): scratch ring (planes), item lag and number of int counters the kernel needs (which must be zero at
 * launch: the launcher enqueues the memset); sb_wz_available: z stage (dimZ == 512, values in stick order). */
int sb_wxy_config(int isFloat, int n, int numPlanes, int* ring, int* lag, int* numCounters);
int sb_launch_wxy_f64(int forward, const sb::XYArgs<double>* args, void* stream);
int sb_launch_wxy_f32(int forward, const sb::XYArgs<float>* args, void* stream); /* local transforms only */
int sb_wz_available(int isFloat, int nz);
int sb_launch_wz_f64(int forward, const sb::ZArgs<double>* args, void* stream);
int sb_launch_wz_f32(int forward, const sb::ZArgs<float>* args, void* stream); /* C2C only (symTile < 0) */
/* Batched multi-transform (band_kernels.cu): one launch per stage over `numBands` <= sb::kMaxBands
 * transforms that share the plan in `args`; the table holds the per-band data pointers.
 * sb_band_kernel_available: does a batched kernel exist for an axis of length n (registerFft: the
 * axis runs the register-FFT kernels, i.e. its ftw table is set)? */
int sb_band_kernel_available(int n, int registerFft);
int sb_launch_z_bands_f64(int forward, const sb::ZArgs<double>* args, const sb::BandTable<double>* table, int numBands, void* stream);
int sb_launch_z_bands_f32(int forward, const sb::ZArgs<float>* args, const sb::BandTable<float>* table, int numBands, void* stream);
int sb_launch_y_bands_f64(int forward, const sb::YArgs<double>* args, const sb::BandTable<double>* table, int numBands, void* stream);
int sb_launch_y_bands_f32(int forward, const sb::YArgs<float>* args, const sb::BandTable<float>* table, int numBands, void* stream);
int sb_launch_x_bands_f64(int forward, const sb::XArgs<double>* args, const sb::BandTable<double>* table, int numBands, void* stream);
int sb_launch_x_bands_f32(int forward, const sb::XArgs<float>* args, const sb::BandTable<float>* table, int numBands, void* stream);
/* total number of kernel launches issued through this file */
long long sb_launch_count(void);
void sb_note_launches(int n);
/* opt-in limit of dynamic shared memory per block on the current device (bytes) */
int sb_max_dynamic_smem(long long* bytes);
}
