/* spfft/b200_ext.h -- entry points this build adds to the SpFFT C ABI.
 *
 * They cover what the reference expresses through MPI (absent here: the distributed transform is
 * sharded over the GPUs of one NVSwitch box and exchanges over NCCL / NVLink peer memory), plus
 * inspection hooks the parity tests and the benchmark need. Plain pointers and sizes only.
 * Everything returns an SpfftError (spfft/errors.h).
 */
#ifndef SPFFT_B200_EXT_H
#define SPFFT_B200_EXT_H
#include "spfft/config.h"
#include "spfft/errors.h"
#include "spfft/grid.h"
#include "spfft/grid_float.h"
#include "spfft/transform.h"
#include "spfft/transform_float.h"
#include "spfft/types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- distributed transforms over the GPUs of one NVSwitch box (NCCL instead of MPI) ----------
 *
 * One process per GPU. Rank 0 obtains an id, the application broadcasts its 128 bytes by any means
 * (torch.distributed, MPI, a file), every rank creates its communicator on its current device, then
 * a distributed grid; transforms created from that grid are distributed: frequency domain = the
 * z-sticks this rank passed, space domain = `localZLength` consecutive xy planes, in rank order.
 * These entry points mirror spfft_grid_create_distributed (include/spfft/grid.h:84-96) and
 * spfft_transform_create_independent_distributed (transform.h:115-128) with SpfftB200Comm in place
 * of MPI_Comm; everything else (spfft_transform_create / backward / forward / getters) is the
 * unchanged SpFFT API. The exchange is one grouped ncclSend/ncclRecv per direction on the
 * transform's stream. */
SPFFT_EXPORT SpfftError spfft_b200_nccl_unique_id(char* id128);
SPFFT_EXPORT SpfftError spfft_b200_comm_create(SpfftB200Comm* comm, int numRanks, int rank,
                                               const char* id128);
SPFFT_EXPORT SpfftError spfft_b200_comm_destroy(SpfftB200Comm comm);
SPFFT_EXPORT SpfftError spfft_b200_comm_size(SpfftB200Comm comm, int* size);
SPFFT_EXPORT SpfftError spfft_b200_comm_rank(SpfftB200Comm comm, int* rank);

SPFFT_EXPORT SpfftError spfft_grid_create_distributed_nccl(
    SpfftGrid* grid, int maxDimX, int maxDimY, int maxDimZ, int maxNumLocalZColumns,
    int maxLocalZLength, SpfftProcessingUnitType processingUnit, int maxNumThreads,
    SpfftB200Comm comm, SpfftExchangeType exchangeType);
SPFFT_EXPORT SpfftError spfft_float_grid_create_distributed_nccl(
    SpfftFloatGrid* grid, int maxDimX, int maxDimY, int maxDimZ, int maxNumLocalZColumns,
    int maxLocalZLength, SpfftProcessingUnitType processingUnit, int maxNumThreads,
    SpfftB200Comm comm, SpfftExchangeType exchangeType);
SPFFT_EXPORT SpfftError spfft_transform_create_independent_distributed_nccl(
    SpfftTransform* transform, int maxNumThreads, SpfftB200Comm comm,
    SpfftExchangeType exchangeType, SpfftProcessingUnitType processingUnit,
    SpfftTransformType transformType, int dimX, int dimY, int dimZ, int localZLength,
    int numLocalElements, SpfftIndexFormatType indexFormat, const int* indices);
SPFFT_EXPORT SpfftError spfft_float_transform_create_independent_distributed_nccl(
    SpfftFloatTransform* transform, int maxNumThreads, SpfftB200Comm comm,
    SpfftExchangeType exchangeType, SpfftProcessingUnitType processingUnit,
    SpfftTransformType transformType, int dimX, int dimY, int dimZ, int localZLength,
    int numLocalElements, SpfftIndexFormatType indexFormat, const int* indices);

/* Whether a distributed transform exchanges through peer-mapped memory (CUDA IPC over NVLink /
 * NVSwitch): the z stage (backward) and the y stage (forward) then store their results straight
 * into the destination rank's buffer and a flag barrier through the same mapped memory orders the
 * stages -- compute and exchange are one kernel. 0: grouped ncclSend/ncclRecv (peer mapping
 * unavailable, more than 8 ranks, or SPFFT_B200_P2P=0 in the environment). */
SPFFT_EXPORT SpfftError spfft_b200_transform_peer_exchange(SpfftTransform transform, int* enabled);
SPFFT_EXPORT SpfftError spfft_b200_float_transform_peer_exchange(SpfftFloatTransform transform,
                                                                 int* enabled);

/* Host-only view of the exchange a rank would perform (no GPU, no NCCL): block offsets / counts
 * (in complex elements) inside the stick-side buffer [dimZ][pitch(rank)] and the plane-side buffer
 * (one block [localPlanes][pitch(r)] per source rank), and the y-stage tables over all ranks'
 * sticks sorted by x*dimY+y (slot = y*Vy + x mod Vy, srcBase + plane*srcPitch = position in the
 * plane-side buffer). sticksAllRanks: the ranks' stick lists back to back. Per-rank outputs have
 * commSize entries, per-stick outputs sum(numSticksPerRank) entries, xtStart numXTiles+1
 * (at most dimX+1). Used by the CPU multi-process tests. */
SPFFT_EXPORT SpfftError spfft_b200_exchange_plan(
    int transformType, int isFloat, int dimX, int dimY, int dimZ, int commSize, int commRank,
    const int* numSticksPerRank, const int* sticksAllRanks, const int* planesPerRank,
    int* pitchPerRank, long long* stickOffset, long long* stickCount, long long* planeOffset,
    long long* planeCount, int* numXTiles, int* log2Vy, int* xtStart, int* stickSlot, int* srcBase,
    int* srcPitch);

/* The peer-memory form of the same exchange (host only): where the fused kernels store.
 * backward: row z of this rank's plane-major stick buffer goes to element rowOff[z] of the
 * plane-side buffer of rank rowRank[z] (dimZ entries each); forward: stick e (global sorted list,
 * same order as stickSlot above) of local plane zl goes to element fwdBase[e] + zl*srcPitch[e] of
 * the stick buffer of rank stickRank[e]; fwdTileRotate = x tile the forward kernels start at. */
SPFFT_EXPORT SpfftError spfft_b200_exchange_plan_peer(
    int transformType, int isFloat, int dimX, int dimY, int dimZ, int commSize, int commRank,
    const int* numSticksPerRank, const int* sticksAllRanks, const int* planesPerRank, int* rowRank,
    long long* rowOff, int* stickRank, int* fwdBase, int* fwdTileRotate);

/* ---- plan inspection (host only, no GPU needed) -------------------------------------------- */

/* The index conversion every transform performs at creation, exposed so that tests can compare it
 * bit for bit with the reference's convert_index_triplets (src/compression/indices.hpp:120-186).
 * triplets: 3*numValues ints. valueIndices: numValues ints out. stickIndices: up to
 * min(numValues, dimX*dimY) ints out. Either output may be NULL. Errors as the reference:
 * SPFFT_INVALID_PARAMETER_ERROR (numValues > dimX*dimY*dimZ), SPFFT_INVALID_INDICES_ERROR. */
SPFFT_EXPORT SpfftError spfft_b200_convert_index_triplets(int hermitianSymmetry, int dimX,
                                                          int dimY, int dimZ, int numValues,
                                                          const int* triplets, int* valueIndices,
                                                          int* stickIndices, int* numSticks);

/* Pointers (owned by the transform, host memory) to the two reference-defined index maps of an
 * existing transform: valueIndices[numValues] = stick*dimZ + z, stickIndices[numSticks] = x*dimY+y
 * ascending. Mirrors Parameters::local_value_indices / z_stick_xy_indices
 * (src/parameters/parameters.hpp). */
SPFFT_EXPORT SpfftError spfft_b200_transform_index_maps(SpfftTransform transform,
                                                        const int** valueIndices, int* numValues,
                                                        const int** stickIndices, int* numSticks);
SPFFT_EXPORT SpfftError spfft_b200_float_transform_index_maps(SpfftFloatTransform transform,
                                                              const int** valueIndices,
                                                              int* numValues,
                                                              const int** stickIndices,
                                                              int* numSticks);

/* ---- execution control / measurement ------------------------------------------------------- */

/* The CUDA stream (cudaStream_t) a transform enqueues its kernels on; the reference keeps it
 * private (src/execution/execution_gpu.cpp:53). Needed to time kernels with CUDA events. */
SPFFT_EXPORT SpfftError spfft_b200_transform_stream(SpfftTransform transform, void** stream);
SPFFT_EXPORT SpfftError spfft_b200_float_transform_stream(SpfftFloatTransform transform,
                                                          void** stream);

/* Per-kernel device timing. With profiling enabled every stage kernel of a transform call is
 * bracketed by CUDA events on the transform's stream. *_stage_times waits for the recorded events
 * and returns, over all backward / forward calls since the previous query, the number of distinct
 * stages, their names (static strings, e.g. "z backward") and the AVERAGE milliseconds per call;
 * it then forgets the recorded events. maxStages bounds the output arrays.
 * Takes the place of the reference's host-side rt_graph timers (src/timing/timing.hpp:34-62). */
SPFFT_EXPORT SpfftError spfft_b200_transform_set_profiling(SpfftTransform transform, int enable);
SPFFT_EXPORT SpfftError spfft_b200_transform_stage_times(SpfftTransform transform, int maxStages,
                                                         int* numStages, const char** names,
                                                         float* milliseconds);
SPFFT_EXPORT SpfftError spfft_b200_float_transform_set_profiling(SpfftFloatTransform transform,
                                                                 int enable);
SPFFT_EXPORT SpfftError spfft_b200_float_transform_stage_times(SpfftFloatTransform transform,
                                                               int maxStages, int* numStages,
                                                               const char** names,
                                                               float* milliseconds);

/* Number of kernels this library has launched since it was loaded (all transforms). */
SPFFT_EXPORT SpfftError spfft_b200_kernel_launch_count(long long int* count);

#ifdef __cplusplus
}
#endif
#endif
