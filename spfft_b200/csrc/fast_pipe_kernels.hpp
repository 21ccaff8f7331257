// fast_pipe_kernels.hpp -- tile bodies of the pipelined persistent xy kernel (fast_pipe.cu).
//
// Same arithmetic as the tiles of fast_stage_kernels.hpp (register radix-8 Stockham, two
// exchanges), but the INPUT of a tile does not come from global memory: it was staged into the
// tile's shared-memory buffer `B` by an asynchronous bulk copy (TMA, cp.async.bulk) that ran
// while the previous tiles were being transformed. A tile therefore
//     1. reads its 8 values per thread from B            (staged layout = global layout)
//     2. barrier                                         (B now becomes the exchange buffer)
//     3. stages 0 .. last-1 with their exchanges in B    (fast_fft_head)
//     4. reads the input of the last stage, barrier      -> B is free: `onFree()` lets one thread
//                                                           start the bulk copy of a later tile
//     5. last stage in registers, stores to global memory
// Layouts of the y<->x hand-off ("scratch" slot of one plane, N x N complex):
//     backward: row major  [y][x]          y tile stores 128-byte row segments, x tile loads
//                                          V whole rows = one contiguous block
//     forward : tile major [x / V][y][V]   x tile stores 128-byte segments, y tile loads one
//                                          contiguous block
// Replaces, like the kernels it supersedes, the two passes of the reference's cuFFT 2-D plans
// (src/fft/transform_2d_gpu.hpp:51-140) and the local transpose
// (src/transpose/gpu_kernels/local_transpose_kernels.cu:48-201).
#pragma once
#include "fast_stage_kernels.hpp"

namespace sb {

struct NoFree {
  SB_DEV void operator()() const {}
};

// Source block of an item's input (one contiguous range of global memory) and its size in bytes;
// 0 bytes: nothing to stage (a y tile without sticks).
template <typename T, int N, bool BWD>
SB_HD const cx<T>* pipe_item_source(const XYArgs<T>& a, const XYItem& it, unsigned* bytes) {
  constexpr int V = 1 << FastLanes<T>::log2V;
  constexpr unsigned tileBytes = (unsigned)(sizeof(cx<T>) * (size_t)N * V);
  const size_t planeElems = (size_t)N * N;
  const cx<T>* slot = a.scratch + (size_t)(it.plane % a.ring) * planeElems;
  const bool yTile = BWD ? it.roleA : !it.roleA;
  if (yTile) {
    const int e0 = a.y.xtStart[it.tile], e1 = a.y.xtStart[it.tile + 1];
    if (BWD) {  // sticks of the x tile, contiguous in the plane's row of the stick buffer
      *bytes = (unsigned)(e1 - e0) * (unsigned)sizeof(cx<T>);
      return a.y.sticks + (size_t)(it.plane + a.y.zRowOffset) * a.y.pitch + e0;
    }
    // forward: tile-major scratch block (only needed if some stick lives in these columns)
    *bytes = e1 > e0 ? tileBytes : 0u;
    return slot + (size_t)it.tile * N * V;
  }
  *bytes = tileBytes;
  if (BWD) return slot + (size_t)it.tile * V * N;  // V rows of the row-major scratch plane
  return static_cast<const cx<T>*>(a.x.spaceIn) + (size_t)it.plane * planeElems + (size_t)it.tile * V * N;
}

// One item on a staged buffer (dispatch on direction and role).
template <typename T, int N, bool BWD, typename OnFree = NoFree>
SB_DEV void pipe_run_item(const XYArgs<T>& a, const XYItem& it, cx<T>* B, const cx<T>* tw, Ctx ctx,
                          OnFree onFree = OnFree());

#define SB_PIPE_ROW_IDS               \
  cx<T>* v = SB_RP(vAll, 8);          \
  const int lane = tid & (V - 1);     \
  const int j = tid >> LOG2V;         \
  (void)nthr;
#define SB_PIPE_COL_IDS               \
  cx<T>* v = SB_RP(vAll, 8);          \
  const int lane = tid / TT;          \
  const int j = tid & (TT - 1);       \
  (void)nthr;

// ---- backward, y tile: staged sticks of x tile `xt` (contiguous, B[i] = i-th stick of the tile at
// this plane) -> y-FFT -> row-major scratch plane ------------------------------------------------
template <typename T, int N, typename OnFree = NoFree>
SB_DEV void pipe_y_backward(const YArgs<T>& a, int xt, cx<T>* B, cx<T>* plane, const cx<T>* tw, Ctx ctx,
                            OnFree onFree = OnFree()) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  constexpr int THREADS = V * TT;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  cx<T>* planeTile = plane + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  if (e0 == e1) {
    // empty x tile: the x stage still reads these columns -> zeros, nothing was staged
    // (barrier first: every thread has observed the buffer's mbarrier phase before it is re-armed)
    SB_PHASE_BEGIN
    (void)tid;
    (void)nthr;
    SB_PHASE_END
    SB_PHASE_BEGIN
    if (tid == 0) onFree();
    for (int i = tid; i < N * V; i += nthr) {
      const int y = i >> LOG2V;
      const int lane = i & (V - 1);
      if (lane < lanesValid) st_g<Mem::L2Only>(planeTile + (size_t)y * a.nxf + lane, mk<T>(0, 0));
    }
    SB_PHASE_END_NOSYNC
    return;
  }
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  SB_PIPE_COL_IDS
  (void)lane;
  (void)j;
  const Inv8 iv = load_inv8(a.inv + ((size_t)xt * THREADS + tid) * 8);
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = iv.i[m] != kNoEntry ? B[iv.i[m]] : mk<T>(0, 0);
  SB_PHASE_END
  SB_MARK(ctx, 2);
  fast_fft_head<T, N, LOG2V, true, SwzCol, true, false, true>(vAll, B, tw, ctx);
  SB_PHASE_BEGIN
  SB_PIPE_ROW_IDS
  fast_fft_tail_read<T, N, LOG2V, SwzCol>(v, B, j, lane);
  SB_PHASE_END
  SB_PHASE_BEGIN
  SB_PIPE_ROW_IDS
  if (tid == 0) onFree();
  fast_stage<T, N, true, FastPlan<N>::numStages - 1, true>(v, j, tw);
  if (lane < lanesValid) {
#pragma unroll
    for (int m = 0; m < 8; ++m) st_g<Mem::L2Only>(planeTile + (size_t)(j + TT * m) * a.nxf + lane, v[m]);
  }
  SB_PHASE_END_NOSYNC
}

// ---- backward, x tile: staged rows y0 .. y0+V-1 of the scratch plane (B[lane*N + x]) -> x-FFT ->
// rows of the space domain -----------------------------------------------------------------------
// ---- forward, x tile: staged rows of the space domain -> x-FFT -> tile-major scratch plane ------
template <typename T, int N, bool BWD, typename OnFree = NoFree>
SB_DEV void pipe_x_tile(cx<T>* B, cx<T>* out, int y0, const cx<T>* tw, Ctx ctx,
                        OnFree onFree = OnFree()) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  SB_PIPE_COL_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = B[(size_t)lane * N + j + TT * m];
  SB_PHASE_END
  SB_MARK(ctx, 2);
  fast_fft_head<T, N, LOG2V, BWD, SwzCol, true, true, true>(vAll, B, tw, ctx);
  SB_PHASE_BEGIN
  SB_PIPE_COL_IDS
  fast_fft_tail_read<T, N, LOG2V, SwzCol>(v, B, j, lane);
  SB_PHASE_END
  SB_PHASE_BEGIN
  SB_PIPE_COL_IDS
  if (tid == 0) onFree();
  fast_stage<T, N, BWD, FastPlan<N>::numStages - 1, true>(v, j, tw);
  if (BWD) {
    cx<T>* dst = out + (size_t)(y0 + lane) * N + j;  // space domain rows
#pragma unroll
    for (int m = 0; m < 8; ++m) dst[TT * m] = v[m];
  } else {
    // tile-major scratch: element (x, y) at (x / V) * N*V + y*V + x % V
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int x = j + TT * m;
      st_g<Mem::L2Only>(out + (size_t)(x >> LOG2V) * (N * V) + (size_t)(y0 + lane) * V + (x & (V - 1)), v[m]);
    }
  }
  SB_PHASE_END_NOSYNC
}

// ---- forward, y tile: staged tile-major block of x tile `xt` (B[y*V + lane]) -> y-FFT -> sticks ---
template <typename T, int N, typename OnFree = NoFree>
SB_DEV void pipe_y_forward(const YArgs<T>& a, int xt, cx<T>* B, cx<T>* stickRow, const cx<T>* tw, Ctx ctx,
                           OnFree onFree = OnFree()) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  constexpr int THREADS = V * TT;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  if (e0 == e1) {  // no stick needs these columns (nothing was staged)
    SB_PHASE_BEGIN
    (void)tid;
    (void)nthr;
    SB_PHASE_END
    SB_PHASE_BEGIN
    (void)nthr;
    if (tid == 0) onFree();
    SB_PHASE_END_NOSYNC
    return;
  }
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  SB_PIPE_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = B[((size_t)(j + TT * m) << LOG2V) + lane];
  SB_PHASE_END
  SB_MARK(ctx, 2);
  fast_fft_head<T, N, LOG2V, false, SwzCol, false, false, true>(vAll, B, tw, ctx);
  SB_PHASE_BEGIN
  SB_PIPE_COL_IDS
  fast_fft_tail_read<T, N, LOG2V, SwzCol>(v, B, j, lane);
  SB_PHASE_END
  SB_PHASE_BEGIN
  SB_PIPE_COL_IDS
  (void)lane;
  if (tid == 0) onFree();
  fast_stage<T, N, false, FastPlan<N>::numStages - 1, true>(v, j, tw);
  const Inv8 iv = load_inv8(a.inv + ((size_t)xt * THREADS + tid) * 8);
  cx<T>* out = stickRow + e0;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    if (iv.i[m] != kNoEntry) out[iv.i[m]] = v[m];
  }
  SB_PHASE_END_NOSYNC
}

template <typename T, int N, bool BWD, typename OnFree>
SB_DEV void pipe_run_item(const XYArgs<T>& a, const XYItem& it, cx<T>* B, const cx<T>* tw, Ctx ctx,
                          OnFree onFree) {
  constexpr int V = 1 << FastLanes<T>::log2V;
  const size_t planeElems = (size_t)N * N;
  cx<T>* slot = a.scratch + (size_t)(it.plane % a.ring) * planeElems;
  cx<T>* stickRow = a.y.sticks + (size_t)(it.plane + a.y.zRowOffset) * a.y.pitch;
  if (BWD) {
    if (it.roleA)
      pipe_y_backward<T, N>(a.y, it.tile, B, slot, tw, ctx, onFree);
    else
      pipe_x_tile<T, N, true>(B, static_cast<cx<T>*>(a.x.spaceOut) + (size_t)it.plane * planeElems,
                              it.tile * V, tw, ctx, onFree);
  } else {
    if (it.roleA)
      pipe_x_tile<T, N, false>(B, slot, it.tile * V, tw, ctx, onFree);
    else
      pipe_y_forward<T, N>(a.y, it.tile, B, stickRow, tw, ctx, onFree);
  }
}

#undef SB_PIPE_ROW_IDS
#undef SB_PIPE_COL_IDS

}  // namespace sb
