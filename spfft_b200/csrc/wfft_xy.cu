// wfft_xy.cu -- fused xy stage on the warp FFT (wfft.hpp, wfft_kernels.cuh): y tiles and x tiles of
// every plane are items of ONE persistent kernel, the y <-> x hand-off plane lives in a small ring
// of scratch planes that stays resident in the 126 MB L2 (written once, read once, then dropped with
// discard.global.L2), so the stage moves only its algorithmic bytes through HBM:
//     backward:  sticks (sparse rows, gathered through the inverse map) -> y-FFT -> ring -> x-FFT -> space
//     forward :  space -> x-FFT -> ring -> y-FFT -> sticks (scattered through the inverse map)
// C2C, dimX == dimY == 512, one process-local slab of planes; double precision (also on distributed transforms) and
// single precision (local transforms; two transforms per warp on the packed fp32 pipe, WUnit in wfft_kernels.cuh).
// Replaces the two passes of the reference's 2-D cuFFT plans (src/fft/transform_2d_gpu.hpp:51-140)
// and its transposing unpack / pack kernels (src/transpose/gpu_kernels/local_transpose_kernels.cu).
//
// Schedule: items in hand-out order (w_decode_dense): for step u, the A tiles of plane u, then the B tiles of plane
// u - lag; an item = 8 columns (y tile) or 8 rows (x tile) of one plane. The CTAs CLAIM their items in that order
// from two interleaved global counters (WQueue), two items ahead; the grid is launched cooperatively, so every CTA
// is resident and waiting on an earlier item cannot deadlock. Inside a CTA the 8 / W groups of W warps (WGeom)
// each take W columns / rows of the CTA's item and run on their own: own sub-tile buffer, own named barrier, own
// flags. A group's B part waits until all A parts of its plane are complete, an A part until the B parts of the
// plane that used its ring slot before are complete (counters count group parts). Completion of a group's part k is
// published while its part k+1 runs (right after the group barrier that part needs anyway, by another warp than
// the one that handles the tensor copies), and the dependency of part k+1 is polled behind the exchange of part k
// with a relaxed load, so in the steady state no warp waits on a flag round trip.
#include <cstdlib>

#include "fast_launch.cuh"
#include "launch.h"
#include "wfft_kernels.cuh"

namespace sb {

#ifdef SB_WTRACE
// experiment builds only (-DSB_WTRACE): clock64 per phase of the backward kernel, summed by lane 0 of warp 0 of
// every CTA: g_wtrace[role][phase] (cycles), g_wtrace[role][15] = number of parts
__device__ unsigned long long g_wtrace[2][16];
#define W_TRACE_DECL long long tr_prev = clock64(); const bool tr_on = threadIdx.x == 0;
#define W_TRACE(role, ph)                                                       \
  if (tr_on) {                                                                  \
    const long long now_ = clock64();                                           \
    atomicAdd(&g_wtrace[role][ph], (unsigned long long)(now_ - tr_prev));       \
    tr_prev = now_;                                                             \
  }
#define W_TRACE_COUNT(role) if (tr_on) atomicAdd(&g_wtrace[role][15], 1ULL);
#define W_TRACE_SLOW(role) if (tr_on) atomicAdd(&g_wtrace[role][14], 1ULL);
__device__ __forceinline__ void w_trace_use(const cx<double>& v) { asm volatile("" ::"d"(v.x), "d"(v.y)); }
__device__ __forceinline__ void w_trace_use(const cx<f2>& v) { asm volatile("" ::"l"(v.x.r), "l"(v.y.r)); }
#define W_TRACE_USE(v) \
  _Pragma("unroll") for (int m_ = 0; m_ < 16; ++m_) w_trace_use((v)[m_]);
#else
#define W_TRACE_DECL
#define W_TRACE(role, ph)
#define W_TRACE_COUNT(role)
#define W_TRACE_SLOW(role)
#define W_TRACE_USE(v)
#endif

namespace {

// Flags. Completion counters are incremented with red.release.gpu after the data of a part has been written;
// they are POLLED with relaxed loads (an acquire load would be followed by an L1 invalidation that stalls the
// polling warp for the whole L2 round trip, ncu: CCTL.IVALL). That is sufficient here because every read of
// hand-off data that follows a successful poll is an L2 access (ld.global.cg / bulk tensor copy), issued after
// the poll in program order, and the GPU does not speculate loads.
__device__ __forceinline__ int w_ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
struct WxyCounters {
  int* aDone;
  int* bDone;
  int nA, nB, ring;
  // counter and threshold of a part's dependency (nullptr: none)
  __device__ __forceinline__ const int* dep_counter(const XYItem& it, int& need) const {
    if (it.roleA) {
      need = nB;
      return it.plane < ring ? nullptr : &bDone[it.plane - ring];
    }
    need = nA;
    return &aDone[it.plane];
  }
  __device__ __forceinline__ bool ready(const XYItem& it) const {
    int need;
    const int* c = dep_counter(it, need);
    return !c || w_ld_relaxed(c) >= need;
  }
};

// Work distribution. The CTAs CLAIM their items in hand-out order from two interleaved global counters (even /
// odd items; a CTA belongs to the queue blockIdx & 1): a CTA that falls behind (far L2 slices, waits) simply
// takes fewer items, so the tiles of a plane complete together instead of waiting for the slowest CTA of a static
// round-robin schedule (profiles/r02_summary.md: 21 % / 46 % of the parts found their dependency unsatisfied).
// All groups of a CTA work on the same item; the k-th item of the CTA is claimed by whichever group needs it
// first, two items ahead (the atomic's round trip hides behind a whole part), and handed to the others through
// a small tagged ring in shared memory.
struct WQueueShared {
  int tag[8];      // ordinal whose item index is in val[ordinal & 7]
  int val[8];
  int claimed;     // ordinals requested so far
  int prog[8];     // per group: ordinal it is working on
};
struct WQueue {
  WQueueShared* q;
  int* counter;  // this CTA's global counter
  int stride, offset, total, groups;
  __device__ __forceinline__ int item_of(int claim) const { return claim * stride + offset; }
  // group leader, at the start of its part k: does ordinal k + 2 still have to be claimed (by this group)?
  __device__ __forceinline__ bool own(int ord) const {
    if (atomicCAS(&q->claimed, ord, ord + 1) != ord) return false;
    // the slot still holds ordinal ord - 8: every group must be past it (a group at part k reads ordinals <= k + 1)
    for (int g = 0; g < groups; ++g)
      while (reinterpret_cast<volatile int*>(q->prog)[g] < ord - 8) {
      }
    return true;
  }
  // the claim itself: the returned value is first touched in post(), a whole part later, so the round trip of
  // the atomic never stalls the leader (inline asm: no phi / select on the result right behind the atomic)
  __device__ __forceinline__ void claim(bool mine, unsigned& raw) const {
    if (mine) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(raw) : "l"(counter) : "memory");
  }
  __device__ __forceinline__ void post(int ord, int claim) const {
    int item = item_of(claim);
    if (item > total) item = total;
    q->val[ord & 7] = item;
    __threadfence_block();
    reinterpret_cast<volatile int*>(q->tag)[ord & 7] = ord;
  }
  // every thread: item index of ordinal ord (>= total: no more work)
  __device__ __forceinline__ int consume(int ord) const {
    while (reinterpret_cast<volatile int*>(q->tag)[ord & 7] != ord) {
    }
    __threadfence_block();
    return reinterpret_cast<volatile int*>(q->val)[ord & 7];
  }
};
__device__ __forceinline__ void w_queue_init(WQueueShared* q, WQueue& wq, int* counters, int total, int groups) {
  wq.q = q;
  wq.stride = gridDim.x > 1 ? 2 : 1;
  wq.offset = wq.stride == 2 ? (blockIdx.x & 1) : 0;
  wq.counter = counters + wq.offset;
  wq.total = total;
  wq.groups = groups;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      q->tag[i] = -1;
      q->prog[i] = 0;
    }
    q->claimed = 2;
    for (int ord = 0; ord < 2; ++ord) {
      int item = wq.item_of(atomicAdd(wq.counter, 1));
      q->val[ord] = item > total ? total : item;
      q->tag[ord] = ord;
    }
  }
}

__device__ __forceinline__ void w_publish(int* counter) { w_red_release(counter); }

}  // namespace

// ------------------------------------------------------------------------------------------------
// Duties inside a group of W warps (lane 0 of ...):
//   warp 0      "leader"    issues and awaits the group's bulk tensor copies
//   warp 1      "poller"    looks at the dependency of the NEXT part while this part's loads are in flight and
//                           posts the answer after its tail (the L2 round trip hides behind the butterflies)
//   warp W - 1  "publisher" publishes the group's PREVIOUS part right after the group barrier of this part
// so that no warp carries more than one flag round trip per part and none of them in front of a barrier.
// ------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------
// backward: A = y tile (gather -> FFT -> transposed into the sub-tile -> TMA store into the ring),
//           B = x tile (one row per warp: ring -> FFT -> space domain)
// ------------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(kWThreads, 2)
    k_wxy_bwd(const __grid_constant__ XYArgs<T> a, const __grid_constant__ TensorMap ringMap,
              const __grid_constant__ WTw4<T> twp) {
  constexpr int N = kWN;
  using G = WGeom<W>;
  using Sc = typename WUnit<T>::Sc;           // scalar of a unit: double, or two floats (two transforms per warp)
  constexpr int kPer = WUnit<T>::kPer;        // transforms per warp
  constexpr int kTiles = kWTiles / kPer;      // items per plane and role
  constexpr int kRows = kWWarps * kPer;       // columns / rows per item
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  __shared__ int sReadySeq[G::kGroups];  // (k << 1) | ready: dependency of the group's part k seen satisfied
  __shared__ __align__(16) cx<Sc> sTw[4 * 32];
  // inverse-map entries of the thread's NEXT y part, fetched asynchronously one part ahead (no registers,
  // no exposed round trip in front of the gather)
  __shared__ __align__(16) uint4 sInv[2][2 * kPer][kWThreads];  // [parity of the part][column, half][thread]: conflict-free LDGSTS
  const int tid = threadIdx.x;
  const int w = tid >> 5, L = tid & 31;  // w: column / row of the item
  const int g = w / W, wl = w % W;       // group, warp inside the group
  const bool leader = L == 0 && wl == 0;
  const bool poller = L == 0 && wl == 1;
  const bool publisher = L == 0 && wl == W - 1;
  cx<Sc>* S = reinterpret_cast<cx<Sc>*>(smemRaw + (size_t)g * G::kSubBytes);
  const WAddr ad = w_addr<W>(wl, L);
  w_stage_twiddles(sTw, twp);
  __syncthreads();
  const int P = a.y.numPlanes;
  WxyCounters dep;
  dep.aDone = a.counters + 1;
  dep.bDone = a.counters + 1 + P;
  dep.nA = kTiles * G::kGroups;
  dep.nB = kTiles * G::kGroups;
  dep.ring = a.ring;
  const int total = 2 * kTiles * P;
  const size_t planeElems = (size_t)N * N;
  const int planLane = w_plan_lane<kPer>(w);  // the warp's (first) column inside its tile of the index plan
  __shared__ WQueueShared sQueue;
  WQueue wq;
  w_queue_init(&sQueue, wq, a.counters + 1 + 2 * P, total, G::kGroups);
  __syncthreads();

  int* pend = nullptr;   // counter of the group's previous part, not yet published (group-uniform)
  bool pendTma = false;  // ... whose output is a bulk tensor store issued by the group leader
  XYItem it, nx;
  int cur = wq.consume(0);
  if (cur < total) it = w_decode_dense<kTiles>(cur, P, a.lag);
  bool invAhead = false;
  int e0 = (cur < total && it.roleA) ? a.y.xtStart[w_plan_tile<kPer>(it.tile, w)] : 0;  // first stick of the part's x tile
  int tileBase = 0, tilePitch = 0;  // distributed: block of a single-source tile in the exchange buffer
  if constexpr (kPer == 1) {
    if (cur < total && it.roleA && a.y.srcBase) {
      tileBase = a.y.tileBase[it.tile];
      tilePitch = a.y.tilePitch[it.tile];
    }
  }
  W_TRACE_DECL
  for (int k = 0; cur < total; ++k) {
    if (leader) reinterpret_cast<volatile int*>(sQueue.prog)[g] = k;
    const bool mine = leader && wq.own(k + 2);
    unsigned claimRaw = 0;
    wq.claim(mine, claimRaw);  // (in flight until it is posted behind the tail)
    const int nxt = wq.consume(k + 1);
    if (nxt < total) nx = w_decode_dense<kTiles>(nxt, P, a.lag);
    const int trRole = it.roleA ? 0 : 1;
    (void)trRole;
    W_TRACE(trRole, 0)  // previous part's stores issued + decode
    // ---- dependency of this part: normally seen satisfied one part ago
    bool depSeen = false;
    if (k > 0) {
      // posted by the poller behind its tail of the previous part (sequence number: no barrier needed)
      int rs;
      do {
        rs = reinterpret_cast<volatile int*>(sReadySeq)[g];
      } while ((rs >> 1) != k);
      depSeen = rs & 1;
    }
    if (!depSeen) {
      W_TRACE_SLOW(trRole)
      w_group_sync<W>(g);
      if (leader) {
        if (pendTma) w_tma_wait_all();
        if (pend) w_publish(pend);
        while (!dep.ready(it)) __nanosleep(64);
      }
      pend = nullptr;
      pendTma = false;
      w_group_sync<W>(g);
    }
    W_TRACE(trRole, 1)  // dependency (slow path)
    // ---- loads
    cx<Sc> v[16];
    const int slot = it.plane % a.ring;
    if (it.roleA) {
      WInvUnit<kPer> iv;
      if (invAhead) {
        w_cp_async_wait();
        iv = w_inv_read<kPer>(sInv[k & 1], tid);
      } else {
        iv = w_inv_load<kPer>(a.y.inv, w_plan_tile<kPer>(it.tile, w), planLane, L);
      }
      bool perStick = false;
      if constexpr (kPer == 1) perStick = a.y.srcBase && tilePitch == 0;
      if (perStick) {
        // distributed, sticks of this tile from several ranks (tiles at a rank boundary): per-stick tables
        if constexpr (kPer == 1) {
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            v[m] = mk<T>(0, 0);
            if (iv.c[0].i[m] != kWNone) v[m] = *y_dist_stick<T, false>(a.y, e0 + iv.c[0].i[m], it.plane);
          }
        }
      } else {
        // local: row of the plane-major stick buffer; distributed: the tile's block inside its source rank's part
        // of the plane-side exchange buffer (read once: evict first, the ring stays in L2)
        const cx<T>* row = (kPer == 1 && a.y.srcBase) ? a.y.sticks + (size_t)tileBase + (size_t)it.plane * tilePitch
                                                      : a.y.sticks + (size_t)(it.plane + a.y.zRowOffset) * a.y.pitch + e0;
        w_gather_cs(v, row, iv);
      }
    } else {
      const cx<T>* src = a.scratch + (size_t)slot * planeElems + (size_t)(it.tile * kRows + w * kPer) * N + L;
#pragma unroll
      for (int m = 0; m < 16; ++m) v[m] = w_row_ldcg(src + 32 * m, N);
    }
    // ---- in flight behind this part's work: inverse map and first stick of the next y part, flag of the next part
    invAhead = nxt < total && nx.roleA;
    int e0Next = 0, tileBaseNext = 0, tilePitchNext = 0;
    if (invAhead) {
      w_inv_prefetch<kPer>(sInv[(k + 1) & 1], a.y.inv, w_plan_tile<kPer>(nx.tile, w), planLane, L, tid);
      e0Next = a.y.xtStart[w_plan_tile<kPer>(nx.tile, w)];
      if constexpr (kPer == 1) {
        if (a.y.srcBase) {
          tileBaseNext = a.y.tileBase[nx.tile];
          tilePitchNext = a.y.tilePitch[nx.tile];
        }
      }
    }
    W_TRACE(trRole, 2)  // loads issued (incl. inverse map wait)
    W_TRACE_USE(v)
    W_TRACE(trRole, 3)  // loads arrived
    w512_head<Sc, true>(v, L);
    W_TRACE_USE(v)
    W_TRACE(trRole, 5)  // head
    if (!it.roleA) {
      // the consumed hand-off row(s) of the warp (8 KB, fully read by now) are dropped from L2, not written back
      const char* rowBytes = reinterpret_cast<const char*>(a.scratch + (size_t)slot * planeElems +
                                                           (size_t)(it.tile * kRows + w * kPer) * N);
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(rowBytes + (size_t)L * 128) : "memory");
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(rowBytes + (size_t)(L + 32) * 128) : "memory");
    }
    // ---- the sub-tile is free (and the previous part complete) once its tensor store has finished; it was
    // issued a gather + head ago
    if (leader && pendTma) w_tma_wait_all();
    W_TRACE(trRole, 6)  // discard + wait for the tensor store + flag of the next part
    w_group_sync<W>(g);
    W_TRACE(trRole, 7)  // group barrier
    if (publisher && pend) w_publish(pend);  // all stores of the previous part precede the barrier
    w512_exchange<Sc, W>(v, S, ad);
    W_TRACE_USE(v)
    W_TRACE(trRole, 8)  // exchange (+ publish)
    // flag of the next part: looked at as late as possible, the round trip hides behind the tail
    int pollNeed = 0, pollSeen = 0;
    if (poller && nxt < total) {
      const int* c = dep.dep_counter(nx, pollNeed);
      if (c) pollSeen = w_ld_relaxed(c); else pollNeed = 0;
    }
    w512_tail<Sc, true>(v, sTw, L);
    W_TRACE_USE(v)
    W_TRACE(trRole, 9)  // tail
    if (poller) reinterpret_cast<volatile int*>(sReadySeq)[g] = ((k + 1) << 1) | (pollSeen >= pollNeed ? 1 : 0);
    if (mine) wq.post(k + 2, (int)claimRaw);
    if (it.roleA) {
      w512_col_store<Sc, W>(v, S, ad);
      fence_async_smem();
      W_TRACE(trRole, 10)  // column store
      w_group_sync<W>(g);
      W_TRACE(trRole, 11)  // group barrier before the tensor store
      if (leader) {
        const int c0 = (it.tile * kWWarps + g * W) * 2;
        const uint64_t keep = l2_policy_evict_last();
        tma_store_3d_hint(&ringMap, c0, 0, slot, S, keep);
        tma_store_3d_hint(&ringMap, c0, 256, slot, S + 256 * W, keep);
        tma_store_commit();
      }
      W_TRACE(trRole, 12)  // tensor store issue
      pend = &dep.aDone[it.plane];
      pendTma = true;
    } else {
      cx<T>* dst = static_cast<cx<T>*>(a.x.spaceOut) + (size_t)it.plane * planeElems +
                   (size_t)(it.tile * kRows + w * kPer) * N + L;
#pragma unroll
      for (int m = 0; m < 16; ++m) w_row_stcs(dst + 32 * m, N, v[m]);  // written once: evict first
      W_TRACE(trRole, 12)  // stores
      pend = &dep.bDone[it.plane];
      pendTma = false;
    }
    W_TRACE_COUNT(trRole)
    cur = nxt;
    it = nx;
    e0 = e0Next;
    tileBase = tileBaseNext;
    tilePitch = tilePitchNext;
  }
  // publish the group's last part (parts of other CTAs may wait for it); the bulk store of the last y
  // tile must have read the sub-tile before the CTA exits
  w_group_sync<W>(g);
  if (leader) {
    if (pendTma) w_tma_wait_all();
    if (pend) w_publish(pend);
  }
}

// ------------------------------------------------------------------------------------------------
// forward: A = x tile (one row per warp: space -> FFT -> ring), B = y tile (TMA load of the group's ring
// columns into its sub-tile -> FFT -> scatter into the stick rows)
// ------------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(kWThreads, 2)
    k_wxy_fwd(const __grid_constant__ XYArgs<T> a, const __grid_constant__ TensorMap ringMap,
              const __grid_constant__ WTw4<T> twp) {
  constexpr int N = kWN;
  using G = WGeom<W>;
  using Sc = typename WUnit<T>::Sc;
  constexpr int kPer = WUnit<T>::kPer;
  constexpr int kTiles = kWTiles / kPer;
  constexpr int kRows = kWWarps * kPer;
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  __shared__ int sReady[G::kGroups][2];
  __shared__ __align__(8) uint64_t full[G::kGroups];
  __shared__ __align__(16) cx<Sc> sTw[4 * 32];
  __shared__ __align__(16) uint4 sInv[2][2 * kPer][kWThreads];  // inverse-map entries of the thread's current y part, by parity
  const int tid = threadIdx.x;
  const int w = tid >> 5, L = tid & 31;
  const int g = w / W, wl = w % W;
  const bool leader = L == 0 && wl == 0;
  const bool poller = L == 0 && wl == 1;
  const bool publisher = L == 0 && wl == W - 1;
  cx<Sc>* S = reinterpret_cast<cx<Sc>*>(smemRaw + (size_t)g * G::kSubBytes);
  const WAddr ad = w_addr<W>(wl, L);
  w_stage_twiddles(sTw, twp);
  if (leader) {
    mbar_init(&full[g], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int P = a.y.numPlanes;
  WxyCounters dep;
  dep.aDone = a.counters + 1;
  dep.bDone = a.counters + 1 + P;
  dep.nA = kTiles * G::kGroups;
  dep.nB = kTiles * G::kGroups;
  dep.ring = a.ring;
  const int total = 2 * kTiles * P;
  const size_t planeElems = (size_t)N * N;
  const int planLane = w_plan_lane<kPer>(w);
  constexpr uint32_t kSubBytes = (uint32_t)G::kSubBytes;
  __shared__ WQueueShared sQueue;
  WQueue wq;
  w_queue_init(&sQueue, wq, a.counters + 1 + 2 * P, total, G::kGroups);
  __syncthreads();

  int* pend = nullptr;
  uint32_t phase = 0;
  bool preloaded = false;  // the sub-tile of the current (B) part is already on its way into S
  XYItem it, nx;
  // y tiles of a plane are visited starting at tile xtRotate (distributed: every rank then stores to a different
  // destination at any time); the ring columns, the inverse map and the sticks all follow the remapped tile
  auto decode = [&](int idx) {
    XYItem d = w_decode_dense<kTiles>(idx, P, a.lag);
    if constexpr (kPer == 1) {
      if (!d.roleA) d.tile = y_forward_tile_at<T>(a.y, d.tile);
    }
    return d;
  };
  int cur = wq.consume(0);
  if (cur < total) it = decode(cur);
  int e0 = (cur < total && !it.roleA) ? a.y.xtStart[w_plan_tile<kPer>(it.tile, w)] : 0;
  for (int k = 0; cur < total; ++k) {
    if (leader) reinterpret_cast<volatile int*>(sQueue.prog)[g] = k;
    const bool mine = leader && wq.own(k + 2);
    unsigned claimRaw = 0;
    wq.claim(mine, claimRaw);  // (in flight until it is posted behind the tail)
    const int nxt = wq.consume(k + 1);
    if (nxt < total) nx = decode(nxt);
    if (k == 0 || !sReady[g][k & 1]) {
      w_group_sync<W>(g);
      if (leader) {
        if (pend) w_publish(pend);
        while (!dep.ready(it)) __nanosleep(64);
      }
      pend = nullptr;
      w_group_sync<W>(g);
    }
    cx<Sc> v[16];
    const int slot = it.plane % a.ring;
    cx<Sc>* R = S + (size_t)wl * N;  // x parts: the warp's flat private region of the sub-tile (its row, or its two rows)
    // every warp of the group is past the barrier of the previous part, i.e. past its last access of S
    if (!preloaded && leader) {
      w_fence_proxy_async();
      mbar_expect_tx(&full[g], kSubBytes);
      if (it.roleA) {
        // the W rows of the group are contiguous in the space domain: one bulk copy
        w_bulk_load(S, static_cast<const cx<T>*>(a.x.spaceIn) + (size_t)it.plane * planeElems + (size_t)(it.tile * kWWarps + g * W) * kPer * N,
                    kSubBytes, &full[g], l2_policy_evict_first());
      } else {
        const int c0 = (it.tile * kWWarps + g * W) * 2;
        tma_load_3d(S, &ringMap, c0, 0, slot, &full[g]);
        tma_load_3d(S + 256 * W, &ringMap, c0, 256, slot, &full[g]);
      }
    }
    if (!it.roleA) w_inv_prefetch<kPer>(sInv[k & 1], a.y.inv, w_plan_tile<kPer>(it.tile, w), planLane, L, tid);
    // in flight behind this part's work: first stick of the next y part, flag of the next part
    const int e0Next = (nxt < total && !nx.roleA) ? a.y.xtStart[w_plan_tile<kPer>(nx.tile, w)] : 0;
    int pollNeed = 0, pollSeen = 0;
    if (poller && nxt < total) {
      const int* c = dep.dep_counter(nx, pollNeed);
      if (c) pollSeen = w_ld_relaxed(c); else pollNeed = 0;
    }
    mbar_wait(&full[g], phase);
    phase ^= 1;
    if (it.roleA) {
      w512_flat_load(v, R, L);
      __syncwarp();
    } else {
      w512_col_load<Sc, W>(v, S, ad);
      __syncwarp();
      // the consumed column segments are dropped from L2 where the group consumes whole 128-byte lines
      // (W == 8); narrower groups share their lines with the other groups of the CTA
      if constexpr (W == 8) {
        const char* tileBytes = reinterpret_cast<const char*>(a.scratch + (size_t)slot * planeElems + (size_t)it.tile * kRows);
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(tileBytes + (size_t)tid * N * sizeof(cx<T>)) : "memory");
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(tileBytes + (size_t)(tid + 256) * N * sizeof(cx<T>)) : "memory");
      }
    }
    w512_head<Sc, false>(v, L);
    if (poller) sReady[g][(k + 1) & 1] = pollSeen >= pollNeed;
    if (it.roleA)
      w512_exchange_flat<Sc>(v, R, L);
    else
      w512_exchange<Sc, W>(v, S, ad);
    // ---- every warp of the group is done with S; all stores of the previous part were issued before this point
    w_group_sync<W>(g);
    if (publisher && pend) w_publish(pend);
    // the input of the next part flows into S behind the tail and the stores of this one: the rows of an x part
    // come from the space domain (always there), the ring columns of a y part once their plane is complete
    preloaded = nxt < total && (nx.roleA || sReady[g][(k + 1) & 1]);
    if (leader && preloaded) {
      w_fence_proxy_async();
      mbar_expect_tx(&full[g], kSubBytes);
      if (nx.roleA) {
        w_bulk_load(S, static_cast<const cx<T>*>(a.x.spaceIn) + (size_t)nx.plane * planeElems + (size_t)(nx.tile * kWWarps + g * W) * kPer * N,
                    kSubBytes, &full[g], l2_policy_evict_first());
      } else {
        const int nslot = nx.plane % a.ring;
        const int c0 = (nx.tile * kWWarps + g * W) * 2;
        tma_load_3d(S, &ringMap, c0, 0, nslot, &full[g]);
        tma_load_3d(S + 256 * W, &ringMap, c0, 256, nslot, &full[g]);
      }
    }
    w512_tail<Sc, false>(v, sTw, L);
    if (mine) wq.post(k + 2, (int)claimRaw);
    if (it.roleA) {
      cx<T>* dst = a.scratch + (size_t)slot * planeElems + (size_t)(it.tile * kRows + w * kPer) * N + L;
      const uint64_t keep = l2_policy_evict_last();
#pragma unroll
      for (int m = 0; m < 16; ++m) w_row_st_hint(dst + 32 * m, N, v[m], keep);  // hand-off: stays in L2 until its y part ran
      pend = &dep.aDone[it.plane];
    } else {
      w_cp_async_wait();
      const WInvUnit<kPer> iv = w_inv_read<kPer>(sInv[k & 1], tid);
      if (kPer == 2 || !a.y.srcBase) {
        // (written once: evict first)
        w_scatter_cs(a.y.sticks + (size_t)(it.plane + a.y.zRowOffset) * a.y.pitch + e0, v, iv);
      } else if constexpr (kPer == 1) {
        if (a.y.tilePitch[it.tile] != 0) {
          // distributed, all sticks of the tile owned by one rank: straight into its stick buffer (NVLink peer
          // memory, or the local exchange buffer of the NCCL path) -- the y stage IS the exchange
          w_scatter_plain(y_dist_tile<T, true>(a.y, it.tile, it.plane), v, iv);
        } else {
#pragma unroll
          for (int m = 0; m < 16; ++m)
            if (iv.c[0].i[m] != kWNone) *y_dist_stick<T, true>(a.y, e0 + iv.c[0].i[m], it.plane) = v[m];
        }
      }
      pend = &dep.bDone[it.plane];
    }
    cur = nxt;
    it = nx;
    e0 = e0Next;
  }
  // publish the group's last part (parts of other CTAs may wait for it)
  w_group_sync<W>(g);
  if (leader && pend) w_publish(pend);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace {

// columns per group: backward 4, forward 8 (the forward y tiles can drop their consumed hand-off lines from L2
// only when one group consumes whole 128-byte lines); SPFFT_B200_WGROUP = 2 / 4 / 8 overrides both (experiments)
int wxy_group(int forward) {
  static const int w = [] {
    const char* e = getenv("SPFFT_B200_WGROUP");
    const int v = e ? atoi(e) : 0;
    return (v == 2 || v == 4 || v == 8) ? v : 0;
  }();
  return w ? w : (forward ? 8 : 4);
}

size_t wxy_smem() { return kWTileBytes; }

template <typename T, int W>
int wxy_grid_w(int* gridOut) {
  static int cached = 0;
  if (cached > 0) {
    *gridOut = cached;
    return 0;
  }
  cudaError_t e = cudaFuncSetAttribute(k_wxy_bwd<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wxy_smem());
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_wxy_fwd<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wxy_smem());
  if (e != cudaSuccess) return (int)e;
  int b0 = 0, b1 = 0, dev = 0, sms = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_wxy_bwd<T, W>, kWThreads, wxy_smem());
  if (e != cudaSuccess) return (int)e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_wxy_fwd<T, W>, kWThreads, wxy_smem());
  if (e != cudaSuccess) return (int)e;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  int coop = 0;
  e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (e != cudaSuccess) return (int)e;
  const int perSm = b0 < b1 ? b0 : b1;
  if (perSm <= 0 || !coop) return (int)cudaErrorInvalidConfiguration;
  cached = perSm * sms;
  *gridOut = cached;
  return 0;
}
template <typename T>
int wxy_grid(int* gridOut) {
  // (same launch bounds and shared memory for every W: one occupancy figure)
  int ga = 0, gb = 0;
  int err = wxy_grid_w<T, 8>(&ga);
  if (err) return err;
  err = wxy_grid_w<T, 4>(&gb);
  if (err) return err;
  int gc = 0;
  err = wxy_grid_w<T, 2>(&gc);
  if (err) return err;
  *gridOut = ga < gb ? (ga < gc ? ga : gc) : (gb < gc ? gb : gc);
  return 0;
}

template <typename T, typename Kernel>
int wxy_launch(Kernel kernel, int grid, const XYArgs<T>& a, const TensorMap& map, const WTw4<T>& tw, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kWThreads);
  cfg.dynamicSmemBytes = wxy_smem();
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kernel, a, map, tw);
}

template <typename T, int W>
int wxy_launch_w(int forward, int grid, const XYArgs<T>& a, const WTw4<T>& tw, cudaStream_t s) {
  // the hand-off ring as planes of 16-byte units (one complex double / two complex floats): the same box of
  // W units x 256 rows in both precisions
  constexpr int kPer = WUnit<T>::kPer;
  TensorMap map;
  const int err = make_tile_map(&map, a.scratch, 16, kWN / kPer, kWN, kWN / kPer, a.ring, (long long)kWN * kWN / kPer, W, 256);
  if (err) return err;
  return forward ? wxy_launch<T>(k_wxy_fwd<T, W>, grid, a, map, tw, s) : wxy_launch<T>(k_wxy_bwd<T, W>, grid, a, map, tw, s);
}

template <typename T>
int wxy_run(int forward, const XYArgs<T>& a, cudaStream_t s) {
  constexpr int kPer = WUnit<T>::kPer;
  if (a.y.numPlanes <= 0) return 0;
  // (the index plan's tiles have 8 columns / rows in both precisions; a single-precision item takes two of them)
  if (a.x.nx != kWN || a.y.ny != kWN || a.y.nxf != kWN || !a.y.inv || a.y.wireF32 || a.y.numXTiles != kWN / kWWarps ||
      a.x.numRowTiles != kWN / kWWarps)
    return (int)cudaErrorInvalidValue;
  if (kPer == 2 && (a.y.srcBase || (a.y.pitch & 1) || (reinterpret_cast<size_t>(a.y.sticks) & 15)))
    return (int)cudaErrorInvalidValue;  // single precision: local transforms only
  int grid = 0;
  int err = wxy_grid<T>(&grid);
  if (err) return err;
  const long long total = 2LL * (kWN / kWWarps / kPer) * a.y.numPlanes;
  if (total > 0x3fffffffLL) return (int)cudaErrorInvalidConfiguration;
  if (grid > total) grid = (int)total;
  static const WTw4<T> tw = [] {
    WTw4<T> t;
    wfft_lane_twiddles<T>(kWN, 32, &t.w[0][0]);
    return t;
  }();
  cudaError_t e = cudaMemsetAsync(a.counters, 0, sizeof(int) * (3 + 2 * (size_t)a.y.numPlanes), s);
  if (e != cudaSuccess) return (int)e;
  sb_note_launches(1);
  switch (wxy_group(forward)) {
    case 8: return wxy_launch_w<T, 8>(forward, grid, a, tw, s);
    case 4: return wxy_launch_w<T, 4>(forward, grid, a, tw, s);
    default: return wxy_launch_w<T, 2>(forward, grid, a, tw, s);
  }
}

}  // namespace
}  // namespace sb

extern "C" {

#ifdef SB_WTRACE
__attribute__((visibility("default"))) int sb_wxy_trace_read(unsigned long long* host /* [2][16] */, int reset) {
  if (cudaMemcpyFromSymbol(host, sb::g_wtrace, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[32] = {0};
    if (cudaMemcpyToSymbol(sb::g_wtrace, z, sizeof z) != cudaSuccess) return -2;
  }
  return 0;
}
#endif

int sb_wxy_config(int isFloat, int n, int numPlanes, int* ring, int* lag, int* numCounters) {
  if (n != sb::kWN) return (int)cudaErrorInvalidValue;
  int grid = 0;
  const int err = isFloat ? sb::wxy_grid<float>(&grid) : sb::wxy_grid<double>(&grid);
  if (err) return err;
  const int perStep = 2 * (n / sb::kWWarps) / (isFloat ? 2 : 1);  // items per plane (single precision: 16 columns / rows each)
  // A part is published one part later than it ran and polled one part before it is needed, and the CTAs drift
  // apart by about one part: the B parts of a plane follow its A parts by the parts in flight (grid / perStep
  // steps) + 5 steps. Measured at 512^3 double (profiles/r02_summary.md): lag 6 / 8 / 10 / 12 = 1.29+1.36 /
  // 1.33+1.31 / 1.38+1.37 / 1.39+1.44 ms (backward + forward), ring = 2 lag + 2; a ring of lag + 4 makes the A
  // parts wait for their slot (1.46+1.44 ms at lag 8).
  int l = (grid + perStep - 1) / perStep + 5;
  if (const char* e = getenv("SPFFT_B200_XY_LAG")) l = atoi(e) > 0 ? atoi(e) : l;
  int r = 2 * l + 2;
  if (const char* e = getenv("SPFFT_B200_XY_RING")) r = atoi(e) > l ? atoi(e) : l + 1;
  if (numPlanes <= r) r = numPlanes > 0 ? numPlanes : 1;  // every plane has its own slot: no reuse waits
  *ring = r;
  *lag = l;
  *numCounters = 3 + 2 * (numPlanes > 0 ? numPlanes : 0);  // [0] unused, 2 P completion counters, 2 queue counters
  return 0;
}

int sb_launch_wxy_f64(int forward, const sb::XYArgs<double>* args, void* stream) {
  return sb::wxy_run<double>(forward, *args, static_cast<cudaStream_t>(stream));
}
int sb_launch_wxy_f32(int forward, const sb::XYArgs<float>* args, void* stream) {
  return sb::wxy_run<float>(forward, *args, static_cast<cudaStream_t>(stream));
}
}
