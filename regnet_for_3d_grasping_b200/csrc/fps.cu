// Farthest point sampling for sm_100a: one thread-block CLUSTER per cloud, points and running min-distances
// resident in registers for the whole kernel, argmax by redux.sync, cluster-wide exchange through distributed
// shared memory with st.async + mbarrier (no barrier.cluster on the critical path).
//
// Replaces csrc/sampling_kernel.cu:47-170 of the reference (grid = B blocks of <= 512 threads, min-distance
// array in global memory, 10-level shared-memory tree per iteration).  Results are bit-identical, including the
// tie rule that the reference's strided scan + tree reduction implies:
//   among the points attaining the maximum min-distance, the winner is the one whose slot t = j mod block has
//   the smallest bit-reversed value (log2(block) bits), and within a slot the smallest j;
//   block = min(nextpow2(N), 512), at least 16;  if the maximum is 0 the previous pick is repeated.
// That rule is folded into one 32-bit key  tie(j) = brev(j mod block) | (j div block)  so that the argmax is
// (max distance bits, min tie): one redux.sync.max per level, plus a redux.sync.min only when two lanes tie.
//
// Per iteration (i = 1 .. M-1), every thread:
//   1. updates min-distance of its PPT register-resident points against the current centroid, keeps its best;
//   2. warp argmax -> the winning lane writes {dist bits} and {tie,x,y,z} for its warp to shared memory;
//   3. __syncthreads; every warp reduces the W warp records redundantly (no second barrier, records are
//      double-buffered by iteration parity);
//   4. (cluster only) lanes 0..CS-1 of warp 0 push the CTA record into every peer's shared memory with
//      st.async.mbarrier::complete_tx; all threads wait on the local mbarrier of this parity, reduce the CS
//      records.  The winner's coordinates travel with the record, so no global load sits on the critical path.
#include <stdlib.h>

#include "common.cuh"

namespace regnet {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t NO_TIE = 0xffffffffu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_u32(uint32_t addr, uint32_t a, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(addr), "r"(a), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();  // protocol bug: fail the launch instead of hanging the GPU
  }
}

// Lane holding the warp-wide winner of (max d, min tie); `wmax` = the maximum on return (all lanes).
__device__ __forceinline__ int pick_lane(uint32_t d, uint32_t tie, uint32_t& wmax) {
  wmax = __reduce_max_sync(FULL, d);
  unsigned vote = __ballot_sync(FULL, d == wmax);
  if (vote & (vote - 1)) {  // several lanes attain the maximum: apply the reference's tie order
    const uint32_t mine = (d == wmax) ? tie : NO_TIE;
    const uint32_t wtie = __reduce_min_sync(FULL, mine);
    vote = __ballot_sync(FULL, d == wmax && mine == wtie);
  }
  return __ffs(vote) - 1;
}

// Coordinates of register-resident point `k` of this thread (k is not a compile-time constant, the arrays live in
// registers): a switch, which ptxas turns into an indexed branch -- executed by a single lane per warp.
template <int PPT>
__device__ __forceinline__ void pick_coords(int k, const float (&px)[PPT], const float (&py)[PPT], const float (&pz)[PPT],
                                            float& x, float& y, float& z) {
#define RN_PC(K) case K: if (K < PPT) { x = px[K < PPT ? K : 0]; y = py[K < PPT ? K : 0]; z = pz[K < PPT ? K : 0]; } break;
  switch (k) {
    RN_PC(0) RN_PC(1) RN_PC(2) RN_PC(3) RN_PC(4) RN_PC(5) RN_PC(6) RN_PC(7) RN_PC(8) RN_PC(9) RN_PC(10) RN_PC(11)
    RN_PC(12) RN_PC(13) RN_PC(14) RN_PC(15) RN_PC(16) RN_PC(17) RN_PC(18) RN_PC(19) RN_PC(20) RN_PC(21) RN_PC(22)
    RN_PC(23) RN_PC(24) RN_PC(25) RN_PC(26) RN_PC(27) RN_PC(28) RN_PC(29) RN_PC(30) RN_PC(31)
    default: break;
  }
#undef RN_PC
}

template <int CS, int T, int PPT, bool MBAR>
__global__ void __launch_bounds__(T, 1)
fps_kernel(const float* __restrict__ pts, Strides3 st, int N, int M, int nbits, int64_t* __restrict__ idx64,
           int32_t* __restrict__ idx32, float* __restrict__ new_xyz, const int32_t* __restrict__ n_var) {
  constexpr int W = T / 32;
  if (n_var) {  // per-cloud point count (region stage: FPS over each cloud's positive points)
    N = n_var[blockIdx.x / CS];
    if (N <= M) return;  // not enough points: the caller does not use FPS for this cloud (whole cluster leaves)
    nbits = N <= 1 ? 0 : 32 - __clz(N - 1);
    nbits = min(9, max(4, nbits));
  }
  // WARP_PUSH (cluster of at most 32 warps): every warp pushes its own record straight into every CTA of the cluster,
  // slot = rank * W + warp, and one 32-lane argmax replaces the CTA-level stage (no __syncthreads, one pick less).
  // The two exchange schemes never coexist in one instance, so their record arrays are sized to zero-ish when unused:
  // the co-running instance must fit the ~2.5 KB a tensor kernel's CTA leaves of the SM's shared memory.
  constexpr bool WARP_PUSH = MBAR && CS > 1 && CS * W <= 32;
  constexpr int NW = WARP_PUSH ? 1 : W, NC = WARP_PUSH ? 1 : 8, NA = WARP_PUSH ? 32 : 1;
  __shared__ uint32_t warp_d[2][NW];
  __shared__ __align__(16) uint4 warp_r[2][NW];  // {tie, x, y, z}
  __shared__ uint32_t cta_d[2][NC];               // slot r is written by cluster rank r (DSMEM)
  __shared__ __align__(16) uint4 cta_r[2][NC];
  __shared__ uint32_t all_d[2][NA];
  __shared__ __align__(16) uint4 all_r[2][NA];
  __shared__ __align__(8) uint64_t bars[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = (CS > 1) ? cluster_ctarank() : 0u;
  const int cloud = blockIdx.x / CS;
  const float* __restrict__ p = pts + (int64_t)cloud * st.b;
  const uint32_t mask = (1u << nbits) - 1u;

  // this thread's points: j = rank*T + tid + k*CS*T.  CS*T is a multiple of the reference block size, so all
  // of them fall into the same reference slot and ascending k == ascending j (the in-slot tie order).
  float px[PPT], py[PPT], pz[PPT], md[PPT];
  const int jbase = (int)rank * T + tid;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int j = jbase + k * CS * T;
    if (j < N) {
      px[k] = p[(int64_t)j * st.n];
      py[k] = p[(int64_t)j * st.n + st.c];
      pz[k] = p[(int64_t)j * st.n + 2 * st.c];
      md[k] = __int_as_float(0x7f800000);  // +inf plays the reference's temp = -1 ("not yet measured")
    } else {
      px[k] = py[k] = pz[k] = 0.f;
      md[k] = 0.f;                         // never strictly greater than anything: cannot win
    }
  }
  float cx = p[0], cy = p[st.c], cz = p[2 * st.c];
  int cur = 0;
  if (rank == 0 && tid == 0) {
    if (idx64) idx64[(int64_t)cloud * M] = 0;
    if (idx32) idx32[(int64_t)cloud * M] = 0;
    if (new_xyz) {
      float* o = new_xyz + (int64_t)cloud * 3 * M;
      o[0] = cx; o[M] = cy; o[2 * (int64_t)M] = cz;
    }
  }
  constexpr uint32_t TX_BYTES = (WARP_PUSH ? CS * W : CS) * 20;  // per iteration each sender's record is 16 + 4 bytes
  if (CS > 1) {
    if (MBAR && tid == 0) {
      mbar_init(smem_u32(&bars[0]), 1);
      mbar_init(smem_u32(&bars[1]), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_arm(smem_u32(&bars[1]), TX_BYTES);           // iteration 1 (odd parity slot)
      if (M > 2) mbar_arm(smem_u32(&bars[0]), TX_BYTES);  // iteration 2
    }
    cluster_sync_all();  // peers' shared memory and barriers exist before the first remote store
  }

  for (int i = 1; i < M; ++i) {
    const int par = i & 1;
    // The scan tracks only (best, slot): 10 instructions per point instead of 14 (four selects for j, x, y, z).  The
    // winner's coordinates are looked up afterwards by ONE lane per warp through a jump table (pick_coords).
    float best = 0.f;
    int bk = -1;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const float d = sqdist_ref(cx, cy, cz, px[k], py[k], pz[k]);
      const float m = fminf(md[k], d);
      md[k] = m;
      if (m > best) { best = m; bk = k; }
    }
    const int bj = bk < 0 ? -1 : jbase + bk * CS * T;
    const uint32_t tie = (bj < 0) ? NO_TIE : (__brev((uint32_t)bj & mask) | ((uint32_t)bj >> nbits));
    uint32_t dmax;
    // -- warp level
    const int src1 = pick_lane(__float_as_uint(best), tie, dmax);
    float bx = 0.f, by = 0.f, bz = 0.f;
    if (lane == src1) pick_coords<PPT>(bk, px, py, pz, bx, by, bz);
    uint4 win;
    if (WARP_PUSH) {
      // the winning lane's record goes round the warp, lanes 0..CS-1 push it to the CS CTAs of the cluster
      uint4 rec;
      rec.x = __shfl_sync(FULL, tie, src1);
      rec.y = __shfl_sync(FULL, __float_as_uint(bx), src1);
      rec.z = __shfl_sync(FULL, __float_as_uint(by), src1);
      rec.w = __shfl_sync(FULL, __float_as_uint(bz), src1);
      if (lane < CS) {
        const uint32_t slot = rank * W + warp;
        const uint32_t dst_bar = mapa_u32(smem_u32(&bars[par]), (uint32_t)lane);
        st_async_v4(mapa_u32(smem_u32(&all_r[par][slot]), (uint32_t)lane), rec, dst_bar);
        st_async_u32(mapa_u32(smem_u32(&all_d[par][slot]), (uint32_t)lane), dmax, dst_bar);
      }
      mbar_wait_cluster(smem_u32(&bars[par]), (uint32_t)((i - 1) >> 1) & 1u);  // k-th use of this parity's barrier
      // Re-arm for iteration i+2.  Every thread has passed the wait of iteration i before any warp of this CTA can send
      // its record of iteration i+1 (sends follow that wait in program order), and a peer's record of iteration i+2
      // needs this CTA's records of iteration i+1 first -- but the arming thread must not be overtaken by its OWN
      // warp-mates' i+1 records, hence warp 0 arms before it goes on (same warp, program order).
      if (tid == 0 && i + 2 < M) mbar_arm(smem_u32(&bars[par]), TX_BYTES);
      const int src3 = pick_lane(lane < CS * W ? all_d[par][lane] : 0u, lane < CS * W ? all_r[par][lane].x : NO_TIE, dmax);
      win = all_r[par][src3];
    } else {
    if (lane == src1) {
      warp_d[par][warp] = dmax;
      warp_r[par][warp] = make_uint4(tie, __float_as_uint(bx), __float_as_uint(by), __float_as_uint(bz));
    }
    __syncthreads();
    // -- CTA level (every warp, redundantly)
    const int src2 = pick_lane(lane < W ? warp_d[par][lane] : 0u, lane < W ? warp_r[par][lane].x : NO_TIE, dmax);
    win = warp_r[par][src2];
    // -- cluster level
    if (CS > 1) {
      if (warp == 0 && lane < CS) {
        const uint32_t dst_r = mapa_u32(smem_u32(&cta_r[par][rank]), (uint32_t)lane);
        const uint32_t dst_d = mapa_u32(smem_u32(&cta_d[par][rank]), (uint32_t)lane);
        if (MBAR) {
          const uint32_t dst_bar = mapa_u32(smem_u32(&bars[par]), (uint32_t)lane);
          st_async_v4(dst_r, win, dst_bar);
          st_async_u32(dst_d, dmax, dst_bar);
        } else {
          st_cluster_v4(dst_r, win);
          st_cluster_u32(dst_d, dmax);
        }
      }
      if (MBAR) {
        mbar_wait_cluster(smem_u32(&bars[par]), (uint32_t)((i - 1) >> 1) & 1u);  // k-th use of this parity's barrier
        // re-arm this parity's barrier for iteration i+2: no peer can send that record before it has received
        // this CTA's record of iteration i+1, which is sent after the next __syncthreads
        if (tid == 0 && i + 2 < M) mbar_arm(smem_u32(&bars[par]), TX_BYTES);
      } else {
        cluster_sync_all();
      }
      const int src3 = pick_lane(lane < CS ? cta_d[par][lane] : 0u, lane < CS ? cta_r[par][lane].x : NO_TIE, dmax);
      win = cta_r[par][src3];
    }
    }
    if (dmax != 0u) {  // otherwise every remaining distance is 0: the reference repeats the previous pick
      const uint32_t t = __brev(win.x) & mask;
      cur = (int)(((win.x & ((1u << (32 - nbits)) - 1u)) << nbits) | t);
      cx = __uint_as_float(win.y); cy = __uint_as_float(win.z); cz = __uint_as_float(win.w);
    }
    if (rank == 0 && tid == 0) {
      if (idx64) idx64[(int64_t)cloud * M + i] = cur;
      if (idx32) idx32[(int64_t)cloud * M + i] = cur;
      if (new_xyz) {
        float* o = new_xyz + (int64_t)cloud * 3 * M + i;
        o[0] = cx; o[M] = cy; o[2 * (int64_t)M] = cz;
      }
    }
  }
  if (CS > 1) cluster_sync_all();  // no CTA leaves while a peer may still address its shared memory
}

// ---- multi-pick rounds -----------------------------------------------------------------------------------------------------
// The kernel above pays one cluster-wide exchange (~1 300 cycles: two warp arg-maxes, st.async to 8 CTAs, an mbarrier
// wake-up) for EVERY pick.  Here one exchange serves several picks, and the result is still the reference's exact sequence:
//
//   round:  every warp publishes its K best points (distance, tie key, coordinates), best first, to every CTA;
//           every warp then replays the selection on the NW * K candidates it now holds (4 per lane):
//             - the best candidate is the next pick (always true for the first pick of a round: it is the old algorithm);
//             - it is applied to the other candidates (their min-distance shrinks exactly as their owners will compute it)
//               and to the warp's own register-resident points;
//             - the next best candidate is again the global arg-max PROVIDED no point that was not published can beat
//               it: every warp also publishes the distance of its best UNPUBLISHED point (its (K+1)-th best when the round
//               began), min-distances only shrink, so a candidate whose distance is strictly greater than
//               F = max_w (that distance) is safe.  The round ends at the first candidate that is not (or after 32 picks).
//
// Min-distances are the same fp32 values in the same order (min is order-independent, the squared distance is computed by
// the same sqdist_ref from the same operands), the arg-max uses the same (distance, tie key) order, so picks are
// bit-identical to the one-pick-per-exchange kernel; with heavy ties (lattices) rounds simply shrink to one pick.
// Used for the shapes whose cluster has at most 32 warps (each lane then holds K = 4 of the <= 128 candidates).
template <int CS, int T, int PPT>
__global__ void __launch_bounds__(T, 1)
fps_multi_kernel(const float* __restrict__ pts, Strides3 st, int N, int M, int nbits, int64_t* __restrict__ idx64,
                 int32_t* __restrict__ idx32, float* __restrict__ new_xyz, const int32_t* __restrict__ n_var) {
  constexpr int W = T / 32, NW = CS * W, K = 4, NC = NW * K, CPL = (NC + 31) / 32;
  static_assert(NW <= 32 && PPT <= 32, "multi-pick FPS: at most 32 warps per cluster and 32 points per thread");
  if (n_var) {
    N = n_var[blockIdx.x / CS];
    if (N <= M) return;
    nbits = N <= 1 ? 0 : 32 - __clz(N - 1);
    nbits = min(9, max(4, nbits));
  }
  __shared__ uint32_t cand_d[2][NC];
  __shared__ __align__(16) uint4 cand_r[2][NC];   // {tie, x, y, z}
  __shared__ uint32_t cand_f[2][NW];              // every warp's best UNPUBLISHED distance (its (K+1)-th best)
  __shared__ __align__(8) uint64_t bars[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = (CS > 1) ? cluster_ctarank() : 0u;
  const int cloud = blockIdx.x / CS;
  const float* __restrict__ p = pts + (int64_t)cloud * st.b;
  const uint32_t mask = (1u << nbits) - 1u;

  float px[PPT], py[PPT], pz[PPT], md[PPT];
  const int jbase = (int)rank * T + tid;
  float cx = p[0], cy = p[st.c], cz = p[2 * st.c];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int j = jbase + k * CS * T;
    if (j < N) {
      px[k] = p[(int64_t)j * st.n];
      py[k] = p[(int64_t)j * st.n + st.c];
      pz[k] = p[(int64_t)j * st.n + 2 * st.c];
      md[k] = sqdist_ref(cx, cy, cz, px[k], py[k], pz[k]);   // first pick (index 0) applied
    } else {
      px[k] = py[k] = pz[k] = 0.f;
      md[k] = 0.f;
    }
  }
  int cur = 0;
  if (rank == 0 && tid == 0) {
    if (idx64) idx64[(int64_t)cloud * M] = 0;
    if (idx32) idx32[(int64_t)cloud * M] = 0;
    if (new_xyz) {
      float* o = new_xyz + (int64_t)cloud * 3 * M;
      o[0] = cx; o[M] = cy; o[2 * (int64_t)M] = cz;
    }
  }
  constexpr uint32_t TX_BYTES = NC * 20 + NW * 4;
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arm(smem_u32(&bars[0]), TX_BYTES);   // round 0
    mbar_arm(smem_u32(&bars[1]), TX_BYTES);   // round 1
  }
  if (CS > 1) cluster_sync_all(); else __syncthreads();

  int done = 1;   // picks made so far (identical in every thread of the cluster)
  // Results leave through a per-lane buffer: pick number q of the current window is kept by lane q (four selects per pick),
  // and warp 0 of rank 0 writes a window of up to 32 picks with coalesced stores.  The per-pick form -- one thread of one
  // warp computing five 64-bit addresses and storing inside a divergent region -- made that warp the slowest of the
  // cluster on EVERY pick, and every round waits for its slowest warp.
  const bool writer = rank == 0 && warp == 0;
  int wq = 0;                                  // picks buffered in this window (warp-uniform)
  int w_cur = 0;
  float w_x = 0.f, w_y = 0.f, w_z = 0.f;
  auto flush = [&]() {
    if (writer && lane < wq) {
      const int64_t at = (int64_t)cloud * M + (done - wq + lane);
      if (idx64) idx64[at] = w_cur;
      if (idx32) idx32[at] = w_cur;
      if (new_xyz) {
        float* o = new_xyz + (int64_t)cloud * 3 * M + (done - wq + lane);
        o[0] = w_x; o[M] = w_y; o[2 * (int64_t)M] = w_z;
      }
    }
    wq = 0;
  };
  for (uint32_t r = 0; done < M; ++r) {
    const uint32_t par = r & 1;
    // ---- this thread's two best points: (b1, k1) then (b2, k2), in (distance desc, k asc) order --------------------------
    float b1 = 0.f, b2 = 0.f;
    int k1 = -1, k2 = -1;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const float m = md[k];
      if (m > b1) { b2 = b1; k2 = k1; b1 = m; k1 = k; }
      else if (m > b2) { b2 = m; k2 = k; }
    }
    // ---- the warp's K best records, best first, pushed to every CTA of the cluster ---------------------------------------
    uint32_t taken = 0;
    int pops = 0;
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
      const int bj = k1 < 0 ? -1 : jbase + k1 * CS * T;
      const uint32_t tie = (bj < 0) ? NO_TIE : (__brev((uint32_t)bj & mask) | ((uint32_t)bj >> nbits));
      uint32_t dmax;
      const int src = pick_lane(__float_as_uint(b1), tie, dmax);
      float bx = 0.f, by = 0.f, bz = 0.f;
      if (lane == src) pick_coords<PPT>(k1, px, py, pz, bx, by, bz);
      uint4 rec;
      rec.x = __shfl_sync(FULL, tie, src);
      rec.y = __shfl_sync(FULL, __float_as_uint(bx), src);
      rec.z = __shfl_sync(FULL, __float_as_uint(by), src);
      rec.w = __shfl_sync(FULL, __float_as_uint(bz), src);
      if (dmax == 0u) rec.x = NO_TIE;   // nothing left in this warp: an empty record
      if (lane < CS) {
        const uint32_t slot = (rank * W + warp) * K + j;
        const uint32_t dst_bar = mapa_u32(smem_u32(&bars[par]), (uint32_t)lane);
        st_async_v4(mapa_u32(smem_u32(&cand_r[par][slot]), (uint32_t)lane), rec, dst_bar);
        st_async_u32(mapa_u32(smem_u32(&cand_d[par][slot]), (uint32_t)lane), dmax, dst_bar);
      }
      // pop the winner's point: its second best moves up; a lane popped twice rescans its points
      bool need = false;
      if (lane == src && k1 >= 0) {
        taken |= 1u << k1;
        if (pops == 0) { b1 = b2; k1 = k2; pops = 1; }
        else need = true;
      }
      if (__any_sync(FULL, need)) {
        float nb = 0.f;
        int nk = -1;
#pragma unroll
        for (int k = 0; k < PPT; ++k)
          if (!((taken >> k) & 1u) && md[k] > nb) { nb = md[k]; nk = k; }
        if (need) { b1 = nb; k1 = nk; }
      }
    }
    {
      // the floor of the round: no unpublished point of this warp is farther than its best remaining one (min-distances
      // only shrink).  Publishing it -- instead of taking the K-th record as the bound -- keeps the K-th record pickable and
      // lowers the floor: on evenly spread inputs (the output of a previous FPS) 3.1 -> 4.4 picks per round.
      const uint32_t fl = __reduce_max_sync(FULL, __float_as_uint(b1));
      if (lane < CS)
        st_async_u32(mapa_u32(smem_u32(&cand_f[par][rank * W + warp]), (uint32_t)lane), fl,
                     mapa_u32(smem_u32(&bars[par]), (uint32_t)lane));
    }
    mbar_wait_cluster(smem_u32(&bars[par]), (r >> 1) & 1u);
    if (tid == 0) mbar_arm(smem_u32(&bars[par]), TX_BYTES);   // for round r + 2 (see the one-pick kernel for why this is safe)
    // ---- replay: lane l holds candidates l, l + 32, ... --------------------------------------------------------------------
    // cd = current min-distance bits (0 = empty or already picked: a zero distance can never be a live pick), ct = ~tie (the
    // low half of the selection key, so that a smaller tie wins), cr = the record as published
    uint32_t cd[CPL], ct[CPL];
    uint4 cr[CPL];
    const uint32_t floor_bits = lane < NW ? cand_f[par][lane] : 0u;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int slot = lane + 32 * c;
      if (slot < NC) {
        cd[c] = cand_d[par][slot];
        cr[c] = cand_r[par][slot];
        if (cr[c].x == NO_TIE) cd[c] = 0u;
      } else {
        cd[c] = 0u;
        cr[c] = make_uint4(NO_TIE, 0u, 0u, 0u);
      }
      ct[c] = ~cr[c].x;
    }
    const uint32_t F = __reduce_max_sync(FULL, floor_bits);
    // A pick is taken while the best candidate's distance is strictly above `lim`: 0 for the first pick of a round (always
    // valid -- unless every distance is 0, the reference's "repeat the previous pick" case, handled at the loop's exit),
    // F afterwards.  A round takes at most 32 picks: one window of the output buffer, flushed once per round.
    uint32_t lim = 0u;
    const int cap = min(M - done, 32);
    // The update of this thread's own points with pick i (25 independent distance / min pairs, ~175 instructions) does not
    // feed the selection of pick i + 1 (that runs on the candidates), so it is applied ONE PICK LATE, at the top of the next
    // trip: it then shares a basic block with the arg-max / redux / ballot chain of the next selection and fills that
    // chain's stalls instead of sitting serially behind it.  (cx, cy, cz) always holds the last pick; re-applying it is
    // idempotent, which keeps the update unconditional (first trip of a round: the previous round's last pick again).
#pragma unroll 1
    for (wq = 0; wq < cap;) {
#pragma unroll
      for (int k = 0; k < PPT; ++k) md[k] = fminf(md[k], sqdist_ref(cx, cy, cz, px[k], py[k], pz[k]));
      // this lane's best candidate by (distance desc, tie asc) as ONE 64-bit key: the two-condition form compiled into a
      // tree of divergent branches (9 BRA + 3 BSSY / BSYNC per pick), the key compare into selects
      unsigned long long key[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) key[c] = ((unsigned long long)cd[c] << 32) | ct[c];
      unsigned long long lk;
      int lc;
      if (CPL == 4) {                          // two levels instead of three dependent compares
        const bool b01 = key[1] > key[0], b23 = key[3 % CPL] > key[2 % CPL];
        const unsigned long long k01 = b01 ? key[1] : key[0], k23 = b23 ? key[3 % CPL] : key[2 % CPL];
        const bool hi = k23 > k01;
        lk = hi ? k23 : k01;
        lc = hi ? (b23 ? 3 : 2) : (b01 ? 1 : 0);
      } else {
        lk = key[0];
        lc = 0;
#pragma unroll
        for (int c = 1; c < CPL; ++c) {
          const bool better = key[c] > lk;
          lk = better ? key[c] : lk;
          lc = better ? c : lc;
        }
      }
      const uint32_t ld = (uint32_t)(lk >> 32), lt = ~(uint32_t)lk;
      uint32_t dmax;
      const int src = pick_lane(ld, ld ? lt : NO_TIE, dmax);
      if (!(dmax > lim)) {                   // an unpublished point could rank before this candidate: next round
        if (wq == 0) {                         // first pick and every distance is 0: the previous pick is repeated
          w_cur = lane == 0 ? cur : w_cur;
          w_x = lane == 0 ? cx : w_x;
          w_y = lane == 0 ? cy : w_y;
          w_z = lane == 0 ? cz : w_z;
          wq = 1;
        }
        break;
      }
      lim = F;
      {
        uint4 win = cr[0];
#pragma unroll
        for (int c = 1; c < CPL; ++c)
          if (lc == c) win = cr[c];
        win.x = __shfl_sync(FULL, win.x, src);
        win.y = __shfl_sync(FULL, win.y, src);
        win.z = __shfl_sync(FULL, win.z, src);
        win.w = __shfl_sync(FULL, win.w, src);
        const uint32_t t = __brev(win.x) & mask;
        cur = (int)(((win.x & ((1u << (32 - nbits)) - 1u)) << nbits) | t);
        cx = __uint_as_float(win.y); cy = __uint_as_float(win.z); cz = __uint_as_float(win.w);
        // the pick leaves the candidate set; the others see their min-distance shrink like their owners will compute it
#pragma unroll
        // (a picked or empty candidate keeps distance 0 under the update: min(0, d) = 0, no marker needed)
        for (int c = 0; c < CPL; ++c) {
          if (lane == src && lc == c) cd[c] = 0u;
          const float d = sqdist_ref(cx, cy, cz, __uint_as_float(cr[c].y), __uint_as_float(cr[c].z), __uint_as_float(cr[c].w));
          cd[c] = __float_as_uint(fminf(__uint_as_float(cd[c]), d));
        }
      }
      {
        const bool mine = lane == wq;
        w_cur = mine ? cur : w_cur;
        w_x = mine ? cx : w_x;
        w_y = mine ? cy : w_y;
        w_z = mine ? cz : w_z;
      }
      ++wq;
    }
    done += wq;
    flush();
    // the round's last pick reaches the own points before the next round's scan
#pragma unroll
    for (int k = 0; k < PPT; ++k) md[k] = fminf(md[k], sqdist_ref(cx, cy, cz, px[k], py[k], pz[k]));
  }
  if (CS > 1) cluster_sync_all();
}

template <int CS, int T, int PPT>
int launch_fps_multi(const float* pts, Strides3 st, int B, int N, int M, int nbits, int64_t* idx64, int32_t* idx32,
                     float* new_xyz, const int32_t* n_var, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RN_CUDA(cudaFuncSetAttribute(fps_multi_kernel<CS, T, PPT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                               (int)cudaSharedmemCarveoutMaxShared));
  RN_CUDA(cudaLaunchKernelEx(&cfg, fps_multi_kernel<CS, T, PPT>, pts, st, N, M, nbits, idx64, idx32, new_xyz, n_var));
  return REGNET_OK;
}

template <int CS, int T, int PPT>
int launch_fps(const float* pts, Strides3 st, int B, int N, int M, int nbits, int64_t* idx64, int32_t* idx32,
               float* new_xyz, const int32_t* n_var, bool mbar, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // Ask for the maximum shared-memory carve-out although this kernel needs ~2 KB: the L1/shared split of an SM can
  // only change while the SM is empty, and the tcgen05 GEMM CTAs that should share the SM with this (latency-bound,
  // minutes-of-cycles-long) kernel need ~217 KB.  With the default small carve-out they could not become resident.
  if (mbar || CS == 1) {
    RN_CUDA(cudaFuncSetAttribute(fps_kernel<CS, T, PPT, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 (int)cudaSharedmemCarveoutMaxShared));
    RN_CUDA(cudaLaunchKernelEx(&cfg, fps_kernel<CS, T, PPT, true>, pts, st, N, M, nbits, idx64, idx32, new_xyz, n_var));
  } else {
    RN_CUDA(cudaFuncSetAttribute(fps_kernel<CS, T, PPT, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 (int)cudaSharedmemCarveoutMaxShared));
    RN_CUDA(cudaLaunchKernelEx(&cfg, fps_kernel<CS, T, PPT, false>, pts, st, N, M, nbits, idx64, idx32, new_xyz, n_var));
  }
  return REGNET_OK;
}

// Instantiated launch shapes: the ones the policy below and the ScoreNet plan's co-running FPS select, with the point counts
// per thread they actually use, plus two barrier.cluster instances for A/B measurements.  Any other (cluster, threads)
// request is served by fps_generic_kernel (same indices) -- the library used to carry 176 instances.
struct FpsShape { int cs, t, ppt; bool mbar; };
#define RN_FPS_SHAPES(X)                                                                                          \
  X(1, 256, 1, true) X(1, 512, 1, true) X(1, 512, 2, true) X(1, 512, 4, true)                                        \
  X(4, 256, 4, true) X(4, 256, 8, true) X(4, 256, 16, true)                                                          \
  X(8, 128, 4, true) X(8, 128, 8, true) X(8, 128, 16, true) X(8, 128, 20, true) X(8, 128, 25, true) X(8, 128, 32, true) \
  X(8, 512, 16, true) X(4, 128, 1, true) X(4, 128, 2, true) X(4, 128, 4, true)                                       \
  X(8, 256, 16, true) X(4, 256, 25, true) X(4, 512, 16, true) X(8, 128, 32, false) X(8, 512, 16, false)

// shapes with a multi-pick instance (clusters of at most 32 warps); REGNET_FPS_SINGLE=1 keeps one pick per exchange
#define RN_FPS_MULTI_SHAPES(X)                                                                                      \
  X(8, 128, 4) X(8, 128, 8) X(8, 128, 16) X(8, 128, 20) X(8, 128, 25) X(8, 128, 32)                                    \
  X(4, 256, 4) X(4, 256, 8) X(4, 256, 16) X(4, 256, 25) X(4, 128, 1) X(4, 128, 2) X(4, 128, 4)

int dispatch_shape(int cs, int t, int ppt, bool mbar, const float* pts, Strides3 st, int B, int N, int M, int nbits,
                   int64_t* idx64, int32_t* idx32, float* new_xyz, const int32_t* n_var, cudaStream_t stream,
                   bool single_pick) {
  if (cs == 1) mbar = true;   // a single CTA exchanges nothing: one variant
  static const int forced = getenv("REGNET_FPS_SINGLE") ? atoi(getenv("REGNET_FPS_SINGLE")) : -1;   // 1 / 0 force one way
  const bool single = forced >= 0 ? forced != 0 : single_pick;
  if (mbar && !single) {
    int bestm = 1 << 30;
#define RN_FPS_MPICK(CS, T, PPT) if (cs == CS && t == T && PPT >= ppt && PPT < bestm) bestm = PPT;
    RN_FPS_MULTI_SHAPES(RN_FPS_MPICK)
#undef RN_FPS_MPICK
#define RN_FPS_MGO(CS, T, PPT)                                                                                      \
    if (cs == CS && t == T && bestm == PPT)                                                                            \
      return launch_fps_multi<CS, T, PPT>(pts, st, B, N, M, nbits, idx64, idx32, new_xyz, n_var, stream);
    RN_FPS_MULTI_SHAPES(RN_FPS_MGO)
#undef RN_FPS_MGO
  }
  // smallest instantiated point count >= ppt for this (cluster, threads, exchange)
  int best = 1 << 30;
#define RN_FPS_PICK(CS, T, PPT, MB) if (cs == CS && t == T && mbar == MB && PPT >= ppt && PPT < best) best = PPT;
  RN_FPS_SHAPES(RN_FPS_PICK)
#undef RN_FPS_PICK
#define RN_FPS_GO(CS, T, PPT, MB)                                                                                  \
  if (cs == CS && t == T && mbar == MB && best == PPT)                                                               \
    return launch_fps<CS, T, PPT>(pts, st, B, N, M, nbits, idx64, idx32, new_xyz, n_var, MB, stream);
  RN_FPS_SHAPES(RN_FPS_GO)
#undef RN_FPS_GO
  return -1;   // no such instance
}

// ---- any size: the min-distance array in global memory ---------------------------------------------------------------------
// Clouds beyond the register-resident limit (> 65 536 points), or launch shapes without an instance: one CTA of 1024
// threads per cloud, the reference's scheme (sampling_kernel.cu:47-117) with the same squared distance and the same tie
// order -- (max distance, min brev(j mod block) | j div block) -- so the indices are bit-identical here too.
__global__ void __launch_bounds__(1024, 1)
fps_generic_kernel(const float* __restrict__ pts, Strides3 st, int N, int M, int nbits, float* __restrict__ mind,
                   int64_t* __restrict__ idx64, int32_t* __restrict__ idx32, float* __restrict__ new_xyz) {
  const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ p = pts + (int64_t)cloud * st.b;
  float* __restrict__ md = mind + (int64_t)cloud * N;
  const uint32_t mask = (1u << nbits) - 1u;
  __shared__ uint32_t sd[2][32], stie[2][32];
  for (int j = tid; j < N; j += 1024) md[j] = __int_as_float(0x7f800000);
  int cur = 0;
  if (tid == 0) {
    if (idx64) idx64[(int64_t)cloud * M] = 0;
    if (idx32) idx32[(int64_t)cloud * M] = 0;
    if (new_xyz) {
      float* o = new_xyz + (int64_t)cloud * 3 * M;
      o[0] = p[0]; o[M] = p[st.c]; o[2 * (int64_t)M] = p[2 * st.c];
    }
  }
  __syncthreads();
  for (int i = 1; i < M; ++i) {
    const int par = i & 1;
    const float cx = p[(int64_t)cur * st.n], cy = p[(int64_t)cur * st.n + st.c], cz = p[(int64_t)cur * st.n + 2 * st.c];
    float best = 0.f;
    int bj = -1;
    for (int j = tid; j < N; j += 1024) {   // 1024 is a multiple of the reference block: one slot, ascending j
      const float d = sqdist_ref(cx, cy, cz, p[(int64_t)j * st.n], p[(int64_t)j * st.n + st.c], p[(int64_t)j * st.n + 2 * st.c]);
      const float m = fminf(md[j], d);
      md[j] = m;
      if (m > best) { best = m; bj = j; }
    }
    const uint32_t tie = (bj < 0) ? NO_TIE : (__brev((uint32_t)bj & mask) | ((uint32_t)bj >> nbits));
    uint32_t dmax;
    const int src1 = pick_lane(__float_as_uint(best), tie, dmax);
    if (lane == src1) { sd[par][warp] = dmax; stie[par][warp] = tie; }
    __syncthreads();
    const int src2 = pick_lane(sd[par][lane], stie[par][lane], dmax);
    const uint32_t wtie = stie[par][src2];
    if (dmax != 0u) cur = (int)(((wtie & ((1u << (32 - nbits)) - 1u)) << nbits) | (__brev(wtie) & mask));
    if (tid == 0) {
      if (idx64) idx64[(int64_t)cloud * M + i] = cur;
      if (idx32) idx32[(int64_t)cloud * M + i] = cur;
      if (new_xyz) {
        float* o = new_xyz + (int64_t)cloud * 3 * M + i;
        o[0] = p[(int64_t)cur * st.n]; o[M] = p[(int64_t)cur * st.n + st.c]; o[2 * (int64_t)M] = p[(int64_t)cur * st.n + 2 * st.c];
      }
    }
  }
}

int fps_generic_launch(const float* pts, Strides3 st, int B, int N, int M, int nbits, int64_t* idx64, int32_t* idx32,
                       float* new_xyz, cudaStream_t stream) {
  float* mind = nullptr;
  RN_CUDA(cudaMallocAsync(&mind, sizeof(float) * (size_t)B * N, stream));
  fps_generic_kernel<<<B, 1024, 0, stream>>>(pts, st, N, M, nbits, mind, idx64, idx32, new_xyz);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(mind, stream);
  if (e != cudaSuccess) return cuda_fail(e, "fps_generic_kernel");
  return REGNET_OK;
}

}  // namespace


int fps_block_log2(int N) {  // sampling_kernel.cu:32-40 get_block + the switch's floor of 16
  int cnt = 0;
  int64_t x = (int64_t)N - 1;
  while (x > 0) { x >>= 1; ++cnt; }
  if (cnt > 9) cnt = 9;
  if (cnt < 4) cnt = 4;
  return cnt;
}

// `threads` selects the CTA size (0 = auto); a NEGATIVE value selects the barrier.cluster exchange variant with
// |threads| threads (kept for A/B measurements against the st.async + mbarrier exchange).
static int fps_launch_impl(const float* pts, Strides3 st, int B, int N, int M, int64_t* idx64, int32_t* idx32,
                           float* new_xyz, const int32_t* n_var, int cluster_size, int threads, cudaStream_t stream,
                           bool single_pick = false);

int fps_launch(const float* pts, Strides3 st, int B, int N, int M, int64_t* idx64, int32_t* idx32, float* new_xyz,
               int cluster_size, int threads, cudaStream_t stream, bool single_pick) {
  RN_CHECK_ARG(N >= M, "farthest_point_sample: num_points (%d) must be >= num_centroids (%d)", N, M);
  return fps_launch_impl(pts, st, B, N, M, idx64, idx32, new_xyz, nullptr, cluster_size, threads, stream, single_pick);
}

// per-cloud point counts n_per_cloud[b] <= Nmax (device array); clouds with n <= M are skipped (output untouched)
int fps_launch_var(const float* pts, Strides3 st, int B, int Nmax, int M, const int32_t* n_per_cloud, int32_t* idx32,
                   cudaStream_t stream) {
  RN_CHECK_ARG(n_per_cloud != nullptr, "farthest_point_sample: null per-cloud counts");
  return fps_launch_impl(pts, st, B, Nmax, M, nullptr, idx32, nullptr, n_per_cloud, 0, 0, stream);
}

static int fps_launch_impl(const float* pts, Strides3 st, int B, int N, int M, int64_t* idx64, int32_t* idx32,
                           float* new_xyz, const int32_t* n_var, int cluster_size, int threads, cudaStream_t stream,
                           bool single_pick) {
  RN_CHECK_ARG(B > 0 && N > 0, "farthest_point_sample: empty input (B=%d, N=%d)", B, N);
  RN_CHECK_ARG(M > 0, "farthest_point_sample: num_centroids must be > 0 (got %d)", M);
  RN_CHECK_ARG(idx64 || idx32, "farthest_point_sample: no index output");
  bool mbar = true;
  if (threads < 0) { mbar = false; threads = -threads; }
  // auto policy from the B=15 sweep on B200 (profiles/r01_fps_sweep.txt): the per-iteration cost is latency, not
  // arithmetic, so small clouds stay in one CTA (no DSMEM round trip) and large ones spread over 8 CTAs
  if (cluster_size == 0 && threads == 0) {
    // debug knob for overlap experiments: REGNET_FPS_FORCE="<cluster>,<threads>" (negative threads = barrier variant)
    if (const char* f = getenv("REGNET_FPS_FORCE")) {
      int c = 0, t = 0;
      if (sscanf(f, "%d,%d", &c, &t) == 2 && N > 2048) { cluster_size = c; threads = t; if (t < 0) { mbar = false; threads = -t; } }
    }
  }
  if (cluster_size == 0 && threads == 0) {
    // (re-measured with the warp-push exchange, profiles/r01_fps_sweep_push.txt: 25 600 points 8x128 4.19 ms vs 8x512
    // 4.52 ms; 5 120 points 4x256 0.58 ms vs 8x256 0.70 ms -- both clusters of 32 warps, the push exchange's limit)
    if (N <= 512) { cluster_size = 1; threads = 256; }
    else if (N <= 2048) { cluster_size = 4; threads = 128; }   // multi-pick rounds: 1 024 -> 256 in 0.07 ms (one CTA: 0.12)
    else if (N <= 12288) { cluster_size = 4; threads = 256; }
    else if (N <= 32768) { cluster_size = 8; threads = 128; }   // 32 register-resident points per thread at most
    else { cluster_size = 8; threads = 512; }
  }
  if (cluster_size == 0) cluster_size = (N > 2048) ? 8 : 1;
  if (threads == 0) threads = 512;
  RN_CHECK_ARG(cluster_size == 1 || cluster_size == 2 || cluster_size == 4 || cluster_size == 8,
               "farthest_point_sample: cluster_size must be 1, 2, 4 or 8");
  RN_CHECK_ARG(threads == 128 || threads == 256 || threads == 512 || threads == 1024,
               "farthest_point_sample: threads must be 128, 256, 512 or 1024");
  // the in-thread tie order relies on CS*T being a multiple of the reference block size (<= 512): only
  // CS=1,T=256 with more than 256 points violates it
  const int nbits = fps_block_log2(N);
  if ((cluster_size * threads) % (1 << nbits) != 0) threads = 512;
  // grow the cluster until the cloud fits in registers; beyond that (and for per-cloud counts never) the generic kernel
  const int max_ppt = threads <= 256 ? 32 : threads <= 512 ? 16 : 8;
  while (cluster_size < 8 && ceil_div(N, cluster_size * threads) > max_ppt) cluster_size *= 2;
  int ppt = ceil_div(N, cluster_size * threads);
  if (ppt > max_ppt && !n_var) {
    // try the two largest register-resident shapes before giving up on registers
    if (ceil_div(N, 8 * 512) <= 16) { cluster_size = 8; threads = 512; ppt = ceil_div(N, 8 * 512); }
    else return fps_generic_launch(pts, st, B, N, M, nbits, idx64, idx32, new_xyz, stream);
  }
  if (ppt > max_ppt && n_var) {
    set_error("farthest_point_sample: N=%d exceeds the register-resident limit of %d points per cloud", N,
              8 * threads * max_ppt);
    return REGNET_ELIMIT;
  }
  const int rc = dispatch_shape(cluster_size, threads, ppt, mbar, pts, st, B, N, M, nbits, idx64, idx32, new_xyz, n_var, stream,
                                single_pick);
  if (rc >= 0) return rc;
  if (!n_var) return fps_generic_launch(pts, st, B, N, M, nbits, idx64, idx32, new_xyz, stream);
  set_error("farthest_point_sample: no instance for cluster %d x %d threads x %d points", cluster_size, threads, ppt);
  return REGNET_ELIMIT;
}

}  // namespace regnet
