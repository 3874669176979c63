// Farthest point sampling for sm_100a: one thread-block CLUSTER per cloud, points and running min-distances
// resident in registers for the whole kernel, argmax by redux.sync + DSMEM exchange.
//
// Replaces csrc/sampling_kernel.cu:47-170 of the reference (grid = B blocks of <= 512 threads, min-distance
// array in global memory, 10-level shared-memory tree per iteration).  Results are bit-identical, including the
// tie rule that the reference's strided scan + tree reduction implies:
//   among the points attaining the maximum min-distance, the winner is the one whose slot t = j mod block has
//   the smallest bit-reversed value (log2(block) bits), and within a slot the smallest j;
//   block = min(nextpow2(N), 512), at least 16;  if the maximum is 0 the previous pick is repeated.
// That rule is folded into one 32-bit key  tie(j) = brev(j mod block) | (j div block)  so that the argmax is
// (max distance bits, min tie) -- two redux.sync per level instead of a 64-bit shuffle tree.
#include "common.cuh"

namespace regnet {

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct __align__(16) FpsRecord {
  uint32_t dbits;  // bits of the (non-negative) distance: unsigned order == float order
  uint32_t tie;    // smaller wins
  float x, y;
  float z;
  uint32_t pad[3];
};
static_assert(sizeof(FpsRecord) == 32, "record is two 16-byte words");

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
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Warp-wide argmax of (dbits, -tie); on return every lane holds the winner's key and coordinates.
__device__ __forceinline__ void warp_pick(uint32_t& dbits, uint32_t& tie, float& x, float& y, float& z) {
  const uint32_t wmax = __reduce_max_sync(FULL, dbits);
  const uint32_t mine = (dbits == wmax) ? tie : 0xffffffffu;
  const uint32_t wtie = __reduce_min_sync(FULL, mine);
  const unsigned vote = __ballot_sync(FULL, dbits == wmax && mine == wtie);
  const int src = __ffs(vote) - 1;
  x = __shfl_sync(FULL, x, src);
  y = __shfl_sync(FULL, y, src);
  z = __shfl_sync(FULL, z, src);
  dbits = wmax;
  tie = wtie;
}

template <int CS, int T, int PPT>
__global__ void __launch_bounds__(T, 1)
fps_kernel(const float* __restrict__ pts, Strides3 st, int N, int M, int nbits, int64_t* __restrict__ idx64,
           int32_t* __restrict__ idx32, float* __restrict__ new_xyz) {
  constexpr int W = T / 32;
  __shared__ FpsRecord warp_rec[2][W];
  __shared__ FpsRecord cta_rec[2][CS];  // slot r is written by cluster rank r (through DSMEM when r != me)

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
  if (CS > 1) cluster_sync_all();  // peers' shared memory must exist before the first remote store

  for (int i = 1; i < M; ++i) {
    const int par = i & 1;
    float best = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
    int bj = -1;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const float d = sqdist_ref(cx, cy, cz, px[k], py[k], pz[k]);
      const float m = fminf(md[k], d);
      md[k] = m;
      if (m > best) { best = m; bj = jbase + k * CS * T; bx = px[k]; by = py[k]; bz = pz[k]; }
    }
    uint32_t dbits = __float_as_uint(best);
    uint32_t tie = (bj < 0) ? 0xffffffffu : (__brev((uint32_t)bj & mask) | ((uint32_t)bj >> nbits));
    warp_pick(dbits, tie, bx, by, bz);
    if (lane == 0) {
      FpsRecord r;
      r.dbits = dbits; r.tie = tie; r.x = bx; r.y = by; r.z = bz; r.pad[0] = r.pad[1] = r.pad[2] = 0;
      warp_rec[par][warp] = r;
    }
    __syncthreads();
    {
      FpsRecord r;
      if (lane < W) r = warp_rec[par][lane];
      else { r.dbits = 0; r.tie = 0xffffffffu; r.x = r.y = r.z = 0.f; }
      dbits = r.dbits; tie = r.tie; bx = r.x; by = r.y; bz = r.z;
      warp_pick(dbits, tie, bx, by, bz);
    }
    if (CS > 1) {
      if (warp == 0 && lane < CS) {
        const uint32_t dst = mapa_u32(smem_u32(&cta_rec[par][rank]), (uint32_t)lane);
        st_cluster_v4(dst, dbits, tie, __float_as_uint(bx), __float_as_uint(by));
        st_cluster_u32(dst + 16, __float_as_uint(bz));
      }
      cluster_sync_all();
      FpsRecord r;
      if (lane < CS) r = cta_rec[par][lane];
      else { r.dbits = 0; r.tie = 0xffffffffu; r.x = r.y = r.z = 0.f; }
      dbits = r.dbits; tie = r.tie; bx = r.x; by = r.y; bz = r.z;
      warp_pick(dbits, tie, bx, by, bz);
    }
    if (dbits != 0u) {  // otherwise every remaining distance is 0: the reference repeats the previous pick
      const uint32_t t = __brev(tie) & mask;
      cur = (int)(((tie & ((1u << (32 - nbits)) - 1u)) << nbits) | t);
      cx = bx; cy = by; cz = bz;
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
}

template <int CS, int T, int PPT>
int launch_fps(const float* pts, Strides3 st, int B, int N, int M, int nbits, int64_t* idx64, int32_t* idx32,
               float* new_xyz, cudaStream_t stream) {
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
  RN_CUDA(cudaLaunchKernelEx(&cfg, fps_kernel<CS, T, PPT>, pts, st, N, M, nbits, idx64, idx32, new_xyz));
  return REGNET_OK;
}

template <int CS, int T>
int dispatch_ppt(int ppt, const float* pts, Strides3 st, int B, int N, int M, int nbits, int64_t* idx64,
                 int32_t* idx32, float* new_xyz, cudaStream_t stream) {
  if (ppt <= 1) return launch_fps<CS, T, 1>(pts, st, B, N, M, nbits, idx64, idx32, new_xyz, stream);
  if (ppt <= 2) return launch_fps<CS, T, 2>(pts, st, B, N, M, nbits, idx64, idx32, new_xyz, stream);
  if (ppt <= 4) return launch_fps<CS, T, 4>(pts, st, B, N, M, nbits, idx64, idx32, new_xyz, stream);
  if (ppt <= 8) return launch_fps<CS, T, 8>(pts, st, B, N, M, nbits, idx64, idx32, new_xyz, stream);
  if constexpr (T == 512) {
    if (ppt <= 16) return launch_fps<CS, T, 16>(pts, st, B, N, M, nbits, idx64, idx32, new_xyz, stream);
  }
  set_error("farthest_point_sample: %d points per thread exceeds the register-resident limit", ppt);
  return REGNET_ELIMIT;
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

int fps_launch(const float* pts, Strides3 st, int B, int N, int M, int64_t* idx64, int32_t* idx32, float* new_xyz,
               int cluster_size, int threads, cudaStream_t stream) {
  RN_CHECK_ARG(B > 0 && N > 0, "farthest_point_sample: empty input (B=%d, N=%d)", B, N);
  RN_CHECK_ARG(M > 0, "farthest_point_sample: num_centroids must be > 0 (got %d)", M);
  RN_CHECK_ARG(N >= M, "farthest_point_sample: num_points (%d) must be >= num_centroids (%d)", N, M);
  RN_CHECK_ARG(idx64 || idx32, "farthest_point_sample: no index output");
  if (cluster_size == 0) cluster_size = (N > 8192) ? 8 : (N > 2048) ? 4 : (N > 1024) ? 2 : 1;
  if (threads == 0) threads = 512;
  RN_CHECK_ARG(cluster_size == 1 || cluster_size == 2 || cluster_size == 4 || cluster_size == 8,
               "farthest_point_sample: cluster_size must be 1, 2, 4 or 8");
  RN_CHECK_ARG(threads == 512 || threads == 1024, "farthest_point_sample: threads must be 512 or 1024");
  // grow the cluster until the cloud fits in registers
  const int max_ppt = threads == 512 ? 16 : 8;
  while (cluster_size < 8 && ceil_div(N, cluster_size * threads) > max_ppt) cluster_size *= 2;
  const int ppt = ceil_div(N, cluster_size * threads);
  if (ppt > max_ppt) {
    set_error("farthest_point_sample: N=%d exceeds the register-resident limit of %d points per cloud", N,
              8 * threads * max_ppt);
    return REGNET_ELIMIT;
  }
  const int nbits = fps_block_log2(N);
#define RN_FPS_CASE(CS, T)                                                                        \
  if (cluster_size == CS && threads == T)                                                         \
    return dispatch_ppt<CS, T>(ppt, pts, st, B, N, M, nbits, idx64, idx32, new_xyz, stream);
  RN_FPS_CASE(1, 512) RN_FPS_CASE(2, 512) RN_FPS_CASE(4, 512) RN_FPS_CASE(8, 512)
  RN_FPS_CASE(1, 1024) RN_FPS_CASE(2, 1024) RN_FPS_CASE(4, 1024) RN_FPS_CASE(8, 1024)
#undef RN_FPS_CASE
  set_error("farthest_point_sample: unreachable configuration");
  return REGNET_EINVAL;
}

}  // namespace regnet
