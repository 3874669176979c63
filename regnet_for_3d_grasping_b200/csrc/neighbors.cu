// Ball query and 3-NN search for sm_100a.
//
// Both replace "one thread per centroid, grid = B blocks, AoS global loads" kernels of the reference
// (csrc/ball_query_kernel.cu:31-74, csrc/interpolate_kernel.cu:28-77) with: grid = (centroid tiles, B) so all
// 148 SMs are busy, candidate points staged once per CTA into shared memory as float4 and read back as
// one broadcast LDS.128 per candidate per warp, distances with the reference's exact rounding
// (common.cuh sqdist_ref), results staged in shared memory and written out coalesced.
// The scan stays in ascending index order per centroid, which is what makes the outputs bit-identical
// ("first K in index order", "earliest index wins ties").
#include "common.cuh"

namespace regnet {

namespace {

constexpr int BQ_THREADS = 128;   // centroids per CTA
constexpr int TILE_PTS = 1024;    // candidates staged per round (16 KB as float4)

__device__ __forceinline__ void stage_points(float4* tile, const float* __restrict__ p, Strides3 st, int base,
                                             int n, int N) {
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int64_t j = (int64_t)(base + t) * st.n;
    float4 v;
    v.x = p[j];
    v.y = p[j + st.c];
    v.z = p[j + 2 * st.c];
    v.w = 0.f;
    tile[t] = v;
  }
  (void)N;
}

// index (B,M,K): first K hits ascending, first hit replicated into the unused slots, zeros if none.
// hits[k * (BQ_THREADS+1) + tid] keeps the per-thread list conflict-free for both the scan (column access)
// and the coalesced write-out (row access).  HitT = uint16_t when N <= 65536 (halves shared memory => 6 CTAs/SM).
// The scan is unrolled by 4 with the (rare) hit handling kept in index order, so the common path is four
// independent LDS.128 + distance chains per trip.
template <typename HitT>
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(const float* __restrict__ pts, Strides3 pst, const float* __restrict__ ctr, Strides3 cst, int N,
                  int M, float radius, int K, int64_t* __restrict__ index, int64_t* __restrict__ count,
                  int32_t* __restrict__ index32) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);
  HitT* hits = reinterpret_cast<HitT*>(smem_raw + sizeof(float4) * TILE_PTS);
  constexpr int LD = BQ_THREADS + 1;

  const int b = blockIdx.y;
  const int m0 = blockIdx.x * BQ_THREADS;
  const int m = m0 + threadIdx.x;
  const bool live = m < M;
  const float* __restrict__ p = pts + (int64_t)b * pst.b;
  const float* __restrict__ c = ctr + (int64_t)b * cst.b;
  const float r2 = __fmul_rn(radius, radius);  // ball_query_kernel.cu:47, fp32
  float x1 = 0.f, y1 = 0.f, z1 = 0.f;
  if (live) {
    x1 = c[(int64_t)m * cst.n];
    y1 = c[(int64_t)m * cst.n + cst.c];
    z1 = c[(int64_t)m * cst.n + 2 * cst.c];
  }
  int cnt = live ? 0 : K;  // dead threads count as full so the block can leave early
  for (int base = 0; base < N; base += TILE_PTS) {
    const int n = min(TILE_PTS, N - base);
    __syncthreads();
    stage_points(tile, p, pst, base, n, N);
    // whole block leaves once every centroid is full (checked per tile; also the barrier after staging)
    if (__syncthreads_and(cnt >= K)) break;
    if (cnt < K) {
      int t = 0;
      for (; t + 4 <= n; t += 4) {
        const float4 q0 = tile[t], q1 = tile[t + 1], q2 = tile[t + 2], q3 = tile[t + 3];
        const bool h0 = sqdist_ref(x1, y1, z1, q0.x, q0.y, q0.z) < r2;
        const bool h1 = sqdist_ref(x1, y1, z1, q1.x, q1.y, q1.z) < r2;
        const bool h2 = sqdist_ref(x1, y1, z1, q2.x, q2.y, q2.z) < r2;
        const bool h3 = sqdist_ref(x1, y1, z1, q3.x, q3.y, q3.z) < r2;
        if (h0 | h1 | h2 | h3) {
          if (h0 && cnt < K) hits[(cnt++) * LD + threadIdx.x] = (HitT)(base + t);
          if (h1 && cnt < K) hits[(cnt++) * LD + threadIdx.x] = (HitT)(base + t + 1);
          if (h2 && cnt < K) hits[(cnt++) * LD + threadIdx.x] = (HitT)(base + t + 2);
          if (h3 && cnt < K) hits[(cnt++) * LD + threadIdx.x] = (HitT)(base + t + 3);
          if (cnt >= K) break;
        }
      }
      for (; t < n && cnt < K; ++t) {
        const float4 q = tile[t];
        if (sqdist_ref(x1, y1, z1, q.x, q.y, q.z) < r2) hits[(cnt++) * LD + threadIdx.x] = (HitT)(base + t);
      }
    }
  }
  if (live) {
    const HitT first = cnt > 0 ? hits[threadIdx.x] : (HitT)0;
    for (int k = cnt; k < K; ++k) hits[k * LD + threadIdx.x] = first;
    if (count) count[(int64_t)b * M + m] = cnt;
  }
  __syncthreads();
  const int rows = min(BQ_THREADS, M - m0);
  for (int e = threadIdx.x; e < rows * K; e += BQ_THREADS) {
    const int r = e / K, k = e - r * K;
    const int v = (int)hits[k * LD + r];
    const int64_t o = ((int64_t)b * M + m0 + r) * K + k;
    if (index) index[o] = v;
    if (index32) index32[o] = v;
  }
}

// 3-NN: index (B,Nq,3), squared distances (B,Nq,3); optional normalised inverse-distance weights
// (modules.py:117-122: inv = 1/clamp(d2, 1e-10); w = inv / (inv0+inv1+inv2)), all fp32 IEEE.
__global__ void __launch_bounds__(BQ_THREADS)
three_nn_kernel(const float* __restrict__ qry, Strides3 qst, const float* __restrict__ key, Strides3 kst, int Nq,
                int Nk, int64_t* __restrict__ index, float* __restrict__ dist, int32_t* __restrict__ index32,
                float* __restrict__ weight) {
  __shared__ float4 tile[TILE_PTS];
  const int b = blockIdx.y;
  const int i = blockIdx.x * BQ_THREADS + threadIdx.x;
  const bool live = i < Nq;
  const float* __restrict__ q = qry + (int64_t)b * qst.b;
  const float* __restrict__ kp = key + (int64_t)b * kst.b;
  float x1 = 0.f, y1 = 0.f, z1 = 0.f;
  if (live) {
    x1 = q[(int64_t)i * qst.n];
    y1 = q[(int64_t)i * qst.n + qst.c];
    z1 = q[(int64_t)i * qst.n + 2 * qst.c];
  }
  // interpolate_kernel.cu:49-50: `scalar_t min_dist[K] = {1e40}; int min_ind[K] = {-1};` -> {+inf,0,0}, {-1,0,0}
  float d0 = __int_as_float(0x7f800000), d1 = 0.f, d2 = 0.f;
  int i0 = -1, i1 = 0, i2 = 0;
  for (int base = 0; base < Nk; base += TILE_PTS) {
    const int n = min(TILE_PTS, Nk - base);
    __syncthreads();
    stage_points(tile, kp, kst, base, n, Nk);
    __syncthreads();
    if (live) {
      auto insert = [&](float d, int j) {
        if (d < d2) {                 // d2 is the current 3rd best once the list is warm
          if (d < d0) { d2 = d1; i2 = i1; d1 = d0; i1 = i0; d0 = d; i0 = j; }
          else if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
          else { d2 = d; i2 = j; }
        } else if (d < d0) {          // only reachable while the bogus zeros of the initialiser are present
          d2 = d1; i2 = i1; d1 = d0; i1 = i0; d0 = d; i0 = j;
        } else if (d < d1) {
          d2 = d1; i2 = i1; d1 = d; i1 = j;
        }
      };
      int t = 0;
      if (base == 0) {                // first three keys flush the initialiser; after that d0 <= d1 <= d2 holds
        for (; t < 3 && t < n; ++t) {
          const float4 c = tile[t];
          insert(sqdist_ref(x1, y1, z1, c.x, c.y, c.z), t);
        }
      }
      for (; t + 4 <= n; t += 4) {
        const float4 c0 = tile[t], c1 = tile[t + 1], c2 = tile[t + 2], c3 = tile[t + 3];
        const float e0 = sqdist_ref(x1, y1, z1, c0.x, c0.y, c0.z);
        const float e1 = sqdist_ref(x1, y1, z1, c1.x, c1.y, c1.z);
        const float e2 = sqdist_ref(x1, y1, z1, c2.x, c2.y, c2.z);
        const float e3 = sqdist_ref(x1, y1, z1, c3.x, c3.y, c3.z);
        if (fminf(fminf(e0, e1), fminf(e2, e3)) < d2) {   // sorted list: nothing below d2 => nothing to insert
          insert(e0, base + t);
          insert(e1, base + t + 1);
          insert(e2, base + t + 2);
          insert(e3, base + t + 3);
        }
      }
      for (; t < n; ++t) {
        const float4 c = tile[t];
        insert(sqdist_ref(x1, y1, z1, c.x, c.y, c.z), base + t);
      }
    }
  }
  if (live) {
    const int64_t o = ((int64_t)b * Nq + i) * 3;
    if (index) { index[o] = i0; index[o + 1] = i1; index[o + 2] = i2; }
    if (index32) { index32[o] = i0; index32[o + 1] = i1; index32[o + 2] = i2; }
    if (dist) { dist[o] = d0; dist[o + 1] = d1; dist[o + 2] = d2; }
    if (weight) {
      const float v0 = __fdiv_rn(1.0f, fmaxf(d0, 1e-10f));
      const float v1 = __fdiv_rn(1.0f, fmaxf(d1, 1e-10f));
      const float v2 = __fdiv_rn(1.0f, fmaxf(d2, 1e-10f));
      const float s = __fadd_rn(__fadd_rn(v0, v1), v2);
      weight[o] = __fdiv_rn(v0, s);
      weight[o + 1] = __fdiv_rn(v1, s);
      weight[o + 2] = __fdiv_rn(v2, s);
    }
  }
}

// Any K (the reference has no limit, ball_query_kernel.cu:31-74): hits go straight to the global index array, one thread
// per centroid, candidates staged through the same shared-memory tiles.  Used when K > 128 (the list of the kernel above
// no longer fits in shared memory).
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_anyk_kernel(const float* __restrict__ pts, Strides3 pst, const float* __restrict__ ctr, Strides3 cst, int N,
                       int M, float radius, int K, int64_t* __restrict__ index, int64_t* __restrict__ count,
                       int32_t* __restrict__ index32) {
  __shared__ float4 tile[TILE_PTS];
  const int b = blockIdx.y;
  const int m = blockIdx.x * BQ_THREADS + threadIdx.x;
  const bool live = m < M;
  const float* __restrict__ p = pts + (int64_t)b * pst.b;
  const float* __restrict__ c = ctr + (int64_t)b * cst.b;
  const float r2 = __fmul_rn(radius, radius);
  float x1 = 0.f, y1 = 0.f, z1 = 0.f;
  if (live) {
    x1 = c[(int64_t)m * cst.n];
    y1 = c[(int64_t)m * cst.n + cst.c];
    z1 = c[(int64_t)m * cst.n + 2 * cst.c];
  }
  const int64_t o = ((int64_t)b * M + m) * K;
  int cnt = live ? 0 : K, first = 0;
  for (int base = 0; base < N; base += TILE_PTS) {
    const int n = min(TILE_PTS, N - base);
    __syncthreads();
    stage_points(tile, p, pst, base, n, N);
    if (__syncthreads_and(cnt >= K)) break;
    for (int t = 0; t < n && cnt < K; ++t) {
      const float4 q = tile[t];
      if (sqdist_ref(x1, y1, z1, q.x, q.y, q.z) < r2) {
        if (cnt == 0) first = base + t;
        if (index) index[o + cnt] = base + t;
        if (index32) index32[o + cnt] = base + t;
        ++cnt;
      }
    }
  }
  if (live) {
    for (int k = cnt; k < K; ++k) {
      if (index) index[o + k] = first;
      if (index32) index32[o + k] = first;
    }
    if (count) count[(int64_t)b * M + m] = cnt;
  }
}

}  // namespace

int ball_query_launch(const float* pts, Strides3 pst, const float* ctr, Strides3 cst, int B, int N, int M,
                      float radius, int K, int64_t* index, int64_t* count, int32_t* index32, cudaStream_t stream) {
  RN_CHECK_ARG(B > 0 && N > 0 && M > 0, "ball_query: empty input (B=%d, N=%d, M=%d)", B, N, M);
  RN_CHECK_ARG(K > 0, "ball_query: num_neighbours must be > 0");
  dim3 grid(ceil_div(M, BQ_THREADS), B);
  if (K > 128) {   // the shared-memory hit list holds 128 entries per centroid: beyond that, hits go straight to global
    ball_query_anyk_kernel<<<grid, BQ_THREADS, 0, stream>>>(pts, pst, ctr, cst, N, M, radius, K, index, count, index32);
    RN_LAUNCH_CHECK("ball_query_anyk_kernel");
    return REGNET_OK;
  }
  // per-device attribute; cheap enough to set on every launch (keeps multi-device processes correct)
  if (N <= 65536) {
    const size_t smem = sizeof(float4) * TILE_PTS + sizeof(uint16_t) * (size_t)K * (BQ_THREADS + 1);
    RN_CUDA(cudaFuncSetAttribute(ball_query_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(float4) * TILE_PTS + sizeof(uint16_t) * 128 * (BQ_THREADS + 1))));
    RN_PREFER_MAX_SMEM(ball_query_kernel<uint16_t>);
    ball_query_kernel<uint16_t><<<grid, BQ_THREADS, smem, stream>>>(pts, pst, ctr, cst, N, M, radius, K, index, count, index32);
  } else {
    const size_t smem = sizeof(float4) * TILE_PTS + sizeof(int) * (size_t)K * (BQ_THREADS + 1);
    RN_CUDA(cudaFuncSetAttribute(ball_query_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(float4) * TILE_PTS + sizeof(int) * 128 * (BQ_THREADS + 1))));
    RN_PREFER_MAX_SMEM(ball_query_kernel<int>);
    ball_query_kernel<int><<<grid, BQ_THREADS, smem, stream>>>(pts, pst, ctr, cst, N, M, radius, K, index, count, index32);
  }
  RN_LAUNCH_CHECK("ball_query_kernel");
  return REGNET_OK;
}

int three_nn_launch(const float* qry, Strides3 qst, const float* key, Strides3 kst, int B, int Nq, int Nk,
                    int64_t* index, float* dist, int32_t* index32, float* weight, cudaStream_t stream) {
  RN_CHECK_ARG(B > 0 && Nq > 0, "point_search: empty input (B=%d, Nq=%d)", B, Nq);
  RN_CHECK_ARG(Nk >= 3, "point_search: needs at least 3 key points (got %d)", Nk);
  dim3 grid(ceil_div(Nq, BQ_THREADS), B);
  RN_PREFER_MAX_SMEM(three_nn_kernel);
  three_nn_kernel<<<grid, BQ_THREADS, 0, stream>>>(qry, qst, key, kst, Nq, Nk, index, dist, index32, weight);
  RN_LAUNCH_CHECK("three_nn_kernel");
  return REGNET_OK;
}

}  // namespace regnet
