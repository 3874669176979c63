// Double-precision instantiations of the three search operators.
//
// The reference dispatches its kernels over float AND double (AT_DISPATCH_FLOATING_TYPES, sampling_kernel.cu:149,
// ball_query_kernel.cu:112, interpolate_kernel.cu:113); REGNet itself only ever passes float32, which is what the
// tuned kernels of this library serve.  These kernels make the operator surface complete for float64 callers: plain
// one-thread-per-centroid / one-CTA-per-cloud restatements with the arithmetic done in double in the reference's
// order (DMUL(dy,dy), DFMA(dx,dx,.), DFMA(dz,dz,.) -- oracle/_ref/pn2_ext_ref.sass.txt), the same tie rules and the
// same "first K in index order" / "earliest index wins" semantics, so indices and squared distances are what the
// reference's double kernels return.  Not a hot path: no tiling, no tuning.
#include "internal.cuh"

namespace regnet {

namespace {

constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ double sqdist_f64(double x1, double y1, double z1, double x2, double y2, double z2) {
  const double dx = __dsub_rn(x2, x1), dy = __dsub_rn(y2, y1), dz = __dsub_rn(z2, z1);
  double t = __dmul_rn(dy, dy);
  t = __fma_rn(dx, dx, t);
  t = __fma_rn(dz, dz, t);
  return t;
}

// (max distance, min tie) over a warp, plain shuffles (64-bit keys)
__device__ __forceinline__ void warp_argmax(double& d, uint32_t& tie) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(FULLM, d, o);
    const uint32_t ot = __shfl_xor_sync(FULLM, tie, o);
    if (od > d || (od == d && ot < tie)) { d = od; tie = ot; }
  }
}

__global__ void __launch_bounds__(1024, 1)
fps_f64_kernel(const double* __restrict__ pts, Strides3 st, int N, int M, int nbits, double* __restrict__ mind,
               int64_t* __restrict__ idx64) {
  const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* __restrict__ p = pts + (int64_t)cloud * st.b;
  double* __restrict__ md = mind + (int64_t)cloud * N;
  const uint32_t mask = (1u << nbits) - 1u;
  __shared__ double sd[2][32];
  __shared__ uint32_t stie[2][32];
  for (int j = tid; j < N; j += 1024) md[j] = -1.0;   // the reference's "not yet measured" marker (sampling_kernel.cu:74)
  int cur = 0;
  if (tid == 0) idx64[(int64_t)cloud * M] = 0;
  __syncthreads();
  for (int i = 1; i < M; ++i) {
    const int par = i & 1;
    const double cx = p[(int64_t)cur * st.n], cy = p[(int64_t)cur * st.n + st.c], cz = p[(int64_t)cur * st.n + 2 * st.c];
    double best = 0.0;
    int bj = -1;
    for (int j = tid; j < N; j += 1024) {
      const double d = sqdist_f64(cx, cy, cz, p[(int64_t)j * st.n], p[(int64_t)j * st.n + st.c], p[(int64_t)j * st.n + 2 * st.c]);
      const double old = md[j];
      const double m = (old < 0.0 || d < old) ? d : old;
      md[j] = m;
      if (m > best) { best = m; bj = j; }
    }
    uint32_t tie = (bj < 0) ? 0xffffffffu : (__brev((uint32_t)bj & mask) | ((uint32_t)bj >> nbits));
    warp_argmax(best, tie);
    if (lane == 0) { sd[par][warp] = best; stie[par][warp] = tie; }
    __syncthreads();
    double d2 = sd[par][lane];
    uint32_t t2 = stie[par][lane];
    warp_argmax(d2, t2);
    if (d2 > 0.0) cur = (int)(((t2 & ((1u << (32 - nbits)) - 1u)) << nbits) | (__brev(t2) & mask));
    if (tid == 0) idx64[(int64_t)cloud * M + i] = cur;
  }
}

__global__ void __launch_bounds__(128)
ball_query_f64_kernel(const double* __restrict__ pts, Strides3 pst, const double* __restrict__ ctr, Strides3 cst, int N,
                      int M, double radius, int K, int64_t* __restrict__ index, int64_t* __restrict__ count) {
  const int b = blockIdx.y, m = blockIdx.x * 128 + threadIdx.x;
  if (m >= M) return;
  const double* __restrict__ p = pts + (int64_t)b * pst.b;
  const double* __restrict__ c = ctr + (int64_t)b * cst.b;
  const double r2 = __dmul_rn(radius, radius);
  const double x1 = c[(int64_t)m * cst.n], y1 = c[(int64_t)m * cst.n + cst.c], z1 = c[(int64_t)m * cst.n + 2 * cst.c];
  const int64_t o = ((int64_t)b * M + m) * K;
  int cnt = 0, first = 0;
  for (int j = 0; j < N && cnt < K; ++j) {
    if (sqdist_f64(x1, y1, z1, p[(int64_t)j * pst.n], p[(int64_t)j * pst.n + pst.c], p[(int64_t)j * pst.n + 2 * pst.c]) < r2) {
      if (cnt == 0) first = j;
      index[o + cnt++] = j;
    }
  }
  for (int k = cnt; k < K; ++k) index[o + k] = first;
  count[(int64_t)b * M + m] = cnt;
}

__global__ void __launch_bounds__(128)
three_nn_f64_kernel(const double* __restrict__ qry, Strides3 qst, const double* __restrict__ key, Strides3 kst, int Nq,
                    int Nk, int64_t* __restrict__ index, double* __restrict__ dist) {
  const int b = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
  if (i >= Nq) return;
  const double* __restrict__ q = qry + (int64_t)b * qst.b;
  const double* __restrict__ kp = key + (int64_t)b * kst.b;
  const double x1 = q[(int64_t)i * qst.n], y1 = q[(int64_t)i * qst.n + qst.c], z1 = q[(int64_t)i * qst.n + 2 * qst.c];
  // interpolate_kernel.cu:49-50 in double: {1e40, 0, 0} and {-1, 0, 0}
  double d0 = 1e40, d1 = 0.0, d2 = 0.0;
  int i0 = -1, i1 = 0, i2 = 0;
  for (int j = 0; j < Nk; ++j) {
    // the reference computes (x1 - x2)^2 + ...: same squares as (x2 - x1)^2, same order of accumulation
    const double d = sqdist_f64(x1, y1, z1, kp[(int64_t)j * kst.n], kp[(int64_t)j * kst.n + kst.c], kp[(int64_t)j * kst.n + 2 * kst.c]);
    if (d < d0) { d2 = d1; i2 = i1; d1 = d0; i1 = i0; d0 = d; i0 = j; }
    else if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
    else if (d < d2) { d2 = d; i2 = j; }
  }
  const int64_t o = ((int64_t)b * Nq + i) * 3;
  index[o] = i0; index[o + 1] = i1; index[o + 2] = i2;
  dist[o] = d0; dist[o + 1] = d1; dist[o + 2] = d2;
}

}  // namespace

}  // namespace regnet

using namespace regnet;

extern "C" {

int regnet_farthest_point_sample_f64(const double* points, int64_t sb, int64_t sc, int64_t sn, int B, int N, int M,
                                     int64_t* index, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RN_CHECK_ARG(points && index, "farthest_point_sample_f64: null argument");
  RN_CHECK_ARG(B > 0 && N > 0, "farthest_point_sample: empty input (B=%d, N=%d)", B, N);
  RN_CHECK_ARG(M > 0 && N >= M, "farthest_point_sample: num_points (%d) must be >= num_centroids (%d) > 0", N, M);
  double* mind = nullptr;
  RN_CUDA(cudaMallocAsync(&mind, sizeof(double) * (size_t)B * N, stream));
  fps_f64_kernel<<<B, 1024, 0, stream>>>(points, Strides3{sb, sc, sn}, N, M, fps_block_log2(N), mind, index);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(mind, stream);
  if (e != cudaSuccess) return cuda_fail(e, "fps_f64_kernel");
  return REGNET_OK;
}

int regnet_ball_query_f64(const double* points, int64_t psb, int64_t psc, int64_t psn, const double* centroids, int64_t csb,
                          int64_t csc, int64_t csn, int B, int N, int M, double radius, int K, int64_t* index,
                          int64_t* count, void* stream_) {
  RN_CHECK_ARG(points && centroids && index && count, "ball_query_f64: null argument");
  RN_CHECK_ARG(B > 0 && N > 0 && M > 0 && K > 0, "ball_query: empty input (B=%d, N=%d, M=%d, K=%d)", B, N, M, K);
  ball_query_f64_kernel<<<dim3(ceil_div(M, 128), B), 128, 0, (cudaStream_t)stream_>>>(
      points, Strides3{psb, psc, psn}, centroids, Strides3{csb, csc, csn}, N, M, radius, K, index, count);
  RN_LAUNCH_CHECK("ball_query_f64_kernel");
  return REGNET_OK;
}

int regnet_point_search_f64(const double* query, int64_t qsb, int64_t qsc, int64_t qsn, const double* key, int64_t ksb,
                            int64_t ksc, int64_t ksn, int B, int Nq, int Nk, int k, int64_t* index, double* distance,
                            void* stream_) {
  RN_CHECK_ARG(query && key && index && distance, "point_search_f64: null argument");
  RN_CHECK_ARG(k == 3, "point_search: num_neighbours must be 3 (got %d)", k);
  RN_CHECK_ARG(B > 0 && Nq > 0, "point_search: empty input (B=%d, Nq=%d)", B, Nq);
  RN_CHECK_ARG(Nk >= 3, "point_search: needs at least 3 key points (got %d)", Nk);
  three_nn_f64_kernel<<<dim3(ceil_div(Nq, 128), B), 128, 0, (cudaStream_t)stream_>>>(
      query, Strides3{qsb, qsc, qsn}, key, Strides3{ksb, ksc, ksn}, Nq, Nk, index, distance);
  RN_LAUNCH_CHECK("three_nn_f64_kernel");
  return REGNET_OK;
}

}  // extern "C"
