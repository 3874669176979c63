// extern "C" surface of libregnet_b200.so, section 1 + 3 of include/regnet_b200.h (pn2_ext operator ABI and the
// stand-alone MLP layer).  Section 2 (the fused ScoreNet plan) lives in scorenet.cu.
#include <stdlib.h>
#include <mutex>
#include <string>

#include "internal.cuh"

namespace regnet {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  (void)cudaGetLastError();
  return REGNET_ECUDA;
}

int* oob_flag() {
  static std::mutex mu;
  static int* flags[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!flags[dev]) {
    if (cudaMalloc(&flags[dev], sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flags[dev], 0, sizeof(int));
  }
  return flags[dev];
}

// 128 -> 1 conv with bias + BN + sigmoid (pointnet2.py:82-84,117-119): one warp per row.
__global__ void __launch_bounds__(256)
score_head_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ shift, int64_t P, int cin, float* __restrict__ score) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < P; r += warps) {
    float acc = 0.f;
    for (int c = lane; c < cin; c += 32) acc = fmaf(X[r * ldx + c], w[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float v = fmaf(acc, scale ? scale[0] : 1.f, shift ? shift[0] : 0.f);
      score[r] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v)));
    }
  }
}

int score_head_launch(const float* X, int ldx, const float* w, const float* scale, const float* shift, int64_t P,
                      int cin, float* score, cudaStream_t stream) {
  if (P == 0) return REGNET_OK;
  int64_t blocks = (P + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  RN_PREFER_MAX_SMEM(score_head_kernel);
  score_head_kernel<<<(int)blocks, 256, 0, stream>>>(X, ldx, w, scale, shift, P, cin, score);
  RN_LAUNCH_CHECK("score_head_kernel");
  return REGNET_OK;
}

}  // namespace regnet

using namespace regnet;

extern "C" {

const char* regnet_last_error(void) { return g_err.c_str(); }

int regnet_abi_version(void) { return 2; }

int regnet_device_arch(int* sm_major, int* sm_minor, int* sm_count) {
  int dev = 0;
  RN_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  RN_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_major) *sm_major = p.major;
  if (sm_minor) *sm_minor = p.minor;
  if (sm_count) *sm_count = p.multiProcessorCount;
  return REGNET_OK;
}

int regnet_check_index_errors(void) {
  int* f = oob_flag();
  if (!f) return REGNET_OK;
  int h = 0;
  RN_CUDA(cudaMemcpy(&h, f, sizeof(int), cudaMemcpyDeviceToHost));
  if (h) {
    cudaMemset(f, 0, sizeof(int));
    set_error("index out of range in a gather/scatter operator (the reference asserts on device here)");
    return REGNET_EINVAL;
  }
  return REGNET_OK;
}

int regnet_farthest_point_sample_ex(const float* points, int64_t sb, int64_t sc, int64_t sn, int B, int N, int M,
                                    int64_t* index64, int32_t* index32, float* new_xyz, int cluster_size,
                                    int threads, void* stream) {
  RN_CHECK_ARG(points != nullptr, "farthest_point_sample: null points");
  return fps_launch(points, Strides3{sb, sc, sn}, B, N, M, index64, index32, new_xyz, cluster_size, threads,
                    (cudaStream_t)stream);
}

int regnet_farthest_point_sample(const float* points, int64_t sb, int64_t sc, int64_t sn, int B, int N, int M,
                                 int64_t* index, float* new_xyz, void* stream) {
  return regnet_farthest_point_sample_ex(points, sb, sc, sn, B, N, M, index, nullptr, new_xyz, 0, 0, stream);
}

int regnet_ball_query(const float* points, int64_t psb, int64_t psc, int64_t psn, const float* centroids,
                      int64_t csb, int64_t csc, int64_t csn, int B, int N, int M, float radius, int K,
                      int64_t* index, int64_t* count, int32_t* index32, void* stream) {
  RN_CHECK_ARG(points && centroids, "ball_query: null input");
  return ball_query_launch(points, Strides3{psb, psc, psn}, centroids, Strides3{csb, csc, csn}, B, N, M, radius, K,
                           index, count, index32, (cudaStream_t)stream);
}

// Grid-accelerated forms of the two search operators (same results, see grid.cu): the caller lends
// regnet_search_workspace_bytes(B, N) bytes of device scratch.  They apply when the uniform grid pays off and its
// limits hold; otherwise (or with a null workspace) the call falls through to the brute-force scan.
int64_t regnet_search_workspace_bytes(int B, int N) { return grid_workspace_bytes(B, N); }

static bool grid_applies(int n_indexed, int K, const void* ws, int64_t ws_bytes, int B) {
  return ws != nullptr && K == 64 && n_indexed >= 4096 && n_indexed <= 65536 &&
         ws_bytes >= grid_workspace_bytes(B, n_indexed) && !getenv("REGNET_API_BRUTE");
}

int regnet_ball_query_ws(const float* points, int64_t psb, int64_t psc, int64_t psn, const float* centroids,
                         int64_t csb, int64_t csc, int64_t csn, int B, int N, int M, float radius, int K,
                         int64_t* index, int64_t* count, void* workspace, int64_t workspace_bytes, void* stream) {
  RN_CHECK_ARG(points && centroids && index && count, "ball_query: null argument");
  if (!grid_applies(N, K, workspace, workspace_bytes, B) || !(radius > 0.f))
    return regnet_ball_query(points, psb, psc, psn, centroids, csb, csc, csn, B, N, M, radius, K, index, count, nullptr,
                             stream);
  RN_CHECK_ARG(B > 0 && N > 0 && M > 0, "ball_query: empty input");
  const Strides3 pst{psb, psc, psn}, cst{csb, csc, csn};
  RN_TRY(grid_build_launch(points, pst, B, N, radius * 1.001f + 1e-7f, workspace, (cudaStream_t)stream));
  return ball_query_grid_launch(points, pst, centroids, cst, B, N, M, radius, workspace, nullptr, (cudaStream_t)stream,
                                index, count);
}

int regnet_point_search_ws(const float* query, int64_t qsb, int64_t qsc, int64_t qsn, const float* key, int64_t ksb,
                           int64_t ksc, int64_t ksn, int B, int Nq, int Nk, int k, int64_t* index, float* distance,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  RN_CHECK_ARG(query && key && index && distance, "point_search: null argument");
  if (k != 3 || !grid_applies(Nk, 64, workspace, workspace_bytes, B))
    return regnet_point_search(query, qsb, qsc, qsn, key, ksb, ksc, ksn, B, Nq, Nk, k, index, distance, stream);
  RN_CHECK_ARG(B > 0 && Nq > 0, "point_search: empty input");
  const Strides3 qst{qsb, qsc, qsn}, kst{ksb, ksc, ksn};
  RN_TRY(grid_build_launch(key, kst, B, Nk, 0.f, workspace, (cudaStream_t)stream));
  return three_nn_grid_launch(query, qst, key, kst, B, Nq, Nk, workspace, nullptr, nullptr, (cudaStream_t)stream, index,
                              distance);
}

int regnet_group_points_forward(const float* input, int64_t sb, int64_t sc, int64_t sn, const int64_t* index,
                                int B, int C, int N, int M, int K, float* out, void* stream) {
  RN_CHECK_ARG(input && index && out, "group_points_forward: null argument");
  return group_forward_launch(input, Strides3{sb, sc, sn}, index, B, C, N, M, K, out, oob_flag(), (cudaStream_t)stream);
}

int regnet_group_points_backward(const float* grad_out, const int64_t* index, int B, int C, int N, int M, int K,
                                 float* grad_in, void* stream) {
  RN_CHECK_ARG(grad_out && index && grad_in, "group_points_backward: null argument");
  return group_backward_launch(grad_out, index, B, C, N, M, K, grad_in, oob_flag(), (cudaStream_t)stream);
}

int regnet_point_search(const float* query, int64_t qsb, int64_t qsc, int64_t qsn, const float* key, int64_t ksb,
                        int64_t ksc, int64_t ksn, int B, int Nq, int Nk, int k, int64_t* index, float* distance,
                        void* stream) {
  RN_CHECK_ARG(query && key, "point_search: null input");
  RN_CHECK_ARG(k == 3, "point_search: only 3 neighbours are supported (got %d)", k);  // interpolate_kernel.cu:101
  return three_nn_launch(query, Strides3{qsb, qsc, qsn}, key, Strides3{ksb, ksc, ksn}, B, Nq, Nk, index, distance,
                         nullptr, nullptr, (cudaStream_t)stream);
}

int regnet_interpolate_forward(const float* input, int64_t sb, int64_t sc, int64_t sn, const int64_t* index,
                               const float* weight, int B, int C, int Ns, int Nd, float* out, void* stream) {
  RN_CHECK_ARG(input && index && weight && out, "interpolate_forward: null argument");
  return interp_forward_launch(input, Strides3{sb, sc, sn}, index, weight, B, C, Ns, Nd, out, oob_flag(),
                               (cudaStream_t)stream);
}

int regnet_interpolate_backward(const float* grad_out, const int64_t* index, const float* weight, int B, int C,
                                int Ns, int Nd, float* grad_in, void* stream) {
  RN_CHECK_ARG(grad_out && index && weight && grad_in, "interpolate_backward: null argument");
  return interp_backward_launch(grad_out, index, weight, B, C, Ns, Nd, grad_in, oob_flag(), (cudaStream_t)stream);
}

int regnet_gather_knn_forward(const float* input, int64_t sb, int64_t sc, int64_t sn, const int64_t* index, int B,
                              int C, int N, int M, int K, float* out, void* stream) {
  return regnet_group_points_forward(input, sb, sc, sn, index, B, C, N, M, K, out, stream);
}

int regnet_gather_knn_backward(const float* grad_out, const int64_t* index, int B, int C, int N, int M, int K,
                               float* grad_in, void* stream) {
  return regnet_group_points_backward(grad_out, index, B, C, N, M, K, grad_in, stream);
}

int regnet_mlp_layer(const float* X, const float* W, const float* scale, const float* shift, int64_t P, int cin,
                     int cout, int pool, int act, int engine, float* Y, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RN_CHECK_ARG(X && W && Y, "mlp_layer: null argument");
  RN_CHECK_ARG(P > 0 && cin > 0 && cout > 0, "mlp_layer: empty problem");
  RN_CHECK_ARG(pool == 0 || pool == 64, "mlp_layer: pool must be 0 or 64");
  RN_CHECK_ARG(engine == REGNET_ENGINE_TC || engine == REGNET_ENGINE_SIMT, "mlp_layer: unknown engine %d", engine);
  const int kpad = round_up(cin, 16);
  Epilogue ep;
  ep.scale = scale; ep.shift = shift; ep.act = act; ep.pool = pool;
  // pooled epilogues need scale >= 0: negate the weight rows whose scale is negative (same products, see internal.cuh)
  float* abs_scale = nullptr;
  const float* row_sign = (pool && scale) ? scale : nullptr;
  if (row_sign) {
    RN_CUDA(cudaMalloc(&abs_scale, sizeof(float) * cout));
    int rc0 = abs_copy_launch(scale, abs_scale, cout, stream);
    if (rc0) { cudaFree(abs_scale); return rc0; }
    ep.scale = abs_scale;
  }
  const int ldo = round_up(cout, 4);
  float* ytmp = nullptr;  // padded output when cout % 4 != 0
  const int64_t out_rows = pool ? P / pool : P;
  if (ldo != cout) RN_CUDA(cudaMalloc(&ytmp, sizeof(float) * (size_t)out_rows * ldo));
  ep.out_f32 = ytmp ? ytmp : Y;
  ep.ld_f32 = ldo;
  int rc = REGNET_OK;
  void* scratch = nullptr;
  if (engine == REGNET_ENGINE_SIMT) {
    const size_t bytes = sizeof(float) * ((size_t)P + cout) * kpad;
    RN_CUDA(cudaMalloc(&scratch, bytes));
    float* Xp = (float*)scratch;
    float* Wp = Xp + (size_t)P * kpad;
    rc = split_rows_launch(X, P, cin, cin, kpad, nullptr, nullptr, Xp, stream);
    if (!rc) rc = split_rows_launch(W, cout, cin, cin, kpad, nullptr, nullptr, Wp, stream, 0, row_sign);
    if (!rc) rc = gemm_simt_launch(Xp, kpad, Wp, kpad, P, kpad, cout, ep, stream);
  } else {
    const size_t bytes = 2 * sizeof(__nv_bfloat16) * ((size_t)P + cout) * kpad;
    RN_CUDA(cudaMalloc(&scratch, bytes));
    __nv_bfloat16* Xh = (__nv_bfloat16*)scratch;
    __nv_bfloat16* Xl = Xh + (size_t)P * kpad;
    __nv_bfloat16* Wh = Xl + (size_t)P * kpad;
    __nv_bfloat16* Wl = Wh + (size_t)cout * kpad;
    rc = split_rows_launch(X, P, cin, cin, kpad, Xh, Xl, nullptr, stream);
    if (!rc) rc = split_rows_launch(W, cout, cin, cin, kpad, Wh, Wl, nullptr, stream, 0, row_sign);
    if (!rc) rc = gemm_tc_launch(Xh, Xl, kpad, Wh, Wl, kpad, P, cin, cout, ep, stream);
  }
  if (!rc && ytmp) {
    cudaError_t e = cudaMemcpy2DAsync(Y, sizeof(float) * cout, ytmp, sizeof(float) * ldo, sizeof(float) * cout,
                                      (size_t)out_rows, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy2DAsync");
  }
  cudaError_t e = cudaStreamSynchronize(stream);
  if (!rc && e != cudaSuccess) rc = cuda_fail(e, "mlp_layer kernels");
  if (scratch) cudaFree(scratch);
  if (ytmp) cudaFree(ytmp);
  if (abs_scale) cudaFree(abs_scale);
  return rc;
}

// One 1x1-conv / linear layer with folded BatchNorm on the tcgen05 engine, operands and results as caller-owned planes:
// the non-allocating, non-synchronising form of regnet_mlp_layer (used by the region / refine heads, region_heads.py).
int regnet_linear_planes(const void* x_hi, const void* x_lo, int ldx, int64_t P, int K, const void* w_hi, const void* w_lo,
                         int ldw, int cout, const float* scale, const float* shift, int act, float* out_f32, int ld_f32,
                         void* out_hi, void* out_lo, int ld_split, void* stream) {
  RN_CHECK_ARG(x_hi && x_lo && w_hi && w_lo, "linear_planes: null operand");
  RN_CHECK_ARG(out_f32 || (out_hi && out_lo), "linear_planes: no output");
  RN_CHECK_ARG(P >= 0 && K > 0 && cout > 0, "linear_planes: empty problem");
  Epilogue ep;
  ep.scale = scale; ep.shift = shift; ep.act = act; ep.pool = 0;
  ep.out_f32 = out_f32; ep.ld_f32 = ld_f32;
  ep.out_hi = (__nv_bfloat16*)out_hi; ep.out_lo = (__nv_bfloat16*)out_lo; ep.ld_split = ld_split;
  return gemm_tc_launch((const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, ldx, (const __nv_bfloat16*)w_hi,
                        (const __nv_bfloat16*)w_lo, ldw, P, K, cout, ep, (cudaStream_t)stream);
}

int regnet_sa0_chain(const float* pc, const float* new_xyz, const int32_t* nbr, int B, int N, int M, const float* W0,
                     const float* scale0, const float* shift0, const float* W1, const float* scale1,
                     const float* shift1, const float* W2, const float* scale2, const float* shift2, float* out,
                     float* dbg, int variant, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RN_CHECK_ARG(pc && new_xyz && nbr && W0 && W1 && W2 && out, "sa0_chain: null argument");
  RN_CHECK_ARG(B > 0 && N > 0 && M > 0, "sa0_chain: empty problem");
  RN_CHECK_ARG(gemm_tc_supported(), "sa0_chain: tensor maps are not available from this driver");
  void* scratch = nullptr;
  const size_t nw = (size_t)(128 + 256) * 128;
  RN_CUDA(cudaMalloc(&scratch, 2 * sizeof(__nv_bfloat16) * nw + sizeof(float) * 256));
  float* abs2 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + 2 * sizeof(__nv_bfloat16) * nw);
  __nv_bfloat16* w1h = (__nv_bfloat16*)scratch;
  __nv_bfloat16* w1l = w1h + 128 * 128;
  __nv_bfloat16* w2h = w1l + 128 * 128;
  __nv_bfloat16* w2l = w2h + 256 * 128;
  int rc = split_rows_launch(W1, 128, 128, 128, 128, w1h, w1l, nullptr, stream);
  if (!rc) rc = split_rows_launch(W2, 256, 128, 128, 128, w2h, w2l, nullptr, stream, 0, scale2);   // pooled layer: scale >= 0
  if (!rc) rc = abs_copy_launch(scale2, abs2, 256, stream);
  if (!rc)
    rc = sa0_chain_launch(pc, Strides3{(int64_t)N * 6, 1, 6}, new_xyz, pc + 3, (int64_t)N * 6, 6, nbr, W0, 6, scale0,
                          shift0, w1h, w1l, 128, scale1, shift1, w2h, w2l, 128, abs2, shift2, B, M, out, 256, nullptr, nullptr, dbg,
                          nullptr, variant, stream);
  cudaError_t e = cudaStreamSynchronize(stream);
  if (!rc && e != cudaSuccess) rc = cuda_fail(e, "sa0_chain kernel");
  cudaFree(scratch);
  return rc;
}

}  // extern "C"
