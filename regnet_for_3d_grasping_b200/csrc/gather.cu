// Gather / scatter kernels.
//
// (1) The reference's stand-alone operators, same semantics and layouts:
//       group_points forward   csrc/grouping_kernel.cu:29-51   (ATen expand+gather there)
//       group_points backward  csrc/grouping_kernel.cu:54-149  (atomicAdd scatter)
//       interpolate forward    csrc/interpolate_kernel.cu:134-232
//       interpolate backward   csrc/interpolate_kernel.cu:239-337
// (2) The fused producers of the shared-MLP operand, which replace the reference's
//       group_points(xyz) - new_xyz ; group_points(feature) ; torch.cat      (modules.py:39-56)
//       feature_interpolate ; torch.cat([interp, dense])                      (modules.py:104-131)
//     by one pass that reads point-major features (rows are contiguous => coalesced row gathers) and writes the
//     MLP operand once, directly in the layout the GEMM engine wants (fp32 rows, or bf16 hi/lo planes).
#include "common.cuh"

namespace regnet {

namespace {

constexpr int THREADS = 256;

inline int grid_for(int64_t work, int threads = THREADS) {
  int64_t g = (work + threads - 1) / threads;
  const int64_t cap = 148LL * 32;  // grid-stride beyond this: whole waves of 148 SMs
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__global__ void __launch_bounds__(THREADS)
group_forward_kernel(const float* __restrict__ in, Strides3 st, const int64_t* __restrict__ index, int C, int N,
                     int M, int K, int64_t total, float* __restrict__ out, int* __restrict__ oob) {
  const int64_t MK = (int64_t)M * K;
  for (int64_t e = blockIdx.x * (int64_t)THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * THREADS) {
    const int64_t mk = e % MK;
    const int64_t bc = e / MK;
    const int c = (int)(bc % C);
    const int64_t b = bc / C;
    const int64_t j = index[b * MK + mk];
    if (j < 0 || j >= N) { *oob = 1; out[e] = 0.f; continue; }
    out[e] = in[b * st.b + c * st.c + j * st.n];
  }
}

__global__ void __launch_bounds__(THREADS)
group_backward_kernel(const float* __restrict__ gout, const int64_t* __restrict__ index, int C, int N, int M, int K,
                      int64_t total, float* __restrict__ gin, int* __restrict__ oob) {
  const int64_t MK = (int64_t)M * K;
  for (int64_t e = blockIdx.x * (int64_t)THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * THREADS) {
    const int64_t mk = e % MK;
    const int64_t bc = e / MK;
    const int64_t b = bc / C;
    const int64_t j = index[b * MK + mk];
    if (j < 0 || j >= N) { *oob = 1; continue; }
    atomicAdd(gin + bc * N + j, gout[e]);  // RED.ADD.F32, order-nondeterministic like the reference
  }
}

__global__ void __launch_bounds__(THREADS)
interp_forward_kernel(const float* __restrict__ in, Strides3 st, const int64_t* __restrict__ index,
                      const float* __restrict__ weight, int C, int Ns, int Nd, int64_t total, float* __restrict__ out,
                      int* __restrict__ oob) {
  for (int64_t e = blockIdx.x * (int64_t)THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * THREADS) {
    const int n = (int)(e % Nd);
    const int64_t bc = e / Nd;
    const int c = (int)(bc % C);
    const int64_t b = bc / C;
    const int64_t o = (b * Nd + n) * 3;
    const float* __restrict__ src = in + b * st.b + c * st.c;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t j = index[o + k];
      if (j < 0 || j >= Ns) { *oob = 1; continue; }
      acc = __fmaf_rn(src[j * st.n], weight[o + k], acc);  // interpolate_kernel.cu:165-170 after fma contraction
    }
    out[e] = acc;
  }
}

__global__ void __launch_bounds__(THREADS)
interp_backward_kernel(const float* __restrict__ gout, const int64_t* __restrict__ index,
                       const float* __restrict__ weight, int C, int Ns, int Nd, int64_t total,
                       float* __restrict__ gin, int* __restrict__ oob) {
  for (int64_t e = blockIdx.x * (int64_t)THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * THREADS) {
    const int n = (int)(e % Nd);
    const int64_t bc = e / Nd;
    const int64_t b = bc / C;
    const int64_t o = (b * Nd + n) * 3;
    const float g = gout[e];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t j = index[o + k];
      if (j < 0 || j >= Ns) { *oob = 1; continue; }
      atomicAdd(gin + bc * Ns + j, __fmul_rn(g, weight[o + k]));
    }
  }
}

// ---- fused operand producers ----------------------------------------------------------------------------------
// Output row r (one grouped position / one dense point) has `kpad` columns; two adjacent columns per thread.
struct OperandOut {
  float* f32;            // (rows, kpad) fp32, or nullptr
  __nv_bfloat16* hi;     // (rows, kpad) bf16 planes, or nullptr
  __nv_bfloat16* lo;
};

__device__ __forceinline__ void store_pair(const OperandOut& o, int64_t off, float a, float b) {
  if (o.f32) *reinterpret_cast<float2*>(o.f32 + off) = make_float2(a, b);
  if (o.hi) {
    __nv_bfloat16 ah, al, bh, bl;
    split_bf16(a, ah, al);
    split_bf16(b, bh, bl);
    *reinterpret_cast<__nv_bfloat162*>(o.hi + off) = __halves2bfloat162(ah, bh);
    *reinterpret_cast<__nv_bfloat162*>(o.lo + off) = __halves2bfloat162(al, bl);
  }
}

// SA operand: row (b,m,k) = [ xyz[j]-new_xyz[m] (3) | feat[j, 0..C) | 0 ... ], j = nbr[b,m,k]      (modules.py:44-52)
__global__ void __launch_bounds__(THREADS)
sa_operand_kernel(const float* __restrict__ xyz, Strides3 xst, const float* __restrict__ new_xyz,
                  const float* __restrict__ feat, int64_t feat_bstride, int feat_ld, int C,
                  const int32_t* __restrict__ nbr, int N, int M, int K, int kpad, int64_t total_pairs, OperandOut out) {
  const int half = kpad >> 1;
  for (int64_t e = blockIdx.x * (int64_t)THREADS + threadIdx.x; e < total_pairs; e += (int64_t)gridDim.x * THREADS) {
    const int cp = (int)(e % half);
    const int64_t row = e / half;
    const int64_t bm = row / K;
    const int m = (int)(bm % M);
    const int64_t b = bm / M;
    const int j = nbr[row];
    float v[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int col = cp * 2 + t;
      float x = 0.f;
      if (col < 3) {
        x = __fsub_rn(xyz[b * xst.b + col * xst.c + (int64_t)j * xst.n], new_xyz[(b * 3 + col) * M + m]);
      } else if (col < 3 + C) {
        x = feat[b * feat_bstride + (int64_t)j * feat_ld + (col - 3)];
      }
      v[t] = x;
    }
    store_pair(out, row * kpad + cp * 2, v[0], v[1]);
  }
  (void)N;
}

// FP operand: row (b,n) = [ sum_k w[b,n,k]*sparse[idx[b,n,k], 0..C2) | dense[n, 0..C1) | 0 ... ]   (modules.py:117-127)
__global__ void __launch_bounds__(THREADS)
fp_operand_kernel(const float* __restrict__ sparse, int64_t sparse_bstride, int sparse_ld, int C2,
                  const float* __restrict__ dense, int64_t dense_bstride, int dense_ld, int C1,
                  const int32_t* __restrict__ idx, const float* __restrict__ w, int Nd, int kpad, int64_t total_pairs,
                  OperandOut out) {
  const int half = kpad >> 1;
  for (int64_t e = blockIdx.x * (int64_t)THREADS + threadIdx.x; e < total_pairs; e += (int64_t)gridDim.x * THREADS) {
    const int cp = (int)(e % half);
    const int64_t row = e / half;
    const int n = (int)(row % Nd);
    const int64_t b = row / Nd;
    const int i0 = idx[row * 3], i1 = idx[row * 3 + 1], i2 = idx[row * 3 + 2];
    const float w0 = w[row * 3], w1 = w[row * 3 + 1], w2 = w[row * 3 + 2];
    const float* __restrict__ sp = sparse + b * sparse_bstride;
    float v[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int col = cp * 2 + t;
      float x = 0.f;
      if (col < C2) {
        x = __fmaf_rn(sp[(int64_t)i0 * sparse_ld + col], w0, 0.f);
        x = __fmaf_rn(sp[(int64_t)i1 * sparse_ld + col], w1, x);
        x = __fmaf_rn(sp[(int64_t)i2 * sparse_ld + col], w2, x);
      } else if (col < C2 + C1) {
        x = dense[b * dense_bstride + (int64_t)n * dense_ld + (col - C2)];
      }
      v[t] = x;
    }
    store_pair(out, row * kpad + cp * 2, v[0], v[1]);
  }
}

// fp32 rows -> bf16 hi/lo planes (used for weights and by regnet_mlp_layer)
__global__ void __launch_bounds__(THREADS)
split_rows_kernel(const float* __restrict__ src, int64_t rows, int cols, int ld_src, int kpad,
                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ f32) {
  const int64_t total = rows * kpad;
  for (int64_t e = blockIdx.x * (int64_t)THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * THREADS) {
    const int col = (int)(e % kpad);
    const int64_t r = e / kpad;
    const float x = col < cols ? src[r * ld_src + col] : 0.f;
    if (f32) f32[e] = x;
    if (hi) {
      __nv_bfloat16 h, l;
      split_bf16(x, h, l);
      hi[e] = h;
      lo[e] = l;
    }
  }
}

}  // namespace

int check_oob_flag(int* d_flag, const char* what, cudaStream_t stream);

int group_forward_launch(const float* in, Strides3 st, const int64_t* index, int B, int C, int N, int M, int K,
                         float* out, int* d_oob, cudaStream_t stream) {
  const int64_t total = (int64_t)B * C * M * K;
  if (total == 0) return REGNET_OK;
  group_forward_kernel<<<grid_for(total), THREADS, 0, stream>>>(in, st, index, C, N, M, K, total, out, d_oob);
  RN_LAUNCH_CHECK("group_forward_kernel");
  return REGNET_OK;
}

int group_backward_launch(const float* gout, const int64_t* index, int B, int C, int N, int M, int K, float* gin,
                          int* d_oob, cudaStream_t stream) {
  RN_CUDA(cudaMemsetAsync(gin, 0, sizeof(float) * (size_t)B * C * N, stream));
  const int64_t total = (int64_t)B * C * M * K;
  if (total == 0) return REGNET_OK;
  group_backward_kernel<<<grid_for(total), THREADS, 0, stream>>>(gout, index, C, N, M, K, total, gin, d_oob);
  RN_LAUNCH_CHECK("group_backward_kernel");
  return REGNET_OK;
}

int interp_forward_launch(const float* in, Strides3 st, const int64_t* index, const float* weight, int B, int C,
                          int Ns, int Nd, float* out, int* d_oob, cudaStream_t stream) {
  const int64_t total = (int64_t)B * C * Nd;
  if (total == 0) return REGNET_OK;
  interp_forward_kernel<<<grid_for(total), THREADS, 0, stream>>>(in, st, index, weight, C, Ns, Nd, total, out, d_oob);
  RN_LAUNCH_CHECK("interp_forward_kernel");
  return REGNET_OK;
}

int interp_backward_launch(const float* gout, const int64_t* index, const float* weight, int B, int C, int Ns,
                           int Nd, float* gin, int* d_oob, cudaStream_t stream) {
  RN_CUDA(cudaMemsetAsync(gin, 0, sizeof(float) * (size_t)B * C * Ns, stream));
  const int64_t total = (int64_t)B * C * Nd;
  if (total == 0) return REGNET_OK;
  interp_backward_kernel<<<grid_for(total), THREADS, 0, stream>>>(gout, index, weight, C, Ns, Nd, total, gin, d_oob);
  RN_LAUNCH_CHECK("interp_backward_kernel");
  return REGNET_OK;
}

int sa_operand_launch(const float* xyz, Strides3 xst, const float* new_xyz, const float* feat, int64_t feat_bstride,
                      int feat_ld, int C, const int32_t* nbr, int B, int N, int M, int K, int kpad, float* out_f32,
                      __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream) {
  RN_CHECK_ARG(kpad % 2 == 0 && kpad >= 3 + C, "sa_operand: bad kpad %d for C=%d", kpad, C);
  const int64_t total = (int64_t)B * M * K * (kpad / 2);
  OperandOut o{out_f32, out_hi, out_lo};
  sa_operand_kernel<<<grid_for(total), THREADS, 0, stream>>>(xyz, xst, new_xyz, feat, feat_bstride, feat_ld, C, nbr, N,
                                                             M, K, kpad, total, o);
  RN_LAUNCH_CHECK("sa_operand_kernel");
  return REGNET_OK;
}

int fp_operand_launch(const float* sparse, int64_t sparse_bstride, int sparse_ld, int C2, const float* dense,
                      int64_t dense_bstride, int dense_ld, int C1, const int32_t* idx, const float* w, int B, int Nd,
                      int kpad, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream) {
  RN_CHECK_ARG(kpad % 2 == 0 && kpad >= C1 + C2, "fp_operand: bad kpad %d for C1+C2=%d", kpad, C1 + C2);
  const int64_t total = (int64_t)B * Nd * (kpad / 2);
  OperandOut o{out_f32, out_hi, out_lo};
  fp_operand_kernel<<<grid_for(total), THREADS, 0, stream>>>(sparse, sparse_bstride, sparse_ld, C2, dense,
                                                             dense_bstride, dense_ld, C1, idx, w, Nd, kpad, total, o);
  RN_LAUNCH_CHECK("fp_operand_kernel");
  return REGNET_OK;
}

int split_rows_launch(const float* src, int64_t rows, int cols, int ld_src, int kpad, __nv_bfloat16* hi,
                      __nv_bfloat16* lo, float* f32, cudaStream_t stream) {
  const int64_t total = rows * kpad;
  if (total == 0) return REGNET_OK;
  split_rows_kernel<<<grid_for(total), THREADS, 0, stream>>>(src, rows, cols, ld_src, kpad, hi, lo, f32);
  RN_LAUNCH_CHECK("split_rows_kernel");
  return REGNET_OK;
}

}  // namespace regnet
