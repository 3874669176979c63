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
#include <algorithm>
#include <stdlib.h>

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

// Row-accumulating forms of the two scatter-add backward kernels: one CTA owns one (b, c) row of the gradient w.r.t. the
// source points (N floats, <= 48 KB), accumulates it in SHARED memory (RED.ADD on smem instead of one L2 atomic per
// contribution: FP level 2 issues 590 M of them) and writes it once; the index / weight rows are shared by the C rows
// of a cloud and stay in L2.  No memset needed.  Summation order still varies between runs, like the reference's.
constexpr int ROW_SMEM_MAX = 12288;   // floats per row

__global__ void __launch_bounds__(THREADS)
interp_backward_rows_kernel(const float* __restrict__ gout, const int64_t* __restrict__ index,
                            const float* __restrict__ weight, int C, int Ns, int Nd, float* __restrict__ gin,
                            int* __restrict__ oob) {
  extern __shared__ float acc[];
  const int64_t bc = blockIdx.x, b = bc / C;
  for (int j = threadIdx.x; j < Ns; j += THREADS) acc[j] = 0.f;
  __syncthreads();
  const float* __restrict__ g = gout + bc * Nd;
  const int64_t* __restrict__ idx = index + b * Nd * 3;
  const float* __restrict__ w = weight + b * Nd * 3;
  for (int n = threadIdx.x; n < Nd; n += THREADS) {
    const float gv = g[n];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t j = idx[(int64_t)n * 3 + k];
      if (j < 0 || j >= Ns) { *oob = 1; continue; }
      atomicAdd(acc + j, __fmul_rn(gv, w[(int64_t)n * 3 + k]));
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Ns; j += THREADS) gin[bc * Ns + j] = acc[j];
}

__global__ void __launch_bounds__(THREADS)
group_backward_rows_kernel(const float* __restrict__ gout, const int64_t* __restrict__ index, int C, int N, int MK,
                           float* __restrict__ gin, int* __restrict__ oob) {
  extern __shared__ float acc[];
  const int64_t bc = blockIdx.x, b = bc / C;
  for (int j = threadIdx.x; j < N; j += THREADS) acc[j] = 0.f;
  __syncthreads();
  const float* __restrict__ g = gout + bc * MK;
  const int64_t* __restrict__ idx = index + b * MK;
  for (int e = threadIdx.x; e < MK; e += THREADS) {
    const int64_t j = idx[e];
    if (j < 0 || j >= N) { *oob = 1; continue; }
    atomicAdd(acc + j, g[e]);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += THREADS) gin[bc * N + j] = acc[j];
}

// ---- fused operand producers ----------------------------------------------------------------------------------
// Output row r (one grouped position / one dense point) has `kpad` columns; FOUR adjacent columns per thread so that
// feature rows move as 16-byte vectors and enough bytes are in flight per thread to cover the gather latency.
struct OperandOut {
  float* f32;            // (rows, kpad) fp32, or nullptr
  __nv_bfloat16* hi;     // (rows, kpad) bf16 planes, or nullptr
  __nv_bfloat16* lo;
};

__device__ __forceinline__ void store_quad(const OperandOut& o, int64_t off, float4 v) {
  if (o.f32) *reinterpret_cast<float4*>(o.f32 + off) = v;
  if (o.hi) {
    __nv_bfloat16 h[4], l[4];
    split_bf16(v.x, h[0], l[0]);
    split_bf16(v.y, h[1], l[1]);
    split_bf16(v.z, h[2], l[2]);
    split_bf16(v.w, h[3], l[3]);
    __nv_bfloat162 h01 = __halves2bfloat162(h[0], h[1]), h23 = __halves2bfloat162(h[2], h[3]);
    __nv_bfloat162 l01 = __halves2bfloat162(l[0], l[1]), l23 = __halves2bfloat162(l[2], l[3]);
    uint2 ph, pl;
    ph.x = *reinterpret_cast<uint32_t*>(&h01); ph.y = *reinterpret_cast<uint32_t*>(&h23);
    pl.x = *reinterpret_cast<uint32_t*>(&l01); pl.y = *reinterpret_cast<uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(o.hi + off) = ph;
    *reinterpret_cast<uint2*>(o.lo + off) = pl;
  }
}

// Index arithmetic in these producers is 32-bit with power-of-two splits: a first version used 64-bit '/' and '%' by
// run-time divisors per element and spent more instructions on (software) division than on the data.
//   block = 256 threads = (256/TQ) rows x TQ quads;  grid.x over row groups (grid-stride), grid.y over quad tiles.

// SA operand: row (b,m,k) = [ feat[j, 0..C) | xyz[j]-new_xyz[m] (3) | 0 ... ], j = nbr[b,m,k], K = 64 neighbours.
// The reference concatenates [xyz_rel, feature] (modules.py:44-52); the rotation by three columns is folded into the
// first layer's weights when they are uploaded (regnet_scorenet_set_layer), which keeps feature quads 16-byte aligned.
template <int TQ>
__global__ void __launch_bounds__(THREADS)
sa_operand_kernel(const float* __restrict__ xyz, Strides3 xst, const float* __restrict__ new_xyz,
                  const float* __restrict__ feat, int64_t feat_bstride, int feat_ld, int C, int vec_ok,
                  const int32_t* __restrict__ nbr, uint32_t M, int kpad, uint32_t rows, OperandOut out) {
  constexpr uint32_t RPB = THREADS / TQ;
  const uint32_t q = blockIdx.y * TQ + (threadIdx.x % TQ);
  if ((int)(q * 4) >= kpad) return;
  const int c0 = (int)q * 4;
  for (uint32_t row = blockIdx.x * RPB + threadIdx.x / TQ; row < rows; row += gridDim.x * RPB) {
    const uint32_t bm = row >> 6;             // K == 64
    const uint32_t b = bm / M, m = bm - b * M;
    const int j = nbr[row];
    float4 v;
    if (vec_ok && c0 + 4 <= C) {
      v = *reinterpret_cast<const float4*>(feat + (int64_t)b * feat_bstride + (int64_t)j * feat_ld + c0);
    } else {
      float t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int col = c0 + u;
        float x = 0.f;
        if (col < C) {
          x = feat[(int64_t)b * feat_bstride + (int64_t)j * feat_ld + col];
        } else if (col < C + 3) {
          const int a = col - C;
          x = __fsub_rn(xyz[(int64_t)b * xst.b + a * xst.c + (int64_t)j * xst.n], new_xyz[((int64_t)b * 3 + a) * M + m]);
        }
        t[u] = x;
      }
      v = make_float4(t[0], t[1], t[2], t[3]);
    }
    store_quad(out, (int64_t)row * kpad + c0, v);
  }
}

// SA level 0 with its first MLP layer fused: the grouped input has only 6 channels (rgb + xyz_rel), so the layer
// 6 -> COUT is ~770 FMAs per position -- cheaper on the SIMT pipes than a tensor-core launch whose epilogue has to
// write the same 128-channel activation anyway.  One thread = (position, 4 output channels); its 24 weights and
// scale/shift stay in registers (the quad index is constant per thread because THREADS % (COUT/4) == 0).
// y = relu(scale * (W v) + shift), v = [feat[j,0..3) | xyz[j]-new_xyz[m]] (the rotated order of sa_operand_kernel).
template <int COUT>
__global__ void __launch_bounds__(THREADS, 3)
sa0_fused_kernel(const float* __restrict__ xyz, Strides3 xst, const float* __restrict__ new_xyz,
                 const float* __restrict__ feat, int64_t feat_bstride, int feat_ld, const int32_t* __restrict__ nbr,
                 const float* __restrict__ W, int ldw, const float* __restrict__ scale, const float* __restrict__ shift,
                 uint32_t M, uint32_t rows, OperandOut out, int ld_out) {
  constexpr int QUADS = COUT / 4;
  static_assert(THREADS % QUADS == 0, "quad index must be constant per thread");
  const int q = threadIdx.x % QUADS;
  float w[4][6], sc[4], sh[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
#pragma unroll
    for (int c = 0; c < 6; ++c) w[u][c] = W[(q * 4 + u) * ldw + c];
    sc[u] = scale[q * 4 + u];
    sh[u] = shift[q * 4 + u];
  }
  // R rows per thread per trip, all their gathers issued before any arithmetic (latency of index -> point loads)
  constexpr int R = 4;
  constexpr uint32_t RPB = THREADS / QUADS;
  const uint32_t sub = threadIdx.x / QUADS;
  for (uint32_t base = blockIdx.x * RPB * R; base < rows; base += gridDim.x * RPB * R) {
    int j[R];
    uint32_t row[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      row[r] = base + r * RPB + sub;
      j[r] = row[r] < rows ? nbr[row[r]] : 0;
    }
    float v[R][6];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint32_t bm = (row[r] < rows ? row[r] : 0u) >> 6;   // K == 64
      const uint32_t b = bm / M, m = bm - b * M;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[r][c] = feat[(int64_t)b * feat_bstride + (int64_t)j[r] * feat_ld + c];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        v[r][3 + a] = __fsub_rn(xyz[(int64_t)b * xst.b + a * xst.c + (int64_t)j[r] * xst.n],
                                new_xyz[((int64_t)b * 3 + a) * M + m]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row[r] >= rows) continue;
      float y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc = fmaf(w[u][c], v[r][c], acc);
        y[u] = fmaxf(fmaf(acc, sc[u], sh[u]), 0.f);
      }
      store_quad(out, (int64_t)row[r] * ld_out + q * 4, make_float4(y[0], y[1], y[2], y[3]));
    }
  }
}

// FP operand: row (b,n) = [ sum_k w[b,n,k]*sparse[idx[b,n,k], 0..C2) | dense[n, 0..C1) | 0 ... ]   (modules.py:117-127)
__global__ void __launch_bounds__(THREADS)
fp_operand_kernel(const float* __restrict__ sparse, int64_t sparse_bstride, int sparse_ld, int C2,
                  const float* __restrict__ dense, int64_t dense_bstride, int dense_ld, int C1, int sparse_vec,
                  int dense_vec, const int32_t* __restrict__ idx, const float* __restrict__ w, uint32_t Nd, int kpad,
                  uint32_t rows, OperandOut out) {
  constexpr uint32_t TQ = 32, RPB = THREADS / TQ;
  const uint32_t q = blockIdx.y * TQ + (threadIdx.x % TQ);
  if ((int)(q * 4) >= kpad) return;
  const int c0 = (int)q * 4;
  for (uint32_t row = blockIdx.x * RPB + threadIdx.x / TQ; row < rows; row += gridDim.x * RPB) {
    const uint32_t b = row / Nd, n = row - b * Nd;
    float4 v;
    if (sparse_vec && c0 + 4 <= C2) {
      const int i0 = idx[row * 3], i1 = idx[row * 3 + 1], i2 = idx[row * 3 + 2];
      const float w0 = w[row * 3], w1 = w[row * 3 + 1], w2 = w[row * 3 + 2];
      const float* __restrict__ sp = sparse + (int64_t)b * sparse_bstride + c0;
      const float4 a0 = *reinterpret_cast<const float4*>(sp + (int64_t)i0 * sparse_ld);
      const float4 a1 = *reinterpret_cast<const float4*>(sp + (int64_t)i1 * sparse_ld);
      const float4 a2 = *reinterpret_cast<const float4*>(sp + (int64_t)i2 * sparse_ld);
      v.x = __fmaf_rn(a2.x, w2, __fmaf_rn(a1.x, w1, __fmaf_rn(a0.x, w0, 0.f)));   // k = 0,1,2 in order, as the
      v.y = __fmaf_rn(a2.y, w2, __fmaf_rn(a1.y, w1, __fmaf_rn(a0.y, w0, 0.f)));   // reference's fma chain
      v.z = __fmaf_rn(a2.z, w2, __fmaf_rn(a1.z, w1, __fmaf_rn(a0.z, w0, 0.f)));
      v.w = __fmaf_rn(a2.w, w2, __fmaf_rn(a1.w, w1, __fmaf_rn(a0.w, w0, 0.f)));
    } else if (dense_vec && c0 >= C2 && c0 + 4 <= C2 + C1) {
      v = *reinterpret_cast<const float4*>(dense + (int64_t)b * dense_bstride + (int64_t)n * dense_ld + (c0 - C2));
    } else {
      float t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int col = c0 + u;
        float x = 0.f;
        if (col < C2) {
          const float* __restrict__ sp = sparse + (int64_t)b * sparse_bstride + col;
          x = __fmaf_rn(sp[(int64_t)idx[row * 3] * sparse_ld], w[row * 3], 0.f);
          x = __fmaf_rn(sp[(int64_t)idx[row * 3 + 1] * sparse_ld], w[row * 3 + 1], x);
          x = __fmaf_rn(sp[(int64_t)idx[row * 3 + 2] * sparse_ld], w[row * 3 + 2], x);
        } else if (col < C2 + C1) {
          x = dense[(int64_t)b * dense_bstride + (int64_t)n * dense_ld + (col - C2)];
        }
        t[u] = x;
      }
      v = make_float4(t[0], t[1], t[2], t[3]);
    }
    store_quad(out, (int64_t)row * kpad + c0, v);
  }
}


// SA level >= 1 with the FEATURE part of its first 1x1 convolution applied per point, before the grouping
// (W [f_j | x_j - c_m] = W_f f_j + W_x (x_j - c_m): the feature term does not depend on the centroid):
//   Z = f W_f^T for the N_prev points of the previous level (tensor cores, 12.8x fewer rows than M*64 grouped positions),
//   y[b,m,k] = relu(scale * (Z[j] + W_x (x_j - c_m)) + shift),  j = nbr[b,m,k]
// The xyz term stays in fp32 on the difference, exactly like the reference's (x_j - c_m) operand -- folding W_x x_j into
// Z and subtracting W_x c_m would lose ~|x| / |x - c| of relative accuracy to cancellation.  One thread = (row, 4 ch).
__global__ void __launch_bounds__(THREADS)
sa_gather_affine_kernel(const float* __restrict__ Z, int ldz, uint32_t n_prev, const float* __restrict__ xyz, Strides3 xst,
                        const float* __restrict__ new_xyz, const float* __restrict__ Wx, int ldw,
                        const int32_t* __restrict__ nbr, const float* __restrict__ scale, const float* __restrict__ shift,
                        uint32_t M, int cout, uint32_t rows, OperandOut out) {
  constexpr uint32_t TQ = 32, RPB = THREADS / TQ;
  const uint32_t q = blockIdx.y * TQ + (threadIdx.x % TQ);
  if ((int)(q * 4) >= cout) return;
  const int c0 = (int)q * 4;
  const float4 sc = *reinterpret_cast<const float4*>(scale + c0), sh = *reinterpret_cast<const float4*>(shift + c0);
  float wx[4][3];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int a = 0; a < 3; ++a) wx[u][a] = Wx[(c0 + u) * ldw + a];
  for (uint32_t row = blockIdx.x * RPB + threadIdx.x / TQ; row < rows; row += gridDim.x * RPB) {
    const uint32_t bm = row >> 6;             // 64 neighbours per centroid
    const uint32_t b = bm / M, m = bm - b * M;
    const int j = nbr[row];
    const float4 z = *reinterpret_cast<const float4*>(Z + ((int64_t)b * n_prev + j) * ldz + c0);
    float rel[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      rel[a] = __fsub_rn(xyz[(int64_t)b * xst.b + a * xst.c + (int64_t)j * xst.n], new_xyz[((int64_t)b * 3 + a) * M + m]);
    float4 v;
    v.x = fmaf(wx[0][2], rel[2], fmaf(wx[0][1], rel[1], fmaf(wx[0][0], rel[0], z.x)));
    v.y = fmaf(wx[1][2], rel[2], fmaf(wx[1][1], rel[1], fmaf(wx[1][0], rel[0], z.y)));
    v.z = fmaf(wx[2][2], rel[2], fmaf(wx[2][1], rel[1], fmaf(wx[2][0], rel[0], z.z)));
    v.w = fmaf(wx[3][2], rel[2], fmaf(wx[3][1], rel[1], fmaf(wx[3][0], rel[0], z.w)));
    v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
    v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
    v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
    v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
    store_quad(out, (int64_t)row * cout + c0, v);
  }
}

// Warp-cooperative form of the kernel above (used when cout % 8 == 0): one warp = 32 consecutive rows = half a
// centroid's neighbour list, so the centroid is loaded once per warp and every lane fetches ONE neighbour index and its
// coordinates (coalesced) instead of every thread re-loading all seven scalars of its row; the rows are then walked with
// shuffles, each lane producing 8 consecutive channels: 32-byte loads of Z (1 KB per warp request), 16-byte stores to each
// bf16 plane, four rows in flight.  Same fma chain per element => identical results.
__device__ __forceinline__ void store_oct(const OperandOut& o, int64_t off, float4 a, float4 b) {
  if (o.f32) {
    *reinterpret_cast<float4*>(o.f32 + off) = a;
    *reinterpret_cast<float4*>(o.f32 + off + 4) = b;
  }
  if (o.hi) {
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      split_bf16_pair(v[2 * u], v[2 * u + 1], ph[u], pl[u]);
    }
    *reinterpret_cast<uint4*>(o.hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(o.lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

__global__ void __launch_bounds__(THREADS)
sa_gather_affine_warp_kernel(const float* __restrict__ Z, int ldz, uint32_t n_prev, const float* __restrict__ xyz,
                             Strides3 xst, const float* __restrict__ new_xyz, const float* __restrict__ Wx, int ldw,
                             const int32_t* __restrict__ nbr, const float* __restrict__ scale,
                             const float* __restrict__ shift, uint32_t M, int cout, uint32_t rows, OperandOut out) {
  const uint32_t lane = threadIdx.x & 31;
  // lanes beyond the last channel keep running (the row walk uses full-mask shuffles); they recompute the last octet
  // and skip the store
  const bool live = (int)(blockIdx.y * 256 + lane * 8) < cout;
  const int c0 = live ? (int)(blockIdx.y * 256 + lane * 8) : cout - 8;
  float sc[8], sh[8], wx[8][3];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    sc[u] = scale[c0 + u];
    sh[u] = shift[c0 + u];
#pragma unroll
    for (int a = 0; a < 3; ++a) wx[u][a] = Wx[(c0 + u) * ldw + a];
  }
  const uint32_t nwarps = gridDim.x * (THREADS / 32);
  for (uint32_t base = (blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * 32; base < rows; base += nwarps * 32) {
    const uint32_t bm = base >> 6;             // 64 neighbours per centroid: the 32 rows of a warp share (b, m)
    const uint32_t b = bm / M, m = bm - b * M;
    const int j = nbr[base + lane];
    float rel[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      rel[a] = __fsub_rn(xyz[(int64_t)b * xst.b + a * xst.c + (int64_t)j * xst.n], new_xyz[((int64_t)b * 3 + a) * M + m]);
    const float* __restrict__ zb = Z + (int64_t)b * n_prev * ldz + c0;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const int jr = __shfl_sync(0xffffffffu, j, r);
      const float r0 = __shfl_sync(0xffffffffu, rel[0], r), r1 = __shfl_sync(0xffffffffu, rel[1], r),
                  r2 = __shfl_sync(0xffffffffu, rel[2], r);
      const float4 za = *reinterpret_cast<const float4*>(zb + (int64_t)jr * ldz);
      const float4 zc = *reinterpret_cast<const float4*>(zb + (int64_t)jr * ldz + 4);
      float v[8] = {za.x, za.y, za.z, za.w, zc.x, zc.y, zc.z, zc.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = fmaf(wx[u][2], r2, fmaf(wx[u][1], r1, fmaf(wx[u][0], r0, v[u])));
        v[u] = fmaxf(fmaf(v[u], sc[u], sh[u]), 0.f);
      }
      if (live)
        store_oct(out, (int64_t)(base + r) * cout + c0, make_float4(v[0], v[1], v[2], v[3]),
                  make_float4(v[4], v[5], v[6], v[7]));
    }
  }
}

// FP module with the first 1x1 convolution applied BEFORE the interpolation (both are linear, so they commute):
//   reference (modules.py:117-127 + conv.py:24-36):  y = act(scale * W [ sum_k w_k f[idx_k] | dense ] + shift)
//   here:   Y = f W_s^T at the Ns sparse points (tensor cores, Ns << Nd rows),  D = dense W_d^T (or a 3-channel matvec),
//           y[n] = act(scale * (sum_k w_k Y[idx_k] + D[n]) + shift)
// The GEMM shrinks from Nd x (C2 + C1) x cout to Ns x C2 x cout (+ Nd x C1 x cout), the (Nd, C2 + C1) operand is never
// built, and this kernel reads 3 rows of cout floats per point from an L2-resident table.  One thread = (row, 4 channels).
__global__ void __launch_bounds__(THREADS)
fp_interp_affine_kernel(const float* __restrict__ Y, int64_t y_bstride, int ldy, const float* __restrict__ D, int ldd,
                        const float* __restrict__ dense3, int64_t dense3_bstride, int dense3_ld,
                        const float* __restrict__ Wd3, int ldw3, const int32_t* __restrict__ idx,
                        const float* __restrict__ w, const float* __restrict__ scale, const float* __restrict__ shift,
                        uint32_t Nd, int cout, uint32_t rows, OperandOut out) {
  constexpr uint32_t TQ = 32, RPB = THREADS / TQ;
  const uint32_t q = blockIdx.y * TQ + (threadIdx.x % TQ);
  if ((int)(q * 4) >= cout) return;
  const int c0 = (int)q * 4;
  const float4 sc = *reinterpret_cast<const float4*>(scale + c0), sh = *reinterpret_cast<const float4*>(shift + c0);
  float wd[4][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  if (Wd3) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < 3; ++c) wd[u][c] = Wd3[(c0 + u) * ldw3 + c];
  }
  for (uint32_t row = blockIdx.x * RPB + threadIdx.x / TQ; row < rows; row += gridDim.x * RPB) {
    const uint32_t b = row / Nd, n = row - b * Nd;
    const int i0 = idx[row * 3], i1 = idx[row * 3 + 1], i2 = idx[row * 3 + 2];
    const float w0 = w[row * 3], w1 = w[row * 3 + 1], w2 = w[row * 3 + 2];
    const float* __restrict__ yp = Y + (int64_t)b * y_bstride + c0;
    const float4 a0 = *reinterpret_cast<const float4*>(yp + (int64_t)i0 * ldy);
    const float4 a1 = *reinterpret_cast<const float4*>(yp + (int64_t)i1 * ldy);
    const float4 a2 = *reinterpret_cast<const float4*>(yp + (int64_t)i2 * ldy);
    float4 v;
    v.x = __fmaf_rn(a2.x, w2, __fmaf_rn(a1.x, w1, __fmul_rn(a0.x, w0)));
    v.y = __fmaf_rn(a2.y, w2, __fmaf_rn(a1.y, w1, __fmul_rn(a0.y, w0)));
    v.z = __fmaf_rn(a2.z, w2, __fmaf_rn(a1.z, w1, __fmul_rn(a0.z, w0)));
    v.w = __fmaf_rn(a2.w, w2, __fmaf_rn(a1.w, w1, __fmul_rn(a0.w, w0)));
    if (D) {
      const float4 d = *reinterpret_cast<const float4*>(D + (int64_t)row * ldd + c0);
      v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
    }
    if (Wd3) {
      const float* __restrict__ dp = dense3 + (int64_t)b * dense3_bstride + (int64_t)n * dense3_ld;
      const float r0 = dp[0], r1 = dp[1], r2 = dp[2];
      v.x = fmaf(wd[0][2], r2, fmaf(wd[0][1], r1, fmaf(wd[0][0], r0, v.x)));
      v.y = fmaf(wd[1][2], r2, fmaf(wd[1][1], r1, fmaf(wd[1][0], r0, v.y)));
      v.z = fmaf(wd[2][2], r2, fmaf(wd[2][1], r1, fmaf(wd[2][0], r0, v.z)));
      v.w = fmaf(wd[3][2], r2, fmaf(wd[3][1], r1, fmaf(wd[3][0], r0, v.w)));
    }
    v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
    v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
    v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
    v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
    store_quad(out, (int64_t)row * cout + c0, v);
  }
}

// Warp-cooperative form of fp_interp_affine_kernel (cout % 8 == 0): one warp = 32 consecutive points; lane l loads the
// three neighbour indices / weights (and the 3 dense channels) of point base + l, the points are walked with shuffles,
// each lane producing 8 consecutive channels (32-byte loads, 16-byte plane stores).  Same arithmetic per element.
__global__ void __launch_bounds__(THREADS)
fp_interp_affine_warp_kernel(const float* __restrict__ Y, int64_t y_bstride, int ldy, const float* __restrict__ D, int ldd,
                             const float* __restrict__ dense3, int64_t dense3_bstride, int dense3_ld,
                             const float* __restrict__ Wd3, int ldw3, const int32_t* __restrict__ idx,
                             const float* __restrict__ w, const float* __restrict__ scale,
                             const float* __restrict__ shift, uint32_t Nd, int cout, uint32_t rows, OperandOut out) {
  const uint32_t lane = threadIdx.x & 31;
  // lanes beyond the last channel keep running (the row walk uses full-mask shuffles); they recompute the last octet
  // and skip the store
  const bool live = (int)(blockIdx.y * 256 + lane * 8) < cout;
  const int c0 = live ? (int)(blockIdx.y * 256 + lane * 8) : cout - 8;
  float sc[8], sh[8], wd[8][3];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    sc[u] = scale[c0 + u];
    sh[u] = shift[c0 + u];
#pragma unroll
    for (int c = 0; c < 3; ++c) wd[u][c] = Wd3 ? Wd3[(c0 + u) * ldw3 + c] : 0.f;
  }
  const uint32_t nwarps = gridDim.x * (THREADS / 32);
  for (uint32_t base = (blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * 32; base < rows; base += nwarps * 32) {
    const uint32_t row = min(base + lane, rows - 1);
    const uint32_t b = row / Nd, n = row - b * Nd;
    const int i0 = idx[(int64_t)row * 3], i1 = idx[(int64_t)row * 3 + 1], i2 = idx[(int64_t)row * 3 + 2];
    const float w0 = w[(int64_t)row * 3], w1 = w[(int64_t)row * 3 + 1], w2 = w[(int64_t)row * 3 + 2];
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    if (Wd3) {
      const float* __restrict__ dp = dense3 + (int64_t)b * dense3_bstride + (int64_t)n * dense3_ld;
      d0 = dp[0]; d1 = dp[1]; d2 = dp[2];
    }
    const int cnt = (int)min(32u, rows - base);
#pragma unroll 2
    for (int r = 0; r < cnt; ++r) {
      const uint32_t br = __shfl_sync(0xffffffffu, b, r);
      const int j0 = __shfl_sync(0xffffffffu, i0, r), j1 = __shfl_sync(0xffffffffu, i1, r), j2 = __shfl_sync(0xffffffffu, i2, r);
      const float u0 = __shfl_sync(0xffffffffu, w0, r), u1 = __shfl_sync(0xffffffffu, w1, r), u2 = __shfl_sync(0xffffffffu, w2, r);
      const float* __restrict__ yp = Y + (int64_t)br * y_bstride + c0;
      const float4 a0 = *reinterpret_cast<const float4*>(yp + (int64_t)j0 * ldy), a0b = *reinterpret_cast<const float4*>(yp + (int64_t)j0 * ldy + 4);
      const float4 a1 = *reinterpret_cast<const float4*>(yp + (int64_t)j1 * ldy), a1b = *reinterpret_cast<const float4*>(yp + (int64_t)j1 * ldy + 4);
      const float4 a2 = *reinterpret_cast<const float4*>(yp + (int64_t)j2 * ldy), a2b = *reinterpret_cast<const float4*>(yp + (int64_t)j2 * ldy + 4);
      const float x0[8] = {a0.x, a0.y, a0.z, a0.w, a0b.x, a0b.y, a0b.z, a0b.w};
      const float x1[8] = {a1.x, a1.y, a1.z, a1.w, a1b.x, a1b.y, a1b.z, a1b.w};
      const float x2[8] = {a2.x, a2.y, a2.z, a2.w, a2b.x, a2b.y, a2b.z, a2b.w};
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __fmaf_rn(x2[u], u2, __fmaf_rn(x1[u], u1, __fmul_rn(x0[u], u0)));
      if (D) {
        const float4 da = *reinterpret_cast<const float4*>(D + (int64_t)(base + r) * ldd + c0);
        const float4 db = *reinterpret_cast<const float4*>(D + (int64_t)(base + r) * ldd + c0 + 4);
        v[0] += da.x; v[1] += da.y; v[2] += da.z; v[3] += da.w; v[4] += db.x; v[5] += db.y; v[6] += db.z; v[7] += db.w;
      }
      if (Wd3) {
        const float e0 = __shfl_sync(0xffffffffu, d0, r), e1 = __shfl_sync(0xffffffffu, d1, r), e2 = __shfl_sync(0xffffffffu, d2, r);
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = fmaf(wd[u][2], e2, fmaf(wd[u][1], e1, fmaf(wd[u][0], e0, v[u])));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = fmaxf(fmaf(v[u], sc[u], sh[u]), 0.f);
      if (live)
        store_oct(out, (int64_t)(base + r) * cout + c0, make_float4(v[0], v[1], v[2], v[3]),
                  make_float4(v[4], v[5], v[6], v[7]));
    }
  }
}

// fp32 rows -> bf16 hi/lo planes (used for weights and by regnet_mlp_layer)
__global__ void __launch_bounds__(THREADS)
split_rows_kernel(const float* __restrict__ src, int64_t rows, int cols, int ld_src, int kpad, int rot,
                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ f32,
                  const float* __restrict__ row_sign) {
  const int64_t total = rows * kpad;
  for (int64_t e = blockIdx.x * (int64_t)THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * THREADS) {
    const int col = (int)(e % kpad);
    const int64_t r = e / kpad;
    float x = col < cols ? src[r * ld_src + (col + rot) % cols] : 0.f;   // dst col c <- src col (c+rot) mod cols
    if (row_sign && row_sign[r] < 0.f) x = -x;   // rows whose BN scale is negative are negated (the scale's sign too)
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
  if (N <= ROW_SMEM_MAX && (int64_t)M * K >= 4 * (int64_t)N && (int64_t)M * K < (1LL << 31) && (int64_t)B * C > 0 &&
      !getenv("REGNET_SCATTER_GLOBAL")) {
    group_backward_rows_kernel<<<(unsigned)((int64_t)B * C), THREADS, sizeof(float) * (size_t)N, stream>>>(
        gout, index, C, N, M * K, gin, d_oob);
    RN_LAUNCH_CHECK("group_backward_rows_kernel");
    return REGNET_OK;
  }
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
  if (Ns <= ROW_SMEM_MAX && Nd >= Ns && (int64_t)B * C > 0 && !getenv("REGNET_SCATTER_GLOBAL")) {
    interp_backward_rows_kernel<<<(unsigned)((int64_t)B * C), THREADS, sizeof(float) * (size_t)Ns, stream>>>(
        gout, index, weight, C, Ns, Nd, gin, d_oob);
    RN_LAUNCH_CHECK("interp_backward_rows_kernel");
    return REGNET_OK;
  }
  RN_CUDA(cudaMemsetAsync(gin, 0, sizeof(float) * (size_t)B * C * Ns, stream));
  const int64_t total = (int64_t)B * C * Nd;
  if (total == 0) return REGNET_OK;
  interp_backward_kernel<<<grid_for(total), THREADS, 0, stream>>>(gout, index, weight, C, Ns, Nd, total, gin, d_oob);
  RN_LAUNCH_CHECK("interp_backward_kernel");
  return REGNET_OK;
}

static inline unsigned row_grid(uint32_t rows, uint32_t rows_per_block, unsigned ytiles) {
  uint64_t g = ((uint64_t)rows + rows_per_block - 1) / rows_per_block;
  const uint64_t cap = (148ull * 16 + ytiles - 1) / ytiles;   // ~16 resident blocks per SM overall, then grid-stride
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

int sa_operand_launch(const float* xyz, Strides3 xst, const float* new_xyz, const float* feat, int64_t feat_bstride,
                      int feat_ld, int C, const int32_t* nbr, int B, int N, int M, int K, int kpad, float* out_f32,
                      __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream) {
  RN_CHECK_ARG(kpad % 4 == 0 && kpad >= 3 + C, "sa_operand: bad kpad %d for C=%d", kpad, C);
  RN_CHECK_ARG(K == 64, "sa_operand: 64 neighbours per centroid expected");
  const int64_t rows64 = (int64_t)B * M * K;
  RN_CHECK_ARG(rows64 < (1LL << 31), "sa_operand: too many positions");
  const uint32_t rows = (uint32_t)rows64;
  OperandOut o{out_f32, out_hi, out_lo};
  const int vec_ok = (feat_ld % 4 == 0) && (feat_bstride % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0);
  const int quads = kpad / 4;
  if (quads <= 4) {
    dim3 grid(row_grid(rows, THREADS / 4, 1), 1);
    RN_PREFER_MAX_SMEM(sa_operand_kernel<4>);
    sa_operand_kernel<4><<<grid, THREADS, 0, stream>>>(xyz, xst, new_xyz, feat, feat_bstride, feat_ld, C, vec_ok, nbr,
                                                       (uint32_t)M, kpad, rows, o);
  } else {
    const unsigned yt = (unsigned)ceil_div(quads, 32);
    dim3 grid(row_grid(rows, THREADS / 32, yt), yt);
    RN_PREFER_MAX_SMEM(sa_operand_kernel<32>);
    sa_operand_kernel<32><<<grid, THREADS, 0, stream>>>(xyz, xst, new_xyz, feat, feat_bstride, feat_ld, C, vec_ok, nbr,
                                                        (uint32_t)M, kpad, rows, o);
  }
  RN_LAUNCH_CHECK("sa_operand_kernel");
  (void)N;
  return REGNET_OK;
}

int sa0_fused_launch(const float* xyz, Strides3 xst, const float* new_xyz, const float* feat, int64_t feat_bstride,
                     int feat_ld, const int32_t* nbr, const float* W, int ldw, const float* scale, const float* shift,
                     int cout, int B, int M, int K, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                     int ld_out, cudaStream_t stream) {
  RN_CHECK_ARG(cout == 128 && ld_out % 4 == 0 && K == 64, "sa0_fused: only the 6 -> 128 layer of the reference architecture");
  const int64_t rows64 = (int64_t)B * M * K;
  RN_CHECK_ARG(rows64 < (1LL << 31), "sa0_fused: too many positions");
  OperandOut o{out_f32, out_hi, out_lo};
  const unsigned grid = row_grid((uint32_t)rows64, (THREADS / 32) * 4, 1);
  RN_PREFER_MAX_SMEM(sa0_fused_kernel<128>);
  sa0_fused_kernel<128><<<grid, THREADS, 0, stream>>>(xyz, xst, new_xyz, feat, feat_bstride, feat_ld, nbr, W, ldw, scale,
                                                     shift, (uint32_t)M, (uint32_t)rows64, o, ld_out);
  RN_LAUNCH_CHECK("sa0_fused_kernel");
  return REGNET_OK;
}

int fp_operand_launch(const float* sparse, int64_t sparse_bstride, int sparse_ld, int C2, const float* dense,
                      int64_t dense_bstride, int dense_ld, int C1, const int32_t* idx, const float* w, int B, int Nd,
                      int kpad, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream) {
  RN_CHECK_ARG(kpad % 4 == 0 && kpad >= C1 + C2, "fp_operand: bad kpad %d for C1+C2=%d", kpad, C1 + C2);
  const int64_t rows64 = (int64_t)B * Nd;
  RN_CHECK_ARG(rows64 * 3 < (1LL << 31), "fp_operand: too many points");
  OperandOut o{out_f32, out_hi, out_lo};
  const int sparse_vec = (sparse_ld % 4 == 0) && (sparse_bstride % 4 == 0) && ((reinterpret_cast<uintptr_t>(sparse) & 15) == 0);
  const int dense_vec = (dense_ld % 4 == 0) && (dense_bstride % 4 == 0) && (C2 % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(dense) & 15) == 0);
  const unsigned yt = (unsigned)ceil_div(kpad / 4, 32);
  dim3 grid(row_grid((uint32_t)rows64, THREADS / 32, yt), yt);
  RN_PREFER_MAX_SMEM(fp_operand_kernel);
  fp_operand_kernel<<<grid, THREADS, 0, stream>>>(sparse, sparse_bstride, sparse_ld, C2, dense, dense_bstride, dense_ld,
                                                  C1, sparse_vec, dense_vec, idx, w, (uint32_t)Nd, kpad, (uint32_t)rows64, o);
  RN_LAUNCH_CHECK("fp_operand_kernel");
  return REGNET_OK;
}


// y = relu(scale * (Z[nbr] + Wx (xyz[nbr] - centroid)) + shift) as bf16 hi/lo planes (B*M*64, cout); see the kernel
int sa_gather_affine_launch(const float* Z, int ldz, int n_prev, const float* xyz, Strides3 xst, const float* new_xyz,
                            const float* Wx, int ldw, const int32_t* nbr, const float* scale, const float* shift, int B,
                            int M, int cout, float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                            cudaStream_t stream) {
  const int64_t rows64 = (int64_t)B * M * 64;
  RN_CHECK_ARG(cout % 4 == 0 && ldz % 4 == 0, "sa_gather_affine: unaligned rows");
  RN_CHECK_ARG(rows64 < (1LL << 31), "sa_gather_affine: too many positions");
  if (rows64 == 0) return REGNET_OK;
  const uint32_t rows = (uint32_t)rows64;
  OperandOut o{out_f32, out_hi, out_lo};
  if (cout % 8 == 0 && !getenv("REGNET_AFFINE_V1")) {
    const unsigned yw = (unsigned)ceil_div(cout, 256);
    const unsigned gx = (unsigned)std::min<int64_t>(ceil_div((int)(rows / 32), THREADS / 32), std::max(1, 148 * 8 / (int)yw));
    RN_PREFER_MAX_SMEM(sa_gather_affine_warp_kernel);
    sa_gather_affine_warp_kernel<<<dim3(gx, yw), THREADS, 0, stream>>>(Z, ldz, (uint32_t)n_prev, xyz, xst, new_xyz, Wx, ldw,
                                                                      nbr, scale, shift, (uint32_t)M, cout, rows, o);
    RN_LAUNCH_CHECK("sa_gather_affine_warp_kernel");
    return REGNET_OK;
  }
  const unsigned yt = (unsigned)ceil_div(cout / 4, 32);
  dim3 grid(row_grid(rows, THREADS / 32, yt), yt);
  RN_PREFER_MAX_SMEM(sa_gather_affine_kernel);
  sa_gather_affine_kernel<<<grid, THREADS, 0, stream>>>(Z, ldz, (uint32_t)n_prev, xyz, xst, new_xyz, Wx, ldw, nbr, scale,
                                                        shift, (uint32_t)M, cout, rows, o);
  RN_LAUNCH_CHECK("sa_gather_affine_kernel");
  return REGNET_OK;
}

// y = relu(scale * (3-NN interpolation of Y + D [+ Wd3 . dense3]) + shift) as bf16 hi/lo planes (rows, cout); see the kernel
int fp_interp_affine_launch(const float* Y, int64_t y_bstride, int ldy, const float* D, int ldd, const float* dense3,
                            int64_t dense3_bstride, int dense3_ld, const float* Wd3, int ldw3, const int32_t* idx,
                            const float* w, const float* scale, const float* shift, int B, int Nd, int cout,
                            float* out_f32, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t stream) {
  const int64_t rows64 = (int64_t)B * Nd;
  RN_CHECK_ARG(cout % 4 == 0 && ldy % 4 == 0 && (!D || ldd % 4 == 0) && y_bstride % 4 == 0, "fp_interp_affine: unaligned rows");
  RN_CHECK_ARG(rows64 < (1LL << 31), "fp_interp_affine: too many points");
  if (rows64 == 0) return REGNET_OK;
  const uint32_t rows = (uint32_t)rows64;
  OperandOut o{out_f32, out_hi, out_lo};
  // the warp-per-32-points form needs enough warps to fill the GPU (FP level 2: 12 000); the small levels keep one
  // thread per (point, 4 channels)
  if (cout % 8 == 0 && (int64_t)(rows / 32) * ceil_div(cout, 256) >= 8192 && !getenv("REGNET_AFFINE_V1")) {
    const unsigned yw = (unsigned)ceil_div(cout, 256);
    const unsigned gx = (unsigned)std::min<int64_t>(ceil_div((int)ceil_div((int)rows, 32), THREADS / 32), std::max(1, 148 * 8 / (int)yw));
    RN_PREFER_MAX_SMEM(fp_interp_affine_warp_kernel);
    fp_interp_affine_warp_kernel<<<dim3(gx, yw), THREADS, 0, stream>>>(Y, y_bstride, ldy, D, ldd, dense3, dense3_bstride,
                                                                      dense3_ld, Wd3, ldw3, idx, w, scale, shift,
                                                                      (uint32_t)Nd, cout, rows, o);
    RN_LAUNCH_CHECK("fp_interp_affine_warp_kernel");
    return REGNET_OK;
  }
  const unsigned yt = (unsigned)ceil_div(cout / 4, 32);
  dim3 grid(row_grid(rows, THREADS / 32, yt), yt);
  RN_PREFER_MAX_SMEM(fp_interp_affine_kernel);
  fp_interp_affine_kernel<<<grid, THREADS, 0, stream>>>(Y, y_bstride, ldy, D, ldd, dense3, dense3_bstride, dense3_ld, Wd3,
                                                        ldw3, idx, w, scale, shift, (uint32_t)Nd, cout, rows, o);
  RN_LAUNCH_CHECK("fp_interp_affine_kernel");
  return REGNET_OK;
}

int split_rows_launch(const float* src, int64_t rows, int cols, int ld_src, int kpad, __nv_bfloat16* hi,
                      __nv_bfloat16* lo, float* f32, cudaStream_t stream, int rot, const float* row_sign) {
  const int64_t total = rows * kpad;
  if (total == 0) return REGNET_OK;
  split_rows_kernel<<<grid_for(total), THREADS, 0, stream>>>(src, rows, cols, ld_src, kpad, rot, hi, lo, f32, row_sign);
  RN_LAUNCH_CHECK("split_rows_kernel");
  return REGNET_OK;
}

namespace {
__global__ void abs_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = fabsf(src[i]);
}
}  // namespace

int abs_copy_launch(const float* src, float* dst, int n, cudaStream_t stream) {
  if (n <= 0) return REGNET_OK;
  abs_copy_kernel<<<(n + 255) / 256, 256, 0, stream>>>(src, dst, n);
  RN_LAUNCH_CHECK("abs_copy_kernel");
  return REGNET_OK;
}

}  // namespace regnet
