// Training-mode pieces of the shared MLP (SURVEY.md 8a row A10 in train mode, row A11's max over neighbours):
//
//   bn_relu_train forward / backward   replaces nn.BatchNorm{1,2}d (batch statistics) + nn.ReLU(inplace) of
//                                      nn/modules/conv.py:24-36,64-76 -- in torch: cudnn bn_fw_tr + a ReLU pass forward,
//                                      threshold_backward + cudnn bn_bw backward (58 ms of a 146 ms training step);
//   maxpool_k forward / backward       replaces torch.max(x, 3)[0] of modules.py:245 and its scatter backward
//                                      (at::reduce_kernel over the innermost 64 elements: 15.6 ms of that step).
//
// All four are HBM-bound streaming kernels over (B, C, L) fp32 tensors in torch's native NCHW layout (L = M*K for the
// 2-D blocks, N for the 1-D blocks): forward = 2 reads + 1 write of the activation, backward = 4 reads + 1 write.
// Statistics are accumulated as (count, mean, M2) triples merged with Chan's formula (fp32 per thread over <= a few
// hundred elements, fp64 for the final merge over the partials), so the variance does not suffer the sum-of-squares
// cancellation.  The ReLU mask is never stored: backward recomputes fma(x, scale, shift) > 0 from the very (scale, shift)
// the forward used.
#include <algorithm>

#include "internal.cuh"
#include "train_common.cuh"

namespace regnet {

namespace {

constexpr int TB = 256;
constexpr unsigned FULL = 0xffffffffu;

struct ChunkPlan {
  int sl;          // splits of the L axis
  int64_t chunk;   // elements per split (multiple of 4)
};

__host__ __device__ inline ChunkPlan plan_chunks(int64_t L) {
  ChunkPlan p;
  int64_t sl = (L + 16383) / 16384;
  if (sl < 1) sl = 1;
  if (sl > 64) sl = 64;
  int64_t chunk = (L + sl - 1) / sl;
  chunk = (chunk + 3) / 4 * 4;
  p.sl = (int)((L + chunk - 1) / chunk);
  p.chunk = chunk;
  return p;
}

struct Welford {
  float n, mean, m2;
};

__device__ __forceinline__ Welford merge(Welford a, Welford b) {
  const float n = a.n + b.n;
  if (n == 0.f) return a;
  const float d = b.mean - a.mean;
  const float f = b.n / n;
  Welford r;
  r.n = n;
  r.mean = fmaf(d, f, a.mean);
  r.m2 = a.m2 + b.m2 + d * d * a.n * f;
  return r;
}

// ---- forward, pass 1: per (channel, batch row, L split) partial statistics -------------------------------------------
__global__ void __launch_bounds__(TB)
bn_stats_kernel(const float* __restrict__ x, int C, int64_t L, int64_t chunk, int sl, float* __restrict__ partial) {
  const int c = blockIdx.x, b = blockIdx.y / sl, s = blockIdx.y % sl;
  const int64_t l0 = (int64_t)s * chunk, l1 = min(L, l0 + chunk);
  const float* __restrict__ row = x + ((int64_t)b * C + c) * L;
  // deviations from a per-channel pivot (the channel's first element): keeps sum-of-squares cancellation harmless even
  // for a channel that is nearly constant at a large offset; partial means are stored relative to the pivot
  const float pivot = x[(int64_t)c * L];
  float sum = 0.f, sq = 0.f, cnt = 0.f;
  for (int64_t l = l0 + (int64_t)threadIdx.x * 4; l < l1; l += TB * 4) {
    float4 v = *reinterpret_cast<const float4*>(row + l);
    v.x -= pivot; v.y -= pivot; v.z -= pivot; v.w -= pivot;
    sum += (v.x + v.y) + (v.z + v.w);
    sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, sq))));
    cnt += 4.f;
  }
  Welford w;
  w.n = cnt;
  w.mean = cnt > 0.f ? sum / cnt : 0.f;
  w.m2 = cnt > 0.f ? fmaxf(sq - sum * w.mean, 0.f) : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Welford t;
    t.n = __shfl_xor_sync(FULL, w.n, o);
    t.mean = __shfl_xor_sync(FULL, w.mean, o);
    t.m2 = __shfl_xor_sync(FULL, w.m2, o);
    w = merge(w, t);
  }
  __shared__ Welford sh[TB / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    Welford r = sh[0];
    for (int i = 1; i < TB / 32; ++i) r = merge(r, sh[i]);
    float* o = partial + ((int64_t)c * gridDim.y + blockIdx.y) * 3;
    o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
  }
}

// ---- forward, pass 1b: one warp per channel merges the partials (fp64), writes mean / invstd / scale / shift and
// updates the running statistics exactly like torch (momentum, unbiased variance) ---------------------------------
__global__ void __launch_bounds__(128)
bn_finalize_kernel(const float* __restrict__ x, int64_t L, const float* __restrict__ partial, int P, int C,
                   const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                   float* __restrict__ running_var, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                   float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int p = lane; p < P; p += 32) {
    const float* q = partial + ((int64_t)c * P + p) * 3;
    const double nb = q[0], mb = q[1], m2b = q[2];
    if (nb == 0.0) continue;
    const double nn = n + nb, d = mb - mean;
    mean += d * nb / nn;
    m2 += m2b + d * d * n * nb / nn;
    n = nn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double nb = __shfl_xor_sync(FULL, n, o), mb = __shfl_xor_sync(FULL, mean, o), m2b = __shfl_xor_sync(FULL, m2, o);
    const double nn = n + nb;
    if (nn > 0.0) {
      const double d = mb - mean;
      mean += d * nb / nn;
      m2 += m2b + d * d * n * nb / nn;
      n = nn;
    }
  }
  if (lane == 0) {
    const double var = m2 / n;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float mu = (float)(mean + (double)x[(int64_t)c * L]);   // partial means are relative to the pivot
    save_mean[c] = mu;
    save_invstd[c] = invstd;
    const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    const float sc = g * invstd;
    scale[c] = sc;
    shift[c] = fmaf(-mu, sc, bt);
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(m2 / fmax(n - 1.0, 1.0));
  }
}

// ---- forward, pass 2: y = [relu](fma(x, scale[c], shift[c])) -----------------------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(TB)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, int C,
                int64_t L, float* __restrict__ y) {
  const int bc = blockIdx.y, c = bc % C;
  const float sc = scale[c], sh = shift[c];
  const float* __restrict__ row = x + (int64_t)bc * L;
  float* __restrict__ out = y + (int64_t)bc * L;
  for (int64_t l = ((int64_t)blockIdx.x * TB + threadIdx.x) * 4; l < L; l += (int64_t)gridDim.x * TB * 4) {
    float4 v = *reinterpret_cast<const float4*>(row + l);
    v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
    if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4*>(out + l) = v;
  }
}

// ---- backward, pass 1: partial sums of g = dy * [y > 0] and g * xhat ---------------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(TB)
bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ save_mean,
                     const float* __restrict__ save_invstd, const float* __restrict__ scale,
                     const float* __restrict__ shift, int C, int64_t L, int64_t chunk, int sl,
                     float* __restrict__ partial) {
  const int c = blockIdx.x, b = blockIdx.y / sl, s = blockIdx.y % sl;
  const int64_t l0 = (int64_t)s * chunk, l1 = min(L, l0 + chunk);
  const int64_t base = ((int64_t)b * C + c) * L;
  const float mu = save_mean[c], is = save_invstd[c], sc = scale[c], sh = shift[c];
  float s1 = 0.f, s2 = 0.f;
  for (int64_t l = l0 + (int64_t)threadIdx.x * 4; l < l1; l += TB * 4) {
    const float4 xv = *reinterpret_cast<const float4*>(x + base + l);
    float4 g = *reinterpret_cast<const float4*>(dy + base + l);
    if (RELU) {
      g.x = fmaf(xv.x, sc, sh) > 0.f ? g.x : 0.f; g.y = fmaf(xv.y, sc, sh) > 0.f ? g.y : 0.f;
      g.z = fmaf(xv.z, sc, sh) > 0.f ? g.z : 0.f; g.w = fmaf(xv.w, sc, sh) > 0.f ? g.w : 0.f;
    }
    s1 += (g.x + g.y) + (g.z + g.w);
    s2 = fmaf(g.x, (xv.x - mu) * is, fmaf(g.y, (xv.y - mu) * is, fmaf(g.z, (xv.z - mu) * is, fmaf(g.w, (xv.w - mu) * is, s2))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(FULL, s1, o);
    s2 += __shfl_xor_sync(FULL, s2, o);
  }
  __shared__ float sh1[TB / 32], sh2[TB / 32];
  if ((threadIdx.x & 31) == 0) { sh1[threadIdx.x >> 5] = s1; sh2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bsum = 0.f;
    for (int i = 0; i < TB / 32; ++i) { a += sh1[i]; bsum += sh2[i]; }
    float* o = partial + ((int64_t)c * gridDim.y + blockIdx.y) * 2;
    o[0] = a; o[1] = bsum;
  }
}

__global__ void __launch_bounds__(128)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int P, int C, double count, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ k1, float* __restrict__ k2) {
  const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int p = lane; p < P; p += 32) {
    a += partial[((int64_t)c * P + p) * 2];
    b += partial[((int64_t)c * P + p) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(FULL, a, o);
    b += __shfl_xor_sync(FULL, b, o);
  }
  if (lane == 0) {
    dbeta[c] = (float)a;
    dgamma[c] = (float)b;
    k1[c] = (float)(a / count);
    k2[c] = (float)(b / count);
  }
}

// ---- backward, pass 2: dx = scale * (g - mean(g) - xhat * mean(g * xhat)) ----------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(TB)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ save_mean,
                    const float* __restrict__ save_invstd, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ k1, const float* __restrict__ k2, int C,
                    int64_t L, float* __restrict__ dx) {
  const int bc = blockIdx.y, c = bc % C;
  const float mu = save_mean[c], is = save_invstd[c], sc = scale[c], sh = shift[c], m1 = k1[c], m2 = k2[c];
  const int64_t base = (int64_t)bc * L;
  for (int64_t l = ((int64_t)blockIdx.x * TB + threadIdx.x) * 4; l < L; l += (int64_t)gridDim.x * TB * 4) {
    const float4 xv = *reinterpret_cast<const float4*>(x + base + l);
    float4 g = *reinterpret_cast<const float4*>(dy + base + l);
    if (RELU) {
      g.x = fmaf(xv.x, sc, sh) > 0.f ? g.x : 0.f; g.y = fmaf(xv.y, sc, sh) > 0.f ? g.y : 0.f;
      g.z = fmaf(xv.z, sc, sh) > 0.f ? g.z : 0.f; g.w = fmaf(xv.w, sc, sh) > 0.f ? g.w : 0.f;
    }
    float4 o;
    o.x = sc * (g.x - m1 - (xv.x - mu) * is * m2);
    o.y = sc * (g.y - m1 - (xv.y - mu) * is * m2);
    o.z = sc * (g.z - m1 - (xv.z - mu) * is * m2);
    o.w = sc * (g.w - m1 - (xv.w - mu) * is * m2);
    *reinterpret_cast<float4*>(dx + base + l) = o;
  }
}

// ---- max over the innermost K = 64 (one half-warp per row: 16 lanes x float4) ----------------------------------------------
__global__ void __launch_bounds__(TB)
maxpool64_fwd_kernel(const float* __restrict__ x, int64_t rows, float* __restrict__ out, uint8_t* __restrict__ arg) {
  const int lane16 = threadIdx.x & 15, half = (threadIdx.x >> 4) & 1;
  const int64_t warp = ((int64_t)blockIdx.x * TB + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * TB) >> 5;
  for (int64_t r2 = warp * 2; r2 < rows; r2 += nwarps * 2) {   // warp-uniform trip count (full-mask shuffles below)
    const int64_t r = r2 + half;
    const bool live = r < rows;
    const float4 v = live ? *reinterpret_cast<const float4*>(x + r * 64 + lane16 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    // first maximum wins (torch.max returns one of the maxima; ties are duplicates of one neighbour or zeros, and
    // every choice back-propagates to the same source point)
    float best = v.x; int k = lane16 * 4;
    if (v.y > best) { best = v.y; k = lane16 * 4 + 1; }
    if (v.z > best) { best = v.z; k = lane16 * 4 + 2; }
    if (v.w > best) { best = v.w; k = lane16 * 4 + 3; }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(FULL, best, o);
      const int ok = __shfl_xor_sync(FULL, k, o);
      if (ob > best || (ob == best && ok < k)) { best = ob; k = ok; }
    }
    if (live && lane16 == 0) { out[r] = best; arg[r] = (uint8_t)k; }
  }
}

__global__ void __launch_bounds__(TB)
maxpool64_bwd_kernel(const float* __restrict__ dout, const uint8_t* __restrict__ arg, int64_t rows, float* __restrict__ dx) {
  const int lane16 = threadIdx.x & 15;
  const int64_t hw = ((int64_t)blockIdx.x * TB + threadIdx.x) >> 4, nhw = ((int64_t)gridDim.x * TB) >> 4;
  for (int64_t r = hw; r < rows; r += nhw) {
    const float g = dout[r];
    const int k = arg[r] - lane16 * 4;
    float4 o;
    o.x = k == 0 ? g : 0.f; o.y = k == 1 ? g : 0.f; o.z = k == 2 ? g : 0.f; o.w = k == 3 ? g : 0.f;
    *reinterpret_cast<float4*>(dx + r * 64 + lane16 * 4) = o;
  }
}

// ---- pooled layer (last block of an SA module): BN + ReLU + max over the 64 neighbours without materialising y ----------
// forward pass 2: out[b,c,m] = max_k relu(fma(x[b,c,m,k], scale, shift)), arg-max kept as a byte
template <bool RELU>
__global__ void __launch_bounds__(TB)
bn_apply_max64_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, int C,
                      int64_t M, float* __restrict__ out, uint8_t* __restrict__ arg) {
  const int bc = blockIdx.y, c = bc % C;
  const float sc = scale[c], sh = shift[c];
  const int lane16 = threadIdx.x & 15, half = (threadIdx.x >> 4) & 1;
  const int64_t warp = ((int64_t)blockIdx.x * TB + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * TB) >> 5;
  const float* __restrict__ xr = x + (int64_t)bc * M * 64;
  for (int64_t m2 = warp * 2; m2 < M; m2 += nwarps * 2) {
    const int64_t m = m2 + half;
    const bool live = m < M;
    float4 v = live ? *reinterpret_cast<const float4*>(xr + m * 64 + lane16 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
    if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    float best = v.x; int k = lane16 * 4;
    if (v.y > best) { best = v.y; k = lane16 * 4 + 1; }
    if (v.z > best) { best = v.z; k = lane16 * 4 + 2; }
    if (v.w > best) { best = v.w; k = lane16 * 4 + 3; }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(FULL, best, o);
      const int ok = __shfl_xor_sync(FULL, k, o);
      if (ob > best || (ob == best && ok < k)) { best = ob; k = ok; }
    }
    if (live && lane16 == 0) { out[(int64_t)bc * M + m] = best; arg[(int64_t)bc * M + m] = (uint8_t)k; }
  }
}

// backward pass 1: the gradient w.r.t. y is dout at the arg-max (if it passed the ReLU) and 0 elsewhere
template <bool RELU>
__global__ void __launch_bounds__(TB)
bn_max64_bwd_reduce_kernel(const float* __restrict__ dout, const uint8_t* __restrict__ arg, const float* __restrict__ x,
                           const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                           const float* __restrict__ scale, const float* __restrict__ shift, int C, int64_t M,
                           float* __restrict__ partial) {
  const int c = blockIdx.x, b = blockIdx.y;
  const int64_t bc = (int64_t)b * C + c;
  const float mu = save_mean[c], is = save_invstd[c], sc = scale[c], sh = shift[c];
  float s1 = 0.f, s2 = 0.f;
  for (int64_t m = threadIdx.x; m < M; m += TB) {
    const float xv = x[(bc * M + m) * 64 + arg[bc * M + m]];
    const float g = (!RELU || fmaf(xv, sc, sh) > 0.f) ? dout[bc * M + m] : 0.f;
    s1 += g;
    s2 = fmaf(g, (xv - mu) * is, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(FULL, s1, o);
    s2 += __shfl_xor_sync(FULL, s2, o);
  }
  __shared__ float sh1[TB / 32], sh2[TB / 32];
  if ((threadIdx.x & 31) == 0) { sh1[threadIdx.x >> 5] = s1; sh2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bsum = 0.f;
    for (int i = 0; i < TB / 32; ++i) { a += sh1[i]; bsum += sh2[i]; }
    float* o = partial + ((int64_t)c * gridDim.y + blockIdx.y) * 2;
    o[0] = a; o[1] = bsum;
  }
}

// backward pass 2: dx = scale * (g - mean(g) - xhat * mean(g * xhat)) for every position
template <bool RELU>
__global__ void __launch_bounds__(TB)
bn_max64_bwd_apply_kernel(const float* __restrict__ dout, const uint8_t* __restrict__ arg, const float* __restrict__ x,
                          const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                          const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ k1,
                          const float* __restrict__ k2, int C, int64_t M, float* __restrict__ dx) {
  const int bc = blockIdx.y, c = bc % C;
  const float mu = save_mean[c], is = save_invstd[c], sc = scale[c], sh = shift[c], m1 = k1[c], m2 = k2[c];
  const int lane16 = threadIdx.x & 15;
  const int64_t hw = ((int64_t)blockIdx.x * TB + threadIdx.x) >> 4, nhw = ((int64_t)gridDim.x * TB) >> 4;
  for (int64_t m = hw; m < M; m += nhw) {
    const int64_t r = (int64_t)bc * M + m;
    const float4 xv = *reinterpret_cast<const float4*>(x + r * 64 + lane16 * 4);
    const float dp = dout[r];
    const int k = (int)arg[r] - lane16 * 4;
    float4 g;
    g.x = (k == 0 && (!RELU || fmaf(xv.x, sc, sh) > 0.f)) ? dp : 0.f;
    g.y = (k == 1 && (!RELU || fmaf(xv.y, sc, sh) > 0.f)) ? dp : 0.f;
    g.z = (k == 2 && (!RELU || fmaf(xv.z, sc, sh) > 0.f)) ? dp : 0.f;
    g.w = (k == 3 && (!RELU || fmaf(xv.w, sc, sh) > 0.f)) ? dp : 0.f;
    float4 o;
    o.x = sc * (g.x - m1 - (xv.x - mu) * is * m2);
    o.y = sc * (g.y - m1 - (xv.y - mu) * is * m2);
    o.z = sc * (g.z - m1 - (xv.z - mu) * is * m2);
    o.w = sc * (g.w - m1 - (xv.w - mu) * is * m2);
    *reinterpret_cast<float4*>(dx + r * 64 + lane16 * 4) = o;
  }
}


// ---- fused-chain variants (conv_train.cu produces Z and its moments; the MLP chain keeps activations as bf16 planes);
// the dropout mask and plane helpers live in train_common.cuh -----------------------------------------------------------
// per-channel (sum, sum of squares) in fp64 -> mean / invstd / scale / shift + running statistics (as bn_finalize_kernel)
__global__ void __launch_bounds__(128)
bn_finalize_moments_kernel(const double* __restrict__ moments, int C, double count, const float* __restrict__ gamma,
                           const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                           float* __restrict__ running_var, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                           float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= C) return;
  const double mean = moments[2 * c] / count;
  const double var = fmax(moments[2 * c + 1] / count - mean * mean, 0.0);
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float mu = (float)mean;
  save_mean[c] = mu;
  save_invstd[c] = invstd;
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = fmaf(-mu, sc, bt);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * count / fmax(count - 1.0, 1.0));
}

// sums (C, 2) in fp64 = (sum g, sum g * xhat) accumulated by the dgrad epilogue (conv_train.cu BnReduce) -> dgamma, dbeta and
// the two batch means the apply pass subtracts
__global__ void __launch_bounds__(128)
bn_bwd_finalize_sums_kernel(const double* __restrict__ sums, int C, double count, float* __restrict__ dgamma,
                            float* __restrict__ dbeta, float* __restrict__ k1, float* __restrict__ k2) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= C) return;
  const double a = sums[2 * c], b = sums[2 * c + 1];
  dbeta[c] = (float)a;
  dgamma[c] = (float)b;
  k1[c] = (float)(a / count);
  k2[c] = (float)(b / count);
}

// y = dropout([relu](fma(z, scale, shift))) written as fp32 and / or as bf16 hi/lo planes
template <bool RELU, bool DROP>
__global__ void __launch_bounds__(TB)
bn_apply_ex_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, int C,
                   int64_t L, DropCfg dc, float* __restrict__ y, __nv_bfloat16* __restrict__ hi,
                   __nv_bfloat16* __restrict__ lo) {
  const int bc = blockIdx.y, c = bc % C;
  const float sc = scale[c], sh = shift[c];
  const int64_t base = (int64_t)bc * L;
  for (int64_t l = ((int64_t)blockIdx.x * TB + threadIdx.x) * 4; l < L; l += (int64_t)gridDim.x * TB * 4) {
    float4 v = *reinterpret_cast<const float4*>(x + base + l);
    v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
    if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (DROP) {
      const float4 m = drop_mult((uint64_t)(base + l) >> 2, dc.seed, dc.thr, dc.inv_keep);
      v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
    }
    if (y) *reinterpret_cast<float4*>(y + base + l) = v;
    if (hi) store_planes4(hi, lo, base + l, v);
  }
}

template <bool RELU, bool DROP>
__global__ void __launch_bounds__(TB)
bn_bwd_reduce_ex_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ save_mean,
                        const float* __restrict__ save_invstd, const float* __restrict__ scale,
                        const float* __restrict__ shift, int C, int64_t L, int64_t chunk, int sl, DropCfg dc,
                        float* __restrict__ partial) {
  const int c = blockIdx.x, b = blockIdx.y / sl, s = blockIdx.y % sl;
  const int64_t l0 = (int64_t)s * chunk, l1 = min(L, l0 + chunk);
  const int64_t base = ((int64_t)b * C + c) * L;
  const float mu = save_mean[c], is = save_invstd[c], sc = scale[c], sh = shift[c];
  float s1 = 0.f, s2 = 0.f;
  for (int64_t l = l0 + (int64_t)threadIdx.x * 4; l < l1; l += TB * 4) {
    const float4 xv = *reinterpret_cast<const float4*>(x + base + l);
    float4 g = *reinterpret_cast<const float4*>(dy + base + l);
    if (DROP) {
      const float4 m = drop_mult((uint64_t)(base + l) >> 2, dc.seed, dc.thr, dc.inv_keep);
      g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
    }
    if (RELU) {
      g.x = fmaf(xv.x, sc, sh) > 0.f ? g.x : 0.f; g.y = fmaf(xv.y, sc, sh) > 0.f ? g.y : 0.f;
      g.z = fmaf(xv.z, sc, sh) > 0.f ? g.z : 0.f; g.w = fmaf(xv.w, sc, sh) > 0.f ? g.w : 0.f;
    }
    s1 += (g.x + g.y) + (g.z + g.w);
    s2 = fmaf(g.x, (xv.x - mu) * is, fmaf(g.y, (xv.y - mu) * is, fmaf(g.z, (xv.z - mu) * is, fmaf(g.w, (xv.w - mu) * is, s2))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(FULL, s1, o);
    s2 += __shfl_xor_sync(FULL, s2, o);
  }
  __shared__ float sh1[TB / 32], sh2[TB / 32];
  if ((threadIdx.x & 31) == 0) { sh1[threadIdx.x >> 5] = s1; sh2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bsum = 0.f;
    for (int i = 0; i < TB / 32; ++i) { a += sh1[i]; bsum += sh2[i]; }
    float* o = partial + ((int64_t)c * gridDim.y + blockIdx.y) * 2;
    o[0] = a; o[1] = bsum;
  }
}

template <bool RELU, bool DROP>
__global__ void __launch_bounds__(TB)
bn_bwd_apply_ex_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ save_mean,
                       const float* __restrict__ save_invstd, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ k1, const float* __restrict__ k2, int C,
                       int64_t L, DropCfg dc, float* __restrict__ dx, __nv_bfloat16* __restrict__ dhi,
                       __nv_bfloat16* __restrict__ dlo) {
  const int bc = blockIdx.y, c = bc % C;
  const float mu = save_mean[c], is = save_invstd[c], sc = scale[c], sh = shift[c], m1 = k1[c], m2 = k2[c];
  const int64_t base = (int64_t)bc * L;
  for (int64_t l = ((int64_t)blockIdx.x * TB + threadIdx.x) * 4; l < L; l += (int64_t)gridDim.x * TB * 4) {
    const float4 xv = *reinterpret_cast<const float4*>(x + base + l);
    float4 g = *reinterpret_cast<const float4*>(dy + base + l);
    if (DROP) {
      const float4 m = drop_mult((uint64_t)(base + l) >> 2, dc.seed, dc.thr, dc.inv_keep);
      g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
    }
    if (RELU) {
      g.x = fmaf(xv.x, sc, sh) > 0.f ? g.x : 0.f; g.y = fmaf(xv.y, sc, sh) > 0.f ? g.y : 0.f;
      g.z = fmaf(xv.z, sc, sh) > 0.f ? g.z : 0.f; g.w = fmaf(xv.w, sc, sh) > 0.f ? g.w : 0.f;
    }
    float4 o;
    o.x = sc * (g.x - m1 - (xv.x - mu) * is * m2);
    o.y = sc * (g.y - m1 - (xv.y - mu) * is * m2);
    o.z = sc * (g.z - m1 - (xv.z - mu) * is * m2);
    o.w = sc * (g.w - m1 - (xv.w - mu) * is * m2);
    if (dx) *reinterpret_cast<float4*>(dx + base + l) = o;
    if (dhi) store_planes4(dhi, dlo, base + l, o);
  }
}

// pooled block, backward pass 2 with the gradient written as planes (and / or fp32)
template <bool RELU>
__global__ void __launch_bounds__(TB)
bn_max64_bwd_apply_ex_kernel(const float* __restrict__ dout, const uint8_t* __restrict__ arg, const float* __restrict__ x,
                             const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                             const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ k1,
                             const float* __restrict__ k2, int C, int64_t M, float* __restrict__ dx,
                             __nv_bfloat16* __restrict__ dhi, __nv_bfloat16* __restrict__ dlo) {
  const int bc = blockIdx.y, c = bc % C;
  const float mu = save_mean[c], is = save_invstd[c], sc = scale[c], sh = shift[c], m1 = k1[c], m2 = k2[c];
  const int lane16 = threadIdx.x & 15;
  const int64_t hw = ((int64_t)blockIdx.x * TB + threadIdx.x) >> 4, nhw = ((int64_t)gridDim.x * TB) >> 4;
  for (int64_t m = hw; m < M; m += nhw) {
    const int64_t r = (int64_t)bc * M + m;
    const float4 xv = *reinterpret_cast<const float4*>(x + r * 64 + lane16 * 4);
    const float dp = dout[r];
    const int k = (int)arg[r] - lane16 * 4;
    float4 g;
    g.x = (k == 0 && (!RELU || fmaf(xv.x, sc, sh) > 0.f)) ? dp : 0.f;
    g.y = (k == 1 && (!RELU || fmaf(xv.y, sc, sh) > 0.f)) ? dp : 0.f;
    g.z = (k == 2 && (!RELU || fmaf(xv.z, sc, sh) > 0.f)) ? dp : 0.f;
    g.w = (k == 3 && (!RELU || fmaf(xv.w, sc, sh) > 0.f)) ? dp : 0.f;
    float4 o;
    o.x = sc * (g.x - m1 - (xv.x - mu) * is * m2);
    o.y = sc * (g.y - m1 - (xv.y - mu) * is * m2);
    o.z = sc * (g.z - m1 - (xv.z - mu) * is * m2);
    o.w = sc * (g.w - m1 - (xv.w - mu) * is * m2);
    if (dx) *reinterpret_cast<float4*>(dx + r * 64 + lane16 * 4) = o;
    if (dhi) store_planes4(dhi, dlo, r * 64 + lane16 * 4, o);
  }
}


inline unsigned apply_grid_x(int64_t L, int rows_bc) {
  int64_t gx = (L / 4 + TB - 1) / TB;
  const int64_t cap = std::max<int64_t>(1, (148LL * 16 + rows_bc - 1) / rows_bc);   // ~16 resident blocks per SM overall
  return (unsigned)std::max<int64_t>(1, std::min(gx, cap));
}

int check_shape(int B, int C, int64_t L) {
  RN_CHECK_ARG(B > 0 && C > 0 && L > 0, "bn_relu_train: empty input");
  RN_CHECK_ARG(L % 4 == 0, "bn_relu_train: innermost extent %lld is not a multiple of 4", (long long)L);
  RN_CHECK_ARG((int64_t)B * plan_chunks(L).sl <= 65535 && (int64_t)B * C <= 65535,
               "bn_relu_train: B*C = %lld exceeds the grid limit of this build", (long long)B * C);
  if ((int64_t)B * L <= 1) {
    set_error("Expected more than 1 value per channel when training");   // torch's message
    return REGNET_EINVAL;
  }
  return REGNET_OK;
}

}  // namespace

}  // namespace regnet

using namespace regnet;

extern "C" {

int64_t regnet_bn_workspace_bytes(int B, int C, int64_t L) {
  if (B <= 0 || C <= 0 || L <= 0) return 0;
  return (int64_t)C * B * plan_chunks(L).sl * 3 * sizeof(float) + 2 * (int64_t)C * sizeof(float) + 256;
}

int regnet_bn_relu_train_forward(const float* x, int B, int C, int64_t L, const float* gamma, const float* beta, float eps,
                                 float momentum, int relu, float* running_mean, float* running_var, float* y,
                                 float* save_mean, float* save_invstd, float* scale, float* shift, void* workspace,
                                 int64_t workspace_bytes, void* stream_) {
  RN_CHECK_ARG(x && y && save_mean && save_invstd && scale && shift && workspace, "bn_relu_train_forward: null argument");
  RN_TRY(check_shape(B, C, L));
  RN_CHECK_ARG(workspace_bytes >= regnet_bn_workspace_bytes(B, C, L), "bn_relu_train_forward: workspace too small");
  cudaStream_t s = (cudaStream_t)stream_;
  const ChunkPlan cp = plan_chunks(L);
  const int P = B * cp.sl;
  float* partial = reinterpret_cast<float*>(workspace);
  bn_stats_kernel<<<dim3(C, P), TB, 0, s>>>(x, C, L, cp.chunk, cp.sl, partial);
  RN_LAUNCH_CHECK("bn_stats_kernel");
  bn_finalize_kernel<<<(C + 3) / 4, 128, 0, s>>>(x, L, partial, P, C, gamma, beta, eps, momentum, running_mean, running_var,
                                                 save_mean, save_invstd, scale, shift);
  RN_LAUNCH_CHECK("bn_finalize_kernel");
  const dim3 grid(apply_grid_x(L, B * C), B * C);
  if (relu) bn_apply_kernel<true><<<grid, TB, 0, s>>>(x, scale, shift, C, L, y);
  else bn_apply_kernel<false><<<grid, TB, 0, s>>>(x, scale, shift, C, L, y);
  RN_LAUNCH_CHECK("bn_apply_kernel");
  return REGNET_OK;
}

int regnet_bn_relu_train_backward(const float* dy, const float* x, int B, int C, int64_t L, const float* save_mean,
                                  const float* save_invstd, const float* scale, const float* shift, int relu, float* dx,
                                  float* dgamma, float* dbeta, void* workspace, int64_t workspace_bytes, void* stream_) {
  RN_CHECK_ARG(dy && x && save_mean && save_invstd && scale && shift && dx && dgamma && dbeta && workspace,
               "bn_relu_train_backward: null argument");
  RN_TRY(check_shape(B, C, L));
  RN_CHECK_ARG(workspace_bytes >= regnet_bn_workspace_bytes(B, C, L), "bn_relu_train_backward: workspace too small");
  cudaStream_t s = (cudaStream_t)stream_;
  const ChunkPlan cp = plan_chunks(L);
  const int P = B * cp.sl;
  float* partial = reinterpret_cast<float*>(workspace);
  float* k1 = partial + (int64_t)C * P * 3;
  float* k2 = k1 + C;
  if (relu) bn_bwd_reduce_kernel<true><<<dim3(C, P), TB, 0, s>>>(dy, x, save_mean, save_invstd, scale, shift, C, L, cp.chunk, cp.sl, partial);
  else bn_bwd_reduce_kernel<false><<<dim3(C, P), TB, 0, s>>>(dy, x, save_mean, save_invstd, scale, shift, C, L, cp.chunk, cp.sl, partial);
  RN_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  bn_bwd_finalize_kernel<<<(C + 3) / 4, 128, 0, s>>>(partial, P, C, (double)B * (double)L, dgamma, dbeta, k1, k2);
  RN_LAUNCH_CHECK("bn_bwd_finalize_kernel");
  const dim3 grid(apply_grid_x(L, B * C), B * C);
  if (relu) bn_bwd_apply_kernel<true><<<grid, TB, 0, s>>>(dy, x, save_mean, save_invstd, scale, shift, k1, k2, C, L, dx);
  else bn_bwd_apply_kernel<false><<<grid, TB, 0, s>>>(dy, x, save_mean, save_invstd, scale, shift, k1, k2, C, L, dx);
  RN_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return REGNET_OK;
}

int regnet_bn_relu_max64_train_forward(const float* x, int B, int C, int64_t M, const float* gamma, const float* beta,
                                       float eps, float momentum, int relu, float* running_mean, float* running_var,
                                       float* out, uint8_t* argmax, float* save_mean, float* save_invstd, float* scale,
                                       float* shift, void* workspace, int64_t workspace_bytes, void* stream_) {
  RN_CHECK_ARG(x && out && argmax && save_mean && save_invstd && scale && shift && workspace,
               "bn_relu_max64_train_forward: null argument");
  const int64_t L = M * 64;
  RN_TRY(check_shape(B, C, L));
  RN_CHECK_ARG(workspace_bytes >= regnet_bn_workspace_bytes(B, C, L), "bn_relu_max64_train_forward: workspace too small");
  cudaStream_t s = (cudaStream_t)stream_;
  const ChunkPlan cp = plan_chunks(L);
  const int P = B * cp.sl;
  float* partial = reinterpret_cast<float*>(workspace);
  bn_stats_kernel<<<dim3(C, P), TB, 0, s>>>(x, C, L, cp.chunk, cp.sl, partial);
  RN_LAUNCH_CHECK("bn_stats_kernel");
  bn_finalize_kernel<<<(C + 3) / 4, 128, 0, s>>>(x, L, partial, P, C, gamma, beta, eps, momentum, running_mean, running_var,
                                                 save_mean, save_invstd, scale, shift);
  RN_LAUNCH_CHECK("bn_finalize_kernel");
  const dim3 grid(apply_grid_x(L, B * C), B * C);
  if (relu) bn_apply_max64_kernel<true><<<grid, TB, 0, s>>>(x, scale, shift, C, M, out, argmax);
  else bn_apply_max64_kernel<false><<<grid, TB, 0, s>>>(x, scale, shift, C, M, out, argmax);
  RN_LAUNCH_CHECK("bn_apply_max64_kernel");
  return REGNET_OK;
}

int regnet_bn_relu_max64_train_backward(const float* dout, const uint8_t* argmax, const float* x, int B, int C, int64_t M,
                                        const float* save_mean, const float* save_invstd, const float* scale,
                                        const float* shift, int relu, float* dx, float* dgamma, float* dbeta,
                                        void* workspace, int64_t workspace_bytes, void* stream_) {
  RN_CHECK_ARG(dout && argmax && x && save_mean && save_invstd && scale && shift && dx && dgamma && dbeta && workspace,
               "bn_relu_max64_train_backward: null argument");
  const int64_t L = M * 64;
  RN_TRY(check_shape(B, C, L));
  RN_CHECK_ARG(workspace_bytes >= regnet_bn_workspace_bytes(B, C, L), "bn_relu_max64_train_backward: workspace too small");
  cudaStream_t s = (cudaStream_t)stream_;
  float* partial = reinterpret_cast<float*>(workspace);
  float* k1 = partial + (int64_t)C * B * plan_chunks(L).sl * 3;
  float* k2 = k1 + C;
  if (relu) bn_max64_bwd_reduce_kernel<true><<<dim3(C, B), TB, 0, s>>>(dout, argmax, x, save_mean, save_invstd, scale, shift, C, M, partial);
  else bn_max64_bwd_reduce_kernel<false><<<dim3(C, B), TB, 0, s>>>(dout, argmax, x, save_mean, save_invstd, scale, shift, C, M, partial);
  RN_LAUNCH_CHECK("bn_max64_bwd_reduce_kernel");
  bn_bwd_finalize_kernel<<<(C + 3) / 4, 128, 0, s>>>(partial, B, C, (double)B * (double)L, dgamma, dbeta, k1, k2);
  RN_LAUNCH_CHECK("bn_bwd_finalize_kernel");
  const dim3 grid(apply_grid_x(L, B * C), B * C);
  if (relu) bn_max64_bwd_apply_kernel<true><<<grid, TB, 0, s>>>(dout, argmax, x, save_mean, save_invstd, scale, shift, k1, k2, C, M, dx);
  else bn_max64_bwd_apply_kernel<false><<<grid, TB, 0, s>>>(dout, argmax, x, save_mean, save_invstd, scale, shift, k1, k2, C, M, dx);
  RN_LAUNCH_CHECK("bn_max64_bwd_apply_kernel");
  return REGNET_OK;
}

int regnet_bn_finalize_moments(const double* moments, int C, double count, const float* gamma, const float* beta, float eps,
                               float momentum, float* running_mean, float* running_var, float* save_mean,
                               float* save_invstd, float* scale, float* shift, void* stream_) {
  RN_CHECK_ARG(moments && save_mean && save_invstd && scale && shift && C > 0, "bn_finalize_moments: null argument");
  if (count <= 1.0) {
    set_error("Expected more than 1 value per channel when training");
    return REGNET_EINVAL;
  }
  bn_finalize_moments_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(
      moments, C, count, gamma, beta, eps, momentum, running_mean, running_var, save_mean, save_invstd, scale, shift);
  RN_LAUNCH_CHECK("bn_finalize_moments_kernel");
  return REGNET_OK;
}

int regnet_bn_apply_ex(const float* z, int B, int C, int64_t L, const float* scale, const float* shift, int relu,
                       float drop_p, uint64_t drop_seed, float* y, void* y_hi, void* y_lo, void* stream_) {
  RN_CHECK_ARG(z && scale && shift && (y || (y_hi && y_lo)), "bn_apply_ex: null argument");
  RN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "bn_apply_ex: dropout probability %f", drop_p);
  RN_TRY(check_shape(B, C, L));
  cudaStream_t s = (cudaStream_t)stream_;
  const DropCfg dc = make_drop(drop_p, drop_seed);
  const dim3 grid(apply_grid_x(L, B * C), B * C);
  __nv_bfloat16* hi = (__nv_bfloat16*)y_hi;
  __nv_bfloat16* lo = (__nv_bfloat16*)y_lo;
  if (dc.thr) {
    if (relu) bn_apply_ex_kernel<true, true><<<grid, TB, 0, s>>>(z, scale, shift, C, L, dc, y, hi, lo);
    else bn_apply_ex_kernel<false, true><<<grid, TB, 0, s>>>(z, scale, shift, C, L, dc, y, hi, lo);
  } else {
    if (relu) bn_apply_ex_kernel<true, false><<<grid, TB, 0, s>>>(z, scale, shift, C, L, dc, y, hi, lo);
    else bn_apply_ex_kernel<false, false><<<grid, TB, 0, s>>>(z, scale, shift, C, L, dc, y, hi, lo);
  }
  RN_LAUNCH_CHECK("bn_apply_ex_kernel");
  return REGNET_OK;
}

int regnet_bn_backward_ex(const float* dy, const float* z, int B, int C, int64_t L, const float* save_mean,
                          const float* save_invstd, const float* scale, const float* shift, int relu, float drop_p,
                          uint64_t drop_seed, float* dz, void* dz_hi, void* dz_lo, float* dgamma, float* dbeta,
                          void* workspace, int64_t workspace_bytes, void* stream_) {
  RN_CHECK_ARG(dy && z && save_mean && save_invstd && scale && shift && (dz || (dz_hi && dz_lo)) && dgamma && dbeta && workspace,
               "bn_backward_ex: null argument");
  RN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "bn_backward_ex: dropout probability %f", drop_p);
  RN_TRY(check_shape(B, C, L));
  RN_CHECK_ARG(workspace_bytes >= regnet_bn_workspace_bytes(B, C, L), "bn_backward_ex: workspace too small");
  cudaStream_t s = (cudaStream_t)stream_;
  const DropCfg dc = make_drop(drop_p, drop_seed);
  const ChunkPlan cp = plan_chunks(L);
  const int P = B * cp.sl;
  float* partial = reinterpret_cast<float*>(workspace);
  float* k1 = partial + (int64_t)C * P * 3;
  float* k2 = k1 + C;
  const dim3 rgrid(C, P);
#define RN_BWD_REDUCE(R, D) \
  bn_bwd_reduce_ex_kernel<R, D><<<rgrid, TB, 0, s>>>(dy, z, save_mean, save_invstd, scale, shift, C, L, cp.chunk, cp.sl, dc, partial)
  if (dc.thr) { if (relu) RN_BWD_REDUCE(true, true); else RN_BWD_REDUCE(false, true); }
  else { if (relu) RN_BWD_REDUCE(true, false); else RN_BWD_REDUCE(false, false); }
#undef RN_BWD_REDUCE
  RN_LAUNCH_CHECK("bn_bwd_reduce_ex_kernel");
  bn_bwd_finalize_kernel<<<(C + 3) / 4, 128, 0, s>>>(partial, P, C, (double)B * (double)L, dgamma, dbeta, k1, k2);
  RN_LAUNCH_CHECK("bn_bwd_finalize_kernel");
  const dim3 grid(apply_grid_x(L, B * C), B * C);
  __nv_bfloat16* hi = (__nv_bfloat16*)dz_hi;
  __nv_bfloat16* lo = (__nv_bfloat16*)dz_lo;
#define RN_BWD_APPLY(R, D) \
  bn_bwd_apply_ex_kernel<R, D><<<grid, TB, 0, s>>>(dy, z, save_mean, save_invstd, scale, shift, k1, k2, C, L, dc, dz, hi, lo)
  if (dc.thr) { if (relu) RN_BWD_APPLY(true, true); else RN_BWD_APPLY(false, true); }
  else { if (relu) RN_BWD_APPLY(true, false); else RN_BWD_APPLY(false, false); }
#undef RN_BWD_APPLY
  RN_LAUNCH_CHECK("bn_bwd_apply_ex_kernel");
  return REGNET_OK;
}

int regnet_bn_backward_from_sums(const float* dy, const float* z, int B, int C, int64_t L, const float* save_mean,
                                 const float* save_invstd, const float* scale, const float* shift, int relu, float drop_p,
                                 uint64_t drop_seed, const double* sums, float* dz, void* dz_hi, void* dz_lo, float* dgamma,
                                 float* dbeta, void* workspace, int64_t workspace_bytes, void* stream_) {
  RN_CHECK_ARG(dy && z && save_mean && save_invstd && scale && shift && sums && (dz || (dz_hi && dz_lo)) && dgamma && dbeta &&
                   workspace, "bn_backward_from_sums: null argument");
  RN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "bn_backward_from_sums: dropout probability %f", drop_p);
  RN_TRY(check_shape(B, C, L));
  RN_CHECK_ARG(workspace_bytes >= (int64_t)(2 * C * sizeof(float)), "bn_backward_from_sums: workspace too small");
  cudaStream_t s = (cudaStream_t)stream_;
  const DropCfg dc = make_drop(drop_p, drop_seed);
  float* k1 = reinterpret_cast<float*>(workspace);
  float* k2 = k1 + C;
  bn_bwd_finalize_sums_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, C, (double)B * (double)L, dgamma, dbeta, k1, k2);
  RN_LAUNCH_CHECK("bn_bwd_finalize_sums_kernel");
  const dim3 grid(apply_grid_x(L, B * C), B * C);
  __nv_bfloat16* hi = (__nv_bfloat16*)dz_hi;
  __nv_bfloat16* lo = (__nv_bfloat16*)dz_lo;
#define RN_BWD_APPLY(R, D) \
  bn_bwd_apply_ex_kernel<R, D><<<grid, TB, 0, s>>>(dy, z, save_mean, save_invstd, scale, shift, k1, k2, C, L, dc, dz, hi, lo)
  if (dc.thr) { if (relu) RN_BWD_APPLY(true, true); else RN_BWD_APPLY(false, true); }
  else { if (relu) RN_BWD_APPLY(true, false); else RN_BWD_APPLY(false, false); }
#undef RN_BWD_APPLY
  RN_LAUNCH_CHECK("bn_bwd_apply_ex_kernel");
  return REGNET_OK;
}

int regnet_bn_apply_max64(const float* z, int B, int C, int64_t M, const float* scale, const float* shift, int relu,
                          float* out, uint8_t* argmax, void* stream_) {
  RN_CHECK_ARG(z && scale && shift && out && argmax, "bn_apply_max64: null argument");
  const int64_t L = M * 64;
  RN_TRY(check_shape(B, C, L));
  cudaStream_t s = (cudaStream_t)stream_;
  const dim3 grid(apply_grid_x(L, B * C), B * C);
  if (relu) bn_apply_max64_kernel<true><<<grid, TB, 0, s>>>(z, scale, shift, C, M, out, argmax);
  else bn_apply_max64_kernel<false><<<grid, TB, 0, s>>>(z, scale, shift, C, M, out, argmax);
  RN_LAUNCH_CHECK("bn_apply_max64_kernel");
  return REGNET_OK;
}

int regnet_bn_max64_backward_ex(const float* dout, const uint8_t* argmax, const float* z, int B, int C, int64_t M,
                                const float* save_mean, const float* save_invstd, const float* scale, const float* shift,
                                int relu, float* dz, void* dz_hi, void* dz_lo, float* dgamma, float* dbeta, void* workspace,
                                int64_t workspace_bytes, void* stream_) {
  RN_CHECK_ARG(dout && argmax && z && save_mean && save_invstd && scale && shift && (dz || (dz_hi && dz_lo)) && dgamma &&
                   dbeta && workspace, "bn_max64_backward_ex: null argument");
  const int64_t L = M * 64;
  RN_TRY(check_shape(B, C, L));
  RN_CHECK_ARG(workspace_bytes >= regnet_bn_workspace_bytes(B, C, L), "bn_max64_backward_ex: workspace too small");
  cudaStream_t s = (cudaStream_t)stream_;
  float* partial = reinterpret_cast<float*>(workspace);
  float* k1 = partial + (int64_t)C * B * plan_chunks(L).sl * 3;
  float* k2 = k1 + C;
  if (relu) bn_max64_bwd_reduce_kernel<true><<<dim3(C, B), TB, 0, s>>>(dout, argmax, z, save_mean, save_invstd, scale, shift, C, M, partial);
  else bn_max64_bwd_reduce_kernel<false><<<dim3(C, B), TB, 0, s>>>(dout, argmax, z, save_mean, save_invstd, scale, shift, C, M, partial);
  RN_LAUNCH_CHECK("bn_max64_bwd_reduce_kernel");
  bn_bwd_finalize_kernel<<<(C + 3) / 4, 128, 0, s>>>(partial, B, C, (double)B * (double)L, dgamma, dbeta, k1, k2);
  RN_LAUNCH_CHECK("bn_bwd_finalize_kernel");
  const dim3 grid(apply_grid_x(L, B * C), B * C);
  __nv_bfloat16* hi = (__nv_bfloat16*)dz_hi;
  __nv_bfloat16* lo = (__nv_bfloat16*)dz_lo;
  if (relu) bn_max64_bwd_apply_ex_kernel<true><<<grid, TB, 0, s>>>(dout, argmax, z, save_mean, save_invstd, scale, shift, k1, k2, C, M, dz, hi, lo);
  else bn_max64_bwd_apply_ex_kernel<false><<<grid, TB, 0, s>>>(dout, argmax, z, save_mean, save_invstd, scale, shift, k1, k2, C, M, dz, hi, lo);
  RN_LAUNCH_CHECK("bn_max64_bwd_apply_ex_kernel");
  return REGNET_OK;
}

int regnet_maxpool64_forward(const float* x, int64_t rows, float* out, uint8_t* argmax, void* stream_) {
  RN_CHECK_ARG(x && out && argmax, "maxpool64_forward: null argument");
  if (rows <= 0) return REGNET_OK;
  const int64_t blocks = std::min<int64_t>((rows * 16 + TB - 1) / TB, 148LL * 32);
  maxpool64_fwd_kernel<<<(unsigned)blocks, TB, 0, (cudaStream_t)stream_>>>(x, rows, out, argmax);
  RN_LAUNCH_CHECK("maxpool64_fwd_kernel");
  return REGNET_OK;
}

int regnet_maxpool64_backward(const float* dout, const uint8_t* argmax, int64_t rows, float* dx, void* stream_) {
  RN_CHECK_ARG(dout && argmax && dx, "maxpool64_backward: null argument");
  if (rows <= 0) return REGNET_OK;
  const int64_t blocks = std::min<int64_t>((rows * 16 + TB - 1) / TB, 148LL * 32);
  maxpool64_bwd_kernel<<<(unsigned)blocks, TB, 0, (cudaStream_t)stream_>>>(dout, argmax, rows, dx);
  RN_LAUNCH_CHECK("maxpool64_bwd_kernel");
  return REGNET_OK;
}

}  // extern "C"
