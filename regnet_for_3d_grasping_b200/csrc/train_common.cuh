// Helpers shared by the training kernels (train_ops.cu, conv_train.cu): the counter-based dropout mask and the bf16
// hi/lo plane store.
#pragma once
#include "common.cuh"

namespace regnet {

// Dropout (nn/modules/mlp.py:101-105: F.dropout after every block of the seg MLP) is a counter-based mask: elements
// 4i..4i+3 of the tensor take 16 bits each of splitmix64(i, seed), so forward and backward regenerate the same mask from
// the seed and no mask tensor exists.
__device__ __forceinline__ uint64_t drop_bits(uint64_t i, uint64_t seed) {
  uint64_t z = i + seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// keep-or-zero multipliers of the four elements of vector i: keep when the 16-bit draw is >= thr = round(p * 65536)
__device__ __forceinline__ float4 drop_mult(uint64_t i, uint64_t seed, uint32_t thr, float inv_keep) {
  const uint64_t r = drop_bits(i, seed);
  float4 m;
  m.x = (uint32_t)(r & 0xffff) >= thr ? inv_keep : 0.f;
  m.y = (uint32_t)((r >> 16) & 0xffff) >= thr ? inv_keep : 0.f;
  m.z = (uint32_t)((r >> 32) & 0xffff) >= thr ? inv_keep : 0.f;
  m.w = (uint32_t)(r >> 48) >= thr ? inv_keep : 0.f;
  return m;
}

struct DropCfg {
  uint64_t seed;
  uint32_t thr;     // 0 = no dropout
  float inv_keep;
};

inline DropCfg make_drop(float p, uint64_t seed) {
  DropCfg dc;
  dc.seed = seed;
  dc.thr = p > 0.f ? (uint32_t)(p * 65536.f + 0.5f) : 0u;
  dc.inv_keep = p > 0.f ? 65536.f / (float)(65536u - dc.thr) : 1.f;   // exact keep probability of the 16-bit draw
  return dc;
}

__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t i, float4 v) {
  uint32_t h[2], l[2];
  split_bf16_pair(v.x, v.y, h[0], l[0]);
  split_bf16_pair(v.z, v.w, h[1], l[1]);
  *reinterpret_cast<uint2*>(hi + i) = make_uint2(h[0], h[1]);
  *reinterpret_cast<uint2*>(lo + i) = make_uint2(l[0], l[1]);
}

}  // namespace regnet
