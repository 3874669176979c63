// Shared device/host helpers for libregnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/regnet_b200.h"

namespace regnet {

// ---- error plumbing ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define RN_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::regnet::set_error(__VA_ARGS__);         \
      return REGNET_EINVAL;                     \
    }                                           \
  } while (0)

#define RN_CUDA(call)                                            \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return ::regnet::cuda_fail(e__, #call); \
  } while (0)

#define RN_LAUNCH_CHECK(name)                                    \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::regnet::cuda_fail(e__, name); \
  } while (0)

#define RN_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != REGNET_OK) return rc__; \
  } while (0)

// ---- the reference's squared distance, rounding pinned ------------------------------------------------------
// nvcc -O2 compiles (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1) (sampling_kernel.cu:82,
// ball_query_kernel.cu:60, interpolate_kernel.cu:56) to FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)
// (oracle/_ref/pn2_ext_ref.sass.txt).  Intrinsics keep that order whatever this translation unit's flags are.
__device__ __forceinline__ float sqdist_ref(float x1, float y1, float z1, float x2, float y2, float z2) {
  const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
  float t = __fmul_rn(dy, dy);
  t = __fmaf_rn(dx, dx, t);
  t = __fmaf_rn(dz, dz, t);
  return t;
}

// ---- split-bf16 operand format --------------------------------------------------------------------------------
// x ~= hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi): 16 significant bits, fp32 range.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(__fsub_rn(x, __bfloat162float(hi)));
}

// Two values at once -> one packed bf16x2 hi word and one lo word (element 0 in the low half).  Same roundings as
// split_bf16, but both conversions are the packed F2FP form: the scalar F2F.BF16.F32 the compiler emits for the `lo` of
// split_bf16 issues at a quarter of the rate and was a third of the tensor kernels' epilogue time.
__device__ __forceinline__ void split_bf16_pair(float y0, float y1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);   // .x (low half) = y0
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(__fsub_rn(y0, h0), __fsub_rn(y1, h1));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ReLU folded into the split: hi = bf16_rz(max(y, 0)) -- rounding toward zero keeps y - hi >= 0 for y >= 0, so the second
// relu-conversion lo = bf16_rn(max(y - hi, 0)) is exact for y >= 0 and gives hi = lo = 0 for y < 0.  No FMNMX; hi + lo still
// carries 16 significant bits (residual <= 2^-17 |y|).
__device__ __forceinline__ void relu_split_bf16_pair(float y0, float y1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y1), "f"(y0));   // d = {hi half: a, lo half: b}
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(__fsub_rn(y1, h1)), "f"(__fsub_rn(y0, h0)));
}

__host__ __device__ __forceinline__ int64_t round_up64(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
__host__ __device__ __forceinline__ int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Every kernel of the forward asks for the MAXIMUM shared-memory carve-out, whatever it uses itself: the L1 / shared
// split of an SM can only change while the SM is empty, the register-resident FPS kernel of the NEXT step becomes
// resident at an arbitrary moment of this step and then stays for milliseconds -- if it lands on SMs that a
// no-shared-memory kernel has just configured for a big L1, the tcgen05 kernels (217-225 KB of shared memory) cannot
// join those SMs until it leaves, and run on the ~28 SMs without an FPS CTA (timeline: gemm.sa1.l1 0.40 -> 1.51 ms).
#define RN_PREFER_MAX_SMEM(kernel)                                                                              \
  do {                                                                                                          \
    static bool done__ = false;                                                                                 \
    if (!done__) {                                                                                              \
      (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,                        \
                                 (int)cudaSharedmemCarveoutMaxShared);                                          \
      done__ = true;                                                                                            \
    }                                                                                                           \
  } while (0)

struct Strides3 {
  int64_t b, c, n;
};

}  // namespace regnet
