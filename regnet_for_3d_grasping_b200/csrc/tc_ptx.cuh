// Inline-PTX wrappers shared by the tcgen05 kernels (gemm_tc.cu, sa0_front.cu): mbarrier, TMA, TMEM, UMMA.
// Instruction strings follow the CUTLASS/CuTe sm_100 headers (cute/arch/copy_sm90_tma.hpp, mma_sm100_umma.hpp,
// tmem_allocator_sm100.hpp, cutlass/arch/barrier.h); descriptor layouts follow cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace regnet {
namespace tc {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error surfaced to the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();  // ~4 s at 2 GHz
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// tile::gather4: four arbitrary rows (r0..r3) of a 2-D tensor, columns [c0, c0 + box_cols), land as four consecutive rows
// of the shared-memory tile at dst (swizzled like any other row of that tile).  The tensor map's box must be
// {box_cols, 1 row}; a {box_cols, 4} box raises "illegal instruction" (scripts/ubench/tma_gather4.cu).
__device__ __forceinline__ void tma_gather4_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int r0, int r1,
                                               int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// explicit shared-space accesses (a pointer derived from the manually aligned dynamic-smem base is generic to nvcc,
// which then emits LD.E / ST.E with 64-bit addressing instead of LDS / STS)
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load without the wait: issue, do other work, then tmem_ld_wait() before touching v
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address >> 4 in [0,14), LBO (ignored for swizzled K-major) = 1 in [16,30), SBO = 1024 B >> 4 in [32,46)
// (8 rows x 128 B per swizzle atom), version = 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major, SWIZZLE_128B operand (cute UMMA canonical layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)), T = 8 bf16): the MN
// extent is contiguous in memory, 64 elements (128 bytes) per row; 8 consecutive k rows form one 1024-byte swizzle
// atom (SBO = distance between 8-row k groups), 64-element MN chunks are LBO bytes apart.  That is exactly what TMA
// writes for a {64 (MN, contiguous), kb rows} box with CU_TENSOR_MAP_SWIZZLE_128B: SBO = 1024, LBO = kb * 128.
// One K = 16 instruction consumes two k groups, so the k-step advance of the start address is 2048 bytes.
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major operand with 32-byte rows (K = 16 bf16): the canonical SWIZZLE_32B layout ((8,n),2):((2,SBO),1) in 16-byte
// units -- row r at r * 32 bytes, 16-byte chunk index XOR bit 2 of the row, 8-row groups SBO = 256 bytes apart.
// (The un-swizzled "interleave" layout with LBO = 128 B was verified to work as well during bring-up.)
__device__ __forceinline__ uint32_t k16_offset(uint32_t row, uint32_t chunk) {
  return row * 32u + ((chunk ^ ((row >> 2) & 1u)) << 4);
}
__device__ __forceinline__ uint64_t make_sdesc_k16(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;             // LBO (unused: one swizzle span per row)
  d |= (uint64_t)(256 >> 4) << 32;    // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;             // SWIZZLE_32B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A,B K-major (bits 15,16 = 0), N>>3 in [17,23), M>>4 in [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
constexpr uint32_t IDESC_A_MN_MAJOR = 1u << 15;   // a_major_ / b_major_ bits: 0 = K-major, 1 = MN-major
constexpr uint32_t IDESC_B_MN_MAJOR = 1u << 16;

// ---- 64-row max-pool of an accumulator tile -------------------------------------------------------------------
// tcgen05.ld.16x256b.x4: the warp reads 16 TMEM lanes x 32 columns; thread t receives, for each 8-column group i,
// v[4i+0..1] = row t/4, columns 8i + 2(t%4) + {0,1} and v[4i+2..3] = row t/4 + 8, same columns (cute
// SM100_TMEM_LOAD_16dp256b4x: DstLayout ((4,8),(64,2,4)):((64,1024),(1,8192,256)) bits).  Two such loads (lane
// offsets 0 and 16 of the warp's quarter) give every thread FOUR rows of 8 columns, so two of the five levels of the
// 32-row max are plain in-thread FMNMX and only three shuffle levels remain (7 SHFL per 32 columns).
// redux.sync.max is NOT used: measured on B200 it retires one warp instruction per ~30 cycles per scheduler
// (scripts/ubench/tmem_ld.cu), ~7 700 cycles for a 128 x 256 tile -- more than the tile's MMA time.
__device__ __forceinline__ void tmem_ld_16x256b_x4_async(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// in-thread part: a = rows t/4, t/4+8 (lanes 0..15 of the quarter), b = rows t/4+16, t/4+24 -> m[2i+e] = max over
// the four rows of column 8i + 2(t%4) + e
__device__ __forceinline__ void colmax_rows4(const uint32_t (&a)[16], const uint32_t (&b)[16], float (&m)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e)
      m[2 * i + e] = fmaxf(fmaxf(__uint_as_float(a[4 * i + e]), __uint_as_float(a[4 * i + 2 + e])),
                           fmaxf(__uint_as_float(b[4 * i + e]), __uint_as_float(b[4 * i + 2 + e])));
}
// cross-lane part: butterfly over lane bits 4, 3, 2 (the 8 lanes that hold the same columns), halving the live values
// at every step.  Lane t ends up with the 32-row maximum of column colmax_column(t) of the 32-column chunk.
__device__ __forceinline__ float colmax_lanes8(const float (&m)[8], int lane) {
  float w4[4];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float keep = up ? m[4 + j] : m[j], send = up ? m[j] : m[4 + j];
      w4[j] = fmaxf(keep, __shfl_xor_sync(FULL, send, 16));
    }
  }
  float w2[2];
  {
    const bool up = lane & 8;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float keep = up ? w4[2 + j] : w4[j], send = up ? w4[j] : w4[2 + j];
      w2[j] = fmaxf(keep, __shfl_xor_sync(FULL, send, 8));
    }
  }
  const bool up = lane & 4;
  const float keep = up ? w2[1] : w2[0], send = up ? w2[0] : w2[1];
  return fmaxf(keep, __shfl_xor_sync(FULL, send, 4));
}
// value index kept by lane t: 4 b4 + 2 b3 + b2 = 2i + e  ->  column 8i + 2(t%4) + e = 16 b4 + 8 b3 + 2 (t & 3) + b2
__device__ __forceinline__ int colmax_column(int lane) { return (lane & 24) + 2 * (lane & 3) + ((lane >> 2) & 1); }

// The same for 16 columns at a time (.x2: 8 registers per load instead of 16 -- for kernels short of registers):
// v[4i+0..1] = row t/4, columns 8i + 2(t%4) + {0,1}, v[4i+2..3] = row t/4 + 8, i = 0, 1.
__device__ __forceinline__ void tmem_ld_16x256b_x2_async(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// a = rows t/4, t/4+8, b = rows t/4+16, t/4+24 of 16 columns: in-thread part, m[2i+e] = max over the four rows
__device__ __forceinline__ void colmax16_rows4(const uint32_t (&a)[8], const uint32_t (&b)[8], float (&m)[4]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e)
      m[2 * i + e] = fmaxf(fmaxf(__uint_as_float(a[4 * i + e]), __uint_as_float(a[4 * i + 2 + e])),
                           fmaxf(__uint_as_float(b[4 * i + e]), __uint_as_float(b[4 * i + 2 + e])));
}
// cross-lane part: lane t returns the 32-row maximum of column colmax16_column(t); lanes t and t^4 hold the same
// column (only 4 values per thread enter the 8-lane butterfly)
__device__ __forceinline__ float colmax16_lanes8(const float (&m)[4], int lane) {
  float w2[2];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float keep = up ? m[2 + j] : m[j], send = up ? m[j] : m[2 + j];
      w2[j] = fmaxf(keep, __shfl_xor_sync(FULL, send, 16));
    }
  }
  const bool up = lane & 8;
  const float keep = up ? w2[1] : w2[0], send = up ? w2[0] : w2[1];
  float r = fmaxf(keep, __shfl_xor_sync(FULL, send, 8));
  return fmaxf(r, __shfl_xor_sync(FULL, r, 4));
}
// index kept by lane t: i = b4, e = b3  ->  column 8 b4 + 2 (t & 3) + b3
__device__ __forceinline__ int colmax16_column(int lane) { return ((lane & 16) >> 1) + 2 * (lane & 3) + ((lane >> 3) & 1); }

__device__ __forceinline__ uint32_t order_bits(float v) {  // unsigned order == float order
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_bits(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}


}  // namespace tc
}  // namespace regnet
