// Micro-benchmark: tcgen05.ld / tcgen05.st cost per instruction as a function of width (x16/x32/x64 columns), number of
// warps reading concurrently, and back-to-back vs. waited issue.  One CTA per SM; reports cycles per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W> __device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink);
template <> __device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t& sink) {
  uint32_t v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) sink ^= v[i];
}
template <> __device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t& sink) {
  uint32_t v[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
                 "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
                 "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) sink ^= v[i];
}
// two x16 loads in flight, one wait
__device__ __forceinline__ void ld2x16(uint32_t taddr, uint32_t& sink) {
  uint32_t v[16], w[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
                 "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "r"(taddr + 16) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) sink ^= v[i] ^ w[i];
}
__device__ __forceinline__ void st16(uint32_t taddr, uint32_t x) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(taddr), "r"(x) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint64_t sdesc(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(1u) : "memory");
}
template <> __device__ __forceinline__ void ld<64>(uint32_t taddr, uint32_t& sink) {
  uint32_t v[64];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
                 "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
                 "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]),
                 "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]),
                 "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]),
                 "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]),
                 "=r"(v[63]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 64; ++i) sink ^= v[i];
}
// mode 0: ld x16 waited; 1: ld x32 waited; 2: two x16 then one wait; 3: st x16 waited; 4: 16 x redux.sync.max (no TMEM)
__global__ void bench(int mode, int iters, long long* out, int mma, volatile int* stop) {
  extern __shared__ __align__(1024) unsigned char opnd[];
  __shared__ uint32_t holder;
  __shared__ unsigned long long bar;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = holder + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t sink = threadIdx.x;
  const int nw = blockDim.x / 32 - (mma ? 1 : 0);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  __syncthreads();
  if (mma && warp == nw) {
    // background tensor load: M128 N256 K16 bf16 MMAs (SS) into columns [256,512), until the readers are done
    if (threadIdx.x % 32 == 0) {
      const uint32_t a0 = (smem_u32(opnd) + 1023) & ~1023u;
      const uint64_t da = sdesc(a0), db = sdesc(a0 + 16384);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
      long long n = 0;
      uint32_t ph = 0;
      while (*stop == 0 || n < 64) {
        if (mma == 2) { for (int k = 0; k < 16; ++k) umma_ts(holder + 256, holder + 128 + 8 * (k & 7), db + 2 * (k & 3), idesc); }
        else { for (int k = 0; k < 16; ++k) umma(holder + 256, da + 2 * (k & 3), db + 2 * (k & 3), idesc); }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(ph) : "memory");
        ph ^= 1;
        n += 16;
        if (n > 40000000) break;
      }
      out[blockIdx.x * 32 + 31] = n;
    }
  } else {
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)((i * 32 + (warp >> 2) * 64) & 127);   // columns [0,128+w): not the accumulator
    if (mode == 0) ld<16>(base + col, sink);
    else if (mode == 1) ld<32>(base + col, sink);
    else if (mode == 2) ld2x16(base + col, sink);
    else if (mode == 3) st16(base + col, sink);
    else if (mode == 5) ld<64>(base + (col & 63), sink);
    else {
#pragma unroll
      for (int j = 0; j < 16; ++j) sink += __reduce_max_sync(0xffffffffu, (int)(sink + j));
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0) out[blockIdx.x * 32 + warp] = t1 - t0 + (sink == 0x12345 ? 1 : 0);
  if (mma) {   // readers done: tell the MMA lane of this CTA
    asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) *stop = 1;
  }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(holder), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 32 * sizeof(long long));
  long long h[32];
  const char* names[6] = {"ld.x16+wait", "ld.x32+wait", "2 x ld.x16, 1 wait", "st.x16+wait", "16 x redux.max", "ld.x64+wait"};
  const int iters = 2000;
  int* stop;
  cudaMalloc(&stop, 4);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  for (int mma = 1; mma < 3; ++mma)
  for (int mode = 0; mode < 6; ++mode)
    for (int nw = 1; nw <= 16; nw *= 2) {
      if (nw == 16) nw = 12;
      if (mma == 0 && mode != 5) { if (nw == 12) break; continue; }
      cudaMemset(stop, 0, 4);
      bench<<<148, (nw + (mma ? 1 : 0)) * 32, 66 * 1024>>>(mode, iters, d, mma, stop);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("%-20s mma=%d warps=%2d  %.1f cycles / iteration (slowest warp)", names[mode], mma, nw, (double)mx / iters);
      if (mma) printf("   [%lld MMAs issued meanwhile = %.1f cycles each]", h[31], (double)mx / (double)h[31]);
      printf("\n");
      if (nw == 12) break;
    }
  return 0;
}
