// Micro-test: cp.async.bulk.tensor.2d tile::gather4 (four arbitrary rows of a 2-D bf16 tensor into a SWIZZLE_128B shared
// memory tile) -- which tensor-map box shape it wants and where the rows land.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4 tma_gather4.cu -lcuda && ./tma_gather4
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap map, const int* rows, int nrows, int col0, uint16_t* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t base = (smem_u32(smem) + 1023) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 32 * 128 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem + (base - smem_u32(smem)))[i] = 0xdead;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nrows * 128) : "memory");
    for (int g = 0; g < nrows / 4; ++g)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(base + g * 512), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(col0), "r"(rows[4 * g]),
                     "r"(rows[4 * g + 1]), "r"(rows[4 * g + 2]), "r"(rows[4 * g + 3]) : "memory");
  }
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("timeout waiting for the gather\n"); break; }
  }
  __syncthreads();
  // de-swizzle: row r at r*128, 16-byte chunk c at (c ^ (r & 7))
  for (int i = threadIdx.x; i < nrows * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const uint32_t off = r * 128 + (((c / 8) ^ (r & 7)) << 4) + (c % 8) * 2;
    out[i] = *reinterpret_cast<uint16_t*>(smem + (base - smem_u32(smem)) + off);
  }
}

int main() {
  const int R = 4096, C = 256;
  uint16_t* h = new uint16_t[R * C];
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)((r * 7 + c) & 0xffff);
  uint16_t *d, *dout;
  cudaMalloc(&d, R * C * 2);
  cudaMemcpy(d, h, R * C * 2, cudaMemcpyHostToDevice);
  const int nrows = 16;
  int hrows[nrows] = {5, 1000, 3, 77, 4095, 0, 2048, 9, 100, 101, 102, 103, 3000, 17, 2999, 64};
  int* drows;
  cudaMalloc(&drows, sizeof(hrows));
  cudaMemcpy(drows, hrows, sizeof(hrows), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, nrows * 64 * 2);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Fn enc = (Fn)fp;
  for (int boxrows = 1; boxrows <= 4; boxrows *= 4) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box rows %d: encode rc=%d\n", boxrows, (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    const int col0 = 64;
    k<<<1, 128, 16384>>>(map, drows, nrows, col0, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    uint16_t ho[nrows * 64];
    cudaMemcpy(ho, dout, sizeof(ho), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < nrows; ++r2)
      for (int c = 0; c < 64; ++c)
        if (ho[r2 * 64 + c] != h[hrows[r2] * C + col0 + c]) { if (bad < 4) printf("  mismatch row %d col %d: got %u want %u\n", r2, c, ho[r2 * 64 + c], h[hrows[r2] * C + col0 + c]); ++bad; }
    printf("  box rows %d: %d mismatches of %d\n", boxrows, bad, nrows * 64);
  }
  return 0;
}
