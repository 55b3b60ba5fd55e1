// Microbenchmark: how fast can one CTA per SM stream a [rows x D] fp32 matrix (row stride = D) into
// shared memory with TMA boxes of {W columns x R rows}?  Decides the box shape of the Gram loaders.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_bench tools/tma_bench.cu -lcuda
//   tools/tma_bench <W cols> <R rows> <slots> <ctas_per_sm> <swizzle128 0|1> [d] [batch]
// Prints GB/s.  `batch` = tiles the producer issues back to back before waiting again.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0; long spins = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (++spins > (1l << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

constexpr int kMaxSlots = 32;

__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ CUtensorMap tmap, int W, int tile_bytes, int slots,
                                                    int ntiles, int batch, int Z, int tall_cols) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full[kMaxSlots], empty[kMaxSlots];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int my = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  if (threadIdx.x == 0) {                       // producer
    for (int i = 0; i < my; ++i) {
      const int s = i % slots, u = i / slots;
      if (u > 0 && (i % batch) == 0) {          // wait for a whole batch of slots
        for (int j = 0; j < batch && i + j < my; ++j) mbar_wait(&empty[(i + j) % slots], (((i + j) / slots) - 1) & 1);
      }
      mbar_expect(&full[s], tile_bytes);
      const int t = blockIdx.x + i * gridDim.x;
      if (Z > 0) tma_load_3d(base + s * tile_bytes, &tmap, &full[s], 0, 0, t * Z);
      else if (tall_cols > 0) tma_load_2d(base + s * tile_bytes, &tmap, &full[s], (t % tall_cols) * W, (t / tall_cols) * 100);
      else tma_load_2d(base + s * tile_bytes, &tmap, &full[s], t * W, 0);
    }
  } else if (threadIdx.x == 32) {               // consumer: release immediately
    for (int i = 0; i < my; ++i) {
      const int s = i % slots, u = i / slots;
      mbar_wait(&full[s], u & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

// Reference point: plain vector loads, each warp streams 512 B per row.
__global__ void __launch_bounds__(512) ldg_kernel(const float4* __restrict__ G, int rows, int64_t ld4, int64_t ncol4, float* out) {
  float acc = 0.f;
  for (int64_t c = blockIdx.x * 512 + threadIdx.x; c < ncol4; c += static_cast<int64_t>(gridDim.x) * 512) {
#pragma unroll 4
    for (int r = 0; r < rows; ++r) { const float4 v = __ldcs(G + r * ld4 + c); acc += v.x + v.y + v.z + v.w; }
  }
  if (acc == 123.456f) out[0] = acc;
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int W = argc > 1 ? atoi(argv[1]) : 32, R = argc > 2 ? atoi(argv[2]) : 100;
  int slots = argc > 3 ? atoi(argv[3]) : 8;
  const int per_sm = argc > 4 ? atoi(argv[4]) : 1, sw = argc > 5 ? atoi(argv[5]) : 1;
  const int64_t d = argc > 6 ? atoll(argv[6]) : 11200000;
  const int batch = argc > 7 ? atoi(argv[7]) : 1;
  const int rows = 100;
  const int Z = argc > 9 ? atoi(argv[9]) : 0;    // >0: 3-D map {W, rows, d/W}, box {W, R, Z}
  const int tall = argc > 10 ? atoi(argv[10]) : 0;   // >0: view the buffer as [100*tall rows x d/tall cols]
  const int promo = argc > 11 ? atoi(argv[11]) : 2;  // 0 none, 1 128B, 2 256B
  float* G; CK(cudaMalloc(&G, sizeof(float) * rows * d)); CK(cudaMemset(G, 0, sizeof(float) * rows * d));
  float* out; CK(cudaMalloc(&out, 4));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  CUtensorMap tmap;
  const int64_t dt = tall > 0 ? d / tall : d;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dt), static_cast<cuuint64_t>(tall > 0 ? rows * tall : rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dt) * 4};
  const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(W), static_cast<cuuint32_t>(R)};
  const cuuint32_t es[2] = {1, 1};
  CUresult r;
  if (Z > 0) {
    const cuuint64_t gdim3[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(d / W)};
    const cuuint64_t gstride3[2] = {static_cast<cuuint64_t>(d) * 4, static_cast<cuuint64_t>(W) * 4};
    const cuuint32_t box3[3] = {static_cast<cuuint32_t>(W), static_cast<cuuint32_t>(R), static_cast<cuuint32_t>(Z)};
    const cuuint32_t es3[3] = {1, 1, 1};
    r = reinterpret_cast<EncFn>(fp)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, G, gdim3, gstride3, box3, es3,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                    pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else
  r = reinterpret_cast<EncFn>(fp)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G, gdim, gstride, box, es,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                           pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int tile_bytes = W * R * 4 * (Z > 0 ? Z : 1);
  if (slots > kMaxSlots) slots = kMaxSlots;
  const size_t smem = static_cast<size_t>(slots) * tile_bytes + 1024;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  const int ntiles = static_cast<int>(d / W / (Z > 0 ? Z : 1));
  const int grid = 148 * per_sm;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int it = 0; it < 6; ++it) {
    CK(cudaEventRecord(e0));
    stream_kernel<<<grid, 64, smem>>>(tmap, W, tile_bytes, slots, ntiles, batch, Z, tall > 0 ? static_cast<int>(dt / W) : 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double bytes = static_cast<double>(ntiles) * tile_bytes;
  printf("tall=%d promo=%d ", tall, promo);
  printf("tma W=%d R=%d Z=%d slots=%d per_sm=%d sw=%d batch=%d smem=%zu: %.4f ms  %.0f GB/s\n", W, R, Z, slots, per_sm, sw, batch, smem, best,
         bytes / best * 1e-6);
  if (argc > 8 && atoi(argv[8]) > 0) {
    best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      CK(cudaEventRecord(e0));
      ldg_kernel<<<148 * atoi(argv[8]), 512>>>(reinterpret_cast<const float4*>(G), rows, d / 4, d / 4, out);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it > 0 && ms < best) best = ms;
    }
    printf("ldg grid=148x%s: %.4f ms  %.0f GB/s\n", argv[8], best, 4.0 * rows * d / best * 1e-6);
  }
  return 0;
}
