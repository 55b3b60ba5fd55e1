// Microbenchmark: the L2 -> shared-memory speed of light of gram_pair_kernel's access pattern.
// A [1024 x d] fp32 matrix that FITS the L2 (d = 8192: 32 MB) is streamed the way the tile-pair kernel streams it: one
// CTA per SM owns a pair of 128-row tiles (ti, tj) and, per k-block of W columns, loads the two boxes {W cols x 128 rows}
// into a ring of `slots` slots; a consumer thread releases every slot as soon as it is full (no conversion, no MMA).
// The CTA walks the d columns `reps` times.  Prints the aggregate bytes/s delivered into shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/l2_tma_bench tools/l2_tma_bench.cu -lcuda
//   tools/l2_tma_bench <W cols: 64|128> <slots> [d=8192] [reps=64] [boxes per k-block: 2|1] [ctas per SM=1]
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
    if (++spins > (1l << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

constexpr int kMaxSlots = 16;

__global__ void __launch_bounds__(64) pair_stream_kernel(const __grid_constant__ CUtensorMap tmap, int W, int box_bytes, int slots,
                                                         int kblocks, int reps, int nbx) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full[kMaxSlots], empty[kMaxSlots];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // tile pair of this CTA among the 36 lower-triangular pairs of 8 row tiles (as gram_pair_kernel at N = 1000)
  const int pair = blockIdx.x % 36;
  int ti = 0;
  while ((ti + 1) * (ti + 2) / 2 <= pair) ++ti;
  const int tj = pair - ti * (ti + 1) / 2;
  const int split = blockIdx.x / 36, splits = (gridDim.x + 35) / 36;
  const int total = kblocks * reps * nbx;
  if (threadIdx.x == 0) {
    for (int b = 0; b < total; ++b) {
      const int s = b % slots, u = b / slots;
      if (u > 0) mbar_wait(&empty[s], (u - 1) & 1);
      mbar_expect(&full[s], box_bytes);
      const int kb = ((b / nbx) * splits + split) % kblocks;
      tma_load_2d(base + s * box_bytes, &tmap, &full[s], kb * W, ((b % nbx) == 0 ? ti : tj) * 128);
    }
  } else if (threadIdx.x == 32) {
    for (int b = 0; b < total; ++b) {
      const int s = b % slots, u = b / slots;
      mbar_wait(&full[s], u & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int W = argc > 1 ? atoi(argv[1]) : 64;
  int slots = argc > 2 ? atoi(argv[2]) : 7;
  const int64_t d = argc > 3 ? atoll(argv[3]) : 8192;
  const int reps = argc > 4 ? atoi(argv[4]) : 64;
  const int nbx = argc > 5 ? atoi(argv[5]) : 2;
  const int per_sm = argc > 6 ? atoi(argv[6]) : 1;
  const int rows = 1024;
  float* G; CK(cudaMalloc(&G, sizeof(float) * rows * d)); CK(cudaMemset(G, 0, sizeof(float) * rows * d));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d) * 4};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(W), 128u};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = reinterpret_cast<EncFn>(fp)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G, gdim, gstride, box, es,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int box_bytes = W * 128 * 4;
  if (slots > kMaxSlots) slots = kMaxSlots;
  const size_t smem = static_cast<size_t>(slots) * box_bytes + 1024;
  if (smem > 227 * 1024) { printf("W=%d slots=%d: %zu bytes of shared memory do not fit\n", W, slots, smem); return 1; }
  CK(cudaFuncSetAttribute(pair_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int kblocks = static_cast<int>(d / W);
  const int grid = 148 * per_sm;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(e0));
    pair_stream_kernel<<<grid, 64, smem>>>(tmap, W, box_bytes, slots, kblocks, reps, nbx);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double bytes = static_cast<double>(grid) * kblocks * reps * nbx * box_bytes;
  printf("l2_tma W=%d rows=128 box=%d KB slots=%d boxes/kblock=%d ctas/SM=%d matrix=%.0f MB reps=%d: %.3f ms  %.0f GB/s into smem, %.0f ns per box per CTA\n",
         W, box_bytes / 1024, slots, nbx, per_sm, rows * d * 4.0 / 1e6, reps, best, bytes / best * 1e-6,
         best * 1e6 / (static_cast<double>(kblocks) * reps * nbx));
  return 0;
}
