// Microbenchmark: sustained tcgen05.mma rate for the Gram kernels' shapes (development tool).
// One CTA per SM; one elected thread issues `iters` x 4 MMAs on fixed smem operands (zeros) into 1 or 2
// TMEM accumulators, commits, waits, and reports cycles per MMA.
#include <cstdio>
#include <cstdlib>
#include "../attacking_federate_learning_b200/csrc/afl_common.cuh"
using namespace afl;

__device__ __forceinline__ void umma_f16k(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_bench(int n, int bf16, int naccum, int iters, int kadv, long long* out, int commit_every) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t ring[8];
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&ring[i], 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (threadIdx.x < 32) {
    const uint32_t idesc = bf16 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t(n) >> 3) << 17) | (8u << 24))
                                : umma_idesc_tf32(128, n);
    const uint64_t da = umma_desc_sw128(smem_u32(smem));
    const uint64_t db = umma_desc_sw128(smem_u32(smem) + 128 * 128);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t d = tb + ((naccum == 2 && (ks & 1)) ? 256 : 0);
          const uint64_t adv = kadv ? uint64_t(ks * 2) : 0;
          if (bf16) umma_f16k(d, da + adv, db + adv, idesc, 1);
          else umma_tf32(d, da + adv, db + adv, idesc, 1);
        }
        if (commit_every && (it % commit_every) == commit_every - 1) umma_commit(&ring[it & 7]);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1; }
    if (elect_one() && blockIdx.x == 0) out[1] = t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

namespace afl { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*, const char*, int) { return 3; }
void count_launch(int) {} int sm_count() { return 148; }
ProfScope::ProfScope(const char*, cudaStream_t) {} ProfScope::~ProfScope() {} }

int main() {
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  const int iters = 2000;
  int cfgs[][5] = {{224,0,1,1,0},{224,0,1,1,1},{224,0,1,1,4},{224,1,1,1,1},{112,0,1,1,1},{256,0,1,1,1}};
  for (auto& c : cfgs) {
    for (int grid : {148}) {
      mma_bench<<<grid, 128, 52 * 1024>>>(c[0], c[1], c[2], iters, c[3], out, c[4]);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      const double cyc = double(h[0] - h[1]) / (iters * 4.0);
      const double macs = 128.0 * c[0] * (c[1] ? 16 : 8);
      printf("N=%3d %s accum=%d commit_every=%d grid=%3d : %7.1f cycles/MMA  %7.0f MAC/clk/SM  (%s)\n", c[0], c[1] ? "bf16 K16" : "tf32 K8 ",
             c[2], c[4], grid, cyc, macs / cyc, cudaGetErrorString(e));
    }
  }
  return 0;
}
