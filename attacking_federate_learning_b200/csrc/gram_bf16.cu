// Gram kernel for the streaming regime (64 <= N_pad <= 112 clients, large D): the kernel that serves
// Krum at "N = 100, D = 11.2 M" (reference: defences.py:16-21).
//
// Why a second kernel: a tcgen05.mma microbenchmark (tools/mma_bench.cu) gives t = 43 + N/2 cycles for
// M = 128, for BOTH kind::tf32 (K = 8) and kind::f16 (K = 16).  At N = 2*112 the split-TF32 kernel of
// gram.cu needs >= 620 tensor cycles per 32 columns, more than the ~550 cycles HBM needs for them, while
// bf16 operands need 310.  Here:
//
//   * TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B, zero OOB fill) streams fp32 tiles [N_pad x 32] into
//     a deep ring (8 tiles in flight per SM) — deep asynchronous prefetch without registers;
//   * 4 converter warps turn two fp32 tiles into one bf16 operand stage: every value is split into two
//     bf16 terms, g = b1 + b2 + r (both roundings to nearest, |r| <= 2^-17 |g|), stored as b1 || b2,
//     K-major, 64 columns per 128-byte row, SWIZZLE_128B.  A 32-byte piece of the fp32 tile maps to the
//     16-byte chunk at HALF its byte offset (the two swizzles cancel), so no address math is needed
//     beyond swapping the two halves on odd rows;
//   * S ~= b1 b1^T + b1 b2^T + (b1 b2^T)^T: ONE tcgen05.mma.kind::f16 (K = 16) per 16 columns with
//     A = b1 (M = 128) and B = b1 || b2 (N = 2*N_pad), fp32 accumulation in TMEM.  Dropped terms
//     (b2 b2^T, r) are ~2^-17 relative per product; b2 b2^T is positive on every squared distance and
//     shows up as a uniform ~-1.3e-6 scale of d2;
//   * two MMA issuer warps alternate k-blocks, TMEM accumulators are double-buffered and drained every
//     `flush` k-blocks by 8 epilogue warps into fp32 registers, K is interleaved over the CTAs into
//     private partial slots that gram_reduce_kernel sums in a fixed order in float64 (bit-reproducible;
//     identical rows give exact zeros and identical table rows).
//
// Warp roles (512 threads): warp 0 TMA producer, warps 1,3 MMA issue, warp 2 TMEM alloc, warps 4-7
// converters, warps 8-15 epilogue.
#include "afl_common.cuh"

namespace afl {
namespace gram {

constexpr int kB16Threads = 512;
constexpr int kRawTiles = 8;         // fp32 tiles [nb x 32 cols] in flight (TMA ring)
constexpr int kBfStages = 4;         // bf16 operand stages [2*nb rows x 64 cols]
constexpr int kB16Cols = 64;         // fp32 columns per k-block = 128 bytes of bf16 per row
constexpr int kB16PartElems = 2 * 128 * 128;

struct B16Params {
  int n, nb;            // clients, padded to a multiple of 16 (64..112)
  int splits;           // CTAs; CTA c owns k-blocks c, c+splits, ...
  int kblocks;          // ceil(d / 64)
  int flush;            // k-blocks per TMEM accumulation chain (even)
  float* parts;         // [splits][2][128][128]
  int kc_log2;          // consecutive k-blocks per CTA visit (contiguous 256 B << kc_log2 per row)
};

__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void sts128_u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// c_format F32 (1) [4,6) | a_format BF16 (1) [7,10) | b_format BF16 (1) [10,13) | K-major | n>>3 | m>>4
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__global__ void __launch_bounds__(kB16Threads, 1)
gram_bf16x2_kernel(const __grid_constant__ CUtensorMap tmap, const B16Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t raw_full[kRawTiles], raw_empty[kRawTiles], bf_empty[kBfStages], acc_full[2],
      acc_empty[2], first_issued[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wg = warp >> 2;
  const int split = blockIdx.x;
  // K assignment: chunks of (1 << kc_log2) consecutive k-blocks; CTA `split` owns chunks split, split+splits, ...
  const int kc = 1 << p.kc_log2;
  const int nchunks_total = (p.kblocks + kc - 1) >> p.kc_log2;
  const int my_chunks = split < nchunks_total ? (nchunks_total - split + p.splits - 1) / p.splits : 0;
  int nkb = my_chunks * kc;
  if (my_chunks > 0 && split + (my_chunks - 1) * p.splits == nchunks_total - 1) nkb -= nchunks_total * kc - p.kblocks;
  const int ngroups = (nkb + p.flush - 1) / p.flush;
  const uint32_t raw_bytes = static_cast<uint32_t>(p.nb) * 128u;      // one fp32 tile [nb x 32]
  const uint32_t stage_bytes = static_cast<uint32_t>(p.nb) * 256u;    // b1 (nb rows) || b2 (nb rows)
  const uint32_t raw_base = smem_u32(smem);
  const uint32_t bf_base = raw_base + kRawTiles * raw_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRawTiles; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 1); }
    for (int s = 0; s < kBfStages; ++s) mbar_init(&bf_empty[s], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 2);
      mbar_init(&acc_empty[b], 8);
      mbar_init(&first_issued[b], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (wg == 0) {
    setmaxnreg_dec<64>();
    if (warp == 0) {
      // ===================== TMA producer: 2 fp32 tiles (32 columns each) per k-block =====================
      if (lane == 0) {
        const uint64_t pol = policy_evict_first();
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < 2 * nkb; ++it) {
          mbar_wait(&raw_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&raw_full[s], raw_bytes);
          const int i = it >> 1;                            // k-block index in this CTA's stream
          const int kb = ((split + (i >> p.kc_log2) * p.splits) << p.kc_log2) + (i & (kc - 1));
          tma_load_2d(smem + static_cast<size_t>(s) * raw_bytes, &tmap, &raw_full[s], kb * kB16Cols + (it & 1) * 32, 0, pol);
          if (++s == kRawTiles) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1 || warp == 3) {
      // ===================== MMA issuers (alternating k-blocks) =====================
      const int j = (warp == 3) ? 1 : 0;
      const uint32_t idesc = umma_idesc_bf16(128, 2 * p.nb);
      int s = j;
      for (int g = 0; g < ngroups; ++g) {
        const int b = g & 1;
        const uint32_t gph = (g >> 1) & 1;
        const int it_begin = g * p.flush, it_end = min(it_begin + p.flush, nkb);
        const uint32_t d_acc = tmem_base + static_cast<uint32_t>(b * 256);
        int it = it_begin + j;
        if (it < it_end) {
          mbar_wait_fast(&acc_empty[b], gph ^ 1);
          if (j == 1) mbar_wait_fast(&first_issued[b], gph);
          tc_fence_after();
        }
        for (; it < it_end; it += 2) {
          named_bar_sync(1 + s, 32 + 32);                // converter warp s has written and fenced stage s
          tc_fence_after();
          const uint64_t dab = umma_desc_sw128(bf_base + static_cast<uint32_t>(s) * stage_bytes);
          if (elect_one()) {                             // A = rows 0..127 of the stage, B = rows 0..2nb-1
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)               // 4 x K=16 bf16 = 64 columns; +32 bytes per step
              umma_bf16(d_acc, dab + static_cast<uint64_t>(ks * 2), dab + static_cast<uint64_t>(ks * 2), idesc,
                        (it != it_begin) || (ks != 0));
            umma_commit(&bf_empty[s]);
            if (it == it_begin) mbar_arrive(&first_issued[b]);
          }
          __syncwarp();
          s += 2;
          if (s >= kBfStages) s -= kBfStages;
        }
        if (elect_one()) umma_commit(&acc_full[b]);
        __syncwarp();
      }
    }
  } else if (wg == 1) {
    setmaxnreg_dec<112>();
    // ===================== converters: two fp32 tiles -> one bf16 stage b1 || b2 =====================
    // Piece q = 32 aligned bytes of an fp32 tile = 8 values of row q/4.  TMA's 128-byte swizzle stored the
    // two 16-byte chunks of the piece swapped on odd rows; the bf16 chunk belongs at byte q*16 of the
    // 64-column half-row, i.e. at (row*128 + half*64 + (q%4)*16) XOR-swizzled by (row & 7).
    // One converter warp per k-block (warp w owns k-blocks w, w+4, ... and bf16 stage w): four k-blocks are
    // converted in parallel and a warp pays its barrier latencies once per k-block of its own.
    const int w4 = warp - 4;
    const int npieces = p.nb * 4;                         // per fp32 tile
    const uint32_t dst = bf_base + static_cast<uint32_t>(w4) * stage_bytes;
    for (int kb = w4; kb < nkb; kb += kBfStages) {
      mbar_wait_fast(&bf_empty[w4], ((kb / kBfStages) & 1) ^ 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int tile = 2 * kb + half;                   // position in the TMA stream
        const int rs = tile % kRawTiles;
        mbar_wait_fast(&raw_full[rs], (tile / kRawTiles) & 1);
        const uint32_t src = raw_base + static_cast<uint32_t>(rs) * raw_bytes;
#pragma unroll 1
        for (int q0 = 0; q0 < npieces; q0 += 4 * 32) {
          float4 v0[4], v1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {                   // all loads first (8 LDS.128 in flight)
            const int q = q0 + u * 32 + lane;
            if (q < npieces) { v0[u] = lds128(src + q * 32); v1[u] = lds128(src + q * 32 + 16); }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int q = q0 + u * 32 + lane;
            if (q < npieces) {
              const int row = q >> 2;
              const bool odd = row & 1;
              const float4 a = odd ? v1[u] : v0[u], b = odd ? v0[u] : v1[u];   // logical column order
              const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
              uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h[e] = pack_bf16x2_rn(x[2 * e], x[2 * e + 1]);
                const float r0 = x[2 * e] - __uint_as_float(h[e] << 16);
                const float r1 = x[2 * e + 1] - __uint_as_float(h[e] & 0xFFFF0000u);
                l[e] = pack_bf16x2_rn(r0, r1);
              }
              // logical 16-byte chunk of the bf16 row: c = half*4 + (stored pair index ^ ((row>>1)&3))
              const int c = half * 4 + ((q & 3) ^ ((row >> 1) & 3));
              const uint32_t off = static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(c ^ (row & 7)) << 4);
              sts128_u(dst + off, h[0], h[1], h[2], h[3]);
              sts128_u(dst + static_cast<uint32_t>(p.nb) * 128u + off, l[0], l[1], l[2], l[3]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);       // done reading this fp32 tile
      }
      fence_proxy_async_smem();
      named_bar_arrive(1 + w4, 32 + 32);
    }
  } else {
    setmaxnreg_inc<168>();
    // ===================== epilogue: drain TMEM chains into fp32 registers =====================
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int a = (warp - 8) >> 2;      // 0: b1*b1^T columns, 1: b1*b2^T columns
    float run[112];
#pragma unroll
    for (int i = 0; i < 112; ++i) run[i] = 0.f;
    for (int g = 0; g < ngroups; ++g) {
      const int b = g & 1;
      mbar_wait_fast(&acc_full[b], (g >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(b * 256 + a * p.nb);
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        if (c * 16 < p.nb) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) run[c * 16 + i] += __uint_as_float(v[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
    float* out = p.parts + static_cast<size_t>(split) * kB16PartElems + static_cast<size_t>(a) * 128 * 128 +
                 static_cast<size_t>(q * 32 + lane) * 128;
#pragma unroll
    for (int c = 0; c < 112; c += 4)
      if (c < p.nb) *reinterpret_cast<float4*>(out + c) = make_float4(run[c], run[c + 1], run[c + 2], run[c + 3]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool bf16x2_eligible(int n, int64_t d) {
  const int nb = (n + 15) & ~15;
  return nb >= 64 && nb <= 112 && d >= 32768;
}

int bf16x2_splits(int64_t d) {
  const int64_t kblocks = (d + kB16Cols - 1) / kB16Cols;
  int s = sm_count();
  if (const char* e = getenv("AFL_GRAM_SPLITS")) if (atoi(e) > 0) s = atoi(e);
  if (s > kblocks) s = static_cast<int>(kblocks);
  return s < 1 ? 1 : s;
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// parts must hold bf16x2_splits(d) * 2*128*128 floats (gram_reduce_kernel only reads rows/columns < n).
int launch_bf16x2(const float* G, int n, int64_t d, int64_t ld, float* parts, int splits, int flush,
                  cudaStream_t stream) {
  static EncodeTiledFn2 enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled entry point not found"); return AFL_ERR_CUDA; }
    enc = reinterpret_cast<EncodeTiledFn2>(fp);
  }
  B16Params p{};
  p.n = n; p.nb = (n + 15) & ~15;
  p.kblocks = static_cast<int>((d + kB16Cols - 1) / kB16Cols);
  p.flush = flush < 2 ? 2 : (flush & ~1);
  p.splits = splits;
  p.parts = parts;
  p.kc_log2 = 0;
  if (const char* e = getenv("AFL_GRAM_KCHUNK_LOG2")) { p.kc_log2 = atoi(e); if (p.kc_log2 < 0 || p.kc_log2 > 4) p.kc_log2 = 0; }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(n)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
  const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(p.nb)};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(G), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", static_cast<int>(r)); return AFL_ERR_CUDA; }
  const size_t smem = static_cast<size_t>(kRawTiles) * p.nb * 128 + static_cast<size_t>(kBfStages) * p.nb * 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    AFL_CUDA(cudaFuncSetAttribute(gram_bf16x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  {
    ProfScope ps("gram_tcgen05", stream);
    gram_bf16x2_kernel<<<splits, kB16Threads, smem, stream>>>(tmap, p);
  }
  AFL_LAUNCH_CHECK("gram_bf16x2_kernel");
  return AFL_OK;
}

}  // namespace gram
}  // namespace afl
