// Gram kernel for the streaming regime (64 <= N_pad <= 112 clients, large D): the kernel that serves
// Krum at "N = 100, D = 11.2 M" (reference: defences.py:16-21).
//
// Measured on B200 (profiles/README.md): the TMA + split-TF32 kernel in gram.cu is limited by
// shared-memory traffic (TMA write + split read/write + tensor-core operand reads ~ 86 KB per 32
// columns at an effective ~70 B/clk/SM).  This kernel moves 2.4x fewer shared-memory bytes per column:
//
//   * 8 loader warps read fp32 rows STRAIGHT from global memory into registers (16-byte loads, 256
//     contiguous bytes per row and k-block, up to 16 loads per lane in flight) — no fp32 tile in smem;
//   * every value is split in registers into two bf16 terms, g = b1 + b2 + r, |r| <= 2^-17 |g| (both
//     roundings to nearest, so r has no preferred sign), and only the compact bf16 tiles b1 || b2 are
//     stored (K-major, 128-byte rows, SWIZZLE_128B — the layout tcgen05 reads);
//   * S ~= b1 b1^T + b1 b2^T + (b1 b2^T)^T: ONE tcgen05.mma.kind::f16 per 16 columns with A = b1
//     (M = 128) and B = b1 || b2 (N = 2*N_pad), fp32 accumulation in TMEM — half the tensor time and a
//     quarter of the MMA instructions per column of the TF32 kernel.  Dropped terms (b2 b2^T, r) are
//     ~2^-17 relative per product with random sign, i.e. ~1e-5/sqrt(D) on a distance.
//   * two MMA issuer warps alternate k-blocks (an issuing thread is blocked while its MMAs drain),
//     TMEM accumulators are double-buffered and drained every `flush` k-blocks into fp32 registers
//     (bounded tensor-core accumulation chains), K is split over the 148 CTAs into private partial
//     slots that gram_reduce_kernel sums in a fixed order in float64 (bit-reproducible; identical rows
//     give exact zeros and identical table rows).
//
// Warp roles (512 threads): warps 1,3 MMA issue; warp 2 TMEM alloc; warps 4-11 loaders (each owns one of
// the 8 smem stages); warps 12-15 epilogue (each thread owns one TMEM lane = one client row and keeps its
// 2*N_pad running sums in registers, hence setmaxnreg 248 for that warpgroup).
#include "afl_common.cuh"

namespace afl {
namespace gram {

constexpr int kB16Threads = 512;
constexpr int kB16Stages = 8;        // == number of loader warps
constexpr int kB16Cols = 64;         // fp32 columns per k-block = 128 bytes of bf16 per row
constexpr int kB16PartElems = 2 * 128 * 128;

struct B16Params {
  const float* G;
  int64_t ld, d;
  int n, nb;            // clients, padded to a multiple of 16 (64..112)
  int splits;           // CTAs; CTA c owns k-blocks c, c+splits, ...
  int kblocks;          // ceil(d / 64)
  int flush;            // k-blocks per TMEM accumulation chain (even)
  float* parts;         // [splits][2][128][128]
  int prefetch;         // L2 prefetch distance in this warp's own k-blocks (0 = off)
  int knock;            // debug knock-outs: 1 no MMA, 2 no smem stores, 4 no global loads
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

// Pull one k-block's worth of this warp's rows (256 contiguous bytes each) into L2, without registers.
__device__ __forceinline__ void prefetch_rows_l2(const B16Params& p, int64_t col0, int lane) {
  if (col0 >= p.d) return;
  const int64_t remain = p.d - col0;
  const uint32_t bytes = remain >= kB16Cols ? 256u : static_cast<uint32_t>(remain / 4) * 16u;
  if (bytes == 0) return;
  for (int r = lane; r < p.n; r += 32) {
    const float* src = p.G + static_cast<int64_t>(r) * p.ld + col0;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
  }
}

constexpr int kPassPerBatch = 3;     // 3 passes x 2 x LDG.128 = 6 loads (24 registers) per batch, triple-buffered:
                                     // two batches (12 loads per lane, ~49 KB per SM) are in flight while one is consumed

struct Batch { float4 v[kPassPerBatch][2]; };

// Issue the loads of one batch: pass q covers rows 4*(4*b+q) + sub; lane covers 8 columns (32 bytes).
__device__ __forceinline__ void load_batch(Batch& B, const B16Params& p, const float* gcol, int64_t col, int b, int sub) {
  const bool full = (col + 8 <= p.d);
#pragma unroll
  for (int q = 0; q < kPassPerBatch; ++q) {
    const int row = 4 * (kPassPerBatch * b + q) + sub;
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (row < p.n && !(p.knock & 4)) {
      const float* src = gcol + static_cast<int64_t>(row) * p.ld;
      if (full) {
        x0 = ldg_stream_f4(reinterpret_cast<const float4*>(src));
        x1 = ldg_stream_f4(reinterpret_cast<const float4*>(src) + 1);
      } else if (col < p.d) {                                     // ragged tail of the last k-block
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = (col + e < p.d) ? __ldg(src + e) : 0.f;
        x0 = make_float4(t[0], t[1], t[2], t[3]);
        x1 = make_float4(t[4], t[5], t[6], t[7]);
      }
    }
    B.v[q][0] = x0;
    B.v[q][1] = x1;
  }
}

// Split 8 fp32 into b1 (RN bf16) and b2 = RN bf16(x - b1) and store both 16-byte chunks.
__device__ __forceinline__ void store_batch(const Batch& B, const B16Params& p, uint32_t st, int b, int sub, int c32) {
#pragma unroll
  for (int q = 0; q < kPassPerBatch; ++q) {
    const int row = 4 * (kPassPerBatch * b + q) + sub;
    if (row < p.n && !(p.knock & 2)) {
      const float x[8] = {B.v[q][0].x, B.v[q][0].y, B.v[q][0].z, B.v[q][0].w,
                          B.v[q][1].x, B.v[q][1].y, B.v[q][1].z, B.v[q][1].w};
      uint32_t h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = pack_bf16x2_rn(x[2 * e], x[2 * e + 1]);
        const float r0 = x[2 * e] - __uint_as_float(h[e] << 16);
        const float r1 = x[2 * e + 1] - __uint_as_float(h[e] & 0xFFFF0000u);
        l[e] = pack_bf16x2_rn(r0, r1);
      }
      const uint32_t off = static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(c32 ^ (row & 7)) << 4);
      sts128_u(st + off, h[0], h[1], h[2], h[3]);
      sts128_u(st + static_cast<uint32_t>(p.nb) * 128u + off, l[0], l[1], l[2], l[3]);
    }
  }
}

__global__ void __launch_bounds__(kB16Threads, 1)
gram_bf16x2_kernel(const B16Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t empty_bar[kB16Stages], acc_full[2], acc_empty[2], first_issued[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wg = warp >> 2;
  const int split = blockIdx.x;
  const int nkb = (p.kblocks - split + p.splits - 1) / p.splits;      // k-blocks split, split+splits, ...
  const int ngroups = (nkb + p.flush - 1) / p.flush;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.nb) * 256u;    // b1 (nb rows) || b2 (nb rows)
  const uint32_t smem_base = smem_u32(smem);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kB16Stages; ++s) mbar_init(&empty_bar[s], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 2);
      mbar_init(&acc_empty[b], 4);
      mbar_init(&first_issued[b], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (wg == 0) {
    setmaxnreg_dec<40>();
    if (warp == 1 || warp == 3) {
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
          mbar_wait_warp(&acc_empty[b], gph ^ 1);
          if (j == 1) mbar_wait_warp(&first_issued[b], gph);
          tc_fence_after();
        }
        for (; it < it_end; it += 2) {
          named_bar_sync(1 + s, 64);                     // loader warp s has written and fenced its stage
          tc_fence_after();
          const uint32_t st = smem_base + static_cast<uint32_t>(s) * stage_bytes;
          const uint64_t dab = umma_desc_sw128(st);      // A = rows 0..127 of the stage, B = rows 0..2nb-1
          if (elect_one()) {
            if (!(p.knock & 1))
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)               // 4 x K=16 bf16 = 64 columns; +32 bytes per step
              umma_bf16(d_acc, dab + static_cast<uint64_t>(ks * 2), dab + static_cast<uint64_t>(ks * 2), idesc,
                        (it != it_begin) || (ks != 0));
            umma_commit(&empty_bar[s]);
            if (it == it_begin) mbar_arrive(&first_issued[b]);
          }
          __syncwarp();
          s += 2;
          if (s >= kB16Stages) s -= kB16Stages;
        }
        if (elect_one()) umma_commit(&acc_full[b]);
        __syncwarp();
      }
    }
  } else if (wg == 1 || wg == 2) {
    setmaxnreg_dec<104>();
    // ===================== loaders: global fp32 -> registers -> bf16 b1 || b2 tiles =====================
    const int w = warp - 4;                               // loader index == stage index
    const int sub = lane >> 3, c32 = lane & 7;
    const int passes = (p.n + 3) >> 2;
    const int nbatch = (passes + kPassPerBatch - 1) / kPassPerBatch;
    const int my_kb = (nkb > w) ? (nkb - w + kB16Stages - 1) / kB16Stages : 0;
    const int total = my_kb * nbatch;
    const uint32_t st = smem_base + static_cast<uint32_t>(w) * stage_bytes;
    // flattened (k-block, batch) stream; loads run one batch ahead of the stores across k-block borders
    int kb_l = 0, b_l = 0;                                // position of the batch being LOADED
    int kb_s = 0, b_s = 0;                                // position of the batch being STORED
    auto col_of = [&](int kbi) -> int64_t {
      return (static_cast<int64_t>(split) + static_cast<int64_t>(w + kbi * kB16Stages) * p.splits) * kB16Cols + c32 * 8;
    };
    auto prefetch_for = [&](int kbi) {                    // k-block kbi + distance of THIS warp
      if (p.prefetch > 0 && kbi + p.prefetch < my_kb) prefetch_rows_l2(p, col_of(kbi + p.prefetch) - c32 * 8, lane);
    };
    Batch B0, B1, B2;
    auto issue = [&](Batch& B) {                          // load the next batch of the stream (if any)
      if (kb_l >= my_kb) return;
      if (b_l == 0) prefetch_for(kb_l);
      const int64_t col = col_of(kb_l);
      load_batch(B, p, p.G + (col < p.d ? col : 0), col, b_l, sub);
      if (++b_l == nbatch) { b_l = 0; ++kb_l; }
    };
    auto consume = [&](const Batch& B) {                  // convert + store the oldest batch in flight
      if (b_s == 0) mbar_wait_warp(&empty_bar[w], (kb_s & 1) ^ 1);
      store_batch(B, p, st, b_s, sub, c32);
      if (++b_s == nbatch) {
        b_s = 0; ++kb_s;
        fence_proxy_async_smem();
        named_bar_arrive(1 + w, 64);
      }
    };
    if (total > 0) {
      for (int k = 0; k < p.prefetch && k < my_kb; ++k) prefetch_rows_l2(p, col_of(k) - c32 * 8, lane);
      issue(B0);
      issue(B1);
      for (int g = 0; g < total; g += 3) {
        issue(B2); consume(B0);
        if (g + 1 >= total) break;
        issue(B0); consume(B1);
        if (g + 2 >= total) break;
        issue(B1); consume(B2);
      }
    }
  } else {
    setmaxnreg_inc<248>();
    // ===================== epilogue: drain TMEM chains into fp32 registers =====================
    const int q = warp & 3;                               // TMEM lane quadrant of this warp
    const int ncols = 2 * p.nb;                           // <= 224
    float run[224];
#pragma unroll
    for (int i = 0; i < 224; ++i) run[i] = 0.f;
    for (int g = 0; g < ngroups; ++g) {
      const int b = g & 1;
      mbar_wait_warp(&acc_full[b], (g >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(b * 256);
#pragma unroll
      for (int c = 0; c < 14; ++c) {
        if (c * 16 < ncols) {
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
    float* out = p.parts + static_cast<size_t>(split) * kB16PartElems + static_cast<size_t>(q * 32 + lane) * 128;
#pragma unroll
    for (int c = 0; c < 224; c += 4) {
      if (c < ncols) {
        const int a = (c >= p.nb) ? 1 : 0;                // nb is a multiple of 16, so a float4 never straddles
        *reinterpret_cast<float4*>(out + static_cast<size_t>(a) * 128 * 128 + (c - a * p.nb)) =
            make_float4(run[c], run[c + 1], run[c + 2], run[c + 3]);
      }
    }
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

// parts must hold bf16x2_splits(d) * 2*128*128 floats.  Columns n..nb-1 of the slots are zero-filled by
// nobody: gram_reduce_kernel only reads rows/columns < n.
int launch_bf16x2(const float* G, int n, int64_t d, int64_t ld, float* parts, int splits, int flush,
                  cudaStream_t stream) {
  B16Params p{};
  p.G = G; p.ld = ld; p.d = d; p.n = n; p.nb = (n + 15) & ~15;
  p.splits = splits;
  p.kblocks = static_cast<int>((d + kB16Cols - 1) / kB16Cols);
  p.flush = flush < 2 ? 2 : (flush & ~1);
  p.parts = parts;
  p.prefetch = 0;
  if (const char* e = getenv("AFL_GRAM_KNOCK")) p.knock = atoi(e);
  if (const char* e = getenv("AFL_GRAM_PREFETCH")) p.prefetch = atoi(e) < 0 ? 0 : atoi(e);
  const size_t smem = static_cast<size_t>(kB16Stages) * p.nb * 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    AFL_CUDA(cudaFuncSetAttribute(gram_bf16x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  {
    ProfScope ps("gram_tcgen05", stream);
    gram_bf16x2_kernel<<<splits, kB16Threads, smem, stream>>>(p);
  }
  AFL_LAUNCH_CHECK("gram_bf16x2_kernel");
  return AFL_OK;
}

}  // namespace gram
}  // namespace afl
