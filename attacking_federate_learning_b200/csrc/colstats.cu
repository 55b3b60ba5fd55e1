// HBM-bound column kernels: plain mean (defences.py:13-14), the ALIE mu/sigma/perturb reduction
// (malicious.py:18-19,35), Krum's row gather, and the server momentum step (server.py:89-90).
// All of them stream the row-major [n, d] matrix once with 16-byte loads; a thread owns VEC adjacent
// columns and walks down the rows, so every warp-level load is one contiguous segment of a row.
#include "afl_common.cuh"

namespace afl {
namespace colstats {

constexpr int kBlock = 256;
constexpr int kUnroll = 8;

template <int VEC> struct Pack { float v[VEC]; };

// Load VEC consecutive columns of one row as fp32.  VEC = 4 (fp32, 16 B) or 8 (bf16, 16 B) or 1 (scalar).
template <typename T, int VEC>
__device__ __forceinline__ Pack<VEC> load_pack(const T* p);
template <> __device__ __forceinline__ Pack<4> load_pack<float, 4>(const float* p) {
  const float4 t = ldg_stream_f4(reinterpret_cast<const float4*>(p));
  return Pack<4>{{t.x, t.y, t.z, t.w}};
}
template <> __device__ __forceinline__ Pack<1> load_pack<float, 1>(const float* p) { return Pack<1>{{__ldg(p)}}; }
template <> __device__ __forceinline__ Pack<8> load_pack<__nv_bfloat16, 8>(const __nv_bfloat16* p) {
  const uint4 t = ldg_stream_u4(reinterpret_cast<const uint4*>(p));
  Pack<8> r;
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = bf16_bits_to_f32(w[i] & 0xFFFFu);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
  return r;
}
template <> __device__ __forceinline__ Pack<1> load_pack<__nv_bfloat16, 1>(const __nv_bfloat16* p) {
  return Pack<1>{{__bfloat162float(*p)}};
}

// Coherent variants (no .nc, no L1 bypass hint): used when the kernel also WRITES the matrix it reads
// (ALIE writing the crafted vector back into the malicious rows it just reduced).
template <typename T, int VEC>
__device__ __forceinline__ Pack<VEC> load_pack_coherent(const T* p) {
  Pack<VEC> r;
  if constexpr (sizeof(T) == 4) {
    if constexpr (VEC == 4) {
      float4 t;
      asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(p) : "memory");
      r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
      float t;
      asm volatile("ld.global.f32 %0, [%1];" : "=f"(t) : "l"(p) : "memory");
      r.v[0] = t;
    }
  } else {
    if constexpr (VEC == 8) {
      uint32_t w[4];
      asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p) : "memory");
#pragma unroll
      for (int i = 0; i < 4; ++i) { r.v[2 * i] = bf16_bits_to_f32(w[i] & 0xFFFFu); r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
    } else {
      uint16_t t;
      asm volatile("ld.global.u16 %0, [%1];" : "=h"(t) : "l"(p) : "memory");
      r.v[0] = bf16_bits_to_f32(t);
    }
  }
  return r;
}
template <typename T, int VEC, bool COHERENT>
__device__ __forceinline__ Pack<VEC> load_rows(const T* p) {
  if constexpr (COHERENT) return load_pack_coherent<T, VEC>(p);
  else return load_pack<T, VEC>(p);
}

// ---- mean: sequential fp32 row accumulation, then one IEEE division — the order NumPy uses for
// np.mean(axis=0) on a C-contiguous array, so fp32 results are bit-identical to the reference.
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
mean_kernel(const T* __restrict__ G, int n, int64_t d, int64_t ld, float* __restrict__ out) {
  const int64_t c0 = (static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x) * VEC;
  if (c0 >= d) return;
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  const T* p = G + c0;
  int r = 0;
  for (; r + kUnroll <= n; r += kUnroll) {
    Pack<VEC> t[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) t[u] = load_pack<T, VEC>(p + static_cast<int64_t>(r + u) * ld);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = __fadd_rn(acc[k], t[u].v[k]);
  }
  for (; r < n; ++r) {
    Pack<VEC> t = load_pack<T, VEC>(p + static_cast<int64_t>(r) * ld);
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = __fadd_rn(acc[k], t.v[k]);
  }
  const float fn = static_cast<float>(n);
#pragma unroll
  for (int k = 0; k < VEC; ++k)
    if (c0 + k < d) out[c0 + k] = __fdiv_rn(acc[k], fn);
}

// ---- ALIE: one pass, sum(x) and sum(x^2) per column in FLOAT64 (B200 has a full-rate FP64 pipe: 3 FP64 instructions
// per value hide under the HBM stream).  x*x is exact in float64 and the sums carry ~1e-16 relative error, so
// E[x^2] - E[x]^2 is accurate even when one client is orders of magnitude larger than the column's sigma (the
// round-1 version shifted by the first row in fp32 and lost ~1e-5 relative there).  sigma = sqrt(population
// variance); crafted = mu - z*sigma with the same two fp32 roundings as `grads_mean[:] -= num_std * grads_stdev[:]`.
template <typename T, int VEC, bool COHERENT>
__global__ void __launch_bounds__(kBlock)
alie_kernel(const T* G, int f, int64_t d, int64_t ld, float z, float* __restrict__ mu_out,
            float* __restrict__ sigma_out, float* __restrict__ crafted_out, float* bcast, int64_t bcast_ld) {
  const int64_t c0 = (static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x) * VEC;
  if (c0 >= d) return;
  const T* p = G + c0;
  double s1[VEC], s2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { s1[k] = 0.0; s2[k] = 0.0; }
  int r = 0;
  for (; r + kUnroll <= f; r += kUnroll) {
    Pack<VEC> t[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) t[u] = load_rows<T, VEC, COHERENT>(p + static_cast<int64_t>(r + u) * ld);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const double x = static_cast<double>(t[u].v[k]);
        s1[k] += x;
        s2[k] = fma(x, x, s2[k]);
      }
  }
  for (; r < f; ++r) {
    Pack<VEC> t = load_rows<T, VEC, COHERENT>(p + static_cast<int64_t>(r) * ld);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const double x = static_cast<double>(t.v[k]);
      s1[k] += x;
      s2[k] = fma(x, x, s2[k]);
    }
  }
  const double inv = 1.0 / static_cast<double>(f);
  float crafted[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const double m = s1[k] * inv;
    double var = fma(-m, m, s2[k] * inv);
    var = var > 0.0 ? var : 0.0;
    const float mu = static_cast<float>(m);
    const float sigma = static_cast<float>(sqrt(var));
    crafted[k] = __fsub_rn(mu, __fmul_rn(z, sigma));
    if (c0 + k < d) {
      if (sigma_out) sigma_out[c0 + k] = sigma;
      if (mu_out && mu_out != crafted_out) mu_out[c0 + k] = mu;
      if (crafted_out) crafted_out[c0 + k] = crafted[k];
    }
  }
  if (bcast) {
    // server.py:82-83 copies the one aliased array into every malicious row: f row segments of VEC floats,
    // written as 16-byte stores when the destination allows it (all of this thread's reads are done)
    const bool v16 = (VEC % 4 == 0) && (c0 + VEC <= d) && (bcast_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(bcast) & 15) == 0);
    for (int rr = 0; rr < f; ++rr) {
      float* dst = bcast + static_cast<int64_t>(rr) * bcast_ld + c0;
      if (v16) {
#pragma unroll
        for (int k = 0; k + 4 <= VEC; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(crafted[k], crafted[k + 1], crafted[k + 2], crafted[k + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          if (c0 + k < d) dst[k] = crafted[k];
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
gather_row_kernel(const T* __restrict__ G, int n, int64_t d, int64_t ld, const int* __restrict__ idx_dev,
                  float* __restrict__ out) {
  int idx = *idx_dev;
  if (idx < 0) idx += n;            // the reference's users_grads[-1] when nobody wins
  if (idx < 0 || idx >= n) return;
  const T* row = G + static_cast<int64_t>(idx) * ld;
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x; c < d;
       c += static_cast<int64_t>(gridDim.x) * kBlock)
    out[c] = load_pack<T, 1>(row + c).v[0];
}

__global__ void __launch_bounds__(kBlock)
momentum_kernel(float* __restrict__ w, float* __restrict__ v, const float* __restrict__ g, int64_t d, float momentum,
                float lr) {
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x; c < d;
       c += static_cast<int64_t>(gridDim.x) * kBlock) {
    // server.py:89  velocity = momentum*velocity - lr*grads  (two products, one subtraction, fp32)
    const float nv = __fsub_rn(__fmul_rn(momentum, v[c]), __fmul_rn(lr, g[c]));
    v[c] = nv;
    w[c] = __fadd_rn(w[c], nv);                      // server.py:90
  }
}

// ALIE band: lo = mu - z*sigma, hi = mu + z*sigma (fp32 product, fp32 sum: the roundings of
// `grads_mean -/+ num_std * grads_stdev`).  x == nullptr: out = lo (malicious.py:35, DriftAttack);
// otherwise out = np.clip(x, lo, hi) = minimum(maximum(x, lo), hi) with NumPy's NaN propagation
// (backdoor.py:60-61).  out may alias mu or x (each element is read before it is written).
__global__ void __launch_bounds__(kBlock)
band_kernel(const float* mu, const float* sigma, float z, const float* x, float* out, int64_t d) {
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x; c < d;
       c += static_cast<int64_t>(gridDim.x) * kBlock) {
    const float m = mu[c], zs = __fmul_rn(z, sigma[c]);
    const float lo = __fsub_rn(m, zs);
    if (x == nullptr) { out[c] = lo; continue; }
    const float hi = __fadd_rn(m, zs);
    const float v = x[c];
    float r = fminf(fmaxf(v, lo), hi);
    if (v != v || lo != lo || hi != hi) r = __int_as_float(0x7fc00000);
    out[c] = r;
  }
}

static bool vec_ok(const void* G, int64_t ld, int dtype) {
  const int64_t es = dtype == AFL_F32 ? 4 : 2;
  return (reinterpret_cast<uintptr_t>(G) % 16 == 0) && ((ld * es) % 16 == 0);
}

int mean(const void* G, int n, int64_t d, int64_t ld, int dtype, float* out, cudaStream_t stream) {
  if (!G || !out || n < 1 || d < 1 || ld < d) { set_error("afl_mean: bad argument"); return AFL_ERR_BAD_ARG; }
  if (dtype != AFL_F32 && dtype != AFL_BF16) { set_error("afl_mean: dtype"); return AFL_ERR_UNSUPPORTED; }
  const bool v = vec_ok(G, ld, dtype);
  const int vec = v ? (dtype == AFL_F32 ? 4 : 8) : 1;
  const unsigned grid = static_cast<unsigned>(ceil_div64(ceil_div64(d, vec), kBlock));
  ProfScope ps("mean", stream);
  if (dtype == AFL_F32) {
    if (v) mean_kernel<float, 4><<<grid, kBlock, 0, stream>>>(static_cast<const float*>(G), n, d, ld, out);
    else mean_kernel<float, 1><<<grid, kBlock, 0, stream>>>(static_cast<const float*>(G), n, d, ld, out);
  } else {
    if (v) mean_kernel<__nv_bfloat16, 8><<<grid, kBlock, 0, stream>>>(static_cast<const __nv_bfloat16*>(G), n, d, ld, out);
    else mean_kernel<__nv_bfloat16, 1><<<grid, kBlock, 0, stream>>>(static_cast<const __nv_bfloat16*>(G), n, d, ld, out);
  }
  AFL_LAUNCH_CHECK("mean_kernel");
  return AFL_OK;
}

int alie(const void* G, int f, int64_t d, int64_t ld, int dtype, double z, float* mu_out, float* sigma_out,
         float* crafted_out, float* bcast, int64_t bcast_ld, cudaStream_t stream) {
  if (!G || f < 1 || d < 1 || ld < d || (bcast && bcast_ld < d)) { set_error("afl_alie: bad argument"); return AFL_ERR_BAD_ARG; }
  if (dtype != AFL_F32 && dtype != AFL_BF16) { set_error("afl_alie: dtype"); return AFL_ERR_UNSUPPORTED; }
  const bool v = vec_ok(G, ld, dtype);
  const int vec = v ? (dtype == AFL_F32 ? 4 : 8) : 1;
  const unsigned grid = static_cast<unsigned>(ceil_div64(ceil_div64(d, vec), kBlock));
  const float zf = static_cast<float>(z);
  ProfScope ps("alie", stream);
  // the reference writes the crafted vector back over the malicious rows (main.py:68 -> server.py:82-83): when the
  // broadcast target overlaps the input, read it coherently (no ld.global.nc on memory this launch writes)
  const uintptr_t g0 = reinterpret_cast<uintptr_t>(G), g1 = g0 + static_cast<uintptr_t>((static_cast<int64_t>(f - 1) * ld + d) * (dtype == AFL_F32 ? 4 : 2));
  const uintptr_t b0 = reinterpret_cast<uintptr_t>(bcast), b1 = b0 + static_cast<uintptr_t>((static_cast<int64_t>(f - 1) * bcast_ld + d) * 4);
  const bool coh = bcast && b0 < g1 && g0 < b1;
#define AFL_ALIE_LAUNCH(T, V, C) alie_kernel<T, V, C><<<grid, kBlock, 0, stream>>>(static_cast<const T*>(G), f, d, ld, zf, mu_out, sigma_out, crafted_out, bcast, bcast_ld)
  if (dtype == AFL_F32) {
    if (v) { if (coh) AFL_ALIE_LAUNCH(float, 4, true); else AFL_ALIE_LAUNCH(float, 4, false); }
    else { if (coh) AFL_ALIE_LAUNCH(float, 1, true); else AFL_ALIE_LAUNCH(float, 1, false); }
  } else {
    if (v) { if (coh) AFL_ALIE_LAUNCH(__nv_bfloat16, 8, true); else AFL_ALIE_LAUNCH(__nv_bfloat16, 8, false); }
    else { if (coh) AFL_ALIE_LAUNCH(__nv_bfloat16, 1, true); else AFL_ALIE_LAUNCH(__nv_bfloat16, 1, false); }
  }
#undef AFL_ALIE_LAUNCH
  AFL_LAUNCH_CHECK("alie_kernel");
  return AFL_OK;
}

int gather_row(const void* G, int n, int64_t d, int64_t ld, int dtype, const int* idx_dev, float* out,
               cudaStream_t stream) {
  if (!G || !idx_dev || !out || n < 1 || d < 1 || ld < d) { set_error("afl_gather_row: bad argument"); return AFL_ERR_BAD_ARG; }
  int64_t blocks = ceil_div64(d, kBlock);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == AFL_F32)
    gather_row_kernel<float><<<static_cast<unsigned>(blocks), kBlock, 0, stream>>>(static_cast<const float*>(G), n, d, ld, idx_dev, out);
  else if (dtype == AFL_BF16)
    gather_row_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), kBlock, 0, stream>>>(static_cast<const __nv_bfloat16*>(G), n, d, ld, idx_dev, out);
  else { set_error("afl_gather_row: dtype"); return AFL_ERR_UNSUPPORTED; }
  AFL_LAUNCH_CHECK("gather_row_kernel");
  return AFL_OK;
}

int momentum_step(float* w, float* v, const float* g, int64_t d, float momentum, float lr, cudaStream_t stream) {
  if (!w || !v || !g || d < 1) { set_error("afl_momentum_step: bad argument"); return AFL_ERR_BAD_ARG; }
  int64_t blocks = ceil_div64(d, kBlock);
  if (blocks > 148 * 16) blocks = 148 * 16;
  momentum_kernel<<<static_cast<unsigned>(blocks), kBlock, 0, stream>>>(w, v, g, d, momentum, lr);
  AFL_LAUNCH_CHECK("momentum_kernel");
  return AFL_OK;
}

int alie_band(const float* mu, const float* sigma, double z, const float* x, float* out, int64_t d,
              cudaStream_t stream) {
  if (!mu || !sigma || !out || d < 1) { set_error("afl_alie_band: bad argument"); return AFL_ERR_BAD_ARG; }
  int64_t blocks = ceil_div64(d, kBlock);
  if (blocks > 148 * 16) blocks = 148 * 16;
  band_kernel<<<static_cast<unsigned>(blocks), kBlock, 0, stream>>>(mu, sigma, static_cast<float>(z), x, out, d);
  AFL_LAUNCH_CHECK("band_kernel");
  return AFL_OK;
}

}  // namespace colstats
}  // namespace afl
