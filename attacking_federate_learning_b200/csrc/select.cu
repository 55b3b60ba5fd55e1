// Krum scoring / argmin (defences.py:23-42) and Bulyan's selection loop (defences.py:57-68) on a
// dense n x n fp32 distance table that is small enough (<= 64 MB at n = 4096) to live in L2.
//
//   row_sort_kernel      one CTA per user: bitonic-sort the user's n-1 distances (value, index) in shared
//                        memory; emit the sorted values, the sorted neighbour indices, the inverse
//                        permutation (rank of every neighbour) and the reference's Krum score
//                        (sequential ascending fp32 sum of the first `take` values — Python's
//                        sum(sorted(...)[:m])).
//   krum_argmin_kernel   strict-< argmin from (1e20, -1) in the reference's visit order [1,0,2,...].
//   bulyan_rounds_kernel one persistent CTA runs all theta rounds.  Removal of the selected user is an
//                        O(1) update per remaining user: its kept set (the m_r smallest alive distances)
//                        loses either the removed neighbour or its current largest kept element, tracked
//                        by a boundary pointer that only ever moves left over the pre-sorted row.  Scores
//                        are kept in float64, so every round's score is the exact sum of the fp32
//                        distances (the reference's fp32 sequential sum differs by rounding noise only).
#include "afl_common.cuh"

namespace afl {
namespace select {

constexpr int kMaxN = 4096;

struct SortWs {
  float* sval;      // [n][n]   sorted distances of row u (first n-1 entries valid)
  uint16_t* sidx;   // [n][n]   neighbour index at each sorted position
  uint16_t* rank;   // [n][n]   rank[u][v] = sorted position of neighbour v in row u
  float* score;     // [n]      Krum score (fp32, reference arithmetic)
};

static size_t ws_bytes_for(int n) {
  const size_t nn = static_cast<size_t>(n) * n;
  return align_up(nn * 4, 256) + 2 * align_up(nn * 2, 256) + align_up(static_cast<size_t>(n) * 4, 256) + 256;
}

static SortWs carve(void* ws, int n) {
  const size_t nn = static_cast<size_t>(n) * n;
  uint8_t* p = static_cast<uint8_t*>(ws);
  SortWs w;
  w.sval = reinterpret_cast<float*>(p); p += align_up(nn * 4, 256);
  w.sidx = reinterpret_cast<uint16_t*>(p); p += align_up(nn * 2, 256);
  w.rank = reinterpret_cast<uint16_t*>(p); p += align_up(nn * 2, 256);
  w.score = reinterpret_cast<float*>(p);
  return w;
}

__device__ __forceinline__ int visit_pos(int u) { return u == 1 ? 0 : (u == 0 ? 1 : u); }

// take: number of smallest distances summed (already resolved from Python slice semantics).
__global__ void __launch_bounds__(256)
row_sort_kernel(const float* __restrict__ dist, int n, int take, SortWs w) {
  extern __shared__ unsigned long long keys[];
  const int u = blockIdx.x;
  int P = 1;
  while (P < n) P <<= 1;
  for (int v = threadIdx.x; v < P; v += blockDim.x) {
    unsigned long long k = ~0ull;
    if (v < n && v != u) {
      // distances are >= 0 (or NaN): the IEEE bit pattern orders like the value; NaN sorts last.
      const uint32_t bits = __float_as_uint(dist[static_cast<size_t>(u) * n + v]) & 0x7FFFFFFFu;
      k = (static_cast<unsigned long long>(bits) << 32) | static_cast<unsigned>(v);
    }
    keys[v] = k;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  const size_t base = static_cast<size_t>(u) * n;
  for (int pos = threadIdx.x; pos < n; pos += blockDim.x) {
    const unsigned long long k = keys[pos];
    if (pos < n - 1) {
      const int v = static_cast<int>(k & 0xFFFFFFFFu);
      w.sval[base + pos] = dist[base + v];               // original bits (keeps a NaN a NaN)
      w.sidx[base + pos] = static_cast<uint16_t>(v);
      w.rank[base + v] = static_cast<uint16_t>(pos);
    } else {
      w.sval[base + pos] = 0.f;
      w.sidx[base + pos] = static_cast<uint16_t>(u);
      w.rank[base + u] = static_cast<uint16_t>(pos);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;                                        // Python: sum() starts at int 0
    for (int pos = 0; pos < take; ++pos) s = s + w.sval[base + pos];
    w.score[u] = s;
  }
}

__global__ void __launch_bounds__(1024)
krum_argmin_kernel(const float* __restrict__ score, int n, int* __restrict__ idx_out, float* __restrict__ scores_out) {
  __shared__ float s_val[32];
  __shared__ int s_pos[32];
  float best = __int_as_float(0x7f800000);   // +inf
  int best_pos = 0x7fffffff;
  for (int u = threadIdx.x; u < n; u += blockDim.x) {
    const float s = score[u];
    if (scores_out) scores_out[u] = s;
    if (n >= 2 && static_cast<double>(s) < 1e20) {       // first comparison is against the Python float 1e20
      const int pos = visit_pos(u);
      if (s < best || (s == best && pos < best_pos)) { best = s; best_pos = pos; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int op = __shfl_xor_sync(0xffffffffu, best_pos, o);
    if (ov < best || (ov == best && op < best_pos)) { best = ov; best_pos = op; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_pos[threadIdx.x >> 5] = best_pos; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = s_val[threadIdx.x]; best_pos = s_pos[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int op = __shfl_xor_sync(0xffffffffu, best_pos, o);
      if (ov < best || (ov == best && op < best_pos)) { best = ov; best_pos = op; }
    }
    if (threadIdx.x == 0) {
      int idx = -1;
      if (best_pos != 0x7fffffff) idx = best_pos == 0 ? 1 : (best_pos == 1 ? 0 : best_pos);
      *idx_out = idx;
    }
  }
}

constexpr int kRowsPerThread = kMaxN / 1024;

__global__ void __launch_bounds__(1024, 1)
bulyan_rounds_kernel(const float* __restrict__ dist, int n, int f, int theta, SortWs w, int* __restrict__ sel_out) {
  __shared__ uint8_t alive[kMaxN];
  __shared__ double r_val[32];
  __shared__ int r_pos[32];
  __shared__ int s_winner;

  const int tid = threadIdx.x;
  double score[kRowsPerThread];
  int bptr[kRowsPerThread];

  for (int v = tid; v < kMaxN; v += 1024) alive[v] = (v < n) ? 1 : 0;
  // round 0: keep0 = min(n - f, n - 1) smallest distances, summed ascending in float64
  const int keep0 = min(n - f, n - 1);
#pragma unroll
  for (int r = 0; r < kRowsPerThread; ++r) {
    const int u = tid + r * 1024;
    score[r] = 0.0; bptr[r] = keep0 - 1;
    if (u < n) {
      const float* sv = w.sval + static_cast<size_t>(u) * n;
      double s = 0.0;
      for (int pos = 0; pos < keep0; ++pos) s += static_cast<double>(sv[pos]);
      score[r] = s;
    }
  }
  __syncthreads();

  for (int round = 0; round < theta; ++round) {
    // ---- argmin over alive users: (score, visit position), strict < with earliest-visited tie-break
    double best = __longlong_as_double(0x7ff0000000000000ll);
    int best_pos = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
      const int u = tid + r * 1024;
      if (u < n && alive[u] && score[r] < 1e20) {
        const int pos = visit_pos(u);
        if (score[r] < best || (score[r] == best && pos < best_pos)) { best = score[r]; best_pos = pos; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int op = __shfl_xor_sync(0xffffffffu, best_pos, o);
      if (ov < best || (ov == best && op < best_pos)) { best = ov; best_pos = op; }
    }
    if ((tid & 31) == 0) { r_val[tid >> 5] = best; r_pos[tid >> 5] = best_pos; }
    __syncthreads();
    if (tid < 32) {
      best = r_val[tid]; best_pos = r_pos[tid];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int op = __shfl_xor_sync(0xffffffffu, best_pos, o);
        if (ov < best || (ov == best && op < best_pos)) { best = ov; best_pos = op; }
      }
      if (tid == 0) {
        int idx = -1;
        if (best_pos != 0x7fffffff) idx = best_pos == 0 ? 1 : (best_pos == 1 ? 0 : best_pos);
        s_winner = idx;
        sel_out[round] = idx;
        if (idx >= 0) alive[idx] = 0;
      }
    }
    __syncthreads();
    const int s = s_winner;
    if (s < 0) {                       // nobody eligible (NaN / >= 1e20 scores): the reference raises here
      for (int r2 = round + 1 + tid; r2 < theta; r2 += 1024) sel_out[r2] = -1;
      return;
    }
    // ---- O(1) update of every surviving user's kept set: size shrinks by one each round
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
      const int u = tid + r * 1024;
      if (u < n && u != s && alive[u]) {
        const size_t base = static_cast<size_t>(u) * n;
        const int pos = w.rank[base + s];
        int b = bptr[r];
        if (b >= 0) {
          if (pos <= b) {
            score[r] -= static_cast<double>(dist[base + s]);
            if (pos == b) { do { --b; } while (b >= 0 && !alive[w.sidx[base + b]]); }
          } else {
            score[r] -= static_cast<double>(w.sval[base + b]);
            do { --b; } while (b >= 0 && !alive[w.sidx[base + b]]);
          }
          bptr[r] = b;
        }
      }
    }
    __syncthreads();
  }
}

size_t workspace_bytes(int n) { return ws_bytes_for(n < 1 ? 1 : n); }

static int python_slice_take(int m, int len) {   // len(errors[:m])
  if (m >= 0) return m < len ? m : len;
  const int t = len + m;
  return t > 0 ? t : 0;
}

static int run_sort(const float* dist, int n, int take, void* ws, size_t ws_bytes, cudaStream_t stream, SortWs* out) {
  if (n > kMaxN) { set_error("selection kernels support n <= %d clients (got %d)", kMaxN, n); return AFL_ERR_UNSUPPORTED; }
  if (!ws || ws_bytes < ws_bytes_for(n) || (reinterpret_cast<uintptr_t>(ws) % 256) != 0) {
    set_error("selection workspace too small or misaligned (%zu < %zu)", ws_bytes, ws_bytes_for(n));
    return AFL_ERR_WORKSPACE;
  }
  SortWs w = carve(ws, n);
  int P = 1; while (P < n) P <<= 1;
  ProfScope ps("row_sort", stream);
  row_sort_kernel<<<n, 256, static_cast<size_t>(P) * sizeof(unsigned long long), stream>>>(dist, n, take, w);
  AFL_LAUNCH_CHECK("row_sort_kernel");
  *out = w;
  return AFL_OK;
}

int krum_select(const float* dist, int n, int users_count, int corrupted_count, int* idx_out, float* scores_out,
                void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!dist || !idx_out || n < 1) { set_error("afl_krum_select: bad argument"); return AFL_ERR_BAD_ARG; }
  const int take = python_slice_take(users_count - corrupted_count, n - 1);
  SortWs w;
  int rc = run_sort(dist, n, take, ws, ws_bytes, stream, &w);
  if (rc) return rc;
  krum_argmin_kernel<<<1, 1024, 0, stream>>>(w.score, n, idx_out, scores_out);
  AFL_LAUNCH_CHECK("krum_argmin_kernel");
  return AFL_OK;
}

int bulyan_select(const float* dist, int n, int users_count, int f, int* sel_out, void* ws, size_t ws_bytes,
                  cudaStream_t stream) {
  if (!dist || !sel_out || n < 1 || f < 0) { set_error("afl_bulyan_select: bad argument"); return AFL_ERR_BAD_ARG; }
  if (users_count < 4 * f + 3) {
    set_error("bulyan: users_count >= 4*corrupted_count + 3 violated (%d, %d)", users_count, f);
    return AFL_ERR_PRECONDITION;
  }
  if (users_count != n) {
    set_error("afl_bulyan_select: users_count (%d) must equal the number of rows (%d)", users_count, n);
    return AFL_ERR_UNSUPPORTED;
  }
  const int theta = users_count - 2 * f;
  SortWs w;
  int rc = run_sort(dist, n, python_slice_take(users_count - f, n - 1), ws, ws_bytes, stream, &w);
  if (rc) return rc;
  ProfScope ps("bulyan_rounds", stream);
  bulyan_rounds_kernel<<<1, 1024, 0, stream>>>(dist, n, f, theta, w, sel_out);
  AFL_LAUNCH_CHECK("bulyan_rounds_kernel");
  return AFL_OK;
}

}  // namespace select
}  // namespace afl
