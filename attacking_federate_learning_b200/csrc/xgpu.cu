// The one exchange step of the D-sharded path over NVLink peer memory, and the fused Krum tail.
//
// Every rank owns ONE cudaMalloc'ed block, exported with cudaIpcGetMemHandle and mapped by its peers:
//     [table 0: n_max^2 float64][table 1][flags: world x u64][done counter]
// afl_sqdist_partial writes the rank's partial squared-distance table straight into table (epoch & 1); a one-thread
// `publish` kernel then stores the epoch into flags[rank] of EVERY rank (release, system scope).  A consumer kernel
// waits until all `world` flags of its own block have reached the epoch (acquire) and then reads the peers' tables
// through the mapped pointers, adding them in rank order -> the sum is bit-identical on every rank (selection stays
// replicated and deterministic, no broadcast), and there is no NCCL launch, no host round trip and no reduction
// tree on the latency path of an 80 KB (N = 100) .. 8 MB (N = 1000) table.  Two tables alternate by epoch: a rank can
// run at most one step ahead of the slowest peer (its consumer kernel waits for everybody's publish of that epoch).
//
//   krum_tail_kernel   one CTA per client u: [wait] -> sum of the ranks' d2 rows -> sqrt -> bitonic sort of the row ->
//                      ascending sequential fp32 sum of the `take` smallest (defences.py:33-34) -> score[u]; the last
//                      CTA to finish does the strict-< argmin in the dict order [1, 0, 2, ...] (defences.py:35-37) and
//                      writes the index to the device AND to mapped pinned host memory, so a Krum step ends with one
//                      stream synchronisation instead of a blocking 4-byte memcpy.
//   xgpu_sum_kernel    [wait] -> elementwise sum of the ranks' tables into a local table (Bulyan keeps its own selection
//                      kernels).
#include <string.h>

#include "afl_common.cuh"

namespace afl {
namespace gram {
int sqdist_partial(const void* G, int n, int64_t d, int64_t ld, int dtype, double* d2_out, void* ws, size_t ws_bytes,
                   int flags, cudaStream_t stream);
int sqdist_partial_ex(const void* G, int n, int64_t d, int64_t ld, int dtype, double* d2_out, void* ws, size_t ws_bytes,
                      int flags, cudaStream_t stream, const PubHook* hook, bool* hook_done);
}
namespace xgpu {

struct Ctx {
  int world, rank, n_max, device;
  size_t table_bytes, block_bytes;
  uint8_t* block;                       // own allocation
  uint8_t* peer[kMaxWorld];             // mapped blocks (peer[rank] == block)
  bool opened[kMaxWorld];
  unsigned long long epoch;
  float* score;                         // [n_max] local scratch
  int* idx_dev;                         // device copy of the last index
  int* idx_host;                        // mapped pinned host memory
  int* idx_host_devptr;
  int* status_host;                     // 0 ok, 1 = flag wait timed out
  int* status_host_devptr;
};

struct TailParams {
  const double* tab[kMaxWorld];         // the ranks' partial tables of this epoch (tab[0] only when world == 1)
  const unsigned long long* flags;      // own flags
  unsigned long long epoch;
  int world, n, take;
  float* score;
  unsigned int* done;
  int* idx_dev; int* idx_host; int* status_host;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {       // never served from a stale L1 line
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Wait (bounded) until every rank has published `epoch`.  Returns false on timeout.
__device__ __forceinline__ bool wait_flags(const unsigned long long* flags, int world, unsigned long long epoch) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
      if (clock64() - t0 > (1ll << 33)) { s_ok = 0; break; }           // ~4 s: a peer died or never launched
    }
  }
  __syncthreads();
  return s_ok != 0;
}

struct PublishParams { unsigned long long* flag[kMaxWorld]; int world, rank; unsigned long long epoch; };
__global__ void publish_kernel(const PublishParams p) {
  if (threadIdx.x < p.world) {
    __threadfence_system();                                              // the table written by earlier kernels of this stream
    st_release_sys(p.flag[threadIdx.x] + p.rank, p.epoch);
  }
}

__device__ __forceinline__ int visit_pos(int u) { return u == 1 ? 0 : (u == 0 ? 1 : u); }

__global__ void __launch_bounds__(256)
krum_tail_kernel(const TailParams p) {
  extern __shared__ uint32_t keys[];                                     // P2 distance bit patterns
  __shared__ float s_val[8];
  __shared__ int s_pos[8];
  __shared__ int s_last;
  const int u = blockIdx.x, n = p.n;
  if (p.world > 1 && !wait_flags(p.flags, p.world, p.epoch)) {
    if (threadIdx.x == 0 && u == 0) { *p.status_host = 1; *p.idx_host = -1; *p.idx_dev = -1; }
    return;
  }
  int P2 = 1;
  while (P2 < n) P2 <<= 1;
  for (int v = threadIdx.x; v < P2; v += blockDim.x) {
    uint32_t k = 0xFFFFFFFFu;
    if (v < n && v != u) {
      double s = 0.0;
      for (int r = 0; r < p.world; ++r)                                  // fixed rank order: identical sum on every rank
        s += p.world > 1 ? ld_peer_f64(p.tab[r] + static_cast<size_t>(u) * n + v) : p.tab[0][static_cast<size_t>(u) * n + v];
      const float dist = static_cast<float>(sqrt(s > 0.0 ? s : 0.0));   // defences.py:20 (np.float32 norm)
      k = __float_as_uint(dist) & 0x7FFFFFFFu;                           // >= 0 or NaN: bit pattern orders like the value
    }
    keys[v] = k;
  }
  __syncthreads();
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P2 >> 1); t += blockDim.x) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const uint32_t a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    float s = 0.f;                                                       // Python: sum() starts at int 0; ascending fp32 adds
    for (int pos = 0; pos < p.take; ++pos) s = s + __uint_as_float(keys[pos]);
    p.score[u] = s;
    __threadfence();
    s_last = (atomicAdd(p.done, 1u) == static_cast<unsigned>(n - 1)) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  // ---- the last CTA: strict-< argmin from (1e20, -1) in the reference's visit order
  __threadfence();
  float best = __int_as_float(0x7f800000);
  int best_pos = 0x7fffffff;
  for (int v = threadIdx.x; v < n; v += blockDim.x) {
    const float s = __ldcg(p.score + v);
    if (n >= 2 && static_cast<double>(s) < 1e20) {
      const int pos = visit_pos(v);
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
  if (threadIdx.x == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
      if (s_val[w] < best || (s_val[w] == best && s_pos[w] < best_pos)) { best = s_val[w]; best_pos = s_pos[w]; }
    int idx = -1;
    if (best_pos != 0x7fffffff) idx = best_pos == 0 ? 1 : (best_pos == 1 ? 0 : best_pos);
    *p.idx_dev = idx;
    *p.idx_host = idx;
    *p.status_host = 0;
    *p.done = 0u;                                                        // ready for the next step
    __threadfence_system();
  }
}

struct SumParams { const double* tab[kMaxWorld]; const unsigned long long* flags; unsigned long long epoch; int world; size_t count; double* out; int* status_host; };
__global__ void __launch_bounds__(256)
xgpu_sum_kernel(const SumParams p) {
  if (!wait_flags(p.flags, p.world, p.epoch)) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *p.status_host = 1;
    return;
  }
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < p.count; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < p.world; ++r) s += ld_peer_f64(p.tab[r] + e);
    p.out[e] = s;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *p.status_host = 0;
}

static size_t flags_offset(const Ctx* c) { return 2 * c->table_bytes; }

int create(int world, int rank, int n_max, void** out) {
  if (!out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || n_max < 1) { set_error("afl_xgpu_create: bad argument"); return AFL_ERR_BAD_ARG; }
  Ctx* c = new Ctx();
  memset(c, 0, sizeof(*c));
  c->world = world; c->rank = rank; c->n_max = n_max; c->device = current_device();
  c->table_bytes = align_up(static_cast<size_t>(n_max) * n_max * sizeof(double), 256);
  c->block_bytes = 2 * c->table_bytes + align_up(sizeof(unsigned long long) * kMaxWorld + 64, 256);
  AFL_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->block), c->block_bytes));
  AFL_CUDA(cudaMemset(c->block, 0, c->block_bytes));
  AFL_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->score), sizeof(float) * n_max + 256));
  AFL_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->idx_dev), 256));
  AFL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->idx_host), 64, cudaHostAllocMapped));
  AFL_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->idx_host_devptr), c->idx_host, 0));
  c->status_host = c->idx_host + 8;
  c->status_host_devptr = c->idx_host_devptr + 8;
  c->idx_host[0] = -1; c->status_host[0] = 0;
  c->peer[rank] = c->block; c->opened[rank] = false;
  AFL_CUDA(cudaDeviceSynchronize());
  *out = c;
  return AFL_OK;
}

int local_handle(void* ctx, unsigned char* out64) {
  Ctx* c = static_cast<Ctx*>(ctx);
  if (!c || !out64) { set_error("afl_xgpu_handle: bad argument"); return AFL_ERR_BAD_ARG; }
  cudaIpcMemHandle_t h;
  AFL_CUDA(cudaIpcGetMemHandle(&h, c->block));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &h, 64);
  return AFL_OK;
}

int connect(void* ctx, const unsigned char* handles) {
  Ctx* c = static_cast<Ctx*>(ctx);
  if (!c || !handles) { set_error("afl_xgpu_connect: bad argument"); return AFL_ERR_BAD_ARG; }
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + static_cast<size_t>(r) * 64, 64);
    void* p = nullptr;
    AFL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[r] = static_cast<uint8_t*>(p);
    c->opened[r] = true;
  }
  return AFL_OK;
}

int destroy(void* ctx) {
  Ctx* c = static_cast<Ctx*>(ctx);
  if (!c) return AFL_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->block); cudaFree(c->score); cudaFree(c->idx_dev); cudaFreeHost(c->idx_host);
  delete c;
  return AFL_OK;
}

static double* table_of(const Ctx* c, int r, unsigned long long epoch) {
  return reinterpret_cast<double*>(c->peer[r] + (epoch & 1ull) * c->table_bytes);
}

// the same flags as a PubHook for the reduction kernel's last block (counter: second word after the done counter)
static PubHook make_hook(Ctx* c) {
  PubHook h{};
  for (int r = 0; r < c->world; ++r) h.flag[r] = reinterpret_cast<unsigned long long*>(c->peer[r] + flags_offset(c));
  h.world = c->world; h.rank = c->rank; h.epoch = c->epoch;
  h.counter = reinterpret_cast<unsigned int*>(c->block + flags_offset(c) + sizeof(unsigned long long) * kMaxWorld) + 4;
  return h;
}

static int publish(Ctx* c, cudaStream_t stream) {
  if (c->world == 1) return AFL_OK;
  PublishParams pp{};
  for (int r = 0; r < c->world; ++r) pp.flag[r] = reinterpret_cast<unsigned long long*>(c->peer[r] + flags_offset(c));
  pp.world = c->world; pp.rank = c->rank; pp.epoch = c->epoch;
  publish_kernel<<<1, 32, 0, stream>>>(pp);
  AFL_LAUNCH_CHECK("publish_kernel");
  return AFL_OK;
}

static int python_slice_take(int m, int len) {   // len(errors[:m])
  if (m >= 0) return m < len ? m : len;
  const int t = len + m;
  return t > 0 ? t : 0;
}

// Whole sharded Krum step on this rank's [n, d_local] shard: partial table -> publish -> fused tail.  Enqueues only;
// *idx_host_out points to mapped pinned memory holding the index once `stream` has been synchronised.
int krum_step(void* ctx, const void* G, int n, int64_t d, int64_t ld, int dtype, int users_count, int corrupted_count,
              void* ws, size_t ws_bytes, int flags, cudaStream_t stream, int** idx_host_out, int** status_host_out,
              int** idx_dev_out) {
  Ctx* c = static_cast<Ctx*>(ctx);
  if (!c || n < 1 || n > c->n_max) { set_error("afl_krum_sharded: bad context or n > n_max"); return AFL_ERR_BAD_ARG; }
  if (c->device != current_device()) { set_error("afl_krum_sharded: context belongs to device %d", c->device); return AFL_ERR_BAD_ARG; }
  c->epoch += 1;
  double* mine = table_of(c, c->rank, c->epoch);
  const PubHook hook = make_hook(c);
  bool published = false;
  int rc = gram::sqdist_partial_ex(G, n, d, ld, dtype, mine, ws, ws_bytes, flags, stream, &hook, &published);
  if (rc) return rc;
  if (!published) {
    rc = publish(c, stream);
    if (rc) return rc;
  }
  TailParams tp{};
  for (int r = 0; r < c->world; ++r) tp.tab[r] = table_of(c, r, c->epoch);
  tp.flags = reinterpret_cast<const unsigned long long*>(c->block + flags_offset(c));
  tp.epoch = c->epoch; tp.world = c->world; tp.n = n;
  tp.take = python_slice_take(users_count - corrupted_count, n - 1);
  tp.score = c->score;
  tp.done = reinterpret_cast<unsigned int*>(c->block + flags_offset(c) + sizeof(unsigned long long) * kMaxWorld);
  tp.idx_dev = c->idx_dev; tp.idx_host = c->idx_host_devptr; tp.status_host = c->status_host_devptr;
  int P2 = 1; while (P2 < n) P2 <<= 1;
  {
    ProfScope ps("krum_tail", stream);
    krum_tail_kernel<<<n, 256, static_cast<size_t>(P2) * sizeof(uint32_t), stream>>>(tp);
  }
  AFL_LAUNCH_CHECK("krum_tail_kernel");
  if (idx_host_out) *idx_host_out = c->idx_host;
  if (status_host_out) *status_host_out = c->status_host;
  if (idx_dev_out) *idx_dev_out = c->idx_dev;
  return AFL_OK;
}

// Partial table -> publish -> sum of all ranks' tables into d2_total (local device memory, n*n float64).
int sqdist_allreduce(void* ctx, const void* G, int n, int64_t d, int64_t ld, int dtype, double* d2_total, void* ws,
                     size_t ws_bytes, int flags, cudaStream_t stream, int** status_host_out) {
  Ctx* c = static_cast<Ctx*>(ctx);
  if (!c || !d2_total || n < 1 || n > c->n_max) { set_error("afl_sqdist_allreduce: bad context or n > n_max"); return AFL_ERR_BAD_ARG; }
  if (c->device != current_device()) { set_error("afl_sqdist_allreduce: context belongs to device %d", c->device); return AFL_ERR_BAD_ARG; }
  if (c->world == 1) return gram::sqdist_partial(G, n, d, ld, dtype, d2_total, ws, ws_bytes, flags, stream);
  c->epoch += 1;
  const PubHook hook = make_hook(c);
  bool published = false;
  int rc = gram::sqdist_partial_ex(G, n, d, ld, dtype, table_of(c, c->rank, c->epoch), ws, ws_bytes, flags, stream, &hook, &published);
  if (rc) return rc;
  if (!published) {
    rc = publish(c, stream);
    if (rc) return rc;
  }
  SumParams sp{};
  for (int r = 0; r < c->world; ++r) sp.tab[r] = table_of(c, r, c->epoch);
  sp.flags = reinterpret_cast<const unsigned long long*>(c->block + flags_offset(c));
  sp.epoch = c->epoch; sp.world = c->world; sp.count = static_cast<size_t>(n) * n; sp.out = d2_total;
  sp.status_host = c->status_host_devptr;
  const size_t blocks = (sp.count + 255) / 256;
  {
    ProfScope ps("xgpu_sum", stream);
    xgpu_sum_kernel<<<static_cast<unsigned>(blocks > 592 ? 592 : blocks), 256, 0, stream>>>(sp);
  }
  AFL_LAUNCH_CHECK("xgpu_sum_kernel");
  if (status_host_out) *status_host_out = c->status_host;
  return AFL_OK;
}

}  // namespace xgpu
}  // namespace afl

using namespace afl;

extern "C" {
int afl_xgpu_create(int world, int rank, int n_max, void** ctx_out) { return xgpu::create(world, rank, n_max, ctx_out); }
int afl_xgpu_handle(void* ctx, unsigned char* out64) { return xgpu::local_handle(ctx, out64); }
int afl_xgpu_connect(void* ctx, const unsigned char* handles) { return xgpu::connect(ctx, handles); }
int afl_xgpu_destroy(void* ctx) { return xgpu::destroy(ctx); }
int afl_krum_sharded(void* ctx, const void* G, int n, int64_t d, int64_t ld, int dtype, int users_count, int corrupted_count,
                     void* workspace, size_t workspace_bytes, int flags, void* stream, int** idx_host_out,
                     int** status_host_out, int** idx_dev_out) {
  return xgpu::krum_step(ctx, G, n, d, ld, dtype, users_count, corrupted_count, workspace, workspace_bytes, flags,
                         static_cast<cudaStream_t>(stream), idx_host_out, status_host_out, idx_dev_out);
}
int afl_sqdist_allreduce(void* ctx, const void* G, int n, int64_t d, int64_t ld, int dtype, double* d2_total, void* workspace,
                         size_t workspace_bytes, int flags, void* stream, int** status_host_out) {
  return xgpu::sqdist_allreduce(ctx, G, n, d, ld, dtype, d2_total, workspace, workspace_bytes, flags,
                                static_cast<cudaStream_t>(stream), status_host_out);
}
}
