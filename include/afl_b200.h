/*
 * afl_b200.h — C ABI of the B200-native Byzantine-robust aggregation engine.
 *
 * The upstream reference (shaneson0/attacking_federate_learning) has no FFI: its boundary for this
 * path is a Python dict of callables, `defences.defend[name](users_grads, users_count,
 * corrupted_count)` (defences.py:73-75, called from server.py:87), plus the template-method class
 * `malicious.Attack` (malicious.py:4-36, called from main.py:67-68).  Each entry point below states
 * the reference function (file:line) whose arithmetic it replaces.  The Python mirror of the
 * reference interface (attacking_federate_learning_b200/defences.py, malicious.py) binds these
 * symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - Every function returns an afl_status (0 = ok).  Nothing throws across the ABI.
 *     afl_last_error() returns a thread-local, human-readable description of the last failure.
 *   - "device" pointers are CUDA device pointers on the current device; "host" pointers are plain
 *     host memory (pinned or pageable).  `stream` is a cudaStream_t passed as void* (NULL = legacy
 *     default stream).  Device entry points only enqueue work; they never synchronise.
 *   - Matrices are row-major `[n rows = clients][d columns = parameters]`, `ld` = row pitch in
 *     ELEMENTS (server.py:35 allocates users_grads as a C-contiguous N x D fp32 array, ld == d).
 *   - dtype: AFL_F32 (the reference's only type) or AFL_BF16 (config "trimmed_mean, bf16"); outputs
 *     are always fp32.
 *   - No entry point allocates device memory; scratch comes from a caller-owned workspace whose size
 *     the matching *_workspace_bytes() function reports.
 *   - Multi-GPU: the parameter dimension D is sharded; every GPU calls the same entry points on its
 *     own `[n, d_local]` shard.  The only exchange is a sum-all-reduce of the n x n float64 table
 *     produced by afl_sqdist_partial() (done by the host layer over NCCL) before
 *     afl_sqdist_to_dist().
 */
#ifndef AFL_B200_H_
#define AFL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum afl_status {
  AFL_OK = 0,
  AFL_ERR_BAD_ARG = 1,       /* null pointer, negative size, misaligned buffer ...               */
  AFL_ERR_PRECONDITION = 2,  /* the reference's `assert` would fire (n >= 2f+1, n >= 4f+3)         */
  AFL_ERR_CUDA = 3,          /* a CUDA call failed; see afl_last_error()                          */
  AFL_ERR_UNSUPPORTED = 4,   /* dtype / size outside what the kernels implement                   */
  AFL_ERR_WORKSPACE = 5,     /* workspace too small                                               */
  AFL_ERR_NO_WINNER = 6      /* a Bulyan round found no eligible user (NaN / >= 1e20 scores): the      */
                             /* reference raises KeyError(-1) at defences.py:66 (`distances.pop(-1)`)  */
} afl_status;

typedef enum afl_dtype { AFL_F32 = 0, AFL_BF16 = 1 } afl_dtype;

/* flags for afl_sqdist_partial */
enum {
  AFL_GRAM_AUTO = 0,          /* tcgen05 path when the layout allows TMA, otherwise SIMT           */
  AFL_GRAM_FORCE_SIMT = 1,    /* CUDA-core difference kernel (verification / unaligned pitch)      */
  AFL_GRAM_FORCE_TCGEN05 = 2, /* fail with AFL_ERR_UNSUPPORTED instead of falling back to SIMT     */
  AFL_GRAM_SINGLE_PASS = 4,   /* tcgen05: hi*hi only (plain TF32), for measurement                  */
  AFL_GRAM_TF32X2 = 16,       /* tcgen05: always the TMA + split-TF32 kernel (default outside the bf16x2 range) */
  AFL_GRAM_BF16X2 = 32,       /* tcgen05: bf16x2 kernel (gram_bf16.cu), the default when 64<=N_pad<=112, D>=32768 */
  AFL_GRAM_NO_CENTER = 64,    /* bf16x2 kernels: do not subtract the last client's row while converting (the  */
                              /* default makes the table translation invariant; env AFL_GRAM_CENTER=0 too)    */
  AFL_GRAM_REWRITE_HI = 8     /* accepted, ignored (kind::tf32 was measured to ignore the low 13     */
                              /* mantissa bits of fp32 operands, which is what the split relies on)  */
};

/* ---- library / device ------------------------------------------------------------------------ */
const char* afl_version(void);
const char* afl_last_error(void);
/* Number of SMs, compute capability and free/total HBM bytes of the current device. */
int afl_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes);
/* Number of kernel launches this library has issued since load (all threads). */
uint64_t afl_launch_count(void);

/* ---- in-library kernel timing (measurement aid for bench.py) -----------------------------------
 * While enabled, the dominant kernel of every entry point is bracketed by CUDA events on the stream it
 * is launched on.  afl_profile_read(name, ...) waits for the recorded events of kernel `name`
 * ("gram_bf16x2", "gram_tcgen05", "sqdist_simt", "trimmed_mean", "alie", "mean", "row_sort", "bulyan_rounds"),
 * returns their summed duration and launch count, and forgets them.  on = 1: every bracketed kernel; on = 2: only the
 * dominant kernel of a rule (gram_*, sqdist_simt, trimmed_mean, mean, alie) - one pair of events per step. */
int afl_profile_enable(int on);
int afl_profile_read(const char* kernel, double* total_ms, int* launches);

/* ---- plain mean:  defences.py:13-14  no_defense -> np.mean(users_grads, axis=0) --------------- */
int afl_mean(const void* G, int n, int64_t d, int64_t ld, int dtype, float* out, void* stream);

/* ---- pairwise squared distances:  defences.py:16-21  _krum_create_distances ------------------- */
/* Partial squared L2 distances over this shard's d columns, as a dense symmetric n x n float64
 * table (zero diagonal).  fp32 inputs, AFL_GRAM_AUTO: G*G^T on tcgen05 tensor cores (TMA-fed, bf16x2 or
 * split-TF32 operands, TMEM accumulators; n > 128: lower-triangular 128 x 128 tile pairs), d2_ij =
 * s_ii + s_jj - 2 s_ij on rows centred on the last client's row.  Partial tables of different shards
 * ADD; take the square root only after the all-reduce (afl_sqdist_to_dist). */
size_t afl_sqdist_workspace_bytes(int n, int64_t d, int dtype, int flags);
int afl_sqdist_partial(const void* G, int n, int64_t d, int64_t ld, int dtype, double* d2_out,
                       void* workspace, size_t workspace_bytes, int flags, void* stream);
/* dist[i][j] = (float)sqrt(max(d2[i][j], 0)), zero diagonal — the values the reference keeps in its
 * dict-of-dicts (np.float32 scalars, defences.py:20). */
int afl_sqdist_to_dist(const double* d2, int n, float* dist, void* stream);

/* ---- Krum selection:  defences.py:23-42  krum(..., return_index=True) --------------------------
 * score(u) = sum of the (users_count - corrupted_count) smallest of u's n-1 distances (ascending
 * fp32 sequential sum, like Python's sum(sorted(...)[:m])); strict-< argmin from (1e20, -1) visiting
 * users in the reference's dict order [1, 0, 2, 3, ...].  *idx_out (device int) receives the index
 * or -1.  scores_out (device float[n]) may be NULL.  No precondition check here: like the reference,
 * the `users_count >= 2f+1` assert belongs to the caller that wants the row (afl_krum). */
size_t afl_select_workspace_bytes(int n);
int afl_krum_select(const float* dist, int n, int users_count, int corrupted_count, int* idx_out,
                    float* scores_out, void* workspace, size_t workspace_bytes, void* stream);

/* Convenience: afl_sqdist_to_dist + afl_krum_select in one call (one FFI crossing per aggregation after
 * the all-reduce).  dist_scratch: device float[n*n]. */
int afl_krum_from_sqdist(const double* d2, int n, int users_count, int corrupted_count, float* dist_scratch,
                         int* idx_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Bulyan selection:  defences.py:57-68 ------------------------------------------------------
 * theta = users_count - 2f rounds of Krum-with-removal on one distance table; sel_out (device
 * int[theta]) receives the indices in selection order.  Returns AFL_ERR_PRECONDITION unless
 * users_count >= 4f+3 (defences.py:56). */
int afl_bulyan_select(const float* dist, int n, int users_count, int corrupted_count, int* sel_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- trimmed mean around the median:  defences.py:44-52  trimmed_mean --------------------------
 * Per column: med = median (even count -> fl32 mean of the two middle values); keep the
 * k = rows - corrupted_count - 1 values with the smallest |fl32(x - med)| (ties: earlier row first);
 * out = mean(kept deviations) + med.  `row_index` (device int[n_rows], may be NULL = rows 0..n-1)
 * selects and ORDERS the participating rows — Bulyan's second stage passes its selection sequence
 * here so the theta x D gather of defences.py:70 is never materialised. */
int afl_trimmed_mean(const void* G, int n, int64_t d, int64_t ld, int dtype, const int* row_index,
                     int n_rows, int corrupted_count, float* out, void* stream);

/* ---- gather one row chosen on the device (Krum's result as a dense vector) --------------------- */
int afl_gather_row(const void* G, int n, int64_t d, int64_t ld, int dtype, const int* idx_dev,
                   float* out, void* stream);

/* ---- ALIE attack:  malicious.py:10-27,34-36  Attack.attack + DriftAttack._attack_grads ---------
 * Over the f malicious rows: mu = mean, sigma = sqrt(population variance); crafted = mu - z*sigma.
 * mu_out receives the UNPERTURBED mean when crafted_out != mu_out; pass crafted_out == mu_out to
 * reproduce the reference's in-place `grads_mean[:] -= z*stdev` aliasing.  If bcast_rows != NULL the
 * crafted vector is also written into rows 0..f-1 of that fp32 matrix (pitch bcast_ld), which is
 * what server.py:82-83 does next with the f aliased `usr.grads`. */
int afl_alie(const void* G_mal, int f, int64_t d, int64_t ld, int dtype, double z, float* mu_out,
             float* sigma_out, float* crafted_out, float* bcast_rows, int64_t bcast_ld, void* stream);

/* ---- ALIE band:  malicious.py:35 (x == NULL)  and  backdoor.py:60-61 (x != NULL) --------------
 * lo = mu - z*sigma, hi = mu + z*sigma in fp32 (the roundings of `grads_mean -/+ num_std*grads_stdev`).
 * x == NULL: out = lo, DriftAttack._attack_grads on statistics that already exist on the device.
 * x != NULL: out = np.clip(x, lo, hi) (NaN propagates as in NumPy), the clamp BackdoorAttack applies to
 * the gradient its maliciously trained network asks for.  fp32 device vectors of length d; out may
 * alias mu or x. */
int afl_alie_band(const float* mu, const float* sigma, double z, const float* x, float* out, int64_t d,
                  void* stream);

/* ---- server momentum step:  server.py:89-90 ----------------------------------------------------
 * v = momentum*v - lr*g ;  w += v   (fp32, in place). */
int afl_momentum_step(float* weights, float* velocity, const float* grads, int64_t d, float momentum,
                      float learning_rate, void* stream);

/* ---- multi-GPU exchange over NVLink peer memory ("exactly one sum of the n x n partial table") ------
 * One process per GPU.  Every rank creates a context, exports its 64-byte CUDA IPC handle, the host layer gathers
 * the handles of all ranks (any transport) and every rank maps its peers.  world == 1 needs no connect.
 *   afl_krum_sharded     partial table of this rank's [n, d_local] shard (afl_sqdist_partial) -> published to the peers ->
 *                        fused tail: sum of the ranks' tables in rank order, sqrt, per-client sort, Krum score, argmin
 *                        (defences.py:16-42).  Enqueues only.  After `stream` is synchronised *idx_host_out (mapped
 *                        pinned memory) holds the index (identical on every rank), *status_host_out is 0 (1: a peer did
 *                        not publish in time) and *idx_dev_out is the device copy (for afl_gather_row).
 *   afl_sqdist_allreduce the same exchange, result = the summed table in d2_total (device, n*n float64) for callers
 *                        that run their own selection (Bulyan). */
int afl_xgpu_create(int world, int rank, int n_max, void** ctx_out);
int afl_xgpu_handle(void* ctx, unsigned char* out64);
int afl_xgpu_connect(void* ctx, const unsigned char* handles /* world x 64 bytes, rank order */);
int afl_xgpu_destroy(void* ctx);
int afl_krum_sharded(void* ctx, const void* G, int n, int64_t d, int64_t ld, int dtype, int users_count, int corrupted_count,
                     void* workspace, size_t workspace_bytes, int flags, void* stream, int** idx_host_out,
                     int** status_host_out, int** idx_dev_out);
int afl_sqdist_allreduce(void* ctx, const void* G, int n, int64_t d, int64_t ld, int dtype, double* d2_total, void* workspace,
                         size_t workspace_bytes, int flags, void* stream, int** status_host_out);

/* ---- one-call host-buffer API (what a cgo/ctypes binding of server.py:87 would call) ----------
 * rule: "NoDefense" | "Krum" | "TrimmedMean" | "Bulyan" (defences.py:4-8).  G_host: n x d fp32 in
 * host memory, pitch ld.  The call stages column slabs through device memory (H2D copies overlap the
 * kernels), runs the rule, and writes the aggregated gradient to out_host[d] (fp32) and, for Krum,
 * the winning index to *idx_out (may be NULL).  Blocking.  Reference assert failures map to
 * AFL_ERR_PRECONDITION.  `slab_cols` = columns per staging slab (0 = default). */
int afl_defend_host(const char* rule, const float* G_host, int n, int64_t d, int64_t ld,
                    int users_count, int corrupted_count, float* out_host, int* idx_out,
                    int64_t slab_cols);

#ifdef __cplusplus
}
#endif
#endif /* AFL_B200_H_ */
