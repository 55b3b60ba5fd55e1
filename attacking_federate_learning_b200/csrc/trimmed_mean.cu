// Coordinate-wise trimmed mean around the median (reference: defences.py:44-52) — also Bulyan's
// second stage (defences.py:70) through `row_index`.
//
// Per column:  med = median of the N values (even N: fl32 mean of the two middle ones);
//              dev = fl32(x - med); keep the k devs of smallest |dev|, ties in client (row) order;
//              out = fl32(fl32(sum(kept)/k) + med).
//
// Layout / mapping
//   * A CTA owns a tile of 64 bytes per row (16 fp32 or 32 bf16 columns) x all rows.  Rows are read
//     with coalesced 16-byte loads (4 lanes per row segment) and scattered into shared memory in a
//     [word-column][slot-group][lane] order with an XOR on the slot-in-group index, so that both the
//     staging stores (STS.32) and the per-column reads (LDS.128) are bank-conflict free.
//   * One warp then owns one word-column: lane l holds rows l, l+32, l+64, ... in registers
//     (S slots per lane, S = 4, 8, 12, ..., 32: the smallest multiple of 4 with 32 S >= rows), so every pass over
//     the column is pure register arithmetic plus one warp reduction.  Rows past N are staged as +inf.
//   * Selection, fast path (select_fast): ONE fused pass per order statistic with a bracket [a, b) aimed from the
//     column's mean / sigma (median) or sigma and the normal quantile (|dev| threshold): it counts #{key < a}, sums
//     the devs below a, and marks the in-bracket slots in a per-lane bit mask (4-5 instructions per value, in PTX).
//     When the target rank is inside and <= 32 slots are marked, the candidates are re-read from the tile by slot
//     index, compacted to one per lane (shuffle scan) and sorted with a 15-stage shuffle bitonic network; otherwise
//     the bracket is re-aimed from the measured counts (about one column in four needs a second pass).
//   * Selection, general path (warp_select): interpolation search on counts with a min/max bisection fallback that
//     terminates for any data (heavy ties, non-Gaussian columns); brackets of <= 32 elements are compacted by ballot
//     and ranked on (key, row), which makes the reference's stable tie rule exact.  In the fast path a tie group that
//     the keep boundary cuts (ALIE's f identical rows, bf16 value collisions) is resolved in row order with ballots on
//     the register-resident column (tie_sum).
//   * More than 1024 rows: trimmed_mean_large_kernel (shared-memory strip, bisection on the integer image of the keys).
#include "afl_common.cuh"

namespace afl {
namespace tmean {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kWordCols = 16;             // 32-bit words per staged row (64 bytes)
constexpr float kInf = __builtin_huge_valf();

struct Params {
  const void* G;
  const int* row_index;   // may be null
  float* out;
  int64_t d, ld;
  int n_rows;             // participating rows
  int n_total;            // rows of G (bounds for row_index)
  int keep;               // effective number of kept devs (python slice semantics applied), >= 0
  float med_density;      // 0.39894228 * n      (ranks per unit value at the centre of a unit Gaussian)
  float key_q;            // Gaussian guess of the |dev| threshold in sigmas
  float key_density;      // 2 * phi(key_q) * n  (ranks per unit |dev| at that threshold, unit sigma)
  int vec_ok;
};

__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Monotone map float -> uint32 (total order incl. negatives; NaN sorts last for positive payloads).
__device__ __forceinline__ uint32_t ord_bits(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// true row of register index ri (registers hold each 4-slot group permuted by jx, see staging)
__device__ __forceinline__ int row_of(int ri, int jx, int lane) { return (((ri & ~3) | ((ri & 3) ^ jx)) << 5) + lane; }

template <bool KEYS> __device__ __forceinline__ float keyof(float v) { return KEYS ? fabsf(v) : v; }

// One counting pass: c = #{key(v) < p} (warp total); for KEYS also s = this LANE's sum of v over those elements.
template <int S, bool KEYS>
__device__ __forceinline__ void count_pass(const float (&v)[S], float p, int& c, float& s) {
  int cc = 0;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    const bool b = keyof<KEYS>(v[i]) < p;
    cc += b ? 1 : 0;
    if (KEYS) ss += b ? v[i] : 0.f;
  }
  c = warp_sum_i(cc);
  if (KEYS) s = ss;                 // lane-local partial; reduced once at the very end
}

template <int S>
__device__ __forceinline__ float tie_sum(const float (&v)[S], float T, int need, int jx, int lane);

// Selection over a register-resident column (general path: any data, any bracket state).
//   KEYS = false : returns the order statistics of ranks r1 <= r2 (r2 <= r1 + 1) in a, b.
//   KEYS = true  : v holds devs, key = |dev|; r1 == r2 == keep-1; returns in `a` the sum of the `keep`
//                  devs of smallest key with ties resolved in row order.
template <int S, bool KEYS>
__device__ __forceinline__ void warp_select(const float (&v)[S], int n, int r1, int r2, float p0, float density,
                                            int lane, int jx, uint32_t* scratch, float lo, float hi, int c_lo, int c_hi,
                                            float sum_lo, float& a, float& b) {
  bool model = (density > 0.f) && (density < kInf) && (p0 == p0) && (fabsf(p0) < kInf) && (lo == -kInf) && (hi == kInf);
  float p = p0;
  a = b = __int_as_float(0x7fc00000);
  for (int iter = 0; iter < 96; ++iter) {
    const int inb = c_hi - c_lo;
    if (inb <= 32) {
      // ---- compact the bracket into this warp's smem scratch (ballot prefix), then rank every
      // candidate by counting (key, row) pairs below it: independent broadcast loads, no shuffles.
      unsigned long long* sk = reinterpret_cast<unsigned long long*>(scratch);      // [32] (ord(key) << 32) | row
      float* sp = reinterpret_cast<float*>(scratch + 64);                            // [32] payload
      sk[lane] = ~0ull;
      sp[lane] = 0.f;
      __syncwarp();
      int base = 0;
#pragma unroll
      for (int i = 0; i < S; ++i) {
        const float k = keyof<KEYS>(v[i]);
        const bool in = (k >= lo) && (k < hi);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (m) {
          if (in) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            sk[pos] = (static_cast<unsigned long long>(ord_bits(k)) << 32) | static_cast<unsigned>(row_of(i, jx, lane));
            sp[pos] = v[i];
          }
          base += __popc(m);
        }
      }
      __syncwarp();
      const unsigned long long mine = sk[lane];
      const float mypay = sp[lane];
      int rank = 0;
#pragma unroll
      for (int t = 0; t < 32; ++t) rank += (sk[t] < mine) ? 1 : 0;
      __syncwarp();
      if (!KEYS) {
        const unsigned ma = __ballot_sync(0xffffffffu, rank == r1 - c_lo);
        const unsigned mb = __ballot_sync(0xffffffffu, rank == r2 - c_lo);
        a = __shfl_sync(0xffffffffu, mypay, __ffs(ma) - 1);
        b = __shfl_sync(0xffffffffu, mypay, __ffs(mb) - 1);
      } else {
        const int take = r1 + 1 - c_lo;             // kept candidates = first `take` in (key,row) order
        a = warp_sum(sum_lo + ((rank < take) ? mypay : 0.f));
      }
      return;
    }
    if (iter >= 4 || !model) {
      // ---- fallback: bisect between the actual extremes of the bracket (guaranteed progress)
      float vmin = kInf, vmax = -kInf;
#pragma unroll
      for (int i = 0; i < S; ++i) {
        const float k = keyof<KEYS>(v[i]);
        const bool in = (k >= lo) && (k < hi);
        vmin = in ? fminf(vmin, k) : vmin;
        vmax = in ? fmaxf(vmax, k) : vmax;
      }
      vmin = warp_min_f(vmin);
      vmax = warp_max_f(vmax);
      if (!(vmin < vmax)) {
        // every element of the bracket has the same key (a tie group wider than a warp)
        if (!KEYS) { a = b = vmin; return; }
        const float part = tie_sum<S>(v, vmin, r1 + 1 - c_lo, jx, lane);
        a = warp_sum(sum_lo + part);
        return;
      }
      p = 0.5f * vmin + 0.5f * vmax;
      if (!(p > vmin)) p = vmax;
      model = false;
    }
    int c;
    float s = 0.f;
    count_pass<S, KEYS>(v, p, c, s);
    if (c <= r1) { lo = p; c_lo = c; sum_lo = s; }
    else if (c > r2) { hi = p; c_hi = c; }
    else {
      // p separates the two middle order statistics (even N median only)
      float below = -kInf, above = kInf;
#pragma unroll
      for (int i = 0; i < S; ++i) {
        below = (v[i] < p) ? fmaxf(below, v[i]) : below;
        above = (v[i] >= p) ? fminf(above, v[i]) : above;
      }
      a = warp_max_f(below);
      b = warp_min_f(above);
      return;
    }
    if (model) {
      // aim just past the target on the far side so that the next pass closes the bracket
      float step;
      if (c <= r1) { const float gap = static_cast<float>(r2 + 1 - c); step = (gap + 3.f + 0.25f * gap) / density; }
      else { const float gap = static_cast<float>(c - r1); step = -(gap + 3.f + 0.25f * gap) / density; }
      float np = p + step;
      if (!(np > lo && np < hi)) {
        if (lo > -kInf && hi < kInf) np = 0.5f * lo + 0.5f * hi;
        if (!(np > lo && np < hi)) model = false;
      }
      p = np;
    }
  }
}

// Sum, in row order, of the first `need` devs among the elements whose |dev| == T (a tie group that the
// keep boundary cuts through).  Uses ballots per slot; registers hold each 4-slot group permuted by jx.
template <int S>
__device__ __forceinline__ float tie_sum(const float (&v)[S], float T, int need, int jx, int lane) {
  float part = 0.f;
  int before = 0;
#pragma unroll
  for (int g = 0; g < S; g += 4) {
    unsigned bm[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) bm[q] = __ballot_sync(0xffffffffu, fabsf(v[g + q]) == T);
    if ((bm[0] | bm[1] | bm[2] | bm[3]) == 0u) continue;       // no member of the tie group in these 128 rows (warp-uniform)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {           // true slot order inside the group
      const int q = kk ^ jx;
      const unsigned m = q == 0 ? bm[0] : q == 1 ? bm[1] : q == 2 ? bm[2] : bm[3];
      const float val = q == 0 ? v[g] : q == 1 ? v[g + 1] : q == 2 ? v[g + 2] : v[g + 3];
      if ((m >> lane) & 1u) {
        const int rank = before + __popc(m & ((1u << lane) - 1u));
        if (rank < need) part += val;
      }
      before += __popc(m);
    }
  }
  return part;
}

// Ascending 32-lane bitonic sort of one float per lane by keyof<KEYS>() (|v| for KEYS); lanes with equal keys keep
// their own value, so nothing is duplicated or lost.  For KEYS the value is rotated left by one bit first: the
// 31 magnitude bits become the high bits and the sign the lowest, so an unsigned integer min/max orders by |v|
// (then sign) and carries the whole value along: 4 instructions per stage instead of 6 with float compares.
template <bool KEYS>
__device__ __forceinline__ float warp_sort32(float v, int lane) {
  uint32_t u = KEYS ? __funnelshift_l(__float_as_uint(v), __float_as_uint(v), 1) : 0u;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const bool up = (k == 32) ? true : ((lane & k) == 0);
      const bool take_min = ((lane & j) == 0) == up;
      if (KEYS) {
        const uint32_t p = __shfl_xor_sync(0xffffffffu, u, j);
        u = take_min ? min(u, p) : max(u, p);
      } else {
        const float p = __shfl_xor_sync(0xffffffffu, v, j);
        v = take_min ? fminf(v, p) : fmaxf(v, p);
      }
    }
  }
  return KEYS ? __uint_as_float(__funnelshift_r(u, u, 1)) : v;
}

// Where a register-resident column lives in the shared-memory tile, so that a candidate can be re-read by a
// run-time register index (a dynamic index into the register array itself would spill it to local memory).
struct ColRef {
  const uint32_t* words;     // tile word of element 0 for this lane; element i is words[(i >> 2) * 128 + (i & 3)]
  uint32_t unpack_sel;       // bf16: PRMT selector that moves the column's half of the word to the top (0x1044 / 0x3244)
  float med;                 // KEYS: the median that was subtracted from the registers
};
template <bool BF16, bool KEYS>
__device__ __forceinline__ float col_fetch(const ColRef& c, int i) {
  const uint32_t w = c.words[(i >> 2) * 128 + (i & 3)];
  const float x = BF16 ? __uint_as_float(__byte_perm(w, 0u, c.unpack_sel)) : __uint_as_float(w);
  return KEYS ? __fsub_rn(x, c.med) : x;
}

// One element of the fused pass: below = key < a; ca += below; (KEYS) sa += below ? v : 0; inmask |= bit when
// a <= key < b.  Written in PTX so that it stays 4 (5) instructions: the C++ form was compiled to 14 per element
// (the count became a set/clear bit mask, and the list address was rebuilt under every store's predicate).
template <bool KEYS>
__device__ __forceinline__ void fused_step(float v, float a, float b, uint32_t bit, int& ca, uint32_t& inmask, float& sa) {
  if (KEYS) {
    asm("{\n\t.reg .pred p, q;\n\t.reg .f32 k;\n\t"
        "abs.f32 k, %3;\n\t"
        "setp.lt.f32 p, k, %4;\n\t"
        "setp.lt.and.f32 q, k, %5, !p;\n\t"
        "@p add.s32 %0, %0, 1;\n\t"
        "@p add.rn.f32 %2, %2, %3;\n\t"
        "@q or.b32 %1, %1, %6;\n\t}"
        : "+r"(ca), "+r"(inmask), "+f"(sa)
        : "f"(v), "f"(a), "f"(b), "r"(bit));
  } else {
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.lt.f32 p, %2, %3;\n\t"
        "setp.lt.and.f32 q, %2, %4, !p;\n\t"
        "@p add.s32 %0, %0, 1;\n\t"
        "@q or.b32 %1, %1, %5;\n\t}"
        : "+r"(ca), "+r"(inmask)
        : "f"(v), "f"(a), "f"(b), "r"(bit));
  }
}

// Fast path.  One fused pass over the registers with a model bracket [a, b): counts #{key < a} (and the
// lane-local dev sum below a) and marks the in-bracket elements in a per-lane bit mask (no stores, no ballots,
// no branches).  If the target rank(s) fall inside and <= 32 candidates were marked, each lane re-reads its
// few candidates from the tile into a dense 32-entry list (exclusive prefix of the per-lane counts), and the
// list is sorted with a 15-stage shuffle network.  Returns false (state updated: lo/c_lo/sum_lo or hi/c_hi
// tightened where the pass proved a bound) when the general path has to take over.
template <int S, bool KEYS, bool BF16>
__device__ __forceinline__ bool select_fast(const float (&v)[S], const ColRef& col, int n, int r1, int r2, float p0,
                                            float density, int lane, int jx, uint32_t* scratch, float& lo, float& hi,
                                            int& c_lo, int& c_hi, float& sum_lo, float& out_a, float& out_b) {
  static_assert(S <= 32, "one mask bit per register-resident element");
  if (!((density > 0.f) && (density < kInf) && (p0 == p0) && (fabsf(p0) < kInf))) return false;
  float center = p0;
  const float inv_density = __fdividef(1.f, density);       // model only: approximate division is enough
  float halfw = (11.f + 0.5f * static_cast<float>(r2 - r1)) * inv_density;
#pragma unroll 1
  for (int attempt = 0; attempt < 3; ++attempt) {
    float a = center - halfw, b = center + halfw;
    if (KEYS && a < 0.f) a = 0.f;
    if (!(a > lo)) a = lo;
    if (!(b < hi)) b = hi;
    if (!(a < b)) return false;
    int ca = 0;
    float sa = 0.f;
    uint32_t inmask = 0u;
#pragma unroll
    for (int i = 0; i < S; ++i) fused_step<KEYS>(v[i], a, b, 1u << i, ca, inmask, sa);
    const int mine = __popc(inmask);
    const int c_a = warp_sum_i(ca);
    const int cin = warp_sum_i(mine);
    const int c_b = c_a + cin;
    const bool inside = (c_a <= r1) && (r2 < c_b);
    if (inside && cin <= 32) {
      // ---- dense compaction: exclusive prefix of the per-lane counts, then one candidate per lane
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      float* dense = reinterpret_cast<float*>(scratch);              // [32]
      int pos = incl - mine;
      for (uint32_t m = inmask; m != 0u; m &= m - 1u) dense[pos++] = col_fetch<BF16, KEYS>(col, __ffs(m) - 1);
      __syncwarp();
      const bool have = lane < cin;
      const float val = have ? dense[lane] : kInf;
      __syncwarp();
      // 15-stage shuffle bitonic sort of the (<= 32) candidates by key; equal keys may end up in any order, which
      // is fine: for the median equal keys are equal values, for the |dev| threshold a tie group that the keep
      // boundary cuts is resolved in row order by tie_sum() on the register-resident column, not on these lanes.
      // (round 1 ranked every candidate against every other: 31 shuffles + 93 ALU ops, 21 % of the kernel.)
      const float sv = warp_sort32<KEYS>(val, lane);
      if (!KEYS) {
        out_a = __shfl_sync(0xffffffffu, sv, r1 - c_a);
        out_b = __shfl_sync(0xffffffffu, sv, r2 - c_a);
        return true;
      }
      const float skey = fabsf(sv);
      const int take = r1 + 1 - c_a;            // number of candidates kept, in (key, row) order
      const float T = __shfl_sync(0xffffffffu, skey, take - 1);                  // the boundary candidate's key
      const int n_less = __popc(__ballot_sync(0xffffffffu, have && skey < T));
      const int group = __popc(__ballot_sync(0xffffffffu, have && skey == T));
      const int need = take - n_less;
      float part = sa + ((have && skey < T) ? sv : 0.f);
      if (need == group) part += (have && skey == T) ? sv : 0.f;    // whole tie group kept: order irrelevant
      else part += tie_sum<S>(v, T, need, jx, lane);                   // boundary cuts the group: row order
      out_a = warp_sum(part);
      return true;
    }
    if (!KEYS && c_a > r1 && c_a <= r2) {
      // the lower pivot separates the two middle order statistics (even N): read them off directly
      float below = -kInf, above = kInf;
#pragma unroll
      for (int i = 0; i < S; ++i) {
        below = (v[i] < a) ? fmaxf(below, v[i]) : below;
        above = (v[i] >= a) ? fminf(above, v[i]) : above;
      }
      out_a = warp_max_f(below);
      out_b = warp_min_f(above);
      return true;
    }
    // ---- not closed: keep what the pass proved (invariants: c_lo <= r1, r2 < c_hi) and aim again
    const float mid = 0.5f * static_cast<float>(r1 + r2) + 0.5f;
    if (c_a <= r1) { lo = a; c_lo = c_a; sum_lo = sa; } else { hi = a; c_hi = c_a; }
    if (c_b > r2 && b < hi) { hi = b; c_hi = c_b; }
    if (inside) {                                   // too many candidates: shrink around the interpolated rank
      const float w = __fdividef(b - a, static_cast<float>(cin > 0 ? cin : 1));
      center = a + (mid - static_cast<float>(c_a)) * w;
      halfw = 9.f * w;
    } else if (c_a > r2) {                          // target below the bracket
      center = a - (static_cast<float>(c_a) - mid) * inv_density;
      halfw = (5.f + 0.35f * (static_cast<float>(c_a) - mid)) * inv_density;
    } else {                                        // target above the bracket
      center = b + (mid - static_cast<float>(c_b)) * inv_density;
      halfw = (5.f + 0.35f * (mid - static_cast<float>(c_b))) * inv_density;
    }
  }
  return false;
}

// ---------------- staging: coalesced 16-byte row-segment loads -> swizzled smem ----------------
// Thread t loads the 16-byte chunk j = t&3 of row (it*64 + t>>2) in iteration `it`.  Row r lives in
// slot r>>5 of lane r&31; slot-group m = slot>>2 and the XOR-ed slot position are compile-time
// functions of `it`, so every store below has an immediate offset from one of two per-thread bases.
template <int S, bool BF16>
__device__ __forceinline__ void stage_tile(const Params& P, uint32_t* tile, int64_t col0) {
  constexpr int kGroups = S / 4;
  const int tid = threadIdx.x;
  const int es = BF16 ? 2 : 4;
  const int cols_per_tile = BF16 ? 32 : 16;
  const uint32_t sentinel = BF16 ? 0x7F807F80u : 0x7F800000u;
  const uint8_t* base = static_cast<const uint8_t*>(P.G);
  constexpr int kIters = (32 * S * 4) / kThreads;      // S/2
  const bool full_tile = P.vec_ok && (col0 + cols_per_tile <= P.d);
  const int rowq = tid >> 2, j = tid & 3, l = rowq & 31, hi2 = tid >> 7;
  uint32_t* b0 = tile + (4 * j * kGroups) * 128 + l * 4 + (hi2 ^ j);
  uint32_t* b1 = tile + (4 * j * kGroups) * 128 + l * 4 + ((2 + hi2) ^ j);
  const int64_t c = col0 + static_cast<int64_t>(j) * (16 / es);
  if (full_tile) {
    // every 16-byte load of the tile is issued before the first store: one DRAM round trip per CTA instead of
    // one per batch of four (the kernel has the registers: the per-column code needs 80 anyway), and ~8
    // instructions per load so that the whole path stays a few hundred bytes of code
    // (branch-free: rows past n_rows re-read the last row and are replaced by the sentinel at the store)
    int gr[kIters];
    const int last = P.n_rows - 1;
    if (P.row_index) {
#pragma unroll
      for (int it = 0; it < kIters; ++it) gr[it] = P.row_index[min(it * 64 + rowq, last)];
#pragma unroll
      for (int it = 0; it < kIters; ++it) gr[it] = gr[it] < 0 ? gr[it] + P.n_total : gr[it];
    } else {
#pragma unroll
      for (int it = 0; it < kIters; ++it) gr[it] = min(it * 64 + rowq, last);
    }
    const uint8_t* colbase = base + c * es;
    const int64_t row_bytes = P.ld * es;
    uint4 val[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it)
      val[it] = ldg_stream_u4(reinterpret_cast<const uint4*>(colbase + gr[it] * row_bytes));
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const bool pad = it * 64 + rowq > last;
      uint32_t* b = (it & 1) ? b1 : b0;
      const int m = it >> 1;
      b[(0 * kGroups + m) * 128] = pad ? sentinel : val[it].x;
      b[(1 * kGroups + m) * 128] = pad ? sentinel : val[it].y;
      b[(2 * kGroups + m) * 128] = pad ? sentinel : val[it].z;
      b[(3 * kGroups + m) * 128] = pad ? sentinel : val[it].w;
    }
    return;
  }
  // ragged last tile / unaligned matrix: element-wise, one rolled iteration at a time (cold: keep the code small)
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    const int r = it * 64 + rowq;
    uint32_t w[4] = {sentinel, sentinel, sentinel, sentinel};
    if (r < P.n_rows) {
      int gr = P.row_index ? P.row_index[r] : r;
      gr = gr < 0 ? gr + P.n_total : gr;
      const uint8_t* src = base + (static_cast<int64_t>(gr) * P.ld + c) * es;
      w[0] = w[1] = w[2] = w[3] = 0u;
      if (BF16) {
        const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (c + e < P.d) w[e >> 1] |= static_cast<uint32_t>(s16[e]) << ((e & 1) * 16);
      } else {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c + e < P.d) w[e] = s32[e];
      }
    }
    uint32_t* b = ((it & 1) ? b1 : b0) + (it >> 1) * 128;
    b[(0 * kGroups) * 128] = w[0];
    b[(1 * kGroups) * 128] = w[1];
    b[(2 * kGroups) * 128] = w[2];
    b[(3 * kGroups) * 128] = w[3];
  }
}

// ---------------- general per-column path (any data): one warp, one column ----------------
// `half` selects the bf16 column inside the 32-bit word-column cw (ignored for fp32).
template <int S, bool BF16>
__device__ __forceinline__ float general_column_impl(const Params& P, const uint32_t* tile, int cw, int half,
                                                     uint32_t* scratch, int lane) {
  constexpr int kGroups = S / 4;
  const int n = P.n_rows;
  const float fn = static_cast<float>(n);
  const int jx = (cw >> 2) & 3;
  const uint4* t4 = reinterpret_cast<const uint4*>(tile) + (cw * kGroups) * 32 + lane;
  // bf16 -> fp32 is one PRMT with a run-time selector (`half` is a loop variable: a ?: costs two predicated instructions)
  const uint32_t unpack_sel = half ? 0x3244u : 0x1044u;
  ColRef col{reinterpret_cast<const uint32_t*>(t4), unpack_sel, 0.f};
  float x[S];
#pragma unroll
  for (int m = 0; m < kGroups; ++m) {
    const uint4 t = t4[m * 32];
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      x[4 * m + q] = BF16 ? __uint_as_float(__byte_perm(w[q], 0u, unpack_sel)) : __uint_as_float(w[q]);
  }

  // mean / sigma of the column (pivot model only; never enters the result).  The kernel is instantiated with
  // 16 * S < n_rows <= 32 * S (S >= 8), so the first half of the slot groups holds real rows only; in the second half a
  // padded row is +inf and is skipped by predicate (no branches: the warp-uniform "is this group full" tests of
  // round 1 cost an instruction-fetch bubble each).
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    if (S >= 8 && i < S / 2) {             // the launcher picks S with 32 * 4 * ceil(S / 8) <= n_rows for S >= 8
      s1 += x[i]; s2 = fmaf(x[i], x[i], s2);
    } else {
      asm("{\n\t.reg .pred p;\n\t"
          "setp.lt.f32 p, %2, 0f7F800000;\n\t"
          "@p add.rn.f32 %0, %0, %2;\n\t"
          "@p fma.rn.f32 %1, %2, %2, %1;\n\t}"
          : "+f"(s1), "+f"(s2) : "f"(x[i]));
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  const float mean = __fdividef(s1, fn);
  const float var = fmaxf(__fdividef(s2, fn) - mean * mean, 0.f);
  const float inv_sd = rsqrtf(var);              // var == 0: +inf -> the densities below fail the model test
  const float sd = var * inv_sd;                 // (NaN for var == 0: same effect on the |dev| pivot)

  // median (np.median: even N -> mean of the two middle order statistics, fp32)
  float a, b;
  {
    float lo = -kInf, hi = kInf, sl = 0.f;
    int c_lo = 0, c_hi = n;
    if (!select_fast<S, false, BF16>(x, col, n, (n - 1) >> 1, n >> 1, mean, P.med_density * inv_sd, lane, jx, scratch, lo, hi, c_lo,
                               c_hi, sl, a, b))
      warp_select<S, false>(x, n, (n - 1) >> 1, n >> 1, mean, P.med_density * inv_sd, lane, jx, scratch, lo, hi, c_lo,
                            c_hi, 0.f, a, b);
  }
  const float med = ((n & 1) != 0) ? a : __fdiv_rn(__fadd_rn(a, b), 2.0f);

  float res;
  if (P.keep <= 0) {
    res = __int_as_float(0x7fc00000);           // np.mean([]) is nan
  } else {
#pragma unroll
    for (int i = 0; i < S; ++i) x[i] = __fsub_rn(x[i], med);      // devs; padded rows stay +inf
    float total, unused;
    float lo = -kInf, hi = kInf, sl = 0.f;
    int c_lo = 0, c_hi = n;
    col.med = med;
    if (!select_fast<S, true, BF16>(x, col, n, P.keep - 1, P.keep - 1, P.key_q * sd, P.key_density * inv_sd, lane, jx, scratch, lo, hi,
                              c_lo, c_hi, sl, total, unused))
      warp_select<S, true>(x, n, P.keep - 1, P.keep - 1, P.key_q * sd, P.key_density * inv_sd, lane, jx, scratch, lo, hi,
                           c_lo, c_hi, sl, total, unused);
    res = __fadd_rn(__fdiv_rn(total, static_cast<float>(P.keep)), med);
  }
  return res;
}

constexpr int kScratchWords = 96;              // per warp: dense candidate list [32] (fast path) / (key,row) u64[32] + payload[32] (general path)

// S <= 20 (up to 640 rows: Bulyan's second stage at N = 500 and N = 1000) leaves room for four CTAs per SM in shared memory; ask
// the compiler for 64 registers there (resident warps are what hides the shuffle chains of the scans and sorts)
template <int S, bool BF16>
__global__ void __launch_bounds__(kThreads, (S <= 20 ? 4 : 3))      // (S = 24 fits 4 CTAs in shared memory too, but the fp32 instance spills at 64 registers)
trimmed_mean_kernel(const Params P) {
  extern __shared__ __align__(1024) uint32_t tile[];     // [16 word-cols][S/4 groups][32 lanes][4 slots] + scratch
  constexpr int kGroups = S / 4;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int cols_per_tile = BF16 ? 32 : 16;
  const int64_t col0 = static_cast<int64_t>(blockIdx.x) * cols_per_tile;
  stage_tile<S, BF16>(P, tile, col0);
  // read once, after staging: `volatile` keeps ptxas from re-reading the special register (S2R, ~50 cycles of
  // latency) in front of every scan and sort of the per-column code to save one register
  int lane = tid & 31, warp_o = warp;
  if (BF16 || S < 32) {       // (the fp32 S = 32 instance is at the 80-register limit: there it would only add spills)
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("mov.u32 %0, %1;" : "=r"(warp_o) : "r"(warp));
  }
  __syncthreads();
  uint32_t* scratch = tile + kWordCols * kGroups * 128 + warp_o * kScratchWords;
#pragma unroll 1
  for (int cw = warp_o; cw < kWordCols; cw += kWarps) {
#pragma unroll 1
    for (int half = 0; half < (BF16 ? 2 : 1); ++half) {
      const int64_t col = col0 + (BF16 ? 2 * cw + half : cw);
      if (col >= P.d) break;                         // warp-uniform
      const float res = general_column_impl<S, BF16>(P, tile, cw, half, scratch, lane);
      if (lane == 0) P.out[col] = res;
    }
  }
}

// ---------------- more than 1024 participating rows: shared-memory bisection kernel (any n that fits) ----------------
// The register-resident kernels above hold ceil(n/32) <= 32 values per lane.  Beyond that a CTA stages a 16-byte-wide
// column strip of ALL rows in shared memory ([row][4 words], up to kLargeMaxRows rows) and one warp per column finds
// the order statistics by bisection on the order-preserving integer image of the values (32 counting passes over
// shared memory per statistic), then the |dev| threshold the same way, and resolves a tie group at the threshold in row
// order with ballots.  Exact for any data; slow (a fallback: the reference has no client-count limit, defences.py:44-52).
constexpr int kLargeMaxRows = 12288;          // 12288 rows x 16 B = 192 KB of shared memory

__device__ __forceinline__ float strip_value(const uint32_t* strip, int row, int col_in_strip, bool bf16) {
  if (!bf16) return __uint_as_float(strip[row * 4 + col_in_strip]);
  const uint32_t w = strip[row * 4 + (col_in_strip >> 1)];
  return __uint_as_float((col_in_strip & 1) ? (w & 0xFFFF0000u) : (w << 16));
}

// k-th smallest (0-based) key among this warp's column, keys = ord_bits(f(value)); F = 0: value, F = 1: |fl32(value - med)|
template <int F>
__device__ __forceinline__ uint32_t warp_kth_key(const uint32_t* strip, int n, int cis, bool bf16, float med, int k, int lane) {
  uint32_t lo = 0u, hi = 0xFFFFFFFFu;                      // invariant: count(key < lo) <= k < count(key <= hi)
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    int c = 0;
    for (int r = lane; r < n; r += 32) {
      float v = strip_value(strip, r, cis, bf16);
      if (F) v = fabsf(__fsub_rn(v, med));
      c += (ord_bits(v) <= mid) ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (c > k) hi = mid; else lo = mid + 1;
  }
  return lo;
}
__device__ __forceinline__ float from_ord_bits(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

__global__ void __launch_bounds__(256, 1)
trimmed_mean_large_kernel(const Params P, int bf16) {
  extern __shared__ __align__(16) uint32_t strip[];        // [n_rows][4 words]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = P.n_rows;
  const int es = bf16 ? 2 : 4;
  const int cols = bf16 ? 8 : 4;
  const int64_t col0 = static_cast<int64_t>(blockIdx.x) * cols;
  const uint8_t* base = static_cast<const uint8_t*>(P.G);
  for (int r = tid; r < n; r += blockDim.x) {
    int gr = P.row_index ? P.row_index[r] : r;
    gr = gr < 0 ? gr + P.n_total : gr;
    const uint8_t* src = base + (static_cast<int64_t>(gr) * P.ld + col0) * es;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (P.vec_ok && col0 + cols <= P.d) {
      const uint4 t = ldg_stream_u4(reinterpret_cast<const uint4*>(src));
      w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
    } else if (bf16) {
      const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src);
      for (int e = 0; e < 8; ++e)
        if (col0 + e < P.d) w[e >> 1] |= static_cast<uint32_t>(s16[e]) << ((e & 1) * 16);
    } else {
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
      for (int e = 0; e < 4; ++e)
        if (col0 + e < P.d) w[e] = s32[e];
    }
    *reinterpret_cast<uint4*>(strip + r * 4) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  for (int cis = warp; cis < cols; cis += 8) {
    const int64_t col = col0 + cis;
    if (col >= P.d) break;
    const float a = from_ord_bits(warp_kth_key<0>(strip, n, cis, bf16 != 0, 0.f, (n - 1) >> 1, lane));
    const float b = (n & 1) ? a : from_ord_bits(warp_kth_key<0>(strip, n, cis, bf16 != 0, 0.f, n >> 1, lane));
    const float med = (n & 1) ? a : __fdiv_rn(__fadd_rn(a, b), 2.0f);
    float res;
    if (P.keep <= 0) {
      res = __int_as_float(0x7fc00000);
    } else {
      const uint32_t To = warp_kth_key<1>(strip, n, cis, bf16 != 0, med, P.keep - 1, lane);   // threshold key (as ord bits)
      // sum of the devs strictly below the threshold, then the first `need` of the tie group in row order
      float part = 0.f;
      int below = 0;
      for (int r = lane; r < n; r += 32) {
        const float dv = __fsub_rn(strip_value(strip, r, cis, bf16 != 0), med);
        const bool lt = ord_bits(fabsf(dv)) < To;
        part += lt ? dv : 0.f;
        below += lt ? 1 : 0;
      }
      below = __reduce_add_sync(0xffffffffu, below);
      int need = P.keep - below;
      for (int r0 = 0; r0 < n && need > 0; r0 += 32) {       // rows in order: r0 .. r0+31 <-> lanes 0..31
        const int r = r0 + lane;
        float dv = 0.f;
        bool tie = false;
        if (r < n) { dv = __fsub_rn(strip_value(strip, r, cis, bf16 != 0), med); tie = ord_bits(fabsf(dv)) == To; }
        const unsigned m = __ballot_sync(0xffffffffu, tie);
        if (tie && __popc(m & ((1u << lane) - 1u)) < need) part += dv;
        need -= __popc(m);
      }
      const float total = warp_sum(part);
      res = __fadd_rn(__fdiv_rn(total, static_cast<float>(P.keep)), med);
    }
    if (lane == 0) P.out[col] = res;
  }
}

static double norm_ppf(double pr) {   // Acklam's rational approximation, |error| < 1.2e-9
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                             6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double dd[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                              3.754408661907416e+00};
  if (pr <= 0.0) return -8.0;
  if (pr >= 1.0) return 8.0;
  if (pr < 0.02425) {
    const double q = sqrt(-2.0 * log(pr));
    return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((dd[0] * q + dd[1]) * q + dd[2]) * q + dd[3]) * q + 1.0);
  }
  if (pr > 1.0 - 0.02425) {
    const double q = sqrt(-2.0 * log(1.0 - pr));
    return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((dd[0] * q + dd[1]) * q + dd[2]) * q + dd[3]) * q + 1.0);
  }
  const double q = pr - 0.5, r = q * q;
  return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
         (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
}

template <int S>
static int launch(const Params& P, int dtype, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(S) * 2048 + kWarps * kScratchWords * 4;
  const int cols = dtype == AFL_BF16 ? 32 : 16;
  const unsigned grid = static_cast<unsigned>(ceil_div64(P.d, cols));
  ProfScope ps("trimmed_mean", stream);
  if (dtype == AFL_BF16) {
    AFL_CUDA(cudaFuncSetAttribute(trimmed_mean_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    trimmed_mean_kernel<S, true><<<grid, kThreads, smem, stream>>>(P);
  } else {
    AFL_CUDA(cudaFuncSetAttribute(trimmed_mean_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    trimmed_mean_kernel<S, false><<<grid, kThreads, smem, stream>>>(P);
  }
  AFL_LAUNCH_CHECK("trimmed_mean_kernel");
  return AFL_OK;
}

int trimmed_mean(const void* G, int n, int64_t d, int64_t ld, int dtype, const int* row_index, int n_rows,
                 int corrupted_count, float* out, cudaStream_t stream) {
  if (!G || !out || n < 1 || d < 1 || ld < d || n_rows < 1) { set_error("afl_trimmed_mean: bad argument"); return AFL_ERR_BAD_ARG; }
  if (dtype != AFL_F32 && dtype != AFL_BF16) { set_error("afl_trimmed_mean: dtype"); return AFL_ERR_UNSUPPORTED; }
  if (n_rows > kLargeMaxRows) {
    set_error("afl_trimmed_mean: at most %d participating rows fit the shared-memory strip (got %d)", kLargeMaxRows, n_rows);
    return AFL_ERR_UNSUPPORTED;
  }
  // number_to_consider = rows - f - 1, then Python slice semantics for sorted(...)[:k]   (defences.py:45,50)
  const int k = n_rows - corrupted_count - 1;
  const int keep = k >= 0 ? (k < n_rows ? k : n_rows) : (n_rows + k > 0 ? n_rows + k : 0);
  Params P{};
  P.G = G; P.row_index = row_index; P.out = out; P.d = d; P.ld = ld; P.n_rows = n_rows; P.n_total = n; P.keep = keep;
  P.med_density = 0.3989422804f * static_cast<float>(n_rows);
  const double frac = keep > 0 ? (static_cast<double>(keep) - 0.5) / n_rows : 0.5;
  const double q = norm_ppf(0.5 * (1.0 + (frac < 0.999999 ? frac : 0.999999)));
  P.key_q = static_cast<float>(q);
  P.key_density = static_cast<float>(2.0 * 0.3989422804014327 * exp(-0.5 * q * q) * n_rows);
  const int64_t es = dtype == AFL_F32 ? 4 : 2;
  P.vec_ok = (reinterpret_cast<uintptr_t>(G) % 16 == 0) && ((ld * es) % 16 == 0);
  if (n_rows > 1024) {
    const size_t smem = static_cast<size_t>(n_rows) * 16;
    const int cols = dtype == AFL_BF16 ? 8 : 4;
    static int smem_attr_done[kMaxDevices] = {0};
    AFL_CUDA(ensure_dyn_smem(trimmed_mean_large_kernel, static_cast<int>(kLargeMaxRows) * 16, smem_attr_done));
    ProfScope ps("trimmed_mean", stream);
    trimmed_mean_large_kernel<<<static_cast<unsigned>(ceil_div64(d, cols)), 256, smem, stream>>>(P, dtype == AFL_BF16 ? 1 : 0);
    AFL_LAUNCH_CHECK("trimmed_mean_large_kernel");
    return AFL_OK;
  }
  // S = slots per lane, a multiple of 4 with 32 * S >= n_rows: the work per column is proportional to S, not to n_rows
  // (Bulyan's second stage at N = 1000, f = 240 selects 520 rows: S = 20 instead of 32)
  if (n_rows <= 128) return launch<4>(P, dtype, stream);
  if (n_rows <= 256) return launch<8>(P, dtype, stream);
  if (n_rows <= 384) return launch<12>(P, dtype, stream);
  if (n_rows <= 512) return launch<16>(P, dtype, stream);
  if (n_rows <= 640) return launch<20>(P, dtype, stream);
  if (n_rows <= 768) return launch<24>(P, dtype, stream);
  if (n_rows <= 896) return launch<28>(P, dtype, stream);
  return launch<32>(P, dtype, stream);
}

}  // namespace tmean
}  // namespace afl
