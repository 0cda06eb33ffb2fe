// Device replay of the reference's corruption stream (rng_mode = 'mt19937_device'): the keep / drop decision of every
// stored positive of every sampled user, bit-exact with DRecPy/Recommender/cdae.py:63-64, which draws
// rng.uniform(0, 1) once per ITEM (n_items draws = 2 * n_items MT19937 outputs per sampled user) from the one
// sequential random.Random of recommender_abc.py:74 -- 2.2e8 outputs per 4096-user step at the ml-20m shape.
//
// MT19937 is linear over GF(2) (mt_jump.cpp): the 624-word window of the stream that starts J outputs ahead is
//   W_J[w] = XOR_{i : g_J[i] = 1} y[i + w],   g_J(t) = t^J mod phi(t),
// a combination of the first 19937 + 623 untempered words y of the stream.  One CTA per segment of `ups` consecutive
// batch rows: it expands y from the current window (91 rounds of 227 independent words), applies its jump polynomial
// (about 10,000 set bits x 624 words, from shared memory), and then simply runs the generator over its segment,
// regenerating 624 words at a time (three rounds of <= 227 independent words, double buffered).  Every thread owns a
// few of the segment's positives (CSR entries in order) and picks their two output words when their block comes by:
// random() = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 < q  <=>  the 53-bit integer < ceil(q * 2^53) (exact, no doubles).
// The last CTA applies the polynomial of the whole step and writes the window the next step starts from.
#include "kernels.h"

namespace {

constexpr int MTN = 624, MTM = 397, MT_NBITS = 19937, MT_PW = 312;
constexpr int MT_NX = MT_NBITS + MTN;            // words of y a jump reads
constexpr int kMtThreads = 256, kMtMaxUps = 64;

__device__ __forceinline__ uint32_t mt_next(uint32_t a, uint32_t b, uint32_t c) {   // y[k+624] from y[k], y[k+1], y[k+397]
  const uint32_t v = (a & 0x80000000u) | (b & 0x7fffffffu);
  return c ^ (v >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

__global__ void __launch_bounds__(kMtThreads) k_mt_keep(MtKeepArgs a) {
  extern __shared__ uint32_t sm[];               // y[MT_NX] during the jump; afterwards two 624-word windows
  __shared__ int s_pref[kMtMaxUps + 1];          // entries before each user of the segment
  __shared__ int64_t s_lo[kMtMaxUps];            // CSR start of each user's row
  const int tid = threadIdx.x;
  const int c = blockIdx.x;
  const bool advance = c == a.n_cta;             // the extra CTA: window of the next step
  // 1. expand the stream from the current window
  for (int i = tid; i < MTN; i += kMtThreads) sm[i] = a.window_in[i];
  __syncthreads();
  for (int base = 0; base + MTN < MT_NX; base += 227) {
    const int k = base + tid;
    if (tid < 227 && k + MTN < MT_NX) sm[k + MTN] = mt_next(sm[k], sm[k + 1], sm[k + MTM]);
    __syncthreads();
  }
  // 2. this CTA's window: no jump for segment 0, polys[c - 1] for segment c, poly_total for the advance CTA
  uint32_t w0 = 0, w1 = 0, w2 = 0;
  const int i0 = tid, i1 = tid + kMtThreads, i2 = tid + 2 * kMtThreads;   // i2 < 624 only for tid < 112
  if (c == 0) {
    w0 = sm[i0]; w1 = sm[i1]; w2 = i2 < MTN ? sm[i2] : 0u;
  } else {
    const uint64_t* g = advance ? a.poly_total : a.polys + (int64_t)(c - 1) * MT_PW;
    for (int wi = 0; wi < MT_PW; wi++) {
      uint64_t bits = __ldg(g + wi);
      while (bits) {
        const int i = wi * 64 + __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        w0 ^= sm[i + i0];
        w1 ^= sm[i + i1];
        if (i2 < MTN) w2 ^= sm[i + i2];
      }
    }
  }
  __syncthreads();                               // everybody is done reading y
  if (advance) {
    a.window_out[i0] = w0; a.window_out[i1] = w1;
    if (i2 < MTN) a.window_out[i2] = w2;
    return;
  }
  uint32_t* cur = sm;
  uint32_t* nxt = sm + MTN;
  cur[i0] = w0; cur[i1] = w1;
  if (i2 < MTN) cur[i2] = w2;
  // 3. the positives of this segment's users, owned round-robin by the threads in stream order
  const int b_lo = c * a.ups, b_hi = min(a.batch, b_lo + a.ups);
  if (tid == 0) {
    int acc = 0;
    for (int b = b_lo; b < b_hi; b++) {
      const int row = a.uids[b];
      s_pref[b - b_lo] = acc;
      s_lo[b - b_lo] = a.indptr[row];
      acc += (int)(a.indptr[row + 1] - a.indptr[row]);
    }
    s_pref[b_hi - b_lo] = acc;
  }
  __syncthreads();
  const int n_users = b_hi - b_lo, n_ent = s_pref[n_users];
  int e = tid, ub = 0;                            // my current entry and its user (relative)
  int64_t g_e = -1;                               // its draw index inside the segment: ub * n_items + item
  int64_t dst = 0;
  auto load_entry = [&]() {
    if (e >= n_ent) { g_e = INT64_MAX; return; }
    while (s_pref[ub + 1] <= e) ub++;
    const int j = e - s_pref[ub];
    g_e = (int64_t)ub * a.n_items + a.indices[s_lo[ub] + j];
    dst = (int64_t)a.keep_off[b_lo + ub] + j;
  };
  load_entry();
  const int64_t n_draws = (int64_t)n_users * a.n_items;
  for (int64_t d0 = 0; d0 < n_draws; d0 += MTN / 2) {
    // draws [d0, d0 + 312) are the output pairs (cur[2t], cur[2t + 1])
    while (g_e < d0 + MTN / 2) {
      const int t = (int)(g_e - d0);
      const uint64_t hi = mt_temper(cur[2 * t]) >> 5, lo = mt_temper(cur[2 * t + 1]) >> 6;
      a.keep[dst] = (((hi << 26) | lo) < a.threshold) ? 0 : 1;      // u < q: dropped
      e += kMtThreads;
      load_entry();
    }
    if (d0 + MTN / 2 >= n_draws) break;
    // next 624 words: three rounds of independent words, new values into the other buffer
    if (tid < 227) nxt[tid] = mt_next(cur[tid], cur[tid + 1], cur[tid + MTM]);
    __syncthreads();
    if (tid < 227) nxt[tid + 227] = mt_next(cur[tid + 227], cur[tid + 228], nxt[tid]);
    __syncthreads();
    if (tid < 170) {
      const int k = tid + 454;
      nxt[k] = mt_next(cur[k], k + 1 < MTN ? cur[k + 1] : nxt[0], nxt[k - 227]);
    }
    __syncthreads();
    uint32_t* t_ = cur; cur = nxt; nxt = t_;
  }
}

}  // namespace

int launch_mt_keep(drb_ctx* ctx, const MtKeepArgs& a) {
  if (a.batch <= 0) return DRB_OK;
  if (a.ups < 1 || a.ups > kMtMaxUps || a.n_cta != (a.batch + a.ups - 1) / a.ups)
    return drb_fail(DRB_E_INVALID, "mt_keep: bad segmentation (ups %d, n_cta %d, batch %d)", a.ups, a.n_cta, a.batch);
  const int smem = MT_NX * 4 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mt_keep, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return drb_fail(DRB_E_CUDA, "cudaFuncSetAttribute(k_mt_keep) failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  drb_prof_scope prof_(ctx, "k_mt_keep");
  k_mt_keep<<<a.n_cta + 1, kMtThreads, smem, ctx->stream>>>(a);
  DRB_LAUNCH_CHECK(ctx, "k_mt_keep");
  return DRB_OK;
}
