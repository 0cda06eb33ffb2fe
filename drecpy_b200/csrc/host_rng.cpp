// Host-side replay of CPython's random.Random (MT19937) and everything on the hot path that is sequential by
// construction: PointSampler, the CDAE corruption stream, ranking_evaluation's per-user candidate generation.
// Reference behaviour: DRecPy/Sampler/point_sampler.py:19-96, DRecPy/Dataset/mem_dataset.py:101-163,
// DRecPy/Recommender/cdae.py:63-64, DRecPy/Evaluation/Processes/ranking_evaluation.py:108-116,163-219.
// The generator is the published MT19937 with CPython's seeding (init_by_array over the 32-bit limbs of abs(seed)).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <cstdint>
#include <thread>
#include <unordered_set>
#include <vector>

#include "drb_internal.h"

struct drb_rng {
  static constexpr int N = 624, M = 397;
  uint32_t mt[N];
  int idx;

  void init_genrand(uint32_t s) {
    mt[0] = s;
    for (int i = 1; i < N; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = N;
  }
  void init_by_array(const uint32_t* key, int len) {
    init_genrand(19650218u);
    int i = 1, j = 0;
    for (int k = (N > len ? N : len); k; k--) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
      i++; j++;
      if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
      if (j >= len) j = 0;
    }
    for (int k = N - 1; k; k--) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
      i++;
      if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
    }
    mt[0] = 0x80000000u;
  }
  void seed(uint64_t a) {  // random.seed(int): key = little-endian 32-bit limbs of abs(seed), at least one
    uint32_t key[2] = {(uint32_t)(a & 0xffffffffu), (uint32_t)(a >> 32)};
    init_by_array(key, key[1] ? 2 : 1);
  }
  void twist() {
    auto mix = [](uint32_t u, uint32_t v) { return (u & 0x80000000u) | (v & 0x7fffffffu); };
    int kk = 0;
    for (; kk < N - M; kk++) {
      uint32_t y = mix(mt[kk], mt[kk + 1]);
      mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < N - 1; kk++) {
      uint32_t y = mix(mt[kk], mt[kk + 1]);
      mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    uint32_t y = mix(mt[N - 1], mt[0]);
    mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    idx = 0;
  }
  static inline uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  inline uint32_t next32() {
    if (idx >= N) twist();
    return temper(mt[idx++]);
  }
  inline void skip(int64_t n) {  // advance the stream by n outputs without tempering
    while (n > 0) {
      if (idx >= N) twist();
      int64_t take = std::min<int64_t>(n, N - idx);
      idx += (int)take;
      n -= take;
    }
  }
  inline double random() {
    uint32_t a = next32() >> 5, b = next32() >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
  }
  inline uint64_t getrandbits(int k) {
    if (k <= 32) return next32() >> (32 - k);
    uint64_t lo = next32();  // little-endian limbs, the last limb truncated
    uint64_t hi = next32() >> (64 - k);
    return lo | (hi << 32);
  }
  inline int64_t randbelow(int64_t n) {
    int k = 64 - __builtin_clzll((unsigned long long)n);  // n.bit_length()
    uint64_t r = getrandbits(k);
    while (r >= (uint64_t)n) r = getrandbits(k);
    return (int64_t)r;
  }
  inline int64_t randint(int64_t a, int64_t b) { return a + randbelow(b - a + 1); }
};

static void sample_indices(drb_rng& g, int64_t n, int64_t k, int64_t* out) {
  // random.sample(range(n), k), CPython 3.12 (pool path vs selection-set path)
  int64_t setsize = 21;
  if (k > 5) setsize += (int64_t)std::pow(4.0, std::ceil(std::log((double)(k * 3)) / std::log(4.0)));
  if (n <= setsize) {
    std::vector<int64_t> pool(n);
    for (int64_t i = 0; i < n; i++) pool[i] = i;
    for (int64_t i = 0; i < k; i++) {
      int64_t j = g.randbelow(n - i);
      out[i] = pool[j];
      pool[j] = pool[n - i - 1];
    }
  } else {
    std::unordered_set<int64_t> selected;
    for (int64_t i = 0; i < k; i++) {
      int64_t j = g.randbelow(n);
      while (selected.count(j)) j = g.randbelow(n);
      selected.insert(j);
      out[i] = j;
    }
  }
}

static void shuffle_i64(drb_rng& g, int64_t n, int64_t* x) {
  for (int64_t i = n - 1; i >= 1; i--) {
    int64_t j = g.randbelow(i + 1);
    std::swap(x[i], x[j]);
  }
}

extern "C" {

int drb_rng_create(uint64_t abs_seed, drb_rng** out) {
  if (!out) return drb_fail(DRB_E_INVALID, "drb_rng_create: out is NULL");
  drb_rng* g = new (std::nothrow) drb_rng;
  if (!g) return drb_fail(DRB_E_NOMEM, "drb_rng_create: out of memory");
  g->seed(abs_seed);
  *out = g;
  return DRB_OK;
}
int drb_rng_destroy(drb_rng* rng) { delete rng; return DRB_OK; }
int drb_rng_seed(drb_rng* rng, uint64_t abs_seed) {
  if (!rng) return drb_fail(DRB_E_INVALID, "drb_rng_seed: rng is NULL");
  rng->seed(abs_seed);
  return DRB_OK;
}
double drb_rng_random(drb_rng* rng) { return rng->random(); }
uint64_t drb_rng_getrandbits(drb_rng* rng, int k) { return (k < 1 || k > 64) ? 0 : rng->getrandbits(k); }
int64_t drb_rng_randbelow(drb_rng* rng, int64_t n) { return n < 1 ? -1 : rng->randbelow(n); }
int drb_rng_random_fill(drb_rng* rng, int64_t n, double* out) {
  if (!rng || (n > 0 && !out)) return drb_fail(DRB_E_INVALID, "drb_rng_random_fill: NULL argument");
  for (int64_t i = 0; i < n; i++) out[i] = rng->random();
  return DRB_OK;
}
int drb_rng_sample_indices(drb_rng* rng, int64_t n, int64_t k, int64_t* out) {
  if (!rng || k < 0 || k > n) return drb_fail(DRB_E_INVALID, "drb_rng_sample_indices: need 0 <= k <= n");
  sample_indices(*rng, n, k, out);
  return DRB_OK;
}
int drb_rng_shuffle_i64(drb_rng* rng, int64_t n, int64_t* x) {
  if (!rng || (n > 0 && !x)) return drb_fail(DRB_E_INVALID, "drb_rng_shuffle_i64: NULL argument");
  shuffle_i64(*rng, n, x);
  return DRB_OK;
}
// The next 624 UNTEMPERED output words of the generator (it is not advanced): with idx = 0 this is the mt[] array
// itself, so (window, idx = 0) is a complete generator state -- the form the device replay of the corruption stream
// (mt_device.cu) keeps.  drb_rng_set_window is the inverse.
int drb_rng_window(const drb_rng* rng, uint32_t window[624]) {
  if (!rng || !window) return drb_fail(DRB_E_INVALID, "drb_rng_window: NULL argument");
  drb_rng c = *rng;
  if (c.idx >= drb_rng::N) c.twist();
  const int head = drb_rng::N - c.idx;
  std::memcpy(window, c.mt + c.idx, (size_t)head * 4);
  if (c.idx > 0) {
    c.twist();
    std::memcpy(window + head, c.mt, (size_t)(drb_rng::N - head) * 4);
  }
  return DRB_OK;
}
int drb_rng_set_window(drb_rng* rng, const uint32_t window[624]) {
  if (!rng || !window) return drb_fail(DRB_E_INVALID, "drb_rng_set_window: NULL argument");
  std::memcpy(rng->mt, window, sizeof(rng->mt));
  rng->idx = 0;
  return DRB_OK;
}
int drb_rng_skip(drb_rng* rng, int64_t n_outputs) {
  if (!rng || n_outputs < 0) return drb_fail(DRB_E_INVALID, "drb_rng_skip: bad argument");
  rng->skip(n_outputs);
  return DRB_OK;
}
int drb_rng_getstate(const drb_rng* rng, uint32_t state[625]) {
  if (!rng || !state) return drb_fail(DRB_E_INVALID, "drb_rng_getstate: NULL argument");
  std::memcpy(state, rng->mt, sizeof(rng->mt));
  state[624] = (uint32_t)rng->idx;
  return DRB_OK;
}
int drb_rng_setstate(drb_rng* rng, const uint32_t state[625]) {
  if (!rng || !state || state[624] > 624) return drb_fail(DRB_E_INVALID, "drb_rng_setstate: bad state");
  std::memcpy(rng->mt, state, sizeof(rng->mt));
  rng->idx = (int)state[624];
  return DRB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ PointSampler
struct drb_sampler {
  int32_t max_uid, max_iid;
  const int64_t* pos_indptr; const int32_t* pos_iid; const double* pos_val;
  const int64_t* all_indptr; const int32_t* all_iid;
  double neg_ratio;
  drb_rng rng, null_rng, pos_rng;  // point_sampler.py:30, mem_dataset.py:135-136, :113-114
  std::vector<int64_t> slot, pos_slot, pos_row, cand_u, cand_i, cand_lo, cand_len;   // scratch of drb_sampler_sample
};

extern "C" {

int drb_sampler_create(int32_t max_uid, int32_t max_iid, const int64_t* pos_indptr, const int32_t* pos_iid,
                       const double* pos_val, const int64_t* all_indptr, const int32_t* all_iid_sorted,
                       double neg_ratio, uint64_t abs_seed, drb_sampler** out) {
  if (!out || !pos_indptr || !all_indptr || max_uid < 0 || max_iid < 0)
    return drb_fail(DRB_E_INVALID, "drb_sampler_create: bad argument");
  if (pos_indptr[(int64_t)max_uid + 1] <= 0)
    return drb_fail(DRB_E_INVALID, "No records were found to sample from.");  // mem_dataset.py:111
  drb_sampler* s = new (std::nothrow) drb_sampler;
  if (!s) return drb_fail(DRB_E_NOMEM, "drb_sampler_create: out of memory");
  s->max_uid = max_uid; s->max_iid = max_iid;
  s->pos_indptr = pos_indptr; s->pos_iid = pos_iid; s->pos_val = pos_val;
  s->all_indptr = all_indptr; s->all_iid = all_iid_sorted;
  s->neg_ratio = neg_ratio;
  s->rng.seed(abs_seed); s->null_rng.seed(abs_seed); s->pos_rng.seed(abs_seed);
  *out = s;
  return DRB_OK;
}
int drb_sampler_destroy(drb_sampler* s) { delete s; return DRB_OK; }

// The reference draws sample after sample (point_sampler.py:44-96), but its three random.Random streams are independent
// objects: `rng` only decides positive / null per sample, `null_rng` only produces candidate (uid, iid) pairs for the null
// samples (mem_dataset.py:158-163: a pair that is stored is simply skipped -- acceptance never changes what is drawn
// next), `pos_rng` only serves the positive samples (mem_dataset.py:119-129).  So each stream can be consumed on its
// own, in its own order, with the same outputs and the same final states -- and the memory lookups, which dominate
// (a random user's CSR row is a cache miss per binary-search level), can be overlapped instead of sitting between two
// draws: the null candidates are drawn in blocks of at most as many pairs as samples are still open (never one more
// than the sequential loop would draw), their rows are prefetched, then tested; the positives' rows are gathered after
// the stream has been walked.
int drb_sampler_sample(drb_sampler* s, int64_t n, int32_t* uid, int32_t* iid, double* val) {
  if (!s || n < 0 || (n > 0 && (!uid || !iid || !val)))
    return drb_fail(DRB_E_INVALID, "drb_sampler_sample: bad argument");
  const double hi = s->neg_ratio + 1.0;
  // 1. the decisions: slot = positions of the null samples, pos_slot = positions of the positive ones (sample order)
  s->slot.clear(); s->pos_slot.clear();
  for (int64_t t = 0; t < n; t++) {
    const bool null_pair = (0.0 + (hi - 0.0) * s->rng.random()) > 1.0;  // point_sampler.py:58
    (null_pair ? s->slot : s->pos_slot).push_back(t);
  }
  const int64_t n_null = (int64_t)s->slot.size(), n_pos = (int64_t)s->pos_slot.size();
  // 2. null samples: the candidate stream, filtered in order (mem_dataset.py:158-163)
  constexpr int64_t kBlock = 128;
  s->cand_u.resize(kBlock); s->cand_i.resize(kBlock); s->cand_lo.resize(kBlock); s->cand_len.resize(kBlock);
  int64_t filled = 0;
  while (filled < n_null) {
    const int64_t m = std::min<int64_t>(kBlock, n_null - filled);   // every candidate fills at most one open sample
    for (int64_t k = 0; k < m; k++) {
      s->cand_u[(size_t)k] = s->null_rng.randint(0, s->max_uid);
      s->cand_i[(size_t)k] = s->null_rng.randint(0, s->max_iid);
      __builtin_prefetch(s->all_indptr + s->cand_u[(size_t)k]);
    }
    // membership of the m candidates by m interleaved lower-bound searches: one level of every search per sweep, so
    // that the cache misses of a level (one per candidate, all independent) overlap instead of forming a chain
    int64_t longest = 0;
    for (int64_t k = 0; k < m; k++) {
      const int64_t lo = s->all_indptr[s->cand_u[(size_t)k]], len = s->all_indptr[s->cand_u[(size_t)k] + 1] - lo;
      s->cand_lo[(size_t)k] = lo; s->cand_len[(size_t)k] = len;
      longest = std::max(longest, len);
      if (len > 0) __builtin_prefetch(s->all_iid + lo + (len >> 1));
    }
    for (; longest > 0; longest >>= 1) {
      for (int64_t k = 0; k < m; k++) {
        int64_t len = s->cand_len[(size_t)k];
        if (len <= 0) continue;
        int64_t lo = s->cand_lo[(size_t)k];
        const int64_t half = len >> 1;
        const bool right = s->all_iid[lo + half] < (int32_t)s->cand_i[(size_t)k];
        lo = right ? lo + half + 1 : lo;
        len = right ? len - half - 1 : half;
        s->cand_lo[(size_t)k] = lo; s->cand_len[(size_t)k] = len;
        if (len > 0) __builtin_prefetch(s->all_iid + lo + (len >> 1));
      }
    }
    for (int64_t k = 0; k < m; k++) {         // cand_lo = first entry of the row that is >= the item (or the row's end)
      const int64_t u = s->cand_u[(size_t)k], i = s->cand_i[(size_t)k];
      const int64_t pos = s->cand_lo[(size_t)k];
      const bool stored = pos < s->all_indptr[u + 1] && s->all_iid[pos] == (int32_t)i;
      if (!stored) {
        const int64_t t = s->slot[(size_t)filled++];
        uid[t] = (int32_t)u; iid[t] = (int32_t)i; val[t] = 0.0;
      }
    }
  }
  // 3. positive samples (mem_dataset.py:119-129): the stream depends on the row lengths only (a 1 MB array); the rows
  //    themselves are fetched after the stream has been walked
  s->pos_row.resize((size_t)n_pos);
  for (int64_t q = 0; q < n_pos; q++) {
    for (;;) {
      const int64_t u = s->pos_rng.randint(0, s->max_uid);
      const int64_t cnt = s->pos_indptr[u + 1] - s->pos_indptr[u];
      if (cnt == 0) continue;
      const int64_t j = s->pos_rng.randint(0, cnt - 1);
      const int64_t r = s->pos_indptr[u] + j;
      uid[s->pos_slot[(size_t)q]] = (int32_t)u;
      s->pos_row[(size_t)q] = r;
      __builtin_prefetch(s->pos_iid + r);
      __builtin_prefetch(s->pos_val + r);
      break;
    }
  }
  for (int64_t q = 0; q < n_pos; q++) {
    const int64_t t = s->pos_slot[(size_t)q], r = s->pos_row[(size_t)q];
    iid[t] = s->pos_iid[r]; val[t] = s->pos_val[r];
  }
  return DRB_OK;
}

int drb_sampler_getstate(const drb_sampler* s, uint32_t state[3 * 625]) {
  if (!s || !state) return drb_fail(DRB_E_INVALID, "drb_sampler_getstate: NULL argument");
  drb_rng_getstate(&s->rng, state);
  drb_rng_getstate(&s->null_rng, state + 625);
  drb_rng_getstate(&s->pos_rng, state + 1250);
  return DRB_OK;
}
int drb_sampler_setstate(drb_sampler* s, const uint32_t state[3 * 625]) {
  if (!s || !state) return drb_fail(DRB_E_INVALID, "drb_sampler_setstate: NULL argument");
  int r = drb_rng_setstate(&s->rng, state);
  if (!r) r = drb_rng_setstate(&s->null_rng, state + 625);
  if (!r) r = drb_rng_setstate(&s->pos_rng, state + 1250);
  return r;
}

// ------------------------------------------------------------------------------------------ corruption mask
int drb_batch_offsets(const int32_t* uids, int32_t batch, const int64_t* csr_indptr, int32_t* keep_off) {
  if (!uids || !csr_indptr || !keep_off || batch < 0) return drb_fail(DRB_E_INVALID, "drb_batch_offsets: bad argument");
  int64_t acc = 0;
  keep_off[0] = 0;
  for (int32_t b = 0; b < batch; b++) {
    acc += csr_indptr[uids[b] + 1] - csr_indptr[uids[b]];
    if (acc > INT32_MAX) return drb_fail(DRB_E_INVALID, "drb_batch_offsets: batch nnz exceeds int32");
    keep_off[b + 1] = (int32_t)acc;
  }
  return DRB_OK;
}

int drb_cdae_corruption_keep_mt(drb_rng* rng, const int32_t* uids, int32_t batch, int32_t n_items, double q,
                                const int64_t* csr_indptr, const int32_t* csr_indices, int32_t* keep_off,
                                uint8_t* keep, int64_t keep_capacity) {
  if (!rng || !uids || !csr_indptr || !csr_indices || !keep_off || !keep || batch < 0 || n_items <= 0)
    return drb_fail(DRB_E_INVALID, "drb_cdae_corruption_keep_mt: bad argument");
  int r = drb_batch_offsets(uids, batch, csr_indptr, keep_off);
  if (r) return r;
  // users are sampled with replacement, so a batch can hold more positives than any set of distinct users: the
  // caller states how much room `keep` has and nothing is written (and no draw consumed) when it is too small
  if ((int64_t)keep_off[batch] > keep_capacity)
    return drb_fail(DRB_E_INVALID, "drb_cdae_corruption_keep_mt: batch holds %lld positives, keep has room for %lld",
                    (long long)keep_off[batch], (long long)keep_capacity);
  for (int32_t b = 0; b < batch; b++) {
    // cdae.py:63-64: draw i of this user decides item i; every draw is rng.uniform(0,1) = two MT outputs.
    const int64_t lo = csr_indptr[uids[b]], hi = csr_indptr[uids[b] + 1];
    uint8_t* out = keep + keep_off[b];
    int64_t pos = 0;  // next item index whose draw has not been consumed yet
    for (int64_t j = lo; j < hi; j++) {
      int64_t item = csr_indices[j];
      rng->skip(2 * (item - pos));
      double u = 0.0 + (1.0 - 0.0) * rng->random();
      out[j - lo] = (u < q) ? 0 : 1;
      pos = item + 1;
    }
    rng->skip(2 * ((int64_t)n_items - pos));
  }
  return DRB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ evaluation candidates
namespace {

struct EvalShared {
  const int64_t* test_indptr; const int64_t* test_item; const double* test_val;
  const int64_t* black_row; const int64_t* black_indptr; const int32_t* black_iid;
  const int64_t* raw_sorted; const int32_t* raw_to_iid; int64_t n_map;
  int train_evaluation; int64_t n_items; double threshold; int64_t n_pos; double n_neg; int n_neg_is_frac;
  int generate; int64_t seed;
};

struct EvalChunk {   // output of one worker: users [lo, hi)
  int64_t lo, hi;
  std::vector<int64_t> cand, pos, cand_len, pos_len;
  std::vector<uint8_t> skipped;
};

// is raw item `it` one of the training positives of evaluated user u?
inline bool in_train(const EvalShared& S, int64_t u, int64_t it) {
  if (!S.black_row || S.black_row[u] < 0) return false;
  const int64_t* p = std::lower_bound(S.raw_sorted, S.raw_sorted + S.n_map, it);
  if (p == S.raw_sorted + S.n_map || *p != it) return false;
  const int32_t iid = S.raw_to_iid[p - S.raw_sorted];
  const int64_t r = S.black_row[u];
  return std::binary_search(S.black_iid + S.black_indptr[r], S.black_iid + S.black_indptr[r + 1], iid);
}

void eval_worker(const EvalShared& S, EvalChunk& C) {
  drb_rng g;
  std::vector<int64_t> p_items, n_pool, negs, picked, all, chosen, tp;
  // Generated negatives are integers in [0, n_items): two byte maps over that range replace a hash set of the
  // negatives drawn so far and the two binary searches per draw of `in_train` (both cleared per user by walking
  // what was set).  iid_raw = inverse of the sorted raw-id map, built once per worker.
  std::vector<uint8_t> neg_mark((size_t)std::max<int64_t>(S.n_items, 0), 0), train_mark(neg_mark.size(), 0);
  std::vector<int64_t> iid_raw, train_set;
  if (S.black_row && S.n_map > 0) {
    int32_t max_iid = -1;
    for (int64_t k = 0; k < S.n_map; k++) max_iid = std::max(max_iid, S.raw_to_iid[k]);
    iid_raw.assign((size_t)max_iid + 1, -1);
    for (int64_t k = 0; k < S.n_map; k++) iid_raw[S.raw_to_iid[k]] = S.raw_sorted[k];
  }
  for (int64_t u = C.lo; u < C.hi; u++) {
    int64_t s = S.seed + u;  // ranking_evaluation.py:111-116
    g.seed((uint64_t)(s < 0 ? -s : s));
    C.skipped.push_back(1); C.cand_len.push_back(0); C.pos_len.push_back(0);
    const int64_t lo = S.test_indptr[u], hi = S.test_indptr[u + 1];
    p_items.clear(); n_pool.clear(); chosen.clear();
    for (int64_t r = lo; r < hi; r++) (S.test_val[r] >= S.threshold ? p_items : n_pool).push_back(S.test_item[r]);
    if (S.n_pos < 0) {
      chosen = p_items;
    } else {
      if ((int64_t)p_items.size() < S.n_pos) continue;  // :175-176
      picked.resize(S.n_pos);
      sample_indices(g, (int64_t)p_items.size(), S.n_pos, picked.data());
      for (int64_t i : picked) chosen.push_back(p_items[i]);
    }
    negs.clear();
    if (S.n_neg < 0) {
      negs = n_pool;
    } else {
      const int64_t want = S.n_neg_is_frac ? (int64_t)(S.n_neg * (double)chosen.size()) : (int64_t)S.n_neg;  // :187-188
      const int64_t take = std::min<int64_t>(want, (int64_t)n_pool.size());
      picked.resize(take);
      sample_indices(g, (int64_t)n_pool.size(), take, picked.data());
      for (int64_t i : picked) negs.push_back(n_pool[i]);
      if ((int64_t)negs.size() < want && S.generate) {
        // blacklist = train positives (unless evaluating on train) U test positives  (:193-201)
        tp = p_items;
        std::sort(tp.begin(), tp.end());
        tp.erase(std::unique(tp.begin(), tp.end()), tp.end());
        int64_t black_size = (int64_t)tp.size();
        if (!S.train_evaluation && S.black_row && S.black_row[u] >= 0) {
          const int64_t r = S.black_row[u];
          black_size = S.black_indptr[r + 1] - S.black_indptr[r];
          for (int64_t it : tp) black_size += in_train(S, u, it) ? 0 : 1;
        }
        if (S.n_items - black_size < want) continue;  // :202-207
        const size_t first_generated = negs.size();
        for (int64_t it : negs)                          // sampled test negatives that fall inside the drawn range
          if (it >= 0 && it < S.n_items) neg_mark[(size_t)it] = 1;
        train_set.clear();
        if (!S.train_evaluation && S.black_row && S.black_row[u] >= 0) {
          const int64_t r = S.black_row[u];
          for (int64_t k = S.black_indptr[r]; k < S.black_indptr[r + 1]; k++) {
            const int32_t iid = S.black_iid[k];
            const int64_t raw = (iid >= 0 && (size_t)iid < iid_raw.size()) ? iid_raw[iid] : -1;
            if (raw >= 0 && raw < S.n_items) { train_mark[(size_t)raw] = 1; train_set.push_back(raw); }
          }
        }
        while ((int64_t)negs.size() < want) {
          const int64_t it = g.randint(0, S.n_items - 1);  // :211 (raw/internal id confusion kept)
          if (neg_mark[(size_t)it] || train_mark[(size_t)it] || std::binary_search(tp.begin(), tp.end(), it)) continue;
          negs.push_back(it);
          neg_mark[(size_t)it] = 1;
        }
        for (size_t k = 0; k < first_generated; k++)
          if (negs[k] >= 0 && negs[k] < S.n_items) neg_mark[(size_t)negs[k]] = 0;
        for (size_t k = first_generated; k < negs.size(); k++) neg_mark[(size_t)negs[k]] = 0;
        for (int64_t raw : train_set) train_mark[(size_t)raw] = 0;
      }
    }
    all = chosen;
    all.insert(all.end(), negs.begin(), negs.end());
    if (all.empty()) continue;  // :217-218
    shuffle_i64(g, (int64_t)all.size(), all.data());
    C.cand.insert(C.cand.end(), all.begin(), all.end());
    C.pos.insert(C.pos.end(), chosen.begin(), chosen.end());
    C.cand_len.back() = (int64_t)all.size();
    C.pos_len.back() = (int64_t)chosen.size();
    C.skipped.back() = 0;
  }
}

}  // namespace

extern "C" {

int drb_eval_candidates(int64_t n_users, const int64_t* test_indptr, const int64_t* test_item,
                        const double* test_val, const int64_t* black_row, const int64_t* black_indptr,
                        const int32_t* black_iid, const int64_t* raw_sorted, const int32_t* raw_to_iid,
                        int64_t n_map, int32_t train_evaluation, int64_t n_items, double threshold, int64_t n_pos,
                        double n_neg, int32_t n_neg_is_frac, int32_t generate_negative_pairs, int64_t seed,
                        int32_t n_threads, int64_t cand_capacity, int64_t* cand_off, int64_t* cand, int64_t* pos_off,
                        int64_t* pos, uint8_t* skipped) {
  if (n_users < 0 || !test_indptr || !cand_off || !pos_off || !skipped)
    return drb_fail(DRB_E_INVALID, "drb_eval_candidates: bad argument");
  if (black_row && (!black_indptr || !black_iid || (n_map > 0 && (!raw_sorted || !raw_to_iid))))
    return drb_fail(DRB_E_INVALID, "drb_eval_candidates: incomplete training-positives description");
  EvalShared S{test_indptr, test_item, test_val, black_row, black_indptr, black_iid, raw_sorted, raw_to_iid, n_map,
               train_evaluation, n_items, threshold, n_pos, n_neg, n_neg_is_frac, generate_negative_pairs, seed};
  int nt = std::max(1, std::min<int>(n_threads, 64));
  if (n_users < 256) nt = 1;
  std::vector<EvalChunk> chunks(nt);
  const int64_t per = (n_users + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    chunks[t].lo = std::min<int64_t>(n_users, t * per);
    chunks[t].hi = std::min<int64_t>(n_users, (t + 1) * per);
  }
  if (nt == 1) {
    eval_worker(S, chunks[0]);
  } else {   // users are independent (each has its own Random(seed + idx)): one contiguous block per thread
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(eval_worker, std::cref(S), std::ref(chunks[t]));
    for (auto& x : th) x.join();
  }
  int64_t co = 0, po = 0;
  cand_off[0] = 0; pos_off[0] = 0;
  for (auto& C : chunks) {
    if (co + (int64_t)C.cand.size() > cand_capacity)
      return drb_fail(DRB_E_INVALID, "drb_eval_candidates: candidate buffer too small");
    if (cand) std::copy(C.cand.begin(), C.cand.end(), cand + co);
    if (pos) std::copy(C.pos.begin(), C.pos.end(), pos + po);
    for (int64_t u = C.lo; u < C.hi; u++) {
      co += C.cand_len[u - C.lo]; po += C.pos_len[u - C.lo];
      cand_off[u + 1] = co; pos_off[u + 1] = po;
      skipped[u] = C.skipped[u - C.lo];
    }
  }
  return DRB_OK;
}

// ---------------------------------------------------------------------------------------------- grouped key lookup
// For every group g (one evaluated user): out[q] = value of the FIRST table row of the group whose key equals
// q_key[q], else `miss`, for q in [q_beg[g], q_end[g]).  The table rows of group g are [tab_beg[g], tab_end[g]) of
// tab_key / tab_val (tab_val NULL: every row counts 1.0).  This is the per-candidate `ds_test.select_one(uid, iid)`
// of ranking_evaluation.py:222-223 (relevancy of a ranked / candidate item = the first matching test row, else 0) and
// the `item in positives` test of the metrics (Metrics/ranking.py:32-56), for all users in one call.
int drb_eval_lookup(int64_t n_groups, const int64_t* tab_beg, const int64_t* tab_end, const int64_t* tab_key,
                    const double* tab_val, const int64_t* q_beg, const int64_t* q_end, const int64_t* q_key,
                    double miss, int32_t n_threads, double* out) {
  if (n_groups < 0 || !tab_beg || !tab_end || !q_beg || !q_end || !out)
    return drb_fail(DRB_E_INVALID, "drb_eval_lookup: bad argument");
  auto worker = [&](int64_t lo, int64_t hi) {
    std::vector<std::pair<int64_t, int64_t>> idx;   // (key, first row) of a long table, sorted by key
    for (int64_t g = lo; g < hi; g++) {
      const int64_t tb = tab_beg[g], te = tab_end[g], nt = te - tb;
      const bool longtab = nt > 16;
      if (longtab) {
        idx.clear();
        for (int64_t r = tb; r < te; r++) idx.emplace_back(tab_key[r], r);
        std::sort(idx.begin(), idx.end());          // pairs: equal keys ordered by row, so the first row comes first
      }
      for (int64_t q = q_beg[g]; q < q_end[g]; q++) {
        const int64_t key = q_key[q];
        int64_t hit = -1;
        if (longtab) {
          auto it = std::lower_bound(idx.begin(), idx.end(), std::make_pair(key, (int64_t)INT64_MIN));
          if (it != idx.end() && it->first == key) hit = it->second;
        } else {
          for (int64_t r = tb; r < te; r++)
            if (tab_key[r] == key) { hit = r; break; }
        }
        out[q] = hit < 0 ? miss : (tab_val ? tab_val[hit] : 1.0);
      }
    }
  };
  int nt = std::max(1, std::min<int>(n_threads, 64));
  if (n_groups < 1024) nt = 1;
  if (nt == 1) {
    worker(0, n_groups);
  } else {
    std::vector<std::thread> th;
    const int64_t per = (n_groups + nt - 1) / nt;
    for (int t = 0; t < nt; t++)
      th.emplace_back(worker, std::min<int64_t>(n_groups, t * per), std::min<int64_t>(n_groups, (t + 1) * per));
    for (auto& x : th) x.join();
  }
  return DRB_OK;
}

// ---------------------------------------------------------------------------------------------- metric sums
int drb_eval_metrics(int64_t n_groups, const int64_t* tab_beg, const int64_t* tab_end, const int64_t* tab_key,
                     const double* tab_val, const int64_t* ranked, int64_t ld_ranked, const int32_t* n_out,
                     const int64_t* c_beg, const int64_t* c_end, const int64_t* c_key, const int64_t* p_beg,
                     const int64_t* p_end, const int64_t* p_key, const int64_t* ks, int32_t n_ks, int32_t n_threads,
                     double* dcg, double* idcg, int64_t* hits) {
  if (n_groups < 0 || n_ks < 1 || !tab_beg || !tab_end || !tab_key || !tab_val || !ranked || !n_out || !c_beg ||
      !c_end || !c_key || !p_beg || !p_end || !p_key || !ks || !dcg || !idcg || !hits)
    return drb_fail(DRB_E_INVALID, "drb_eval_metrics: bad argument");
  int64_t k_max = 0;
  for (int j = 0; j < n_ks; j++) k_max = std::max(k_max, ks[j]);
  auto worker = [&](int64_t lo, int64_t hi) {
    std::vector<double> rel, gain;
    std::vector<std::pair<int64_t, int64_t>> idx;       // (key, first row) of a long test table, sorted by key
    std::vector<int64_t> psorted;
    for (int64_t g = lo; g < hi; g++) {
      const int64_t tb = tab_beg[g], te = tab_end[g];
      const bool longtab = te - tb > 16;
      if (longtab) {
        idx.clear();
        for (int64_t r = tb; r < te; r++) idx.emplace_back(tab_key[r], r);
        std::sort(idx.begin(), idx.end());
      }
      auto relevancy = [&](int64_t key) -> double {     // first matching test row, else 0
        if (longtab) {
          auto it = std::lower_bound(idx.begin(), idx.end(), std::make_pair(key, (int64_t)INT64_MIN));
          return (it != idx.end() && it->first == key) ? tab_val[it->second] : 0.0;
        }
        for (int64_t r = tb; r < te; r++)
          if (tab_key[r] == key) return tab_val[r];
        return 0.0;
      };
      const int64_t np_ = p_end[g] - p_beg[g];
      const bool longpos = np_ > 16;
      if (longpos) { psorted.assign(p_key + p_beg[g], p_key + p_end[g]); std::sort(psorted.begin(), psorted.end()); }
      auto is_positive = [&](int64_t key) -> bool {
        if (longpos) return std::binary_search(psorted.begin(), psorted.end(), key);
        for (int64_t r = p_beg[g]; r < p_end[g]; r++)
          if (p_key[r] == key) return true;
        return false;
      };
      // ranked list: running DCG and hit count, read off at every cut-off
      const int64_t nr = std::min<int64_t>(n_out[g], k_max);
      double cur = 0.0;
      int64_t nh = 0;
      for (int j = 0; j < n_ks; j++) { dcg[g * n_ks + j] = 0.0; hits[g * n_ks + j] = 0; idcg[g * n_ks + j] = 0.0; }
      for (int64_t i = 0; i < nr; i++) {
        const int64_t item = ranked[g * ld_ranked + i];
        cur += (std::pow(2.0, relevancy(item)) - 1.0) / std::log2(2.0 + (double)i);
        nh += is_positive(item) ? 1 : 0;
        for (int j = 0; j < n_ks; j++)
          if (i < ks[j]) { dcg[g * n_ks + j] = cur; hits[g * n_ks + j] = nh; }
      }
      // ideal list: the candidates' relevancies in descending order
      const int64_t nc = c_end[g] - c_beg[g];
      rel.resize((size_t)nc);
      for (int64_t q = 0; q < nc; q++) rel[(size_t)q] = relevancy(c_key[c_beg[g] + q]);
      const int64_t top = std::min<int64_t>(nc, k_max);
      std::partial_sort(rel.begin(), rel.begin() + top, rel.end(), std::greater<double>());
      double best = 0.0;
      for (int64_t i = 0; i < top; i++) {
        best += (std::pow(2.0, rel[(size_t)i]) - 1.0) / std::log2(2.0 + (double)i);
        for (int j = 0; j < n_ks; j++)
          if (i < ks[j]) idcg[g * n_ks + j] = best;
      }
    }
  };
  int nt = std::max(1, std::min<int>(n_threads, 64));
  if (n_groups < 1024) nt = 1;
  if (nt == 1) {
    worker(0, n_groups);
  } else {
    std::vector<std::thread> th;
    const int64_t per = (n_groups + nt - 1) / nt;
    for (int t = 0; t < nt; t++)
      th.emplace_back(worker, std::min<int64_t>(n_groups, t * per), std::min<int64_t>(n_groups, (t + 1) * per));
    for (auto& x : th) x.join();
  }
  return DRB_OK;
}

// ---------------------------------------------------------------------------------------------- leave-k-out split
// DRecPy/Evaluation/Splits/leave_k_out.py:14-135 for every user at once.  User idx (order of first appearance) draws
// from random.Random(seed + idx + 1) -- the reference increments the seed *before* creating each user's generator
// (leave_k_out.py:68-69) -- and rng.sample(rids, k) picks k of the user's rows, which are given here as positions
// user_indptr[idx] .. user_indptr[idx+1] of the caller's "rows grouped by user, in DataFrame order" permutation.
// flags[pos]: 0 = train, 1 = test, 2 = removed (user with fewer than min_user_interactions rows).
int drb_leave_k_out(int64_t n_users, const int64_t* user_indptr, int64_t k_fixed, double k_ratio, int32_t is_ratio,
                    int64_t min_user_interactions, int64_t seed, int32_t n_threads, uint8_t* flags) {
  if (n_users < 0 || !user_indptr || !flags) return drb_fail(DRB_E_INVALID, "drb_leave_k_out: bad argument");
  if (!is_ratio && k_fixed <= 0) return drb_fail(DRB_E_INVALID, "The value of k (%lld) must be > 0.", (long long)k_fixed);
  auto worker = [&](int64_t lo, int64_t hi) {
    drb_rng g;
    std::vector<int64_t> picked;
    for (int64_t u = lo; u < hi; u++) {
      const int64_t beg = user_indptr[u], n = user_indptr[u + 1] - beg;
      const int64_t k = is_ratio ? (int64_t)((double)n * k_ratio) : k_fixed;   // int(len * k)
      if (n < min_user_interactions) {
        std::fill(flags + beg, flags + beg + n, (uint8_t)2);
      } else if (n > k && k > 0) {
        const int64_t s = seed + u + 1;
        g.seed((uint64_t)(s < 0 ? -s : s));
        picked.resize(k);
        sample_indices(g, n, k, picked.data());
        for (int64_t j = 0; j < k; j++) flags[beg + picked[j]] = 1;
      }
    }
  };
  int nt = std::max(1, std::min<int>(n_threads, 64));
  if (n_users < 1024) nt = 1;
  if (nt == 1) {
    worker(0, n_users);
  } else {
    std::vector<std::thread> th;
    const int64_t per = (n_users + nt - 1) / nt;
    for (int t = 0; t < nt; t++)
      th.emplace_back(worker, std::min<int64_t>(n_users, t * per), std::min<int64_t>(n_users, (t + 1) * per));
    for (auto& x : th) x.join();
  }
  return DRB_OK;
}

}  // extern "C"
