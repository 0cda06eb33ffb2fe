// Jump-ahead tables for MT19937 (host side): what a GPU needs to replay the reference's corruption stream in parallel.
//
// Reference behaviour being reproduced: DRecPy/Recommender/cdae.py:63-64 draws rng.uniform(0, 1) once per ITEM for every
// sampled user from ONE sequential random.Random (recommender_abc.py:74): 2 * n_items MT19937 outputs per user,
// 2.2e8 per 4096-user step at the ml-20m shape.  A sequential replay is 0.3 s per step on the host.  MT19937 is linear
// over GF(2): if phi(t) is the characteristic polynomial of its recurrence (degree 19937) and g_J(t) = t^J mod phi(t),
// then for the untempered output sequence y:  y[n + J] = XOR_{i : g_J[i] = 1} y[n + i]  for every n.  So a block of
// the stream that starts J outputs ahead is a GF(2) combination of the first 19937 + 623 words of the stream, and many
// blocks can be produced at once (mt_device.cu).  This file computes phi (Berlekamp-Massey on one output bit) and the
// polynomials g for the offsets seg_len * p, p = 1 .. n_seg, plus one for the whole step.
#include <cstdint>
#include <cstring>
#include <vector>

#include "drb_internal.h"

namespace {

constexpr int NBITS = 19937, NW = 312;          // polynomials of degree < 19937 in 312 64-bit words
constexpr int MT_N = 624, MT_M = 397;

struct Poly { uint64_t w[NW]; };
struct Wide { uint64_t w[2 * NW + 1]; };

inline bool get_bit(const uint64_t* w, int i) { return (w[i >> 6] >> (i & 63)) & 1u; }
inline void flip_bit(uint64_t* w, int i) { w[i >> 6] ^= 1ull << (i & 63); }

// dst[0 .. n_dst) ^= src[0 .. n_src) << shift (bit shift)
inline void xor_shifted(uint64_t* dst, int n_dst, const uint64_t* src, int n_src, int shift) {
  const int ws = shift >> 6, bs = shift & 63;
  if (bs == 0) {
    for (int i = 0; i < n_src && i + ws < n_dst; i++) dst[i + ws] ^= src[i];
    return;
  }
  for (int i = 0; i < n_src; i++) {
    if (i + ws < n_dst) dst[i + ws] ^= src[i] << bs;
    if (i + ws + 1 < n_dst) dst[i + ws + 1] ^= src[i] >> (64 - bs);
  }
}

// one untempered MT19937 word sequence from a standard seed (any non-degenerate state gives the same phi)
void raw_sequence(std::vector<uint32_t>& y, int count) {
  std::vector<uint32_t> x(MT_N + count);
  x[0] = 5489u;
  for (int i = 1; i < MT_N; i++) x[i] = 1812433253u * (x[i - 1] ^ (x[i - 1] >> 30)) + (uint32_t)i;
  for (int k = 0; k < count; k++) {
    const uint32_t v = (x[k] & 0x80000000u) | (x[k + 1] & 0x7fffffffu);
    x[k + MT_N] = x[k + MT_M] ^ (v >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  }
  y.assign(x.begin() + MT_N, x.end());
}

// Berlekamp-Massey over GF(2) on s[0 .. n): connection polynomial C (C[0] = 1) with s[k] = XOR_{i=1..L} C[i] s[k-i]
int berlekamp_massey(const std::vector<uint8_t>& s, std::vector<uint64_t>& C) {
  const int n = (int)s.size(), words = (n >> 6) + 2;
  std::vector<uint64_t> Cc(words, 0), Bc(words, 0), T(words), srev(words, 0);   // srev bit i = s[k - i]
  Cc[0] = Bc[0] = 1;
  int L = 0, m = 1;
  for (int k = 0; k < n; k++) {
    for (int i = words - 1; i > 0; i--) srev[i] = (srev[i] << 1) | (srev[i - 1] >> 63);   // shift in s[k] at bit 0
    srev[0] = (srev[0] << 1) | (uint64_t)s[k];
    uint64_t acc = 0;
    const int lw = (L >> 6) + 1;
    for (int i = 0; i < lw; i++) acc ^= Cc[i] & srev[i];
    if (__builtin_parityll(acc)) {       // discrepancy (C[0] = 1 pairs with s[k] itself)
      T = Cc;
      xor_shifted(Cc.data(), words, Bc.data(), words, m);
      if (2 * L <= k) { L = k + 1 - L; Bc = T; m = 1; } else { m++; }
    } else {
      m++;
    }
  }
  C = Cc;
  return L;
}

struct Phi {
  uint64_t w[NW + 1];          // phi(t), degree 19937 (bit 19937 set)
  uint64_t sh[64][NW + 2];     // phi << b for b = 0..63 (word-aligned XORs in the reduction)
};

// r (degree < 2 * NBITS) mod phi -> out
void reduce(Wide& r, const Phi& phi, Poly& out) {
  for (int d = 2 * NBITS - 2; d >= NBITS; d--) {
    if (!get_bit(r.w, d)) continue;
    const int shift = d - NBITS, ws = shift >> 6, bs = shift & 63;
    const uint64_t* p = phi.sh[bs];
    for (int i = 0; i < NW + 2 && i + ws < 2 * NW + 1; i++) r.w[i + ws] ^= p[i];
  }
  std::memcpy(out.w, r.w, sizeof(out.w));
  out.w[NW - 1] &= (1ull << (NBITS - 64 * (NW - 1))) - 1;      // bits >= 19937 are zero after the reduction
}

void mulmod(const Poly& a, const Poly& b, const Phi& phi, Poly& out) {
  Wide r;
  std::memset(&r, 0, sizeof(r));
  for (int wi = 0; wi < NW; wi++) {
    uint64_t x = a.w[wi];
    while (x) {
      const int bit = __builtin_ctzll(x);
      x &= x - 1;
      xor_shifted(r.w, 2 * NW + 1, b.w, NW, 64 * wi + bit);
    }
  }
  reduce(r, phi, out);
}

void pow_t(uint64_t J, const Phi& phi, Poly& out) {      // t^J mod phi
  Poly result, base;
  std::memset(&result, 0, sizeof(result));
  std::memset(&base, 0, sizeof(base));
  result.w[0] = 1;      // 1
  base.w[0] = 2;        // t
  while (J) {
    if (J & 1) { Poly t; mulmod(result, base, phi, t); result = t; }
    J >>= 1;
    if (J) { Poly t; mulmod(base, base, phi, t); base = t; }
  }
  out = result;
}

const Phi* get_phi() {
  static Phi* phi = nullptr;
  if (phi) return phi;
  std::vector<uint32_t> y;
  raw_sequence(y, 2 * NBITS + 64);
  std::vector<uint8_t> s(2 * NBITS + 64);
  for (size_t i = 0; i < s.size(); i++) s[i] = y[i] & 1u;
  std::vector<uint64_t> C;
  const int L = berlekamp_massey(s, C);
  if (L != NBITS) return nullptr;
  Phi* p = new Phi;
  std::memset(p, 0, sizeof(Phi));
  for (int i = 0; i <= NBITS; i++)               // phi(t) = t^L C(1/t): coefficient of t^(L-i) is C[i]
    if (get_bit(C.data(), i)) flip_bit(p->w, NBITS - i);
  for (int b = 0; b < 64; b++) xor_shifted(p->sh[b], NW + 2, p->w, NW + 1, b);
  phi = p;
  return phi;
}

}  // namespace

struct drb_mtjump {
  int64_t seg_len; int32_t n_seg;
  std::vector<Poly> polys;      // [n_seg]: polys[p - 1] = t^(seg_len * p) mod phi
};

extern "C" {

int drb_mtjump_create(int64_t seg_len, int32_t n_seg, drb_mtjump** out) {
  if (!out || seg_len < 1 || n_seg < 1) return drb_fail(DRB_E_INVALID, "drb_mtjump_create: bad argument");
  const Phi* phi = get_phi();
  if (!phi) return drb_fail(DRB_E_STATE, "drb_mtjump_create: characteristic polynomial of MT19937 not recovered");
  drb_mtjump* j = new (std::nothrow) drb_mtjump;
  if (!j) return drb_fail(DRB_E_NOMEM, "out of memory");
  j->seg_len = seg_len; j->n_seg = n_seg;
  j->polys.resize(n_seg);
  pow_t((uint64_t)seg_len, *phi, j->polys[0]);
  for (int p = 1; p < n_seg; p++) mulmod(j->polys[p - 1], j->polys[0], *phi, j->polys[p]);
  *out = j;
  return DRB_OK;
}

int drb_mtjump_destroy(drb_mtjump* j) { delete j; return DRB_OK; }

int drb_mtjump_polys(const drb_mtjump* j, uint64_t* out) {
  if (!j || !out) return drb_fail(DRB_E_INVALID, "drb_mtjump_polys: NULL argument");
  std::memcpy(out, j->polys.data(), j->polys.size() * sizeof(Poly));
  return DRB_OK;
}

// Host application of one jump (tests): window_out[w] = XOR_{i : g[i]} y[i + w], y = the stream continued from window_in
int drb_mtjump_apply_host(const drb_mtjump* j, int32_t p, const uint32_t* window_in, uint32_t* window_out) {
  if (!j || !window_in || !window_out || p < 1 || p > j->n_seg) return drb_fail(DRB_E_INVALID, "drb_mtjump_apply_host: bad argument");
  std::vector<uint32_t> y(NBITS + MT_N + MT_N);
  std::memcpy(y.data(), window_in, MT_N * 4);
  for (int k = 0; k + MT_N < (int)y.size(); k++) {
    const uint32_t v = (y[k] & 0x80000000u) | (y[k + 1] & 0x7fffffffu);
    y[k + MT_N] = y[k + MT_M] ^ (v >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  }
  const Poly& g = j->polys[p - 1];
  std::memset(window_out, 0, MT_N * 4);
  for (int i = 0; i < NBITS; i++)
    if (get_bit(g.w, i))
      for (int w = 0; w < MT_N; w++) window_out[w] ^= y[i + w];
  return DRB_OK;
}

}  // extern "C"
