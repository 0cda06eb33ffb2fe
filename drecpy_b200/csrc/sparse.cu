// K1 / K3 / K5 sparse kernels (sm_100a): CSR row gather-sum fused with mask + row bias + bias + activation,
// its transpose (vector-atomic scatter of row gradients), and the per-batch label / mask preparation.
//
// Replaces (reference, DRecPy/):
//   Recommender/cdae.py:59-65,73-75   densify + corrupt + tf.matmul([x~], W) + V_u + b + sigmoid
//   Recommender/dmf.py:75-86,89-90    densify + l2_normalize + first Dense(relu) of each tower
// and the matching half of tape.gradient (recommender_abc.py:203).
//
// Work decomposition: one 128-thread CTA per sampled row.  A row of the table is ld floats (ld % 4 == 0); LPR
// lanes cooperate on one table row with 128-bit loads, so a warp streams 32/LPR table rows at a time and each
// lane keeps NV float4 accumulators.  Indices / weights are fetched 32 at a time per warp and broadcast with
// shuffles.  Masked-out (corrupted) entries are never loaded.  The table (W: 21 MB at the ml-20m shape) is
// L2-resident on B200, so this kernel is bound by L2->SM bandwidth and load issue, not HBM.
#include "kernels.h"

#include <algorithm>

namespace {

constexpr int kGatherThreads = 128;
constexpr int kWarps = kGatherThreads / 32;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == DRB_ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
  if (act == DRB_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

// Inner loop of the gathers: the warp walks `deg` CSR entries starting at `lo`, 32 (index, weight) pairs at a time,
// broadcast by shuffle.  The table-row loads of kGatherRows entries are issued back to back (predicated, no branch)
// before the first FMA consumes one: with one row per iteration the FMAs of row t stalled in front of the loads of
// row t+1 and every warp had a single row in flight (SASS), which left the kernel latency bound at ~5 TB/s from L2.
constexpr int kGatherRows = 4;
template <int LPR, int NV>
__device__ __forceinline__ void gather_accumulate(const GatherArgs& a, int64_t lo, int deg, const uint8_t* keep, int warp,
                                                  int lane, float4 (&acc)[NV]) {
  constexpr int G = 32 / LPR;
  const int sub = lane / LPR, c = lane % LPR;
  const int ld4 = a.ld >> 2;
  for (int base = warp * 32; base < deg; base += kWarps * 32) {
    const int j = base + lane;
    int my_idx = 0;
    float my_w = 0.f;
    if (j < deg) {
      my_idx = a.indices[lo + j];
      my_w = a.values ? a.values[lo + j] : 1.0f;
      if (keep && keep[j] == 0) my_w = 0.f;
    }
    const int cnt = min(32, deg - base);
    for (int t = 0; t < cnt; t += G * kGatherRows) {
      float4 x[kGatherRows][NV];
      float w[kGatherRows];
#pragma unroll
      for (int u = 0; u < kGatherRows; u++) {
        const int src = t + u * G + sub;
        const int idx = __shfl_sync(0xffffffffu, my_idx, src & 31);
        const float ww = __shfl_sync(0xffffffffu, my_w, src & 31);
        const bool ok = src < cnt && ww != 0.f;       // dropped (corrupted) entries are never loaded
        w[u] = ok ? ww : 0.f;
        const float* rp = a.table + (int64_t)idx * a.ld;
#pragma unroll
        for (int v = 0; v < NV; v++) {
          const int c4 = c + v * LPR;
          x[u][v] = (ok && c4 < ld4) ? ldg4(rp + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < kGatherRows; u++) {
#pragma unroll
        for (int v = 0; v < NV; v++) {
          acc[v].x = fmaf(w[u], x[u][v].x, acc[v].x);
          acc[v].y = fmaf(w[u], x[u][v].y, acc[v].y);
          acc[v].z = fmaf(w[u], x[u][v].z, acc[v].z);
          acc[v].w = fmaf(w[u], x[u][v].w, acc[v].w);
        }
      }
    }
  }
}

// blockIdx.y selects one of two argument sets of equal row width: the two DMF towers run as one launch
__device__ __forceinline__ void red_write(float4* red, int col, float y) { reinterpret_cast<float*>(red)[col] = y; }

template <int LPR, int NV>
__global__ void __launch_bounds__(kGatherThreads) k_gather(GatherArgs a0, GatherArgs a1) {
  constexpr int G = 32 / LPR;  // table rows streamed concurrently by one warp
  extern __shared__ float4 red[];  // [kWarps * G][ld4]
  const GatherArgs a = blockIdx.y ? a1 : a0;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, c = lane % LPR;
  const int ld4 = a.ld >> 2;
  const int row = a.rows[b];
  const int64_t lo = a.indptr[row], hi = a.indptr[row + 1];
  const int deg = (int)(hi - lo);
  const uint8_t* keep = a.keep ? a.keep + a.keep_off[b] : nullptr;

  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

  gather_accumulate<LPR, NV>(a, lo, deg, keep, warp, lane, acc);
#pragma unroll
  for (int v = 0; v < NV; v++) {
    const int c4 = c + v * LPR;
    if (c4 < ld4) red[(warp * G + sub) * ld4 + c4] = acc[v];
  }
  __syncthreads();
  // finalize: fixed summation order over the kWarps*G partials -> deterministic
  const float* redf = reinterpret_cast<const float*>(red);
  float rs = a.scale;
  if (a.row_scale) rs *= a.row_scale[row];
  for (int col = threadIdx.x; col < a.ld; col += kGatherThreads) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < kWarps * G; p++) s += redf[p * a.ld + col];
    float z = s * rs;
    const int brow = a.bias_rows ? a.bias_rows[b] : row;     // item-sharded: only the user's owner adds V_u + b
    if (brow >= 0) {
      if (a.rowbias) z += a.rowbias[(int64_t)brow * a.ld + col];
      if (a.bias) z += a.bias[col];
    }
    const float y = (col < a.width) ? apply_act(z, a.act) : 0.f;
    a.out[(int64_t)b * a.ld + col] = y;
    if (a.next_out) red_write(red, col, y);       // this thread is the only reader of column `col` of the partials
  }
  if (!a.next_out) return;
  // fused second Dense layer of a DMF tower (dmf.py:48-58): next_out[b] = act(out[b] . next_k + next_b), exact fp32,
  // fixed summation order
  __syncthreads();
  for (int n = threadIdx.x; n < a.next_ld; n += kGatherThreads) {
    float z = 0.f;
    if (n < a.next_width) {
      for (int cc = 0; cc < a.width; cc++) z = fmaf(redf[cc], __ldg(a.next_k + (int64_t)cc * a.next_ld + n), z);
      z = apply_act(z + a.next_b[n], a.next_act);
    }
    a.next_out[(int64_t)b * a.next_ld + n] = z;
  }
}

template <int LPR, int NV>
__global__ void __launch_bounds__(kGatherThreads) k_scatter(ScatterArgs a0, ScatterArgs a1) {
  constexpr int G = 32 / LPR;
  const ScatterArgs a = blockIdx.y ? a1 : a0;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, c = lane % LPR;
  const int ld4 = a.ld >> 2;
  const int row = a.rows[b];
  const int64_t lo = a.indptr[row], hi = a.indptr[row + 1];
  const int deg = (int)(hi - lo);
  const uint8_t* keep = a.keep ? a.keep + a.keep_off[b] : nullptr;
  float rs = a.scale;
  if (a.row_scale) rs *= a.row_scale[row];

  float4 d[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) {
    const int c4 = c + v * LPR;
    d[v] = (c4 < ld4) ? ldg4(a.d + (int64_t)b * a.ld + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int brow = a.bias_rows ? a.bias_rows[b] : row;
  if (a.growbias && brow >= 0 && warp == 0 && sub == 0) {
    if (a.row_touched && c == 0) atomicOr(a.row_touched + (brow >> 5), 1u << (brow & 31));
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const int c4 = c + v * LPR;
      if (c4 < ld4) atomicAdd(reinterpret_cast<float4*>(a.growbias + (int64_t)brow * a.ld) + c4, d[v]);
    }
  }
  for (int base = warp * 32; base < deg; base += kWarps * 32) {
    const int j = base + lane;
    int my_idx = 0;
    float my_w = 0.f;
    if (j < deg) {
      my_idx = a.indices[lo + j];
      my_w = (a.values ? a.values[lo + j] : 1.0f) * rs;
      if (keep && keep[j] == 0) my_w = 0.f;
    }
    const int cnt = min(32, deg - base);
    for (int t = 0; t < cnt; t += G) {
      const int src = t + sub;
      const int idx = __shfl_sync(0xffffffffu, my_idx, src & 31);
      const float w = __shfl_sync(0xffffffffu, my_w, src & 31);
      if (src < cnt && w != 0.f) {
        float4* gp = reinterpret_cast<float4*>(a.gtable + (int64_t)idx * a.ld);
#pragma unroll
        for (int v = 0; v < NV; v++) {
          const int c4 = c + v * LPR;
          if (c4 < ld4) atomicAdd(gp + c4, make_float4(w * d[v].x, w * d[v].y, w * d[v].z, w * d[v].w));
        }
      }
    }
  }
}

// philox4x32-10, counter = (item, slot, step_lo, step_hi), key = (seed_lo, seed_hi).  Restated in numpy by
// oracle/philox.py; the mask is kept iff u >= q with u = (x0 >> 8) * 2^-24.
__device__ __forceinline__ uint32_t philox_first(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c0;
}

__global__ void __launch_bounds__(128) k_batch_prep(BatchPrepArgs a) {
  const int b = blockIdx.x;
  const int row = a.rows[b];
  const int64_t lo = a.indptr[row], hi = a.indptr[row + 1];
  const int koff = a.keep_off ? a.keep_off[b] : 0;
  const uint64_t step = a.step_dev ? (((uint64_t)a.step_dev[1] << 32) | a.step_dev[0]) : a.step;
  for (int64_t j = lo + threadIdx.x; j < hi; j += blockDim.x) {
    const int item = a.indices[j];
    if (a.count) atomicAdd(a.count + item, 1.0f);
    if (a.label_bits) atomicOr(a.label_bits + (int64_t)b * a.words_per_row + (item >> 5), 1u << (item & 31));
    if (a.keep_out) {
      const uint32_t x = philox_first((uint32_t)(item + a.item_offset), (uint32_t)(b + a.slot_offset), (uint32_t)step, (uint32_t)(step >> 32),
                                      (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      const float u = (float)(x >> 8) * (1.0f / 16777216.0f);
      a.keep_out[koff + (j - lo)] = (u < a.q) ? 0 : 1;
    }
  }
}

__global__ void __launch_bounds__(128) k_row_scatter(const int32_t* __restrict__ ids, const float* __restrict__ rows,
                                                     int n, int ld, float* __restrict__ gtable,
                                                     uint32_t* __restrict__ row_touched) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n) return;
  if (row_touched && lane == 0) atomicOr(row_touched + (ids[r] >> 5), 1u << (ids[r] & 31));
  float4* dst = reinterpret_cast<float4*>(gtable + (int64_t)ids[r] * ld);
  const float4* src = reinterpret_cast<const float4*>(rows + (int64_t)r * ld);
  for (int c = lane; c < (ld >> 2); c += 32) atomicAdd(dst + c, __ldg(src + c));
}

constexpr int kRowsPerBlock = 8;    // large batches; small ones use 4 so that more than a couple of CTAs run
__host__ __device__ inline int rows_per_block(int n) { return n >= 2048 ? kRowsPerBlock : 4; }

// dz1 = dh * h * (1-h) with dh = sum over split-K partials; column partial sums for db (cdae.py b gradient)
__global__ void k_dz1(const float* __restrict__ dh_part, int splits, const float* __restrict__ h,
                      float* __restrict__ dz1, int n, int ld, float* __restrict__ colpart) {
  const int r0 = blockIdx.x * rows_per_block(n);
  const int r1 = min(n, r0 + rows_per_block(n));
  const int64_t plane = (int64_t)n * ld;
  for (int col = threadIdx.x; col < ld; col += blockDim.x) {
    float cs = 0.f;
    for (int r = r0; r < r1; r++) {
      const int64_t o = (int64_t)r * ld + col;
      float dh = 0.f;
      for (int s = 0; s < splits; s++) dh += dh_part[s * plane + o];
      const float hv = h[o];
      const float g = dh * hv * (1.0f - hv);
      dz1[o] = g;
      cs += g;
    }
    colpart[(int64_t)blockIdx.x * ld + col] = cs;
  }
}

__global__ void k_colpart(const float* __restrict__ x, int n, int ld, float* __restrict__ colpart) {
  const int r0 = blockIdx.x * rows_per_block(n);
  const int r1 = min(n, r0 + rows_per_block(n));
  for (int col = threadIdx.x; col < ld; col += blockDim.x) {
    float cs = 0.f;
    for (int r = r0; r < r1; r++) cs += x[(int64_t)r * ld + col];
    colpart[(int64_t)blockIdx.x * ld + col] = cs;
  }
}

__global__ void __launch_bounds__(128) k_zero_rows(float* __restrict__ table, const int32_t* __restrict__ ids, int n,
                                                   int ld, uint32_t* __restrict__ row_touched) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n) return;
  if (row_touched && lane == 0) atomicAnd(row_touched + (ids[r] >> 5), ~(1u << (ids[r] & 31)));
  float4* dst = reinterpret_cast<float4*>(table + (int64_t)ids[r] * ld);
  for (int c = lane; c < (ld >> 2); c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void k_iota(int32_t* __restrict__ out, int n, int start) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = start + i;
}

__global__ void k_sum_planes(const float* __restrict__ part, int planes, int64_t n4, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < planes; s++) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(part) + s * n4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

__global__ void k_sigmoid_rows(float* __restrict__ x, int n, int ld, int width) {
  const int64_t total = (int64_t)n * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % ld);
    x[i] = (col < width) ? 1.0f / (1.0f + expf(-x[i])) : 0.f;
  }
}

// out[j] = sum_p part[p][j]: 32 columns per block, 8 thread rows stride over the partials, fixed-order final sum
__global__ void __launch_bounds__(256) k_reduce_partials(const float* __restrict__ part, int nparts, int ld,
                                                         float* __restrict__ out, int n) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (j < n)
    for (int p = ty; p < nparts; p += 8) s += part[(int64_t)p * ld + j];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < n) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; r++) t += sm[r][tx];
    out[j] = t;
  }
}

// ------------------------------------------------------------------------------------------ chunked (balanced) forms
// One CTA per sampled row leaves the kernel waiting for its longest row: with log-normal user activity the heaviest
// row of a 4096-user batch has ~35x the median degree, and ncu showed the SMs busy 46 % of the gather's duration.
// The chunked forms cut every CSR row into pieces of kChunk entries and let a persistent grid walk the pieces:
//   k_chunk_scan     chunk_off[b] = sum_{b' < b} ceil(deg(b') / kChunk)           (one CTA, running carry)
//   k_gather_chunks  partial sum of one piece, added to acc[b] with vector atomics (acc zeroed by the caller)
//   k_gather_finish  out = act(acc * scale + rowbias + bias)                      (in place)
//   k_scatter_chunks the scatter of one piece (atomics anyway)
// The piece -> row map is a binary search in a shared-memory copy of chunk_off.  Used by the CDAE training step;
// scoring and the DMF towers keep the one-CTA-per-row kernels (deterministic summation order).
constexpr int kChunk = 256;
constexpr int kMaxChunkRows = 16384;   // chunk_off is staged in shared memory: 64 KB at most

__global__ void __launch_bounds__(1024) k_chunk_scan(const int64_t* __restrict__ indptr, const int32_t* __restrict__ rows,
                                                     int n, int32_t* __restrict__ chunk_off) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 4096) {     // four consecutive rows per thread: their loads overlap
    const int b0 = base + 4 * threadIdx.x;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      v[k] = 0;
      if (b0 + k < n) {
        const int row = rows[b0 + k];
        v[k] = max(1, (int)((indptr[row + 1] - indptr[row] + kChunk - 1) / kChunk));   // >= 1: an empty row still owns dV[u]
      }
    }
    const int mine = v[0] + v[1] + v[2] + v[3];
    int x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const int incl = carry_s + x + (warp ? wsum[warp - 1] : 0);
    int run = incl - mine;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (b0 + k < n) chunk_off[b0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) chunk_off[n] = carry_s;
}

// largest b with coff[b] <= c  (coff non-decreasing, coff[0] = 0, c < coff[n])
__device__ __forceinline__ int chunk_row(const int32_t* coff, int n, int c) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (coff[mid] <= c) lo = mid; else hi = mid;
  }
  return lo;
}

template <int LPR, int NV>
__global__ void __launch_bounds__(kGatherThreads) k_gather_chunks(GatherArgs a, const int32_t* __restrict__ chunk_off,
                                                                  int n) {
  constexpr int G = 32 / LPR;
  extern __shared__ float4 red[];  // [kWarps * G][ld4], then coff[n + 1]
  const int ld4 = a.ld >> 2;
  int32_t* coff = reinterpret_cast<int32_t*>(red + kWarps * G * ld4);
  for (int i = threadIdx.x; i <= n; i += kGatherThreads) coff[i] = chunk_off[i];
  __syncthreads();
  const int total = coff[n];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, c = lane % LPR;
  for (int ch = blockIdx.x; ch < total; ch += gridDim.x) {
    const int b = chunk_row(coff, n, ch);
    const int first = (ch - coff[b]) * kChunk;
    const int row = a.rows[b];
    const int64_t lo = a.indptr[row] + first;
    const int deg = (int)min((int64_t)kChunk, (int64_t)(a.indptr[row + 1] - lo));
    const uint8_t* keep = a.keep ? a.keep + a.keep_off[b] + first : nullptr;
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    gather_accumulate<LPR, NV>(a, lo, deg, keep, warp, lane, acc);
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const int c4 = c + v * LPR;
      if (c4 < ld4) red[(warp * G + sub) * ld4 + c4] = acc[v];
    }
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < ld4; c4 += kGatherThreads) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kWarps * G; p++) {
        const float4 x = red[p * ld4 + c4];
        s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
      }
      atomicAdd(reinterpret_cast<float4*>(a.out + (int64_t)b * a.ld) + c4, s);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_gather_finish(GatherArgs a, int n) {
  const int64_t total = (int64_t)n * a.ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / a.ld), col = (int)(i % a.ld);
    const int row = a.rows[b];
    float rs = a.scale;
    if (a.row_scale) rs *= a.row_scale[row];
    float z = a.out[i] * rs;
    const int brow = a.bias_rows ? a.bias_rows[b] : row;
    if (brow >= 0) {
      if (a.rowbias) z += a.rowbias[(int64_t)brow * a.ld + col];
      if (a.bias) z += a.bias[col];
    }
    a.out[i] = (col < a.width) ? apply_act(z, a.act) : 0.f;
  }
}

template <int LPR, int NV>
__global__ void __launch_bounds__(kGatherThreads) k_scatter_chunks(ScatterArgs a, const int32_t* __restrict__ chunk_off,
                                                                   int n) {
  constexpr int G = 32 / LPR;
  extern __shared__ int32_t coff_s[];  // [n + 1]
  for (int i = threadIdx.x; i <= n; i += kGatherThreads) coff_s[i] = chunk_off[i];
  __syncthreads();
  const int total = coff_s[n];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, c = lane % LPR;
  const int ld4 = a.ld >> 2;
  for (int ch = blockIdx.x; ch < total; ch += gridDim.x) {
    const int b = chunk_row(coff_s, n, ch);
    const int piece = ch - coff_s[b];
    const int first = piece * kChunk;
    const int row = a.rows[b];
    const int64_t lo = a.indptr[row] + first;
    const int deg = (int)min((int64_t)kChunk, (int64_t)(a.indptr[row + 1] - lo));
    const uint8_t* keep = a.keep ? a.keep + a.keep_off[b] + first : nullptr;
    float rs = a.scale;
    if (a.row_scale) rs *= a.row_scale[row];
    float4 d[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const int c4 = c + v * LPR;
      d[v] = (c4 < ld4) ? ldg4(a.d + (int64_t)b * a.ld + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int brow = a.bias_rows ? a.bias_rows[b] : row;
    if (a.growbias && brow >= 0 && piece == 0 && warp == 0 && sub == 0) {
      if (a.row_touched && c == 0) atomicOr(a.row_touched + (brow >> 5), 1u << (brow & 31));
#pragma unroll
      for (int v = 0; v < NV; v++) {
        const int c4 = c + v * LPR;
        if (c4 < ld4) atomicAdd(reinterpret_cast<float4*>(a.growbias + (int64_t)brow * a.ld) + c4, d[v]);
      }
    }
    for (int base = warp * 32; base < deg; base += kWarps * 32) {
      const int j = base + lane;
      int my_idx = 0;
      float my_w = 0.f;
      if (j < deg) {
        my_idx = a.indices[lo + j];
        my_w = (a.values ? a.values[lo + j] : 1.0f) * rs;
        if (keep && keep[j] == 0) my_w = 0.f;
      }
      const int cnt = min(32, deg - base);
      for (int t = 0; t < cnt; t += G) {
        const int src = t + sub;
        const int idx = __shfl_sync(0xffffffffu, my_idx, src & 31);
        const float w = __shfl_sync(0xffffffffu, my_w, src & 31);
        if (src < cnt && w != 0.f) {
          float4* gp = reinterpret_cast<float4*>(a.gtable + (int64_t)idx * a.ld);
#pragma unroll
          for (int v = 0; v < NV; v++) {
            const int c4 = c + v * LPR;
            if (c4 < ld4) atomicAdd(gp + c4, make_float4(w * d[v].x, w * d[v].y, w * d[v].z, w * d[v].w));
          }
        }
      }
    }
  }
}

template <template <int, int> class Launcher, typename Args>
int dispatch_lpr(drb_ctx* ctx, const Args& a, int n, int ld, const char* name) {
  const int ld4 = ld >> 2;
  if (ld4 <= 4) return Launcher<4, 1>::run(ctx, a, n, name);
  if (ld4 <= 8) return Launcher<8, 1>::run(ctx, a, n, name);
  if (ld4 <= 16) return Launcher<16, 1>::run(ctx, a, n, name);
  if (ld4 <= 32) return Launcher<32, 1>::run(ctx, a, n, name);
  if (ld4 <= 64) return Launcher<32, 2>::run(ctx, a, n, name);
  if (ld4 <= 96) return Launcher<32, 3>::run(ctx, a, n, name);
  if (ld4 <= 128) return Launcher<32, 4>::run(ctx, a, n, name);
  return drb_fail(DRB_E_INVALID, "%s: row width %d exceeds 512 floats", name, ld);
}

template <int LPR, int NV>
struct GatherLauncher {
  static int run(drb_ctx* ctx, const GatherArgs& a, int n, const char* name) {
    const size_t smem = (size_t)kWarps * (32 / LPR) * a.ld * sizeof(float);
    drb_prof_scope prof_(ctx, "k_gather");
    k_gather<LPR, NV><<<n, kGatherThreads, smem, ctx->stream>>>(a, a);
    DRB_LAUNCH_CHECK(ctx, name);
    return DRB_OK;
  }
};
template <int LPR, int NV>
struct GatherPairLauncher {
  static int run(drb_ctx* ctx, const GatherPair& p, int n, const char* name) {
    const size_t smem = (size_t)kWarps * (32 / LPR) * p.a[0].ld * sizeof(float);
    drb_prof_scope prof_(ctx, "k_gather");
    k_gather<LPR, NV><<<dim3(n, 2), kGatherThreads, smem, ctx->stream>>>(p.a[0], p.a[1]);
    DRB_LAUNCH_CHECK(ctx, name);
    return DRB_OK;
  }
};
template <int LPR, int NV>
struct ScatterPairLauncher {
  static int run(drb_ctx* ctx, const ScatterPair& p, int n, const char* name) {
    drb_prof_scope prof_(ctx, "k_scatter");
    k_scatter<LPR, NV><<<dim3(n, 2), kGatherThreads, 0, ctx->stream>>>(p.a[0], p.a[1]);
    DRB_LAUNCH_CHECK(ctx, name);
    return DRB_OK;
  }
};
template <int LPR, int NV>
struct ScatterLauncher {
  static int run(drb_ctx* ctx, const ScatterArgs& a, int n, const char* name) {
    drb_prof_scope prof_(ctx, "k_scatter");
    k_scatter<LPR, NV><<<n, kGatherThreads, 0, ctx->stream>>>(a, a);
    DRB_LAUNCH_CHECK(ctx, name);
    return DRB_OK;
  }
};

}  // namespace

template <int LPR, int NV>
struct GatherChunksLauncher {
  static int run(drb_ctx* ctx, const GatherArgs& a, int n, const char* name) {
    const size_t smem = (size_t)kWarps * (32 / LPR) * a.ld * sizeof(float) + (size_t)(n + 1) * sizeof(int32_t);
    auto kern = k_gather_chunks<LPR, NV>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return drb_fail(DRB_E_CUDA, "%s: %zu bytes of shared memory: %s", name, smem, cudaGetErrorString(e));
    }
    drb_prof_scope prof_(ctx, "k_gather");
    kern<<<ctx->sm_count * 8, kGatherThreads, smem, ctx->stream>>>(a, a.chunk_off, n);
    DRB_LAUNCH_CHECK(ctx, name);
    return DRB_OK;
  }
};
template <int LPR, int NV>
struct ScatterChunksLauncher {
  static int run(drb_ctx* ctx, const ScatterArgs& a, int n, const char* name) {
    const size_t smem = (size_t)(n + 1) * sizeof(int32_t);
    auto kern = k_scatter_chunks<LPR, NV>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return drb_fail(DRB_E_CUDA, "%s: %zu bytes of shared memory: %s", name, smem, cudaGetErrorString(e));
    }
    drb_prof_scope prof_(ctx, "k_scatter");
    kern<<<ctx->sm_count * 8, kGatherThreads, smem, ctx->stream>>>(a, a.chunk_off, n);
    DRB_LAUNCH_CHECK(ctx, name);
    return DRB_OK;
  }
};

int launch_chunk_scan(drb_ctx* ctx, const int64_t* indptr, const int32_t* rows, int n, int32_t* chunk_off) {
  if (n <= 0) return DRB_OK;
  drb_prof_scope prof_(ctx, "k_chunk_scan");
  k_chunk_scan<<<1, 1024, 0, ctx->stream>>>(indptr, rows, n, chunk_off);
  DRB_LAUNCH_CHECK(ctx, "k_chunk_scan");
  return DRB_OK;
}

int launch_gather(drb_ctx* ctx, const GatherArgs& a, int n) {
  if (n <= 0) return DRB_OK;
  if (a.ld % 4) return drb_fail(DRB_E_INVALID, "gather: ld must be a multiple of 4");
  if (a.chunk_off && n <= kMaxChunkRows) {   // balanced form: zero, accumulate pieces, finish
    cudaError_t e = cudaMemsetAsync(a.out, 0, (size_t)n * a.ld * sizeof(float), ctx->stream);
    if (e != cudaSuccess) { ctx->sticky = (int)e; return drb_fail(DRB_E_CUDA, "gather: memset failed: %s", cudaGetErrorString(e)); }
    int r = dispatch_lpr<GatherChunksLauncher>(ctx, a, n, a.ld, "k_gather_chunks");
    if (r) return r;
    const int64_t total = (int64_t)n * a.ld;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 8);
    drb_prof_scope prof_(ctx, "k_gather_finish");
    k_gather_finish<<<blocks, 256, 0, ctx->stream>>>(a, n);
    DRB_LAUNCH_CHECK(ctx, "k_gather_finish");
    return DRB_OK;
  }
  return dispatch_lpr<GatherLauncher>(ctx, a, n, a.ld, "k_gather");
}

int launch_gather_pair(drb_ctx* ctx, const GatherArgs& a0, const GatherArgs& a1, int n) {
  if (n <= 0) return DRB_OK;
  if (a0.ld != a1.ld || a0.ld % 4 || a0.chunk_off || a1.chunk_off)
    return drb_fail(DRB_E_INVALID, "gather pair: both argument sets need the same row width");
  GatherPair p{{a0, a1}};
  return dispatch_lpr<GatherPairLauncher>(ctx, p, n, a0.ld, "k_gather");
}

int launch_scatter_pair(drb_ctx* ctx, const ScatterArgs& a0, const ScatterArgs& a1, int n) {
  if (n <= 0) return DRB_OK;
  if (a0.ld != a1.ld || a0.ld % 4 || a0.chunk_off || a1.chunk_off)
    return drb_fail(DRB_E_INVALID, "scatter pair: both argument sets need the same row width");
  ScatterPair p{{a0, a1}};
  return dispatch_lpr<ScatterPairLauncher>(ctx, p, n, a0.ld, "k_scatter");
}

int launch_scatter(drb_ctx* ctx, const ScatterArgs& a, int n) {
  if (n <= 0) return DRB_OK;
  if (a.ld % 4) return drb_fail(DRB_E_INVALID, "scatter: ld must be a multiple of 4");
  if (a.chunk_off && n <= kMaxChunkRows) return dispatch_lpr<ScatterChunksLauncher>(ctx, a, n, a.ld, "k_scatter_chunks");
  return dispatch_lpr<ScatterLauncher>(ctx, a, n, a.ld, "k_scatter");
}

int launch_row_scatter(drb_ctx* ctx, const int32_t* ids, const float* rows, int n, int ld, float* gtable,
                       uint32_t* row_touched) {
  if (n <= 0) return DRB_OK;
  drb_prof_scope prof_(ctx, "k_row_scatter");
  k_row_scatter<<<(n * 32 + 127) / 128, 128, 0, ctx->stream>>>(ids, rows, n, ld, gtable, row_touched);
  DRB_LAUNCH_CHECK(ctx, "k_row_scatter");
  return DRB_OK;
}

int launch_batch_prep(drb_ctx* ctx, const BatchPrepArgs& a, int n) {
  if (n <= 0) return DRB_OK;
  drb_prof_scope prof_(ctx, "k_batch_prep");
  k_batch_prep<<<n, 128, 0, ctx->stream>>>(a);
  DRB_LAUNCH_CHECK(ctx, "k_batch_prep");
  return DRB_OK;
}

int launch_dz1(drb_ctx* ctx, const float* dh_part, int splits, const float* h, float* dz1, int n, int ld,
               float* colpart) {
  const int nblk = (n + rows_per_block(n) - 1) / rows_per_block(n);
  drb_prof_scope prof_(ctx, "k_dz1");
  k_dz1<<<nblk, min(512, (int)drb_round_up(ld, 32)), 0, ctx->stream>>>(dh_part, splits, h, dz1, n, ld, colpart);
  DRB_LAUNCH_CHECK(ctx, "k_dz1");
  return nblk;
}

int launch_colpart(drb_ctx* ctx, const float* x, int n, int ld, float* colpart) {
  const int nblk = (n + rows_per_block(n) - 1) / rows_per_block(n);
  drb_prof_scope prof_(ctx, "k_colpart");
  k_colpart<<<nblk, min(512, (int)drb_round_up(ld, 32)), 0, ctx->stream>>>(x, n, ld, colpart);
  DRB_LAUNCH_CHECK(ctx, "k_colpart");
  return nblk;
}

int launch_zero_rows(drb_ctx* ctx, float* table, const int32_t* ids, int n, int ld, uint32_t* row_touched) {
  if (n <= 0) return DRB_OK;
  drb_prof_scope prof_(ctx, "k_zero_rows");
  k_zero_rows<<<(n * 32 + 127) / 128, 128, 0, ctx->stream>>>(table, ids, n, ld, row_touched);
  DRB_LAUNCH_CHECK(ctx, "k_zero_rows");
  return DRB_OK;
}

int launch_iota(drb_ctx* ctx, int32_t* out, int n, int start) {
  if (n <= 0) return DRB_OK;
  drb_prof_scope prof_(ctx, "k_iota");
  k_iota<<<(n + 255) / 256, 256, 0, ctx->stream>>>(out, n, start);
  DRB_LAUNCH_CHECK(ctx, "k_iota");
  return DRB_OK;
}

int launch_sum_planes(drb_ctx* ctx, const float* part, int planes, int64_t n_elems, float* out) {
  const int64_t n4 = n_elems / 4;   // n_elems % 4 == 0 (ld is a multiple of 4)
  const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)ctx->sm_count * 8);
  drb_prof_scope prof_(ctx, "k_sum_planes");
  k_sum_planes<<<std::max(1, blocks), 256, 0, ctx->stream>>>(part, planes, n4, out);
  DRB_LAUNCH_CHECK(ctx, "k_sum_planes");
  return DRB_OK;
}

int launch_sigmoid_rows(drb_ctx* ctx, float* x, int n, int ld, int width) {
  const int64_t total = (int64_t)n * ld;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 8);
  drb_prof_scope prof_(ctx, "k_sigmoid_rows");
  k_sigmoid_rows<<<std::max(1, blocks), 256, 0, ctx->stream>>>(x, n, ld, width);
  DRB_LAUNCH_CHECK(ctx, "k_sigmoid_rows");
  return DRB_OK;
}

int launch_reduce_partials(drb_ctx* ctx, const float* part, int nparts, int ld, float* out, int n) {
  if (n <= 0) return DRB_OK;
  drb_prof_scope prof_(ctx, "k_reduce_partials");
  k_reduce_partials<<<(n + 31) / 32, 256, 0, ctx->stream>>>(part, nparts, ld, out, n);
  DRB_LAUNCH_CHECK(ctx, "k_reduce_partials");
  return DRB_OK;
}
