// Sampled-output layer of the CDAE step for catalogs whose dense output layer is out of reach (BASELINE.json configs[4]:
// 10 M users x 1 M items, item weights sharded over the GPUs): for every sampled user only its positives and a fixed
// number of uniformly drawn items are scored, with per-user labels, and the same kernel runs the backward pass.
//
// This is an EXTENSION of the reference: DRecPy/Recommender/cdae.py:76 always scores all I items (6*K*I flops per
// sampled user = 1.5 GFLOP at K = 256, I = 1 M); its definition is pinned by oracle/cdae.py: CDAESampledOracle.
// What stays the reference's: the hidden layer (cdae.py:59-75), Keras BCE / MSE term by term, dense Adam + L2.
//
// One CTA per sampled user (4 warps).  A warp takes one output item at a time: the lanes cover the item's row of W'^T
// (item-major, K contiguous) with 128-bit loads, the dot product is a shuffle reduction, then the same lanes
//   - accumulate dh_b += dz * W'_j in registers (reduced over the 4 warps at the end, fixed order),
//   - add dz * h_b into dW'^T_j with vector atomics (red.global.add.v4.f32) and dz into db'_j.
// Per output item the kernel moves 4K bytes of W' in and 4K bytes of gradient out: HBM / L2-atomic bound
// (8*K*S bytes per user).  Negatives are philox4x32-10 draws (counter = draw, slot, step; key = seed ^
// group), so the oracle reproduces them; a drawn item that is one of the user's positives is labelled positive.
#include "kernels.h"

namespace {

constexpr float KERAS_EPS = 1e-7f;
constexpr int kSoThreads = 128, kSoWarps = 4;

__device__ __forceinline__ uint32_t so_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
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
__device__ __forceinline__ float so_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ bool so_contains(const int32_t* lo, int n, int32_t x) {
  int a = 0, b = n;
  while (a < b) {
    const int mid = (a + b) >> 1;
    if (lo[mid] < x) a = mid + 1; else b = mid;
  }
  return a < n && lo[a] == x;
}

template <int NV>
__global__ void __launch_bounds__(kSoThreads) k_sampled_out(SampledOutArgs a) {
  __shared__ float4 s_dh[kSoWarps][NV * 32];
  __shared__ float s_loss[kSoWarps];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld4 = a.ld >> 2;
  const int row = a.rows[b];
  const int64_t lo = a.indptr[row];
  const int deg = (int)(a.indptr[row + 1] - lo);
  const int32_t* pos = a.indices + lo;
  const uint64_t step = a.step_dev ? (((uint64_t)a.step_dev[1] << 32) | a.step_dev[0]) : a.step;
  float4 h[NV], dh[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) {
    const int c4 = lane + 32 * v;
    h[v] = c4 < ld4 ? __ldg(reinterpret_cast<const float4*>(a.h + (int64_t)b * a.ld) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    dh[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float loss = 0.f;
  const int n_neg = a.n_groups * a.neg_per_group;
  const int per = (a.n_items_total + a.n_groups_total - 1) / a.n_groups_total;     // items per group (global partition)
  for (int q = warp; q < deg + n_neg; q += kSoWarps) {
    int item;
    float t;
    if (q < deg) {
      item = pos[q];
      t = 1.0f;
    } else {
      const int gl = (q - deg) / a.neg_per_group, d = (q - deg) % a.neg_per_group;
      const int g = a.group_id0 + gl;                                           // global group id
      const int g_lo = min(a.n_items_total, g * per), g_hi = min(a.n_items_total, (g + 1) * per);
      const uint32_t x = so_philox((uint32_t)d, (uint32_t)(b + a.slot_offset), (uint32_t)step, (uint32_t)(step >> 32),
                                   (uint32_t)a.seed ^ (0x9E3779B9u * (uint32_t)(g + 1)), (uint32_t)(a.seed >> 32));
      item = g_lo + (int)(x % (uint32_t)max(g_hi - g_lo, 1)) - a.item_offset;   // local item id
      t = so_contains(pos, deg, item) ? 1.0f : 0.0f;
    }
    const float4* wr = reinterpret_cast<const float4*>(a.w2t + (int64_t)item * a.ld);
    float4 w[NV];
    float dot = 0.f;
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const int c4 = lane + 32 * v;
      w[v] = c4 < ld4 ? __ldg(wr + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      dot = fmaf(h[v].x, w[v].x, dot); dot = fmaf(h[v].y, w[v].y, dot);
      dot = fmaf(h[v].z, w[v].z, dot); dot = fmaf(h[v].w, w[v].w, dot);
    }
    dot = so_warp_sum(dot);
    const float z = dot + __ldg(a.b2 + item);
    const float p = 1.0f / (1.0f + expf(-z));
    float gz, lt;
    if (a.loss_kind == DRB_LOSS_BCE) {      // Keras-2 binary_crossentropy term by term, as in the dense epilogue
      const float one_m = 1.0f - KERAS_EPS;
      const float pc = fminf(fmaxf(p, KERAS_EPS), one_m);
      const float da = pc + KERAS_EPS, db = 1.0f - pc + KERAS_EPS;
      lt = -(t * logf(da) + (1.0f - t) * logf(db));
      const bool inside = (p >= KERAS_EPS) && (p <= one_m);
      gz = inside ? -(t / da - (1.0f - t) / db) * a.inv_count * (p * (1.0f - p)) : 0.f;
    } else {
      lt = (p - t) * (p - t);
      gz = 2.0f * (p - t) * a.inv_count * (p * (1.0f - p));
    }
    loss += lt;
    float4* gw = reinterpret_cast<float4*>(a.g_w2t + (int64_t)item * a.ld);
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const int c4 = lane + 32 * v;
      dh[v].x = fmaf(gz, w[v].x, dh[v].x); dh[v].y = fmaf(gz, w[v].y, dh[v].y);
      dh[v].z = fmaf(gz, w[v].z, dh[v].z); dh[v].w = fmaf(gz, w[v].w, dh[v].w);
      if (c4 < ld4) atomicAdd(gw + c4, make_float4(gz * h[v].x, gz * h[v].y, gz * h[v].z, gz * h[v].w));
    }
    if (lane == 0) atomicAdd(a.g_b2 + item, gz);
  }
#pragma unroll
  for (int v = 0; v < NV; v++) s_dh[warp][lane + 32 * v] = dh[v];
  if (lane == 0) s_loss[warp] = loss;          // every lane of the warp holds the same terms
  __syncthreads();
  for (int c4 = threadIdx.x; c4 < ld4; c4 += kSoThreads) {
    float4 s = s_dh[0][c4];
#pragma unroll
    for (int wv = 1; wv < kSoWarps; wv++) {
      const float4 x = s_dh[wv][c4];
      s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
    }
    reinterpret_cast<float4*>(a.dh + (int64_t)b * a.ld)[c4] = s;
  }
  if (threadIdx.x == 0) a.loss_part[b] = s_loss[0] + s_loss[1] + s_loss[2] + s_loss[3];
}

}  // namespace

int launch_sampled_out(drb_ctx* ctx, const SampledOutArgs& a, int n) {
  if (n <= 0) return DRB_OK;
  if (a.ld % 4 || a.ld > 512) return drb_fail(DRB_E_INVALID, "sampled output layer: hidden width %d not supported", a.ld);
  if (a.n_groups < 1 || a.neg_per_group < 0 || a.n_groups_total < a.n_groups)
    return drb_fail(DRB_E_INVALID, "sampled output layer: bad negative-sampling groups");
  const int nv = (a.ld / 4 + 31) / 32;
  drb_prof_scope prof_(ctx, "k_sampled_out");
  if (nv == 1) k_sampled_out<1><<<n, kSoThreads, 0, ctx->stream>>>(a);
  else if (nv == 2) k_sampled_out<2><<<n, kSoThreads, 0, ctx->stream>>>(a);
  else if (nv == 3) k_sampled_out<3><<<n, kSoThreads, 0, ctx->stream>>>(a);
  else k_sampled_out<4><<<n, kSoThreads, 0, ctx->stream>>>(a);
  DRB_LAUNCH_CHECK(ctx, "k_sampled_out");
  return DRB_OK;
}
