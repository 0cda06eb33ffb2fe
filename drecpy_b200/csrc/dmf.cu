// K5 head of the DMF step (sm_100a): cosine of the two tower outputs, max(1e-6, .), Keras BCE with soft labels,
// and the backward pass through the l2-normalisation and the last ReLU of each tower.  One warp per (u, i) pair.
//
// Replaces (reference, DRecPy/): Recommender/dmf.py:88-99 (_predict_batch / _compute_batch_loss) and the matching
// part of tape.gradient (recommender_abc.py:203).  tf.nn.l2_normalize = x * rsqrt(max(sum x^2, 1e-12));
// tf.maximum(1e-6, c) passes the gradient to c only where c > 1e-6.
#include "kernels.h"

namespace {

constexpr float KERAS_EPS = 1e-7f;
constexpr float L2N_EPS = 1e-12f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(128) k_dmf_head(DmfHeadArgs h) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= h.n) return;
  const float* a = h.a + (int64_t)warp * h.ld;
  const float* e = h.e + (int64_t)warp * h.ld;
  float ssa = 0.f, sse = 0.f, dot = 0.f;
  for (int k = lane; k < h.width; k += 32) {
    const float av = a[k], ev = e[k];
    ssa = fmaf(av, av, ssa);
    sse = fmaf(ev, ev, sse);
    dot = fmaf(av, ev, dot);
  }
  ssa = warp_sum(ssa); sse = warp_sum(sse); dot = warp_sum(dot);
  const float inva = 1.0f / sqrtf(fmaxf(ssa, L2N_EPS)), inve = 1.0f / sqrtf(fmaxf(sse, L2N_EPS));
  const float c = dot * inva * inve;
  const float p = fmaxf(1e-6f, c);
  if (h.p_out && lane == 0) h.p_out[warp] = p;
  if (!h.labels) return;
  const float t = h.labels[warp];
  const float one_m = 1.0f - KERAS_EPS;
  const float pc = fminf(fmaxf(p, KERAS_EPS), one_m);
  const float da_ = pc + KERAS_EPS, db_ = 1.0f - pc + KERAS_EPS;
  if (lane == 0) h.loss_part[warp] = -(t * logf(da_) + (1.0f - t) * logf(db_));
  const bool inside = (p >= KERAS_EPS) && (p <= one_m);
  const float dp = inside ? -(t / da_ - (1.0f - t) / db_) / (float)(h.n_global > 0 ? h.n_global : h.n) : 0.f;
  const float dc = (c > 1e-6f) ? dp : 0.f;
  for (int k = lane; k < h.ld; k += 32) {
    float ga = 0.f, ge = 0.f;
    if (k < h.width) {
      const float av = a[k], ev = e[k];
      const float ah = av * inva, eh = ev * inve;
      // d a = (d a_hat - a_hat * <a_hat, d a_hat>) * inv  with d a_hat = dc * e_hat (and symmetrically)
      ga = (ssa >= L2N_EPS) ? (dc * eh - ah * dc * c) * inva : dc * eh * inva;
      ge = (sse >= L2N_EPS) ? (dc * ah - eh * dc * c) * inve : dc * ah * inve;
      if (!(av > 0.f)) ga = 0.f;  // last layer is Dense(relu) (dmf.py:50-51,57-58)
      if (!(ev > 0.f)) ge = 0.f;
    }
    h.da[(int64_t)warp * h.ld + k] = ga;
    h.de[(int64_t)warp * h.ld + k] = ge;
  }
}

// Dense part of the backward pass of a two-layer tower, both towers in one launch (blockIdx.y), kRows batch rows per CTA:
//   dpre0[r] = (dpre1[r] K1^T) * [act0[r] > 0];  dK1 += act0^T dpre1;  db1 += sum_r dpre1[r];  db0 += sum_r dpre0[r]
// Replaces tape.gradient through the second Dense(relu) of dmf.py:50-51,57-58 (six launches per tower before: two
// small GEMMs, two column-sum pairs).  Partial sums live in registers per CTA and reach the gradients by atomics.
constexpr int kBwdRows = 8, kBwdThreads = 256, kBwdMaxW0 = 128, kBwdMaxW1 = 64;

__global__ void __launch_bounds__(kBwdThreads) k_dmf_tower_bwd(DmfTowerBwd t0, DmfTowerBwd t1, int n) {
  // all kBwdRows rows of the CTA are staged at once (one barrier): the C2 step is latency bound, and a row-by-row
  // loop with two barriers per row was the longest kernel of the step
  __shared__ float s_act[kBwdRows][kBwdMaxW0], s_d1[kBwdRows][kBwdMaxW1], s_gb0[kBwdMaxW0];
  const DmfTowerBwd t = blockIdx.y ? t1 : t0;
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * kBwdRows, nr = min(n, r0 + kBwdRows) - r0;
  for (int i = tid; i < kBwdRows * t.w0; i += kBwdThreads) {
    const int r = i / t.w0, c = i % t.w0;
    s_act[r][c] = r < nr ? t.act0[(int64_t)(r0 + r) * t.ld0 + c] : 0.f;
  }
  for (int i = tid; i < kBwdRows * t.w1; i += kBwdThreads) {
    const int r = i / t.w1, j = i % t.w1;
    s_d1[r][j] = r < nr ? t.dpre1[(int64_t)(r0 + r) * t.ld1 + j] : 0.f;
  }
  if (tid < kBwdMaxW0) s_gb0[tid] = 0.f;
  __syncthreads();
  // dpre0[r][c] = (dpre1[r] . K1[c]) * [act0[r][c] > 0]; pad columns (c >= w0) are written as zeros
  for (int o = tid; o < nr * t.ld0; o += kBwdThreads) {
    const int r = o / t.ld0, c = o % t.ld0;
    float d = 0.f;
    if (c < t.w0 && s_act[r][c] > 0.f) {
      const float* krow = t.k1 + (int64_t)c * t.ld1;
      for (int j = 0; j < t.w1; j++) d = fmaf(s_d1[r][j], __ldg(krow + j), d);
      atomicAdd(&s_gb0[c], d);
    }
    t.dpre0[(int64_t)(r0 + r) * t.ld0 + c] = d;
  }
  // dK1[c][j] += sum_r act0[r][c] dpre1[r][j]  (rows past the batch were staged as zeros)
  const int n_el = t.w0 * t.w1;
  for (int e = tid; e < n_el; e += kBwdThreads) {
    const int c = e / t.w1, j = e % t.w1;
    float g = 0.f;
#pragma unroll
    for (int r = 0; r < kBwdRows; r++) g = fmaf(s_act[r][c], s_d1[r][j], g);
    if (g != 0.f) atomicAdd(t.g_k1 + (int64_t)c * t.ld1 + j, g);
  }
  if (tid < t.w1) {
    float gb1 = 0.f;
#pragma unroll
    for (int r = 0; r < kBwdRows; r++) gb1 += s_d1[r][tid];
    atomicAdd(t.g_b1 + tid, gb1);
  }
  __syncthreads();
  if (tid < t.w0) atomicAdd(t.g_b0 + tid, s_gb0[tid]);
}

}  // namespace

int launch_dmf_tower_bwd(drb_ctx* ctx, const DmfTowerBwd& t0, const DmfTowerBwd& t1, int n) {
  if (n <= 0) return DRB_OK;
  for (const DmfTowerBwd* t : {&t0, &t1})
    if (t->w0 > kBwdMaxW0 || t->w1 > kBwdMaxW1 || t->ld0 > kBwdThreads)
      return drb_fail(DRB_E_INVALID, "dmf tower bwd: layer widths %d -> %d exceed the fused kernel's limits", t->w0, t->w1);
  drb_prof_scope prof_(ctx, "k_dmf_tower_bwd");
  k_dmf_tower_bwd<<<dim3((n + kBwdRows - 1) / kBwdRows, 2), kBwdThreads, 0, ctx->stream>>>(t0, t1, n);
  DRB_LAUNCH_CHECK(ctx, "k_dmf_tower_bwd");
  return DRB_OK;
}

bool dmf_tower_bwd_fits(int w0, int ld0, int w1) { return w0 <= kBwdMaxW0 && w1 <= kBwdMaxW1 && ld0 <= kBwdThreads; }

int launch_dmf_head(drb_ctx* ctx, const DmfHeadArgs& a) {
  if (a.n <= 0) return DRB_OK;
  const int blocks = (a.n * 32 + 127) / 128;
  drb_prof_scope prof_(ctx, "k_dmf_head");
  k_dmf_head<<<blocks, 128, 0, ctx->stream>>>(a);
  DRB_LAUNCH_CHECK(ctx, "k_dmf_head");
  return DRB_OK;
}
