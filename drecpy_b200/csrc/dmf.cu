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
  const float dp = inside ? -(t / da_ - (1.0f - t) / db_) / (float)h.n : 0.f;
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

}  // namespace

int launch_dmf_head(drb_ctx* ctx, const DmfHeadArgs& a) {
  if (a.n <= 0) return DRB_OK;
  const int blocks = (a.n * 32 + 127) / 128;
  drb_prof_scope prof_(ctx, "k_dmf_head");
  k_dmf_head<<<blocks, 128, 0, ctx->stream>>>(a);
  DRB_LAUNCH_CHECK(ctx, "k_dmf_head");
  return DRB_OK;
}
