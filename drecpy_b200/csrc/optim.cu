// K4: fused dense Adam + L2 over the whole parameter arena in one launch (sm_100a, HBM-bound: 28 B / parameter).
//
// Replaces (reference, DRecPy/): Recommender/recommender_abc.py:328-334 (one optimizer.apply_gradients per
// variable => the Adam step counter advances once per variable), Keras Adam's dense update
// (lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2); w -= lr_t*m/(sqrt(v)+eps)), and the L2
// terms of Recommender/cdae.py:81-82 (reg/B * l2_loss) / Keras regularizers.l2 (dmf.py:47), whose gradient is
// folded into g here and whose value (at the pre-update weights) is reduced into reg_part for the reported loss.
#include "kernels.h"

#include <cmath>

namespace {

constexpr int kAdamThreads = 256;

__global__ void __launch_bounds__(kAdamThreads) k_adam(AdamArgs a, int64_t total4) {
  __shared__ float sred[kAdamThreads];
  float reg = 0.f;
  const float omb1 = 1.0f - a.beta1, omb2 = 1.0f - a.beta2;
  for (int64_t i = a.seg[0].off4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (int64_t)gridDim.x * blockDim.x) {
    int s = 0;
#pragma unroll 1
    while (s + 1 < a.nseg && i >= a.seg[s + 1].off4) s++;
    if (i >= a.seg[s].off4 + a.seg[s].n4) continue;  // alignment gap between segments
    const float alpha = a.alpha_dev ? __ldg(a.alpha_dev + a.seg[s].alpha_idx) : a.seg[s].alpha;
    const float l2 = a.seg[s].l2, regw = a.seg[s].regw;
    float4 w = reinterpret_cast<float4*>(a.w)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f);
    bool read_g = true;
    if (s == a.row_seg) {          // rows of this table without a gradient this step hold zeros: do not read them
      const uint32_t row = (uint32_t)(i - a.seg[s].off4) / (uint32_t)a.row_len4;
      read_g = (__ldg(a.row_mask + (row >> 5)) >> (row & 31)) & 1u;
    }
    if (read_g) g0 = __ldg(reinterpret_cast<const float4*>(a.g) + i);
    reg += regw * (w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w);
#define DRB_ADAM1(c)                                           \
    {                                                          \
      const float g = g0.c + l2 * w.c;                         \
      m.c += (g - m.c) * omb1;                                 \
      v.c += (g * g - v.c) * omb2;                             \
      w.c -= alpha * m.c / (sqrtf(v.c) + a.eps);               \
    }
    DRB_ADAM1(x) DRB_ADAM1(y) DRB_ADAM1(z) DRB_ADAM1(w)
#undef DRB_ADAM1
    reinterpret_cast<float4*>(a.w)[i] = w;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
  }
  sred[threadIdx.x] = reg;
  __syncthreads();
  for (int s = kAdamThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sred[threadIdx.x] += sred[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.reg_part[blockIdx.x] = sred[0];
  if (!a.fin_loss_out) return;
  // fused loss finalisation: the last block to get here sums the loss terms and the regularisation partials
  __shared__ bool last;
  __shared__ double sa[kAdamThreads], sb[kAdamThreads];
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicInc(a.fin_ticket, gridDim.x - 1) == gridDim.x - 1;      // wraps to 0: ready for the next launch
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int n_reg = a.fin_n_reg > 0 ? a.fin_n_reg : (int)gridDim.x;
  double la = 0.0, lb = 0.0;
  for (int i = threadIdx.x; i < a.fin_n_loss; i += kAdamThreads) la += (double)__ldcg(a.fin_loss_part + i);
  const float* regs = a.fin_reg_part ? a.fin_reg_part : a.reg_part;
  for (int i = threadIdx.x; i < n_reg; i += kAdamThreads) lb += (double)__ldcg(regs + i);
  sa[threadIdx.x] = la;
  sb[threadIdx.x] = lb;
  __syncthreads();
  for (int s = kAdamThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sa[threadIdx.x] += sa[threadIdx.x + s];
      sb[threadIdx.x] += sb[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.fin_loss_out[0] = (float)(sa[0] * (double)a.fin_scale) + (float)sb[0];
    a.fin_loss_out[1] = (float)(sa[0] * (double)a.fin_scale);
  }
}

__global__ void __launch_bounds__(256) k_finalize_loss(const float* __restrict__ loss_part, int n_loss, float scale,
                                                       const float* __restrict__ reg_part, int n_reg,
                                                       float* __restrict__ loss_out) {
  __shared__ double sa[256], sb[256];
  double la = 0.0, lb = 0.0;
  for (int i = threadIdx.x; i < n_loss; i += 256) la += (double)loss_part[i];
  for (int i = threadIdx.x; i < n_reg; i += 256) lb += (double)reg_part[i];
  sa[threadIdx.x] = la;
  sb[threadIdx.x] = lb;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sa[threadIdx.x] += sa[threadIdx.x + s];
      sb[threadIdx.x] += sb[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    loss_out[0] = (float)(sa[0] * (double)scale) + (float)sb[0];   // reported loss: batch term + regularisation
    loss_out[1] = (float)(sa[0] * (double)scale);                  // batch term alone (summed over ranks when data parallel)
  }
}

struct Scalars8 { float v[8]; };
__global__ void k_set_scalars(float* __restrict__ dst, Scalars8 s, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = s.v[threadIdx.x];
}

}  // namespace

int launch_set_scalars(drb_ctx* ctx, float* dst, const float* vals, int n) {
  if (n < 0 || n > 8) return drb_fail(DRB_E_INVALID, "launch_set_scalars: n must be in [0, 8]");
  Scalars8 s{};
  for (int i = 0; i < n; i++) s.v[i] = vals[i];
  k_set_scalars<<<1, 32, 0, ctx->stream>>>(dst, s, n);
  DRB_LAUNCH_CHECK(ctx, "k_set_scalars");
  return DRB_OK;
}

float drb_adam_alpha(float lr, float beta1, float beta2, int t) {
  // Keras computes this in fp32: lr * sqrt(1 - beta2^t) / (1 - beta1^t)
  const float b1p = powf(beta1, (float)t), b2p = powf(beta2, (float)t);
  return lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
}

int launch_adam(drb_ctx* ctx, const AdamArgs& a, int* n_blocks_out) {
  if (a.nseg < 1 || a.nseg > DRB_MAX_SEGS) return drb_fail(DRB_E_INVALID, "adam: bad segment count");
  const int64_t total4 = a.seg[a.nseg - 1].off4 + a.seg[a.nseg - 1].n4;      // the launch covers [seg[0].off4, total4)
  int64_t want = (total4 - a.seg[0].off4 + kAdamThreads - 1) / kAdamThreads;
  // 16 float4 per thread in flight across the grid keeps HBM busy; grid is a multiple of the SM count
  int blocks = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * 16);
  if (blocks < 1) blocks = 1;
  *n_blocks_out = blocks;
  drb_prof_scope prof_(ctx, "k_adam");
  k_adam<<<blocks, kAdamThreads, 0, ctx->stream>>>(a, total4);
  DRB_LAUNCH_CHECK(ctx, "k_adam");
  return DRB_OK;
}

int launch_finalize_loss(drb_ctx* ctx, const float* loss_part, int n_loss, float scale, const float* reg_part,
                         int n_reg, float* loss_out) {
  drb_prof_scope prof_(ctx, "k_finalize_loss");
  k_finalize_loss<<<1, 256, 0, ctx->stream>>>(loss_part, n_loss, scale, reg_part, n_reg, loss_out);
  DRB_LAUNCH_CHECK(ctx, "k_finalize_loss");
  return DRB_OK;
}
