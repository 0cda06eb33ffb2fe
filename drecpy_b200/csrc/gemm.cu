// K2 / K3 dense GEMMs (sm_100a, fp32 FFMA path): output layer h * W'^T with the fused sigmoid / BCE / dL/dz
// epilogue, its two backward products, the DMF tower layers, and full-catalog scoring.
//
// Replaces (reference, DRecPy/): Recommender/cdae.py:76 (tf.matmul(hidden, W_) + b_ + sigmoid), cdae.py:78-79
// (Keras BinaryCrossentropy / MeanSquaredError on the (B,B,I) broadcast == batch-mean labels),
// the matching tape.gradient products (recommender_abc.py:203) and Recommender/dmf.py:89-90 (Dense stacks).
//
// This is the exact-fp32 CUDA-core path: north_star demands forward scores within 1e-5 relative, which plain
// TF32/BF16 tensor-core MMA cannot give; the tcgen05 3xTF32 split path is the follow-up (DESIGN.md "GEMM").
// 128 x BN x 16 tiles, 256 threads, 8 x TN register micro-tiles, double-buffered shared memory with register
// prefetch, 128-bit global and shared accesses.  Operand layouts are compile-time so every global access is a
// coalesced float4 regardless of which operand is transposed.
#include "kernels.h"

namespace {

constexpr int BM = 128, BK = 16, THREADS = 256, PAD = 4;
constexpr float KERAS_EPS = 1e-7f;

template <int BN, int LAYOUT, int EPI>
__global__ void __launch_bounds__(THREADS) k_sgemm(GemmArgs g) {
  constexpr int TN = BN / 16;  // 8 (BN=128) or 4 (BN=64)
  constexpr bool A_KM = (LAYOUT == LAYOUT_KK || LAYOUT == LAYOUT_KN);
  constexpr bool B_KM = (LAYOUT == LAYOUT_KK);
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // split-K range
  const int kchunk = (int)(((int64_t)(g.Kred + g.splits - 1) / g.splits + BK - 1) / BK * BK);
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(g.Kred, kbeg + kchunk);
  const int ntiles = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  constexpr int A_LD4 = BM * BK / 4 / THREADS;   // 2 float4 per thread
  constexpr int B_LD4 = (BN * BK / 4 + THREADS - 1) / THREADS;  // 2 (BN=128) or 1 (BN=64)
  float4 ra[A_LD4], rb[B_LD4];

  auto load_a = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_LD4; i++) {
      const int f = tid + i * THREADS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A_KM) {  // A[m][k]: float4 along k
        const int row = f / (BK / 4), kq = f % (BK / 4);
        const int m = m0 + row, k = k0 + kq * 4;
        if (m < g.M && k < kend) {
          v = __ldg(reinterpret_cast<const float4*>(g.A + (int64_t)m * g.lda + k));
          if (k + 3 >= kend) {  // ragged end of the reduction range (lda padding is readable)
            if (k + 1 >= kend) v.y = 0.f;
            if (k + 2 >= kend) v.z = 0.f;
            v.w = 0.f;
          }
        }
      } else {     // A[k][m]: float4 along m
        const int kr = f / (BM / 4), mq = f % (BM / 4);
        const int k = k0 + kr, m = m0 + mq * 4;
        if (k < kend && m < g.M) v = __ldg(reinterpret_cast<const float4*>(g.A + (int64_t)k * g.lda + m));
      }
      ra[i] = v;
    }
  };
  auto load_b = [&](int k0) {
#pragma unroll
    for (int i = 0; i < B_LD4; i++) {
      const int f = tid + i * THREADS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < BN * BK / 4) {
        if (B_KM) {
          const int row = f / (BK / 4), kq = f % (BK / 4);
          const int n = n0 + row, k = k0 + kq * 4;
          if (n < g.N && k < kend) {
            v = __ldg(reinterpret_cast<const float4*>(g.B + (int64_t)n * g.ldb + k));
            if (k + 3 >= kend) {
              if (k + 1 >= kend) v.y = 0.f;
              if (k + 2 >= kend) v.z = 0.f;
              v.w = 0.f;
            }
          }
        } else {
          const int kr = f / (BN / 4), nq = f % (BN / 4);
          const int k = k0 + kr, n = n0 + nq * 4;
          if (k < kend && n < g.N) v = __ldg(reinterpret_cast<const float4*>(g.B + (int64_t)k * g.ldb + n));
        }
      }
      rb[i] = v;
    }
  };
  auto store_a = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD4; i++) {
      const int f = tid + i * THREADS;
      if (A_KM) {
        const int row = f / (BK / 4), kq = f % (BK / 4);
        As[buf][kq * 4 + 0][row] = ra[i].x;
        As[buf][kq * 4 + 1][row] = ra[i].y;
        As[buf][kq * 4 + 2][row] = ra[i].z;
        As[buf][kq * 4 + 3][row] = ra[i].w;
      } else {
        const int kr = f / (BM / 4), mq = f % (BM / 4);
        *reinterpret_cast<float4*>(&As[buf][kr][mq * 4]) = ra[i];
      }
    }
  };
  auto store_b = [&](int buf) {
#pragma unroll
    for (int i = 0; i < B_LD4; i++) {
      const int f = tid + i * THREADS;
      if (f < BN * BK / 4) {
        if (B_KM) {
          const int row = f / (BK / 4), kq = f % (BK / 4);
          Bs[buf][kq * 4 + 0][row] = rb[i].x;
          Bs[buf][kq * 4 + 1][row] = rb[i].y;
          Bs[buf][kq * 4 + 2][row] = rb[i].z;
          Bs[buf][kq * 4 + 3][row] = rb[i].w;
        } else {
          const int kr = f / (BN / 4), nq = f % (BN / 4);
          *reinterpret_cast<float4*>(&Bs[buf][kr][nq * 4]) = rb[i];
        }
      }
    }
  };

  if (ntiles > 0) {
    load_a(kbeg);
    load_b(kbeg);
    store_a(0);
    store_b(0);
  }
  __syncthreads();
  for (int t = 0; t < ntiles; t++) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_a(kbeg + (t + 1) * BK);
      load_b(kbeg + (t + 1) * BK);
    }
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      if (TN == 8) {
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][(BN / 2) + tx * 4]);
        b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      store_a(buf ^ 1);
      store_b(buf ^ 1);
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  float* C = g.C + (int64_t)blockIdx.z * g.M * g.ldc;
  float loss_local = 0.f;
  float colsum[TN];
#pragma unroll
  for (int j = 0; j < TN; j++) colsum[j] = 0.f;

#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int jh = 0; jh < TN / 4; jh++) {
      const int n = n0 + jh * (BN / 2) + tx * 4;
      if (n >= g.ldc) continue;  // ldc % 4 == 0, so a float4 is either fully inside the row or outside
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float accv = acc[i][jh * 4 + j];
        const int nn = n + j;
        float r = accv;
        if (EPI == EPI_BIAS_ACT) {
          r = (nn < g.N) ? accv + (g.bias ? g.bias[nn] : 0.f) : 0.f;
          if (nn < g.N) {
            if (g.act == DRB_ACT_SIGMOID) r = 1.0f / (1.0f + expf(-r));
            else if (g.act == DRB_ACT_RELU) r = fmaxf(r, 0.f);
          }
        } else if (EPI == EPI_MASK_POS) {
          r = (nn < g.N && g.mask[(int64_t)m * g.ldc + nn] > 0.f) ? accv : 0.f;
        } else if (EPI == EPI_CDAE_LOSS) {
          r = 0.f;
          if (nn < g.N) {
            const float z = accv + g.bias[nn];
            const float p = 1.0f / (1.0f + expf(-z));
            float tgt;
            if (g.label_count) tgt = g.label_count[nn] / (float)g.batch;
            else tgt = (float)((g.label_bits[(int64_t)m * g.words_per_row + (nn >> 5)] >> (nn & 31)) & 1u);
            float dp;
            if (g.loss_kind == DRB_LOSS_BCE) {
              const float one_m = 1.0f - KERAS_EPS;
              const float pc = fminf(fmaxf(p, KERAS_EPS), one_m);
              const float da = pc + KERAS_EPS, db = 1.0f - pc + KERAS_EPS;
              loss_local -= tgt * logf(da) + (1.0f - tgt) * logf(db);
              const bool inside = (p >= KERAS_EPS) && (p <= one_m);
              dp = inside ? -(tgt / da - (1.0f - tgt) / db) * g.inv_count : 0.f;
            } else {
              if (g.label_count) loss_local += p * p - 2.0f * p * tgt + tgt;
              else loss_local += (p - tgt) * (p - tgt);
              dp = 2.0f * (p - tgt) * g.inv_count;
            }
            r = dp * p * (1.0f - p);
            colsum[jh * 4 + j] += r;
          }
        }
        o[j] = r;
      }
      *reinterpret_cast<float4*>(C + (int64_t)m * g.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }

  if (EPI == EPI_CDAE_LOSS) {
    // deterministic per-block reductions: loss partial and column partial sums (db' of cdae.py's b_)
    __syncthreads();
    float* sred = &As[0][0][0];  // reuse: need 16 * BN floats  (<= 2*16*132)
#pragma unroll
    for (int jh = 0; jh < TN / 4; jh++)
#pragma unroll
      for (int j = 0; j < 4; j++) sred[ty * BN + jh * (BN / 2) + tx * 4 + j] = colsum[jh * 4 + j];
    __syncthreads();
    if (tid < BN) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 16; r++) s += sred[r * BN + tid];
      const int n = n0 + tid;
      if (n < g.ldc) g.col_part[(int64_t)blockIdx.y * g.ldc + n] = s;
    }
    __syncthreads();
    float* lred = &Bs[0][0][0];
    lred[tid] = loss_local;
    __syncthreads();
    for (int s = THREADS / 2; s > 0; s >>= 1) {
      if (tid < s) lred[tid] += lred[tid + s];
      __syncthreads();
    }
    if (tid == 0) g.loss_part[blockIdx.y * gridDim.x + blockIdx.x] = lred[0];
  }
}

template <int BN, int LAYOUT, int EPI>
int run(drb_ctx* ctx, const GemmArgs& g, int* n_mtiles_out, int* n_blocks_out) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.splits);
  if (grid.y > 65535 || grid.z > 65535) return drb_fail(DRB_E_INVALID, "gemm: grid too large");
  if (n_mtiles_out) *n_mtiles_out = grid.y;
  if (n_blocks_out) *n_blocks_out = grid.x * grid.y;
  drb_prof_scope prof_(ctx, LAYOUT == LAYOUT_KK ? (EPI == EPI_CDAE_LOSS ? "k_sgemm_kk_loss" : "k_sgemm_kk")
                             : (LAYOUT == LAYOUT_MN ? "k_sgemm_mn" : "k_sgemm_kn"));
  k_sgemm<BN, LAYOUT, EPI><<<grid, THREADS, 0, ctx->stream>>>(g);
  DRB_LAUNCH_CHECK(ctx, "k_sgemm");
  return DRB_OK;
}

template <int LAYOUT, int EPI>
int pick_bn(drb_ctx* ctx, const GemmArgs& g, int* a, int* b) {
  // narrow outputs (hidden width 50..256) waste less with 64-wide tiles
  const int waste128 = (int)drb_round_up(g.N, 128) - g.N, waste64 = (int)drb_round_up(g.N, 64) - g.N;
  if (g.N <= 64 || waste64 + 32 < waste128) return run<64, LAYOUT, EPI>(ctx, g, a, b);
  return run<128, LAYOUT, EPI>(ctx, g, a, b);
}

}  // namespace

int launch_gemm(drb_ctx* ctx, int layout, int epi, const GemmArgs& g, int* n_mtiles_out, int* n_blocks_out) {
  if (g.M <= 0 || g.N <= 0) return DRB_OK;
  if ((g.lda | g.ldb | g.ldc) & 3) return drb_fail(DRB_E_INVALID, "gemm: leading dimensions must be multiples of 4");
  if (g.splits < 1) return drb_fail(DRB_E_INVALID, "gemm: splits < 1");
#define DRB_GEMM_CASE(L, E) if (layout == L && epi == E) return pick_bn<L, E>(ctx, g, n_mtiles_out, n_blocks_out)
  DRB_GEMM_CASE(LAYOUT_KK, EPI_CDAE_LOSS);
  DRB_GEMM_CASE(LAYOUT_KK, EPI_BIAS_ACT);
  DRB_GEMM_CASE(LAYOUT_KK, EPI_MASK_POS);
  DRB_GEMM_CASE(LAYOUT_KK, EPI_STORE);
  DRB_GEMM_CASE(LAYOUT_MN, EPI_STORE);
  DRB_GEMM_CASE(LAYOUT_KN, EPI_STORE);
  DRB_GEMM_CASE(LAYOUT_KN, EPI_BIAS_ACT);
#undef DRB_GEMM_CASE
  return drb_fail(DRB_E_INVALID, "gemm: unsupported layout/epilogue combination (%d, %d)", layout, epi);
}
