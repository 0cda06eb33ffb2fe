// K3 on the 5th-generation tensor cores (sm_100a): the two backward products of the CDAE output layer,
// dW'^T = dz^T h and dh = dz W'^T, as fp32-accurate 3 x TF32 split GEMMs (A_lo*B_hi + A_hi*B_lo + A_hi*B_hi) issued
// with tcgen05.mma.kind::tf32, operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle),
// accumulator in TMEM, epilogue through tcgen05.ld.  (The forward product with the fused loss epilogue is the
// persistent kernel in umma_loss.cu.)
//
// Replaces the tape.gradient products of DRecPy/Recommender/recommender_abc.py:203 for cdae.py:76.  Why 3xTF32:
// north_star asks for fp32 parity (forward scores within 1e-5 relative); a single TF32/BF16 MMA rounds operands to
// 2^-11 / 2^-8.  Splitting every fp32 operand into hi = rna_tf32(x), lo = x - hi keeps ~22 mantissa bits.
//
// Warp roles (192 threads, one output tile per CTA): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
// (one elected lane), warps 2..5 = epilogue (each owns the 32 TMEM lanes of its sub-partition).
// Tile 128 x BN x 32 floats.  B is K-major ([rows][32 floats], SWIZZLE_128B).  A is dz: K-major for dh, or -- for the
// transposed use in dW'^T -- MN-major, which tcgen05 accepts for 32-bit operands only in the SWIZZLE_128B_BASE32B
// layout (four 32x32 TMA boxes per tile).  dz normally arrives in the 128x32 tile-major layout written by the loss
// kernel, so every box is one contiguous 16 KB / 4 KB read.
#include "umma_common.cuh"

namespace {

constexpr int THREADS = 192;

// KB = floats of the reduction dimension per pipeline stage.  Wide tiles (BN > 128) use 16: with 32 only two ~90 KB
// stages fit, one in flight while the other is consumed, and the tensor pipe idled half the time waiting for TMA
// (ncu: sm__pipe_tensor_cycles_active 53 %); 16-deep stages give five stages of the same shared memory.
//
// CL = 2 runs two adjacent row tiles (blockIdx.x even / odd) as one CTA pair with cta_group::2 MMAs (M = 256): every CTA stages its
// own 128 rows of A but only half of the B rows, which cuts the shared-memory traffic of the MMAs (A + B read per
// instruction, 96 B/clk at full rate with N = 256, next to the TMA writes) -- the actual limit of the one-CTA form.
template <int BN, int KB, int CL>
struct Smem {
  static constexpr int A_BYTES = BM * KB * 4;           // one of A_hi / A_lo
  static constexpr int B_BYTES = (BN / CL) * KB * 4;    // this CTA's rows of B
  static constexpr int STAGES = (CL == 2) ? ((KB == 32) ? 3 : (BN <= 208 ? 7 : 6))
                              : (KB == 32) ? ((BN <= 128) ? 3 : 2) : ((BN <= 128) ? 6 : (BN <= 208) ? 5 : 4);
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

struct UmmaParams {
  int M, N, Kred;           // logical problem; N <= BN * grid.y
  int splits;               // reduction split over blockIdx.z, partial z stored at C + z * M * ldc
  float* C; int ldc;        // EPI_STORE: C[m][n] for n < n_store
  int n_store; int n_valid;  // columns in [n_valid, n_store) are written as 0 (row padding of C)
  float* extra_col; int extra_col_index;   // EPI_STORE: column `extra_col_index` of the product goes to extra_col[m]
  int atomic_out;           // splits > 1: accumulate all partials into C / extra_col with vector atomics (C pre-zeroed)
  int a_tiled_nib;          // > 0: A lives in a tile-major dz layout (128-row tiles of 32 floats / 64 halfs) with this many
                            // column blocks per row tile
  float out_scale; const float* out_scale_dev;   // accumulators are multiplied by out_scale * (*out_scale_dev)
  int stages;               // pipeline stages in use (<= Smem::STAGES): 3 lets two CTAs share an SM (DRB_UMMA_STAGES)
};

template <int BN, bool A_MN, int KB, int CL, bool H>
__global__ void __launch_bounds__(THREADS, 1)
k_umma_gemm(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, UmmaParams p) {
  using S = Smem<BN, KB, CL>;
  constexpr int BK = H ? 2 * KB : KB;            // elements of the reduction dimension per stage (KB: 4-byte units)
  constexpr int UK = H ? 16 : UMMA_K;            // elements one MMA consumes (32 bytes either way)
  constexpr int TW = H ? 64 : 32;                // width of a dz tile in elements (128 bytes)
  const uint32_t cta_rank = (CL > 1) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024-byte alignment
  const int NS = p.stages;
  const uint32_t bars = base + NS * S::STAGE_BYTES;              // full[NS], empty[NS], tmem_full, tmem_ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NS + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * NS);
  const uint32_t tmem_ptr_addr = bars + 8u * (2 * NS + 1);
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;   // row tiles along x: a CTA pair (cluster {2,1,1}) is two adjacent row tiles
  const int nkb_total = (p.Kred + BK - 1) / BK;
  const int kb_per = (nkb_total + p.splits - 1) / p.splits;
  const int kb_beg = blockIdx.z * kb_per;
  const int kb_end = min(nkb_total, kb_beg + kb_per);
  const int nkb = max(0, kb_end - kb_beg);
  constexpr uint32_t TMEM_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CL > 1) {   // collective: the same warp of both CTAs
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_generic;
  // every TMA load of the pair reports to the even CTA's full barrier (only that CTA issues MMAs)
  auto load = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if (CL > 1) tma_load_2d_pair(dst, map, mapa_rank(bar, 0), c0, c1);
    else tma_load_2d(dst, map, bar, c0, c1);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < nkb; i++) {
        const int s = i % NS;
        const uint32_t ph = (i / NS) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (leader) mbar_expect_tx(full_bar(s), CL * S::STAGE_BYTES);
        const uint32_t sa_hi = base + s * S::STAGE_BYTES, sa_lo = sa_hi + S::A_BYTES;
        const uint32_t sb_hi = sa_lo + S::A_BYTES, sb_lo = sb_hi + S::B_BYTES;
        const int k0 = (kb_beg + i) * BK;
        if (p.a_tiled_nib > 0) {
          // dz tile-major layout: tile (row tile rt, column block cb) = 128 rows x 32 floats, contiguous 16 KB at
          // row ((rt * nib + cb) * 128) of a [*, 32] tensor
          if (A_MN) {   // M runs over dz columns, K over dz rows: one box {TW cols (128 bytes), BK rows} per MN atom
#pragma unroll
            for (int j = 0; j < BM / TW; j++) {
              const int row = ((k0 >> 7) * p.a_tiled_nib + (m0 / TW) + j) * 128 + (k0 & 127);
              load(sa_hi + j * (BK * 128), &map_a_hi, full_bar(s), 0, row);
              load(sa_lo + j * (BK * 128), &map_a_lo, full_bar(s), 0, row);
            }
          } else {      // M runs over dz rows, K over dz columns: a whole 16 KB tile (128-byte stages) or its left / right half
            const int row = ((m0 >> 7) * p.a_tiled_nib + (k0 / TW)) * 128;
            load(sa_hi, &map_a_hi, full_bar(s), k0 % TW, row);
            load(sa_lo, &map_a_lo, full_bar(s), k0 % TW, row);
          }
        } else if (A_MN) {
          // A[m][k] stored as G[k][m] (m contiguous): boxes of {TW m (128 bytes), BK k}, one per MN atom column
#pragma unroll
          for (int j = 0; j < BM / TW; j++) {
            load(sa_hi + j * (BK * 128), &map_a_hi, full_bar(s), m0 + TW * j, k0);
            load(sa_lo + j * (BK * 128), &map_a_lo, full_bar(s), m0 + TW * j, k0);
          }
        } else {
          load(sa_hi, &map_a_hi, full_bar(s), k0, m0);
          load(sa_lo, &map_a_lo, full_bar(s), k0, m0);
        }
        load(sb_hi, &map_b_hi, full_bar(s), k0, n0 + (int)cta_rank * (BN / CL));
        load(sb_lo, &map_b_lo, full_bar(s), k0, n0 + (int)cta_rank * (BN / CL));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0 && leader) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
      // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = make_idesc<H>(BM * CL, BN, A_MN);
      for (int i = 0; i < nkb; i++) {
        const int s = i % NS;
        const uint32_t ph = (i / NS) & 1;
        mbar_wait(full_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa_hi = base + s * S::STAGE_BYTES, sa_lo = sa_hi + S::A_BYTES;
        const uint32_t sb_hi = sa_lo + S::A_BYTES, sb_lo = sb_hi + S::B_BYTES;
        // last k-block of the reduction: MMAs over the zero-filled tail of the box are not issued
        const int kk_n = min(BK / UK, (p.Kred - (kb_beg + i) * BK + UK - 1) / UK);
#pragma unroll
        for (int kk = 0; kk < BK / UK; kk++) {
          if (kk >= kk_n) break;
          uint64_t a_hi, a_lo;
          if (A_MN && H) {
            // MN-major fp16: rows = k (128 B = 64 m each), SWIZZLE_128B; K atoms of 8 rows are 1024 B apart, one MMA
            // (K = 16) spans two of them; MN atoms (64 halfs, one TMA box of BK rows) are BK * 128 B apart
            a_hi = make_desc(sa_hi + kk * 2048, BK * 128, 1024, 2);
            a_lo = make_desc(sa_lo + kk * 2048, BK * 128, 1024, 2);
          } else if (A_MN) {   // MN-major tf32: rows = k (128 B = 32 m each); K atoms of 4 rows are 512 B apart, one MMA (K=8)
                        // spans two of them; MN atoms (32 floats, one TMA box of BK rows) are BK * 128 B apart
            a_hi = make_desc(sa_hi + kk * 1024, KB * 128, 512, 1);
            a_lo = make_desc(sa_lo + kk * 1024, KB * 128, 512, 1);
          } else {
            a_hi = make_desc_kmajor<KB>(sa_hi, kk);
            a_lo = make_desc_kmajor<KB>(sa_lo, kk);
          }
          const uint64_t b_hi = make_desc_kmajor<KB>(sb_hi, kk);
          const uint64_t b_lo = make_desc_kmajor<KB>(sb_lo, kk);
          umma_split<H, CL>(tmem_base, a_lo, b_hi, idesc, (i | kk) != 0);   // small terms first
          umma_split<H, CL>(tmem_base, a_hi, b_lo, idesc, 1u);
          umma_split<H, CL>(tmem_base, a_hi, b_hi, idesc, 1u);
        }
        if (CL > 1) umma_commit_pair(empty_bar(s));   // frees the smem stage (both CTAs) once these MMAs have read it
        else umma_commit(empty_bar(s));
      }
      if (CL > 1) umma_commit_pair(tmem_full_bar);    // accumulator complete
      else umma_commit(tmem_full_bar);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                       // TMEM sub-partition of this warp
    const int m = m0 + q * 32 + lane;             // one accumulator row per thread
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* crow = p.C + (p.atomic_out ? 0 : (int64_t)blockIdx.z * p.M * p.ldc) + (int64_t)m * p.ldc;
    const float osc = p.out_scale * (p.out_scale_dev ? __ldg(p.out_scale_dev) : 1.0f);
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t r[16];
      if (nkb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      } else {
#pragma unroll
        for (int j = 0; j < 16; j++) r[j] = 0u;
      }
      if (H) {
#pragma unroll
        for (int j = 0; j < 16; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) * osc);
      }
      const int n = n0 + c;
      if (m >= p.M) continue;
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        const int nn = n + j4 * 4;
        if (nn < p.n_store) {   // n_store % 4 == 0
          const float4 v = make_float4(nn + 0 < p.n_valid ? __uint_as_float(r[j4 * 4]) : 0.f,
                                       nn + 1 < p.n_valid ? __uint_as_float(r[j4 * 4 + 1]) : 0.f,
                                       nn + 2 < p.n_valid ? __uint_as_float(r[j4 * 4 + 2]) : 0.f,
                                       nn + 3 < p.n_valid ? __uint_as_float(r[j4 * 4 + 3]) : 0.f);
          if (p.atomic_out) atomicAdd(reinterpret_cast<float4*>(crow + nn), v);
          else *reinterpret_cast<float4*>(crow + nn) = v;
        }
      }
      if (p.extra_col && p.extra_col_index >= n && p.extra_col_index < n + 16 && (p.atomic_out || blockIdx.z == 0)) {
#pragma unroll
        for (int j = 0; j < 16; j++)
          if (n + j == p.extra_col_index) {
            if (p.atomic_out) atomicAdd(p.extra_col + m, __uint_as_float(r[j]));
            else p.extra_col[m] = __uint_as_float(r[j]);
          }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CL > 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------ operand preparation
// src [rows][ld] -> hi/lo [rows][ld] and, optionally, transposed hi/lo [ld_t rows >= cols][ldt] (tile transpose)
__global__ void __launch_bounds__(256) k_split_tf32(const float* __restrict__ src, int rows, int cols, int ld,
                                                   float* __restrict__ hi, float* __restrict__ lo,
                                                   float* __restrict__ t_hi, float* __restrict__ t_lo, int ldt,
                                                   int ones_row) {
  __shared__ float th[32][33], tl[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    float h = 0.f, l = 0.f;
    if (r < rows && c < cols) {
      split_tf32(src[(int64_t)r * ld + c], h, l);
      if (hi) { hi[(int64_t)r * ld + c] = h; lo[(int64_t)r * ld + c] = l; }
    }
    th[ty + i * 8][tx] = h;
    tl[ty + i * 8][tx] = l;
  }
  if (!t_hi) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = c0 + ty + i * 8, r = r0 + tx;   // transposed: row index c, column index r
    if (c < cols && r < rows) {
      t_hi[(int64_t)c * ldt + r] = th[tx][ty + i * 8];
      t_lo[(int64_t)c * ldt + r] = tl[tx][ty + i * 8];
    }
    if (ones_row >= 0 && c == ones_row && r < rows) {   // constant-one feature: folds the bias gradient into the GEMM
      t_hi[(int64_t)c * ldt + r] = 1.0f;
      t_lo[(int64_t)c * ldt + r] = 0.0f;
    }
  }
}


// fp16 hi / lo split of a scaled fp32 matrix (see launch_split_f16 in kernels.h); same 32 x 32 tile transpose
__global__ void __launch_bounds__(256) k_split_f16(const float* __restrict__ src, int rows, int cols, int ld, float alpha,
                                                  const float* __restrict__ alpha_dev, __half* __restrict__ hi,
                                                  __half* __restrict__ lo, int ldh, __half* __restrict__ t_hi,
                                                  __half* __restrict__ t_lo, int ldt, int ones_row) {
  __shared__ __half th[32][34], tl[32][34];
  const float sc = alpha * (alpha_dev ? __ldg(alpha_dev) : 1.0f);
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    __half h = __float2half_rn(0.f), l = h;
    if (r < rows && c < cols) {
      split_f16(src[(int64_t)r * ld + c] * sc, h, l);
      if (hi) { hi[(int64_t)r * ldh + c] = h; lo[(int64_t)r * ldh + c] = l; }
    }
    th[ty + i * 8][tx] = h;
    tl[ty + i * 8][tx] = l;
  }
  if (!t_hi) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = c0 + ty + i * 8, r = r0 + tx;   // transposed: row index c, column index r
    if (c < cols && r < rows) {
      t_hi[(int64_t)c * ldt + r] = th[tx][ty + i * 8];
      t_lo[(int64_t)c * ldt + r] = tl[tx][ty + i * 8];
    }
    if (ones_row >= 0 && c == ones_row && r < rows) {   // constant-one feature, in the scaled units of this operand
      t_hi[(int64_t)c * ldt + r] = __float2half_rn(sc);
      t_lo[(int64_t)c * ldt + r] = __float2half_rn(0.f);
    }
  }
}

// max |x| over the first `cols` columns of x[rows][ld], as the bit pattern of a non-negative float (atomicMax on uint)
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ x, int64_t rows, int cols, int ld,
                                                uint32_t* __restrict__ out_bits) {
  float m = 0.f;
  const int64_t total = rows * (int64_t)ld;
  (void)cols;   // pad columns of the arenas are exactly zero, so the whole pitch can be scanned
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));
}
// scales[0] = alpha = 2^(14 - floor(log2 max)), scales[1] = 1 / alpha; max == 0 (or not finite) -> 1
__global__ void k_publish_scale(const uint32_t* __restrict__ max_bits, float* __restrict__ scales) {
  const float m = __uint_as_float(*max_bits);
  float alpha = 1.0f;
  if (m > 0.f && isfinite(m)) {
    const int e = (int)((*max_bits >> 23) & 0xffu) - 127;     // floor(log2 m) for normal m
    const int ea = max(-100, min(100, 14 - e));
    alpha = exp2f((float)ea);
  }
  scales[0] = alpha;
  scales[1] = 1.0f / alpha;
}

template <int BN, bool A_MN, int KB, int CL, bool H>
int run_umma(drb_ctx* ctx, const UmmaOperands& o, const UmmaParams& p_in, int* n_blocks_out) {
  constexpr int BK = KB;
  using SM = Smem<BN, KB, CL>;
  static const int st_env = getenv("DRB_UMMA_STAGES") ? atoi(getenv("DRB_UMMA_STAGES")) : 0;
  UmmaParams p = p_in;
  p.stages = (st_env >= 2 && st_env < SM::STAGES) ? st_env : SM::STAGES;
  const int SMEM = p.stages * SM::STAGE_BYTES + 1024 + 256;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int r;
  if (H) {      // fp16 hi / lo operands; tile-major dz = a [tiles * 128, 64 halfs] tensor
    if (o.a_tiled_nib > 0 && A_MN) {          // MN-major: boxes of {64 halfs of M, 2 * KB rows of K}
      if ((r = make_map_h(&ma_hi, o.a_hi, 64, o.a_tiled_rows, 64, 64, 2 * KB))) return r;
      if ((r = make_map_h(&ma_lo, o.a_lo, 64, o.a_tiled_rows, 64, 64, 2 * KB))) return r;
    } else if (A_MN) {                        // A given as G[k][m]
      if ((r = make_map_h(&ma_hi, o.a_hi, p.M, p.Kred, o.lda, 64, 2 * KB))) return r;
      if ((r = make_map_h(&ma_lo, o.a_lo, p.M, p.Kred, o.lda, 64, 2 * KB))) return r;
    } else if (o.a_tiled_nib > 0) {
      if ((r = make_map_h(&ma_hi, o.a_hi, 64, o.a_tiled_rows, 64, 2 * KB, BM))) return r;
      if ((r = make_map_h(&ma_lo, o.a_lo, 64, o.a_tiled_rows, 64, 2 * KB, BM))) return r;
    } else {
      if ((r = make_map_h(&ma_hi, o.a_hi, p.Kred, p.M, o.lda, 2 * KB, BM))) return r;
      if ((r = make_map_h(&ma_lo, o.a_lo, p.Kred, p.M, o.lda, 2 * KB, BM))) return r;
    }
    if ((r = make_map_h(&mb_hi, o.b_hi, p.Kred, o.b_rows, o.ldb, 2 * KB, BN / CL))) return r;
    if ((r = make_map_h(&mb_lo, o.b_lo, p.Kred, o.b_rows, o.ldb, 2 * KB, BN / CL))) return r;
  } else {
  const float* fa_hi = static_cast<const float*>(o.a_hi); const float* fa_lo = static_cast<const float*>(o.a_lo);
  const float* fb_hi = static_cast<const float*>(o.b_hi); const float* fb_lo = static_cast<const float*>(o.b_lo);
  if (o.a_tiled_nib > 0) {   // tile-major dz: a [tiles * 128, 32] tensor
    if ((r = make_map(&ma_hi, fa_hi, 32, o.a_tiled_rows, 32, A_MN ? 32 : BK, A_MN ? BK : BM, A_MN))) return r;
    if ((r = make_map(&ma_lo, fa_lo, 32, o.a_tiled_rows, 32, A_MN ? 32 : BK, A_MN ? BK : BM, A_MN))) return r;
  } else if (A_MN) {   // A given as G[k][m]
    if ((r = make_map(&ma_hi, fa_hi, p.M, p.Kred, o.lda, 32, BK, true))) return r;
    if ((r = make_map(&ma_lo, fa_lo, p.M, p.Kred, o.lda, 32, BK, true))) return r;
  } else {      // A given as G[m][k]
    if ((r = make_map(&ma_hi, fa_hi, p.Kred, p.M, o.lda, BK, BM))) return r;
    if ((r = make_map(&ma_lo, fa_lo, p.Kred, p.M, o.lda, BK, BM))) return r;
  }
  if ((r = make_map(&mb_hi, fb_hi, p.Kred, o.b_rows, o.ldb, BK, BN / CL))) return r;
  if ((r = make_map(&mb_lo, fb_lo, p.Kred, o.b_rows, o.ldb, BK, BN / CL))) return r;
  }
  dim3 grid(((p.M + BM - 1) / BM + CL - 1) / CL * CL, (p.N + BN - 1) / BN, p.splits);   // whole pairs of row tiles
  if (n_blocks_out) *n_blocks_out = grid.x * grid.y;
  auto kern = k_umma_gemm<BN, A_MN, KB, CL, H>;
  static bool attr_set = false;   // per template instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return drb_fail(DRB_E_CUDA, "cudaFuncSetAttribute(smem=%d) failed: %s", SM::TOTAL, cudaGetErrorString(e));
    attr_set = true;
  }
  drb_prof_scope prof_(ctx, o.name ? o.name : (A_MN ? "k_umma_gemm_mn" : "k_umma_gemm_kk"));
  if (CL > 1) {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = ctx->stream;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mb_hi, mb_lo, p);
    if (e != cudaSuccess) return drb_fail(DRB_E_CUDA, "cluster launch of k_umma_gemm failed: %s", cudaGetErrorString(e));
  } else {
    kern<<<grid, THREADS, SMEM, ctx->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  }
  DRB_LAUNCH_CHECK(ctx, "k_umma_gemm");
  return DRB_OK;
}

}  // namespace

int launch_split_tf32(drb_ctx* ctx, const float* src, int rows, int cols, int ld, float* hi, float* lo, float* t_hi,
                      float* t_lo, int ldt, int ones_row) {
  int ccols = cols;
  if (ones_row >= cols) ccols = ones_row + 1;   // make sure a block visits the ones row
  dim3 grid((ccols + 31) / 32, (rows + 31) / 32);
  drb_prof_scope prof_(ctx, "k_split_tf32");
  k_split_tf32<<<grid, 256, 0, ctx->stream>>>(src, rows, cols, ld, hi, lo, t_hi, t_lo, ldt, ones_row);
  DRB_LAUNCH_CHECK(ctx, "k_split_tf32");
  return DRB_OK;
}

int launch_split_f16(drb_ctx* ctx, const float* src, int rows, int cols, int ld, float alpha, const float* alpha_dev,
                     void* hi, void* lo, int ldh, void* t_hi, void* t_lo, int ldt, int ones_row) {
  if (rows <= 0 || cols <= 0) return DRB_OK;
  int ccols = cols;
  if (ones_row >= cols) ccols = ones_row + 1;   // make sure a block visits the ones row
  dim3 grid((ccols + 31) / 32, (rows + 31) / 32);
  drb_prof_scope prof_(ctx, "k_split_f16");
  k_split_f16<<<grid, 256, 0, ctx->stream>>>(src, rows, cols, ld, alpha, alpha_dev, static_cast<__half*>(hi),
                                             static_cast<__half*>(lo), ldh, static_cast<__half*>(t_hi),
                                             static_cast<__half*>(t_lo), ldt, ones_row);
  DRB_LAUNCH_CHECK(ctx, "k_split_f16");
  return DRB_OK;
}

int launch_absmax_scale(drb_ctx* ctx, const float* x, int64_t rows, int cols, int ld, float* scales) {
  uint32_t* bits = reinterpret_cast<uint32_t*>(scales + 2);
  cudaError_t e = cudaMemsetAsync(bits, 0, 4, ctx->stream);
  if (e != cudaSuccess) return drb_fail(DRB_E_CUDA, "absmax: memset failed: %s", cudaGetErrorString(e));
  const int64_t total = rows * (int64_t)ld;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 8));
  {
    drb_prof_scope prof_(ctx, "k_absmax");
    k_absmax<<<blocks, 256, 0, ctx->stream>>>(x, rows, cols, ld, bits);
    DRB_LAUNCH_CHECK(ctx, "k_absmax");
  }
  k_publish_scale<<<1, 1, 0, ctx->stream>>>(bits, scales);
  DRB_LAUNCH_CHECK(ctx, "k_publish_scale");
  return DRB_OK;
}

int launch_umma_store(drb_ctx* ctx, const UmmaOperands& o, bool a_mn_major, int M, int N, int Kred, int splits,
                      float* C, int ldc, int n_store, int n_valid, float* extra_col, int extra_col_index,
                      bool atomic_out) {
  UmmaParams p{};
  p.M = M; p.N = N; p.Kred = Kred; p.splits = splits; p.C = C; p.ldc = ldc; p.n_store = n_store; p.n_valid = n_valid;
  p.extra_col = extra_col; p.extra_col_index = extra_col_index;
  p.a_tiled_nib = o.a_tiled_nib;
  p.atomic_out = atomic_out ? 1 : 0;
  p.out_scale = o.out_scale; p.out_scale_dev = o.out_scale_dev;
  if (N > 256) return drb_fail(DRB_E_INVALID, "umma store GEMM: N must be <= 256 (hidden width)");
  static const int bk_env = getenv("DRB_UMMA_BK") ? atoi(getenv("DRB_UMMA_BK")) : 0;   // profiling override: 16 | 32
  static const int cl_env = getenv("DRB_UMMA_CLUSTER") ? atoi(getenv("DRB_UMMA_CLUSTER")) : 0;   // override: 1 | 2
#define DRB_UMMA_CASE(BN, KB_DEFAULT, CL_DEFAULT)                                       \
  if (N <= BN) {                                                                        \
    if (o.half) {   /* fp16 hi / lo operands */                                         \
      if ((bk_env ? bk_env : KB_DEFAULT) == 16) {                                       \
        if ((cl_env ? cl_env : CL_DEFAULT) == 2) {                                      \
          if (a_mn_major) return run_umma<BN, true, 16, 2, true>(ctx, o, p, nullptr);   \
          return run_umma<BN, false, 16, 2, true>(ctx, o, p, nullptr);                  \
        }                                                                               \
        if (a_mn_major) return run_umma<BN, true, 16, 1, true>(ctx, o, p, nullptr);     \
        return run_umma<BN, false, 16, 1, true>(ctx, o, p, nullptr);                    \
      }                                                                                 \
      if (a_mn_major) return run_umma<BN, true, 32, 1, true>(ctx, o, p, nullptr);       \
      return run_umma<BN, false, 32, 1, true>(ctx, o, p, nullptr);                      \
    }                                                                                   \
    if ((bk_env ? bk_env : KB_DEFAULT) == 16) {                                         \
      if ((cl_env ? cl_env : CL_DEFAULT) == 2) {                                        \
        if (a_mn_major) return run_umma<BN, true, 16, 2, false>(ctx, o, p, nullptr);    \
        return run_umma<BN, false, 16, 2, false>(ctx, o, p, nullptr);                   \
      }                                                                                 \
      if (a_mn_major) return run_umma<BN, true, 16, 1, false>(ctx, o, p, nullptr);      \
      return run_umma<BN, false, 16, 1, false>(ctx, o, p, nullptr);                     \
    }                                                                                   \
    if (a_mn_major) return run_umma<BN, true, 32, 1, false>(ctx, o, p, nullptr);        \
    return run_umma<BN, false, 32, 1, false>(ctx, o, p, nullptr);                       \
  }
  DRB_UMMA_CASE(64, 32, 1)
  DRB_UMMA_CASE(128, 32, 1)
  DRB_UMMA_CASE(208, 16, 2)
  DRB_UMMA_CASE(256, 16, 2)
#undef DRB_UMMA_CASE
  return drb_fail(DRB_E_INVALID, "umma store GEMM: unsupported N");
}

bool umma_available() { return get_encode() != nullptr; }
