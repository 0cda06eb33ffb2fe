// Shared device / host helpers of the tcgen05 kernels (umma.cu, umma_loss.cu): mbarrier, TMA, tcgen05 wrappers,
// UMMA shared-memory descriptors and TMA tensor-map encoding.  Internal to libdrb.
#ifndef DRB_UMMA_COMMON_CUH
#define DRB_UMMA_COMMON_CUH

#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace {

constexpr int BM = 128, UMMA_K = 8;   // the k-block depth (32 or 16 floats per stage) is a kernel template parameter
constexpr float KERAS_EPS = 1e-7f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// The same wait for the many epilogue warps of a persistent kernel: between polls the warp sleeps, so that sixteen
// spinning warps do not take the issue slots of the one thread that issues the MMAs and the one that issues the TMA
// loads (ncu, score filter: 21 M of 52 M executed instructions were this spin; the MMA thread was starved)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  while (!done) {
    asm volatile("nanosleep.u32 %0;" ::"r"(ns));
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// ---- CTA-pair forms (cta_group::2, cluster of two CTAs on one TPC): one MMA spans both CTAs' shared and tensor
// memory (M = 256: 128 rows of A and of D per CTA, each CTA holds half of the N rows of B), issued by the even CTA.
// TMA loads of either CTA signal the even CTA's mbarrier; commits arrive on the same mbarrier offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {   // same offset in CTA `rank` of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp16 operands (kind::f16): the same shared-memory descriptors (byte based), one MMA covers K = 16 halfs = 32 bytes
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one split-precision MMA of either kind (H: fp16 hi/lo operands, else tf32 hi/lo), one CTA or a CTA pair
template <bool H, int CL>
__device__ __forceinline__ void umma_split(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  if (H) {
    if (CL > 1) umma_f16_pair(tmem_d, adesc, bdesc, idesc, accumulate);
    else umma_f16(tmem_d, adesc, bdesc, idesc, accumulate);
  } else {
    if (CL > 1) umma_tf32_pair(tmem_d, adesc, bdesc, idesc, accumulate);
    else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
  }
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1; A / B format [7,10) / [10,13): TF32 = 2
// (kind::tf32), F16 = 0 (kind::f16); a_major [15] (1 = MN-major); N >> 3 [17,23); M >> 4 [24,29)
template <bool H>
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool a_mn_major = false) {
  const uint32_t fmt = H ? 0u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// fp16 split of a value already scaled into the fp16 range: x = hi + lo up to ~2^-22 |x| (both round-to-nearest; lo is
// a subnormal half only for elements 2^-11 below the tensor's maximum, where its absolute error is irrelevant)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// asynchronous form: issue the load now, make the registers visible later with tmem_ld16_wait (which names the
// registers as in/out operands so the compiler cannot move or spill them across the wait)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type [61,64): 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B -- the only
// layout the hardware accepts for MN-major 32-bit (tf32) operands: 128-byte rows swizzled in 32-byte units over
// 4-row atoms (cute::UMMA::Layout_MN_SW128_32B_Atom, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint64_t layout_type = 2) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46) | (layout_type << 61);
}

// K-major operand tile [rows][KB floats] as written by TMA: 128-byte rows / SWIZZLE_128B for KB = 32, 64-byte rows /
// SWIZZLE_64B (layout_type 4) for KB = 16; 8-row groups are 8 * row bytes apart, one MMA (K = 8) advances 32 bytes
template <int KB>
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t saddr, int kk) {
  return KB == 32 ? make_desc(saddr + kk * 32, 16, 1024, 2) : make_desc(saddr + kk * 32, 16, 512, 4);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D fp32 tensor [outer][inner] with row pitch `pitch_floats`; box {box_inner, box_outer}; OOB = 0.  Swizzle: 128-byte
// (16-byte units) for 32-float box rows, 64-byte for 16-float box rows, or the 32-byte-unit 128-byte form (atom32)
int make_map(CUtensorMap* map, const float* ptr, int64_t inner, int64_t outer, int64_t pitch_floats, int box_inner,
             int box_outer, bool atom32 = false) {
  const CUtensorMapSwizzle swz = atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                        : (box_inner == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  EncodeTiledFn enc = get_encode();
  if (!enc) return drb_fail(DRB_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_floats * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return drb_fail(DRB_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DRB_OK;
}

// The same for an fp16 tensor [outer][inner halfs]: 128-byte box rows (64 halfs, SWIZZLE_128B) or 64-byte (32 halfs,
// SWIZZLE_64B).  pitch_halfs must be a multiple of 8 (16-byte global strides).
int make_map_h(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t pitch_halfs, int box_inner,
               int box_outer) {
  const CUtensorMapSwizzle swz = (box_inner == 32) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  EncodeTiledFn enc = get_encode();
  if (!enc) return drb_fail(DRB_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if (pitch_halfs % 8) return drb_fail(DRB_E_INVALID, "fp16 tensor map: row pitch %lld is not a multiple of 8", (long long)pitch_halfs);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_halfs * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return drb_fail(DRB_E_CUDA, "cuTensorMapEncodeTiled (fp16) failed with CUresult %d", (int)r);
  return DRB_OK;
}

// operand map of either kind: KB is the stage depth in 4-byte units (32 -> 128-byte rows, 16 -> 64-byte rows), i.e.
// KB floats or 2*KB halfs; inner / pitch are in elements of the operand type
template <bool H>
int make_operand_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t pitch, int KB,
                     int box_outer) {
  if (H) return make_map_h(map, ptr, inner, outer, pitch, 2 * KB, box_outer);
  return make_map(map, static_cast<const float*>(ptr), inner, outer, pitch, KB, box_outer);
}

}  // namespace

#endif
