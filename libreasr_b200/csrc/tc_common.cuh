// sm_100a tensor-core building blocks: mbarrier, TMA bulk copy, tcgen05 (alloc / mma /
// commit / ld), UMMA descriptors, and the "operand image" layout shared by every
// tensor-core kernel of the path.
//
// Operand image.  Every GEMM operand (activations = A, weights = B; both K-major) is kept
// in HBM as a sequence of ready-to-use shared-memory tiles:
//     image[row_tile][k_block][part] = TR rows x 64 halves, 128-byte swizzled (SW128)
// part 0 = hi, part 1 = lo of the 3xFP16 split (x ~= hi + lo * 2^-11, both fp16).  A tile
// is a contiguous TR*128 bytes, so the producer warp moves it with ONE cp.async.bulk (TMA
// bulk engine, no tensor map) and the swizzle is already baked in; the UMMA descriptor
// then only needs the tile's shared-memory address.  Byte offset of element (r, k) inside
// a tile:  r*128 + (((k>>3) ^ (r&7)) << 4) + (k&7)*2      (Swizzle<3,4,3>, atoms of 8 rows).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace rnnt {

constexpr int kImgK = 64;              // halves per k-block (128 bytes per row)
constexpr float kLoScale = 2048.0f;    // lo part is stored multiplied by 2^11
constexpr float kLoInv = 1.0f / 2048.0f;

__host__ __device__ inline size_t img_tile_bytes(int TR) { return (size_t)TR * 128; }
// bytes of a whole image: ceil(R/TR) row tiles x KB k-blocks x 2 parts
__host__ __device__ inline size_t img_bytes(int64_t R, int K, int TR) {
  return (size_t)ceil_div(R, TR) * (size_t)ceil_div(K, kImgK) * 2 * img_tile_bytes(TR);
}
__host__ __device__ inline size_t img_tile_offset(int64_t row_tile, int kb, int part, int KB, int TR) {
  return ((size_t)(row_tile * KB + kb) * 2 + part) * img_tile_bytes(TR);
}
__host__ __device__ inline uint32_t img_elem_offset(int r, int k) {
  return (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}
__device__ __forceinline__ void split_f16x3(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * kLoScale);
}

// ---------------- mbarrier ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// one lane of a fully converged warp (lets the issuing code stay warp-uniform, so descriptors live
// in uniform registers: a diverged `if (lane == 0)` loop costs ~150 cycles per tcgen05.mma in R2UR traffic)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---------------- TMA bulk copy (global -> shared, completes on an mbarrier) ----------------
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------- tcgen05 ----------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// all previously issued MMAs of this thread arrive on `bar` when they have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate; one thread issues
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SW128 shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major: 1) | [32,46) SBO>>4 = 1024>>4
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp16 x fp16 -> fp32, both K-major
__host__ __device__ inline uint32_t umma_idesc_f16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                     // c_format = F32
  d |= 0u << 7;                     // a_format = F16
  d |= 0u << 10;                    // b_format = F16
  d |= (uint32_t)(N >> 3) << 17;    // n_dim
  d |= (uint32_t)(M >> 4) << 24;    // m_dim
  return d;
}

}  // namespace rnnt
