// Two-CTA (cta_group::2) variant of the 3xFP16 tcgen05 GEMM of gemm_tc.cu:
//     C[M,N] = A[M,K] * W[N,K]^T + bias[N]
// The single-CTA kernel moves 96 KB of operands per 64-deep k-block for 12 M128 N256 K16 MMAs (1536 tensor
// cycles); with every SM pulling at the chip-level L2->SM rate (~43 B/clk/SM) that is 2300 cycles of ingest --
// ncu shows 781 MB of L2->SM traffic at 6.3 TB/s and the tensor pipe 48 % active.  Here a CTA PAIR (one cluster
// of two SMs) owns a 256 x 256 output tile: each CTA stages its own 128 rows of A and only HALF of the W tile
// (128 of the 256 rows), 64 KB per k-block, and the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) that
// read A and W halves from both CTAs' shared memory; each CTA keeps its 128 rows of the accumulators
// (D0 | D1, 512 TMEM columns) and runs its own epilogue.  One third fewer bytes per FLOP, one more pipeline stage.
//
//   warp 0      producer (both CTAs): four 2-D tensor-map TMA loads per stage ([A hi][A lo][W-half hi][W-half lo], 16 KB
//               boxes of 128 image rows x 128 B); `.cta_group::2` lets every load complete on the LEADER's mbarrier, so
//               one barrier per stage tells the leader that both CTAs' operands have landed
//   warp 1      leader CTA: MMA issuer (idle in the peer)
//   warps 2-5   epilogue (both CTAs)
// Stage hand-back and "accumulators ready" use tcgen05.commit ... multicast::cluster to the same barrier in both CTAs.
#include <cuda.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace rnnt {
namespace {

constexpr int G2_BM = 128;            // rows of A per CTA (pair: 256)
constexpr int G2_BN = 256;            // output columns of the pair tile
constexpr int G2_STAGES = 3;
constexpr int G2_A_BYTES = 2 * G2_BM * 128;          // hi + lo, 32 KB
constexpr int G2_BH_BYTES = (G2_BN / 2) * 128;       // one part (hi or lo) of this CTA's half of the W tile, 16 KB
constexpr int G2_STAGE_BYTES = G2_A_BYTES + 2 * G2_BH_BYTES;   // 64 KB
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + 1024 + 256;
constexpr int G2_THREADS = 192;

struct Gemm2Args {
  const uint8_t* a_img;  // image(TR=128) of A, an even number of row tiles
  const uint8_t* b_img;  // image(TR=256) of W
  const float* bias;
  float* C;
  int ldc;
  int64_t M;
  int N, KB;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// 2-D tensor-map load of one 128 x 128 B box into this CTA's shared memory; the transaction bytes are credited to the
// barrier at the same offset in the even (leader) CTA of the pair (peer bit of the shared::cluster address cleared)
__device__ __forceinline__ void tma_2d_pair(void* smem_dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar) & 0xFEFFFFFFu)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {  // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// all MMAs issued so far by this thread arrive on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_tc2_f16x3_kernel(Gemm2Args p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES);   // (leader) both CTAs' stage landed
  uint64_t* empty = full + G2_STAGES;    // the MMAs that read the stage (in both CTAs) have completed
  uint64_t* tfull = empty + G2_STAGES;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int nt = blockIdx.y;
  const int64_t mt = (int64_t)(blockIdx.x >> 1) * 2 + rank;   // this CTA's 128-row tile of A / C

  if (threadIdx.x == 0) {
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(tptr, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers exist before any TMA completion / multicast commit targets them
  tc_fence_after();
  const uint32_t tmem = *tptr;

  if (warp == 0) {
    for (int kb = 0; kb < p.KB; ++kb) {
      const int s = kb % G2_STAGES;
      const uint32_t ph = (kb / G2_STAGES) & 1;
      mbar_wait_cluster(&empty[s], ph ^ 1);
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * G2_STAGE_BYTES);   // both CTAs' four boxes
        uint8_t* dst = smem + s * G2_STAGE_BYTES;
        const int32_t rowA = (int32_t)(img_tile_offset(mt, kb, 0, p.KB, G2_BM) >> 7);                       // image rows of 128 B
        const int32_t rowW = (int32_t)(img_tile_offset(nt, kb, 0, p.KB, G2_BN) >> 7) + (int32_t)rank * (G2_BN / 2);
        tma_2d_pair(dst, &tmA, 0, rowA, &full[s]);                                       // A hi
        tma_2d_pair(dst + G2_BM * 128, &tmA, 0, rowA + G2_BM, &full[s]);                 // A lo
        tma_2d_pair(dst + G2_A_BYTES, &tmW, 0, rowW, &full[s]);                          // this CTA's half of W, hi part
        tma_2d_pair(dst + G2_A_BYTES + G2_BH_BYTES, &tmW, 0, rowW + G2_BN, &full[s]);    // ... lo part
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_f16(2 * G2_BM, G2_BN);
      for (int kb = 0; kb < p.KB; ++kb) {
        const int s = kb % G2_STAGES;
        const uint32_t ph = (kb / G2_STAGES) & 1;
        mbar_wait_cluster(&full[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + s * G2_STAGE_BYTES);
        const uint32_t b_base = a_base + G2_A_BYTES;
        const uint64_t a_hi = umma_desc_sw128(a_base), a_lo = umma_desc_sw128(a_base + G2_BM * 128);
        const uint64_t b_hi = umma_desc_sw128(b_base), b_lo = umma_desc_sw128(b_base + G2_BH_BYTES);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint32_t acc = (kb > 0 || k4 > 0) ? 1u : 0u;
          if (elect_one()) {
            tc_mma2_f16(tmem, a_hi + 2 * k4, b_hi + 2 * k4, idesc, acc);
            tc_mma2_f16(tmem + G2_BN, a_hi + 2 * k4, b_lo + 2 * k4, idesc, acc);
            tc_mma2_f16(tmem + G2_BN, a_lo + 2 * k4, b_hi + 2 * k4, idesc, 1u);
          }
        }
        if (elect_one()) tc_commit2(&empty[s]);
        __syncwarp();
      }
      if (elect_one()) tc_commit2(tfull);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int64_t row = mt * G2_BM + q * 32 + lane;
    mbar_wait_cluster(tfull, 0);
    tc_fence_after();
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < G2_BN; c0 += 16) {
      float d0[16], d1[16];
      tmem_ld16(tbase + c0, d0);
      tmem_ld16(tbase + G2_BN + c0, d1);
      tmem_ld_wait();
      const int col = nt * G2_BN + c0;
      if (row < p.M && col < p.N) {
        float* out = p.C + row * p.ldc + col;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 v;
          v.x = fmaf(d1[j + 0], kLoInv, d0[j + 0]);
          v.y = fmaf(d1[j + 1], kLoInv, d0[j + 1]);
          v.z = fmaf(d1[j + 2], kLoInv, d0[j + 2]);
          v.w = fmaf(d1[j + 3], kLoInv, d0[j + 3]);
          if (col + j + 3 < p.N) {
            if (p.bias) {
              const float4 b = *reinterpret_cast<const float4*>(p.bias + col + j);
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            *reinterpret_cast<float4*>(out + j) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the other may still signal or read it
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

}  // namespace

cudaError_t configure_gemm_tc2() {
  return cudaFuncSetAttribute(gemm_tc2_f16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES);
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
// an operand image seen as a 2-D byte array [rows][128]; boxes of 128 rows (the swizzle is already in the data)
bool image_map(CUtensorMap* m, const uint8_t* img, size_t bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {128, (cuuint64_t)(bytes >> 7)};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {128, 128};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(img), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// a_img must hold an EVEN number of 128-row tiles (gemm_tc_a_image_bytes rounds up)
cudaError_t launch_gemm_tc2(const uint8_t* a_img, const uint8_t* w_img, const float* bias, float* C, int ldc, int64_t M, int N,
                            int K, cudaStream_t st) {
  if ((ldc & 3) || (N & 3)) return cudaErrorInvalidValue;
  Gemm2Args a;
  a.a_img = a_img; a.b_img = w_img; a.bias = bias; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.KB = (int)ceil_div(K, kImgK);
  alignas(64) CUtensorMap tmA, tmW;
  if (!image_map(&tmA, a_img, gemm_tc_a_image_bytes(M, K)) || !image_map(&tmW, w_img, gemm_tc_w_image_bytes(N, K)))
    return cudaErrorNotSupported;
  dim3 grid((unsigned)(2 * ceil_div(M, 2 * G2_BM)), (unsigned)ceil_div(N, G2_BN));
  gemm_tc2_f16x3_kernel<<<grid, G2_THREADS, G2_SMEM_BYTES, st>>>(a, tmA, tmW);
  return cudaGetLastError();
}

}  // namespace rnnt
