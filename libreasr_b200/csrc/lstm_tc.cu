// Persistent tcgen05 LSTM layer: the recurrent half of the LSTM gate GEMMs on the 5th-gen
// tensor cores, one launch per layer for ALL time steps.
// Replaces torch.nn.LSTM as driven by CustomRNN.forward_one_rnn (reference
// libreasr/lib/layers/custom_rnn.py:140-175, arithmetic haste/lstm.py:51-60, gate order
// i,f,g,o) + the per-layer BatchNorm1d eval (custom_rnn.py:210-213).
//
// Decomposition.  CTA c owns U hidden units = NC = 4U interleaved gate rows of W_hh and
// keeps that slice (hi+lo fp16 operand image, NC x H x 4 bytes) RESIDENT in shared memory
// for the whole layer when it fits (H = 1024: 128 KB); otherwise it is streamed per step.
// Each step is one skinny GEMM with the BATCH on the MMA M axis and the CTA's gate rows on
// N:   D[128 x NC] = h_{t-1}[B x H] * W_slice[NC x H]^T   (3xFP16 split, see gemm_tc.cu),
// so all H/U CTAs (128 of the 148 SMs at H = 1024) share every step.
//   warp 0     producer: waits for the grid-wide "h_{t-1} complete" counter, then streams the
//              h_{t-1} operand image (B rows x 64 k, hi|lo) per k-block with TMA bulk copies
//   warp 1     one thread issues 12 tcgen05.mma (M128 x NC x K16) per k-block into TMEM
//   warps 2-5  epilogue: thread (q, lane) owns batch row 32q+lane: tcgen05.ld its NC gate
//              pre-activations, adds the hoisted input projection xp, applies the cell with
//              c (and h) kept in REGISTERS across steps, writes h_t straight into the next
//              step's operand image (fp16 hi/lo, swizzled) and BatchNorm(h_t) both as fp32
//              and as the operand image of the next layer's input GEMM; then arrives on the
//              grid counter (release) that the producers of all CTAs poll (acquire).
// The only grid-wide synchronisation is that one counter per step.
#include "kernels.h"
#include "tc_common.cuh"

namespace rnnt {
namespace {

constexpr int LT_THREADS = 192;
constexpr int LT_MAX_U = 16;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct Smem {
  uint8_t* ring;
  uint8_t* wres;
  uint64_t *full, *empty, *tfull, *tempty, *wfull;
  uint32_t* tptr;
};

__global__ void __launch_bounds__(LT_THREADS, 1) lstm_layer_tc_kernel(LstmTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int NC = p.NC, U = p.U, KB = p.KB, S = p.stages;
  const uint32_t xrows_bytes = (uint32_t)p.Bpad8 * 128;      // valid rows of one part of the h tile
  const uint32_t wtile = (uint32_t)NC * 256;                 // hi + lo weight tile of one k-block
  const uint32_t stage_bytes = 2 * xrows_bytes + (p.w_resident ? 0 : wtile);
  Smem sm;
  sm.ring = base;
  sm.wres = base + (size_t)S * stage_bytes;
  uint8_t* after = sm.wres + (p.w_resident ? (size_t)KB * wtile : 0);
  // the MMA reads 128 rows per A tile; rows >= Bpad8 alias whatever follows (never used): keep >= 16 KB mapped
  uint8_t* bars = base + p.bar_offset;
  (void)after;
  sm.full = reinterpret_cast<uint64_t*>(bars);
  sm.empty = sm.full + S;
  sm.tfull = sm.empty + S;
  sm.tempty = sm.tfull + 1;
  sm.wfull = sm.tempty + 1;
  sm.tptr = reinterpret_cast<uint32_t*>(sm.wfull + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(sm.tfull, 1);
    mbar_init(sm.tempty, 128);
    mbar_init(sm.wfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(sm.tptr, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sm.tptr;

  if (warp == 0) {
    // =========================== producer ===========================
    if (lane == 0) {
      if (p.w_resident) {
        mbar_arrive_expect_tx(sm.wfull, (uint32_t)KB * wtile);
        for (int kb = 0; kb < KB; ++kb)
          tma_bulk_g2s(sm.wres + (size_t)kb * wtile, p.w_img + img_tile_offset(cta, kb, 0, KB, NC), wtile, sm.wfull);
      }
      uint32_t g = 0;
      for (int t = 0; t < p.T; ++t) {
        const unsigned target = (unsigned)(t + 1) * (unsigned)G;
        while (ld_acquire_u32(p.barrier) < target) {
        }
        fence_proxy_async_global();
        const uint8_t* ximg = p.x_img[t & 1];
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(&sm.empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&sm.full[s], stage_bytes);
          uint8_t* dst = sm.ring + (size_t)s * stage_bytes;
          tma_bulk_g2s(dst, ximg + img_tile_offset(0, kb, 0, KB, 128), xrows_bytes, &sm.full[s]);
          tma_bulk_g2s(dst + xrows_bytes, ximg + img_tile_offset(0, kb, 1, KB, 128), xrows_bytes, &sm.full[s]);
          if (!p.w_resident)
            tma_bulk_g2s(dst + 2 * xrows_bytes, p.w_img + img_tile_offset(cta, kb, 0, KB, NC), wtile, &sm.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, NC);
      if (p.w_resident) mbar_wait(sm.wfull, 0);
      uint32_t g = 0;
      for (int t = 0; t < p.T; ++t) {
        if (t > 0) mbar_wait(sm.tempty, (t - 1) & 1);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(&sm.full[s], ph);
          tc_fence_after();
          const uint32_t a_hi0 = smem_u32(sm.ring + (size_t)s * stage_bytes);
          const uint32_t a_lo0 = a_hi0 + xrows_bytes;
          const uint32_t b_hi0 = p.w_resident ? smem_u32(sm.wres + (size_t)kb * wtile) : a_hi0 + 2 * xrows_bytes;
          const uint32_t b_lo0 = b_hi0 + (uint32_t)NC * 128;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t a_hi = umma_desc_sw128(a_hi0 + k4 * 32), a_lo = umma_desc_sw128(a_lo0 + k4 * 32);
            const uint64_t b_hi = umma_desc_sw128(b_hi0 + k4 * 32), b_lo = umma_desc_sw128(b_lo0 + k4 * 32);
            const uint32_t acc = (kb > 0 || k4 > 0) ? 1u : 0u;
            tc_mma_f16(tmem, a_hi, b_hi, idesc, acc);
            tc_mma_f16(tmem + NC, a_hi, b_lo, idesc, acc);
            tc_mma_f16(tmem + NC, a_lo, b_hi, idesc, 1u);
          }
          tc_commit(&sm.empty[s]);
        }
        tc_commit(sm.tfull);
      }
    }
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;
    const int b = q * 32 + lane;        // batch row == TMEM lane
    const bool valid = b < p.B;
    const int et = (warp - 2) * 32 + lane;  // 0..127
    const int unit0 = cta * U;
    const int H = p.H;
    float c[LT_MAX_U], h[LT_MAX_U];
    const int len = valid ? (p.lens_T ? min(p.lens_T[b], p.T) : p.T) : 0;
#pragma unroll
    for (int j = 0; j < LT_MAX_U; ++j) {
      if (j < U && valid) {
        h[j] = p.state_h_in ? p.state_h_in[(size_t)b * H + unit0 + j] : p.h_init_vec[unit0 + j];
        c[j] = p.state_c_in ? p.state_c_in[(size_t)b * H + unit0 + j] : p.c_init_vec[unit0 + j];
      } else {
        h[j] = 0.f; c[j] = 0.f;
      }
    }
    // writes U consecutive values (k = unit0 ..) of row `r` of a TR=128 image tile set as hi/lo halves
    auto store_img = [&](uint8_t* img, int64_t row_tile, int r, const float* v) {
#pragma unroll
      for (int j0 = 0; j0 < LT_MAX_U; j0 += 4) {
        if (j0 < U) {
          const int k = unit0 + j0;
          __align__(8) __half hi[4];
          __align__(8) __half lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split_f16x3(v[j0 + j], hi[j], lo[j]);
          const uint32_t off = img_elem_offset(r, k & 63);
          *reinterpret_cast<uint2*>(img + img_tile_offset(row_tile, k >> 6, 0, KB, 128) + off) = *reinterpret_cast<const uint2*>(hi);
          *reinterpret_cast<uint2*>(img + img_tile_offset(row_tile, k >> 6, 1, KB, 128) + off) = *reinterpret_cast<const uint2*>(lo);
        }
      }
    };
    if (valid) store_img(p.x_img[0], 0, b, h);
    fence_proxy_async_global();
    named_bar_sync(1, 128);
    if (et == 0) {
      __threadfence();
      atomicAdd(p.barrier, 1u);
    }

    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < p.T; ++t) {
      const int64_t row = (int64_t)b * p.T + t;
      mbar_wait(sm.tfull, t & 1);
      tc_fence_after();
      float hy[LT_MAX_U];
#pragma unroll
      for (int j0 = 0; j0 < LT_MAX_U; j0 += 4) {  // 16 gate columns = 4 units per chunk
        if (j0 < U) {
          float d0[16], d1[16];
          tmem_ld16(tlane + j0 * 4, d0);
          tmem_ld16(tlane + NC + j0 * 4, d1);
          tmem_ld_wait();
          if (valid) {
            const float4* xr = reinterpret_cast<const float4*>(p.xp + row * (size_t)(4 * H) + (size_t)(unit0 + j0) * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 x = xr[j];
              const float vi = fmaf(d1[4 * j + 0], kLoInv, d0[4 * j + 0]) + x.x;
              const float vf = fmaf(d1[4 * j + 1], kLoInv, d0[4 * j + 1]) + x.y;
              const float vg = fmaf(d1[4 * j + 2], kLoInv, d0[4 * j + 2]) + x.z;
              const float vo = fmaf(d1[4 * j + 3], kLoInv, d0[4 * j + 3]) + x.w;
              if (t < len) {
                const float cn = sigmoidf_acc(vf) * c[j0 + j] + sigmoidf_acc(vi) * tanhf(vg);
                c[j0 + j] = cn;
                h[j0 + j] = sigmoidf_acc(vo) * tanhf(cn);
              }
              hy[j0 + j] = h[j0 + j] * p.bn_scale[unit0 + j0 + j] + p.bn_shift[unit0 + j0 + j];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(sm.tempty);
      if (valid) {
        store_img(p.x_img[(t + 1) & 1], 0, b, h);
        if (p.y) {
#pragma unroll
          for (int j0 = 0; j0 < LT_MAX_U; j0 += 4)
            if (j0 < U)
              *reinterpret_cast<float4*>(p.y + row * H + unit0 + j0) = make_float4(hy[j0], hy[j0 + 1], hy[j0 + 2], hy[j0 + 3]);
        }
        if (p.y_img) store_img(p.y_img, row >> 7, (int)(row & 127), hy);
      }
      fence_proxy_async_global();
      named_bar_sync(1, 128);
      if (et == 0) {
        __threadfence();
        atomicAdd(p.barrier, 1u);
      }
    }
    if (valid) {
#pragma unroll
      for (int j = 0; j < LT_MAX_U; ++j) {
        if (j < U) {
          if (p.state_h_out) p.state_h_out[(size_t)b * H + unit0 + j] = h[j];
          if (p.state_c_out) p.state_c_out[(size_t)b * H + unit0 + j] = c[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, p.tmem_cols);
}

}  // namespace

cudaError_t configure_lstm_tc() {
  return cudaFuncSetAttribute(lstm_layer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

// Chooses the decomposition for hidden size H on a device with `sms` SMs; returns false if none fits.
bool lstm_tc_plan(int H, int B, int sms, LstmTcPlan* pl) {
  if (B < 1 || B > 128 || H % 64) return false;
  int U = 0;
  for (int u : {4, 8, 16})
    if (H % u == 0 && H / u <= sms) { U = u; break; }
  if (!U) return false;
  pl->U = U;
  pl->NC = 4 * U;
  pl->grid = H / U;
  pl->KB = H / 64;
  pl->Bpad8 = (int)round_up(B, 8);
  const size_t xstage = (size_t)2 * pl->Bpad8 * 128, wtile = (size_t)pl->NC * 256;
  const size_t budget = 227 * 1024 - 2048 /*align + barriers*/ - 16384 /*guard rows*/;
  pl->w_resident = ((size_t)pl->KB * wtile + 4 * xstage <= budget) ? 1 : 0;
  const size_t stage = xstage + (pl->w_resident ? 0 : wtile);
  size_t avail = budget - (pl->w_resident ? (size_t)pl->KB * wtile : 0);
  int S = (int)(avail / stage);
  if (S > 8) S = 8;
  if (S < 2) return false;
  pl->stages = S;
  const size_t used = (size_t)S * stage + (pl->w_resident ? (size_t)pl->KB * wtile : 0) + 16384;
  pl->bar_offset = (int)round_up(used, 1024);
  pl->smem_bytes = pl->bar_offset + 1024 + 1024;
  int cols = 32;
  while (cols < 2 * pl->NC) cols *= 2;
  pl->tmem_cols = cols;
  return pl->smem_bytes <= 227 * 1024;
}

cudaError_t launch_lstm_layer_tc(const LstmTcArgs& a, const LstmTcPlan& pl, cudaStream_t st) {
  LstmTcArgs args = a;
  args.U = pl.U; args.NC = pl.NC; args.KB = pl.KB; args.Bpad8 = pl.Bpad8; args.stages = pl.stages;
  args.w_resident = pl.w_resident; args.bar_offset = pl.bar_offset; args.tmem_cols = pl.tmem_cols;
  void* kargs[] = {&args};
  return cudaLaunchCooperativeKernel((void*)lstm_layer_tc_kernel, dim3(pl.grid), dim3(LT_THREADS), kargs, pl.smem_bytes, st);
}

}  // namespace rnnt
