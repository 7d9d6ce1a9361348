// Skinny-GEMM tile primitive shared by the LSTM step kernel and the decode-loop phases.
//
// One CTA (256 threads) computes, for a tile of 8 "units" (G weight rows each) and 32
// batch columns,
//     out[u][g][b] = sum_k Wt[k][wcol0 + u*G + g] * Xt[k][b0 + b]        (fp32 FMA)
// Wt is the weight matrix stored k-major ([K][ldw], repacked once at finalize), Xt the
// activation vector stored feature-major ([K][ldx], ldx = padded batch), so both stage
// into shared memory with plain 16-byte cp.async rows (no transposition) through a
// 4-stage ring.  The 8 warps split every 64-deep k chunk 8 ways (split-K inside the
// CTA: at batch 32 the tile has too few outputs to occupy 256 threads otherwise); lane
// (u, tc) keeps a G x 8 register tile; the 8 partial tiles are reduced through shared
// memory and thread tid ends up owning out[u = tid/32][0..G)[b = tid%32].
#pragma once
#include "common.cuh"

namespace rnnt {

constexpr int TG_THREADS = 256;
constexpr int TG_UNITS = 8;
constexpr int TG_KC = 64;
constexpr int TG_STAGES = 4;
constexpr int TG_STAGE_FLOATS = TG_KC * 32 * 2;                  // W region (<=32 floats/row) + X region
constexpr int TG_SMEM_FLOATS = TG_STAGES * TG_STAGE_FLOATS;      // 64 KB
constexpr int TG_SMEM_BYTES = TG_SMEM_FLOATS * 4;

template <int G>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ Wt, int ldw, int wcol0,
                                          const float* __restrict__ Xt, int ldx, int b0, int K,
                                          float* smem, float (&out)[G]) {
  static_assert(G == 3 || G == 4, "3 (GRU) or 4 (LSTM / plain) rows per unit");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u = lane >> 2, tc = lane & 3;
  constexpr int WROW = TG_UNITS * G;
  constexpr int WPIECES = WROW / 4;
  float acc[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[g][j] = 0.f;
  const int nchunks = K / TG_KC;

  auto load_stage = [&](int chunk, int stage) {
    float* ws = smem + stage * TG_STAGE_FLOATS;
    float* xs = ws + TG_KC * 32;
    const int k0 = chunk * TG_KC;
    for (int p = tid; p < TG_KC * WPIECES; p += TG_THREADS) {
      const int r = p / WPIECES, c = p % WPIECES;
      cp_async16(ws + r * WROW + c * 4, Wt + (size_t)(k0 + r) * ldw + wcol0 + c * 4);
    }
    for (int p = tid; p < TG_KC * 8; p += TG_THREADS) {
      const int r = p >> 3, c = p & 7;
      cp_async16(xs + r * 32 + c * 4, Xt + (size_t)(k0 + r) * ldx + b0 + c * 4);
    }
  };

#pragma unroll
  for (int s = 0; s < TG_STAGES - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int i = 0; i < nchunks; ++i) {
    cp_async_wait<TG_STAGES - 2>();
    __syncthreads();
    {
      const int nxt = i + TG_STAGES - 1;
      if (nxt < nchunks) load_stage(nxt, nxt % TG_STAGES);
      cp_async_commit();
    }
    const float* ws = smem + (i % TG_STAGES) * TG_STAGE_FLOATS;
    const float* xs = ws + TG_KC * 32;
#pragma unroll
    for (int j = 0; j < TG_KC / 8; ++j) {
      const int k = warp * (TG_KC / 8) + j;
      float w[G];
      if (G == 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(ws + k * WROW + u * 4);
        w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[G - 1] = w4.w;
      } else {
#pragma unroll
        for (int g = 0; g < G; ++g) w[g] = ws[k * WROW + u * G + g];
      }
      const float4 x0 = *reinterpret_cast<const float4*>(xs + k * 32 + tc * 8);
      const float4 x1 = *reinterpret_cast<const float4*>(xs + k * 32 + tc * 8 + 4);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        acc[g][0] = fmaf(w[g], x0.x, acc[g][0]);
        acc[g][1] = fmaf(w[g], x0.y, acc[g][1]);
        acc[g][2] = fmaf(w[g], x0.z, acc[g][2]);
        acc[g][3] = fmaf(w[g], x0.w, acc[g][3]);
        acc[g][4] = fmaf(w[g], x1.x, acc[g][4]);
        acc[g][5] = fmaf(w[g], x1.y, acc[g][5]);
        acc[g][6] = fmaf(w[g], x1.z, acc[g][6]);
        acc[g][7] = fmaf(w[g], x1.w, acc[g][7]);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // cross-warp (split-K) reduction through shared memory; red aliases the stage ring
  float* red = smem;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float* dst = red + ((warp * WROW + u * G + g) * 32 + tc * 8);
    *reinterpret_cast<float4*>(dst) = make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[g][4], acc[g][5], acc[g][6], acc[g][7]);
  }
  __syncthreads();
  const int uu = tid >> 5, bb = tid & 31;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[(w * WROW + uu * G + g) * 32 + bb];
    out[g] = s;
  }
  __syncthreads();
}

}  // namespace rnnt
