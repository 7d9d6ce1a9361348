// Encoder LSTM recurrence (fp32 CUDA-core variant): one launch per (layer, time step).
// Replaces torch.nn.LSTM as driven by CustomRNN.forward_one_rnn (reference
// libreasr/lib/layers/custom_rnn.py:140-175); arithmetic per haste/lstm.py:51-60 with
// the native gate order i,f,g,o, plus the per-layer BatchNorm1d in eval mode
// (custom_rnn.py:210-213) folded into the epilogue.
//
// The hoisted input projection (x_t * W_ih^T + b_ih + b_hh for all t, one big GEMM per
// layer) arrives in `xp`; this kernel adds the recurrent product h_{t-1} * W_hh^T and
// applies the cell.  CTA (ut, bt) owns hidden units [8*ut, 8*ut+8) -- all four gates of a
// unit sit in adjacent columns of the repacked k-major W_hh -- and 32 batch columns.
#include "kernels.h"
#include "tile_gemm.cuh"

namespace rnnt {
namespace {

__global__ void __launch_bounds__(TG_THREADS) lstm_step_kernel(LstmStepArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int ut = blockIdx.x, b0 = blockIdx.y * kBatchTile;
  float v[4];
  tile_gemm<4>(p.Whh_t, 4 * p.H, ut * TG_UNITS * 4, p.hT_in, p.Bp, b0, p.H, smem, v);
  const int unit = ut * TG_UNITS + (threadIdx.x >> 5);
  const int b = b0 + (threadIdx.x & 31);
  if (b >= p.B) return;
  const size_t row = (size_t)b * p.T + p.t;
  const float4 x = *reinterpret_cast<const float4*>(p.xp + row * (size_t)(4 * p.H) + unit * 4);
  const size_t si = (size_t)unit * p.Bp + b;
  const float h_old = p.hT_in[si], c_old = p.cT[si];
  const float gi = sigmoidf_acc(v[0] + x.x);
  const float gf = sigmoidf_acc(v[1] + x.y);
  const float gg = tanhf(v[2] + x.z);
  const float go = sigmoidf_acc(v[3] + x.w);
  float c = gf * c_old + gi * gg;
  float h = go * tanhf(c);
  if (p.lens_T && p.t >= p.lens_T[b]) {  // past the end of a ragged utterance: state frozen
    c = c_old;
    h = h_old;
  }
  p.cT[si] = c;
  p.hT_out[si] = h;
  p.y[row * p.H + unit] = h * p.bn_scale[unit] + p.bn_shift[unit];
}

__global__ void state_to_T_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Bp, int H) {
  // in [B][H] -> out [H][Bp], padding columns zeroed
  __shared__ float tile[32][33];
  const int h0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int b = b0 + i, hh = h0 + threadIdx.x;
    tile[i][threadIdx.x] = (b < B && hh < H) ? in[(size_t)b * H + hh] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int hh = h0 + i, b = b0 + threadIdx.x;
    if (hh < H && b < Bp) out[(size_t)hh * Bp + b] = tile[threadIdx.x][i];
  }
}

__global__ void state_from_T_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Bp, int H) {
  __shared__ float tile[32][33];
  const int h0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int hh = h0 + i, b = b0 + threadIdx.x;
    tile[i][threadIdx.x] = (hh < H && b < Bp) ? in[(size_t)hh * Bp + b] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int b = b0 + i, hh = h0 + threadIdx.x;
    if (b < B && hh < H) out[(size_t)b * H + hh] = tile[threadIdx.x][i];
  }
}

__global__ void state_broadcast_T_kernel(const float* __restrict__ vec, float* __restrict__ out, int B, int Bp, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)H * Bp) return;
  const int hh = (int)(i / Bp), b = (int)(i % Bp);
  out[i] = (b < B) ? vec[hh] : 0.f;
}

__global__ void lens_to_steps_kernel(const int32_t* __restrict__ lens, int32_t* __restrict__ steps, int B, int hop,
                                     int n_stack, int D, int T_max) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int F = lens[b] / hop + 1;
  const int T = F >= n_stack ? (F - n_stack) / D + 1 : 0;
  steps[b] = min(T, T_max);
}

}  // namespace

cudaError_t launch_lens_to_steps(const int32_t* lens, int32_t* steps, int B, int hop, int n_stack, int D, int T_max,
                                 cudaStream_t st) {
  lens_to_steps_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, st>>>(lens, steps, B, hop, n_stack, D, T_max);
  return cudaGetLastError();
}

cudaError_t configure_lstm() {
  return cudaFuncSetAttribute(lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES);
}

cudaError_t launch_lstm_step(const LstmStepArgs& a, cudaStream_t st) {
  dim3 grid(a.H / TG_UNITS, a.Bp / kBatchTile);
  lstm_step_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_state_to_T(const float* in_BH, float* out_HB, int B, int Bp, int H, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(H, 32), (unsigned)ceil_div(Bp, 32));
  state_to_T_kernel<<<grid, dim3(32, 8), 0, st>>>(in_BH, out_HB, B, Bp, H);
  return cudaGetLastError();
}

cudaError_t launch_state_from_T(const float* in_HB, float* out_BH, int B, int Bp, int H, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(H, 32), (unsigned)ceil_div(Bp, 32));
  state_from_T_kernel<<<grid, dim3(32, 8), 0, st>>>(in_HB, out_BH, B, Bp, H);
  return cudaGetLastError();
}

cudaError_t launch_state_broadcast_T(const float* vec_H, float* out_HB, int B, int Bp, int H, cudaStream_t st) {
  const int64_t n = (int64_t)H * Bp;
  state_broadcast_T_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(vec_H, out_HB, B, Bp, H);
  return cudaGetLastError();
}

}  // namespace rnnt
