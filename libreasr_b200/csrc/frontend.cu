// Feature front-end: framed STFT (hann(400) in a 1024-point frame, hop 160, reflect
// padding) -> power -> HTK mel filterbank -> log(x + 1e-6) -> 10-frame stack / 8-frame
// hop, written directly in the reference's (mel-major, stack-minor) feature layout.
// Replaces TransformTime.encodes + StackDownsample.encodes (+ StreamPostprocess for the
// serving window): reference libreasr/lib/transforms.py:301-323, 335-342, 436-441.
//
// One CTA per output row (b, t): warp s computes frame t*D + frame0 + s entirely in
// shared memory (real 1024-point FFT as a 512-point complex FFT + untangle; only the 400
// windowed samples are non-zero), the CTA then writes the X = n_mels*n_stack floats of
// the row with one coalesced pass.  Audio is read once per frame (25% of frames are
// shared by two rows and recomputed: the kernel is FFT-, not HBM-bound).
#include "kernels.h"

namespace rnnt {

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(kFrontendMaxWarps * 32)
mel_stack_kernel(FrontendArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x, b = blockIdx.y;

  float2* tw = reinterpret_cast<float2*>(smem);                 // [512] exp(-2 pi i j / 1024)
  float* lm = smem + 1024;                                      // [n_stack][n_mels]
  float* wbase = lm + p.n_stack * p.n_mels;
  wbase += (4 - ((p.n_stack * p.n_mels) & 3)) & 3;              // keep float2/float4 alignment
  float2* z = reinterpret_cast<float2*>(wbase + warp * (1024 + 520));  // [512] complex
  float* P = wbase + warp * (1024 + 520) + 1024;                // [513] power spectrum

  for (int i = threadIdx.x; i < 512; i += blockDim.x) tw[i] = p.tw[i];

  const int len = p.lens ? p.lens[b] : (int)p.n;
  const int F = len / p.hop + 1;
  const int Tb = (F >= p.n_stack) ? (F - p.n_stack) / p.D + 1 : 0;
  float* orow = p.out + ((size_t)b * p.T_out + t) * (size_t)(p.n_mels * p.n_stack);
  const bool valid = p.is_stream ? true : (t < Tb);
  if (!valid) {  // uniform per CTA
    for (int i = threadIdx.x; i < p.n_mels * p.n_stack; i += blockDim.x) orow[i] = 0.f;
    return;
  }
  __syncthreads();

  // ---- load + window, bit-reversed placement (z[m] = x[2m] + i x[2m+1]) ----
  const float* x = p.audio + (size_t)b * p.n;
  const int f = t * p.D + p.frame0 + warp;
  const int start = f * p.hop - p.win / 2;
  const int half_win = p.win / 2;
  for (int m = lane; m < 512; m += 32) {
    float2 v = make_float2(0.f, 0.f);
    if (m < half_win) {
      int i0 = start + 2 * m, i1 = i0 + 1;
      if (i0 < 0) i0 = -i0;
      if (i1 < 0) i1 = -i1;
      if (i0 >= len) i0 = 2 * (len - 1) - i0;
      if (i1 >= len) i1 = 2 * (len - 1) - i1;
      i0 = min(max(i0, 0), len - 1);
      i1 = min(max(i1, 0), len - 1);
      v.x = __ldg(x + i0) * __ldg(p.window + 2 * m);
      v.y = __ldg(x + i1) * __ldg(p.window + 2 * m + 1);
    }
    z[__brev((unsigned)m) >> 23] = v;
  }
  __syncwarp();

  // ---- 512-point complex radix-2 DIT FFT, 9 stages, 8 butterflies per lane per stage ----
#pragma unroll 1
  for (int s = 0; s < 9; ++s) {
    const int half = 1 << s;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int q = lane + 32 * r;
      const int pos = q & (half - 1);
      const int i = ((q >> s) << (s + 1)) + pos;
      const int j = i + half;
      const float2 w = tw[(pos << (8 - s)) * 2];  // exp(-2 pi i pos / (2 half)) from the 1024-table
      const float2 a = z[i];
      const float2 bb = cmul(z[j], w);
      z[i] = make_float2(a.x + bb.x, a.y + bb.y);
      z[j] = make_float2(a.x - bb.x, a.y - bb.y);
    }
    __syncwarp();
  }

  // ---- untangle the real FFT and take the power: P[k], P[512-k] for k = 0..256 ----
  for (int k = lane; k <= 256; k += 32) {
    const float2 zk = z[k];
    const float2 zn = z[(512 - k) & 511];
    const float2 xe = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
    const float2 xo = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
    const float2 wx = cmul(tw[k], xo);
    const float ar = xe.x + wx.x, ai = xe.y + wx.y;
    const float br = xe.x - wx.x, bi = xe.y - wx.y;
    P[k] = ar * ar + ai * ai;
    P[512 - k] = br * br + bi * bi;
  }
  __syncwarp();

  // ---- sparse triangular mel filterbank + log ----
  for (int m = lane; m < p.n_mels; m += 32) {
    const int k0 = p.mel_start[m], cnt = p.mel_count[m];
    const float* w = p.mel_w + p.mel_off[m];
    float s = 0.f;
    for (int i = 0; i < cnt; ++i) s = fmaf(w[i], P[k0 + i], s);
    lm[warp * p.n_mels + m] = logf(s + p.log_offset);
  }
  __syncthreads();

  // ---- stacked row: feature index = m * n_stack + s (transforms.py:439-440) ----
  const int X = p.n_mels * p.n_stack;
  for (int i = threadIdx.x; i < X; i += blockDim.x) {
    const int m = i / p.n_stack, s = i - m * p.n_stack;
    orow[i] = lm[s * p.n_mels + m];
  }
}

// LayerNorm over the feature axis (Encoder.input_norm, models.py:84,107): one warp per row.
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             int64_t rows, int X, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* r = in + (size_t)warp * X;
  float s = 0.f;
  for (int i = lane; i < X; i += 32) s += r[i];
  const float mean = warp_sum(s) / (float)X;
  float v = 0.f;
  for (int i = lane; i < X; i += 32) {
    const float d = r[i] - mean;
    v = fmaf(d, d, v);
  }
  const float inv = 1.0f / sqrtf(warp_sum(v) / (float)X + eps);
  float* o = out + (size_t)warp * X;
  for (int i = lane; i < X; i += 32) o[i] = (r[i] - mean) * inv * gamma[i] + beta[i];
}

}  // namespace

size_t frontend_smem_bytes(int n_stack, int n_mels) {
  size_t fl = 1024 + (size_t)n_stack * n_mels + 4 + (size_t)n_stack * (1024 + 520);
  return fl * sizeof(float);
}

cudaError_t launch_mel_stack(const FrontendArgs& a, int B, cudaStream_t st) {
  if (a.n_stack > kFrontendMaxWarps) return cudaErrorInvalidValue;
  const size_t smem = frontend_smem_bytes(a.n_stack, a.n_mels);
  cudaError_t e = cudaFuncSetAttribute(mel_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(a.T_out, B);
  mel_stack_kernel<<<grid, a.n_stack * 32, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_layernorm(const float* in, float* out, const float* gamma, const float* beta, int64_t rows, int X,
                             float eps, cudaStream_t st) {
  const int64_t blocks = ceil_div(rows, 8);
  layernorm_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, gamma, beta, rows, X, eps);
  return cudaGetLastError();
}

}  // namespace rnnt
