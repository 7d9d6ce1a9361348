// fp32 CUDA-core GEMM used by gemm_mode FP32_SIMT (the bit-level parity mode) and by the
// one-off weight repack at finalize():  C[M,N] = A[M,K] * W[N,K]^T + bias[N].
// 128x128x16 CTA tile, 256 threads, 8x8 register tile, double-buffered shared memory with
// register prefetch.  The tensor-core path (gemm_tc.cu) replaces it for the hoisted LSTM
// input projections and the joint encoder projection in the TC modes.
#include "kernels.h"

namespace rnnt {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

__global__ void __launch_bounds__(256) gemm_nt_f32_kernel(const float* __restrict__ A, int lda,
                                                          const float* __restrict__ W, int ldw,
                                                          const float* __restrict__ bias, float* __restrict__ C,
                                                          int ldc, int64_t M, int N, int K) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  // loader mapping: each thread loads two float4 of A and two of W per k-tile
  const int lrow = tid >> 2;          // 0..63 (+64)
  const int lk = (tid & 3) * 4;       // 0,4,8,12
  // compute mapping
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = m0 + lrow + 64 * h;
      const int k = k0 + lk;
      ra[h] = (r < M && k < K) ? *reinterpret_cast<const float4*>(A + r * lda + k) : make_float4(0, 0, 0, 0);
      const int c = n0 + lrow + 64 * h;
      rb[h] = (c < N && k < K) ? *reinterpret_cast<const float4*>(W + (int64_t)c * ldw + k) : make_float4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + 64 * h;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y; Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  const int ntiles = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) gload((t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue: rows {ty*4..+3, 64+ty*4..+3}, cols {tx*4..+3, 64+tx*4..+3}
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = acc[i][jh * 4 + j];
        if (bias && c + j < N) v[j] += bias[c + j];
      }
      if (c + 3 < N && ((ldc & 3) == 0)) {
        *reinterpret_cast<float4*>(C + r * ldc + c) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < N) C[r * ldc + c + j] = v[j];
      }
    }
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, int ld_in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)r * ld_in + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ src_row,
                                   int rows_out, int cols) {
  const int r = blockIdx.x;
  if (r >= rows_out) return;
  const float* s = in + (size_t)src_row[r] * cols;
  float* d = out + (size_t)r * cols;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) d[c] = s[c];
}

}  // namespace

cudaError_t launch_gemm_nt_f32(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                               int64_t M, int N, int K, cudaStream_t st) {
  if ((K & 3) || (lda & 3) || (ldw & 3)) return cudaErrorInvalidValue;
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
  gemm_nt_f32_kernel<<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
  return cudaGetLastError();
}

cudaError_t launch_transpose(const float* in, int ld_in, float* out, int rows, int cols, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(in, ld_in, out, rows, cols);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* in, float* out, const int* src_row, int rows_out, int cols, cudaStream_t st) {
  gather_rows_kernel<<<rows_out, 256, 0, st>>>(in, out, src_row, rows_out, cols);
  return cudaGetLastError();
}

}  // namespace rnnt
