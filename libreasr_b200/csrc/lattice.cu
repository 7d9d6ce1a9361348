// Training-time forward of the transducer: the joint network on the full T x U lattice and the
// RNN-T loss on top of it, in eval mode (no autograd: loss values / validation, not a training step).
// Replaces Transducer.forward (reference libreasr/lib/models.py:308-359: encoder, teacher-forced predictor over cat(bos, y),
// joint on the broadcast [N,T,U,H] pair, log_softmax) and what get_loss_func("rnnt") computes from its output
// (libreasr/lib/loss.py:72-110 -> warp_rnnt.rnnt_loss(..., average_frames=False): per-sequence negative log-likelihood by the
// forward recursion over the lattice; restated from its definition -- warp_rnnt is not vendored, oracle/rnnt_loss.py).
//
// The [N,T,U,V] tensor is never needed as a whole for the loss: rows (n, t, u) are processed in chunks that stay L2-resident --
// z = tanh(ep[n,t] + pp[n,u]) is written straight into a tcgen05 operand image, one 3xFP16 GEMM with W2 gives the chunk's
// logits, and one pass reduces every row to (log p(blank), log p(y_{u+1})); only those two lattices ([N,T,U] each) reach the
// alpha recursion.  The full log_softmax lattice is written only when the caller asks for Transducer.forward's return value.
#include "kernels.h"
#include "tc_common.cuh"

namespace rnnt {
namespace {

// rows [row0, row0 + rows) of the lattice (row = (n*T + t)*U + u) -> operand image (128-row tiles) of z = tanh(ep[n,t] + pp[n,u])
__global__ void __launch_bounds__(256) lattice_z_image_kernel(const float* __restrict__ pp, const float* __restrict__ ep, int T, int U, int64_t row0,
                                                              int rows, int J, uint8_t* __restrict__ img) {
  const int KB = J / kImgK;
  const int rt = blockIdx.y, kb = blockIdx.x;
  uint8_t* hi_t = img + img_tile_offset(rt, kb, 0, KB, 128);
  uint8_t* lo_t = img + img_tile_offset(rt, kb, 1, KB, 128);
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const int lr = rt * 128 + r;
    const int k0 = kb * kImgK + c * 8;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.f;
    if (lr < rows) {
      const int64_t row = row0 + lr;
      const int64_t nt = row / U;                  // n*T + t
      const int u = (int)(row - nt * U);
      const int64_t n = nt / T;
      const float* p0 = pp + (size_t)(n * U + u) * J + k0;
      const float* e0 = ep + (size_t)nt * J + k0;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = tanhf(p0[j] + e0[j]);
    }
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_f16x3(x[j], h[j], l[j]);
    const uint32_t off = img_elem_offset(r, c * 8);
    *reinterpret_cast<uint4*>(hi_t + off) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo_t + off) = *reinterpret_cast<const uint4*>(l);
  }
}

// one block per lattice row: log-sum-exp of the V logits; log p(blank), log p(label u) of the row; optionally the whole
// log_softmax row (Transducer.forward's return value)
__global__ void __launch_bounds__(256) lattice_lse_kernel(const float* __restrict__ logits, int V, int64_t row0, const int32_t* __restrict__ labels, int T,
                                                          int U, int Umax, int blank, float* __restrict__ lp_blank, float* __restrict__ lp_label,
                                                          float* __restrict__ lattice_out) {
  const int lr = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (size_t)lr * V;
  const int64_t row = row0 + lr;
  __shared__ float red[8];
  __shared__ float bc[2];
  float m = -INFINITY;
  for (int k = tid; k < V; k += 256) m = fmaxf(m, x[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float mm = red[0];
    for (int w = 1; w < 8; ++w) mm = fmaxf(mm, red[w]);
    bc[0] = mm;
  }
  __syncthreads();
  m = bc[0];
  float s = 0.f;
  for (int k = tid; k < V; k += 256) s += expf(x[k] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) {
    float ss = 0.f;
    for (int w = 0; w < 8; ++w) ss += red[w];
    bc[1] = m + logf(ss);
  }
  __syncthreads();
  const float lse = bc[1];
  if (tid == 0) {
    const int64_t nt = row / U;
    const int u = (int)(row - nt * U);
    const int64_t n = nt / T;
    lp_blank[row] = x[blank] - lse;
    lp_label[row] = u < Umax ? x[labels[(size_t)n * Umax + u]] - lse : -INFINITY;
  }
  if (lattice_out)
    for (int k = tid; k < V; k += 256) lattice_out[(size_t)row * V + k] = x[k] - lse;
}

// (log p(blank), log p(label)) of every lattice point from a full log-probability lattice [N,T,U,V] (loss of a given `inp`)
__global__ void lattice_gather_kernel(const float* __restrict__ lat, int64_t rows, int V, const int32_t* __restrict__ labels, int T, int U, int Umax,
                                      int blank, float* __restrict__ lp_blank, float* __restrict__ lp_label) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int64_t nt = row / U;
  const int u = (int)(row - nt * U);
  const int64_t n = nt / T;
  lp_blank[row] = lat[(size_t)row * V + blank];
  lp_label[row] = u < Umax ? lat[(size_t)row * V + labels[(size_t)n * Umax + u]] : -INFINITY;
}

__device__ __forceinline__ double logaddexp_d(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const double hi = a > b ? a : b, lo = a > b ? b : a;
  return hi + log1p(exp(lo - hi));
}

// RNN-T forward recursion, one block per sequence, anti-diagonal wavefront (fp64):
//   alpha(t,u) = logaddexp(alpha(t-1,u) + lp_blank(t-1,u), alpha(t,u-1) + lp_label(t,u-1));  loss = -(alpha(Tn-1,Un) + lp_blank(Tn-1,Un))
__global__ void rnnt_alpha_kernel(const float* __restrict__ lp_blank, const float* __restrict__ lp_label, const int32_t* __restrict__ xl,
                                  const int32_t* __restrict__ yl, int T, int U, double* __restrict__ loss) {
  extern __shared__ double sm[];   // two diagonals of U entries
  const int n = blockIdx.x;
  const int Tn = xl ? min(xl[n], T) : T, Un = min(yl[n], U - 1);
  if (Tn < 1) { if (threadIdx.x == 0) loss[n] = 0.0; return; }
  const float* B = lp_blank + (size_t)n * T * U;
  const float* L = lp_label + (size_t)n * T * U;
  double* prev = sm;
  double* cur = sm + U;
  for (int u = threadIdx.x; u < U; u += blockDim.x) { prev[u] = -INFINITY; cur[u] = -INFINITY; }
  __syncthreads();
  for (int d = 0; d <= Tn - 1 + Un; ++d) {   // diagonal d holds (t, u) with t + u = d; entry index = u
    for (int u = threadIdx.x; u <= Un; u += blockDim.x) {
      const int t = d - u;
      if (t < 0 || t >= Tn) continue;
      double v;
      if (d == 0) {
        v = 0.0;
      } else {
        const double a = t > 0 ? prev[u] + (double)B[(size_t)(t - 1) * U + u] : -INFINITY;          // from (t-1, u): blank
        const double b = u > 0 ? prev[u - 1] + (double)L[(size_t)t * U + (u - 1)] : -INFINITY;       // from (t, u-1): label
        v = logaddexp_d(a, b);
      }
      cur[u] = v;
    }
    __syncthreads();
    double* tmp = prev; prev = cur; cur = tmp;
    for (int u = threadIdx.x; u < U; u += blockDim.x) cur[u] = -INFINITY;
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[n] = -(prev[Un] + (double)B[(size_t)(Tn - 1) * U + Un]);
}

// column u of the teacher-forcing input cat(bos, y): tokens[n] = u == 0 ? bos : y[n][u-1]
__global__ void lattice_tokens_kernel(const int32_t* __restrict__ labels, int N, int Umax, int u, int bos, int32_t* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) out[n] = u == 0 ? bos : labels[(size_t)n * Umax + (u - 1)];
}

}  // namespace

cudaError_t launch_lattice_tokens(const int32_t* labels, int N, int Umax, int u, int bos, int32_t* out, cudaStream_t st) {
  lattice_tokens_kernel<<<(unsigned)ceil_div(N, 128), 128, 0, st>>>(labels, N, Umax, u, bos, out);
  return cudaGetLastError();
}
cudaError_t launch_lattice_z_image(const float* pp, const float* ep, int T, int U, int64_t row0, int rows, int J, uint8_t* img, cudaStream_t st) {
  dim3 grid((unsigned)(J / kImgK), (unsigned)ceil_div(rows, 128));
  lattice_z_image_kernel<<<grid, 256, 0, st>>>(pp, ep, T, U, row0, rows, J, img);
  return cudaGetLastError();
}
cudaError_t launch_lattice_lse(const float* logits, int rows, int V, int64_t row0, const int32_t* labels, int T, int U, int Umax, int blank,
                               float* lp_blank, float* lp_label, float* lattice_out, cudaStream_t st) {
  lattice_lse_kernel<<<rows, 256, 0, st>>>(logits, V, row0, labels, T, U, Umax, blank, lp_blank, lp_label, lattice_out);
  return cudaGetLastError();
}
cudaError_t launch_lattice_gather(const float* lat, int64_t rows, int V, const int32_t* labels, int T, int U, int Umax, int blank, float* lp_blank,
                                  float* lp_label, cudaStream_t st) {
  lattice_gather_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, st>>>(lat, rows, V, labels, T, U, Umax, blank, lp_blank, lp_label);
  return cudaGetLastError();
}
cudaError_t launch_rnnt_alpha(const float* lp_blank, const float* lp_label, const int32_t* xl, const int32_t* yl, int N, int T, int U, double* loss,
                              cudaStream_t st) {
  rnnt_alpha_kernel<<<N, 256, (size_t)2 * U * sizeof(double), st>>>(lp_blank, lp_label, xl, yl, T, U, loss);
  return cudaGetLastError();
}

}  // namespace rnnt
