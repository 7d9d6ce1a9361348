// Small per-tick kernels of the streaming session (rnnt_b200_stream_*): the audio windows, the pending feature rows and
// all recurrent state of every stream stay in HBM; these two kernels move one tick's worth of data for all streams.
#include "kernels.h"

namespace rnnt {
namespace {

// window slide of the serving loop (reference api-server.py:95-102: frames.append(chunk); cat; frames.pop(0)):
// new = [old[chunk:], chunk] for the streams whose bit is set in `active`, new = old for the others
__global__ void __launch_bounds__(256) slide_window_kernel(const float* __restrict__ w_old, float* __restrict__ w_new,
                                                           const float* __restrict__ chunks, int W, int ck, StreamMask active) {
  const int b = blockIdx.y;
  const bool on = (active.bits[b >> 5] >> (b & 31)) & 1u;
  const float* src = w_old + (size_t)b * W;
  float* dst = w_new + (size_t)b * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W; i += gridDim.x * blockDim.x) {
    float v;
    if (!on) v = src[i];
    else v = (i < W - ck) ? src[i + ck] : chunks[(size_t)b * ck + (i - (W - ck))];
    dst[i] = v;
  }
}

// Buffer (transforms.py:455-471): append this tick's stacked row of stream b at position pos[b] (< 0: not appended)
__global__ void __launch_bounds__(256) store_rows_kernel(const float* __restrict__ row, float* __restrict__ rows,
                                                         const int32_t* __restrict__ pos, int X, int n_buffer) {
  const int b = blockIdx.y;
  const int p = pos[b];
  if (p < 0) return;
  const float* src = row + (size_t)b * X;
  float* dst = rows + ((size_t)b * n_buffer + p) * X;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < X; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace

cudaError_t launch_slide_window(const float* w_old, float* w_new, const float* chunks, int B, int W, int ck, const StreamMask& active,
                                cudaStream_t st) {
  dim3 grid((unsigned)((W + 255) / 256 < 8 ? (W + 255) / 256 : 8), (unsigned)B);
  slide_window_kernel<<<grid, 256, 0, st>>>(w_old, w_new, chunks, W, ck, active);
  return cudaGetLastError();
}

cudaError_t launch_store_rows(const float* row, float* rows, const int32_t* pos, int B, int X, int n_buffer, cudaStream_t st) {
  dim3 grid((unsigned)((X + 255) / 256 < 4 ? (X + 255) / 256 : 4), (unsigned)B);
  store_rows_kernel<<<grid, 256, 0, st>>>(row, rows, pos, X, n_buffer);
  return cudaGetLastError();
}

}  // namespace rnnt
