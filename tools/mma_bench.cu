// Micro-benchmark: cycles per tcgen05.mma (kind::f16, SS) for small tiles -- tuning aid, not product code.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I libreasr_b200/csrc -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace rnnt;

// issue `8*reps` MMAs with compile-time-constant descriptor offsets; NACC accumulators 64 columns apart
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int M, int N, int NACC, int ADV, int NW>
__global__ void bench(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(&bar, NW); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tptr;
  const int warp = threadIdx.x >> 5;
  if (warp < NW) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
    const uint32_t id = umma_idesc_f16(M, N);
    const uint64_t ad = umma_desc_sw128(a0), bd = umma_desc_sw128(b0);
    const uint32_t tw = tmem + warp * (NACC * 64 > 128 ? 128 : NACC * 64);
    for (int rep = 0; rep < 2; ++rep) {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint64_t off = ADV ? (uint64_t)(((i & 3) * 32 + (i >> 2) * 16384) >> 4) : 0;
          if (elect_one()) tc_mma_f16(tw + (i % NACC) * 64, ad + off, bd + off, id, 1u);
        }
      }
      long long t1 = clock64();
      if (elect_one()) tc_commit(&bar);
      mbar_wait(&bar, rep & 1);
      long long t2 = clock64();
      if (rep == 1 && threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int M, int N, int NACC, int ADV, int NW = 1>
void run(int grid, long long* d) {
  const int reps = 16;
  cudaFuncSetAttribute(bench<M, N, NACC, ADV, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  bench<M, N, NACC, ADV, NW><<<grid, 128, 200 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("grid %3d M %3d N %3d nacc %d adv %d warps %d : issue %6.1f cyc/mma(per warp)  complete %6.1f cyc per mma overall  (%s)\n", grid, M, N, NACC, ADV, NW,
         (double)h[0] / (8 * reps), (double)h[1] / (8 * reps * NW), cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16);
  for (int grid : {1}) {
    run<128, 256, 1, 1>(grid, d); run<128, 128, 1, 1>(grid, d); run<128, 64, 1, 1>(grid, d);
    run<128, 32, 1, 1>(grid, d); run<128, 32, 2, 1>(grid, d); run<128, 16, 2, 1>(grid, d);
    run<64, 128, 1, 1>(grid, d); run<64, 64, 1, 1>(grid, d); run<64, 64, 2, 1>(grid, d); run<64, 32, 1, 1>(grid, d);
    run<64, 32, 2, 1>(grid, d); run<64, 16, 2, 1>(grid, d); run<64, 8, 2, 1>(grid, d); run<64, 96, 1, 1>(grid, d);
    run<64, 32, 2, 1, 2>(grid, d); run<64, 32, 2, 1, 4>(grid, d); run<64, 64, 2, 1, 2>(grid, d); run<64, 64, 1, 1, 4>(grid, d);
    run<128, 32, 2, 1, 2>(grid, d); run<128, 32, 1, 1, 4>(grid, d); run<128, 256, 1, 1, 2>(grid, d);
  }
  return 0;
}
