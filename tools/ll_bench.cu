// Latency microbenchmarks behind the tagged-chunk exchange and the DSMEM partial tiles (tuning aid, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ll_bench tools/ll_bench.cu && tools/ll_bench
//   T1  round trip of a dependent ld.relaxed.gpu.v4 chain on an L2-resident buffer
//   T2  ping-pong through global memory between CTA 0 and CTA k (st.relaxed.gpu / ld.relaxed.gpu polling), idle chip
//   T3  the same ping-pong while 126 other CTAs (4 warps each) poll lines of the same buffer
//   T4  ping-pong through distributed shared memory inside a 2-CTA cluster: st.async + mbarrier complete_tx
//   T5  ... with st.shared::cluster + mbarrier.arrive.release.cluster (remote) instead
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// T1: buf[i] holds the byte offset of the next element (stride 4 KB): one lane chases
__global__ void chase_kernel(const uint8_t* buf, int n, unsigned long long* out) {
  if (threadIdx.x != 0) return;
  uint32_t off = 0;
  for (int i = 0; i < 64; ++i) off = ld_relaxed_v4(buf + off).x;   // warm
  const unsigned long long t0 = gtimer();
  const long long c0 = clock64();
  for (int i = 0; i < n; ++i) off = ld_relaxed_v4(buf + off).x;
  const long long c1 = clock64();
  const unsigned long long t1 = gtimer();
  out[0] = t1 - t0; out[1] = (unsigned long long)(c1 - c0); out[2] = off;
}

// T2 / T3: CTA 0 <-> CTA `peer`; the others either exit (noise = 0) or poll (noise = 1)
__global__ void pingpong_kernel(uint8_t* flags, uint8_t* noise_buf, int peer, int n, int noise, unsigned long long* out, volatile int* stop) {
  const int cta = blockIdx.x;
  uint8_t* fa = flags, *fb = flags + 4096;
  if (cta == 0) {
    if (threadIdx.x == 0) {
      const unsigned long long t0 = gtimer();
      for (int i = 1; i <= n; ++i) {
        st_relaxed_v4(fa, i, i, i, i);
        while (ld_relaxed_v4(fb).w != (uint32_t)i) {
        }
      }
      out[0] = gtimer() - t0;
      *stop = 1;
      __threadfence();
    }
  } else if (cta == peer) {
    if (threadIdx.x == 0) {
      for (int i = 1; i <= n; ++i) {
        while (ld_relaxed_v4(fa).w != (uint32_t)i) {
        }
        st_relaxed_v4(fb, i, i, i, i);
      }
    }
  } else if (noise) {
    // 4 warps polling 512 B each, like the loader warps of the persistent kernels
    const uint8_t* src = noise_buf + (size_t)(cta % 32) * 32768 + (size_t)threadIdx.x * 16;
    uint32_t acc = 0;
    while (!*stop) acc += ld_relaxed_v4(src).w;
    if (acc == 0x12345678u) out[7] = acc;
  }
}

// T4 / T5: 2-CTA cluster ping-pong through DSMEM
__global__ void __cluster_dims__(2, 1, 1) dsm_pingpong_kernel(int n, int mode, unsigned long long* out) {
  __shared__ __align__(16) float data[32];
  __shared__ __align__(8) unsigned long long bar;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t other = rank ^ 1u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  uint32_t rdata, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdata) : "r"(smem_u32(data)), "r"(other));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&bar)), "r"(other));
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned long long t0 = 0;
    if (rank == 0 && lane == 0) t0 = gtimer();
    for (int i = 0; i < n; ++i) {
      const uint32_t parity = i & 1;
      auto send = [&]() {
        if (mode == 0) {   // st.async, bytes counted on the receiver's barrier (armed by the receiver)
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(rdata + lane * 4), "f"((float)i), "r"(rbar) : "memory");
        } else {
          asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(rdata + lane * 4), "f"((float)i) : "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
        }
      };
      auto recv = [&]() {
        if (mode == 0 && lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128) : "memory");
        uint32_t ok = 0;
        while (!ok) {
          if (mode == 0)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
        }
        __syncwarp();
      };
      if (rank == 0) { send(); recv(); } else { recv(); send(); }
    }
    if (rank == 0 && lane == 0) out[0] = gtimer() - t0;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  unsigned long long *out, hout[8];
  CK(cudaMalloc(&out, 64));
  uint8_t* buf;
  const size_t N = 1 << 22;   // 4 MB: L2 resident
  CK(cudaMalloc(&buf, N));
  {   // chase ring with 4 KB stride (different lines / slices)
    uint8_t* h = (uint8_t*)calloc(N, 1);
    const int cnt = (int)(N / 4096);
    for (int i = 0; i < cnt; ++i) *(uint32_t*)(h + (size_t)i * 4096) = (uint32_t)(((i * 37 + 11) % cnt) * 4096);
    CK(cudaMemcpy(buf, h, N, cudaMemcpyHostToDevice));
    free(h);
  }
  const int n = 2000;
  chase_kernel<<<1, 32>>>(buf, n, out);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hout, out, 64, cudaMemcpyDeviceToHost));
  printf("T1 dependent ld.relaxed.gpu.v4 (L2 hit): %.1f ns, %.0f cycles per load\n", (double)hout[0] / n, (double)hout[1] / n);
  uint8_t* flags;
  int* stop;
  CK(cudaMalloc(&flags, 8192));
  CK(cudaMalloc(&stop, 4));
  for (int noise = 0; noise < 2; ++noise) {
    for (int peer : {1, 37, 74, 127}) {
      CK(cudaMemset(flags, 0, 8192));
      CK(cudaMemset(stop, 0, 4));
      void* args[] = {&flags, &buf, (void*)&peer, (void*)&n, (void*)&noise, &out, &stop};
      CK(cudaLaunchCooperativeKernel((void*)pingpong_kernel, dim3(128), dim3(128), args, 0, 0));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hout, out, 64, cudaMemcpyDeviceToHost));
      printf("T%d global ping-pong CTA0<->CTA%d%s: %.1f ns per round trip (one way = store visible + detected: %.1f ns)\n", noise ? 3 : 2, peer,
             noise ? " with 126 polling CTAs" : "", (double)hout[0] / n, (double)hout[0] / n / 2);
    }
  }
  for (int mode = 0; mode < 2; ++mode) {
    dsm_pingpong_kernel<<<2, 64>>>(n, mode, out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hout, out, 64, cudaMemcpyDeviceToHost));
    printf("T%d DSMEM ping-pong (%s): %.1f ns per round trip (one way %.1f ns)\n", 4 + mode, mode == 0 ? "st.async + complete_tx" : "st.shared::cluster + arrive.release.cluster",
           (double)hout[0] / n, (double)hout[0] / n / 2);
  }
  return 0;
}
