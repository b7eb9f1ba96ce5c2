// Micro-benchmark: how fast can the 8 CTAs of a cluster all-gather a [128 x 256] bf16 (hi | lo) activation (16 KB slice per CTA,
// 128 KB landing in every CTA) by pushing 16-byte pieces into the peers' shared memory with st.async (complete_tx on the
// receiver's mbarrier)?  Compared in DESIGN.md 4.2 with the L2 round trip the persistent sampler uses (~7 k cycles per exchange).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/dsmem_probe tools/dsmem_probe.cu && tools/_bin/dsmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CL = 8, ROUNDS = 200;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }

template <int MODE>   // 0: st.async 16 B, 1: st.shared::cluster.v4 + one remote arrive per warp
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];          // 2 x 128 KB receive buffers would not fit: 1 x 128 KB, rounds alternate barriers
  __shared__ __align__(8) uint64_t bar[2];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(MODE == 0 ? 1 : 4 * CL));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(MODE == 0 ? 1 : 4 * CL));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int row = threadIdx.x;
  const uint32_t base = smem_u32(smem);
  long long t0 = clock64();
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t b = smem_u32(&bar[r & 1]);
    if (MODE == 0 && threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(131072) : "memory");
    // this CTA's slice: K-block rank / 2, 16-byte chunks (rank & 1) * 4 .. + 3 of the row, hi and lo images (swizzled positions)
    const uint32_t line = base + (rank >> 1) * 32768 + row * 128;
#pragma unroll
    for (int p = 0; p < CL; ++p) {
      const uint32_t rb = mapa(b, p);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t c = ((((rank & 1) * 4 + j) ^ (row & 7)) << 4);
        const uint32_t a0 = mapa(line + c, p), a1 = mapa(line + 16384 + c, p);
        if (MODE == 0) {
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%2, %2, %2, %2}, [%1];" ::"r"(a0), "r"(rb), "r"(r) : "memory");
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%2, %2, %2, %2}, [%1];" ::"r"(a1), "r"(rb), "r"(r) : "memory");
        } else {
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a0), "r"(r) : "memory");
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a1), "r"(r) : "memory");
        }
      }
    }
    if (MODE == 1) {
      __syncwarp();
      if ((threadIdx.x & 31) == 0) {
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(b, 0)) : "memory");
        for (int p = 1; p < CL; ++p) asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(b, p)) : "memory");
      }
    }
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"((r >> 1) & 1) : "memory");
    // a round may not start before every peer has finished reading the previous one in the real kernel; here the buffers are only
    // written, so the rounds run back to back: this measures the push itself
  }
  long long t1 = clock64();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / ROUNDS;
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * 8);
  for (int mode = 0; mode < 2; ++mode)
    for (int clusters : {1, 4}) {
      if (mode == 0) cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
      else cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) probe<0><<<clusters * CL, 128, 131072>>>(d); else probe<1><<<clusters * CL, 128, 131072>>>(d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      }
      long long h[64];
      cudaMemcpy(h, d, clusters * CL * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < clusters * CL; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%s, %d cluster(s) of 8: %lld cycles per all-gather round (128 KB into every CTA, 128 KB out of every CTA) = %.1f B/clk per SM inbound\n",
             mode == 0 ? "st.async 16 B + complete_tx" : "st.shared::cluster.v4 + release/relaxed arrives", clusters, mx, 131072.0 / mx);
    }
  return 0;
}
