// Which scheduler (SM sub-partition) does warp w of a CTA run on? Warps selected by a mask run an ALU-pipe-bound loop;
// two of them on one scheduler take twice as long. Also prints %warpid of every CTA warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smsp_map smsp_map.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t mask, long long* cycles, uint32_t* hw, uint32_t* sink) {
  const uint32_t warp = threadIdx.x >> 5;
  uint32_t wid, smid;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if ((threadIdx.x & 31) == 0) hw[blockIdx.x * 32 + warp] = wid | (smid << 16);
  __syncthreads();
  if (!((mask >> warp) & 1u)) return;
  uint32_t a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3, e = a + 4, f = a + 5, g = a + 6, h = a + 7;
  const uint32_t s = (mask & 7u) + 1u;
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 20000; ++i) {  // eight independent funnel-shift chains: ALU pipe bound
    a = __funnelshift_r(a, b, s); b = __funnelshift_r(b, c, s); c = __funnelshift_r(c, d, s); d = __funnelshift_r(d, e, s);
    e = __funnelshift_r(e, f, s); f = __funnelshift_r(f, g, s); g = __funnelshift_r(g, h, s); h = __funnelshift_r(h, a, s);
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 32 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

int main() {
  long long* d_c; uint32_t *d_hw, *d_s;
  cudaMalloc(&d_c, 64 * 8); cudaMalloc(&d_hw, 64 * 4); cudaMalloc(&d_s, 2 * 512 * 4);
  const uint32_t masks[] = {0x1, 0x3, 0x11, 0x5, 0xF, 0x33, 0x55, 0x1111, 0xFF};
  for (uint32_t m : masks) {
    cudaMemset(d_c, 0, 64 * 8);
    probe<<<1, 512>>>(m, d_c, d_hw, d_s);
    long long c[64]; uint32_t hw[64];
    cudaMemcpy(c, d_c, 64 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hw, d_hw, 64 * 4, cudaMemcpyDeviceToHost);
    printf("mask %04x:", m);
    for (int w = 0; w < 16; ++w) if ((m >> w) & 1) printf("  w%d(hw %u) %lld", w, hw[w] & 0xFFFF, c[w]);
    printf("\n");
  }
  // two 128-thread CTAs: do they share an SM, and where do their warps sit?
  probe<<<296, 128>>>(0x1, d_c, d_hw, d_s);
  cudaDeviceSynchronize();
  uint32_t hw[64];
  cudaMemcpy(hw, d_hw, 64 * 4, cudaMemcpyDeviceToHost);
  printf("CTA 0: sm %u warps hw", hw[0] >> 16); for (int w = 0; w < 4; ++w) printf(" %u", hw[w] & 0xFFFF);
  printf("\nCTA 1: sm %u warps hw", hw[32] >> 16); for (int w = 0; w < 4; ++w) printf(" %u", hw[32 + w] & 0xFFFF);
  printf("\n");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
