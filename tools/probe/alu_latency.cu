// Dependent-issue latency of the integer ops on the rANS consumer's critical path (one warp, clock64 around
// an unrolled dependent chain). nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_latency alu_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 4096;
template <int OP>
__global__ void chain(uint32_t seed, uint32_t m, uint32_t sh, uint32_t thr, long long* cycles, uint32_t* sink) {
  uint32_t x = seed + threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = x * m + sh;                                  // IMAD
    if (OP == 1) x = __umulhi(x, m) + sh;                         // IMAD.HI (+c folded)
    if (OP == 2) x = (x >> (sh & 31)) | 0x40000000u;              // SHF by a register (+LOP)
    if (OP == 3) x = (x >= thr ? 24u : 8u) + x;                   // ISETP -> SEL -> IADD
    if (OP == 4) {                                                // the step as in rans_encode_range (rows in registers)
      const uint32_t hi = __umulhi(x, m) + 0u;
      const uint32_t q0 = hi >> sh;
      const bool p1 = x >= thr, p2 = x >= (thr << 8), p3 = x >= (thr << 16);
      const uint32_t k8 = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
      x = (q0 >> k8) * 1000003u + ((x >> k8) + 12345u);
      x = (x & 0x3FFFFFFFu) | 0x00400000u;                        // keep the state in range for the probe
    }
    if (OP == 5) {                                                // same, k8 by arithmetic on the compare results
      const uint32_t hi = __umulhi(x, m);
      const uint32_t q0 = hi >> sh;
      const uint32_t k8 = ((x >= thr) + (x >= (thr << 8)) + (x >= (thr << 16))) * 8u;
      x = (q0 >> k8) * 1000003u + ((x >> k8) + 12345u);
      x = (x & 0x3FFFFFFFu) | 0x00400000u;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  sink[threadIdx.x] = x;
}

template <int OP>
void run(const char* name) {
  long long* d_c; uint32_t* d_s;
  cudaMalloc(&d_c, 8); cudaMalloc(&d_s, 128);
  for (int r = 0; r < 2; ++r) chain<OP><<<1, 32>>>(12345u, 0x9E3779B1u, 7u, 1u << 12, d_c, d_s);
  long long c = 0;
  cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %6.2f cycles per iteration\n", name, (double)c / N);
  cudaFree(d_c); cudaFree(d_s);
}

int main() {
  run<0>("IMAD  x = x*m + c");
  run<1>("IMAD.HI  x = umulhi(x, m) + c");
  run<2>("SHF  x = (x >> r) | k");
  run<3>("ISETP+SEL+IADD");
  run<4>("rANS step, rows in registers");
  run<5>("rANS step, arithmetic k8");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
