// Probe: is an external event-record node of a CUDA graph "pending" for cudaEventQuery / cudaEventSynchronize as soon as
// cudaGraphLaunch returns (like cudaEventRecord on a stream), also on the second launch when the event completed before?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long cycles, int* out) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} *out = 1; }
int main() {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  int* d; cudaMalloc(&d, 4);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  spin<<<1, 1, 0, s>>>(200000000LL, d);  // ~0.1 s
  cudaEventRecordWithFlags(ev, s, cudaEventRecordExternal);
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int it = 0; it < 3; ++it) {
    cudaGraphLaunch(ge, s);
    cudaError_t q = cudaEventQuery(ev);
    printf("launch %d: query right after launch -> %s\n", it, q == cudaErrorNotReady ? "not ready (pending)" : cudaGetErrorName(q));
    cudaEventSynchronize(ev);
    printf("launch %d: after synchronize stream query -> %s\n", it, cudaGetErrorName(cudaStreamQuery(s)));
    cudaStreamSynchronize(s);
  }
  return 0;
}
