// Kernel-dispatch-rate probe for the sampler design: S CUDA graphs, each a dependent chain of N small kernels
// (grid G x 192 threads, each CTA spinning ~D ns), replayed concurrently on S streams.  Prints kernels/s per (S, G, D):
// if the rate saturates near what 8 concurrent 50-step sampler chains reach (~740 k kernels/s) independent of the kernels'
// own work, the pipelined sampler cost is bound by kernel dispatch, and only fewer kernels per step can lower it.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/launch_rate tools/launch_rate.cu && tools/_bin/launch_rate
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

__global__ void spin_kernel(long long ns, int* sink) {
  if (ns > 0) {
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    long long t = t0;
    while (t - t0 < ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t == 0) *sink = 1;
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main() {
  const int N = 3750;
  int* sink;
  CK(cudaMalloc(&sink, 4));
  const int Ss[] = {1, 2, 4, 8, 16};
  const int Gs[] = {16, 64};
  const long long Ds[] = {0, 2000, 4000};
  printf("%8s %6s %8s %12s %14s\n", "streams", "grid", "spin_ns", "ms/chain", "kernels/s");
  for (int G : Gs) for (long long D : Ds) for (int S : Ss) {
    std::vector<cudaStream_t> st(S);
    std::vector<cudaGraphExec_t> ex(S);
    for (int i = 0; i < S; ++i) {
      CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
      cudaGraph_t g;
      CK(cudaStreamBeginCapture(st[i], cudaStreamCaptureModeThreadLocal));
      for (int k = 0; k < N; ++k) spin_kernel<<<G, 192, 0, st[i]>>>(D, sink);
      CK(cudaStreamEndCapture(st[i], &g));
      CK(cudaGraphInstantiate(&ex[i], g, 0));
      CK(cudaGraphDestroy(g));
    }
    for (int i = 0; i < S; ++i) CK(cudaGraphLaunch(ex[i], st[i]));
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int reps = 3;
    CK(cudaEventRecord(a, st[0]));
    for (int r = 0; r < reps; ++r) for (int i = 0; i < S; ++i) CK(cudaGraphLaunch(ex[i], st[i]));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(b, st[0]));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    printf("%8d %6d %8lld %12.2f %14.0f\n", S, G, D, ms / reps, (double)N * S * reps / (ms / 1e3));
    for (int i = 0; i < S; ++i) { cudaGraphExecDestroy(ex[i]); cudaStreamDestroy(st[i]); }
    cudaEventDestroy(a); cudaEventDestroy(b);
  }
  return 0;
}
