// micro-benchmark: tail cost of per-channel fp32 atomics issued by every block at the end of a kernel
// (the BatchNorm-statistics flush of the reducing accx kernels), plain vs sharded over S accumulators.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void flush(float* out, int C, int shards, int reps) {
  // every thread < C adds one value per statistic (2 statistics), like reduce_lanes_atomic's last stage
  float* dst = out + (size_t)(blockIdx.x % shards) * 2 * C;
  for (int r = 0; r < reps; ++r)
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) atomicAdd(dst + c, 1.0f);
}
__global__ void empty_k(float* out) { if (out == nullptr) printf("x"); }
int main() {
  float* d;
  cudaMalloc(&d, 64 << 20);
  cudaMemset(d, 0, 64 << 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int Cs[] = {32, 96, 256, 4352};
  const int Bs[] = {148, 296, 592, 1184, 4736};
  const int Ss[] = {1, 4, 16, 64};
  for (int C : Cs) for (int blocks : Bs) for (int S : Ss) {
    for (int w = 0; w < 3; ++w) flush<<<blocks, 256>>>(d, C, S, 1);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) flush<<<blocks, 256>>>(d, C, S, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("C=%5d blocks=%5d shards=%3d : %7.2f us per launch\n", C, blocks, S, ms / 20 * 1e3);
  }
  cudaEventRecord(e0);
  for (int i = 0; i < 20; ++i) empty_k<<<296, 256>>>(d);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("empty kernel: %.2f us\n", ms / 20 * 1e3);
  return 0;
}
