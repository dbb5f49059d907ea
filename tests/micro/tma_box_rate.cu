// micro-benchmark: how fast can ONE thread per SM pull [rows x 64 ch] bf16 boxes (128B swizzle) of a [P, C] matrix
// through TMA when nobody does anything with the data?  (the input path of pw_fwd_tc / pw_wgrad_tc)  Also: the same
// bytes as 1-D bulk copies (cp.async.bulk) of the contiguous tile.   nvcc -arch=sm_100a -lcuda -o tma_box_rate ...
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define ACCX_NO_LIB
#include "../../acc-unet-unext_b200/csrc/tc_common.cuh"
using namespace accx;

struct alignas(64) P {
  CUtensorMap map;
  const __nv_bfloat16* x;
  int64_t n_rows;
  int C, box_rows, stages, tiles, mode, boxes_per_tile;
};

__global__ void __launch_bounds__(64, 1) pull(const __grid_constant__ P p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)p.box_rows * 128;
  const uint32_t bar0 = base + p.stages * stage_bytes;      // full[S], empty[S]
  const int S = p.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * (S + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_it = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x * p.boxes_per_tile;
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
      for (int b = 0; b < p.boxes_per_tile; ++b, ++it) {
        mbar_wait(bar0 + 8 * (S + stage), ph ^ 1);
        const uint32_t bar = bar0 + 8 * stage;
        const uint32_t bytes = p.mode == 0 ? stage_bytes : (uint32_t)p.box_rows * p.C * 2;
        mbar_expect_tx(bar, bytes);
        if (p.mode == 0) tma_load_2d(base + stage * stage_bytes, &p.map, b * 64, tile * p.box_rows, bar);
        else bulk_g2s(base + stage * stage_bytes, p.x + (int64_t)tile * p.box_rows * p.C, bytes, bar);
        if (++stage == S) { stage = 0; ph ^= 1; }
      }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t ph = 0;
    for (int it = 0; it < n_it; ++it) {
      mbar_wait(bar0 + 8 * stage, ph);
      mbar_arrive(bar0 + 8 * (S + stage));
      if (++stage == S) { stage = 0; ph ^= 1; }
    }
  }
}

int main() {
  const int64_t rows = 16 * 224 * 224;
  __nv_bfloat16* x;
  cudaMalloc(&x, rows * 256 * 2);
  cudaMemset(x, 0, rows * 256 * 2);
  cudaFuncSetAttribute(pull, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int Cs[] = {32, 64, 128, 256};
  for (int mode = 0; mode < 2; ++mode)
    for (int C : Cs)
      for (int box_rows : {64, 128, 256})
        for (int S : {2, 4, 8}) {
          if (mode == 1 && (box_rows * C * 2 > box_rows * 128)) continue;   // bulk copy of the tile must fit the stage
          P p;
          p.x = x; p.n_rows = rows; p.C = C; p.box_rows = box_rows; p.stages = S; p.mode = mode;
          p.tiles = (int)(rows / box_rows);
          p.boxes_per_tile = mode == 0 ? (C + 63) / 64 : 1;
          if (mode == 0 && !encode_2d_bf16(&p.map, x, C, rows, C, box_rows)) { printf("encode failed\n"); return 1; }
          const size_t smem = 1024 + (size_t)S * box_rows * 128 + 16 * S + 64;
          if (smem > 200 * 1024) continue;
          for (int w = 0; w < 2; ++w) pull<<<148, 64, smem>>>(p);
          cudaEventRecord(e0);
          for (int i = 0; i < 5; ++i) pull<<<148, 64, smem>>>(p);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          ms /= 5;
          const double bytes = (double)rows * C * 2;
          const double boxes = (double)p.tiles * p.boxes_per_tile;
          printf("%s C=%3d box_rows=%3d stages=%d : %7.1f us  %6.0f GB/s  %5.0f ns per box per SM, %5.2f ns per row\n",
                 mode == 0 ? "tensor-2d" : "bulk-1d  ", C, box_rows, S, ms * 1e3, bytes / ms / 1e6, ms * 1e6 / (boxes / 148),
                 ms * 1e6 / (boxes / 148) / box_rows);
        }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
