// UNeXt shifted tokenized-MLP block (/root/reference/Experiments/nets/UNext.py:38-160):
//   shiftedBlock(x) = x + shiftmlp(LayerNorm(x));  shiftmlp = shift_H -> fc1 -> DWConv(3x3, bias) -> GELU -> shift_W -> fc2
// Tokens [B, N = H*W, C] ARE an NHWC tensor, so the shifts (pad / chunk(5) / roll / narrow, :78-84,:97-103) are
// shifted-operand index arithmetic inside the fc1 / fc2 contractions (accx_pw_fwd: five operands with (dy, dx) offsets,
// no data movement), DWConv is accx_dw3x3_fwd, and what is left are the two element-wise pieces in this file:
//   accx_layernorm_fwd / _bwd   nn.LayerNorm(dim) of shiftedBlock.norm2 (:151, :156)  -- one warp per token
//   accx_gelu_fwd / _bwd        nn.GELU (exact erf form, :47,:90)
// Memory-bound: each tensor once.
#include "common.cuh"

namespace accx {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// y = (x - mean) * rstd * gamma + beta over the C channels of every row; mean / rstd [R] saved for the backward
template <typename T>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(int64_t R, int C, const T* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, T* __restrict__ y, float* __restrict__ mean,
                                                            float* __restrict__ rstd) {
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * nw + warp; r < R; r += (int64_t)gridDim.x * nw) {
    const T* xr = x + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += to_f(xr[c]);
    const float mu = warp_sum(s) / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = to_f(xr[c]) - mu; q = fmaf(d, d, q); }
    const float rs = rsqrtf(warp_sum(q) / C + eps);
    T* yr = y + r * C;
    for (int c = lane; c < C; c += 32) yr[c] = from_f<T>((to_f(xr[c]) - mu) * rs * gamma[c] + beta[c]);
    if (lane == 0) { if (mean) mean[r] = mu; if (rstd) rstd[r] = rs; }
  }
}

// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)), g = dy * gamma;  dgamma += sum_r dy * xhat, dbeta += sum_r dy.
// A lane owns channels lane, lane + 32, ..: its per-channel sums stay in registers over all rows of the warp, are
// added across the warps of the block in warp order through shared memory, then one atomicAdd per channel and block
// (deterministic mode: one block).
template <typename T, int NJ>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(int64_t R, int C, const T* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const T* __restrict__ dy,
                                                            T* __restrict__ dx, float* dgamma, float* dbeta) {
  pdl_sync();
  extern __shared__ float red[];        // [nw][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float gm[NJ], ag[NJ], ab[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = lane + 32 * j;
    gm[j] = c < C ? gamma[c] : 0.f;
    ag[j] = ab[j] = 0.f;
  }
  for (int64_t r = (int64_t)blockIdx.x * nw + warp; r < R; r += (int64_t)gridDim.x * nw) {
    const float mu = mean[r], rs = rstd[r];
    float xh[NJ], g[NJ];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      const float d = c < C ? to_f(dy[r * C + c]) : 0.f;
      xh[j] = c < C ? (to_f(x[r * C + c]) - mu) * rs : 0.f;
      g[j] = d * gm[j];
      s1 += g[j];
      s2 = fmaf(g[j], xh[j], s2);
      ag[j] = fmaf(d, xh[j], ag[j]);
      ab[j] += d;
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      if (c < C) dx[r * C + c] = from_f<T>(rs * (g[j] - s1 - xh[j] * s2));
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = lane + 32 * j;
    if (c < C) { red[(warp * 2) * C + c] = ag[j]; red[(warp * 2 + 1) * C + c] = ab[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < nw; ++w) { a += red[(w * 2) * C + c]; b += red[(w * 2 + 1) * C + c]; }
    if (dgamma) atomicAdd(dgamma + c, a);
    if (dbeta) atomicAdd(dbeta + c, b);
  }
}

__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float v) {
  return 0.5f * (1.f + erff(v * 0.70710678118654752f)) + v * 0.3989422804014327f * __expf(-0.5f * v * v);
}

// y = gelu(x)  (dy == nullptr)   or   y = dy * gelu'(x)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) gelu_kernel(int64_t n_vec, const T* __restrict__ x, const T* __restrict__ dy,
                                                   T* __restrict__ y) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[VEC], d[VEC];
    ldv<T, VEC>(x + i * VEC, v);
    if (dy) {
      ldv<T, VEC>(dy + i * VEC, d);
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] = d[e] * gelu_grad_f(v[e]);
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] = gelu_f(v[e]);
    }
    stv<T, VEC>(y + i * VEC, v);
  }
}

template <typename T>
static int launch_gelu(int64_t n, const void* x, const void* dy, void* y, cudaStream_t st) {
  const bool vec = n % DT<T>::VEC == 0 && aligned16(x) && aligned16(y) && (!dy || aligned16(dy));
  if (vec) {
    const int64_t nv = n / DT<T>::VEC;
    launch_k(gelu_kernel<T, DT<T>::VEC>, grid_x_for(nv, 256, 148 * 8), 256, 0, st, nv, (const T*)x, (const T*)dy, (T*)y);
  } else {
    launch_k(gelu_kernel<T, 1>, grid_x_for(n, 256, 148 * 8), 256, 0, st, n, (const T*)x, (const T*)dy, (T*)y);
  }
  return check_launch("gelu");
}

}  // namespace accx

using namespace accx;

extern "C" {

int accx_layernorm_fwd(int dtype, int64_t R, int C, const void* x, const float* gamma, const float* beta, float eps,
                       void* y, float* mean, float* rstd, void* stream) {
  ACCX_REQUIRE(R > 0 && C > 0 && x && gamma && beta && y, "layernorm_fwd: bad arguments");
  const int grid = grid_x_for(R, 8, 148 * 8);
  ACCX_DISPATCH_T(dtype, {
    launch_k(layernorm_fwd_kernel<T>, grid, 256, 0, (cudaStream_t)stream, R, C, (const T*)x, gamma, beta, eps, (T*)y, mean, rstd);
  });
  return check_launch("layernorm_fwd");
}

int accx_layernorm_bwd(int dtype, int64_t R, int C, const void* x, const float* gamma, const float* mean,
                       const float* rstd, const void* dy, void* dx, float* dgamma, float* dbeta, void* stream) {
  ACCX_REQUIRE(R > 0 && C > 0 && C <= 1024 && x && gamma && mean && rstd && dy && dx,
               "layernorm_bwd: bad arguments (C = %d must be <= 1024)", C);
  int grid = grid_x_for(R, 8 * 8, 148 * 2);
  if (det_on()) grid = 1;               // deterministic mode: one contribution per parameter gradient
  const size_t sm = (size_t)8 * 2 * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  ACCX_DISPATCH_T(dtype, {
    if (C <= 128)
      launch_k(layernorm_bwd_kernel<T, 4>, grid, 256, sm, st, R, C, (const T*)x, gamma, mean, rstd, (const T*)dy, (T*)dx, dgamma, dbeta);
    else if (C <= 256)
      launch_k(layernorm_bwd_kernel<T, 8>, grid, 256, sm, st, R, C, (const T*)x, gamma, mean, rstd, (const T*)dy, (T*)dx, dgamma, dbeta);
    else if (C <= 512)
      launch_k(layernorm_bwd_kernel<T, 16>, grid, 256, sm, st, R, C, (const T*)x, gamma, mean, rstd, (const T*)dy, (T*)dx, dgamma, dbeta);
    else {
      cudaFuncSetAttribute(layernorm_bwd_kernel<T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      launch_k(layernorm_bwd_kernel<T, 32>, grid, 256, sm, st, R, C, (const T*)x, gamma, mean, rstd, (const T*)dy, (T*)dx, dgamma, dbeta);
    }
  });
  return check_launch("layernorm_bwd");
}

int accx_gelu_fwd(int dtype, int64_t n, const void* x, void* y, void* stream) {
  ACCX_REQUIRE(n > 0 && x && y, "gelu_fwd: bad arguments");
  ACCX_DISPATCH_T(dtype, { return launch_gelu<T>(n, x, nullptr, y, (cudaStream_t)stream); });
  return ACCX_OK;
}

int accx_gelu_bwd(int dtype, int64_t n, const void* x, const void* dy, void* dx, void* stream) {
  ACCX_REQUIRE(n > 0 && x && dy && dx, "gelu_bwd: bad arguments");
  ACCX_DISPATCH_T(dtype, { return launch_gelu<T>(n, x, dy, dx, (cudaStream_t)stream); });
  return ACCX_OK;
}

}  // extern "C"
