// Memory-bound channel-lane kernels: BatchNorm finalisation / backward, lazy materialisation,
// residual add, block-sum pooling, nearest upsample-add, layout changes.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace accx {

static thread_local char g_err[512] = "";
int g_knobs[KNOB_COUNT] = {0};
Det g_det = {nullptr, nullptr};
int64_t g_det_floats = 0;
int g_det_ctrs = 0;
static cudaStream_t g_det_streams[DET_SLOTS];
static int g_det_n_streams = 0;

int det_slot(cudaStream_t st) {
  for (int i = 0; i < g_det_n_streams; ++i)
    if (g_det_streams[i] == st) return i;
  if (g_det_n_streams >= DET_SLOTS) return -1;
  g_det_streams[g_det_n_streams] = st;
  return g_det_n_streams++;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return ACCX_ERR_CUDA;
  }
  return ACCX_OK;
}

// --------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(int C, double count, const float* __restrict__ stats,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ conv_bias, float eps,
                                   float momentum, int training, float* running_mean, float* running_var,
                                   int64_t* nbt, float* scale, float* shift, float* mean_o, float* rstd_o) {
  pdl_sync();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt) *nbt += 1;
  if (c >= C) return;
  const float cb = conv_bias ? conv_bias[c] : 0.f;   // bias of the producing conv, folded here
  float mean, var;
  if (training) {
    double m = (double)stats[c] / count;
    double v = (double)stats[C + c] / count - m * m;
    if (v < 0) v = 0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) {
      double unb = count > 1 ? v * count / (count - 1) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + cb);
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c] - cb;
    var = running_var[c];
  }
  float rstd = rsqrtf(var + eps);
  float s = gamma[c] * rstd;
  scale[c] = s;
  shift[c] = beta[c] - mean * s;
  if (mean_o) mean_o[c] = mean;
  if (rstd_o) rstd_o[c] = rstd;
}

// --------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 2) act_apply_kernel(int64_t P, int C, const T* __restrict__ x, const float* scale, const float* shift,
                                 int act, const float* scale2, const float* shift2, const T* __restrict__ residual,
                                 T* __restrict__ out, float* stats, Det det) {
  pdl_sync();
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  const int c0 = active ? cv * VEC : 0;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float s2[VEC], t2[VEC];
  if (scale2) { ldf<VEC>(scale2 + c0, s2); ldf<VEC>(shift2 + c0, t2); }
  float acc[2][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (active) {
    constexpr int U = 8;
    RawVec<T, VEC> rx[U], rr[U];
    pixel_loop<U>((int64_t)blockIdx.x * blockDim.y + threadIdx.y, P, (int64_t)gridDim.x * blockDim.y,
        [&](int u, int64_t p) {
          rx[u].load(x + p * C + c0);
          if (residual) rr[u].load(residual + p * C + c0);
        },
        [&](int u, int64_t p) {
          float v[VEC];
          rx[u].unpack(v);
          lz.apply(v);
          if (scale2) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[i] = fmaf(v[i], s2[i], t2[i]);
          }
          if (residual) {
            float r[VEC];
            rr[u].unpack(r);
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[i] += r[i];
          }
#pragma unroll
          for (int i = 0; i < VEC; ++i) { acc[0][i] += v[i]; acc[1][i] += v[i] * v[i]; }
          if (out) stv<T, VEC>(out + p * C + c0, v);
        });
  }
  if (stats) reduce_lanes_atomic<2, VEC>(acc, smem, stats, C, C, det, blockIdx.y, blockIdx.x, gridDim.x);
}

// z = act(a) + r
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 2) add_fwd_kernel(int64_t P, int C, const T* __restrict__ a, const float* scale, const float* shift,
                               int act, const T* __restrict__ r, T* __restrict__ z, float* stats, Det det) {
  pdl_sync();
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  const int c0 = active ? cv * VEC : 0;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float acc[2][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (active) {
    constexpr int U = 8;
    RawVec<T, VEC> ra[U], rr[U];
    pixel_loop<U>((int64_t)blockIdx.x * blockDim.y + threadIdx.y, P, (int64_t)gridDim.x * blockDim.y,
        [&](int u, int64_t p) {
          ra[u].load(a + p * C + c0);
          rr[u].load(r + p * C + c0);
        },
        [&](int u, int64_t p) {
          float v[VEC], w[VEC];
          ra[u].unpack(v);
          rr[u].unpack(w);
          lz.apply(v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            v[i] += w[i];
            acc[0][i] += v[i];
            acc[1][i] += v[i] * v[i];
          }
          stv<T, VEC>(z + p * C + c0, v);
        });
  }
  if (stats) reduce_lanes_atomic<2, VEC>(acc, smem, stats, C, C, det, blockIdx.y, blockIdx.x, gridDim.x);
}

// --------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void bn_bwd_reduce_kernel(int64_t P, int C, const T* __restrict__ y, const float* scale,
                                     const float* shift, int act, const float* mean, const float* rstd,
                                     const T* __restrict__ da, float* sums, Det det) {
  pdl_sync();
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  const int c0 = active ? cv * VEC : 0;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float mu[VEC], rs[VEC];
  ldf<VEC>(mean + c0, mu);
  ldf<VEC>(rstd + c0, rs);
  float acc[2][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (active) {
    constexpr int U = 8;
    RawVec<T, VEC> ry[U], rg[U];
    pixel_loop<U>((int64_t)blockIdx.x * blockDim.y + threadIdx.y, P, (int64_t)gridDim.x * blockDim.y,
        [&](int u, int64_t p) {
          ry[u].load(y + p * C + c0);
          rg[u].load(da + p * C + c0);
        },
        [&](int u, int64_t p) {
          float yv[VEC], g[VEC];
          ry[u].unpack(yv);
          rg[u].unpack(g);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            float gi = g[i] * lz.dact(yv[i], i);
            acc[0][i] += gi;
            acc[1][i] += gi * (yv[i] - mu[i]) * rs[i];
          }
        });
  }
  reduce_lanes_atomic<2, VEC>(acc, smem, sums, C, C, det, blockIdx.y, blockIdx.x, gridDim.x);
}

template <typename T, int VEC>
__global__ void bn_bwd_apply_kernel(int64_t P, int C, const T* __restrict__ y, const float* scale, const float* shift,
                                    int act, const float* mean, const float* rstd, const float* gamma,
                                    const T* __restrict__ da, const float* __restrict__ sums, float inv_count,
                                    T* __restrict__ dy, float* dgamma, float* dbeta) {
  pdl_sync();
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * VEC >= C) return;
  const int c0 = cv * VEC;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float mu[VEC], rs[VEC], gm[VEC], s1[VEC], s2[VEC];
  ldf<VEC>(mean + c0, mu);
  ldf<VEC>(rstd + c0, rs);
  ldf<VEC>(gamma + c0, gm);
  ldf<VEC>(sums + c0, s1);
  ldf<VEC>(sums + C + c0, s2);
  if (blockIdx.x == 0 && threadIdx.y == 0) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      if (dgamma) atomicAdd(dgamma + c0 + i, s2[i]);
      if (dbeta) atomicAdd(dbeta + c0 + i, s1[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) { s1[i] *= inv_count; s2[i] *= inv_count; gm[i] *= rs[i]; }
  constexpr int U = 4;
  RawVec<T, VEC> ry[U], rg[U];
  pixel_loop<U>((int64_t)blockIdx.x * blockDim.y + threadIdx.y, P, (int64_t)gridDim.x * blockDim.y,
      [&](int u, int64_t p) {
        ry[u].load(y + p * C + c0);
        rg[u].load(da + p * C + c0);
      },
      [&](int u, int64_t p) {
        float yv[VEC], g[VEC];
        ry[u].unpack(yv);
        rg[u].unpack(g);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float gi = g[i] * lz.dact(yv[i], i);
          float xh = (yv[i] - mu[i]) * rs[i];
          g[i] = gm[i] * (gi - s1[i] - xh * s2[i]);
        }
        stv<T, VEC>(dy + p * C + c0, g);
      });
}

// --------------------------------------------------------------------------------------
// out[b,ho,wo,c] = mul * sum_{s x s} x[b, ho*s+i, wo*s+j, c]
template <typename TI, typename TO, int VEC, int S>
__global__ void pool_sum_kernel(int B, int H, int W, int C, int log2s, float mul, const TI* __restrict__ x,
                                TO* __restrict__ out, int64_t out_ld) {
  pdl_sync();
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * VEC >= C) return;
  const int c0 = cv * VEC;
  const int s = S > 0 ? S : 1 << log2s, Ho = H >> log2s, Wo = W >> log2s;
  const int64_t Po = (int64_t)B * Ho * Wo;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; q < Po; q += (int64_t)gridDim.x * blockDim.y) {
    int wo = (int)(q % Wo);
    int64_t t = q / Wo;
    int ho = (int)(t % Ho);
    int b = (int)(t / Ho);
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    if constexpr (S == 8) {     // 8 x 8 window: one window row (8 loads) in flight at a time, two rows per step
#pragma unroll 1
      for (int i = 0; i < 8; i += 2) {
        RawVec<TI, VEC> rx[16];
        const TI* row = x + (((int64_t)b * H + (ho * 8 + i)) * W + (int64_t)wo * 8) * C + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          rx[j].load(row + (int64_t)j * C);
          rx[8 + j].load(row + ((int64_t)W + j) * C);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float v[VEC];
          rx[k].unpack(v);
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[e] += v[e];
        }
      }
    } else if constexpr (S > 0) {      // compile-time window: all S*S loads in flight before the first add
      RawVec<TI, VEC> rx[S * S];
#pragma unroll
      for (int i = 0; i < S; ++i) {
        const TI* row = x + (((int64_t)b * H + (ho * S + i)) * W + (int64_t)wo * S) * C + c0;
#pragma unroll
        for (int j = 0; j < S; ++j) rx[i * S + j].load(row + (int64_t)j * C);
      }
#pragma unroll
      for (int k = 0; k < S * S; ++k) {
        float v[VEC];
        rx[k].unpack(v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += v[e];
      }
    } else {
      for (int i = 0; i < s; ++i) {
        const TI* row = x + (((int64_t)b * H + (ho * s + i)) * W + (int64_t)wo * s) * C + c0;
        for (int j = 0; j < s; ++j) {
          float v[VEC];
          ldv<TI, VEC>(row + (int64_t)j * C, v);
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[e] += v[e];
        }
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] *= mul;
    stv<TO, VEC>(out + q * out_ld + c0, acc);
  }
}

// dst[b,h,w,c] (+)= mul * src[b, h>>l, w>>l, c]
template <typename TI, typename TO, int VEC>
__global__ void upsample_add_kernel(int B, int H, int W, int C, int log2s, float mul, const TI* __restrict__ src,
                                    int64_t src_ld, TO* __restrict__ dst, int accumulate) {
  pdl_sync();
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * VEC >= C) return;
  const int c0 = cv * VEC;
  const int Hs = H >> log2s, Ws = W >> log2s;
  const int64_t P = (int64_t)B * H * W;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p < P; p += (int64_t)gridDim.x * blockDim.y) {
    int w = (int)(p % W);
    int64_t t = p / W;
    int h = (int)(t % H);
    int b = (int)(t / H);
    int64_t q = ((int64_t)b * Hs + (h >> log2s)) * Ws + (w >> log2s);
    float v[VEC];
    ldv<TI, VEC>(src + q * src_ld + c0, v);
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] *= mul;
    if (accumulate) {
      float d[VEC];
      ldv<TO, VEC>(dst + p * C + c0, d);
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] += d[e];
    }
    stv<TO, VEC>(dst + p * C + c0, v);
  }
}

// --------------------------------------------------------------------------------------
// [B, C, HW] <-> [B, HW, C] through a 32x32 shared tile
template <typename TI, typename TO>
__global__ void transpose_kernel(int rows, int cols, const TI* __restrict__ src, TO* __restrict__ dst) {
  pdl_sync();
  // per batch: src [rows, cols] -> dst [cols, rows]
  __shared__ float tile[32][33];
  const TI* s = src + (int64_t)blockIdx.z * rows * cols;
  TO* d = dst + (int64_t)blockIdx.z * rows * cols;
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = blockIdx.y * 32 + i;
    if (r < rows && c < cols) tile[i][threadIdx.x] = to_f(s[(int64_t)r * cols + c]);
  }
  __syncthreads();
  int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c2 = blockIdx.x * 32 + i;
    if (r2 < rows && c2 < cols) d[(int64_t)c2 * rows + r2] = from_f<TO>(tile[threadIdx.x][i]);
  }
}

template <typename TI, typename TO>
static int launch_transpose(int B, int rows, int cols, const void* src, void* dst, cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, B), block(32, 8);
  launch_k(transpose_kernel<TI, TO>, grid, block, 0, st, rows, cols, (const TI*)src, (TO*)dst);
  return check_launch("transpose");
}

static int transpose_dispatch(int in_dtype, int out_dtype, int B, int rows, int cols, const void* src, void* dst,
                              void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == ACCX_F32 && out_dtype == ACCX_F32) return launch_transpose<float, float>(B, rows, cols, src, dst, st);
  if (in_dtype == ACCX_F32 && out_dtype == ACCX_BF16) return launch_transpose<float, bf16>(B, rows, cols, src, dst, st);
  if (in_dtype == ACCX_BF16 && out_dtype == ACCX_F32) return launch_transpose<bf16, float>(B, rows, cols, src, dst, st);
  if (in_dtype == ACCX_BF16 && out_dtype == ACCX_BF16) return launch_transpose<bf16, bf16>(B, rows, cols, src, dst, st);
  set_error("transpose: bad dtypes %d %d", in_dtype, out_dtype);
  return ACCX_ERR_INVALID;
}

template <typename TI, typename TO>
static int launch_pool_sum(int B, int H, int W, int C, int log2s, float mul, const void* x, void* out, int64_t out_ld,
                           cudaStream_t st) {
  const int64_t Po = (int64_t)B * (H >> log2s) * (W >> log2s);
  // vector lanes only when both sides have the same element size (else scalar lanes; rare)
  constexpr bool same = sizeof(TI) == sizeof(TO);
  const bool al = same && aligned16(x) && aligned16(out) && out_ld % 8 == 0 && C % 8 == 0;
  Lanes l = make_lanes(C, DT<TI>::VEC, al);
  // large windows: few output pixels, many loads per pixel -> one output pixel per thread
  dim3 block(l.tx, l.ty), grid(grid_x_for(Po, l.ty * (log2s >= 3 ? 1 : 4), 148 * 8), l.gy);
  if constexpr (same) {
    if (l.vec != 1) {
      constexpr int V = DT<TI>::VEC;
      if (log2s == 1)
        launch_k(pool_sum_kernel<TI, TO, V, 2>, grid, block, 0, st, B, H, W, C, log2s, mul, (const TI*)x, (TO*)out, out_ld);
      else if (log2s == 2)
        launch_k(pool_sum_kernel<TI, TO, V, 4>, grid, block, 0, st, B, H, W, C, log2s, mul, (const TI*)x, (TO*)out, out_ld);
      else if (log2s == 3)
        launch_k(pool_sum_kernel<TI, TO, V, 8>, grid, block, 0, st, B, H, W, C, log2s, mul, (const TI*)x, (TO*)out, out_ld);
      else
        launch_k(pool_sum_kernel<TI, TO, V, 0>, grid, block, 0, st, B, H, W, C, log2s, mul, (const TI*)x, (TO*)out, out_ld);
      return check_launch("pool_sum");
    }
  }
  launch_k(pool_sum_kernel<TI, TO, 1, 0>, grid, block, 0, st, B, H, W, C, log2s, mul, (const TI*)x, (TO*)out, out_ld);
  return check_launch("pool_sum");
}

}  // namespace accx

using namespace accx;

extern "C" {

int accx_set_knob(int index, int value) {
  ACCX_REQUIRE(index >= 0 && index < KNOB_COUNT, "set_knob: index %d out of range [0, %d)", index, (int)KNOB_COUNT);
  g_knobs[index] = value;
  return ACCX_OK;
}

int accx_set_deterministic(void* workspace, int64_t workspace_bytes, unsigned int* counters, int n_counters) {
  g_det_n_streams = 0;
  if (!workspace) {           // back to the default (atomics)
    g_det.ws = nullptr;
    g_det.ctr = nullptr;
    g_det_floats = 0;
    g_det_ctrs = 0;
    return ACCX_OK;
  }
  ACCX_REQUIRE(workspace_bytes >= (16 << 20) && counters && n_counters >= 16384 && aligned16(workspace),
               "set_deterministic: needs a 16-byte aligned workspace of >= 16 MiB and >= 16384 zeroed counters");
  g_det.ws = (float*)workspace;
  g_det.ctr = counters;
  g_det_floats = workspace_bytes / 4;
  g_det_ctrs = n_counters;
  return ACCX_OK;
}

const char* accx_last_error(void) { return g_err; }
int accx_version(void) { return 100; }

int accx_bn_finalize(int C, double count, const float* stats, const float* gamma, const float* beta,
                     const float* conv_bias, float eps, float momentum, int training, float* running_mean, float* running_var, int64_t* nbt,
                     float* scale, float* shift, float* mean, float* rstd, void* stream) {
  ACCX_REQUIRE(C > 0 && gamma && beta && scale && shift, "bn_finalize: bad arguments");
  ACCX_REQUIRE(training ? (stats != nullptr && count > 0) : (running_mean && running_var),
               "bn_finalize: missing statistics");
  launch_k(bn_finalize_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, 
      C, count, stats, gamma, beta, conv_bias, eps, momentum, training, running_mean, running_var, nbt, scale, shift,
      mean, rstd);
  return check_launch("bn_finalize");
}

int accx_act_apply(int dtype, int64_t P, int C, const void* x, const float* scale, const float* shift, int act,
                   const float* scale2, const float* shift2, const void* residual, void* out, float* stats,
                   void* stream) {
  ACCX_REQUIRE(P > 0 && C > 0 && x, "act_apply: bad arguments");
  ACCX_REQUIRE(act == 0 || (scale && shift), "act_apply: act %d needs scale/shift", act);
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(x) && (!out || aligned16(out)) && (!residual || aligned16(residual)));
    // statistics end in one atomicAdd per channel per block: same-address atomics serialise in L2 (~50 ns each),
    // so the reducing variants run few, long-lived blocks (8 pixels in flight per thread keep HBM busy)
    dim3 block(l.tx, l.ty), grid(grid_x_for(P, l.ty * 8, stats ? 148 * 2 : 148 * 8), l.gy);
    size_t sm = (size_t)l.tx * l.ty * l.vec * sizeof(float);
    Det det;
    if (!det_handle(stats ? (int64_t)grid.x * grid.y * 2 * l.tx * l.vec : 0, grid.y, (cudaStream_t)stream, det)) return ACCX_ERR_INVALID;
    ACCX_DISPATCH_VEC(l, {
      launch_k(act_apply_kernel<T, VEC>, grid, block, sm, (cudaStream_t)stream, 
          P, C, (const T*)x, scale, shift, act, scale2, shift2, (const T*)residual, (T*)out, stats, det);
    });
  });
  return check_launch("act_apply");
}

int accx_add_fwd(int dtype, int64_t P, int C, const void* a, const float* scale, const float* shift, int act,
                 const void* r, void* z, float* stats, void* stream) {
  ACCX_REQUIRE(P > 0 && C > 0 && a && r && z, "add_fwd: bad arguments");
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(a) && aligned16(r) && aligned16(z));
    dim3 block(l.tx, l.ty), grid(grid_x_for(P, l.ty * 8, stats ? 148 * 2 : 148 * 8), l.gy);
    size_t sm = (size_t)l.tx * l.ty * l.vec * sizeof(float);
    Det det;
    if (!det_handle(stats ? (int64_t)grid.x * grid.y * 2 * l.tx * l.vec : 0, grid.y, (cudaStream_t)stream, det)) return ACCX_ERR_INVALID;
    ACCX_DISPATCH_VEC(l, {
      launch_k(add_fwd_kernel<T, VEC>, grid, block, sm, (cudaStream_t)stream, P, C, (const T*)a, scale, shift, act,
                                                                        (const T*)r, (T*)z, stats, det);
    });
  });
  return check_launch("add_fwd");
}

int accx_bn_bwd_reduce(int dtype, int64_t P, int C, const void* y, const float* scale, const float* shift, int act,
                       const float* mean, const float* rstd, const void* da, float* sums, void* stream) {
  ACCX_REQUIRE(P > 0 && C > 0 && y && da && sums && mean && rstd, "bn_bwd_reduce: bad arguments");
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(y) && aligned16(da));
    dim3 block(l.tx, l.ty), grid(grid_x_for(P, l.ty * 8, 148 * knob(KNOB_BN_REDUCE_BLOCKS, 2)), l.gy);
    size_t sm = (size_t)l.tx * l.ty * l.vec * sizeof(float);
    Det det;
    if (!det_handle((int64_t)grid.x * grid.y * 2 * l.tx * l.vec, grid.y, (cudaStream_t)stream, det)) return ACCX_ERR_INVALID;
    ACCX_DISPATCH_VEC(l, {
      launch_k(bn_bwd_reduce_kernel<T, VEC>, grid, block, sm, (cudaStream_t)stream, P, C, (const T*)y, scale, shift, act,
                                                                              mean, rstd, (const T*)da, sums, det);
    });
  });
  return check_launch("bn_bwd_reduce");
}

int accx_bn_bwd_apply(int dtype, int64_t P, int C, const void* y, const float* scale, const float* shift, int act,
                      const float* mean, const float* rstd, const float* gamma, const void* da, const float* sums,
                      double count, void* dy, float* dgamma, float* dbeta, void* stream) {
  ACCX_REQUIRE(P > 0 && C > 0 && y && da && sums && dy && gamma, "bn_bwd_apply: bad arguments");
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(y) && aligned16(da) && aligned16(dy));
    dim3 block(l.tx, l.ty), grid(grid_x_for(P, l.ty * 8, 148 * 8), l.gy);
    ACCX_DISPATCH_VEC(l, {
      launch_k(bn_bwd_apply_kernel<T, VEC>, grid, block, 0, (cudaStream_t)stream, 
          P, C, (const T*)y, scale, shift, act, mean, rstd, gamma, (const T*)da, sums, (float)(1.0 / count), (T*)dy,
          dgamma, dbeta);
    });
  });
  return check_launch("bn_bwd_apply");
}

#define ACCX_DISPATCH_IO(in_dtype, out_dtype, ...)                                  \
  do {                                                                              \
    if ((in_dtype) == ACCX_F32 && (out_dtype) == ACCX_F32) {                        \
      typedef float TI; typedef float TO; __VA_ARGS__                               \
    } else if ((in_dtype) == ACCX_BF16 && (out_dtype) == ACCX_BF16) {               \
      typedef accx::bf16 TI; typedef accx::bf16 TO; __VA_ARGS__                     \
    } else if ((in_dtype) == ACCX_BF16 && (out_dtype) == ACCX_F32) {                \
      typedef accx::bf16 TI; typedef float TO; __VA_ARGS__                          \
    } else if ((in_dtype) == ACCX_F32 && (out_dtype) == ACCX_BF16) {                \
      typedef float TI; typedef accx::bf16 TO; __VA_ARGS__                          \
    } else {                                                                        \
      accx::set_error("bad dtypes %d %d", (int)(in_dtype), (int)(out_dtype));       \
      return ACCX_ERR_INVALID;                                                      \
    }                                                                               \
  } while (0)

int accx_pool_sum(int in_dtype, int out_dtype, int B, int H, int W, int C, int log2s, float mul, const void* x,
                  void* out, int64_t out_ld, void* stream) {
  ACCX_REQUIRE(B > 0 && C > 0 && x && out && log2s >= 0, "pool_sum: bad arguments");
  ACCX_REQUIRE((H >> log2s) << log2s == H && (W >> log2s) << log2s == W, "pool_sum: %dx%d not divisible by %d", H, W,
               1 << log2s);
  ACCX_DISPATCH_IO(in_dtype, out_dtype,
                   { return launch_pool_sum<TI, TO>(B, H, W, C, log2s, mul, x, out, out_ld, (cudaStream_t)stream); });
  return ACCX_OK;
}

int accx_upsample_add(int in_dtype, int out_dtype, int B, int H, int W, int C, int log2s, float mul, const void* src,
                      int64_t src_ld, void* dst, int accumulate, void* stream) {
  ACCX_REQUIRE(B > 0 && C > 0 && src && dst && log2s >= 0, "upsample_add: bad arguments");
  int64_t P = (int64_t)B * H * W;
  ACCX_DISPATCH_IO(in_dtype, out_dtype, {
    bool same = sizeof(TI) == sizeof(TO);
    bool al = same && aligned16(src) && aligned16(dst) && src_ld % 8 == 0 && C % 8 == 0;
    Lanes l = make_lanes(C, DT<TI>::VEC, al);
    dim3 block(l.tx, l.ty), grid(grid_x_for(P, l.ty * 8, 148 * 8), l.gy);
    if (l.vec == 1) {
      launch_k(upsample_add_kernel<TI, TO, 1>, grid, block, 0, (cudaStream_t)stream, B, H, W, C, log2s, mul, (const TI*)src,
                                                                              src_ld, (TO*)dst, accumulate);
    } else {
      launch_k(upsample_add_kernel<TI, TI, DT<TI>::VEC>, grid, block, 0, (cudaStream_t)stream, 
          B, H, W, C, log2s, mul, (const TI*)src, src_ld, (TI*)dst, accumulate);
    }
  });
  return check_launch("upsample_add");
}

int accx_nchw_to_nhwc(int in_dtype, int out_dtype, int B, int C, int HW, const void* src, void* dst, void* stream) {
  ACCX_REQUIRE(B > 0 && C > 0 && HW > 0 && src && dst, "nchw_to_nhwc: bad arguments");
  return transpose_dispatch(in_dtype, out_dtype, B, C, HW, src, dst, stream);
}

int accx_nhwc_to_nchw(int in_dtype, int out_dtype, int B, int C, int HW, const void* src, void* dst, void* stream) {
  ACCX_REQUIRE(B > 0 && C > 0 && HW > 0 && src && dst, "nhwc_to_nchw: bad arguments");
  return transpose_dispatch(in_dtype, out_dtype, B, HW, C, src, dst, stream);
}

}  // extern "C"
