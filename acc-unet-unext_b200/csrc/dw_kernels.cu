// Depthwise 3x3 (stride 1, zero pad 1) on NHWC lazy inputs: forward / input-gradient (flip)
// and weight gradient.  HANCBlock.conv2, /root/reference/ACC_UNet/ACC_UNet.py:246-252.
//
// A thread owns one channel vector and walks strips of STRIP pixels along W, so the 3x(STRIP+2)
// input window is read once per strip (4.5 loads per output instead of 9) and the nine filter
// taps of its channels stay in registers.  HBM-bound: reads C*P, writes C*P.
#include "common.cuh"

namespace accx {

constexpr int STRIP = 4;

template <typename T, int VEC>
__global__ void __launch_bounds__(256, 2) dw3x3_fwd_kernel(int B, int H, int W, int C, const T* __restrict__ x, const float* scale,
                                 const float* shift, int act, const float* __restrict__ w,
                                 const float* __restrict__ bias, int flip, T* __restrict__ y, float* stats) {
  pdl_sync();
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  const int c0 = active ? cv * VEC : 0;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float wt[9][VEC], bs[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t][i] = w[(int64_t)(c0 + i) * 9 + (flip ? 8 - t : t)];
    bs[i] = bias ? bias[c0 + i] : 0.f;
  }
  float acc[2][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
  const int strips_w = (W + STRIP - 1) / STRIP;
  const int64_t n_strips = (int64_t)B * H * strips_w;
  if (active) {
    for (int sidx = blockIdx.x * blockDim.y + threadIdx.y; sidx < (int)n_strips; sidx += gridDim.x * blockDim.y) {
      const int w0 = (sidx % strips_w) * STRIP;
      const int t = sidx / strips_w;
      const int h = t % H;
      const int b = t / H;
      // 1) all 18 window loads first (independent, raw storage type: 2 registers each) ...
      RawVec<T, VEC> raw[3][STRIP + 2];
      unsigned inb = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
        const T* row = x + (((int64_t)b * H + hh) * W) * C + c0;
#pragma unroll
        for (int cidx = 0; cidx < STRIP + 2; ++cidx) {
          const int ww = w0 + cidx - 1;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
            raw[r][cidx].load(row + (int64_t)ww * C);
            inb |= 1u << (r * (STRIP + 2) + cidx);
          }
        }
      }
      // 2) ... then normalise + accumulate (zero padding applies to the ACTIVATED tensor: skip outside taps)
      float o[STRIP][VEC];
#pragma unroll
      for (int j = 0; j < STRIP; ++j)
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[j][i] = bs[i];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cidx = 0; cidx < STRIP + 2; ++cidx) {
          if (!(inb & (1u << (r * (STRIP + 2) + cidx)))) continue;
          float v[VEC];
          raw[r][cidx].unpack(v);
          lz.apply(v);
#pragma unroll
          for (int j = 0; j < STRIP; ++j) {
            const int tap = cidx - j;  // column tap 0..2 of output j
            if (tap < 0 || tap > 2) continue;
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[j][i] = fmaf(wt[r * 3 + tap][i], v[i], o[j][i]);
          }
        }
      }
      T* orow = y + (((int64_t)b * H + h) * W) * C + c0;
#pragma unroll
      for (int j = 0; j < STRIP; ++j) {
        if (w0 + j >= W) break;
#pragma unroll
        for (int i = 0; i < VEC; ++i) { acc[0][i] += o[j][i]; acc[1][i] += o[j][i] * o[j][i]; }
        stv<T, VEC>(orow + (int64_t)(w0 + j) * C, o[j]);
      }
    }
  }
  if (stats) reduce_lanes_atomic<2, VEC>(acc, smem, stats, C, C);
}

// dw[c, r, t] += sum_p dy[p] * a[p + (r-1, t-1)]
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 2) dw3x3_wgrad_kernel(int B, int H, int W, int C, const T* __restrict__ x, const float* scale,
                                   const float* shift, int act, const T* __restrict__ dy, float* dw) {
  pdl_sync();
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  const int c0 = active ? cv * VEC : 0;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float acc[9][VEC];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[t][i] = 0.f;
  const int strips_w = (W + STRIP - 1) / STRIP;
  const int64_t n_strips = (int64_t)B * H * strips_w;
  if (active) {
    for (int sidx = blockIdx.x * blockDim.y + threadIdx.y; sidx < (int)n_strips; sidx += gridDim.x * blockDim.y) {
      const int w0 = (sidx % strips_w) * STRIP;
      const int t = sidx / strips_w;
      const int h = t % H;
      const int b = t / H;
      RawVec<T, VEC> graw[STRIP], raw[3][STRIP + 2];
      unsigned inb = 0;
      const T* grow = dy + (((int64_t)b * H + h) * W) * C + c0;
#pragma unroll
      for (int j = 0; j < STRIP; ++j) {
        if (w0 + j < W) graw[j].load(grow + (int64_t)(w0 + j) * C);
        else graw[j].zero();
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
        const T* row = x + (((int64_t)b * H + hh) * W) * C + c0;
#pragma unroll
        for (int cidx = 0; cidx < STRIP + 2; ++cidx) {
          const int ww = w0 + cidx - 1;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
            raw[r][cidx].load(row + (int64_t)ww * C);
            inb |= 1u << (r * (STRIP + 2) + cidx);
          }
        }
      }
      float g[STRIP][VEC];
#pragma unroll
      for (int j = 0; j < STRIP; ++j) graw[j].unpack(g[j]);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cidx = 0; cidx < STRIP + 2; ++cidx) {
          if (!(inb & (1u << (r * (STRIP + 2) + cidx)))) continue;
          float v[VEC];
          raw[r][cidx].unpack(v);
          lz.apply(v);
#pragma unroll
          for (int j = 0; j < STRIP; ++j) {
            const int tap = cidx - j;
            if (tap < 0 || tap > 2) continue;
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[r * 3 + tap][i] = fmaf(g[j][i], v[i], acc[r * 3 + tap][i]);
          }
        }
      }
    }
  }
  // reduce over the block's pixel lanes, then one atomic per (channel, tap)
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VEC; ++i) smem[(ty * VEC + i) * TX + tx] = acc[t][i];
    __syncthreads();
    for (int i = ty; i < VEC; i += TY) {
      float sum = 0.f;
      for (int r = 0; r < TY; ++r) sum += smem[(r * VEC + i) * TX + tx];
      if (active) atomicAdd(dw + (int64_t)(c0 + i) * 9 + t, sum);
    }
  }
}

// dw_tiled.cu: TMA halo-tile kernels (the fast path whenever the tensor is TMA-addressable)
bool dw_tiled_ok(int dtype, int C, const void* x, const void* y_or_dy);
int dw_tiled_fwd(int dtype, int B, int H, int W, int C, const void* x, const float* scale, const float* shift, int act,
                 const float* w, const float* bias, int flip, void* y, float* stats, cudaStream_t st);
int dw_tiled_wgrad(int dtype, int B, int H, int W, int C, const void* x, const float* scale, const float* shift,
                   int act, const void* dy, float* dw, cudaStream_t st);

int dw_tiled_dgrad_bnred(int dtype, int B, int H, int W, int C, const void* dy, const float* w, void* da, const void* y1,
                         const float* bn_scale, const float* bn_shift, int bn_act, const float* bn_mean,
                         const float* bn_rstd, float* sums, cudaStream_t st);

}  // namespace accx

using namespace accx;

extern "C" {

int accx_dw3x3_dgrad_bnred(int dtype, int B, int H, int W, int C, const void* dy, const float* w, void* da,
                           const void* y1, const float* bn_scale, const float* bn_shift, int bn_act,
                           const float* bn_mean, const float* bn_rstd, float* sums, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && dy && w && da && y1 && bn_mean && bn_rstd && sums,
               "dw3x3_dgrad_bnred: bad arguments");
  ACCX_REQUIRE(dtype == ACCX_F32 || dtype == ACCX_BF16, "dw3x3_dgrad_bnred: unsupported dtype %d", dtype);
  ACCX_REQUIRE(bn_act == 0 || (bn_scale && bn_shift), "dw3x3_dgrad_bnred: act %d needs scale/shift", bn_act);
  ACCX_REQUIRE(dw_tiled_ok(dtype, C, dy, da) && dw_tiled_ok(dtype, C, y1, da),
               "dw3x3_dgrad_bnred: tensors must be TMA-addressable (16-byte aligned, C*elem %% 16 == 0); use "
               "accx_dw3x3_fwd(flip) + accx_bn_bwd_reduce");
  return dw_tiled_dgrad_bnred(dtype, B, H, W, C, dy, w, da, y1, bn_scale, bn_shift, bn_act, bn_mean, bn_rstd, sums,
                              (cudaStream_t)stream);
}

int accx_dw3x3_fwd(int dtype, int B, int H, int W, int C, const void* x, const float* scale, const float* shift,
                   int act, const float* w, const float* bias, int flip, void* y, float* stats, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && x && w && y, "dw3x3_fwd: bad arguments");
  ACCX_REQUIRE(act == 0 || (scale && shift), "dw3x3_fwd: act %d needs scale/shift", act);
  ACCX_REQUIRE(dtype == ACCX_F32 || dtype == ACCX_BF16, "dw3x3_fwd: unsupported dtype %d", dtype);
  if (dw_tiled_ok(dtype, C, x, y))
    return dw_tiled_fwd(dtype, B, H, W, C, x, scale, shift, act, w, bias, flip, y, stats, (cudaStream_t)stream);
  const int64_t n_strips = (int64_t)B * H * ((W + STRIP - 1) / STRIP);
  ACCX_DISPATCH_T(dtype, {
    // 4 channels per thread for both dtypes: 9 taps x 4 weights + 4 x 4 accumulators stay under ~80 registers,
    // so three 256-thread CTAs fit on an SM (8-wide bf16 lanes needed 179 registers: one CTA per SM)
    Lanes l = make_lanes(C, 4, aligned16(x) && aligned16(y));
    dim3 block(l.tx, l.ty), grid((stats && det_on()) ? 1 : grid_x_for(n_strips, l.ty * 2, 148 * 12), l.gy);   // det: one add per address
    size_t sm = (size_t)l.tx * l.ty * l.vec * sizeof(float);
    if (l.vec == 1)
      launch_k(dw3x3_fwd_kernel<T, 1>, grid, block, sm, (cudaStream_t)stream, B, H, W, C, (const T*)x, scale, shift, act, w,
                                                                        bias, flip, (T*)y, stats);
    else
      launch_k(dw3x3_fwd_kernel<T, 4>, grid, block, sm, (cudaStream_t)stream, B, H, W, C, (const T*)x, scale, shift, act, w,
                                                                        bias, flip, (T*)y, stats);
  });
  return check_launch("dw3x3_fwd");
}

int accx_dw3x3_wgrad(int dtype, int B, int H, int W, int C, const void* x, const float* scale, const float* shift,
                     int act, const void* dy, float* dw, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && x && dy && dw, "dw3x3_wgrad: bad arguments");
  ACCX_REQUIRE(dtype == ACCX_F32 || dtype == ACCX_BF16, "dw3x3_wgrad: unsupported dtype %d", dtype);
  if (dw_tiled_ok(dtype, C, x, dy))
    return dw_tiled_wgrad(dtype, B, H, W, C, x, scale, shift, act, dy, dw, (cudaStream_t)stream);
  const int64_t n_strips = (int64_t)B * H * ((W + STRIP - 1) / STRIP);
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, 4, aligned16(x) && aligned16(dy));
    dim3 block(l.tx, l.ty), grid(det_on() ? 1 : grid_x_for(n_strips, l.ty * 4, 148 * 3), l.gy);   // det: one add per address
    size_t sm = (size_t)l.tx * l.ty * l.vec * sizeof(float);
    if (l.vec == 1)
      launch_k(dw3x3_wgrad_kernel<T, 1>, grid, block, sm, (cudaStream_t)stream, B, H, W, C, (const T*)x, scale, shift, act,
                                                                          (const T*)dy, dw);
    else
      launch_k(dw3x3_wgrad_kernel<T, 4>, grid, block, sm, (cudaStream_t)stream, B, H, W, C, (const T*)x, scale, shift, act,
                                                                          (const T*)dy, dw);
  });
  return check_launch("dw3x3_wgrad");
}

}  // extern "C"
