// HANC pyramid pooling (HANCLayer.forward, /root/reference/ACC_UNet/ACC_UNet.py:83-136).
//
// The reference materialises [x, up2(avg2 x), up4(avg4 x), .., up2(max2 x), ..] (2k-1 maps of the
// full resolution) and feeds the interleaved concat to a 1x1 conv.  Here only the POOLED maps are
// produced (1/4, 1/16, .. of the pixels); the 1x1 conv is applied to them at their own resolution
// and the results are nearest-upsample-added in the epilogue of the full-resolution contraction
// (accx_pw_fwd), using  W . cat_j(up(p_j)) = sum_j up(W_j . p_j).
#include <float.h>

#include "common.cuh"

namespace accx {

// one 2x level: out[b,ho,wo, 0:C] = avg 2x2, out[.., C:2C] = max 2x2
template <typename T, int VEC>
__global__ void hanc_pool_kernel(int B, int H, int W, int C, int first, const T* __restrict__ x, const float* scale,
                                 const float* shift, int act, T* __restrict__ out) {
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * VEC >= C) return;
  const int c0 = cv * VEC;
  Lazy<VEC> lz;
  lz.init(scale, shift, first ? act : 0, c0);
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t Po = (int64_t)B * Ho * Wo;
  const int64_t ld = first ? C : 2 * C;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; q < Po; q += (int64_t)gridDim.x * blockDim.y) {
    const int wo = (int)(q % Wo);
    const int64_t t = q / Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float sa[VEC], mx[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { sa[i] = 0.f; mx[i] = -FLT_MAX; }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const T* row = x + (((int64_t)b * H + (2 * ho + i)) * W + 2 * wo) * ld + c0;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[VEC];
        ldv<T, VEC>(row + j * ld, v);
        if (first) {
          lz.apply(v);
#pragma unroll
          for (int e = 0; e < VEC; ++e) { sa[e] += v[e]; mx[e] = fmaxf(mx[e], v[e]); }
        } else {
          float m[VEC];
          ldv<T, VEC>(row + j * ld + C, m);
#pragma unroll
          for (int e = 0; e < VEC; ++e) { sa[e] += v[e]; mx[e] = fmaxf(mx[e], m[e]); }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) sa[e] *= 0.25f;
    stv<T, VEC>(out + q * 2 * C + c0, sa);
    stv<T, VEC>(out + q * 2 * C + C + c0, mx);
  }
}

// da[p] (+)= davg[blk]/s^2 + [p == first row-major argmax of blk] * dmax[blk]
// S > 0: compile-time window (all S*S loads of a block are issued before the first compare); S == 0: generic.
template <typename T, int VEC, int S>
__global__ void hanc_unpool_kernel(int B, int H, int W, int C, int log2s, const T* __restrict__ x, const float* scale,
                                   const float* shift, int act, const float* __restrict__ dpool, T* __restrict__ da,
                                   int accumulate) {
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * VEC >= C) return;
  const int c0 = cv * VEC;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  const int s = S > 0 ? S : 1 << log2s, Ho = H >> log2s, Wo = W >> log2s;
  const float inv = 1.f / (float)(s * s);
  const int64_t Po = (int64_t)B * Ho * Wo;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; q < Po; q += (int64_t)gridDim.x * blockDim.y) {
    const int wo = (int)(q % Wo);
    const int64_t t = q / Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int64_t base = (((int64_t)b * H + (int64_t)ho * s) * W + (int64_t)wo * s) * C + c0;
    float mx[VEC];
    int arg[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { mx[e] = -FLT_MAX; arg[e] = 0; }
    float ga[VEC], gm[VEC];
    if constexpr (S > 0) {
      constexpr bool PRE = S == 2;        // 2x2: the read-modify-write operand is prefetched as well
      RawVec<T, VEC> rx[S * S], rd[PRE ? S * S : 1];
#pragma unroll
      for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) {
          rx[i * S + j].load(x + base + ((int64_t)i * W + j) * C);
          if (PRE && accumulate) rd[PRE ? i * S + j : 0].load(da + base + ((int64_t)i * W + j) * C);
        }
      ldf<VEC>(dpool + q * 2 * C + c0, ga);
      ldf<VEC>(dpool + q * 2 * C + C + c0, gm);
#pragma unroll
      for (int k = 0; k < S * S; ++k) {
        float v[VEC];
        rx[k].unpack(v);
        lz.apply(v);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (v[e] > mx[e]) { mx[e] = v[e]; arg[e] = k; }   // strict > keeps the FIRST maximum
      }
#pragma unroll
      for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) {
          float g[VEC];
          if (accumulate) {
            if constexpr (PRE) rd[i * S + j].unpack(g);
            else ldv<T, VEC>(da + base + ((int64_t)i * W + j) * C, g);
          } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) g[e] = 0.f;
          }
#pragma unroll
          for (int e = 0; e < VEC; ++e) g[e] += ga[e] * inv + (arg[e] == i * S + j ? gm[e] : 0.f);
          stv<T, VEC>(da + base + ((int64_t)i * W + j) * C, g);
        }
    } else {
      for (int i = 0; i < s; ++i)
        for (int j = 0; j < s; ++j) {
          float v[VEC];
          ldv<T, VEC>(x + base + ((int64_t)i * W + j) * C, v);
          lz.apply(v);
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            if (v[e] > mx[e]) { mx[e] = v[e]; arg[e] = i * s + j; }
        }
      ldf<VEC>(dpool + q * 2 * C + c0, ga);
      ldf<VEC>(dpool + q * 2 * C + C + c0, gm);
      for (int i = 0; i < s; ++i)
        for (int j = 0; j < s; ++j) {
          float g[VEC];
          T* dst = da + base + ((int64_t)i * W + j) * C;
          if (accumulate) {
            ldv<T, VEC>(dst, g);
          } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) g[e] = 0.f;
          }
#pragma unroll
          for (int e = 0; e < VEC; ++e) g[e] += ga[e] * inv + (arg[e] == i * s + j ? gm[e] : 0.f);
          stv<T, VEC>(dst, g);
        }
    }
  }
}

}  // namespace accx

using namespace accx;

extern "C" {

int accx_hanc_pool_fwd(int dtype, int B, int H, int W, int C, int first, const void* x, const float* scale,
                       const float* shift, int act, void* out, void* stream) {
  ACCX_REQUIRE(B > 0 && C > 0 && x && out, "hanc_pool_fwd: bad arguments");
  ACCX_REQUIRE(H % 2 == 0 && W % 2 == 0, "hanc_pool_fwd: H, W must be even (got %dx%d)", H, W);
  const int64_t Po = (int64_t)B * (H / 2) * (W / 2);
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(x) && aligned16(out));
    dim3 block(l.tx, l.ty), grid(grid_x_for(Po, l.ty * 4, 148 * 8), l.gy);
    ACCX_DISPATCH_VEC(l, {
      hanc_pool_kernel<T, VEC><<<grid, block, 0, (cudaStream_t)stream>>>(B, H, W, C, first, (const T*)x, scale, shift,
                                                                         act, (T*)out);
    });
  });
  return check_launch("hanc_pool_fwd");
}

int accx_hanc_unpool_bwd(int dtype, int B, int H, int W, int C, int log2s, const void* x, const float* scale,
                         const float* shift, int act, const float* dpool, void* da, int accumulate, void* stream) {
  ACCX_REQUIRE(B > 0 && C > 0 && x && dpool && da && log2s >= 1, "hanc_unpool_bwd: bad arguments");
  ACCX_REQUIRE((H >> log2s) << log2s == H && (W >> log2s) << log2s == W, "hanc_unpool_bwd: %dx%d not divisible by %d",
               H, W, 1 << log2s);
  const int64_t Po = (int64_t)B * (H >> log2s) * (W >> log2s);
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(x) && aligned16(da) && aligned16(dpool));
    dim3 block(l.tx, l.ty), grid(grid_x_for(Po, l.ty * 2, 148 * 8), l.gy);
    ACCX_DISPATCH_VEC(l, {
      if (log2s == 1)
        hanc_unpool_kernel<T, VEC, 2><<<grid, block, 0, (cudaStream_t)stream>>>(B, H, W, C, log2s, (const T*)x, scale,
                                                                                shift, act, dpool, (T*)da, accumulate);
      else if (log2s == 2)
        hanc_unpool_kernel<T, VEC, 4><<<grid, block, 0, (cudaStream_t)stream>>>(B, H, W, C, log2s, (const T*)x, scale,
                                                                                shift, act, dpool, (T*)da, accumulate);
      else
        hanc_unpool_kernel<T, VEC, 0><<<grid, block, 0, (cudaStream_t)stream>>>(B, H, W, C, log2s, (const T*)x, scale,
                                                                                shift, act, dpool, (T*)da, accumulate);
    });
  });
  return check_launch("hanc_unpool_bwd");
}

}  // extern "C"
