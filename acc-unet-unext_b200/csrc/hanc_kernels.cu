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
  pdl_sync();
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
  pdl_sync();
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

// ---------------------------------------------------------------------------------------------------
// Fused HANC backward for k = 2 / 3 (LEVELS = 1 / 2) + BatchNorm-backward reduction of the layer in front:
//   da (+)= sum over levels of [ davg[blk]/s^2 + (p is the FIRST row-major maximum of blk ? dmax[blk] : 0) ]
//   sums[c] += sum_p g, sums[C+c] += sum_p g * xhat,   g = da_total * act'(y),  xhat = (y - mean) * rstd
// One thread owns a 2^LEVELS-square window of 4 channels: y and da are read ONCE (the per-level kernels read y
// and read-modify-write da once per level, and the BN reduction reads both again), every load of the window is
// in flight before the first compare.  bf16 storage.
template <int LEVELS, int VEC, int OCC = 1>
__global__ void __launch_bounds__(128, OCC) hanc_unpool_bnred_kernel(int B, int H, int W, int C, const bf16* __restrict__ y,
                                                                const float* scale, const float* shift, int act,
                                                                const float* __restrict__ dp1,
                                                                const float* __restrict__ dp2, bf16* __restrict__ da,
                                                                const float* mean, const float* rstd, float* sums,
                                                                Det det) {
  pdl_sync();
  constexpr int S = 1 << LEVELS, NPX = S * S;
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  const int c0 = active ? cv * VEC : 0;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float mu[VEC];
  ldf<VEC>(mean + c0, mu);
  float acc[2][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
  const int Ho = H / S, Wo = W / S;
  const int64_t n_win = (int64_t)B * Ho * Wo;
  if (active) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; q < n_win; q += (int64_t)gridDim.x * blockDim.y) {
      const int wo = (int)(q % Wo);
      const int64_t t = q / Wo;
      const int ho = (int)(t % Ho);
      const int b = (int)(t / Ho);
      const int64_t base = (((int64_t)b * H + (int64_t)ho * S) * W + (int64_t)wo * S) * C + c0;
      RawVec<bf16, VEC> ry[NPX], rd[NPX];
#pragma unroll
      for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) {
          ry[i * S + j].load(y + base + ((int64_t)i * W + j) * C);
          rd[i * S + j].load(da + base + ((int64_t)i * W + j) * C);
        }
      // level-1 gradients of the (S/2)^2 2x2 sub-windows, level-2 gradient of the whole 4x4 window
      constexpr int NSUB = (S / 2) * (S / 2);
      float ga1[NSUB][VEC], gm1[NSUB][VEC], ga2[VEC], gm2[VEC];
#pragma unroll
      for (int si = 0; si < S / 2; ++si)
#pragma unroll
        for (int sj = 0; sj < S / 2; ++sj) {
          const int64_t q1 = (((int64_t)b * (H >> 1) + (ho * (S / 2) + si)) * (W >> 1) + (wo * (S / 2) + sj)) * 2 * C + c0;
          ldf<VEC>(dp1 + q1, ga1[si * (S / 2) + sj]);
          ldf<VEC>(dp1 + q1 + C, gm1[si * (S / 2) + sj]);
        }
      if constexpr (LEVELS == 2) {
        ldf<VEC>(dp2 + q * 2 * C + c0, ga2);
        ldf<VEC>(dp2 + q * 2 * C + C + c0, gm2);
      }
      // ---- pass 1: arg-maxima (first maximum in row-major order of each window) ----
      float mx4[VEC];
      int arg4[VEC], arg2[NSUB][VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) { mx4[e] = -FLT_MAX; arg4[e] = 0; }
#pragma unroll
      for (int sub = 0; sub < NSUB; ++sub) {
        const int si = sub / (S / 2), sj = sub % (S / 2);
        float mx2[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) { mx2[e] = -FLT_MAX; arg2[sub][e] = 0; }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int idx = (2 * si + (qq >> 1)) * S + 2 * sj + (qq & 1);
          float v[VEC];
          ry[idx].unpack(v);
          lz.apply(v);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            if (v[e] > mx2[e]) { mx2[e] = v[e]; arg2[sub][e] = qq; }
            if (LEVELS == 2 && (v[e] > mx4[e] || (v[e] == mx4[e] && idx < arg4[e]))) { mx4[e] = v[e]; arg4[e] = idx; }
          }
        }
      }
      // ---- pass 2: total gradient, store, BatchNorm-backward partial sums ----
#pragma unroll
      for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) {
          const int idx = i * S + j, sub = (i >> 1) * (S / 2) + (j >> 1), qq = (i & 1) * 2 + (j & 1);
          float g[VEC], yv[VEC];
          rd[idx].unpack(g);
          ry[idx].unpack(yv);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            g[e] += ga1[sub][e] * 0.25f + (arg2[sub][e] == qq ? gm1[sub][e] : 0.f);
            if (LEVELS == 2) g[e] += ga2[e] * 0.0625f + (arg4[e] == idx ? gm2[e] : 0.f);
          }
          stv<bf16, VEC>(da + base + ((int64_t)i * W + j) * C, g);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float gi = g[e] * lz.dact(yv[e], e);
            acc[0][e] += gi;
            acc[1][e] = fmaf(gi, yv[e] - mu[e], acc[1][e]);
          }
        }
    }
  }
  float rs[VEC];
  ldf<VEC>(rstd + c0, rs);
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[1][e] *= rs[e];
  reduce_lanes_atomic<2, VEC>(acc, smem, sums, C, C, det, blockIdx.y, blockIdx.x, gridDim.x);
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
      launch_k(hanc_pool_kernel<T, VEC>, grid, block, 0, (cudaStream_t)stream, B, H, W, C, first, (const T*)x, scale, shift,
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
        launch_k(hanc_unpool_kernel<T, VEC, 2>, grid, block, 0, (cudaStream_t)stream, B, H, W, C, log2s, (const T*)x, scale,
                                                                                shift, act, dpool, (T*)da, accumulate);
      else if (log2s == 2)
        launch_k(hanc_unpool_kernel<T, VEC, 4>, grid, block, 0, (cudaStream_t)stream, B, H, W, C, log2s, (const T*)x, scale,
                                                                                shift, act, dpool, (T*)da, accumulate);
      else
        launch_k(hanc_unpool_kernel<T, VEC, 0>, grid, block, 0, (cudaStream_t)stream, B, H, W, C, log2s, (const T*)x, scale,
                                                                                shift, act, dpool, (T*)da, accumulate);
    });
  });
  return check_launch("hanc_unpool_bwd");
}

int accx_hanc_unpool_bnred(int dtype, int B, int H, int W, int C, int levels, const void* y, const float* scale,
                           const float* shift, int act, const float* dpool1, const float* dpool2, void* da,
                           const float* mean, const float* rstd, float* sums, void* stream) {
  ACCX_REQUIRE(B > 0 && C > 0 && y && dpool1 && da && mean && rstd && sums, "hanc_unpool_bnred: bad arguments");
  ACCX_REQUIRE(dtype == ACCX_BF16, "hanc_unpool_bnred: bf16 storage only (use accx_hanc_unpool_bwd + accx_bn_bwd_reduce)");
  ACCX_REQUIRE(levels == 1 || (levels == 2 && dpool2), "hanc_unpool_bnred: levels must be 1 or 2 (k = 2 or 3)");
  const int S = 1 << levels;
  ACCX_REQUIRE(H % S == 0 && W % S == 0, "hanc_unpool_bnred: %dx%d not divisible by %d", H, W, S);
  ACCX_REQUIRE(C % 4 == 0 && aligned16(dpool1) && (!dpool2 || aligned16(dpool2)) &&
                   (reinterpret_cast<uintptr_t>(y) & 7) == 0 && (reinterpret_cast<uintptr_t>(da) & 7) == 0,
               "hanc_unpool_bnred: needs C %% 4 == 0 and aligned tensors");
  ACCX_REQUIRE(act == 0 || (scale && shift), "hanc_unpool_bnred: act %d needs scale/shift", act);
  const int64_t n_win = (int64_t)B * (H / S) * (W / S);
  // 128-thread blocks: the window state is register heavy.  4 x 4 windows (two levels) with 4 channels per thread hold 255
  // registers (two blocks = 8 warps per SM, whose load / compute / store phases leave HBM idle half of the time); with 2
  // channels per thread (4-byte accesses, a warp still covers whole 128-byte lines) 168 registers and three blocks, or --
  // the default -- held to 128 registers (24 bytes of spills) and four blocks: 16x56x56x4352 683 -> 437 -> 396 us,
  // 16x224x224x192 450 -> 333 -> 310 us (profiles/r02_unpool_channels_per_thread.txt).
  // 2 x 2 windows (one level, 118 registers) are faster with 4 channels per thread (71 vs 95 us at 16x28x28x1536).
  Lanes l;
  const int vk = knob(KNOB_UNPOOL_VEC, levels == 2 ? 3 : 4);      // 3 = two channels held to 128 registers (four blocks per SM)
  l.vec = (vk == 2 || vk == 3) ? 2 : 4;
  l.cvn = C / l.vec;
  l.tx = l.cvn <= 128 ? l.cvn : 128;
  for (int d = 128; l.cvn > 128 && d >= 32; --d)
    if (l.cvn % d == 0) { l.tx = d; break; }
  l.ty = 128 / l.tx;
  l.gy = (l.cvn + l.tx - 1) / l.tx;
  dim3 block(l.tx, l.ty), grid(grid_x_for(n_win, l.ty, 148 * (vk == 3 ? 8 : (l.vec == 2 ? 6 : 3))), l.gy);
  const size_t sm = (size_t)l.tx * l.ty * l.vec * sizeof(float);
  Det det;
  if (!det_handle((int64_t)grid.x * grid.y * 2 * l.tx * l.vec, grid.y, (cudaStream_t)stream, det)) return ACCX_ERR_INVALID;
#define ACCX_UNPOOL_LAUNCH(LV, VC)                                                                                          \
  launch_k(hanc_unpool_bnred_kernel<LV, VC>, grid, block, sm, (cudaStream_t)stream, B, H, W, C, (const bf16*)y, scale, shift, \
           act, dpool1, dpool2, (bf16*)da, mean, rstd, sums, det)
  if (levels == 1) {
    if (l.vec == 2) ACCX_UNPOOL_LAUNCH(1, 2); else ACCX_UNPOOL_LAUNCH(1, 4);
  } else if (vk == 3) {
    launch_k(hanc_unpool_bnred_kernel<2, 2, 4>, grid, block, sm, (cudaStream_t)stream, B, H, W, C, (const bf16*)y, scale, shift,
             act, dpool1, dpool2, (bf16*)da, mean, rstd, sums, det);
  } else {
    if (l.vec == 2) ACCX_UNPOOL_LAUNCH(2, 2); else ACCX_UNPOOL_LAUNCH(2, 4);
  }
#undef ACCX_UNPOOL_LAUNCH
  return check_launch("hanc_unpool_bnred");
}


}  // extern "C"
