// The steps either side of the HANC/MLFC blocks inside one training step:
//   * MaxPool2d(2) between encoder levels            /root/reference/ACC_UNet/ACC_UNet.py:552,608-618
//   * WeightedDiceBCE(0.5, 0.5) on logits, fwd + bwd  /root/reference/Experiments/utils.py:21-74,109-171
//   * Adam over ONE flat parameter / gradient buffer  /root/reference/Experiments/train_model.py:647
// All HBM-bound, each tensor touched once.
#include "common.cuh"

namespace accx {

// ---------------------------------------------------------------------------------------------------
// MaxPool2d(kernel 2, stride 2), NHWC.  Thread = one output pixel x VEC channels; the four window loads are
// issued before the first compare.  Backward recomputes the arg-max from the input (no index tensor) and
// routes the gradient to the FIRST maximum in row-major window order (ATen's tie rule: strict '>').
template <typename T, int VEC>
__global__ void maxpool2_fwd_kernel(int B, int H, int W, int C, const T* __restrict__ x, T* __restrict__ out) {
  pdl_sync();
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * VEC >= C) return;
  const int c0 = cv * VEC;
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t Po = (int64_t)B * Ho * Wo;
  constexpr int U = 2;
  RawVec<T, VEC> r[U][4];
  pixel_loop<U>((int64_t)blockIdx.x * blockDim.y + threadIdx.y, Po, (int64_t)gridDim.x * blockDim.y,
      [&](int u, int64_t q) {
        const int wo = (int)(q % Wo);
        const int64_t t = q / Wo;
        const int ho = (int)(t % Ho);
        const int64_t b = t / Ho;
        const T* p = x + (((b * H + 2 * ho) * W) + 2 * wo) * C + c0;
        r[u][0].load(p);
        r[u][1].load(p + C);
        r[u][2].load(p + (int64_t)W * C);
        r[u][3].load(p + (int64_t)W * C + C);
      },
      [&](int u, int64_t q) {
        float m[VEC], v[VEC];
        r[u][0].unpack(m);
#pragma unroll
        for (int j = 1; j < 4; ++j) {
          r[u][j].unpack(v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) m[i] = v[i] > m[i] ? v[i] : m[i];
        }
        stv<T, VEC>(out + q * C + c0, m);
      });
}

template <typename T, int VEC>
__global__ void maxpool2_bwd_kernel(int B, int H, int W, int C, const T* __restrict__ x, const T* __restrict__ dy,
                                    T* __restrict__ dx) {
  pdl_sync();
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * VEC >= C) return;
  const int c0 = cv * VEC;
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t Po = (int64_t)B * Ho * Wo;
  constexpr int U = 2;
  RawVec<T, VEC> r[U][4], rd[U];
  pixel_loop<U>((int64_t)blockIdx.x * blockDim.y + threadIdx.y, Po, (int64_t)gridDim.x * blockDim.y,
      [&](int u, int64_t q) {
        const int wo = (int)(q % Wo);
        const int64_t t = q / Wo;
        const int ho = (int)(t % Ho);
        const int64_t b = t / Ho;
        const T* p = x + (((b * H + 2 * ho) * W) + 2 * wo) * C + c0;
        r[u][0].load(p);
        r[u][1].load(p + C);
        r[u][2].load(p + (int64_t)W * C);
        r[u][3].load(p + (int64_t)W * C + C);
        rd[u].load(dy + q * C + c0);
      },
      [&](int u, int64_t q) {
        const int wo = (int)(q % Wo);
        const int64_t t = q / Wo;
        const int ho = (int)(t % Ho);
        const int64_t b = t / Ho;
        float m[VEC], v[4][VEC], d[VEC];
        int arg[VEC];
        rd[u].unpack(d);
#pragma unroll
        for (int j = 0; j < 4; ++j) r[u][j].unpack(v[j]);
#pragma unroll
        for (int i = 0; i < VEC; ++i) { m[i] = v[0][i]; arg[i] = 0; }
#pragma unroll
        for (int j = 1; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < VEC; ++i)
            if (v[j][i] > m[i]) { m[i] = v[j][i]; arg[i] = j; }
        T* o = dx + (((b * H + 2 * ho) * W) + 2 * wo) * C + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float g[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) g[i] = arg[i] == j ? d[i] : 0.f;
          stv<T, VEC>(o + (int64_t)(j >> 1) * W * C + (j & 1) * C, g);
        }
      });
}

// ---------------------------------------------------------------------------------------------------
// WeightedDiceBCE on logits, one class.  Per image b over its N pixels (p = sigmoid(x)/2, t = y/2: the class
// weights [0.5, 0.5] of the reference enter as these factors):
//   sums[b][0] = sum p*t   [1] = sum p*p   [2] = sum t*t
//   sums[b][3] = sum_{y>0.5} bce(x,y)   [4] = sum_{y<=0.5} bce(x,y)   [5] = #{y>0.5}
//   dice = mean_b 1 - (2*s0 + 1e-5)/(s1 + s2 + 1e-5)
//   bce  = 0.5 * S3/max(S5,1) + 0.5 * S4/max(BN - S5, 1)          (S = sums over all images)
//   loss = dice_w*dice + bce_w*bce
constexpr int LOSS_NS = 6, LOSS_STRIDE = 8;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

struct LossTotals { float dice, bce, npos, nneg; };

__device__ __forceinline__ LossTotals loss_totals(const float* sums, int B, int64_t N) {
  double dice = 0, s3 = 0, s4 = 0, s5 = 0;
  for (int b = 0; b < B; ++b) {
    const float* s = sums + b * LOSS_STRIDE;      // written by atomics of other blocks: read through L2
    dice += 1.0 - (2.0 * __ldcg(s) + 1e-5) / ((double)__ldcg(s + 1) + __ldcg(s + 2) + 1e-5);
    s3 += __ldcg(s + 3); s4 += __ldcg(s + 4); s5 += __ldcg(s + 5);
  }
  LossTotals t;
  const double total = (double)B * (double)N;
  t.npos = (float)(s5 < 1.0 ? 1.0 : s5);
  t.nneg = (float)((total - s5) < 1.0 ? 1.0 : (total - s5));
  t.dice = (float)(dice / B);
  t.bce = (float)(0.5 * s3 / t.npos + 0.5 * s4 / t.nneg);
  return t;
}

template <typename T>
__global__ void __launch_bounds__(256) dice_bce_fwd_kernel(int B, int64_t N, const T* __restrict__ logit,
                                                           const float* __restrict__ truth, float dice_w, float bce_w,
                                                           float* sums, unsigned int* counter, float* loss) {
  pdl_sync();
  __shared__ float red[8][LOSS_NS];
  __shared__ bool is_last;
  const int b = blockIdx.y;
  const T* x = logit + (int64_t)b * N;
  const float* y = truth + (int64_t)b * N;
  float a[LOSS_NS];
#pragma unroll
  for (int i = 0; i < LOSS_NS; ++i) a[i] = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = to_f(x[i]), yv = y[i];
    const float p = 0.5f * sigmoidf_(xv), t = 0.5f * yv;
    const float l = fmaxf(xv, 0.f) - xv * yv + log1pf(__expf(-fabsf(xv)));
    const bool pos = yv > 0.5f;
    a[0] += p * t;
    a[1] += p * p;
    a[2] += t * t;
    a[3] += pos ? l : 0.f;
    a[4] += pos ? 0.f : l;
    a[5] += pos ? 1.f : 0.f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < LOSS_NS; ++i) {
    float v = a[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < LOSS_NS) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
    atomicAdd(sums + b * LOSS_STRIDE + threadIdx.x, s);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    is_last = atomicAdd(counter, 1u) == total - 1;
    if (is_last) *counter = 0;          // re-arm (CUDA-graph replays reuse the buffer)
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    LossTotals t = loss_totals((const float*)sums, B, N);
    *loss = dice_w * t.dice + bce_w * t.bce;
  }
}

// Per-step segmentation metrics of the training loop, left ON THE DEVICE (Experiments/Train_one_epoch.py:134-135 calls
// iou_on_batch -> .cpu().numpy() + sklearn every step: a device-to-host sync per step).  One pass over logits and masks:
//   pred = sigmoid(logit) >= 0.5, mask = truth > 0                                  (Experiments/utils.py:478-494)
//   counts[b] = {TP, #pred, #mask} as integers (order-free, exact)
//   iou_b  = TP / (#pred + #mask - TP)   (sklearn.metrics.jaccard_score, binary; 0 when the union is empty)
//   dice_b = WeightedDiceBCE._show_dice (utils.py:148-157): 1 - WeightedDiceLoss(binarised pred, mask) -- which
//            passes the 0/1 prediction through sigmoid AGAIN (utils.py:121), so p is 0.5*sigmoid(1) or 0.5*sigmoid(0)
// out[0] = mean_b iou_b, out[1] = mean_b dice_b, written by the last block.
template <typename T>
__global__ void __launch_bounds__(256) seg_metrics_kernel(int B, int64_t N, const T* __restrict__ logit,
                                                          const float* __restrict__ truth, unsigned int* counts,
                                                          float* out) {
  pdl_sync();
  __shared__ unsigned int red[8][3];
  __shared__ bool is_last;
  const int b = blockIdx.y;
  const T* x = logit + (int64_t)b * N;
  const float* y = truth + (int64_t)b * N;
  unsigned int tp = 0, np = 0, nm = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const bool p = 1.f / (1.f + expf(-to_f(x[i]))) >= 0.5f;
    const bool m = y[i] > 0.f;
    tp += (p && m) ? 1u : 0u;
    np += p ? 1u : 0u;
    nm += m ? 1u : 0u;
  }
  tp = __reduce_add_sync(0xffffffffu, tp);
  np = __reduce_add_sync(0xffffffffu, np);
  nm = __reduce_add_sync(0xffffffffu, nm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[warp][0] = tp; red[warp][1] = np; red[warp][2] = nm; }
  __syncthreads();
  if (threadIdx.x < 3) {
    unsigned int s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
    atomicAdd(counts + b * 4 + threadIdx.x, s);
  }
  __threadfence();
  __syncthreads();
  unsigned int* counter = counts + (int64_t)B * 4;
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    is_last = atomicAdd(counter, 1u) == total - 1;
    if (is_last) *counter = 0;
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    const double s1 = 1.0 / (1.0 + exp(-1.0));       // sigmoid(1): the second sigmoid of _show_dice on a positive
    double iou = 0, dice = 0;
    for (int i = 0; i < B; ++i) {
      const double TP = __ldcg(counts + i * 4), NP = __ldcg(counts + i * 4 + 1), NM = __ldcg(counts + i * 4 + 2);
      const double uni = NP + NM - TP;
      iou += uni > 0 ? TP / uni : 0.0;
      const double FN = NM - TP, N0 = (double)N - NP;
      const double inter = 0.25 * (s1 * TP + 0.5 * FN);
      const double pp = 0.25 * (s1 * s1 * NP + 0.25 * N0), tt = 0.25 * NM;
      dice += (2.0 * inter + 1e-5) / (pp + tt + 1e-5);
    }
    out[0] = (float)(iou / B);
    out[1] = (float)(dice / B);
  }
}

// dlogit = gscale * d loss / d logit (TG = gradient storage type); dbias += sum dlogit (the final 1x1 conv's bias)
template <typename T, typename TG>
__global__ void __launch_bounds__(256) dice_bce_bwd_kernel(int B, int64_t N, const T* __restrict__ logit,
                                                           const float* __restrict__ truth,
                                                           const float* __restrict__ sums, float dice_w, float bce_w,
                                                           const float* gscale, TG* __restrict__ dlogit) {
  pdl_sync();
  __shared__ LossTotals tot;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) tot = loss_totals(sums, B, N);
  __syncthreads();
  const float gs = gscale ? *gscale : 1.f;
  const float* s = sums + b * LOSS_STRIDE;
  const float I2 = 2.f * s[0] + 1e-5f, U = s[1] + s[2] + 1e-5f;
  const float kd = gs * dice_w / (float)B / (U * U);
  const float kpos = gs * bce_w * 0.5f / tot.npos, kneg = gs * bce_w * 0.5f / tot.nneg;
  const T* x = logit + (int64_t)b * N;
  const float* y = truth + (int64_t)b * N;
  TG* d = dlogit + (int64_t)b * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = to_f(x[i]), yv = y[i];
    const float sg = sigmoidf_(xv);
    const float p = 0.5f * sg, t = 0.5f * yv, dp = 0.5f * sg * (1.f - sg);
    // d/dx [1 - I2/U] = -(2 t dp U - I2 * 2 p dp) / U^2
    const float gd = -kd * (2.f * t * dp * U - I2 * 2.f * p * dp);
    const float gb = (sg - yv) * (yv > 0.5f ? kpos : kneg);
    d[i] = from_f<TG>(gd + gb);
  }
}

// ---------------------------------------------------------------------------------------------------
// Adam over one flat fp32 buffer (all parameters of the model back to back): 4 reads + 3 writes per element in
// ONE launch instead of a multi-tensor launch chain over ~900 tensors.  state[0] = step count (device side, so
// the launch is graph-capturable); same arithmetic as torch.optim.Adam (no amsgrad, L2 weight decay).
__global__ void adam_tick_kernel(float* state) {
  pdl_sync(); state[0] += 1.f; }

__global__ void __launch_bounds__(256) adam_flat_kernel(int64_t n4, float4* __restrict__ p, const float4* __restrict__ g,
                                                        float4* __restrict__ m, float4* __restrict__ v,
                                                        const float* __restrict__ state, float lr, float b1, float b2,
                                                        float eps, float wd, float gscale) {
  pdl_sync();
  const double t = (double)state[0];
  const float bc1 = (float)(1.0 - pow((double)b1, t));
  const float bc2s = (float)sqrt(1.0 - pow((double)b2, t));
  // lr < 0: the learning rate lives on the device (state[1]) -- a captured step graph then follows an LR schedule
  // (CosineAnnealingWarmRestarts, Experiments/train_model.py:738) without being re-captured
  const float step_size = (lr >= 0.f ? lr : state[1]) / bc1;
  constexpr int U = 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += U * stride) {
    float4 P[U], G[U], M[U], V[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) { P[u] = p[i]; G[u] = g[i]; M[u] = m[i]; V[u] = v[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n4) continue;
      float* pp = reinterpret_cast<float*>(&P[u]);
      float* gg = reinterpret_cast<float*>(&G[u]);
      float* mm = reinterpret_cast<float*>(&M[u]);
      float* vv = reinterpret_cast<float*>(&V[u]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float gr = gg[j] * gscale;
        if (wd != 0.f) gr = fmaf(pp[j], wd, gr);
        mm[j] = mm[j] + (1.f - b1) * (gr - mm[j]);
        vv[j] = b2 * vv[j] + (1.f - b2) * gr * gr;
        const float denom = sqrtf(vv[j]) / bc2s + eps;
        pp[j] -= step_size * mm[j] / denom;
      }
      p[i] = P[u]; m[i] = M[u]; v[i] = V[u];
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// ConvTranspose2d(kernel 2, stride 2) + skip concat (ACC_UNet.py:578-599,620-631).  The transposed conv is ONE
// pointwise contraction [P, Cin] x [Cin, 4*Co] whose column co*4 + (ky*2+kx) is output pixel (2h+ky, 2w+kx),
// channel co (the reference's weight layout [Cin, Co, 2, 2] IS that matrix, so the contraction reads it through a
// strided view); these kernels interleave its result into the left Co columns of the [B, 2H, 2W, ld_out] concat
// buffer (+ bias) and back, and copy column ranges (the skip half of the concat, its gradient).
//   thread = one input pixel x 2 channels: a 16-byte load of (2 channels x 4 taps), four 4-byte stores (bf16) --
//   consecutive threads write consecutive channel pairs of the same output pixel.
template <typename T, bool FWD>
__global__ void __launch_bounds__(256) upshuffle_kernel(int B, int H, int W, int Co, T* __restrict__ temp,
                                                        const float* __restrict__ bias, T* __restrict__ out,
                                                        int64_t ld_out, float* dbias) {
  pdl_sync();
  extern __shared__ float smem[];
  const int pairs = Co >> 1;
  const int cp = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cp < pairs;
  const int co = active ? cp * 2 : 0;
  float bacc[1][2] = {{0.f, 0.f}};                     // backward: the transposed conv's bias gradient
  const float b0 = (FWD && bias) ? bias[co] : 0.f, b1 = (FWD && bias) ? bias[co + 1] : 0.f;
  const int64_t P = (int64_t)B * H * W;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; active && q < P; q += (int64_t)gridDim.x * blockDim.y) {
    const int w = (int)(q % W);
    const int64_t t = q / W;
    const int h = (int)(t % H);
    const int64_t b = t / H;
    T* trow = temp + q * (4 * (int64_t)Co) + co * 4;
    T* o00 = out + (((b * 2 * H + 2 * h) * 2 * W) + 2 * w) * ld_out + co;
    const int64_t dnx = ld_out, dny = 2 * (int64_t)W * ld_out;
    if constexpr (FWD) {
      float v[8];
      if constexpr (sizeof(T) == 2) {
        ldv<T, 8>(trow, v);
      } else {
        ldv<T, 4>(trow, *reinterpret_cast<float(*)[4]>(v));
        ldv<T, 4>(trow + 4, *reinterpret_cast<float(*)[4]>(v + 4));
      }
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        T* o = o00 + (tap >> 1) * dny + (tap & 1) * dnx;        // ld_out and co are even: one 4 / 8-byte store
        if constexpr (sizeof(T) == 2) *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(v[tap] + b0, v[4 + tap] + b1);
        else *reinterpret_cast<float2*>(o) = make_float2(v[tap] + b0, v[4 + tap] + b1);
      }
    } else {
      float v[8];
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        const T* o = o00 + (tap >> 1) * dny + (tap & 1) * dnx;
        if constexpr (sizeof(T) == 2) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(o));
          v[tap] = f.x;
          v[4 + tap] = f.y;
        } else {
          const float2 f = *reinterpret_cast<const float2*>(o);
          v[tap] = f.x;
          v[4 + tap] = f.y;
        }
      }
      if constexpr (sizeof(T) == 2) {
        stv<T, 8>(trow, v);
      } else {
        stv<T, 4>(trow, *reinterpret_cast<float(*)[4]>(v));
        stv<T, 4>(trow + 4, *reinterpret_cast<float(*)[4]>(v + 4));
      }
      bacc[0][0] += (v[0] + v[1]) + (v[2] + v[3]);
      bacc[0][1] += (v[4] + v[5]) + (v[6] + v[7]);
    }
  }
  if constexpr (!FWD) {
    if (dbias) reduce_lanes_atomic<1, 2>(bacc, smem, dbias, 0, Co);
  }
}

// dst[p, 0..C) = src[p, 0..C) for row pitches ld_src / ld_dst (column ranges of wider matrices)
template <typename T, int VEC>
__global__ void copy_cols_kernel(int64_t P, int C, const T* __restrict__ src, int64_t ld_src, T* __restrict__ dst,
                                 int64_t ld_dst) {
  pdl_sync();
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * VEC >= C) return;
  const int c0 = cv * VEC;
  constexpr int U = 4;
  RawVec<T, VEC> r[U];
  pixel_loop<U>((int64_t)blockIdx.x * blockDim.y + threadIdx.y, P, (int64_t)gridDim.x * blockDim.y,
      [&](int u, int64_t p) { r[u].load(src + p * ld_src + c0); },
      [&](int u, int64_t p) {
        float v[VEC];
        r[u].unpack(v);
        stv<T, VEC>(dst + p * ld_dst + c0, v);
      });
}

}  // namespace accx

using namespace accx;

extern "C" {

int accx_maxpool2_fwd(int dtype, int B, int H, int W, int C, const void* x, void* out, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 1 && W > 1 && C > 0 && x && out, "maxpool2_fwd: bad arguments");
  const int64_t Po = (int64_t)B * (H >> 1) * (W >> 1);
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(x) && aligned16(out));
    dim3 block(l.tx, l.ty), grid(grid_x_for(Po, l.ty * 2, 148 * 16), l.gy);
    ACCX_DISPATCH_VEC(l, {
      launch_k(maxpool2_fwd_kernel<T, VEC>, grid, block, 0, (cudaStream_t)stream, B, H, W, C, (const T*)x, (T*)out);
    });
  });
  return check_launch("maxpool2_fwd");
}

int accx_maxpool2_bwd(int dtype, int B, int H, int W, int C, const void* x, const void* dy, void* dx, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 1 && W > 1 && C > 0 && x && dy && dx, "maxpool2_bwd: bad arguments");
  ACCX_REQUIRE(H % 2 == 0 && W % 2 == 0, "maxpool2_bwd: H and W must be even (got %dx%d)", H, W);
  const int64_t Po = (int64_t)B * (H >> 1) * (W >> 1);
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(x) && aligned16(dy) && aligned16(dx));
    dim3 block(l.tx, l.ty), grid(grid_x_for(Po, l.ty * 2, 148 * 16), l.gy);
    ACCX_DISPATCH_VEC(l, {
      launch_k(maxpool2_bwd_kernel<T, VEC>, grid, block, 0, (cudaStream_t)stream, B, H, W, C, (const T*)x, (const T*)dy, (T*)dx);
    });
  });
  return check_launch("maxpool2_bwd");
}

static inline int loss_grid_x(int B, int64_t N) {
  int64_t per = (N + 256 * 8 - 1) / (256 * 8);           // ~8 elements per thread
  int64_t want = (148 * 4 + B - 1) / B;
  int64_t g = per < want ? per : want;
  return (int)(g < 1 ? 1 : g);
}

int accx_dice_bce_fwd(int dtype, int B, int64_t N, const void* logit, const float* truth, float dice_w, float bce_w,
                      float* sums, float* loss, void* stream) {
  ACCX_REQUIRE(B > 0 && N > 0 && logit && truth && sums && loss, "dice_bce_fwd: bad arguments");
  dim3 grid(det_on() ? 1 : loss_grid_x(B, N), B);       // deterministic mode: one block (= one contribution) per image
  unsigned int* counter = reinterpret_cast<unsigned int*>(sums + (int64_t)B * LOSS_STRIDE);
  ACCX_DISPATCH_T(dtype, {
    launch_k(dice_bce_fwd_kernel<T>, grid, 256, 0, (cudaStream_t)stream, B, N, (const T*)logit, truth, dice_w, bce_w, sums,
                                                                    counter, loss);
  });
  return check_launch("dice_bce_fwd");
}

int accx_seg_metrics(int dtype, int B, int64_t N, const void* logit, const float* truth, unsigned int* counts, float* out,
                     void* stream) {
  ACCX_REQUIRE(B > 0 && N > 0 && N < ((int64_t)1 << 31) && logit && truth && counts && out, "seg_metrics: bad arguments");
  dim3 grid(loss_grid_x(B, N), B);
  ACCX_DISPATCH_T(dtype, {
    launch_k(seg_metrics_kernel<T>, grid, 256, 0, (cudaStream_t)stream, B, N, (const T*)logit, truth, counts, out);
  });
  return check_launch("seg_metrics");
}

int accx_dice_bce_bwd(int dtype, int grad_dtype, int B, int64_t N, const void* logit, const float* truth,
                      const float* sums, float dice_w, float bce_w, const float* gscale, void* dlogit, void* stream) {
  ACCX_REQUIRE(B > 0 && N > 0 && logit && truth && sums && dlogit, "dice_bce_bwd: bad arguments");
  dim3 grid(loss_grid_x(B, N), B);
  cudaStream_t st = (cudaStream_t)stream;
  ACCX_DISPATCH_T(dtype, {
    if (grad_dtype == ACCX_F32)
      launch_k(dice_bce_bwd_kernel<T, float>, grid, 256, 0, st, B, N, (const T*)logit, truth, sums, dice_w, bce_w, gscale,
                                                          (float*)dlogit);
    else if (grad_dtype == ACCX_BF16)
      launch_k(dice_bce_bwd_kernel<T, bf16>, grid, 256, 0, st, B, N, (const T*)logit, truth, sums, dice_w, bce_w, gscale,
                                                         (bf16*)dlogit);
    else {
      set_error("dice_bce_bwd: unsupported gradient dtype %d", grad_dtype);
      return ACCX_ERR_INVALID;
    }
  });
  return check_launch("dice_bce_bwd");
}

int accx_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* state,
                   float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  ACCX_REQUIRE(n > 0 && n % 4 == 0 && param && grad && exp_avg && exp_avg_sq && state,
               "adam_step: bad arguments (n = %lld must be a positive multiple of 4)", (long long)n);
  ACCX_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq),
               "adam_step: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(adam_tick_kernel, 1, 1, 0, st, state);
  const int64_t n4 = n / 4;
  int grid = grid_x_for(n4, 256 * 2, 148 * 8);
  launch_k(adam_flat_kernel, grid, 256, 0, st, n4, (float4*)param, (const float4*)grad, (float4*)exp_avg, (float4*)exp_avg_sq,
                                         state, lr, beta1, beta2, eps, weight_decay, grad_scale);
  return check_launch("adam_step");
}


int accx_upshuffle(int dtype, int forward, int B, int H, int W, int Co, void* temp, const float* bias, void* out,
                   int64_t ld_out, float* dbias, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && Co > 0 && Co % 2 == 0 && temp && out && ld_out >= Co,
               "upshuffle: bad arguments (Co = %d must be even)", Co);
  ACCX_REQUIRE(aligned16(temp) && aligned16(out) && ld_out % 2 == 0, "upshuffle: buffers must be 16-byte aligned");
  const int pairs = Co / 2;
  const int tx = pairs < 128 ? pairs : 128;
  int ty = 256 / tx;
  if (ty < 1) ty = 1;
  const int64_t P = (int64_t)B * H * W;
  // the backward variant also reduces the bias gradient: few long-lived blocks (one atomic per block and channel)
  dim3 block(tx, ty), grid(grid_x_for(P, ty * 2, forward || !dbias ? 148 * 16 : 148 * 4), (pairs + tx - 1) / tx);
  const size_t sm = (forward || !dbias) ? 0 : (size_t)tx * ty * 2 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (!forward && dbias && det_on()) grid.x = 1;       // deterministic mode: one contribution per bias gradient
  ACCX_DISPATCH_T(dtype, {
    if (forward)
      launch_k(upshuffle_kernel<T, true>, grid, block, sm, st, B, H, W, Co, (T*)temp, bias, (T*)out, ld_out, (float*)nullptr);
    else
      launch_k(upshuffle_kernel<T, false>, grid, block, sm, st, B, H, W, Co, (T*)temp, bias, (T*)out, ld_out, dbias);
  });
  return check_launch("upshuffle");
}

int accx_copy_cols(int dtype, int64_t P, int C, const void* src, int64_t ld_src, void* dst, int64_t ld_dst,
                   void* stream) {
  ACCX_REQUIRE(P > 0 && C > 0 && src && dst && ld_src >= C && ld_dst >= C, "copy_cols: bad arguments");
  ACCX_DISPATCH_T(dtype, {
    const int esz = (int)sizeof(T);
    const bool al = aligned16(src) && aligned16(dst) && (ld_src * esz) % 16 == 0 && (ld_dst * esz) % 16 == 0;
    Lanes l = make_lanes(C, DT<T>::VEC, al);
    dim3 block(l.tx, l.ty), grid(grid_x_for(P, l.ty * 4, 148 * 16), l.gy);
    ACCX_DISPATCH_VEC(l, {
      launch_k(copy_cols_kernel<T, VEC>, grid, block, 0, (cudaStream_t)stream, P, C, (const T*)src, ld_src, (T*)dst, ld_dst);
    });
  });
  return check_launch("copy_cols");
}

}  // extern "C"
