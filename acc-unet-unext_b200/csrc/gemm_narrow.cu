// Pointwise contraction and its weight gradient for TINY channel counts (sum K <= 32, N <= 32, no shifted taps):
// the first HANCBlock of ACC-UNet works on the 3-channel image (cnv11: C = 3, 3C = 9, K = 45 split into 9 + 18 + 18;
// /root/reference/ACC_UNet/ACC_UNet.py:554).  A 128-pixel GEMM tile is >90 % padding there; these kernels give
// every thread whole pixels instead: the K inputs of a pixel live in registers, the (N x K) weights in shared
// memory (broadcast reads), outputs / statistics / weight-gradient partial sums in registers.
// Same contract as accx_pw_fwd / accx_pw_wgrad (lazy operands, strided weight views, bias, upsample-adds, stats).
// HBM-bound: reads P * sum K, writes P * N elements.
#include "common.cuh"

namespace accx {

constexpr int NAR_THREADS = 256;

// one warp-shuffle tree per value, then shared memory across the warps of the block, then atomics
template <int NV>
__device__ __forceinline__ void narrow_block_reduce_atomic(float (&v)[NV], float* red, float* out, const int* index,
                                                           int n_valid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp * NV + i] = x;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w * NV + i];
    if (i < n_valid && index[i] >= 0) atomicAdd(out + index[i], s);
  }
  __syncthreads();
}

template <typename T, typename TO, int KT, int NT>
__global__ void __launch_bounds__(NAR_THREADS) pw_fwd_narrow_kernel(const __grid_constant__ NarrowParams prm) {
  pdl_sync();
  __shared__ float Ws[NT][KT];                 // weights, zero padded
  __shared__ float sc[KT], sh[KT], slope[KT];  // pending affine + LeakyReLU slope per concatenated k
  __shared__ const T* kptr[KT];                // operand base (+ column) per concatenated k
  __shared__ int64_t kld[KT];
  __shared__ float bs[NT];
  __shared__ float red[(NAR_THREADS / 32) * 2 * NT];
  __shared__ int sidx[2 * NT];
  const int N = prm.N, Kt = prm.k_total;
  for (int i = threadIdx.x; i < NT * KT; i += blockDim.x) Ws[i / KT][i % KT] = 0.f;
  __syncthreads();
  if (threadIdx.x < KT) {
    const int kk = threadIdx.x;
    sc[kk] = 1.f; sh[kk] = 0.f; slope[kk] = 1.f; kptr[kk] = nullptr; kld[kk] = 0;
    int o = 0, k = kk;
    while (o < prm.n_ops && k >= prm.op[o].K) { k -= prm.op[o].K; ++o; }
    if (o < prm.n_ops) {
      const accx_operand_t& op = prm.op[o];
      if (op.act != 0) { sc[kk] = op.scale[k]; sh[kk] = op.shift[k]; }
      if (op.act == 2) slope[kk] = ACCX_LRELU;
      kptr[kk] = (const T*)op.data + k;
      kld[kk] = op.ld;
      for (int n = 0; n < N; ++n) Ws[n][kk] = op.w[(int64_t)n * op.w_ld + (int64_t)k * op.w_ks];
    }
  }
  if (threadIdx.x < NT) bs[threadIdx.x] = (prm.bias && threadIdx.x < N) ? prm.bias[threadIdx.x] : 0.f;
  if (threadIdx.x < 2 * NT) {
    const int n = threadIdx.x % NT;
    sidx[threadIdx.x] = n < N ? (threadIdx.x < NT ? n : N + n) : -1;
  }
  __syncthreads();
  float s1[NT], s2[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) s1[n] = s2[n] = 0.f;
  const int HWp = prm.H * prm.W;
  const bool vec_in = prm.n_ops == 1 && Kt % (16 / (int)sizeof(T)) == 0 && (prm.op[0].ld * sizeof(T)) % 16 == 0 &&
                      (reinterpret_cast<uintptr_t>(prm.op[0].data) & 15) == 0;
  const bool vec_out = N % (16 / (int)sizeof(TO)) == 0 && (prm.ldy * sizeof(TO)) % 16 == 0 &&
                       (reinterpret_cast<uintptr_t>(prm.y) & 15) == 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < prm.P; p += (int64_t)gridDim.x * blockDim.x) {
    float a[KT];
    constexpr int IPV = 16 / sizeof(T);                  // input elements per 16-byte load
    if (KT % IPV == 0 && vec_in) {                       // one operand, contiguous channels: 16-byte loads
      const T* arow = kptr[0] + p * kld[0];
#pragma unroll
      for (int k0 = 0; k0 < KT; k0 += IPV) {
        float t[IPV];
#pragma unroll
        for (int e = 0; e < IPV; ++e) t[e] = 0.f;
        if (k0 < Kt) ldv<T, IPV>(arow + k0, t);
#pragma unroll
        for (int e = 0; e < IPV; ++e) {
          float v = fmaf(t[e], sc[k0 + e], sh[k0 + e]);
          a[k0 + e] = (k0 + e < Kt) ? fmaxf(v, v * slope[k0 + e]) : 0.f;
        }
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        float v = 0.f;
        if (kk < Kt) {
          v = to_f(kptr[kk][p * kld[kk]]);
          v = fmaf(v, sc[kk], sh[kk]);
          v = fmaxf(v, v * slope[kk]);
        }
        a[kk] = v;
      }
    }
    float y[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      float acc = bs[n];
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) acc = fmaf(a[kk], Ws[n][kk], acc);
      y[n] = acc;
    }
    if (prm.n_add > 0) {
      const int b = (int)(p / HWp), rem = (int)(p % HWp);
      const int h = rem / prm.W, w = rem % prm.W;
      for (int j = 0; j < prm.n_add; ++j) {
        const int l = prm.add_log2s[j];
        const float* ar = prm.add[j] + (((int64_t)b * (prm.H >> l) + (h >> l)) * (prm.W >> l) + (w >> l)) * N;
#pragma unroll
        for (int n = 0; n < NT; ++n)
          if (n < N) y[n] += __ldg(ar + n);
      }
    }
    TO* dst = (TO*)prm.y + p * prm.ldy;
    constexpr int EPV = 16 / sizeof(TO);                 // output elements per 16-byte store
    if constexpr (NT % EPV == 0) {
      if (vec_out) {                                     // a pixel's outputs as 16-byte stores (scalar 2-byte stores at a
#pragma unroll                                           // 64-byte pitch cost one LSU transaction per element)
        for (int n0 = 0; n0 < NT; n0 += EPV) {
          if (n0 < N) {
            float t[EPV];
#pragma unroll
            for (int e = 0; e < EPV; ++e) t[e] = y[n0 + e];
            stv<TO, EPV>(dst + n0, t);
          }
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (n < N) {
        if (!(NT % EPV == 0 && vec_out)) dst[n] = from_f<TO>(y[n]);
        s1[n] += y[n];
        s2[n] = fmaf(y[n], y[n], s2[n]);
      }
    }
  }
  if (prm.stats) {
    float v[2 * NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) { v[n] = s1[n]; v[NT + n] = s2[n]; }
    narrow_block_reduce_atomic<2 * NT>(v, red, prm.stats, sidx, 2 * NT);
  }
}

// dW[n*w_ld + k*w_ks] += sum_p dY[p, n] * value(p, k);   N <= NT, K <= KT, NT * KT <= 128
template <typename T, typename TG, int KT, int NT>
__global__ void __launch_bounds__(NAR_THREADS) pw_wgrad_narrow_kernel(accx_operand_t op, int N, int64_t P,
                                                                      const TG* __restrict__ dy, int64_t ldy,
                                                                      float* dw) {
  pdl_sync();
  __shared__ float sc[KT], sh[KT];
  __shared__ float red[(NAR_THREADS / 32) * NT * KT];
  __shared__ int widx[NT * KT];
  const int K = op.K;
  if (threadIdx.x < KT) {
    sc[threadIdx.x] = (op.act != 0 && threadIdx.x < K) ? op.scale[threadIdx.x] : 1.f;
    sh[threadIdx.x] = (op.act != 0 && threadIdx.x < K) ? op.shift[threadIdx.x] : 0.f;
  }
  for (int i = threadIdx.x; i < NT * KT; i += blockDim.x) {
    const int n = i / KT, k = i % KT;
    widx[i] = (n < N && k < K) ? (int)(n * op.w_ld + k * op.w_ks) : -1;
  }
  __syncthreads();
  const float slope = op.act == 2 ? ACCX_LRELU : 1.f;
  const bool vec_a = K % (16 / (int)sizeof(T)) == 0 && (op.ld * sizeof(T)) % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(op.data) & 15) == 0 && op.dy == 0 && op.dx == 0;
  const bool vec_g = N % (16 / (int)sizeof(TG)) == 0 && (ldy * sizeof(TG)) % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
  float acc[NT * KT];
#pragma unroll
  for (int i = 0; i < NT * KT; ++i) acc[i] = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    float a[KT], g[NT];
    const T* arow = (const T*)op.data + p * op.ld;
    constexpr int IPV = 16 / sizeof(T);
    if (KT % IPV == 0 && vec_a) {
#pragma unroll
      for (int k0 = 0; k0 < KT; k0 += IPV) {
        float t[IPV];
#pragma unroll
        for (int e = 0; e < IPV; ++e) t[e] = 0.f;
        if (k0 < K) ldv<T, IPV>(arow + k0, t);
#pragma unroll
        for (int e = 0; e < IPV; ++e) {
          float v = fmaf(t[e], sc[k0 + e], sh[k0 + e]);
          a[k0 + e] = (k0 + e < K) ? fmaxf(v, v * slope) : 0.f;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        float v = 0.f;
        if (k < K) {
          v = fmaf(to_f(arow[k]), sc[k], sh[k]);
          v = fmaxf(v, v * slope);
        }
        a[k] = v;
      }
    }
    const TG* grow = dy + p * ldy;
    constexpr int GPV = 16 / sizeof(TG);                 // dY elements per 16-byte load
    if (NT % GPV == 0 && vec_g) {
#pragma unroll
      for (int n0 = 0; n0 < NT; n0 += GPV) {
        float t[GPV];
#pragma unroll
        for (int e = 0; e < GPV; ++e) t[e] = 0.f;
        if (n0 < N) ldv<TG, GPV>(grow + n0, t);
#pragma unroll
        for (int e = 0; e < GPV; ++e) g[n0 + e] = t[e];
      }
    } else {
#pragma unroll
      for (int n = 0; n < NT; ++n) g[n] = n < N ? to_f(grow[n]) : 0.f;
    }
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int k = 0; k < KT; ++k) acc[n * KT + k] = fmaf(g[n], a[k], acc[n * KT + k]);
  }
  narrow_block_reduce_atomic<NT * KT>(acc, red, dw, widx, NT * KT);
}

template <typename T, typename TO>
static int launch_fwd_narrow(const NarrowParams& prm, cudaStream_t st) {
  const int64_t blocks64 = (prm.P + NAR_THREADS - 1) / NAR_THREADS;
  const int cap = prm.stats ? 148 * 2 : 148 * 8;          // reducing variant: few blocks (same-address atomics)
  int blocks = (int)(blocks64 < cap ? blocks64 : cap);
  if (prm.stats && det_on()) blocks = 1;                  // deterministic mode: one contribution per statistic
  const int K = prm.k_total, N = prm.N;
#define ACCX_NARROW_FWD(KT, NT)                                                        \
  if (K <= KT && N <= NT) {                                                            \
    launch_k(pw_fwd_narrow_kernel<T, TO, KT, NT>, blocks, NAR_THREADS, 0, st, prm);          \
    return check_launch("pw_fwd(narrow)");                                             \
  }
  ACCX_NARROW_FWD(4, 12)
  ACCX_NARROW_FWD(12, 4)
  ACCX_NARROW_FWD(20, 4)
  ACCX_NARROW_FWD(4, 32)
  ACCX_NARROW_FWD(32, 4)
  ACCX_NARROW_FWD(12, 12)
  ACCX_NARROW_FWD(32, 12)
  ACCX_NARROW_FWD(12, 32)
#undef ACCX_NARROW_FWD
  set_error("pw_fwd(narrow): no instantiation for K=%d N=%d", K, N);
  return ACCX_ERR_INVALID;
}

// shapes the narrow kernels are instantiated for
bool narrow_fwd_ok(int k_total, int N) {
  return (k_total <= 32 && N <= 12) || (k_total <= 12 && N <= 32);
}
bool narrow_wgrad_ok(int K, int N) {
  return (K <= 4 && N <= 32) || (K <= 32 && N <= 4) || (K <= 12 && N <= 12);
}

int pw_fwd_narrow(int dtype, int out_dtype, const NarrowParams& prm, cudaStream_t st) {
  if (dtype == ACCX_BF16 && out_dtype == ACCX_BF16) return launch_fwd_narrow<bf16, bf16>(prm, st);
  if (dtype == ACCX_BF16 && out_dtype == ACCX_F32) return launch_fwd_narrow<bf16, float>(prm, st);
  if (dtype == ACCX_F32 && out_dtype == ACCX_F32) return launch_fwd_narrow<float, float>(prm, st);
  set_error("pw_fwd(narrow): unsupported dtypes %d %d", dtype, out_dtype);
  return ACCX_ERR_INVALID;
}

template <typename T, typename TG>
static int launch_wgrad_narrow(const accx_operand_t& op, int N, int64_t P, const void* dy, int64_t ldy, float* dw,
                               cudaStream_t st) {
  const int64_t blocks64 = (P + NAR_THREADS - 1) / NAR_THREADS;
  int blocks = (int)(blocks64 < 148 * 2 ? blocks64 : 148 * 2);
  if (det_on()) blocks = 1;                               // deterministic mode: one contribution per dW element
  const int K = op.K;
#define ACCX_NARROW_WG(KT, NT)                                                                               \
  if (K <= KT && N <= NT) {                                                                                  \
    launch_k(pw_wgrad_narrow_kernel<T, TG, KT, NT>, blocks, NAR_THREADS, 0, st, op, N, P, (const TG*)dy, ldy, dw); \
    return check_launch("pw_wgrad(narrow)");                                                                 \
  }
  ACCX_NARROW_WG(4, 12)
  ACCX_NARROW_WG(12, 4)
  ACCX_NARROW_WG(4, 32)
  ACCX_NARROW_WG(32, 4)
  ACCX_NARROW_WG(12, 12)
#undef ACCX_NARROW_WG
  set_error("pw_wgrad(narrow): no instantiation for K=%d N=%d", K, N);
  return ACCX_ERR_INVALID;
}

int pw_wgrad_narrow(int dtype, int dy_f32, const accx_operand_t& op, int N, int64_t P, const void* dy, int64_t ldy,
                    float* dw, cudaStream_t st) {
  if (dtype == ACCX_F32) return launch_wgrad_narrow<float, float>(op, N, P, dy, ldy, dw, st);
  if (dtype == ACCX_BF16 && dy_f32) return launch_wgrad_narrow<bf16, float>(op, N, P, dy, ldy, dw, st);
  if (dtype == ACCX_BF16) return launch_wgrad_narrow<bf16, bf16>(op, N, P, dy, ldy, dw, st);
  set_error("pw_wgrad(narrow): unsupported dtype %d", dtype);
  return ACCX_ERR_INVALID;
}

}  // namespace accx
