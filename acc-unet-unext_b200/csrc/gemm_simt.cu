// Pointwise contraction on CUDA cores (fp32 FMA), any shape, any number of lazy / shifted
// operands with strided weight views.  This is the exact-fp32 path (parity mode, odd channel
// counts such as C=3/9/45) and the reference implementation for the tcgen05 kernel in
// gemm_tc.cu, which takes over the large bf16 shapes.
//
//   Y[p, n] = sum_op sum_k act(A_op[p', k]*s[k] + t[k]) * W_op[n, k] + bias[n] + sum_j up(add_j)[p, n]
//
// Tiling: 128 pixels x BN outputs per CTA, BK = 16, 256 threads, 8 x TN register tile.
#include "common.cuh"

namespace accx {

constexpr int BM = 128, BK = 16, TM = 8, NTHREADS = 256;

struct PwParams {
  accx_operand_t op[ACCX_MAX_OPERANDS];
  int n_ops;
  int B, H, W, N;
  int64_t P;
  const float* bias;
  const float* add[ACCX_MAX_ADDENDS];
  int add_log2s[ACCX_MAX_ADDENDS];
  int n_add;
  void* y;
  int64_t ldy;
  float* stats;
  Det det;
};

template <typename T>
__device__ __forceinline__ void load8(const T* __restrict__ src, int k, int K, bool vec_ok, float (&v)[8]) {
  if (vec_ok) {
    if constexpr (sizeof(T) == 2) {
      ldv<T, 8>(src + k, v);
    } else {
      float a[4], b[4];
      ldv<T, 4>(src + k, a);
      ldv<T, 4>(src + k + 4, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i] = a[i]; v[4 + i] = b[i]; }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (k + i < K) ? to_f(src[k + i]) : 0.f;
  }
}

template <typename T, typename TO, int TN>
__global__ void __launch_bounds__(NTHREADS) pw_fwd_kernel(const __grid_constant__ PwParams prm) {
  pdl_sync();
  constexpr int BN = 16 * TN;
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Ws[BK][BN];
  __shared__ float red[2][16][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int HW = prm.H * prm.W;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // A-loader role: row r, k-half kh
  const int r = tid & (BM - 1), kh = tid >> 7;
  const int64_t prow = m0 + r;
  int ph = 0, pw = 0;
  if (prow < prm.P) {
    int rem = (int)(prow % HW);
    ph = rem / prm.W;
    pw = rem % prm.W;
  }
  // W-loader role
  constexpr int KPER = BK * BN / NTHREADS;
  const int wn = tid % BN, wk0 = (tid / BN) * KPER;

  for (int o = 0; o < prm.n_ops; ++o) {
    const accx_operand_t& op = prm.op[o];
    const int K = op.K;
    bool valid = prow < prm.P;
    int64_t psrc = prow;
    if (op.dy != 0 || op.dx != 0) {
      const int hh = ph + op.dy, ww = pw + op.dx;
      valid = valid && hh >= 0 && hh < prm.H && ww >= 0 && ww < prm.W;
      psrc = prow + (int64_t)op.dy * prm.W + op.dx;
    }
    const T* arow = (const T*)op.data + (valid ? psrc : 0) * op.ld;
    const bool vec_ok = (K % 8 == 0) && (op.ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(op.data) & 15) == 0);
    const bool wvalid = (n0 + wn) < prm.N;
    const float* wrow = op.w + (int64_t)(wvalid ? n0 + wn : 0) * op.w_ld;

    for (int k0 = 0; k0 < K; k0 += BK) {
      // ---- stage A (transform on the way in) ----
      {
        const int k = k0 + kh * 8;
        float v[8];
        if (valid && k < K) {
          load8<T>(arow, k, K, vec_ok, v);
          if (op.act != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (k + i < K) {
                float u = fmaf(v[i], __ldg(op.scale + k + i), __ldg(op.shift + k + i));
                v[i] = (op.act == 2) ? lrelu(u) : u;
              }
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) As[kh * 8 + i][r] = v[i];
      }
      // ---- stage W ----
#pragma unroll
      for (int i = 0; i < KPER; ++i) {
        const int k = k0 + wk0 + i;
        Ws[wk0 + i][wn] = (wvalid && k < K) ? __ldg(wrow + (int64_t)k * op.w_ks) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          float4 t = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
          a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Ws[k][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue: bias, upsample-adds, store, statistics ----
  float s1[TN], s2[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) s1[j] = s2[j] = 0.f;
  TO* yout = (TO*)prm.y;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t p = m0 + ty * TM + i;
    if (p >= prm.P) continue;
    int64_t addrow[ACCX_MAX_ADDENDS];
    if (prm.n_add > 0) {
      const int b = (int)(p / HW);
      const int rem = (int)(p % HW);
      const int h = rem / prm.W, w = rem % prm.W;
      for (int a = 0; a < prm.n_add; ++a) {
        const int l = prm.add_log2s[a];
        addrow[a] = (((int64_t)b * (prm.H >> l) + (h >> l)) * (prm.W >> l) + (w >> l)) * prm.N;
      }
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= prm.N) continue;
      float v = acc[i][j];
      if (prm.bias) v += __ldg(prm.bias + n);
      for (int a = 0; a < prm.n_add; ++a) v += __ldg(prm.add[a] + addrow[a] + n);
      s1[j] += v;
      s2[j] += v * v;
      yout[p * prm.ldy + n] = from_f<TO>(v);
    }
  }
  if (prm.stats) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      red[0][ty][tx * TN + j] = s1[j];
      red[1][ty][tx * TN + j] = s2[j];
    }
    __syncthreads();
    float s = 0.f;
    const int which = tid / BN, n = tid % BN;
    if (tid < 2 * BN) {
#pragma unroll
      for (int q = 0; q < 16; ++q) s += red[which][q][n];
    }
    if (prm.det.ws == nullptr) {
      if (tid < 2 * BN && n0 + n < prm.N) atomicAdd(prm.stats + (int64_t)which * prm.N + n0 + n, s);
    } else {
      // deterministic mode: the pixel tiles of one column block are added in tile order by the last one to arrive
      det_commit(prm.det, blockIdx.y, blockIdx.x, gridDim.x, 2 * BN,
                 [&](float* slot) { if (tid < 2 * BN) slot[tid] = s; },
                 [&](int i, float total) {
                   const int wh = i / BN, nn = i % BN;
                   if (n0 + nn < prm.N) atomicAdd(prm.stats + (int64_t)wh * prm.N + n0 + nn, total);
                 });
    }
  }
}

// ---------------------------------------------------------------------------------------
// dW[n, k] += sum_p dY[p, n] * value(p, k): 64 x 64 output tile per CTA, pixels split over
// gridDim.z, 16 pixels per smem stage, 4 x 4 register tile.
constexpr int WG_T = 64, WG_P = 16;

template <typename T, typename TG>
__global__ void __launch_bounds__(NTHREADS) pw_wgrad_kernel(accx_operand_t op, int B, int H, int W, int N, int64_t P,
                                                            const TG* __restrict__ dy, int64_t ldy, float* dw,
                                                            int64_t p_per_split) {
  pdl_sync();
  __shared__ __align__(16) float Gs[WG_P][WG_T];   // dY tile  [pixel][n]
  __shared__ __align__(16) float Xs[WG_P][WG_T];   // A tile   [pixel][k]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;          // tx -> k quad, ty -> n quad
  const int k0 = blockIdx.x * WG_T, n0 = blockIdx.y * WG_T;
  const int64_t pbeg = (int64_t)blockIdx.z * p_per_split;
  const int64_t pend = (pbeg + p_per_split < P) ? pbeg + p_per_split : P;
  const int HW = H * W;
  const int K = op.K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader role: pixel lp (0..15), 4 consecutive columns lc..lc+3
  const int lp = tid >> 4, lc = (tid & 15) * 4;
  float sc[4], sh[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + lc + i;
    sc[i] = (op.act != 0 && k < K) ? op.scale[k] : 1.f;
    sh[i] = (op.act != 0 && k < K) ? op.shift[k] : 0.f;
  }

  for (int64_t pb = pbeg; pb < pend; pb += WG_P) {
    const int64_t p = pb + lp;
    const bool pv = p < pend;
    // dY
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + lc + i;
      Gs[lp][lc + i] = (pv && n < N) ? to_f(dy[p * ldy + n]) : 0.f;
    }
    // A (shifted, lazy)
    bool valid = pv;
    int64_t psrc = p;
    if (pv && (op.dy != 0 || op.dx != 0)) {
      const int rem = (int)(p % HW);
      const int hh = rem / W + op.dy, ww = rem % W + op.dx;
      valid = hh >= 0 && hh < H && ww >= 0 && ww < W;
      psrc = p + (int64_t)op.dy * W + op.dx;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lc + i;
      float v = 0.f;
      if (valid && k < K) {
        v = to_f(((const T*)op.data)[psrc * op.ld + k]);
        if (op.act != 0) {
          v = fmaf(v, sc[i], sh[i]);
          if (op.act == 2) v = lrelu(v);
        }
      }
      Xs[lp][lc + i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < WG_P; ++q) {
      const float4 g = *reinterpret_cast<const float4*>(&Gs[q][ty * 4]);
      const float4 x = *reinterpret_cast<const float4*>(&Xs[q][tx * 4]);
      const float gv[4] = {g.x, g.y, g.z, g.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], xv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dw + (int64_t)n * op.w_ld + (int64_t)k * op.w_ks, acc[i][j]);
    }
  }
}

template <typename T, typename TO>
static int launch_pw(PwParams& prm, cudaStream_t st) {
  const int gx = (int)((prm.P + BM - 1) / BM);
  {
    const int bn = prm.N <= 32 ? 32 : ((prm.N <= 64 || prm.N % 128 != 0) ? 64 : 128);
    const int gy = (prm.N + bn - 1) / bn;
    if (!det_handle(prm.stats ? (int64_t)gx * gy * 2 * bn : 0, gy, st, prm.det)) return ACCX_ERR_INVALID;
  }
  if (prm.N <= 32) {
    dim3 grid(gx, (prm.N + 31) / 32);
    launch_k(pw_fwd_kernel<T, TO, 2>, grid, NTHREADS, 0, st, prm);
  } else if (prm.N <= 64 || prm.N % 128 != 0) {
    dim3 grid(gx, (prm.N + 63) / 64);
    launch_k(pw_fwd_kernel<T, TO, 4>, grid, NTHREADS, 0, st, prm);
  } else {
    dim3 grid(gx, (prm.N + 127) / 128);
    launch_k(pw_fwd_kernel<T, TO, 8>, grid, NTHREADS, 0, st, prm);
  }
  return check_launch("pw_fwd");
}

}  // namespace accx

using namespace accx;

extern "C" {

int accx_pw_fwd(int dtype, int out_dtype, int B, int H, int W, int N, const accx_operand_t* ops, int n_ops,
                const float* bias, const float* const* add, const int* add_log2s, int n_add, void* y, int64_t ldy,
                float* stats, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0 && ops && y, "pw_fwd: bad arguments");
  ACCX_REQUIRE(n_ops >= 1 && n_ops <= ACCX_MAX_OPERANDS, "pw_fwd: n_ops %d out of range", n_ops);
  ACCX_REQUIRE(n_add >= 0 && n_add <= ACCX_MAX_ADDENDS, "pw_fwd: n_add %d out of range", n_add);
  ACCX_REQUIRE(ldy >= N, "pw_fwd: ldy %lld < N %d", (long long)ldy, N);
  PwParams prm;
  prm.n_ops = n_ops;
  for (int i = 0; i < n_ops; ++i) {
    prm.op[i] = ops[i];
    ACCX_REQUIRE(ops[i].data && ops[i].w && ops[i].K > 0 && ops[i].ld >= ops[i].K, "pw_fwd: operand %d malformed", i);
    ACCX_REQUIRE(ops[i].act == 0 || (ops[i].scale && ops[i].shift), "pw_fwd: operand %d act without scale/shift", i);
  }
  prm.B = B; prm.H = H; prm.W = W; prm.N = N;
  prm.P = (int64_t)B * H * W;
  prm.bias = bias;
  prm.n_add = n_add;
  for (int i = 0; i < n_add; ++i) {
    prm.add[i] = add[i];
    prm.add_log2s[i] = add_log2s[i];
    ACCX_REQUIRE(add[i] && add_log2s[i] >= 0 && (H >> add_log2s[i]) << add_log2s[i] == H &&
                     (W >> add_log2s[i]) << add_log2s[i] == W,
                 "pw_fwd: addend %d does not tile %dx%d", i, H, W);
  }
  prm.y = y; prm.ldy = ldy; prm.stats = stats;
  cudaStream_t st = (cudaStream_t)stream;
  {   // tiny channel counts (ACC-UNet's first block, C = 3): whole pixels per thread instead of GEMM tiles
    int k_total = 0;
    bool plain = !(dtype == ACCX_F32 && out_dtype == ACCX_BF16);
    for (int i = 0; i < n_ops; ++i) {
      k_total += ops[i].K;
      plain = plain && ops[i].dy == 0 && ops[i].dx == 0;
    }
    if (plain && narrow_fwd_ok(k_total, N)) {
      NarrowParams np;
      for (int i = 0; i < n_ops; ++i) np.op[i] = ops[i];
      np.n_ops = n_ops; np.k_total = k_total;
      np.B = B; np.H = H; np.W = W; np.N = N; np.P = prm.P;
      np.bias = bias; np.n_add = n_add;
      for (int i = 0; i < ACCX_MAX_ADDENDS; ++i) { np.add[i] = i < n_add ? add[i] : nullptr; np.add_log2s[i] = i < n_add ? add_log2s[i] : 0; }
      np.y = y; np.ldy = ldy; np.stats = stats;
      return pw_fwd_narrow(dtype, out_dtype, np, st);
    }
  }
  if (dtype == ACCX_F32 && out_dtype == ACCX_F32) return launch_pw<float, float>(prm, st);
  if (dtype == ACCX_BF16 && out_dtype == ACCX_BF16) return launch_pw<bf16, bf16>(prm, st);
  if (dtype == ACCX_BF16 && out_dtype == ACCX_F32) return launch_pw<bf16, float>(prm, st);
  if (dtype == ACCX_F32 && out_dtype == ACCX_BF16) return launch_pw<float, bf16>(prm, st);
  set_error("pw_fwd: bad dtypes %d %d", dtype, out_dtype);
  return ACCX_ERR_INVALID;
}

int accx_pw_wgrad(int dtype, int B, int H, int W, int N, const accx_operand_t* op, float* dw, const void* dy,
                  int64_t ldy, int dy_f32, void* stream) {
  ACCX_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0 && op && op->data && dw && dy, "pw_wgrad: bad arguments");
  ACCX_REQUIRE(op->act == 0 || (op->scale && op->shift), "pw_wgrad: act without scale/shift");
  const int64_t P = (int64_t)B * H * W;
  if (op->dy == 0 && op->dx == 0 && narrow_wgrad_ok(op->K, N) && (dtype == ACCX_F32 || dtype == ACCX_BF16))
    return pw_wgrad_narrow(dtype, dy_f32, *op, N, P, dy, ldy, dw, (cudaStream_t)stream);
  const int gx = (op->K + WG_T - 1) / WG_T, gy = (N + WG_T - 1) / WG_T;
  int64_t splits = (148 * 4 + gx * gy - 1) / (gx * gy);
  if (det_on()) splits = 1;          // deterministic mode: one contribution per dW element, pixels in order
  const int64_t max_splits = (P + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t per = (P + splits - 1) / splits;
  per = (per + WG_P - 1) / WG_P * WG_P;
  splits = (P + per - 1) / per;
  dim3 grid(gx, gy, (unsigned)splits);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == ACCX_F32) {
    launch_k(pw_wgrad_kernel<float, float>, grid, NTHREADS, 0, st, *op, B, H, W, N, P, (const float*)dy, ldy, dw, per);
  } else if (dtype == ACCX_BF16 && dy_f32) {
    launch_k(pw_wgrad_kernel<bf16, float>, grid, NTHREADS, 0, st, *op, B, H, W, N, P, (const float*)dy, ldy, dw, per);
  } else if (dtype == ACCX_BF16) {
    launch_k(pw_wgrad_kernel<bf16, bf16>, grid, NTHREADS, 0, st, *op, B, H, W, N, P, (const bf16*)dy, ldy, dw, per);
  } else {
    set_error("pw_wgrad: bad dtype %d", dtype);
    return ACCX_ERR_INVALID;
  }
  return check_launch("pw_wgrad");
}

}  // extern "C"
