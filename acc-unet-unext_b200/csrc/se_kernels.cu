// ChannelSELayer (/root/reference/ACC_UNet/ACC_UNet.py:37-49):
//   m = mean_hw(a); g = sigmoid(W2 lrelu(W1 m + b1) + b2); z = a*g; out = lrelu(BN(z)).
// One read pass (squeeze) yields per-(b,c) sum and sum of squares of the lazy input a; the gate
// AND the batch statistics of z follow from them (mean_c = sum_b g*S1/n, E[z^2]_c = sum_b g^2*S2/n),
// so the layer costs: read a (squeeze) + read a, write out (apply).  Backward mirrors it:
// one reduce pass, a tiny per-batch kernel, one apply pass.
#include "common.cuh"

namespace accx {

// grid.x = B * chunks; each block reduces a slice of one image
template <typename T, int VEC, int U>
__global__ void se_squeeze_kernel(int B, int HW, int C, int chunks, const T* __restrict__ x, const float* scale,
                                  const float* shift, int act, float* S, Det det) {
  pdl_sync();
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  const int c0 = active ? cv * VEC : 0;
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float acc[2][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (active) {
    const T* img = x + (int64_t)b * HW * C + c0;
    RawVec<T, VEC> rx[U];
    pixel_loop<U>(chunk * blockDim.y + threadIdx.y, HW, chunks * blockDim.y,
        [&](int u, int64_t p) { rx[u].load(img + p * C); },
        [&](int u, int64_t p) {
          float v[VEC];
          rx[u].unpack(v);
          lz.apply(v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) { acc[0][i] += v[i]; acc[1][i] += v[i] * v[i]; }
        });
  }
  reduce_lanes_atomic<2, VEC>(acc, smem, S + (int64_t)b * C, (int64_t)B * C, C, det, b * gridDim.y + blockIdx.y, chunk,
                              chunks);
}

// The two per-batch kernels are chains of dependent L2 round trips over tiny matrices (one image per block); wider
// blocks (1024 threads) and deeper unrolling were measured and did not help (B200, round 1).
constexpr int SE_GATE_THREADS = 256;

// grid = B blocks; the last block to finish derives the BatchNorm affine.
__global__ void __launch_bounds__(SE_GATE_THREADS) se_gate_kernel(int B, int C, int Cr, double HW, const float* __restrict__ S,
                               const float* __restrict__ w1, const float* __restrict__ b1,
                               const float* __restrict__ w2, const float* __restrict__ b2,
                               const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                               float momentum, int training, float* running_mean, float* running_var, int64_t* nbt,
                               float* gate, float* hidden, float* scale, float* shift, float* mean_o, float* rstd_o,
                               unsigned int* counter) {
  pdl_sync();
  extern __shared__ float sm[];
  float* m = sm;            // [C]
  float* h = sm + C;        // [Cr]
  __shared__ bool is_last;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float inv_hw = (float)(1.0 / HW);
  for (int c = tid; c < C; c += nt) m[c] = S[(int64_t)b * C + c] * inv_hw;
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int j = warp; j < Cr; j += nw) {
    float a = 0.f;
#pragma unroll 4
    for (int c = lane; c < C; c += 32) a = fmaf(__ldg(w1 + (int64_t)j * C + c), m[c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      a += b1[j];
      hidden[(int64_t)b * Cr + j] = a;
      h[j] = lrelu(a);
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += nt) {
    float a = b2[c];
    const float* wr = w2 + (int64_t)c * Cr;
    if ((Cr & 3) == 0) {        // row of w2 as independent 16-byte loads (the scalar loop is one dependent chain)
      float a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
      for (int j = 0; j < Cr; j += 4) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wr + j));
        a = fmaf(w.x, h[j], a);
        a1 = fmaf(w.y, h[j + 1], a1);
        a2 = fmaf(w.z, h[j + 2], a2);
        a3 = fmaf(w.w, h[j + 3], a3);
      }
      a += a1 + a2 + a3;
    } else {
      for (int j = 0; j < Cr; ++j) a = fmaf(wr[j], h[j], a);
    }
    gate[(int64_t)b * C + c] = 1.f / (1.f + __expf(-a));
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int prev = atomicAdd(counter, 1u);
    is_last = (prev == (unsigned)(B - 1));
    if (is_last) *counter = 0;   // re-arm for the next launch (graph replays)
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const double n = HW * (double)B;
  if (tid == 0 && training && nbt) *nbt += 1;
  for (int c = tid; c < C; c += nt) {
    float mean, var;
    if (training) {
      double s1 = 0, s2 = 0;
#pragma unroll 4
      for (int bb = 0; bb < B; ++bb) {
        float g = __ldcg(gate + (int64_t)bb * C + c);
        s1 += (double)g * S[(int64_t)bb * C + c];
        s2 += (double)g * g * S[(int64_t)B * C + (int64_t)bb * C + c];
      }
      double mu = s1 / n, v = s2 / n - mu * mu;
      if (v < 0) v = 0;
      mean = (float)mu;
      var = (float)v;
      if (running_mean) {
        double unb = n > 1 ? v * n / (n - 1) : v;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
      }
    } else {
      mean = running_mean[c];
      var = running_var[c];
    }
    float rstd = rsqrtf(var + eps);
    float s = gamma[c] * rstd;
    scale[c] = s;
    shift[c] = beta[c] - mean * s;
    if (mean_o) mean_o[c] = mean;
    if (rstd_o) rstd_o[c] = rstd;
  }
}

// out = mixf( lrelu(a*gate[b,c]*se_scale[c] + se_shift[c]) , residual )
// grid.x = B * chunks: a block works inside ONE image, so the gate (folded with the BN scale) lives in registers
template <typename T, int VEC, int U>
__global__ void se_apply_kernel(int B, int HW, int C, int chunks, const T* __restrict__ x, const float* scale,
                                const float* shift, int act, const float* __restrict__ gate, const float* se_scale,
                                const float* se_shift, const T* __restrict__ residual, const float* mix,
                                T* __restrict__ out, float* stats, Det det) {
  pdl_sync();
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  const int c0 = active ? cv * VEC : 0;
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float gs[VEC], st[VEC];
  ldf<VEC>(se_scale + c0, gs);
  ldf<VEC>(se_shift + c0, st);
  {
    float g[VEC];
    ldf<VEC>(gate + (int64_t)b * C + c0, g);
#pragma unroll
    for (int i = 0; i < VEC; ++i) gs[i] *= g[i];
  }
  const float mx = mix ? *mix : 1.f, rx = mix ? 1.f - mx : 1.f;
  float acc[2][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (active) {
    const int64_t off = (int64_t)b * HW * C + c0;
    RawVec<T, VEC> rv[U], rr[U];
    pixel_loop<U>(chunk * blockDim.y + threadIdx.y, HW, chunks * blockDim.y,
        [&](int u, int64_t p) {
          rv[u].load(x + off + p * C);
          if (residual) rr[u].load(residual + off + p * C);
        },
        [&](int u, int64_t p) {
          float v[VEC];
          rv[u].unpack(v);
          lz.apply(v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) v[i] = lrelu(fmaf(v[i], gs[i], st[i]));
          if (residual) {
            float r[VEC];
            rr[u].unpack(r);
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[i] = v[i] * mx + r[i] * rx;
          }
#pragma unroll
          for (int i = 0; i < VEC; ++i) { acc[0][i] += v[i]; acc[1][i] += v[i] * v[i]; }
          stv<T, VEC>(out + off + p * C, v);
        });
  }
  if (stats) reduce_lanes_atomic<2, VEC>(acc, smem, stats, C, C, det, blockIdx.y, blockIdx.x, gridDim.x);
}

// G[0,b,c] += sum_hw g', G[1,b,c] += sum_hw g'*a;  g' = dout*mix*lrelu'(v);  dmix += sum dout*(v_act - r)
template <typename T, int VEC, int U>
__global__ void se_bwd_reduce_kernel(int B, int HW, int C, int chunks, const T* __restrict__ x, const float* scale,
                                     const float* shift, int act, const float* __restrict__ gate,
                                     const float* se_scale, const float* se_shift, const T* __restrict__ dout,
                                     const float* mix, const T* __restrict__ residual, float* dmix, float* G,
                                     Det det, float* dmix_part) {
  pdl_sync();
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  const int c0 = active ? cv * VEC : 0;
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float ss[VEC], st[VEC], g[VEC];
  ldf<VEC>(se_scale + c0, ss);
  ldf<VEC>(se_shift + c0, st);
  ldf<VEC>(gate + (int64_t)b * C + c0, g);
  const float mx = mix ? *mix : 1.f;
  float acc[2][VEC];
  float dm = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (active) {
    const int64_t off = (int64_t)b * HW * C + c0;
    RawVec<T, VEC> rv[U], rd[U], rr[U];
    pixel_loop<U>(chunk * blockDim.y + threadIdx.y, HW, chunks * blockDim.y,
        [&](int u, int64_t p) {
          rv[u].load(x + off + p * C);
          rd[u].load(dout + off + p * C);
          if (dmix) rr[u].load(residual + off + p * C);
        },
        [&](int u, int64_t p) {
          float v[VEC], d[VEC], r[VEC];
          rv[u].unpack(v);
          rd[u].unpack(d);
          if (dmix) rr[u].unpack(r);
          lz.apply(v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            float uu = fmaf(v[i] * g[i], ss[i], st[i]);
            if (dmix) dm += d[i] * (lrelu(uu) - r[i]);
            float gp = d[i] * mx * (uu > 0.f ? 1.f : ACCX_LRELU);
            acc[0][i] += gp;
            acc[1][i] += gp * v[i];
          }
        });
  }
  reduce_lanes_atomic<2, VEC>(acc, smem, G + (int64_t)b * C, (int64_t)B * C, C, det, b * gridDim.y + blockIdx.y, chunk,
                              chunks);
  if (dmix) {   // block size need not be a multiple of 32: reduce through shared memory
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
    __syncthreads();
    smem[tid] = dm;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int i = 0; i < nth; ++i) s += smem[i];
      // deterministic mode: one partial per block, added in block order by se_dmix_fold_kernel
      if (dmix_part) dmix_part[blockIdx.y * gridDim.x + blockIdx.x] = s;
      else atomicAdd(dmix, s);
    }
  }
}

// One block per image.  Every block first derives the per-channel BN-backward constants
// (c1 = mean g', c2 = mean g'*zhat) from G over ALL images (cheap: B*C values), then runs the
// two tiny FC backward products for its image and emits the per-(b,c) apply coefficients.
__global__ void __launch_bounds__(SE_GATE_THREADS) se_bwd_gate_kernel(int B, int C, int Cr, double HW, const float* __restrict__ S,
                                   const float* __restrict__ G, const float* __restrict__ gate,
                                   const float* __restrict__ hidden, const float* __restrict__ w1,
                                   const float* __restrict__ w2, const float* __restrict__ gamma,
                                   const float* __restrict__ mean, const float* __restrict__ rstd, float* dw1,
                                   float* db1, float* dw2, float* db2, float* dgamma, float* dbeta, float* PQR,
                                   int training, int b_off) {
  pdl_sync();
  extern __shared__ float sm[];
  float* dpre2 = sm;            // [C]
  float* m = sm + C;            // [C]
  float* hact = sm + 2 * C;     // [Cr]
  float* dpre1 = hact + Cr;     // [Cr]
  const int b = blockIdx.x + b_off, tid = threadIdx.x, nt = blockDim.x;   // b_off: deterministic mode, one image per launch
  const double n = HW * (double)B;
  const float inv_hw = (float)(1.0 / HW);
  const int64_t BC = (int64_t)B * C;
  for (int j = tid; j < Cr; j += nt) hact[j] = lrelu(hidden[(int64_t)b * Cr + j]);
  for (int c = tid; c < C; c += nt) {
    const float mu = mean[c], rs = rstd[c];
    double a1 = 0, a2 = 0;
#pragma unroll 4
    for (int bb = 0; bb < B; ++bb) {
      const float g1 = G[(int64_t)bb * C + c], g2 = G[BC + (int64_t)bb * C + c];
      a1 += g1;
      a2 += (double)rs * ((double)gate[(int64_t)bb * C + c] * g2 - (double)mu * g1);
    }
    if (b == 0) {
      if (dbeta) atomicAdd(dbeta + c, (float)a1);
      if (dgamma) atomicAdd(dgamma + c, (float)a2);
    }
    // eval mode: the BatchNorm is a fixed affine of the running statistics, no batch-mean terms in its backward
    const float c1 = training ? (float)(a1 / n) : 0.f, c2 = training ? (float)(a2 / n) : 0.f;
    const float g = gate[(int64_t)b * C + c];
    const float s1 = S[(int64_t)b * C + c], s2 = S[BC + (int64_t)b * C + c];
    const float g2 = G[BC + (int64_t)b * C + c];
    const float s = gamma[c] * rs;
    // dgate = sum_hw dz * a with dz = s*(g' - c1 - zhat*c2), zhat = rs*(a*g - mu)
    const float dgate = s * (g2 - c1 * s1 - c2 * rs * (g * s2 - mu * s1));
    const float dp = dgate * g * (1.f - g);
    dpre2[c] = dp;
    m[c] = s1 * inv_hw;
    if (db2) atomicAdd(db2 + c, dp);
    PQR[(int64_t)b * C + c] = s * g;                                  // P
    PQR[BC + (int64_t)b * C + c] = -s * rs * g * g * c2;              // Q
    PQR[2 * BC + (int64_t)b * C + c] = -s * g * (c1 - c2 * rs * mu);  // R (dm/HW added below)
  }
  __syncthreads();
  // dw2[c, j] += dpre2[c] * h[j];  dh[j] = sum_c w2[c, j] * dpre2[c]
  for (int idx = tid; idx < C * Cr; idx += nt) {
    const int c = idx / Cr, j = idx % Cr;
    if (dw2) atomicAdd(dw2 + idx, dpre2[c] * hact[j]);
  }
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int j = warp; j < Cr; j += nw) {
    float a = 0.f;
#pragma unroll 4
    for (int c = lane; c < C; c += 32) a = fmaf(__ldg(w2 + (int64_t)c * Cr + j), dpre2[c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      const float d = a * (hidden[(int64_t)b * Cr + j] > 0.f ? 1.f : ACCX_LRELU);
      dpre1[j] = d;
      if (db1) atomicAdd(db1 + j, d);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < C * Cr; idx += nt) {
    const int j = idx / C, c = idx % C;
    if (dw1) atomicAdd(dw1 + idx, dpre1[j] * m[c]);
  }
  for (int c = tid; c < C; c += nt) {
    float dm = 0.f;
#pragma unroll 4
    for (int j = 0; j < Cr; ++j) dm = fmaf(__ldg(w1 + (int64_t)j * C + c), dpre1[j], dm);
    PQR[2 * BC + (int64_t)b * C + c] += dm * inv_hw;
  }
}

// da (+)= P*g' + Q*a + R;  grid.x = B * chunks (one image per block: gate and P, Q, R in registers)
template <typename T, int VEC, int U>
__global__ void se_bwd_apply_kernel(int B, int HW, int C, int chunks, const T* __restrict__ x, const float* scale,
                                    const float* shift, int act, const float* __restrict__ gate,
                                    const float* se_scale, const float* se_shift, const T* __restrict__ dout,
                                    const float* mix, const float* __restrict__ PQR, T* __restrict__ da,
                                    int accumulate, const float* bn_mean, const float* bn_rstd, float* bn_sums,
                                    Det det) {
  pdl_sync();
  // bn_sums != NULL: also the BatchNorm-backward reduction of the lazy input's own BatchNorm on the gradient
  // just produced (sum g, sum g*xhat with g = da*act'(x)) -- saves the separate accx_bn_bwd_reduce pass
  extern __shared__ float smem[];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv * VEC < C;
  if (!active && bn_sums == nullptr) return;
  const int c0 = active ? cv * VEC : 0;
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  float bacc[2][VEC], bmu[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { bacc[0][i] = bacc[1][i] = 0.f; bmu[i] = 0.f; }
  if (bn_sums) ldf<VEC>(bn_mean + c0, bmu);
  Lazy<VEC> lz;
  lz.init(scale, shift, act, c0);
  float gs[VEC], st[VEC], cp[VEC], cq[VEC], cr[VEC];
  const int64_t BC = (int64_t)B * C;
  const float mx = mix ? *mix : 1.f;
  ldf<VEC>(se_scale + c0, gs);
  ldf<VEC>(se_shift + c0, st);
  ldf<VEC>(PQR + (int64_t)b * C + c0, cp);
  ldf<VEC>(PQR + BC + (int64_t)b * C + c0, cq);
  ldf<VEC>(PQR + 2 * BC + (int64_t)b * C + c0, cr);
  {
    float g[VEC];
    ldf<VEC>(gate + (int64_t)b * C + c0, g);
#pragma unroll
    for (int i = 0; i < VEC; ++i) { gs[i] *= g[i]; cp[i] *= mx; }
  }
  const int64_t off = (int64_t)b * HW * C + c0;
  RawVec<T, VEC> rv[U], rd[U], ro[U];
  if (active)
  pixel_loop<U>(chunk * blockDim.y + threadIdx.y, HW, chunks * blockDim.y,
      [&](int u, int64_t p) {
        rv[u].load(x + off + p * C);
        rd[u].load(dout + off + p * C);
        if (accumulate) ro[u].load(da + off + p * C);
      },
      [&](int u, int64_t p) {
        float v[VEC], d[VEC], o[VEC], raw[VEC];
        rv[u].unpack(v);
        rd[u].unpack(d);
#pragma unroll
        for (int i = 0; i < VEC; ++i) raw[i] = v[i];
        lz.apply(v);
        if (accumulate) {
          ro[u].unpack(o);
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) o[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float uu = fmaf(v[i], gs[i], st[i]);
          const float gp = d[i] * (uu > 0.f ? 1.f : ACCX_LRELU);
          o[i] += fmaf(cp[i], gp, fmaf(cq[i], v[i], cr[i]));
        }
        stv<T, VEC>(da + off + p * C, o);
        if (bn_sums) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float gi = o[i] * lz.dact(raw[i], i);
            bacc[0][i] += gi;
            bacc[1][i] = fmaf(gi, raw[i] - bmu[i], bacc[1][i]);
          }
        }
      });
  if (bn_sums) {
    float rs[VEC];
    ldf<VEC>(bn_rstd + c0, rs);
#pragma unroll
    for (int i = 0; i < VEC; ++i) bacc[1][i] *= rs[i];
    reduce_lanes_atomic<2, VEC>(bacc, smem, bn_sums, C, C, det, blockIdx.y, blockIdx.x, gridDim.x);
  }
}

// pixels in flight per thread: a compile-time unroll chosen at launch (tuning knob, else the per-kernel default)
#define ACCX_DISPATCH_U(u, ...)                         \
  do {                                                  \
    if ((u) >= 16) { constexpr int U = 16; __VA_ARGS__ }  \
    else if ((u) >= 8) { constexpr int U = 8; __VA_ARGS__ } \
    else { constexpr int U = 4; __VA_ARGS__ }           \
  } while (0)

// channel-vector width: 1 (scalar fallback), the full 16-byte vector, or -- bf16 only, for the register-heavy
// backward kernels -- half of it (8-byte accesses, half the per-channel constants per thread: 126 instead of
// 203-239 registers, two blocks per SM instead of one; B200, 51 MB tensors: se_bwd_reduce 45 -> 37 us,
// se_bwd_apply 61 -> 46 us; KNOB_SE_BWD_VEC = 8 restores the full vector)
#define ACCX_DISPATCH_VEC_H(lanes, ...)                                  \
  do {                                                                   \
    if ((lanes).vec == 1) {                                              \
      constexpr int VEC = 1;                                             \
      __VA_ARGS__                                                        \
    } else if ((lanes).vec == 4) {                                       \
      constexpr int VEC = 4;                                             \
      __VA_ARGS__                                                        \
    } else if ((lanes).vec == 2) {                                       \
      constexpr int VEC = 2;                                             \
      __VA_ARGS__                                                        \
    } else {                                                             \
      constexpr int VEC = accx::DT<T>::VEC;                              \
      __VA_ARGS__                                                        \
    }                                                                    \
  } while (0)

// deterministic mode: dmix += partial[0] + partial[1] + ..  in block order
__global__ void se_dmix_fold_kernel(const float* __restrict__ part, int n, float* dmix) {
  pdl_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += part[i];
    *dmix += s;
  }
}

static inline int se_chunks(int B, int HW, int ty, int target_blocks = 148 * 4) {
  // enough blocks to fill the machine, but keep the number of atomics per (b,c) small
  int per_img = (HW + ty * 8 - 1) / (ty * 8);
  int want = (target_blocks + B - 1) / B;
  int c = per_img < want ? per_img : want;
  return c < 1 ? 1 : c;
}

}  // namespace accx

using namespace accx;

extern "C" {

int accx_se_squeeze(int dtype, int B, int HW, int C, const void* x, const float* scale, const float* shift, int act,
                    float* S, void* stream) {
  ACCX_REQUIRE(B > 0 && HW > 0 && C > 0 && x && S, "se_squeeze: bad arguments");
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(x));
    int chunks = se_chunks(B, HW, l.ty, 148 * knob(KNOB_SE_SQUEEZE_BLOCKS, 4));
    dim3 block(l.tx, l.ty), grid(B * chunks, l.gy);
    size_t sm = (size_t)l.tx * l.ty * l.vec * sizeof(float);
    const int u = knob(KNOB_SE_SQUEEZE_U, 4);
    Det det;
    if (!det_handle((int64_t)grid.x * grid.y * 2 * l.tx * l.vec, (int64_t)B * grid.y, (cudaStream_t)stream, det)) return ACCX_ERR_INVALID;
    ACCX_DISPATCH_VEC(l, {
      ACCX_DISPATCH_U(u, {
        launch_k(se_squeeze_kernel<T, VEC, U>, grid, block, sm, (cudaStream_t)stream, B, HW, C, chunks, (const T*)x, scale,
                                                                                shift, act, S, det);
      });
    });
  });
  return check_launch("se_squeeze");
}

int accx_se_gate(int B, int C, int Cr, double HW, const float* S, const float* w1, const float* b1, const float* w2,
                 const float* b2, const float* gamma, const float* beta, float eps, float momentum, int training,
                 float* running_mean, float* running_var, int64_t* nbt, float* gate, float* hidden, float* scale,
                 float* shift, float* mean, float* rstd, unsigned int* counter, void* stream) {
  ACCX_REQUIRE(B > 0 && C > 0 && Cr > 0 && S && w1 && b1 && w2 && b2 && gate && hidden && scale && shift && counter,
               "se_gate: bad arguments (C=%d Cr=%d)", C, Cr);
  size_t sm = (size_t)(C + Cr) * sizeof(float);
  launch_k(se_gate_kernel, B, SE_GATE_THREADS, sm, (cudaStream_t)stream, B, C, Cr, HW, S, w1, b1, w2, b2, gamma, beta, eps, momentum,
                                                       training, running_mean, running_var, nbt, gate, hidden, scale,
                                                       shift, mean, rstd, counter);
  return check_launch("se_gate");
}

int accx_se_apply(int dtype, int B, int HW, int C, const void* x, const float* scale, const float* shift, int act,
                  const float* gate, const float* se_scale, const float* se_shift, const void* residual,
                  const float* mix, void* out, float* stats, void* stream) {
  ACCX_REQUIRE(B > 0 && HW > 0 && C > 0 && x && gate && se_scale && se_shift && out, "se_apply: bad arguments");
  ACCX_DISPATCH_T(dtype, {
    Lanes l = make_lanes(C, DT<T>::VEC, aligned16(x) && aligned16(out) && (!residual || aligned16(residual)));
    const int chunks = se_chunks(B, HW, l.ty, 148 * (stats ? knob(KNOB_SE_APPLY_STATS_BLOCKS, 4) : knob(KNOB_SE_APPLY_BLOCKS, 4)));
    dim3 block(l.tx, l.ty), grid(B * chunks, l.gy);
    size_t sm = (size_t)l.tx * l.ty * l.vec * sizeof(float);
    const int u = knob(KNOB_SE_APPLY_U, 4);
    Det det;
    if (!det_handle(stats ? (int64_t)grid.x * grid.y * 2 * l.tx * l.vec : 0, grid.y, (cudaStream_t)stream, det)) return ACCX_ERR_INVALID;
    ACCX_DISPATCH_VEC(l, {
      ACCX_DISPATCH_U(u, {
        launch_k(se_apply_kernel<T, VEC, U>, grid, block, sm, (cudaStream_t)stream, B, HW, C, chunks, (const T*)x, scale, shift, act,
                                                                              gate, se_scale, se_shift, (const T*)residual,
                                                                              mix, (T*)out, stats, det);
      });
    });
  });
  return check_launch("se_apply");
}

int accx_se_bwd_reduce(int dtype, int B, int HW, int C, const void* x, const float* scale, const float* shift, int act,
                       const float* gate, const float* se_scale, const float* se_shift, const void* dout,
                       const float* mix, const void* residual, float* dmix, float* G, void* stream) {
  ACCX_REQUIRE(B > 0 && HW > 0 && C > 0 && x && gate && dout && G, "se_bwd_reduce: bad arguments");
  ACCX_REQUIRE(!dmix || (mix && residual), "se_bwd_reduce: dmix needs mix and residual");
  ACCX_DISPATCH_T(dtype, {
    const int fv = knob(KNOB_SE_BWD_VEC, 4) == 4 ? 4 : (knob(KNOB_SE_BWD_VEC, 4) == 2 && sizeof(T) == 2 ? 2 : DT<T>::VEC);
    Lanes l = make_lanes(C, fv, aligned16(x) && aligned16(dout) && (!residual || aligned16(residual)));
    int chunks = se_chunks(B, HW, l.ty, 148 * knob(KNOB_SE_BWD_REDUCE_BLOCKS, 4));
    dim3 block(l.tx, l.ty), grid(B * chunks, l.gy);
    size_t sm = (size_t)l.tx * l.ty * l.vec * sizeof(float);
    const int u = knob(KNOB_SE_BWD_REDUCE_U, 8);
    Det det;
    const int64_t lanes_floats = (int64_t)grid.x * grid.y * 2 * l.tx * l.vec;
    const int n_blocks = (int)(grid.x * grid.y);
    if (!det_handle(lanes_floats + (dmix ? n_blocks : 0), (int64_t)B * grid.y, (cudaStream_t)stream, det)) return ACCX_ERR_INVALID;
    float* dmix_part = (det.ws && dmix) ? det.ws + lanes_floats : nullptr;
    ACCX_DISPATCH_VEC_H(l, {
      ACCX_DISPATCH_U(u, {
        launch_k(se_bwd_reduce_kernel<T, VEC, U>, grid, block, sm, (cudaStream_t)stream, 
            B, HW, C, chunks, (const T*)x, scale, shift, act, gate, se_scale, se_shift, (const T*)dout, mix,
            (const T*)residual, dmix, G, det, dmix_part);
      });
    });
    if (dmix_part) launch_k(se_dmix_fold_kernel, 1, 32, 0, (cudaStream_t)stream, (const float*)dmix_part, n_blocks, dmix);
  });
  return check_launch("se_bwd_reduce");
}

int accx_se_bwd_gate(int B, int C, int Cr, double HW, const float* S, const float* G, const float* gate,
                     const float* hidden, const float* w1, const float* w2, const float* gamma, const float* mean,
                     const float* rstd, float* dw1, float* db1, float* dw2, float* db2, float* dgamma, float* dbeta,
                     float* PQR, int training, void* stream) {
  ACCX_REQUIRE(B > 0 && C > 0 && Cr > 0 && S && G && gate && hidden && w1 && w2 && gamma && mean && rstd && PQR,
               "se_bwd_gate: bad arguments");
  size_t sm = (size_t)(2 * C + 2 * Cr) * sizeof(float);
  if (det_on()) {
    // the per-image blocks add into the shared FC gradients: one single-block launch per image, in image order
    for (int b = 0; b < B; ++b)
      launch_k(se_bwd_gate_kernel, 1, SE_GATE_THREADS, sm, (cudaStream_t)stream, B, C, Cr, HW, S, G, gate, hidden, w1, w2, gamma,
               mean, rstd, dw1, db1, dw2, db2, dgamma, dbeta, PQR, training, b);
    return check_launch("se_bwd_gate");
  }
  launch_k(se_bwd_gate_kernel, B, SE_GATE_THREADS, sm, (cudaStream_t)stream, B, C, Cr, HW, S, G, gate, hidden, w1, w2, gamma, mean, rstd,
                                                           dw1, db1, dw2, db2, dgamma, dbeta, PQR, training, 0);
  return check_launch("se_bwd_gate");
}

int accx_se_bwd_apply(int dtype, int B, int HW, int C, const void* x, const float* scale, const float* shift, int act,
                      const float* gate, const float* se_scale, const float* se_shift, const void* dout,
                      const float* mix, const float* PQR, void* da, int accumulate, const float* bn_mean,
                      const float* bn_rstd, float* bn_sums, void* stream) {
  ACCX_REQUIRE(B > 0 && HW > 0 && C > 0 && x && gate && dout && PQR && da, "se_bwd_apply: bad arguments");
  ACCX_REQUIRE(!bn_sums || (bn_mean && bn_rstd), "se_bwd_apply: bn_sums needs bn_mean and bn_rstd");
  ACCX_DISPATCH_T(dtype, {
    // small maps (levels 3-5: <= 56 x 56 per image) in bf16: two channels per thread, four pixels in flight (72 registers,
    // three blocks per SM) -- 24.9 -> 20.7 us at 16 x 3136 x 128, 22.9 -> 15.8 us at 16 x 784 x 256 with two blocks per SM;
    // the large maps keep four channels (47 vs 57 us at 16 x 50176 x 32): profiles/r02_se_bwd_channels_per_thread.txt
    const bool small_map = sizeof(T) == 2 && HW <= 3136 && g_knobs[KNOB_SE_BWD_VEC] == 0;
    const int fv = small_map ? 2 : (knob(KNOB_SE_BWD_VEC, 4) == 4 ? 4 : (knob(KNOB_SE_BWD_VEC, 4) == 2 && sizeof(T) == 2 ? 2 : DT<T>::VEC));
    Lanes l = make_lanes(C, fv, aligned16(x) && aligned16(dout) && aligned16(da));
    // reducing variant: few blocks (atomics)
    const int chunks = se_chunks(B, HW, l.ty, 148 * (bn_sums ? knob(KNOB_SE_BWD_APPLY_BN_BLOCKS, (small_map && HW <= 784) ? 2 : 4)
                                                             : knob(KNOB_SE_BWD_APPLY_BLOCKS, 8)));
    dim3 block(l.tx, l.ty), grid(B * chunks, l.gy);
    const size_t sm = bn_sums ? (size_t)l.tx * l.ty * l.vec * sizeof(float) : 0;
    const int u = knob(KNOB_SE_BWD_APPLY_U, small_map ? 4 : 8);
    Det det;
    if (!det_handle(bn_sums ? (int64_t)grid.x * grid.y * 2 * l.tx * l.vec : 0, grid.y, (cudaStream_t)stream, det)) return ACCX_ERR_INVALID;
    ACCX_DISPATCH_VEC_H(l, {
      ACCX_DISPATCH_U(u, {
        launch_k(se_bwd_apply_kernel<T, VEC, U>, grid, block, sm, (cudaStream_t)stream, 
            B, HW, C, chunks, (const T*)x, scale, shift, act, gate, se_scale, se_shift, (const T*)dout, mix, PQR, (T*)da,
            accumulate, bn_mean, bn_rstd, bn_sums, det);
      });
    });
  });
  return check_launch("se_bwd_apply");
}

}  // extern "C"
